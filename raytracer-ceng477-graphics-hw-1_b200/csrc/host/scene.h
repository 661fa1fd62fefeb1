// scene.h — host-side scene surface, field-compatible with the reference's parser.h.
//
// The reference's driver (raytracer.cpp:487-525) talks to its renderer through plain structs in
// namespace `parser` (parser.h:170-266).  A maintainer switching to the B200 path keeps that code:
// the same struct and field names exist here, with only what the host needs (no per-pixel math —
// that lives in the CUDA kernels).  `flatten()` produces the flat RtSceneDesc the C-ABI takes.
#pragma once

#include <string>
#include <vector>

#include "rt_b200.h"

namespace parser {

typedef unsigned char Pixel[3];  // parser.h:12
typedef Pixel *Image;            // parser.h:13

struct Vec3f { float x, y, z; };      // parser.h:18
struct Vec3i { int x, y, z; };        // parser.h:107
struct Vec4f { float x, y, z, w; };   // parser.h:166

struct Camera {  // parser.h:170-178
    Vec3f position;
    Vec3f gaze;
    Vec3f up;
    Vec4f near_plane;  // l r b t
    float near_distance;
    int image_width, image_height;
    std::string image_name;
};

struct PointLight {  // parser.h:180-183
    Vec3f position;
    Vec3f intensity;
};

struct Material {  // parser.h:185-192
    bool is_mirror;
    Vec3f ambient;
    Vec3f diffuse;
    Vec3f specular;
    Vec3f mirror;
    float phong_exponent;
};

struct Face {  // parser.h:194-198
    int v0_id;
    int v1_id;
    int v2_id;
};

struct Sphere {  // parser.h:200-204
    int material_id;
    int center_vertex_id;
    float radius;
};

struct Mesh {  // parser.h:238-241
    int material_id;
    std::vector<Face> faces;
};

struct Triangle {  // parser.h:243-251 (normal/center are device-side business here)
    int material_id;
    Face indices;
};

struct Scene {  // parser.h:254-269
    Vec3i background_color;
    float shadow_ray_epsilon;
    int max_recursion_depth;
    std::vector<Camera> cameras;
    Vec3f ambient_light;
    std::vector<PointLight> point_lights;
    std::vector<Material> materials;
    std::vector<Vec3f> vertex_data;
    std::vector<Mesh> meshes;
    std::vector<Triangle> triangles;
    std::vector<Sphere> spheres;

    // Same contract as parser.cpp:6 — throws std::runtime_error when the file cannot be read or has
    // no root element.  Unlike the reference it also throws (instead of dereferencing null) when a
    // mandatory element is missing.
    void loadFromXml(const std::string &filepath);
};

// Flat, C-ABI view of a Scene.  Owns the arrays the RtSceneDesc points into.
struct FlatScene {
    std::vector<RtVec3> vertices;
    std::vector<RtTriangle> triangles;  // <Triangle>s first, then every mesh face in file order (raytracer.cpp:336-341)
    std::vector<RtSphere> spheres;
    std::vector<RtMaterial> materials;
    std::vector<RtPointLight> lights;
    RtSceneDesc desc;
};

void flatten(const Scene &scene, FlatScene &out);
RtCamera to_rt_camera(const Camera &camera);

}  // namespace parser

// ASCII P3 writer, byte-identical to the reference's ppm.cpp:4-39 ("P3\n%d %d\n255\n", "%d " per
// channel, no trailing space after the last value of a row, "\n" per row).  Throws
// std::runtime_error when the file cannot be opened, like the reference.
void write_ppm(const char *filename, unsigned char *data, int width, int height);
