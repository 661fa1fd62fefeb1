// render_params.h — the argument block of the render kernels (passed by value, __grid_constant__).
#pragma once

#include <cstdint>

#include <cuda_runtime.h>

namespace rtb {

enum OutMode : int {
    kOutFrame = 0,   // row-major RGB8 frame [ny][nx][3] (possibly peer memory)
    kOutPacked = 1,  // this part's tiles back to back, [local_tile][RT_TILE][RT_TILE][3]
};

struct RenderParams {
    // scene (device pointers; layouts in rt_internal.h)
    const float4 *nodes;
    const float4 *prims;
    const float4 *tri_nm;
    const float4 *tri_nn;  // normalize(normal) per triangle (raytracer.cpp:414, 432 re-normalise the stored normal)
    const float4 *sph_cr;
    const int *sph_mat;
    const uint32_t *ranks;
    const float4 *ref_nodes;        // reference tree (exact culling replay)
    const int *ref_leaf_prims;
    const float4 *prim_bounds;      // [2 * n_prims] un-padded bounds of each primitive as the reference computes them
    const int *slot_of_prim;        // prim id -> slot in `prims`
    int exact_culling;              // 1: reproduce the reference's box-culling decisions (default)
    const float4 *materials;
    const float4 *lights;
    int n_nodes, n_tris, n_prims, n_lights;
    int max_depth;       // Scene::max_recursion_depth
    int brute_force;
    float eps;           // Scene::shadow_ray_epsilon
    float ambient[3];    // Scene::ambient_light
    float background[3]; // (float) Scene::background_color
    // camera, precomputed on the host exactly as EyeRayGenerator::init does (raytracer.cpp:292-314)
    float e[3], q[3], u[3], v[3];
    float su_mul, sv_mul;
    // frame
    int nx, ny;          // output resolution
    int f;               // supersampling factor (sub-sample grid is nx*f by ny*f)
    int P;               // output pixels per work-item side
    int Ph;              // work-item height in pixels (= P except on small frames: one 8x4 round per item)
    int items_x;         // work items per tile row = ceil(RT_TILE / P)
    int items_y;         // work items per tile column = ceil(RT_TILE / Ph)
    int tiles_x, tiles_y;
    int part_rank, part_world;
    unsigned int n_items;  // work items of this part
    int out_mode;
    int refill_threshold;  // kernel 2: refill idle lanes once <= this many lanes are busy
    unsigned char *out;
    unsigned int *work_counter;      // zeroed before launch
    unsigned long long *stats;       // [6] primary, reflection, shadow, occluded, replayed closest, replayed any
};

}  // namespace rtb
