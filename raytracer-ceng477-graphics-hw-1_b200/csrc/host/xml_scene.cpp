// xml_scene.cpp — the scene loader behind parser::Scene::loadFromXml, with the reference's
// semantics (parser.cpp:6-218) on top of a ~150-line XML reader of our own (the reference vendors
// tinyxml2 4.0.1; the scene grammar needs only elements, one attribute and text).
//
// Semantics kept:
//   * root = first element of the document (parser.cpp:17)
//   * defaults: BackgroundColor "0 0 0", ShadowRayEpsilon 0.001, MaxRecursionDepth 0 (parser.cpp:24-57)
//   * BackgroundColor parsed as ints (parser.cpp:33)
//   * children are looked up BY NAME, first match (order inside <Camera>/<Material> is free)
//   * <Material type="mirror"> sets is_mirror (parser.cpp:119); every id= attribute is ignored, ids are
//     positional and 1-based
//   * VertexData / Faces are whitespace-separated numbers until the text ends (parser.cpp:146-151, 166-171)
//   * objects are collected in three passes: Mesh, Triangle, Sphere (parser.cpp:154-217), all from the
//     first <Objects> element
//   * numbers are read the way operator>> reads them (leading whitespace skipped, "1e-3" accepted)
#include <cctype>
#include <cerrno>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <stdexcept>

#include "scene.h"

namespace {

struct XmlNode {
    std::string name;
    std::vector<std::pair<std::string, std::string>> attrs;
    std::string text;  // text that precedes the first child element (what tinyxml2's GetText() returns)
    bool text_first = false;
    std::vector<std::unique_ptr<XmlNode>> children;

    const XmlNode *child(const char *n) const {
        for (auto &c: children)
            if (c->name == n) return c.get();
        return nullptr;
    }
    const char *attr(const char *n) const {
        for (auto &a: attrs)
            if (a.first == n) return a.second.c_str();
        return nullptr;
    }
};

class XmlReader {
public:
    explicit XmlReader(const std::string &s) : s_(s), p_(0) {}

    std::unique_ptr<XmlNode> parse_document() {
        skip_misc();
        if (p_ >= s_.size() || s_[p_] != '<') return nullptr;
        return parse_element();
    }

private:
    const std::string &s_;
    size_t p_;

    [[noreturn]] void fail(const char *what) { throw std::runtime_error(std::string("Error: The xml file cannot be loaded. (") + what + ")"); }
    bool starts(const char *lit) const { return s_.compare(p_, strlen(lit), lit) == 0; }
    void skip_ws() {
        while (p_ < s_.size() && isspace((unsigned char) s_[p_])) p_++;
    }
    void skip_until(const char *lit) {
        size_t e = s_.find(lit, p_);
        if (e == std::string::npos) fail("unterminated markup");
        p_ = e + strlen(lit);
    }
    // whitespace, comments, <?...?> and <!DOCTYPE ...> between elements
    void skip_misc() {
        for (;;) {
            skip_ws();
            if (starts("<!--")) skip_until("-->");
            else if (starts("<?")) skip_until("?>");
            else if (starts("<!") && !starts("<![CDATA[")) skip_until(">");
            else return;
        }
    }
    std::string parse_name() {
        size_t b = p_;
        while (p_ < s_.size() && !isspace((unsigned char) s_[p_]) && s_[p_] != '>' && s_[p_] != '/' && s_[p_] != '=') p_++;
        if (p_ == b) fail("empty name");
        return s_.substr(b, p_ - b);
    }
    static void append_decoded(std::string &out, const std::string &raw) {
        for (size_t i = 0; i < raw.size(); i++) {
            if (raw[i] == '&') {
                static const struct { const char *ent; char ch; } E[] = {{"&lt;", '<'}, {"&gt;", '>'}, {"&amp;", '&'}, {"&quot;", '"'}, {"&apos;", '\''}};
                bool done = false;
                for (auto &e: E)
                    if (raw.compare(i, strlen(e.ent), e.ent) == 0) {
                        out.push_back(e.ch);
                        i += strlen(e.ent) - 1;
                        done = true;
                        break;
                    }
                if (done) continue;
            }
            if (raw[i] == '\r') {  // line-ending normalisation, as tinyxml2 does
                out.push_back('\n');
                if (i + 1 < raw.size() && raw[i + 1] == '\n') i++;
                continue;
            }
            out.push_back(raw[i]);
        }
    }
    std::unique_ptr<XmlNode> parse_element() {
        std::unique_ptr<XmlNode> n(new XmlNode());
        p_++;  // '<'
        n->name = parse_name();
        for (;;) {
            skip_ws();
            if (p_ >= s_.size()) fail("unterminated tag");
            if (s_[p_] == '/') {
                if (!starts("/>")) fail("bad tag");
                p_ += 2;
                return n;
            }
            if (s_[p_] == '>') {
                p_++;
                break;
            }
            std::string an = parse_name();
            skip_ws();
            if (p_ >= s_.size() || s_[p_] != '=') fail("attribute without value");
            p_++;
            skip_ws();
            char q = p_ < s_.size() ? s_[p_] : 0;
            if (q != '"' && q != '\'') fail("unquoted attribute");
            size_t e = s_.find(q, p_ + 1);
            if (e == std::string::npos) fail("unterminated attribute");
            std::string v;
            append_decoded(v, s_.substr(p_ + 1, e - p_ - 1));
            n->attrs.emplace_back(an, v);
            p_ = e + 1;
        }
        bool first = true;
        for (;;) {  // content
            size_t lt = s_.find('<', p_);
            if (lt == std::string::npos) fail("unterminated element");
            if (lt > p_) {
                if (first) {
                    append_decoded(n->text, s_.substr(p_, lt - p_));
                    n->text_first = true;
                }
                p_ = lt;
            }
            if (starts("</")) {
                p_ += 2;
                std::string close = parse_name();
                if (close != n->name) fail("mismatched end tag");
                skip_ws();
                if (p_ >= s_.size() || s_[p_] != '>') fail("bad end tag");
                p_++;
                return n;
            }
            if (starts("<!--")) { skip_until("-->"); continue; }
            if (starts("<?")) { skip_until("?>"); continue; }
            if (starts("<![CDATA[")) {
                size_t e = s_.find("]]>", p_);
                if (e == std::string::npos) fail("unterminated CDATA");
                if (first) n->text += s_.substr(p_ + 9, e - p_ - 9);
                p_ = e + 3;
                continue;
            }
            first = false;  // text after a child element is not GetText() material
            n->children.push_back(parse_element());
        }
    }
};

// operator>>-style number scanning over one element's text
struct Scanner {
    const char *p;
    const char *what;
    explicit Scanner(const XmlNode *n, const char *what) : p(n->text.c_str()), what(what) {}
    void skip() {
        while (*p && isspace((unsigned char) *p)) p++;
    }
    bool at_end() {
        skip();
        return *p == 0;
    }
    [[noreturn]] void fail() { throw std::runtime_error(std::string("Error: malformed number in <") + what + ">"); }
    float f() {
        skip();
        char *e;
        float v = strtof(p, &e);
        if (e == p) fail();
        p = e;
        return v;
    }
    int i() {
        skip();
        char *e;
        long v = strtol(p, &e, 10);
        if (e == p) fail();
        p = e;
        return (int) v;
    }
    std::string word() {
        skip();
        const char *b = p;
        while (*p && !isspace((unsigned char) *p)) p++;
        return std::string(b, p);
    }
    parser::Vec3f v3() {
        parser::Vec3f v;
        v.x = f(); v.y = f(); v.z = f();
        return v;
    }
};

const XmlNode *need(const XmlNode *parent, const char *name) {
    const XmlNode *c = parent ? parent->child(name) : nullptr;
    if (!c) throw std::runtime_error(std::string("Error: element <") + name + "> is missing.");
    return c;
}

}  // namespace

void parser::Scene::loadFromXml(const std::string &filepath) {
    std::string data;
    {
        FILE *f = fopen(filepath.c_str(), "rb");
        if (!f) throw std::runtime_error("Error: The xml file cannot be loaded.");
        char buf[1 << 16];
        size_t n;
        while ((n = fread(buf, 1, sizeof buf, f)) > 0) data.append(buf, n);
        fclose(f);
    }
    XmlReader reader(data);
    std::unique_ptr<XmlNode> root = reader.parse_document();
    if (!root) throw std::runtime_error("Error: Root is not found.");

    if (auto e = root->child("BackgroundColor")) {
        Scanner s(e, "BackgroundColor");
        background_color.x = s.i(); background_color.y = s.i(); background_color.z = s.i();
    } else {
        background_color = {0, 0, 0};
    }
    if (auto e = root->child("ShadowRayEpsilon")) shadow_ray_epsilon = Scanner(e, "ShadowRayEpsilon").f();
    else shadow_ray_epsilon = strtof("0.001", nullptr);
    if (auto e = root->child("MaxRecursionDepth")) max_recursion_depth = Scanner(e, "MaxRecursionDepth").i();
    else max_recursion_depth = 0;

    for (auto &c: need(root.get(), "Cameras")->children) {
        if (c->name != "Camera") continue;
        Camera cam;
        cam.position = Scanner(need(c.get(), "Position"), "Position").v3();
        cam.gaze = Scanner(need(c.get(), "Gaze"), "Gaze").v3();
        cam.up = Scanner(need(c.get(), "Up"), "Up").v3();
        {
            Scanner s(need(c.get(), "NearPlane"), "NearPlane");
            cam.near_plane.x = s.f(); cam.near_plane.y = s.f(); cam.near_plane.z = s.f(); cam.near_plane.w = s.f();
        }
        cam.near_distance = Scanner(need(c.get(), "NearDistance"), "NearDistance").f();
        {
            Scanner s(need(c.get(), "ImageResolution"), "ImageResolution");
            cam.image_width = s.i(); cam.image_height = s.i();
        }
        cam.image_name = Scanner(need(c.get(), "ImageName"), "ImageName").word();
        cameras.push_back(cam);
    }

    const XmlNode *lights = need(root.get(), "Lights");
    ambient_light = Scanner(need(lights, "AmbientLight"), "AmbientLight").v3();
    for (auto &c: lights->children) {
        if (c->name != "PointLight") continue;
        PointLight pl;
        pl.position = Scanner(need(c.get(), "Position"), "Position").v3();
        pl.intensity = Scanner(need(c.get(), "Intensity"), "Intensity").v3();
        point_lights.push_back(pl);
    }

    for (auto &c: need(root.get(), "Materials")->children) {
        if (c->name != "Material") continue;
        Material m;
        const char *type = c->attr("type");
        m.is_mirror = type && strcmp(type, "mirror") == 0;
        m.ambient = Scanner(need(c.get(), "AmbientReflectance"), "AmbientReflectance").v3();
        m.diffuse = Scanner(need(c.get(), "DiffuseReflectance"), "DiffuseReflectance").v3();
        m.specular = Scanner(need(c.get(), "SpecularReflectance"), "SpecularReflectance").v3();
        m.mirror = Scanner(need(c.get(), "MirrorReflectance"), "MirrorReflectance").v3();
        m.phong_exponent = Scanner(need(c.get(), "PhongExponent"), "PhongExponent").f();
        materials.push_back(m);
    }

    {
        Scanner s(need(root.get(), "VertexData"), "VertexData");
        while (!s.at_end()) vertex_data.push_back(s.v3());
    }

    const XmlNode *objects = need(root.get(), "Objects");
    for (auto &c: objects->children) {
        if (c->name != "Mesh") continue;
        Mesh mesh;
        mesh.material_id = Scanner(need(c.get(), "Material"), "Material").i();
        Scanner s(need(c.get(), "Faces"), "Faces");
        while (!s.at_end()) {
            Face f;
            f.v0_id = s.i(); f.v1_id = s.i(); f.v2_id = s.i();
            mesh.faces.push_back(f);
        }
        meshes.push_back(mesh);
    }
    for (auto &c: objects->children) {
        if (c->name != "Triangle") continue;
        Triangle t;
        t.material_id = Scanner(need(c.get(), "Material"), "Material").i();
        Scanner s(need(c.get(), "Indices"), "Indices");
        t.indices.v0_id = s.i(); t.indices.v1_id = s.i(); t.indices.v2_id = s.i();
        triangles.push_back(t);
    }
    for (auto &c: objects->children) {
        if (c->name != "Sphere") continue;
        Sphere sp;
        sp.material_id = Scanner(need(c.get(), "Material"), "Material").i();
        sp.center_vertex_id = Scanner(need(c.get(), "Center"), "Center").i();
        sp.radius = Scanner(need(c.get(), "Radius"), "Radius").f();
        spheres.push_back(sp);
    }
}

void parser::flatten(const Scene &scene, FlatScene &out) {
    out.vertices.clear(); out.triangles.clear(); out.spheres.clear(); out.materials.clear(); out.lights.clear();
    for (auto &v: scene.vertex_data) out.vertices.push_back(RtVec3{v.x, v.y, v.z});
    for (auto &t: scene.triangles) out.triangles.push_back(RtTriangle{t.indices.v0_id, t.indices.v1_id, t.indices.v2_id, t.material_id});
    for (auto &m: scene.meshes)
        for (auto &f: m.faces) out.triangles.push_back(RtTriangle{f.v0_id, f.v1_id, f.v2_id, m.material_id});
    for (auto &s: scene.spheres) out.spheres.push_back(RtSphere{s.material_id, s.center_vertex_id, s.radius});
    for (auto &m: scene.materials) {
        RtMaterial r;
        r.ambient = RtVec3{m.ambient.x, m.ambient.y, m.ambient.z};
        r.diffuse = RtVec3{m.diffuse.x, m.diffuse.y, m.diffuse.z};
        r.specular = RtVec3{m.specular.x, m.specular.y, m.specular.z};
        r.mirror = RtVec3{m.mirror.x, m.mirror.y, m.mirror.z};
        r.phong_exponent = m.phong_exponent;
        r.is_mirror = m.is_mirror ? 1 : 0;
        out.materials.push_back(r);
    }
    for (auto &l: scene.point_lights)
        out.lights.push_back(RtPointLight{RtVec3{l.position.x, l.position.y, l.position.z}, RtVec3{l.intensity.x, l.intensity.y, l.intensity.z}});
    RtSceneDesc &d = out.desc;
    d.vertices = out.vertices.data(); d.n_vertices = (int32_t) out.vertices.size();
    d.triangles = out.triangles.data(); d.n_triangles = (int32_t) out.triangles.size();
    d.spheres = out.spheres.data(); d.n_spheres = (int32_t) out.spheres.size();
    d.materials = out.materials.data(); d.n_materials = (int32_t) out.materials.size();
    d.lights = out.lights.data(); d.n_lights = (int32_t) out.lights.size();
    d.ambient_light = RtVec3{scene.ambient_light.x, scene.ambient_light.y, scene.ambient_light.z};
    d.background[0] = scene.background_color.x; d.background[1] = scene.background_color.y; d.background[2] = scene.background_color.z;
    d.shadow_ray_epsilon = scene.shadow_ray_epsilon;
    d.max_recursion_depth = scene.max_recursion_depth;
}

RtCamera parser::to_rt_camera(const Camera &c) {
    RtCamera r;
    r.position = RtVec3{c.position.x, c.position.y, c.position.z};
    r.gaze = RtVec3{c.gaze.x, c.gaze.y, c.gaze.z};
    r.up = RtVec3{c.up.x, c.up.y, c.up.z};
    r.l = c.near_plane.x; r.r = c.near_plane.y; r.b = c.near_plane.z; r.t = c.near_plane.w;
    r.near_distance = c.near_distance;
    r.image_width = c.image_width; r.image_height = c.image_height;
    return r;
}
