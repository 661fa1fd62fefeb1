// host_api.cpp — C entry points over the host-side scene surface (XML reader, flattening, PPM
// writer) so that Python (tests, bench.py) drives exactly the code the `raytracer` CLI uses.
// Host only: no CUDA here.
#include <cstring>
#include <stdexcept>
#include <string>

#include "scene.h"

namespace {
struct HostScene {
    parser::Scene scene;
    parser::FlatScene flat;
};
thread_local std::string g_host_err;
}  // namespace

extern "C" {

const char *rth_last_error() { return g_host_err.c_str(); }

// parser::Scene::loadFromXml + flatten; NULL on failure
void *rth_scene_load_xml(const char *path) {
    try {
        auto *h = new HostScene();
        h->scene.loadFromXml(path);
        parser::flatten(h->scene, h->flat);
        return h;
    } catch (std::exception &e) {
        g_host_err = e.what();
        return nullptr;
    }
}

void rth_scene_free(void *h) { delete (HostScene *) h; }

const RtSceneDesc *rth_scene_desc(void *h) { return &((HostScene *) h)->flat.desc; }

int rth_scene_num_cameras(void *h) { return (int) ((HostScene *) h)->scene.cameras.size(); }

int rth_scene_num_meshes(void *h) { return (int) ((HostScene *) h)->scene.meshes.size(); }

int rth_scene_camera(void *h, int i, RtCamera *out, char *name, int name_cap) {
    auto &cams = ((HostScene *) h)->scene.cameras;
    if (i < 0 || i >= (int) cams.size()) return -1;
    *out = parser::to_rt_camera(cams[i]);
    if (name && name_cap > 0) snprintf(name, (size_t) name_cap, "%s", cams[i].image_name.c_str());
    return 0;
}

int rth_write_ppm(const char *path, unsigned char *rgb, int w, int h) {
    try {
        write_ppm(path, rgb, w, h);
        return 0;
    } catch (std::exception &e) {
        g_host_err = e.what();
        return -1;
    }
}

}  // extern "C"
