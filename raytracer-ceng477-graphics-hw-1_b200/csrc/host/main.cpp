// main.cpp — the reference's command line on the B200 path:  raytracer scene.xml
//
// Same surface as the reference's main (raytracer.cpp:487-525): reads the XML scene, "plants the
// trees" (here: uploads the scene and builds the BVH on the GPU), renders every <Camera> with the
// supersampling factor, writes one ASCII-P3 PPM per camera under its <ImageName> into the current
// directory and prints the same three timing lines.  The reference fixes its AA factor at compile
// time (2, raytracer.cpp:26-28); here 2 is the default and `--aa N` selects it at run time.
//
//   raytracer scene.xml [--aa N] [--res WxH] [--gpus N] [--builder ploc|sah_gpu|lbvh|sah] [--stats]
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "rt_b200.h"
#include "scene.h"

static double seconds_since(std::chrono::high_resolution_clock::time_point t0) {
    return std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - t0).count();
}

int main(int argc, char *argv[]) {
    const char *xml = nullptr;
    int aa = 2, gpus = 1, res_w = 0, res_h = 0, builder = RT_BUILD_DEFAULT;
    bool want_stats = false;
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        if (a == "--aa" && i + 1 < argc) aa = atoi(argv[++i]);
        else if (a == "--gpus" && i + 1 < argc) gpus = atoi(argv[++i]);
        else if (a == "--res" && i + 1 < argc) {
            if (sscanf(argv[++i], "%dx%d", &res_w, &res_h) != 2) {
                fprintf(stderr, "--res expects WxH\n");
                return 2;
            }
        } else if (a == "--builder" && i + 1 < argc) {
            std::string b = argv[++i];
            builder = b == "lbvh" ? RT_BUILD_LBVH_GPU : b == "sah" ? RT_BUILD_SAH_HOST : b == "ploc" ? RT_BUILD_PLOC_GPU : b == "sah_gpu" ? RT_BUILD_SAH_GPU : RT_BUILD_DEFAULT;
        } else if (a == "--stats") want_stats = true;
        else if (a == "--help" || a == "-h") {
            xml = nullptr;
            break;
        }
        else if (!xml) xml = argv[i];
    }
    if (!xml) {
        fprintf(stderr, "usage: raytracer scene.xml [--aa N] [--res WxH] [--gpus N] [--builder ploc|sah_gpu|lbvh|sah] [--stats]\n");
        return 2;
    }
    try {
        parser::Scene scene;
        scene.loadFromXml(xml);

        auto begin1 = std::chrono::high_resolution_clock::now();
        parser::FlatScene flat;
        parser::flatten(scene, flat);
        RtBuildOptions opts;
        memset(&opts, 0, sizeof opts);
        opts.builder = builder;
        int have = rt_device_count();
        if (have < 1) throw std::runtime_error("Error: no CUDA device (this build has no CPU fallback).");
        if (gpus > have) gpus = have;
        std::vector<RtScene *> handles;
        for (int g = 0; g < gpus; g++) {
            if (rt_set_device(g) != RT_OK) throw std::runtime_error(rt_last_error());
            RtScene *h = nullptr;
            if (rt_scene_create(&flat.desc, &opts, &h) != RT_OK) throw std::runtime_error(rt_last_error());
            handles.push_back(h);
        }
        double elapsed1 = seconds_since(begin1);
        printf("Planted trees in %.3f seconds.\n", elapsed1);
        if (aa > 1) printf("Super Sampling Anti aliasing is enabled. (%d*%dx)\n", aa, aa);

        auto begin2 = std::chrono::high_resolution_clock::now();
        // Cameras are rendered back to back on the resident scene; the P3 file of camera i is written by a host
        // thread while the GPUs render camera i+1 (the reference spends half of its "Rendered in"
        // time in fprintf).
        struct Joiner {  // joins on every exit path, including exceptions from the render loop
            std::vector<std::thread> threads;
            ~Joiner() {
                for (auto &t: threads)
                    if (t.joinable()) t.join();
            }
        } joiner;
        std::vector<std::thread> &writers = joiner.threads;
        std::vector<std::string> write_errors(scene.cameras.size());
        size_t cam_index = 0;
        for (auto camera: scene.cameras) {
            if (res_w > 0) camera.image_width = res_w, camera.image_height = res_h;
            RtCamera cam = parser::to_rt_camera(camera);
            auto *image_ptr = new std::vector<unsigned char>((size_t) camera.image_width * camera.image_height * 3);
            std::vector<unsigned char> &image = *image_ptr;
            printf("Rendering %s with %d B200 GPU%s...\n", camera.image_name.c_str(), gpus, gpus > 1 ? "s" : "");
            fflush(stdout);
            RtStats st;
            int rc = rt_render_multi(handles.data(), gpus, &cam, aa, image.data(), &st);
            if (rc != RT_OK) throw std::runtime_error(rt_last_error());
            if (want_stats) {
                unsigned long long rays = st.primary_rays + st.reflection_rays + st.shadow_rays;
                printf("  rays: %llu primary, %llu reflection, %llu shadow (%llu occluded); %.3f ms render, %.3f ms to host, %.1f Mrays/s\n",
                       (unsigned long long) st.primary_rays, (unsigned long long) st.reflection_rays,
                       (unsigned long long) st.shadow_rays, (unsigned long long) st.shadow_occluded, st.ms_render, st.ms_d2h,
                       st.ms_render > 0 ? rays / (st.ms_render * 1e3) : 0.0);
            }
            std::string *err = &write_errors[cam_index++];
            writers.emplace_back([image_ptr, camera, err]() {
                try {
                    write_ppm(camera.image_name.c_str(), image_ptr->data(), camera.image_width, camera.image_height);
                } catch (std::exception &e) {
                    *err = e.what();
                }
                delete image_ptr;
            });
        }
        for (auto &w: writers) w.join();
        for (auto &e: write_errors)
            if (!e.empty()) throw std::runtime_error(e);
        double elapsed2 = seconds_since(begin2);
        printf("Rendered in %.3f seconds.\n", elapsed2);
        printf("Total: %.3f seconds.\n", elapsed2 + elapsed1);
        for (auto h: handles) rt_scene_destroy(h);
    } catch (std::exception &e) {
        fprintf(stderr, "%s\n", e.what());
        return 1;
    }
    return 0;
}
