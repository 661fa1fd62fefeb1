// main.cpp — the reference's command line on the B200 path:  raytracer scene.xml
//
// Same surface as the reference's main (raytracer.cpp:487-525): reads the XML scene, "plants the
// trees" (here: uploads the scene and builds the BVH on the GPU), renders every <Camera> with the
// supersampling factor, writes one ASCII-P3 PPM per camera under its <ImageName> into the current
// directory and prints the same three timing lines.  The reference fixes its AA factor at compile
// time (2, raytracer.cpp:26-28); here 2 is the default and `--aa N` selects it at run time.
//
//   raytracer scene.xml [more.xml ...] [--aa N] [--res WxH] [--gpus N] [--builder ploc|sah_gpu|lbvh|sah] [--stats]
//
// What is arranged around the C-ABI calls so that the process as a whole is fast:
//   * the CUDA context and the kernels come up on a helper thread (rt_warmup) while the main thread parses the
//     XML file — on a B200 box the driver needs 0.6-1 s for that, far more than anything else this program does;
//   * cameras are rendered back to back on the resident scene with rt_render_async into page-locked frames: the
//     device-to-host copy and the P3 file of camera i overlap the render of camera i+1 (raytracer.cpp:505-519
//     renders and writes them one after the other);
//   * with --gpus N and at least N cameras, cameras are dealt to the GPUs (camera i on GPU i % N), each GPU
//     rendering whole frames; with fewer cameras than GPUs every frame is split into row bands (rt_render_multi);
//   * several scene files may be given: they share one process, i.e. one CUDA start-up.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <future>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "rt_b200.h"
#include "scene.h"

static double seconds_since(std::chrono::high_resolution_clock::time_point t0) {
    return std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - t0).count();
}

struct Options {
    int aa = 2, gpus = 1, res_w = 0, res_h = 0, builder = RT_BUILD_DEFAULT;
    bool want_stats = false;
};

struct Frame {  // one camera's page-locked image and its pending work
    parser::Camera camera;
    unsigned char *pixels = nullptr;
    int gpu = 0, ticket = -1;
    RtStats stats;
    std::thread writer;
    std::string write_error;
};

static void print_stats(const RtStats &st) {
    unsigned long long rays = st.primary_rays + st.reflection_rays + st.shadow_rays;
    printf("  rays: %llu primary, %llu reflection, %llu shadow (%llu occluded); %.3f ms render, %.3f ms to host, %.1f Mrays/s\n",
           (unsigned long long) st.primary_rays, (unsigned long long) st.reflection_rays, (unsigned long long) st.shadow_rays,
           (unsigned long long) st.shadow_occluded, st.ms_render, st.ms_d2h, st.ms_render > 0 ? rays / (st.ms_render * 1e3) : 0.0);
}

static void render_scene(const parser::Scene &scene, const parser::FlatScene &flat, const Options &opt, int gpus) {
    auto begin1 = std::chrono::high_resolution_clock::now();
    RtBuildOptions bo;
    memset(&bo, 0, sizeof bo);
    bo.builder = opt.builder;
    std::vector<RtScene *> handles;
    std::vector<Frame> frames;
    struct Cleanup {  // every exit path: writers joined, GPU work drained (rt_scene_destroy), then the frames freed
        std::vector<RtScene *> &h;
        std::vector<Frame> &frames;
        ~Cleanup() {
            for (auto &f: frames)
                if (f.writer.joinable()) f.writer.join();
            for (auto p: h) rt_scene_destroy(p);
            for (auto &f: frames)
                if (f.pixels) rt_host_free(f.pixels);
        }
    } cleanup{handles, frames};
    for (int g = 0; g < gpus; g++) {
        if (rt_set_device(g) != RT_OK) throw std::runtime_error(rt_last_error());
        RtScene *h = nullptr;
        if (rt_scene_create(&flat.desc, &bo, &h) != RT_OK) throw std::runtime_error(rt_last_error());
        handles.push_back(h);
    }
    rt_set_device(0);
    double elapsed1 = seconds_since(begin1);
    printf("Planted trees in %.3f seconds.\n", elapsed1);
    if (opt.aa > 1) printf("Super Sampling Anti aliasing is enabled. (%d*%dx)\n", opt.aa, opt.aa);

    auto begin2 = std::chrono::high_resolution_clock::now();
    frames.resize(scene.cameras.size());
    const bool deal_cameras = gpus > 1 && (int) frames.size() >= gpus;  // whole frames per GPU instead of bands
    for (size_t i = 0; i < frames.size(); i++) {
        Frame &f = frames[i];
        f.camera = scene.cameras[i];
        if (opt.res_w > 0) f.camera.image_width = opt.res_w, f.camera.image_height = opt.res_h;
        void *p = nullptr;
        if (rt_host_alloc((int64_t) f.camera.image_width * f.camera.image_height * 3, &p) != RT_OK) throw std::runtime_error(rt_last_error());
        f.pixels = (unsigned char *) p;
        f.gpu = deal_cameras ? (int) (i % (size_t) gpus) : 0;
    }
    auto finish = [&](Frame &f) {  // wait for the frame, then hand it to a writer thread
        if (rt_wait(handles[f.gpu], f.ticket, &f.stats) != RT_OK) throw std::runtime_error(rt_last_error());
        f.ticket = -1;
        if (opt.want_stats) print_stats(f.stats);
        f.writer = std::thread([&f]() {
            try {
                write_ppm(f.camera.image_name.c_str(), f.pixels, f.camera.image_width, f.camera.image_height);
            } catch (std::exception &e) {
                f.write_error = e.what();
            }
        });
    };
    const size_t in_flight = 2;  // per GPU
    std::vector<std::vector<size_t>> pending((size_t) gpus);
    for (size_t i = 0; i < frames.size(); i++) {
        Frame &f = frames[i];
        RtCamera cam = parser::to_rt_camera(f.camera);
        printf("Rendering %s with %d B200 GPU%s...\n", f.camera.image_name.c_str(), gpus, gpus > 1 ? "s" : "");
        fflush(stdout);
        if (gpus > 1 && !deal_cameras) {
            if (rt_render_multi(handles.data(), gpus, &cam, opt.aa, f.pixels, &f.stats) != RT_OK) throw std::runtime_error(rt_last_error());
            if (opt.want_stats) print_stats(f.stats);
            f.writer = std::thread([&f]() {
                try {
                    write_ppm(f.camera.image_name.c_str(), f.pixels, f.camera.image_width, f.camera.image_height);
                } catch (std::exception &e) {
                    f.write_error = e.what();
                }
            });
            continue;
        }
        auto &q = pending[(size_t) f.gpu];
        if (q.size() >= in_flight) {
            finish(frames[q.front()]);
            q.erase(q.begin());
        }
        if (rt_render_async(handles[f.gpu], &cam, opt.aa, f.pixels, &f.ticket) != RT_OK) throw std::runtime_error(rt_last_error());
        q.push_back(i);
    }
    for (auto &q: pending)
        for (size_t i: q) finish(frames[i]);
    for (auto &f: frames)
        if (f.writer.joinable()) f.writer.join();
    for (auto &f: frames)
        if (!f.write_error.empty()) throw std::runtime_error(f.write_error);
    double elapsed2 = seconds_since(begin2);
    printf("Rendered in %.3f seconds.\n", elapsed2);
    printf("Total: %.3f seconds.\n", elapsed2 + elapsed1);
}

int main(int argc, char *argv[]) {
    std::vector<const char *> xmls;
    Options opt;
    bool help = false;
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        if (a == "--aa" && i + 1 < argc) opt.aa = atoi(argv[++i]);
        else if (a == "--gpus" && i + 1 < argc) opt.gpus = atoi(argv[++i]);
        else if (a == "--res" && i + 1 < argc) {
            if (sscanf(argv[++i], "%dx%d", &opt.res_w, &opt.res_h) != 2) {
                fprintf(stderr, "--res expects WxH\n");
                return 2;
            }
        } else if (a == "--builder" && i + 1 < argc) {
            std::string b = argv[++i];
            opt.builder = b == "lbvh" ? RT_BUILD_LBVH_GPU : b == "sah" ? RT_BUILD_SAH_HOST : b == "ploc" ? RT_BUILD_PLOC_GPU : b == "sah_gpu" ? RT_BUILD_SAH_GPU : RT_BUILD_DEFAULT;
        } else if (a == "--stats") opt.want_stats = true;
        else if (a == "--probe") {
            // nothing but CUDA start-up (driver, context, kernel load): what a process pays before any of its own work
            // (tools/cli_wall.py reports it next to the whole-process wall times)
            return rt_device_count() > 0 && rt_warmup(0) == RT_OK ? 0 : 1;
        }
        else if (a == "--help" || a == "-h") help = true;
        else xmls.push_back(argv[i]);
    }
    if (help || xmls.empty()) {
        fprintf(stderr, "usage: raytracer scene.xml [more.xml ...] [--aa N] [--res WxH] [--gpus N] [--builder ploc|sah_gpu|lbvh|sah] [--stats]\n");
        return 2;
    }
    try {
        // CUDA start-up (driver, context, kernels) on helper threads, one per GPU, while the first file is parsed
        std::future<int> warm = std::async(std::launch::async, [&opt]() {
            int have = rt_device_count();
            if (have < 1) return RT_ERR_CUDA;
            int want = opt.gpus < 1 ? 1 : (opt.gpus > have ? have : opt.gpus);
            std::vector<std::future<int>> rest;
            for (int g = 1; g < want; g++) rest.push_back(std::async(std::launch::async, [g]() { return rt_warmup(g); }));
            int rc = rt_warmup(0);
            for (auto &r: rest) {
                int e = r.get();
                if (rc == RT_OK) rc = e;
            }
            return rc;
        });
        int gpus = -1;
        for (const char *xml: xmls) {
            parser::Scene scene;  // parsing the first file overlaps the warm-up
            scene.loadFromXml(xml);
            parser::FlatScene flat;
            parser::flatten(scene, flat);
            if (gpus < 0) {
                if (warm.get() != RT_OK) throw std::runtime_error("Error: no usable CUDA device (this build has no CPU fallback).");
                const int have = rt_device_count();
                gpus = opt.gpus < 1 ? 1 : (opt.gpus > have ? have : opt.gpus);
            }
            render_scene(scene, flat, opt, gpus);
        }
    } catch (std::exception &e) {
        fprintf(stderr, "%s\n", e.what());
        return 1;
    }
    return 0;
}
