// scene_build.h — what api.cu sees of the GPU-resident scene build (scene_build.cu).
#pragma once

#include <cstdint>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "build_device.h"

namespace rtb {

// render "slots": frames in flight on one scene handle (rt_render_async double-buffers); each slot owns 64 bytes of
// the control block: [0] work counter (low 32 bits), [1..6] ray statistics — zeroed with ONE memset per frame
constexpr int kControlSlots = 4;

// device pointers into the persistent arena (layouts in rt_internal.h)
struct SceneBuffers {
    unsigned long long *control = nullptr;
    BuildResult *result = nullptr;
    float4 *nodes = nullptr, *prims = nullptr, *tri_nm = nullptr, *tri_nn = nullptr, *sph_cr = nullptr, *prim_bounds = nullptr,
           *ref_nodes = nullptr, *materials = nullptr, *lights = nullptr;
    int *sph_mat = nullptr, *ref_leaf_prims = nullptr, *slot_of_prim = nullptr;
    uint32_t *ranks = nullptr;
};

// Device blocks of finished builds and destroyed scenes are kept per device and handed out again (first fit within 4x
// the requested size): creating scene after scene — a dynamic scene rebuilt every frame — then makes no cudaMalloc /
// cudaFree call at all (both cost milliseconds on a busy box, and cudaFree synchronises the device).  rt_trim() frees
// the cache.
void *block_acquire(size_t bytes, size_t *cap);   // nullptr on failure
void block_release(void *ptr, size_t cap);
void block_cache_trim();

struct SceneBuild {
    void *arena = nullptr;    // persistent: owned by the RtScene afterwards
    void *scratch = nullptr;  // builders' scratch: released by release_scratch()
    size_t arena_bytes = 0, scratch_bytes = 0;  // capacities of the two blocks (block_acquire)
    BuildResult result;       // host copy of the device result block
    bool used_host_builder = false;
    float ms_device = 0;            // CUDA events around uploads + every build kernel
    float ms_host_before_sync = 0;  // host time until everything was enqueued
    float ms_wall = 0;
    int last_builder = 0, last_ploc_grid = 1, last_reinsert_grid = 0;

    // Enqueues and completes the whole build on `stream`.  0 on success; < 0 with `err` set.
    // reinsert_rounds: rounds of insertion-based optimisation on the top-down tree (0 = none); reinsert_accept: the
    // optimised tree is kept when its SAH cost < reinsert_accept x the builder's
    int run(const RtSceneDesc &d, int builder, int ploc_radius, float ploc_leaf_cost, int reinsert_rounds, float reinsert_accept,
            int n_sms, cudaStream_t stream, SceneBuffers &out, std::string &err);
    // test hook: the reference-order tree as ref_order.cpp numbers it, read back from the scratch arena
    int read_reference_tree(const RtSceneDesc &d, cudaStream_t stream, const SceneBuffers &out, std::vector<uint32_t> &ranks,
                            RefTreeStats &stats, RefTree &tree, std::string &err);
    void release_scratch();
};

}  // namespace rtb
