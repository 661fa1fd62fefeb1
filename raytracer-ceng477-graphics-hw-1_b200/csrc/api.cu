// api.cu — the C-ABI of include/rt_b200.h over the CUDA kernels.  No CPU fallback: without a
// usable GPU every compute entry point fails with RT_ERR_CUDA.
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <cstring>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "exact_math.h"
#include "render_params.h"
#include "rt_b200.h"
#include "rt_internal.h"

namespace rtb {
int launch_render(const RenderParams &p, int n_ctas, cudaStream_t stream);
int launch_assemble(const unsigned char *parts, long long part_stride, int part_world, int nx, int ny, int tiles_x,
                    int n_tiles, unsigned char *frame, cudaStream_t stream);
int render_kernel_occupancy(int *ctas_per_sm);
int launch_render_v2(const RenderParams &p, int n_ctas, cudaStream_t stream);
int render_kernel_v2_occupancy(int *ctas_per_sm, int *warps_per_cta);
int launch_render_v3(const RenderParams &p, int n_ctas, cudaStream_t stream);
int render_kernel_v3_occupancy(int *ctas_per_sm, int *warps_per_cta);
int build_bvh_device(const std::vector<Aabb> &bounds, HostBvh &out, float *ms_device, int use_ploc);
int build_bvh_sah_device(const std::vector<Aabb> &bounds, HostBvh &out, float *ms_device);
int build_reference_ranks_device(const RtSceneDesc &d, const std::vector<Aabb> &bounds, std::vector<uint32_t> &ranks,
                                 RefTreeStats &stats, RefTree &tree, float *ms_device);
}  // namespace rtb

using namespace rtb;

struct RtScene {
    int device = 0;
    int n_sms = 0;
    int ctas_per_sm = 1;
    int kernel = 2;  // 1: CTA-tile megakernel (render.cu), 2: warp-tile state machine (render_v2.cu); env RT_B200_KERNEL
    int ctas_per_sm2 = 1, warps_per_cta2 = 4;
    int refill_threshold = 0;  // env RT_B200_REFILL
    int ctas_per_sm3 = 1, warps_per_cta3 = 4;
    // device buffers
    float4 *d_tri_nn = nullptr;
    float4 *d_nodes = nullptr, *d_prims = nullptr, *d_tri_nm = nullptr, *d_sph_cr = nullptr, *d_materials = nullptr,
           *d_lights = nullptr;
    int *d_sph_mat = nullptr;
    uint32_t *d_ranks = nullptr;
    float4 *d_ref_nodes = nullptr, *d_prim_bounds = nullptr;
    int *d_ref_leaf_prims = nullptr, *d_slot_of_prim = nullptr;
    unsigned int *d_counter = nullptr;
    unsigned long long *d_stats = nullptr;
    unsigned char *d_frame = nullptr;  // grows on demand
    size_t frame_cap = 0;
    unsigned char *d_parts = nullptr;  // rt_render_multi gather buffer on the root device
    size_t parts_cap = 0;
    unsigned char *h_pinned = nullptr;  // staging for D2H into pageable caller memory
    size_t pinned_cap = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    // scene constants
    RenderParams base;
    RtSceneInfo info;
};

namespace {

thread_local std::string g_err;

int fail(int code, const std::string &msg) {
    g_err = msg;
    return code;
}

#define CU(call)                                                                                                    \
    do {                                                                                                            \
        cudaError_t e_ = (call);                                                                                    \
        if (e_ != cudaSuccess)                                                                                      \
            return fail(RT_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));                           \
    } while (0)

template <typename T>
int upload(T **dst, const void *src, size_t count) {
    size_t bytes = sizeof(T) * (count ? count : 1);
    CU(cudaMalloc((void **) dst, bytes));
    if (count) CU(cudaMemcpy(*dst, src, sizeof(T) * count, cudaMemcpyHostToDevice));
    return RT_OK;
}

double now_ms() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

int validate(const RtSceneDesc *d) {
    if (!d) return fail(RT_ERR_INVALID, "scene description is NULL");
    if (d->n_vertices < 0 || d->n_triangles < 0 || d->n_spheres < 0 || d->n_materials < 0 || d->n_lights < 0)
        return fail(RT_ERR_INVALID, "negative count in scene description");
    if ((d->n_vertices && !d->vertices) || (d->n_triangles && !d->triangles) || (d->n_spheres && !d->spheres) ||
        (d->n_materials && !d->materials) || (d->n_lights && !d->lights))
        return fail(RT_ERR_INVALID, "NULL array with non-zero count");
    if (d->max_recursion_depth > kMaxSupportedDepth)
        return fail(RT_ERR_INVALID, "max_recursion_depth above the supported 32");
    auto vid = [&](int id) { return id >= 1 && id <= d->n_vertices; };
    auto mid = [&](int id) { return id >= 1 && id <= d->n_materials; };
    for (int i = 0; i < d->n_triangles; i++) {
        const RtTriangle &t = d->triangles[i];
        if (!vid(t.v0_id) || !vid(t.v1_id) || !vid(t.v2_id)) return fail(RT_ERR_INVALID, "triangle vertex id out of range");
        if (!mid(t.material_id)) return fail(RT_ERR_INVALID, "triangle material id out of range");
    }
    for (int i = 0; i < d->n_spheres; i++) {
        if (!vid(d->spheres[i].center_vertex_id)) return fail(RT_ERR_INVALID, "sphere centre id out of range");
        if (!mid(d->spheres[i].material_id)) return fail(RT_ERR_INVALID, "sphere material id out of range");
    }
    if ((long long) d->n_triangles + d->n_spheres >= (1LL << 27)) return fail(RT_ERR_INVALID, "too many primitives");
    return RT_OK;
}

int tree_depth(const HostBvh &b) {
    if (b.nodes.empty()) return 0;
    int best = 0;
    std::vector<std::pair<int, int>> st;
    st.emplace_back(0, 1);
    while (!st.empty()) {
        auto [n, dep] = st.back();
        st.pop_back();
        if (dep > best) best = dep;
        if (b.nodes[n].child0 >= 0 && b.nodes[n].child0 != kEmptyChild) st.emplace_back(b.nodes[n].child0, dep + 1);
        if (b.nodes[n].child1 >= 0 && b.nodes[n].child1 != kEmptyChild) st.emplace_back(b.nodes[n].child1, dep + 1);
    }
    return best;
}

// EyeRayGenerator::init (raytracer.cpp:292-314) in host fp32, operation for operation; `width`
// and `height` are the SUB-SAMPLE grid dimensions (main multiplies the camera by the AA factor
// before render, raytracer.cpp:506-509).
void camera_setup(const RtCamera &c, int width, int height, RenderParams &p) {
    const float e[3] = {c.position.x, c.position.y, c.position.z};
    const float w[3] = {-c.gaze.x, -c.gaze.y, -c.gaze.z};
    const float v[3] = {c.up.x, c.up.y, c.up.z};
    const float u[3] = {v[1] * w[2] - v[2] * w[1], v[2] * w[0] - v[0] * w[2], v[0] * w[1] - v[1] * w[0]};
    for (int k = 0; k < 3; k++) {
        volatile float m = e[k] + (-w[k]) * c.near_distance;
        volatile float q = m + u[k] * c.l;
        q = q + v[k] * c.t;
        p.e[k] = e[k];
        p.q[k] = q;
        p.u[k] = u[k];
        p.v[k] = v[k];
    }
    p.su_mul = (c.r - c.l) / (float) width;
    p.sv_mul = (c.t - c.b) / (float) height;
}

int ensure(unsigned char **buf, size_t *cap, size_t need, bool pinned) {
    if (*cap >= need) return RT_OK;
    if (*buf) {
        if (pinned) cudaFreeHost(*buf);
        else cudaFree(*buf);
        *buf = nullptr;
        *cap = 0;
    }
    if (pinned) CU(cudaMallocHost((void **) buf, need));
    else CU(cudaMalloc((void **) buf, need));
    *cap = need;
    return RT_OK;
}

struct FrameGeom {
    int tiles_x, tiles_y, n_tiles;
};
// Tiles are numbered row-major over a grid whose row length is made co-prime-ish with the number of parts: when
// tiles_x is a multiple of `world`, one phantom (empty) column is appended so that tile k -> part k % world walks
// diagonally over the image instead of giving every part fixed vertical stripes (load balance: +6 % at 8 GPUs).
FrameGeom geom(const RtCamera *cam, int world) {
    FrameGeom g;
    g.tiles_x = (cam->image_width + RT_TILE - 1) / RT_TILE;
    if (world > 1 && g.tiles_x % world == 0) g.tiles_x += 1;
    g.tiles_y = (cam->image_height + RT_TILE - 1) / RT_TILE;
    g.n_tiles = g.tiles_x * g.tiles_y;
    return g;
}
int64_t part_tiles(const FrameGeom &g, int rank, int world) {
    if (rank >= g.n_tiles) return 0;
    return (g.n_tiles - rank + world - 1) / world;
}

int check_render_args(RtScene *s, const RtCamera *cam, int aa, int rank, int world) {
    if (!s || !cam) return fail(RT_ERR_INVALID, "NULL scene or camera");
    if (aa < 1 || aa > 64) return fail(RT_ERR_INVALID, "aa_factor must be in [1, 64]");
    if (cam->image_width < 1 || cam->image_height < 1) return fail(RT_ERR_INVALID, "empty image");
    if ((long long) cam->image_width * aa > (1 << 24) || (long long) cam->image_height * aa > (1 << 24))
        return fail(RT_ERR_INVALID, "sub-sample grid wider than 2^24 (pixel centres would not be exact floats)");
    if (world < 1 || rank < 0 || rank >= world) return fail(RT_ERR_INVALID, "bad part_rank / part_world");
    int dev;
    CU(cudaGetDevice(&dev));
    if (dev != s->device) CU(cudaSetDevice(s->device));
    return RT_OK;
}

// enqueue one part on `stream`; stats are accumulated into s->d_stats (zeroed here)
int enqueue_part(RtScene *s, const RtCamera *cam, int aa, int rank, int world, unsigned char *d_out, int out_mode,
                 cudaStream_t stream, int *launches) {
    RenderParams p = s->base;
    camera_setup(*cam, cam->image_width * aa, cam->image_height * aa, p);
    const FrameGeom g = geom(cam, world);
    p.nx = cam->image_width;
    p.ny = cam->image_height;
    p.f = aa;
    // work item: P x P output pixels with P*f ~ 32 sub-samples a side (1024 sub-samples): a CTA tile with 4
    // sub-samples per thread (kernel 1) or a warp tile refilled lane by lane (kernel 2, P <= 16 when f > 1)
    int P = (32 + aa - 1) / aa;
    if (P > RT_TILE) P = RT_TILE;
    if (P < 1) P = 1;
    if (s->kernel >= 2) {
        if (const char *tp = getenv("RT_B200_TILE_P")) P = atoi(tp) > 0 ? atoi(tp) : P;  // experiments only
        if (aa > 1 && P > 16) P = 16;
        // small frames: shrink the warp tile (down to 8 sub-samples a side) until there are enough tiles to
        // give every resident warp a few dozen of them (dynamic load balance: tiles differ a lot in cost)
        const long long slots = s->kernel == 2 ? (long long) s->n_sms * s->ctas_per_sm2 * s->warps_per_cta2
                                               : (long long) s->n_sms * s->ctas_per_sm3 * s->warps_per_cta3;
        for (;;) {
            const long long ix = (RT_TILE + P - 1) / P;
            if (part_tiles(g, rank, world) * ix * ix >= 32 * slots || P * aa <= 8 || P == 1) break;
            P = (P + 1) / 2;
        }
    }
    int Ph = P;
    if (s->kernel >= 2 && P * aa == 8 && 4 % aa == 0) {
        // still too few tiles for the machine (a 1440x720 frame without AA has 16 K tiles of 64 pixels for 4144
        // resident warps): halve the tile once more to a single 8x4 round of sub-samples
        const long long slots = s->kernel == 2 ? (long long) s->n_sms * s->ctas_per_sm2 * s->warps_per_cta2
                                               : (long long) s->n_sms * s->ctas_per_sm3 * s->warps_per_cta3;
        const long long ix = (RT_TILE + P - 1) / P;
        if (part_tiles(g, rank, world) * ix * ix < 32 * slots) Ph = 4 / aa;
    }
    p.P = P;
    p.Ph = Ph;
    p.items_x = (RT_TILE + P - 1) / P;
    p.items_y = (RT_TILE + Ph - 1) / Ph;
    p.tiles_x = g.tiles_x;
    p.tiles_y = g.tiles_y;
    p.part_rank = rank;
    p.part_world = world;
    const long long n_items = part_tiles(g, rank, world) * (long long) p.items_x * p.items_y;
    if (n_items >= (1LL << 32) - (1 << 22)) return fail(RT_ERR_INVALID, "too many work items");
    p.n_items = (unsigned) n_items;
    p.out_mode = out_mode;
    p.refill_threshold = s->refill_threshold;
    p.out = d_out;
    p.work_counter = s->d_counter;
    p.stats = s->d_stats;
    CU(cudaMemsetAsync(s->d_counter, 0, sizeof(unsigned int), stream));
    CU(cudaMemsetAsync(s->d_stats, 0, 6 * sizeof(unsigned long long), stream));
    if (n_items == 0) return RT_OK;
    cudaError_t e;
    if (s->kernel == 2) {
        long long ctas = (long long) s->n_sms * s->ctas_per_sm2;
        const long long need = (n_items + s->warps_per_cta2 - 1) / s->warps_per_cta2;
        if (ctas > need) ctas = need;
        e = (cudaError_t) launch_render_v2(p, (int) ctas, stream);
    } else if (s->kernel == 3) {
        long long ctas = (long long) s->n_sms * s->ctas_per_sm3;
        const long long need = (n_items + s->warps_per_cta3 - 1) / s->warps_per_cta3;
        if (ctas > need) ctas = need;
        e = (cudaError_t) launch_render_v3(p, (int) ctas, stream);
    } else {
        long long ctas = (long long) s->n_sms * s->ctas_per_sm;
        if (ctas > n_items) ctas = n_items;
        e = (cudaError_t) launch_render(p, (int) ctas, stream);
    }
    if (e != cudaSuccess) return fail(RT_ERR_CUDA, std::string("render kernel launch: ") + cudaGetErrorString(e));
    if (launches) (*launches)++;
    return RT_OK;
}

int fetch_stats(RtScene *s, RtStats *stats) {
    unsigned long long h[6];
    CU(cudaMemcpy(h, s->d_stats, sizeof h, cudaMemcpyDeviceToHost));
    stats->primary_rays = h[0];
    stats->reflection_rays = h[1];
    stats->shadow_rays = h[2];
    stats->shadow_occluded = h[3];
    stats->replayed_closest = h[4];
    stats->replayed_any = h[5];
    return RT_OK;
}

}  // namespace

extern "C" {

const char *rt_last_error(void) { return g_err.c_str(); }
int rt_abi_version(void) { return RT_B200_ABI_VERSION; }

int rt_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

// ---- peer memory for the fused gather ---------------------------------------------------------------

int rt_device_alloc(int64_t bytes, void **d_ptr) {
    if (!d_ptr || bytes <= 0) return fail(RT_ERR_INVALID, "bad argument");
    CU(cudaMalloc(d_ptr, (size_t) bytes));
    CU(cudaMemset(*d_ptr, 0, (size_t) bytes));
    return RT_OK;
}

int rt_device_free(void *d_ptr) {
    CU(cudaFree(d_ptr));
    return RT_OK;
}

int rt_ipc_export(void *d_ptr, unsigned char handle[RT_IPC_HANDLE_BYTES]) {
    static_assert(sizeof(cudaIpcMemHandle_t) == RT_IPC_HANDLE_BYTES, "IPC handle size");
    if (!d_ptr || !handle) return fail(RT_ERR_INVALID, "bad argument");
    cudaIpcMemHandle_t h;
    CU(cudaIpcGetMemHandle(&h, d_ptr));
    memcpy(handle, &h, sizeof h);
    return RT_OK;
}

int rt_ipc_open(const unsigned char handle[RT_IPC_HANDLE_BYTES], void **d_ptr) {
    if (!d_ptr || !handle) return fail(RT_ERR_INVALID, "bad argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof h);
    CU(cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return RT_OK;
}

int rt_ipc_close(void *d_ptr) {
    CU(cudaIpcCloseMemHandle(d_ptr));
    return RT_OK;
}

// ---- host-only hooks (no CUDA call inside): let CPU tests pin host-side logic and closed forms ----

int rt_host_reference_ranks(const RtSceneDesc *desc, uint32_t *ranks_out, int32_t *stats4) {
    int rc = validate(desc);
    if (rc != RT_OK) return rc;
    std::vector<uint32_t> ranks;
    RefTreeStats st;
    build_reference_ranks(*desc, ranks, st);
    if (ranks_out) memcpy(ranks_out, ranks.data(), ranks.size() * sizeof(uint32_t));
    if (stats4) stats4[0] = st.nodes, stats4[1] = st.leaves, stats4[2] = st.max_leaf, stats4[3] = st.max_depth;
    return RT_OK;
}

// builds the host SAH BVH and checks its invariants: every primitive in exactly one leaf, every child box
// (after padding) contains the bounds of everything below it.  Returns the node count or a negative error.
int rt_host_check_bvh(const RtSceneDesc *desc, float *sah_cost, int32_t *max_depth) {
    int rc = validate(desc);
    if (rc != RT_OK) return rc;
    std::vector<Aabb> bounds;
    primitive_bounds(*desc, bounds);
    HostBvh bvh;
    build_bvh_sah_host(bounds, bvh);
    if (sah_cost) *sah_cost = bvh_sah_cost(bvh);
    if (max_depth) *max_depth = tree_depth(bvh);
    compact_dfs(bvh, 7);  // the layout pass every tree goes through (here with a breadth-first prefix) must keep the tree intact
    pad_boxes(bvh, bounds);
    const int np = (int) bounds.size();
    std::vector<int> seen((size_t) np, 0);
    if (np == 0) return bvh.nodes.empty() ? 0 : fail(RT_ERR_STATE, "nodes without primitives");
    struct Item { int ref; Aabb box; };
    std::vector<Item> st;
    auto child_box = [](const HostNode &n, int c) {
        Aabb b;
        for (int k = 0; k < 3; k++) b.mn[k] = c ? n.c1mn[k] : n.c0mn[k], b.mx[k] = c ? n.c1mx[k] : n.c0mx[k];
        return b;
    };
    auto inside = [](const Aabb &in, const Aabb &out) {
        for (int k = 0; k < 3; k++)
            if (in.mn[k] < out.mn[k] || in.mx[k] > out.mx[k]) return false;
        return true;
    };
    Aabb all = {{-INFINITY, -INFINITY, -INFINITY}, {INFINITY, INFINITY, INFINITY}};
    st.push_back({0, all});
    while (!st.empty()) {
        Item it = st.back();
        st.pop_back();
        if (it.ref == kEmptyChild) continue;
        if (it.ref < 0) {
            int enc = ~it.ref, first = enc >> 3, count = (enc & 7) + 1;
            for (int s = first; s < first + count; s++) {
                if (s >= np) return fail(RT_ERR_STATE, "leaf range out of bounds");
                int id = bvh.prim_order[s];
                seen[id]++;
                if (!inside(bounds[id], it.box)) return fail(RT_ERR_STATE, "primitive outside its leaf box");
            }
            continue;
        }
        const HostNode &n = bvh.nodes[it.ref];
        for (int c = 0; c < 2; c++) {
            int ch = c ? n.child1 : n.child0;
            if (ch == kEmptyChild) continue;
            st.push_back({ch, child_box(n, c)});
        }
    }
    for (int i = 0; i < np; i++)
        if (seen[i] != 1) return fail(RT_ERR_STATE, "primitive not in exactly one leaf");
    return (int) bvh.nodes.size();
}

// FNV-1a over the reference tree (boxes, axes, child links, leaf ranges, leaf order): equal hashes <=> equal trees
static uint64_t ref_tree_hash(const RefTree &t) {
    uint64_t h = 1469598103934665603ull;
    auto mix = [&](const void *p, size_t n) {
        const unsigned char *b = (const unsigned char *) p;
        for (size_t i = 0; i < n; i++) h = (h ^ b[i]) * 1099511628211ull;
    };
    for (auto &n: t.nodes) {
        mix(n.mn, sizeof n.mn);
        mix(n.mx, sizeof n.mx);
        const int v[5] = {n.is_leaf ? 0 : n.axis, n.is_leaf, n.is_leaf ? -1 : n.right, n.first, n.count};
        mix(v, sizeof v);
    }
    mix(t.leaf_prims.data(), t.leaf_prims.size() * sizeof(int));
    mix(t.leaf_of_prim.data(), t.leaf_of_prim.size() * sizeof(int));
    return h;
}

int rt_host_reference_tree_hash(const RtSceneDesc *desc, uint64_t *hash) {
    int rc = validate(desc);
    if (rc != RT_OK) return rc;
    std::vector<uint32_t> ranks;
    RefTreeStats st;
    RefTree tree;
    build_reference_ranks(*desc, ranks, st, &tree);
    *hash = ref_tree_hash(tree);
    return RT_OK;
}

// the GPU build of the same tree and ranks (ref_order_device.cu); needs a device
int rt_device_reference_ranks(const RtSceneDesc *desc, uint32_t *ranks_out, int32_t *stats4, uint64_t *tree_hash) {
    int rc = validate(desc);
    if (rc != RT_OK) return rc;
    std::vector<Aabb> bounds;
    primitive_bounds(*desc, bounds);
    std::vector<uint32_t> ranks;
    RefTreeStats st;
    RefTree tree;
    if (build_reference_ranks_device(*desc, bounds, ranks, st, tree, nullptr) != 0) return fail(RT_ERR_CUDA, "device build failed");
    if (ranks_out) memcpy(ranks_out, ranks.data(), ranks.size() * sizeof(uint32_t));
    if (stats4) stats4[0] = st.nodes, stats4[1] = st.leaves, stats4[2] = st.max_leaf, stats4[3] = st.max_depth;
    if (tree_hash) *tree_hash = ref_tree_hash(tree);
    return RT_OK;
}

float rt_host_pow_ref(float base, float e) { return pow_ref(base, e); }
int rt_host_specular_gate(float cos_theta) { return specular_gate(cos_theta) ? 1 : 0; }

int rt_set_device(int device) {
    CU(cudaSetDevice(device));
    return RT_OK;
}

int rt_scene_create(const RtSceneDesc *desc, const RtBuildOptions *opts, RtScene **out) {
    if (!out) return fail(RT_ERR_INVALID, "out is NULL");
    *out = nullptr;
    int rc = validate(desc);
    if (rc != RT_OK) return rc;
    int dev = 0;
    CU(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, dev));

    RtScene *s = new RtScene();
    s->device = dev;
    s->n_sms = prop.multiProcessorCount;
    memset(&s->info, 0, sizeof s->info);
    memset(&s->base, 0, sizeof s->base);
    const int nt = desc->n_triangles, ns = desc->n_spheres, np = nt + ns;
    int builder = opts ? opts->builder : RT_BUILD_DEFAULT;
    if (builder == RT_BUILD_DEFAULT) builder = RT_BUILD_AUTO;

    const double t0 = now_ms();
    // reference-order tie ranks
    std::vector<uint32_t> ranks;
    RefTreeStats rstats;
    RefTree rtree;
    std::vector<Aabb> bounds;
    primitive_bounds(*desc, bounds);
    float ms_ranks_device = 0;
    double ms_ranks_wall = 0;
    if (getenv("RT_B200_HOST_RANKS")) {  // the host implementation (ref_order.cpp) stays as the cross-check
        build_reference_ranks(*desc, ranks, rstats, &rtree);
    } else {
        const double tr0 = now_ms();
        if (build_reference_ranks_device(*desc, bounds, ranks, rstats, rtree, &ms_ranks_device) != 0) {
            delete s;
            return fail(RT_ERR_CUDA, "device build of the reference-order tree failed");
        }
        ms_ranks_wall = now_ms() - tr0;
    }

    HostBvh bvh;
    float ms_device = 0;
    double ms_device_wall = 0;  // includes CUDA context creation on the first call; not host build work
    if (builder == RT_BUILD_LBVH_GPU || builder == RT_BUILD_PLOC_GPU || builder == RT_BUILD_AUTO) {
        const double td0 = now_ms();
        int e = build_bvh_device(bounds, bvh, &ms_device, builder != RT_BUILD_LBVH_GPU);
        ms_device_wall = now_ms() - td0;
        if (e != 0) {
            delete s;
            return fail(RT_ERR_CUDA, "device BVH build failed");
        }
        if (builder == RT_BUILD_AUTO) {
            // AUTO: both trees are built on the GPU.  The PLOC tree is kept when its SAH cost is clearly lower than
            // the top-down binned-SAH tree's (< 0.8x: scenes with huge primitives next to dense meshes, e.g.
            // horse_and_mug 4.4 vs 7.2); otherwise the shallower top-down tree traverses 4-12 % faster
            // (tools/ploc_tune.py).
            HostBvh sah_tree;
            float ms_sah = 0;
            if (build_bvh_sah_device(bounds, sah_tree, &ms_sah) != 0) {
                delete s;
                return fail(RT_ERR_CUDA, "device SAH build failed");
            }
            ms_device += ms_sah;
            if (tree_depth(bvh) > 60 || !(bvh_sah_cost(bvh) < 0.8f * bvh_sah_cost(sah_tree))) {
                bvh = sah_tree;
                builder = RT_BUILD_SAH_GPU;
            } else {
                builder = RT_BUILD_PLOC_GPU;
            }
            ms_device_wall = now_ms() - td0;
        }
    } else if (builder == RT_BUILD_SAH_GPU) {
        const double td0 = now_ms();
        if (build_bvh_sah_device(bounds, bvh, &ms_device) != 0) {
            delete s;
            return fail(RT_ERR_CUDA, "device SAH build failed");
        }
        ms_device_wall = now_ms() - td0;
    } else {
        build_bvh_sah_host(bounds, bvh);
    }
    bvh.max_depth = tree_depth(bvh);
    if (bvh.max_depth > 60 && builder != RT_BUILD_SAH_HOST) {
        // a pathological primitive order made the clustered tree too deep for the traversal stack: the host
        // builder splits at the median when the SAH finds nothing and stays O(log n) deep
        builder = RT_BUILD_SAH_HOST;
        build_bvh_sah_host(bounds, bvh);
        bvh.max_depth = tree_depth(bvh);
    }
    if (bvh.max_depth > 60) {
        delete s;
        return fail(RT_ERR_STATE, "BVH deeper than the traversal stack");
    }
    compact_dfs(bvh, getenv("RT_B200_BFS_TOP") ? atoi(getenv("RT_B200_BFS_TOP")) : 0);
    const float sah = bvh_sah_cost(bvh);
    pad_boxes(bvh, bounds);
    // A single-primitive scene has a root with one real child.  The missing child becomes a leaf over a dummy
    // all-zero triangle (slot np: detA = 0, every comparison on NaN fails, it can never report a hit), so the
    // traversal loop needs no "empty child" test.
    for (auto &n: bvh.nodes) {
        if (n.child1 == kEmptyChild) {
            n.child1 = ~((np << 3) | 0);
            for (int k = 0; k < 3; k++) n.c1mn[k] = n.c0mn[k], n.c1mx[k] = n.c0mx[k];
        }
        if (n.child0 == kEmptyChild) {
            n.child0 = ~((np << 3) | 0);
            for (int k = 0; k < 3; k++) n.c0mn[k] = n.c1mn[k], n.c0mx[k] = n.c1mx[k];
        }
    }

    // stage SoA buffers
    std::vector<float4> nodes(bvh.nodes.size() * 4);
    for (size_t i = 0; i < bvh.nodes.size(); i++) {
        const HostNode &n = bvh.nodes[i];
        // per axis: (centre, half-extent) of the padded child box; the half-extent is rounded up so that the
        // stored box still contains the padded one (device_common.cuh slab())
        float c0[3], h0[3], c1[3], h1[3];
        for (int k = 0; k < 3; k++) {
            c0[k] = 0.5f * (n.c0mn[k] + n.c0mx[k]);
            c1[k] = 0.5f * (n.c1mn[k] + n.c1mx[k]);
            h0[k] = std::max(n.c0mx[k] - c0[k], c0[k] - n.c0mn[k]) * 1.000001f + std::fabs(c0[k]) * 2e-7f;
            h1[k] = std::max(n.c1mx[k] - c1[k], c1[k] - n.c1mn[k]) * 1.000001f + std::fabs(c1[k]) * 2e-7f;
        }
        nodes[4 * i + 0] = make_float4(c0[0], h0[0], c0[1], h0[1]);
        nodes[4 * i + 1] = make_float4(c1[0], h1[0], c1[1], h1[1]);
        nodes[4 * i + 2] = make_float4(c0[2], h0[2], c1[2], h1[2]);
        nodes[4 * i + 3] = make_float4(__builtin_bit_cast(float, n.child0), __builtin_bit_cast(float, n.child1), 0.f, 0.f);
    }
    auto bits = [](int v) { return __builtin_bit_cast(float, v); };
    std::vector<float4> prims((size_t) (np + 1) * 3, make_float4(0.f, 0.f, 0.f, 0.f));
    for (int sidx = 0; sidx < np; sidx++) {
        const int id = bvh.prim_order[sidx];
        if (id < nt) {
            const RtTriangle &t = desc->triangles[id];
            const RtVec3 &a = desc->vertices[t.v0_id - 1], &b = desc->vertices[t.v1_id - 1], &c = desc->vertices[t.v2_id - 1];
            // raytracer.cpp:135-138: a - b and a - c, the same fp32 subtractions done once
            volatile float abx = a.x - b.x, aby = a.y - b.y, abz = a.z - b.z;
            volatile float acx = a.x - c.x, acy = a.y - c.y, acz = a.z - c.z;
            volatile float p1 = aby * acz, p2 = acy * abz;
            volatile float mn = p1 - p2;  // det()'s m10*m21 - m11*m20 of raytracer.cpp:18
            prims[3 * (size_t) sidx + 0] = make_float4(a.x, a.y, a.z, bits(id));
            prims[3 * (size_t) sidx + 1] = make_float4(abx, aby, abz, bits(0));
            prims[3 * (size_t) sidx + 2] = make_float4(acx, acy, acz, mn);
        } else {
            const RtSphere &sp = desc->spheres[id - nt];
            const RtVec3 &c = desc->vertices[sp.center_vertex_id - 1];
            prims[3 * (size_t) sidx + 0] = make_float4(c.x, c.y, c.z, bits(id));
            prims[3 * (size_t) sidx + 1] = make_float4(sp.radius, 0.f, 0.f, bits(1));
            prims[3 * (size_t) sidx + 2] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    std::vector<float4> prim_bounds((size_t) np * 2);
    for (int i = 0; i < np; i++) {
        prim_bounds[2 * (size_t) i] = make_float4(bounds[i].mn[0], bounds[i].mn[1], bounds[i].mn[2], 0.f);
        prim_bounds[2 * (size_t) i + 1] = make_float4(bounds[i].mx[0], bounds[i].mx[1], bounds[i].mx[2], 0.f);
    }
    std::vector<int> slot_of_prim((size_t) np);
    for (int sidx = 0; sidx < np; sidx++) slot_of_prim[bvh.prim_order[sidx]] = sidx;
    std::vector<float4> ref_nodes(rtree.nodes.size() * 3);
    for (size_t i = 0; i < rtree.nodes.size(); i++) {
        const RefTreeNode &n = rtree.nodes[i];
        ref_nodes[3 * i + 0] = make_float4(n.mn[0], n.mn[1], n.mn[2], bits(n.axis | (n.is_leaf ? 4 : 0)));
        ref_nodes[3 * i + 1] = make_float4(n.mx[0], n.mx[1], n.mx[2], bits(n.right));
        ref_nodes[3 * i + 2] = make_float4(bits(n.first), bits(n.count), 0.f, 0.f);
    }
    std::vector<float4> tri_nm((size_t) nt), tri_nn((size_t) nt);
    for (int i = 0; i < nt; i++) {
        const RtTriangle &t = desc->triangles[i];
        const RtVec3 &a = desc->vertices[t.v0_id - 1], &b = desc->vertices[t.v1_id - 1], &c = desc->vertices[t.v2_id - 1];
        // raytracer.cpp:346  ((b - a) x (c - a)).normalize()
        volatile float bax = b.x - a.x, bay = b.y - a.y, baz = b.z - a.z;
        volatile float cax = c.x - a.x, cay = c.y - a.y, caz = c.z - a.z;
        volatile float x1 = bay * caz, x2 = baz * cay, y1 = baz * cax, y2 = bax * caz, z1 = bax * cay, z2 = bay * cax;
        volatile float nx = x1 - x2, ny = y1 - y2, nz = z1 - z2;
        volatile float xx = nx * nx, yy = ny * ny, zz = nz * nz;
        volatile float s2 = xx + yy;
        s2 = s2 + zz;
        const float len = (float) std::sqrt((double) s2);
        tri_nm[i] = make_float4(nx / len, ny / len, nz / len, bits(t.material_id));
        // intersection.normal.normalize() of an already unit-length normal (raytracer.cpp:414, :432): same ops, once
        volatile float ux = nx / len, uy = ny / len, uz = nz / len;
        volatile float uxx = ux * ux, uyy = uy * uy, uzz = uz * uz;
        volatile float u2 = uxx + uyy;
        u2 = u2 + uzz;
        const float ulen = (float) std::sqrt((double) u2);
        tri_nn[i] = make_float4(ux / ulen, uy / ulen, uz / ulen, 0.f);
    }
    std::vector<float4> sph_cr((size_t) ns);
    std::vector<int> sph_mat((size_t) ns);
    for (int i = 0; i < ns; i++) {
        const RtVec3 &c = desc->vertices[desc->spheres[i].center_vertex_id - 1];
        sph_cr[i] = make_float4(c.x, c.y, c.z, desc->spheres[i].radius);
        sph_mat[i] = desc->spheres[i].material_id;
    }
    std::vector<float4> mats((size_t) desc->n_materials * 4);
    for (int i = 0; i < desc->n_materials; i++) {
        const RtMaterial &m = desc->materials[i];
        mats[4 * (size_t) i + 0] = make_float4(m.ambient.x, m.ambient.y, m.ambient.z, m.phong_exponent);
        mats[4 * (size_t) i + 1] = make_float4(m.diffuse.x, m.diffuse.y, m.diffuse.z, bits(m.is_mirror ? 1 : 0));
        mats[4 * (size_t) i + 2] = make_float4(m.specular.x, m.specular.y, m.specular.z, 0.f);
        mats[4 * (size_t) i + 3] = make_float4(m.mirror.x, m.mirror.y, m.mirror.z, 0.f);
    }
    std::vector<float4> lights((size_t) desc->n_lights * 2);
    for (int i = 0; i < desc->n_lights; i++) {
        const RtPointLight &l = desc->lights[i];
        lights[2 * (size_t) i + 0] = make_float4(l.position.x, l.position.y, l.position.z, 0.f);
        lights[2 * (size_t) i + 1] = make_float4(l.intensity.x, l.intensity.y, l.intensity.z, 0.f);
    }
    const double t1 = now_ms();

    rc = upload(&s->d_nodes, nodes.data(), nodes.size());
    if (rc == RT_OK) rc = upload(&s->d_prims, prims.data(), prims.size());
    if (rc == RT_OK) rc = upload(&s->d_tri_nm, tri_nm.data(), tri_nm.size());
    if (rc == RT_OK) rc = upload(&s->d_tri_nn, tri_nn.data(), tri_nn.size());
    if (rc == RT_OK) rc = upload(&s->d_sph_cr, sph_cr.data(), sph_cr.size());
    if (rc == RT_OK) rc = upload(&s->d_sph_mat, sph_mat.data(), sph_mat.size());
    if (rc == RT_OK) rc = upload(&s->d_ranks, ranks.data(), ranks.size());
    if (rc == RT_OK) rc = upload(&s->d_ref_nodes, ref_nodes.data(), ref_nodes.size());
    if (rc == RT_OK) rc = upload(&s->d_ref_leaf_prims, rtree.leaf_prims.data(), rtree.leaf_prims.size());
    if (rc == RT_OK) rc = upload(&s->d_prim_bounds, prim_bounds.data(), prim_bounds.size());
    if (rc == RT_OK) rc = upload(&s->d_slot_of_prim, slot_of_prim.data(), slot_of_prim.size());
    if (rc == RT_OK) rc = upload(&s->d_materials, mats.data(), mats.size());
    if (rc == RT_OK) rc = upload(&s->d_lights, lights.data(), lights.size());
    if (rc == RT_OK && cudaMalloc((void **) &s->d_counter, sizeof(unsigned int)) != cudaSuccess) rc = fail(RT_ERR_CUDA, "cudaMalloc");
    if (rc == RT_OK && cudaMalloc((void **) &s->d_stats, 6 * sizeof(unsigned long long)) != cudaSuccess) rc = fail(RT_ERR_CUDA, "cudaMalloc");
    if (rc == RT_OK && cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking) != cudaSuccess) rc = fail(RT_ERR_CUDA, "cudaStreamCreate");
    for (int i = 0; i < 4 && rc == RT_OK; i++)
        if (cudaEventCreate(&s->ev[i]) != cudaSuccess) rc = fail(RT_ERR_CUDA, "cudaEventCreate");
    if (rc == RT_OK) {
        int occ = 0;
        if (render_kernel_occupancy(&occ) != 0 || occ < 1) rc = fail(RT_ERR_CUDA, "render kernel cannot be resident on this device (built for sm_100a)");
        s->ctas_per_sm = occ;
        int occ2 = 0;
        if (render_kernel_v2_occupancy(&occ2, &s->warps_per_cta2) != 0 || occ2 < 1) rc = fail(RT_ERR_CUDA, "render kernel v2 cannot be resident on this device");
        s->ctas_per_sm2 = occ2;
        const char *kv = getenv("RT_B200_KERNEL");
        if (kv && kv[0] == '1') s->kernel = 1;
        if (kv && kv[0] == '3') s->kernel = 3;
        int occ3 = 0;
        if (render_kernel_v3_occupancy(&occ3, &s->warps_per_cta3) != 0 || occ3 < 1) rc = fail(RT_ERR_CUDA, "render kernel v3 cannot be resident on this device");
        s->ctas_per_sm3 = occ3;
        const char *rv = getenv("RT_B200_REFILL");
        if (rv) s->refill_threshold = atoi(rv);
    }
    if (rc != RT_OK) {
        rt_scene_destroy(s);
        return rc;
    }

    RenderParams &b = s->base;
    b.nodes = s->d_nodes;
    b.prims = s->d_prims;
    b.tri_nm = s->d_tri_nm;
    b.tri_nn = s->d_tri_nn;
    b.sph_cr = s->d_sph_cr;
    b.sph_mat = s->d_sph_mat;
    b.ranks = s->d_ranks;
    b.ref_nodes = s->d_ref_nodes;
    b.ref_leaf_prims = s->d_ref_leaf_prims;
    b.prim_bounds = s->d_prim_bounds;
    b.slot_of_prim = s->d_slot_of_prim;
    b.exact_culling = (opts && opts->no_exact_culling) ? 0 : 1;
    b.materials = s->d_materials;
    b.lights = s->d_lights;
    b.n_nodes = (int) bvh.nodes.size();
    b.n_tris = nt;
    b.n_prims = np;
    b.n_lights = desc->n_lights;
    b.max_depth = desc->max_recursion_depth;
    b.brute_force = opts ? opts->brute_force : 0;
    b.eps = desc->shadow_ray_epsilon;
    b.ambient[0] = desc->ambient_light.x, b.ambient[1] = desc->ambient_light.y, b.ambient[2] = desc->ambient_light.z;
    for (int k = 0; k < 3; k++) b.background[k] = (float) desc->background[k];  // raytracer.cpp:446-447

    RtSceneInfo &inf = s->info;
    inf.n_triangles = nt;
    inf.n_spheres = ns;
    inf.bvh_nodes = (int) bvh.nodes.size();
    inf.bvh_max_depth = bvh.max_depth;
    inf.ref_tree_nodes = rstats.nodes;
    inf.ref_tree_leaves = rstats.leaves;
    inf.ref_tree_max_leaf = rstats.max_leaf;
    inf.ref_tree_max_depth = rstats.max_depth;
    inf.ms_build_host = (float) (t1 - t0 - ms_device_wall - ms_ranks_wall);
    inf.ms_build_device = ms_device + ms_ranks_device;
    inf.bvh_sah_cost = sah;
    inf.builder = builder;
    inf.device = dev;
    *out = s;
    return RT_OK;
}

void rt_scene_destroy(RtScene *s) {
    if (!s) return;
    int prev = 0;
    cudaGetDevice(&prev);
    cudaSetDevice(s->device);
    cudaFree(s->d_nodes);
    cudaFree(s->d_prims);
    cudaFree(s->d_tri_nm);
    cudaFree(s->d_tri_nn);
    cudaFree(s->d_sph_cr);
    cudaFree(s->d_sph_mat);
    cudaFree(s->d_ranks);
    cudaFree(s->d_ref_nodes);
    cudaFree(s->d_ref_leaf_prims);
    cudaFree(s->d_prim_bounds);
    cudaFree(s->d_slot_of_prim);
    cudaFree(s->d_materials);
    cudaFree(s->d_lights);
    cudaFree(s->d_counter);
    cudaFree(s->d_stats);
    cudaFree(s->d_frame);
    cudaFree(s->d_parts);
    if (s->h_pinned) cudaFreeHost(s->h_pinned);
    if (s->stream) cudaStreamDestroy(s->stream);
    for (auto &e: s->ev)
        if (e) cudaEventDestroy(e);
    cudaSetDevice(prev);
    delete s;
}

int rt_scene_info(const RtScene *s, RtSceneInfo *info) {
    if (!s || !info) return fail(RT_ERR_INVALID, "NULL argument");
    *info = s->info;
    return RT_OK;
}

int64_t rt_part_tiles(const RtCamera *cam, int part_rank, int part_world) {
    if (!cam || part_world < 1 || part_rank < 0 || part_rank >= part_world) return -1;
    return part_tiles(geom(cam, part_world), part_rank, part_world);
}

int64_t rt_part_bytes(const RtCamera *cam, int part_rank, int part_world) {
    int64_t t = rt_part_tiles(cam, part_rank, part_world);
    return t < 0 ? t : t * RT_TILE * RT_TILE * 3;
}

int rt_render(RtScene *s, const RtCamera *cam, int aa, unsigned char *rgb_out, RtStats *stats) {
    int rc = check_render_args(s, cam, aa, 0, 1);
    if (rc != RT_OK) return rc;
    if (!rgb_out) return fail(RT_ERR_INVALID, "rgb_out is NULL");
    const size_t bytes = (size_t) cam->image_width * cam->image_height * 3;
    rc = ensure(&s->d_frame, &s->frame_cap, bytes, false);
    if (rc != RT_OK) return rc;
    cudaPointerAttributes attr;
    bool pinned_dst = cudaPointerGetAttributes(&attr, rgb_out) == cudaSuccess && attr.type == cudaMemoryTypeHost;
    cudaGetLastError();
    if (!pinned_dst) {
        rc = ensure(&s->h_pinned, &s->pinned_cap, bytes, true);
        if (rc != RT_OK) return rc;
    }
    int launches = 0;
    CU(cudaEventRecord(s->ev[0], s->stream));
    rc = enqueue_part(s, cam, aa, 0, 1, s->d_frame, kOutFrame, s->stream, &launches);
    if (rc != RT_OK) return rc;
    CU(cudaEventRecord(s->ev[1], s->stream));
    CU(cudaMemcpyAsync(pinned_dst ? rgb_out : s->h_pinned, s->d_frame, bytes, cudaMemcpyDeviceToHost, s->stream));
    CU(cudaEventRecord(s->ev[2], s->stream));
    CU(cudaStreamSynchronize(s->stream));
    if (!pinned_dst) memcpy(rgb_out, s->h_pinned, bytes);
    if (stats) {
        memset(stats, 0, sizeof *stats);
        rc = fetch_stats(s, stats);
        if (rc != RT_OK) return rc;
        cudaEventElapsedTime(&stats->ms_render, s->ev[0], s->ev[1]);
        cudaEventElapsedTime(&stats->ms_d2h, s->ev[1], s->ev[2]);
        cudaEventElapsedTime(&stats->ms_total, s->ev[0], s->ev[2]);
        stats->n_launches = launches;
    }
    return RT_OK;
}

static int render_part_common(RtScene *s, const RtCamera *cam, int aa, int rank, int world, void *d_out, int mode,
                              void *cuda_stream, RtStats *stats) {
    int rc = check_render_args(s, cam, aa, rank, world);
    if (rc != RT_OK) return rc;
    if (!d_out) return fail(RT_ERR_INVALID, "device output pointer is NULL");
    cudaStream_t st = (cudaStream_t) cuda_stream;
    int launches = 0;
    if (stats) CU(cudaEventRecord(s->ev[0], st));
    rc = enqueue_part(s, cam, aa, rank, world, (unsigned char *) d_out, mode, st, &launches);
    if (rc != RT_OK) return rc;
    if (stats) {
        CU(cudaEventRecord(s->ev[1], st));
        CU(cudaStreamSynchronize(st));
        memset(stats, 0, sizeof *stats);
        rc = fetch_stats(s, stats);
        if (rc != RT_OK) return rc;
        cudaEventElapsedTime(&stats->ms_render, s->ev[0], s->ev[1]);
        stats->ms_total = stats->ms_render;
        stats->n_launches = launches;
    }
    return RT_OK;
}

int rt_render_part(RtScene *s, const RtCamera *cam, int aa, int rank, int world, void *d_tiles, void *cuda_stream,
                   RtStats *stats) {
    return render_part_common(s, cam, aa, rank, world, d_tiles, kOutPacked, cuda_stream, stats);
}

int rt_render_part_into_frame(RtScene *s, const RtCamera *cam, int aa, int rank, int world, void *d_frame,
                              void *cuda_stream, RtStats *stats) {
    return render_part_common(s, cam, aa, rank, world, d_frame, kOutFrame, cuda_stream, stats);
}

int rt_assemble_tiles(const RtCamera *cam, int part_world, const void *d_parts, int64_t part_stride_bytes, void *d_frame,
                      void *cuda_stream) {
    if (!cam || !d_parts || !d_frame || part_world < 1) return fail(RT_ERR_INVALID, "bad argument");
    const FrameGeom g = geom(cam, part_world);
    cudaError_t e = (cudaError_t) launch_assemble((const unsigned char *) d_parts, part_stride_bytes, part_world, cam->image_width,
                                                  cam->image_height, g.tiles_x, g.n_tiles, (unsigned char *) d_frame,
                                                  (cudaStream_t) cuda_stream);
    if (e != cudaSuccess) return fail(RT_ERR_CUDA, std::string("assemble kernel launch: ") + cudaGetErrorString(e));
    return RT_OK;
}

int rt_render_multi(RtScene *const *scenes, int n, const RtCamera *cam, int aa, unsigned char *rgb_out, RtStats *stats) {
    if (!scenes || n < 1) return fail(RT_ERR_INVALID, "no scenes");
    if (n == 1) return rt_render(scenes[0], cam, aa, rgb_out, stats);
    for (int i = 0; i < n; i++) {
        int rc = check_render_args(scenes[i], cam, aa, i, n);
        if (rc != RT_OK) return rc;
    }
    if (!rgb_out) return fail(RT_ERR_INVALID, "rgb_out is NULL");
    RtScene *root = scenes[0];
    const size_t bytes = (size_t) cam->image_width * cam->image_height * 3;
    const int64_t stride = rt_part_bytes(cam, 0, n);  // part 0 owns the most tiles
    int prev = 0;
    cudaGetDevice(&prev);
    int launches = 0;
    // per-device packed buffers (reuse d_frame of each handle), gather buffer on the root
    CU(cudaSetDevice(root->device));
    int rc = ensure(&root->d_parts, &root->parts_cap, (size_t) stride * n, false);
    if (rc == RT_OK) rc = ensure(&root->d_frame, &root->frame_cap, bytes, false);
    if (rc == RT_OK) rc = ensure(&root->h_pinned, &root->pinned_cap, bytes, true);
    if (rc != RT_OK) return rc;
    CU(cudaEventRecord(root->ev[0], root->stream));
    for (int i = 0; i < n; i++) {
        RtScene *s = scenes[i];
        CU(cudaSetDevice(s->device));
        unsigned char *dst;
        if (i == 0) {
            dst = root->d_parts;
        } else {
            rc = ensure(&s->d_frame, &s->frame_cap, (size_t) stride, false);
            if (rc != RT_OK) return rc;
            dst = s->d_frame;
        }
        rc = enqueue_part(s, cam, aa, i, n, dst, kOutPacked, s->stream, &launches);
        if (rc != RT_OK) return rc;
        if (i != 0) {
            // one peer copy per GPU over NVLink, ordered after that GPU's render
            CU(cudaMemcpyPeerAsync(root->d_parts + (size_t) stride * i, root->device, s->d_frame, s->device,
                                   (size_t) rt_part_bytes(cam, i, n), s->stream));
            CU(cudaEventRecord(s->ev[3], s->stream));
        }
    }
    CU(cudaSetDevice(root->device));
    for (int i = 1; i < n; i++) CU(cudaStreamWaitEvent(root->stream, scenes[i]->ev[3], 0));
    CU(cudaEventRecord(root->ev[1], root->stream));
    rc = rt_assemble_tiles(cam, n, root->d_parts, stride, root->d_frame, root->stream);
    if (rc != RT_OK) return rc;
    launches++;
    CU(cudaMemcpyAsync(root->h_pinned, root->d_frame, bytes, cudaMemcpyDeviceToHost, root->stream));
    CU(cudaEventRecord(root->ev[2], root->stream));
    CU(cudaStreamSynchronize(root->stream));
    memcpy(rgb_out, root->h_pinned, bytes);
    if (stats) {
        memset(stats, 0, sizeof *stats);
        for (int i = 0; i < n; i++) {
            RtStats part;
            memset(&part, 0, sizeof part);
            CU(cudaSetDevice(scenes[i]->device));
            rc = fetch_stats(scenes[i], &part);
            if (rc != RT_OK) return rc;
            stats->primary_rays += part.primary_rays;
            stats->reflection_rays += part.reflection_rays;
            stats->shadow_rays += part.shadow_rays;
            stats->shadow_occluded += part.shadow_occluded;
            stats->replayed_closest += part.replayed_closest;
            stats->replayed_any += part.replayed_any;
        }
        CU(cudaSetDevice(root->device));
        cudaEventElapsedTime(&stats->ms_render, root->ev[0], root->ev[1]);
        cudaEventElapsedTime(&stats->ms_d2h, root->ev[1], root->ev[2]);
        cudaEventElapsedTime(&stats->ms_total, root->ev[0], root->ev[2]);
        stats->n_launches = launches;
    }
    cudaSetDevice(prev);
    return RT_OK;
}

}  // extern "C"
