// api.cu — the C-ABI of include/rt_b200.h over the CUDA kernels.  No CPU fallback: without a
// usable GPU every compute entry point fails with RT_ERR_CUDA.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <string>
#include <vector>

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cuda_runtime.h>

#include "exact_math.h"
#include "render_params.h"
#include "rt_b200.h"
#include "rt_internal.h"
#include "reinsert_core.h"
#include "scene_build.h"

namespace rtb {
int launch_assemble(const unsigned char *parts, long long part_stride, int part_world, int nx, int ny, int band_h,
                    unsigned char *frame, cudaStream_t stream);
int launch_render_v2(const RenderParams &p, int n_ctas, cudaStream_t stream);
int render_kernel_v2_occupancy(int *ctas_per_sm, int *warps_per_cta);
}  // namespace rtb

using namespace rtb;

namespace {

// one frame in flight on a scene handle (rt_render_async double-buffers; rt_render uses the next free slot)
struct RenderSlot {
    cudaEvent_t ev_start = nullptr, ev_kernel = nullptr, ev_done = nullptr;
    unsigned char *d_frame = nullptr;  // grows on demand
    size_t frame_cap = 0;
    unsigned char *h_pinned = nullptr;  // staging for D2H into pageable caller memory
    size_t pinned_cap = 0;
    unsigned long long *h_stats = nullptr;  // pinned, 8 words
    // pending frame
    bool busy = false;
    unsigned char *dst = nullptr;
    size_t bytes = 0;
    bool staged = false;
    int launches = 0;
};

constexpr int kFrameSlots = kControlSlots - 1;  // the last control slot belongs to the part calls

}  // namespace

struct RtScene {
    int device = 0;
    int n_sms = 0;
    int ctas_per_sm[2] = {1, 1};  // shared-accumulator / register-accumulator instantiation
    int warps_per_cta = 4;
    int refill_threshold = 0;
    int max_ctas_per_sm = 0;  // experiments: 0 = occupancy limit
    void *arena = nullptr;
    size_t arena_cap = 0;
    SceneBuffers buf;
    RenderSlot slot[kFrameSlots];
    unsigned char *d_parts = nullptr;  // rt_render_multi / rt_render_part_to_host: this device's packed bands
    size_t parts_cap = 0;
    cudaStream_t stream = nullptr, copy_stream = nullptr;
    cudaEvent_t ev_part[2] = {nullptr, nullptr};
    float scene_center[3] = {0, 0, 0};
    float scene_reach = 0;  // scene diagonal + largest |coordinate|
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

// switches the current device and restores the caller's on every exit path
struct DeviceGuard {
    int prev = -1, cur = -1;
    DeviceGuard() {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        cur = prev;
    }
    int use(int dev) {
        if (dev == cur) return 0;
        if (cudaSetDevice(dev) != cudaSuccess) return -1;
        cur = dev;
        return 0;
    }
    ~DeviceGuard() {
        if (prev >= 0 && cur != prev) cudaSetDevice(prev);
    }
};

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
    if ((long long) d->n_triangles + d->n_spheres >= (1LL << 26)) return fail(RT_ERR_INVALID, "too many primitives");
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

// ---- work decomposition ----------------------------------------------------------------------------------------
// A pure function of (camera, aa, world): every rank, the gathering side and the host-side assembly agree on it
// without talking to each other.
struct ItemGeom {
    int acc_mode;  // 1: f % 8 == 0, register accumulators, strips of 32 x 1 pixels
    int P, Ph;     // item size in pixels
    int rpb;       // item rows per band; the band height is Ph * rpb pixel rows
    int items_x;
    int n_bands_total;
    int group_bands, tile_items, tiles_per_group;
};
constexpr long long kNominalWarps = 148LL * 7 * 4;  // resident warps of one B200 at the kernel's occupancy
#ifndef RT_ACC_REGS
#define RT_ACC_REGS 1  // A/B builds (tools/build_variants.sh): 0 = shared-memory accumulators at every AA factor
#endif

// experiments (rt_set_partition): 0 = the defaults computed below
int g_band_rows = 0, g_strip_width = 0;
// rt_render / rt_render_async into page-locked memory: frames of at least this many bytes are written by the kernel itself
// (rt_set_zero_copy).  Measured (tools/e2e_zero_copy.py): the kernel's stores cross PCIe at ~15 GB/s against ~32 GB/s for
// the copy engine, so it only pays when the kernel runs much longer than the transfer — 8K 16x: 655.9 -> 654.6 ms,
// horse_and_mug 1440x720: 0.468 -> 0.446 ms, but simple 800x800: 0.104 -> 0.151 ms.
long long g_zero_copy_min = 32LL << 20;
constexpr int kZeroCopyPartAA = 8;  // rt_render_part_to_host / rt_render_multi: parts of frames supersampled at least this much

int g_block_width = 0;  // experiments (rt_set_block_width): 0 = block_width_log2()'s choice
int block_width_log2(int world) {
    if (g_block_width >= 4 && g_block_width <= 512) return 31 - __builtin_clz((unsigned) g_block_width);
    (void) world;
    return 5;  // 32 x 32 pixels.  Flatter blocks for multi-GPU parts (whose rows are `world` image rows apart) were tried: no gain
}

ItemGeom item_geometry(const RtCamera *cam, int aa, int world) {
    ItemGeom g;
    const int nx = cam->image_width, ny = cam->image_height;
    if (aa % 8 == 0 && RT_ACC_REGS) {
        // one pixel row per item row; the kernel claims runs of up to P pixels of a row with guided self-scheduling
        // (render_v2.cu).  Measured (tools/partition_experiment.py, 8K 16x): runs of 4 pixels beat 32 by 0.6 % on the whole
        // frame and by 3 % on a 1/8 part (neighbouring warps then work on neighbouring pixels of the same rows)
        g.acc_mode = 1;
        g.P = 4;
        g.Ph = 1;
        if (g_strip_width >= 1 && g_strip_width <= 32) g.P = 1 << (31 - __builtin_clz((unsigned) g_strip_width));
    } else {
        g.acc_mode = 0;
        // P x P output pixels with P*f ~ 32 sub-samples a side (1024 sub-samples per item, P <= 16 when f > 1: the
        // accumulators live in shared memory)
        int P = (32 + aa - 1) / aa;
        if (P > 32) P = 32;
        if (aa > 1 && P > 16) P = 16;
        if (P < 1) P = 1;
        // small frames: shrink the item (down to 8 sub-samples a side) until there are enough of them to give every
        // resident warp a few dozen (dynamic load balance: items differ a lot in cost)
        auto items = [&](int p, int ph) { return (long long) ((nx + p - 1) / p) * ((ny + ph - 1) / ph) / world; };
        while (items(P, P) < 32 * kNominalWarps && P * aa > 8 && P > 1) P = (P + 1) / 2;
        if (g_strip_width >= 1 && g_strip_width < P) P = g_strip_width;  // experiments
        int Ph = P;
        // whole 8x4 blocks of sub-samples: P*f a multiple of 8 and Ph*f a multiple of 4 where the factor allows it
        // (f = 3: 11x11 pixels = 33x33 sub-samples left a fifth block column with one live lane in eight)
        {
            const int mx = 8 / std::gcd(8, aa), my = 4 / std::gcd(4, aa);
            P = std::max(mx, P / mx * mx);
            Ph = std::max(my, Ph / my * my);
            if (aa > 1 && P > 16) P = 16 / mx * mx > 0 ? 16 / mx * mx : mx;
            if (aa > 1 && Ph > 16) Ph = 16 / my * my > 0 ? 16 / my * my : my;
        }
        // still too few (a 1440x720 frame without AA has 16 K items of 64 pixels for 4144 resident warps): halve the
        // item once more to a single 8x4 round of sub-samples
        if (P * aa == 8 && 4 % aa == 0 && items(P, P) < 32 * kNominalWarps) Ph = 4 / aa;
        g.P = P;
        g.Ph = Ph;
    }
    g.rpb = g_band_rows > 0 ? std::max(1, g_band_rows / g.Ph) : 1;
    g.items_x = (nx + g.P - 1) / g.P;
    g.n_bands_total = (ny + g.Ph * g.rpb - 1) / (g.Ph * g.rpb);
    g.group_bands = std::max(1, 32 / g.Ph);
    g.tile_items = std::max(1, 32 / g.P);
    g.tiles_per_group = (g.items_x + g.tile_items - 1) / g.tile_items;
    return g;
}
int64_t part_bands(const ItemGeom &g, int rank, int world) {
    if (rank >= g.n_bands_total) return 0;
    return (g.n_bands_total - rank + world - 1) / world;
}

int check_render_args(RtScene *s, const RtCamera *cam, int aa, int rank, int world) {
    if (!s || !cam) return fail(RT_ERR_INVALID, "NULL scene or camera");
    if (aa < 1 || aa > 64) return fail(RT_ERR_INVALID, "aa_factor must be in [1, 64]");
    if (cam->image_width < 1 || cam->image_height < 1) return fail(RT_ERR_INVALID, "empty image");
    // (col + 0.5) must be an exact fp32 number (the reference forms it in double, raytracer.cpp:320): col < 2^23
    if ((long long) cam->image_width * aa > (1 << 23) || (long long) cam->image_height * aa > (1 << 23))
        return fail(RT_ERR_INVALID, "sub-sample grid wider than 2^23 (pixel centres col + 0.5 would not be exact floats)");
    if (world < 1 || rank < 0 || rank >= world) return fail(RT_ERR_INVALID, "bad part_rank / part_world");
    return RT_OK;
}

// enqueue one part on `stream`; statistics accumulate into control block `slot` (zeroed here with one memset)
int enqueue_part(RtScene *s, const RtCamera *cam, int aa, int rank, int world, unsigned char *d_out, int out_mode, int slot,
                 cudaStream_t stream, int *launches) {
    RenderParams p = s->base;
    camera_setup(*cam, cam->image_width * aa, cam->image_height * aa, p);
    const ItemGeom g = item_geometry(cam, aa, world);
    p.nx = cam->image_width;
    p.ny = cam->image_height;
    p.f = aa;
    p.P = g.P;
    p.Ph = g.Ph;
    p.items_x = g.items_x;
    p.n_bands = (int) part_bands(g, rank, world) * g.rpb;
    p.rows_per_band = g.rpb;
    p.group_bands = g.group_bands;
    p.tile_items = g.tile_items;
    p.tiles_per_group = g.tiles_per_group;
    p.part_rank = rank;
    p.part_world = world;
    const long long n_groups = (p.n_bands + g.group_bands - 1) / g.group_bands;  // p.n_bands counts item rows
    // register-accumulator mode: the work counter runs over pixel slots in block order (32 rows x 32 pixels per block)
    // (the block shape is a knob of tools/partition_experiment.py: rt_set_block_width)
    p.blk_w_log2 = block_width_log2(world);
    const int blk_w = 1 << p.blk_w_log2, blk_h = 1024 >> p.blk_w_log2;
    const long long n_items = g.acc_mode == 1 ? (long long) ((p.n_bands + blk_h - 1) / blk_h) * ((cam->image_width + blk_w - 1) / blk_w) * 1024
                                              : n_groups * g.tiles_per_group * g.tile_items * g.group_bands;
    if (n_items >= (1LL << 32) - (1 << 22)) return fail(RT_ERR_INVALID, "too many work items");
    p.n_items = (unsigned) n_items;
    p.out_mode = out_mode;
    p.acc_mode = (g.acc_mode == 1 && s->refill_threshold == 0) ? 1 : 0;
    if (g.acc_mode == 1 && p.acc_mode == 0) return fail(RT_ERR_STATE, "a non-zero refill threshold needs an AA factor that is not a multiple of 8");
    p.refill_threshold = s->refill_threshold;
    // camera far outside the scene (beyond 4x its reach): the slab arithmetic needs the widened test (render_v2.cu)
    {
        double d2 = 0;
        const float e[3] = {cam->position.x, cam->position.y, cam->position.z};
        for (int k = 0; k < 3; k++) d2 += ((double) e[k] - s->scene_center[k]) * ((double) e[k] - s->scene_center[k]);
        p.far_camera = std::sqrt(d2) > 4.0 * (double) s->scene_reach ? 1 : 0;
    }
    p.out = d_out;
    p.control = s->buf.control + 8 * slot;
    CU(cudaMemsetAsync(p.control, 0, 8 * sizeof(unsigned long long), stream));
    if (n_items == 0) return RT_OK;
    int per_sm = s->ctas_per_sm[p.acc_mode];
    if (s->max_ctas_per_sm > 0 && s->max_ctas_per_sm < per_sm) per_sm = s->max_ctas_per_sm;
    long long ctas = (long long) s->n_sms * per_sm;
    const long long units = p.acc_mode == 1 ? (long long) p.n_bands * cam->image_width : n_items;  // pixels / items to hand out
    const long long need = (units + s->warps_per_cta - 1) / s->warps_per_cta;
    if (ctas > need) ctas = need;
    if (p.acc_mode == 1) {
        p.tiles_per_group = (cam->image_width + blk_w - 1) / blk_w;
        p.guide = (unsigned) (4 * ctas * s->warps_per_cta);
    }
    const cudaError_t e = (cudaError_t) launch_render_v2(p, (int) ctas, stream);
    if (e != cudaSuccess) return fail(RT_ERR_CUDA, std::string("render kernel launch: ") + cudaGetErrorString(e));
    if (launches) (*launches)++;
    return RT_OK;
}

void stats_from_words(const unsigned long long *h, RtStats *stats) {
    stats->primary_rays = h[1];
    stats->reflection_rays = h[2];
    stats->shadow_rays = h[3];
    stats->shadow_occluded = h[4];
    stats->replayed_closest = h[5];
    stats->replayed_any = h[6];
}

// strided device-to-host copy of one part's packed bands into the rows of a row-major host frame
int copy_part_to_frame(const ItemGeom &g, const RtCamera *cam, int rank, int world, const unsigned char *d_part, unsigned char *frame,
                       cudaStream_t stream) {
    const int band_h = g.Ph * g.rpb;
    const size_t row_bytes = (size_t) cam->image_width * 3, band_bytes = row_bytes * band_h;
    const int64_t nb = part_bands(g, rank, world);
    if (nb == 0) return RT_OK;
    // the last band of the frame may be cut short by the image height
    const int64_t last_band = rank + (nb - 1) * world;
    const int64_t last_rows = std::min<int64_t>(band_h, cam->image_height - last_band * band_h);
    const int64_t full = last_rows == band_h ? nb : nb - 1;
    if (full > 0) {
        if (world == 1) CU(cudaMemcpyAsync(frame, d_part, band_bytes * full, cudaMemcpyDeviceToHost, stream));
        else CU(cudaMemcpy2DAsync(frame + band_bytes * rank, band_bytes * world, d_part, band_bytes, band_bytes, (size_t) full,
                                  cudaMemcpyDeviceToHost, stream));
    }
    if (full < nb)
        CU(cudaMemcpyAsync(frame + band_bytes * last_band, d_part + band_bytes * (nb - 1), row_bytes * last_rows, cudaMemcpyDeviceToHost, stream));
    return RT_OK;
}

bool is_pinned(const void *p) {
    cudaPointerAttributes attr;
    const bool yes = cudaPointerGetAttributes(&attr, p) == cudaSuccess && attr.type == cudaMemoryTypeHost;
    cudaGetLastError();
    return yes;
}

// FNV-1a over the reference tree (boxes, axes, child links, leaf ranges, leaf order): equal hashes <=> equal trees
uint64_t ref_tree_hash(const RefTree &t) {
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

int scene_create_impl(const RtSceneDesc *desc, const RtBuildOptions *opts, RtScene **out, SceneBuild *keep_build) {
    if (!out) return fail(RT_ERR_INVALID, "out is NULL");
    *out = nullptr;
    int rc = validate(desc);
    if (rc != RT_OK) return rc;
    const double t0 = now_ms();
    int dev = 0;
    CU(cudaGetDevice(&dev));
    RtScene *s = new RtScene();
    s->device = dev;
    memset(&s->info, 0, sizeof s->info);
    memset(&s->base, 0, sizeof s->base);
    auto bail = [&](int code, const std::string &msg) {
        rt_scene_destroy(s);
        return fail(code, msg);
    };
    if (cudaDeviceGetAttribute(&s->n_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
        return bail(RT_ERR_CUDA, "cudaDeviceGetAttribute failed (no usable CUDA device?)");
    const int nt = desc->n_triangles, ns = desc->n_spheres, np = nt + ns;
    int builder = opts ? opts->builder : RT_BUILD_DEFAULT;
    if (builder == RT_BUILD_DEFAULT) builder = RT_BUILD_AUTO;
    if (builder < RT_BUILD_LBVH_GPU || builder > RT_BUILD_SAH_GPU) return bail(RT_ERR_INVALID, "unknown builder");
    if (cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&s->copy_stream, cudaStreamNonBlocking) != cudaSuccess)
        return bail(RT_ERR_CUDA, "cudaStreamCreate failed (no usable CUDA device?)");
    int occ[2] = {0, 0};
    if (render_kernel_v2_occupancy(occ, &s->warps_per_cta) != 0 || occ[0] < 1 || occ[1] < 1)
        return bail(RT_ERR_CUDA, "render kernel cannot be resident on this device (built for sm_100a)");
    s->ctas_per_sm[0] = occ[0];
    s->ctas_per_sm[1] = occ[1];
    if (opts && opts->max_ctas_per_sm > 0) s->max_ctas_per_sm = opts->max_ctas_per_sm;
    if (opts && opts->refill_threshold > 0) s->refill_threshold = opts->refill_threshold > 31 ? 31 : opts->refill_threshold;

    SceneBuild local_build;
    SceneBuild &build = keep_build ? *keep_build : local_build;
    std::string err;
    const int radius = (opts && opts->ploc_radius > 0) ? opts->ploc_radius : 16;
    const float leaf_cost = (opts && opts->ploc_leaf_cost > 0) ? opts->ploc_leaf_cost : 1.0f;  // tools/ploc_tune.py: 1.0 beats 1.6 and 2.5
    // insertion-based optimisation of the top-down tree (reinsert_core.h): on unless switched off
    const int reinsert_rounds = !opts || opts->reinsert_rounds == 0 ? kReinsertDefaultRounds : (opts->reinsert_rounds < 0 ? 0 : opts->reinsert_rounds);
    const float reinsert_accept = (opts && opts->reinsert_accept > 0) ? opts->reinsert_accept : kReinsertAccept;
    rc = build.run(*desc, builder, radius, leaf_cost, reinsert_rounds, reinsert_accept, s->n_sms, s->stream, s->buf, err);
    s->arena = build.arena;
    s->arena_cap = build.arena_bytes;
    build.arena = nullptr;
    if (rc != 0) cudaStreamSynchronize(s->stream);  // nothing may still be running in the blocks that go back to the cache
    if (!keep_build) build.release_scratch();
    if (rc != 0) return bail(rc == -2 ? RT_ERR_STATE : RT_ERR_CUDA, "scene build: " + err);
    for (int i = 0; i < 2; i++)
        if (cudaEventCreate(&s->ev_part[i]) != cudaSuccess) return bail(RT_ERR_CUDA, "cudaEventCreate");

    const BuildResult &r = build.result;
    int kept = builder;
    if (build.used_host_builder) kept = RT_BUILD_SAH_HOST;
    else if (builder == RT_BUILD_AUTO) kept = np <= 1 ? RT_BUILD_SAH_GPU : (r.chosen == 0 ? RT_BUILD_PLOC_GPU : RT_BUILD_SAH_GPU);

    RenderParams &b = s->base;
    b.nodes = s->buf.nodes;
    b.prims = s->buf.prims;
    b.tri_nm = s->buf.tri_nm;
    b.tri_nn = s->buf.tri_nn;
    b.sph_cr = s->buf.sph_cr;
    b.sph_mat = s->buf.sph_mat;
    b.ranks = s->buf.ranks;
    b.ref_nodes = s->buf.ref_nodes;
    b.ref_leaf_prims = s->buf.ref_leaf_prims;
    b.prim_bounds = s->buf.prim_bounds;
    b.slot_of_prim = s->buf.slot_of_prim;
    b.exact_culling = (opts && opts->no_exact_culling) ? 0 : ((opts && opts->force_replay) ? 2 : 1);
    b.materials = s->buf.materials;
    b.lights = s->buf.lights;
    b.n_nodes = np > 0 ? r.n_nodes[r.chosen] : 0;
    b.n_tris = nt;
    b.n_prims = np;
    b.n_lights = desc->n_lights;
    b.max_depth = desc->max_recursion_depth;
    b.brute_force = opts ? opts->brute_force : 0;
    b.eps = desc->shadow_ray_epsilon;
    b.ambient[0] = desc->ambient_light.x, b.ambient[1] = desc->ambient_light.y, b.ambient[2] = desc->ambient_light.z;
    for (int k = 0; k < 3; k++) b.background[k] = (float) desc->background[k];  // raytracer.cpp:446-447
    if (desc->max_recursion_depth < 0) {
        // raytracer.cpp:387-389: depth 0 > max_recursion_depth, every primary ray returns black without being traced
        b.n_nodes = 0;
        b.background[0] = b.background[1] = b.background[2] = 0.0f;
    }
    if (np > 0) {
        double diag2 = 0;
        float reach = 0;
        for (int k = 0; k < 3; k++) {
            const float lo = ord2f(r.scene_bounds[k]), hi = ord2f(r.scene_bounds[3 + k]);
            s->scene_center[k] = 0.5f * (lo + hi);
            diag2 += ((double) hi - lo) * ((double) hi - lo);
            reach = std::max(reach, std::max(std::fabs(lo), std::fabs(hi)));
        }
        s->scene_reach = (float) std::sqrt(diag2) + reach;
    } else {
        s->scene_reach = INFINITY;
    }

    RtSceneInfo &inf = s->info;
    inf.n_triangles = nt;
    inf.n_spheres = ns;
    inf.bvh_nodes = b.n_nodes;
    inf.bvh_max_depth = np > 0 ? r.height[r.chosen] : 0;
    inf.ref_tree_nodes = r.ref_nodes;
    inf.ref_tree_leaves = r.ref_leaves;
    inf.ref_tree_max_leaf = r.ref_max_leaf;
    inf.ref_tree_max_depth = r.ref_max_depth;
    inf.ms_build_host = build.ms_host_before_sync;
    inf.ms_build_device = build.ms_device;
    inf.bvh_sah_cost = np > 0 ? r.sah_cost[r.chosen] : 0.0f;
    inf.builder = kept;
    inf.device = dev;
    inf.sah_cost_ploc = (builder == RT_BUILD_AUTO && np > 1) ? r.sah_cost[0] : 0.0f;
    inf.sah_cost_sah = (builder == RT_BUILD_AUTO && np > 1) ? r.sah_cost[1] : 0.0f;
    inf.reinsert_cost_before = r.reinsert_cost_before;
    inf.reinsert_cost_after = r.reinsert_cost_after;
    inf.reinsert_moves = r.reinsert_moves;
    inf.reinsert_rounds = r.reinsert_rounds;
    inf.reinsert_accepted = r.reinsert_accepted;
    inf.ms_create_wall = (float) (now_ms() - t0);
    *out = s;
    return RT_OK;
}

int slot_prepare(RenderSlot &sl) {
    if (!sl.ev_start) {
        CU(cudaEventCreate(&sl.ev_start));
        CU(cudaEventCreate(&sl.ev_kernel));
        CU(cudaEventCreate(&sl.ev_done));
        CU(cudaMallocHost((void **) &sl.h_stats, 8 * sizeof(unsigned long long)));
    }
    return RT_OK;
}

int render_part_common(RtScene *s, const RtCamera *cam, int aa, int rank, int world, void *d_out, int mode, void *cuda_stream,
                       RtStats *stats) {
    int rc = check_render_args(s, cam, aa, rank, world);
    if (rc != RT_OK) return rc;
    if (!d_out) return fail(RT_ERR_INVALID, "device output pointer is NULL");
    DeviceGuard g;
    if (g.use(s->device) != 0) return fail(RT_ERR_CUDA, "cudaSetDevice failed");
    cudaStream_t st = (cudaStream_t) cuda_stream;
    int launches = 0;
    const int k = kControlSlots - 1;  // parts use the last control slot
    if (stats) CU(cudaEventRecord(s->ev_part[0], st));
    rc = enqueue_part(s, cam, aa, rank, world, (unsigned char *) d_out, mode, k, st, &launches);
    if (rc != RT_OK) return rc;
    if (stats) {
        CU(cudaEventRecord(s->ev_part[1], st));
        unsigned long long h[8];
        CU(cudaMemcpyAsync(h, s->buf.control + 8 * k, sizeof h, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        memset(stats, 0, sizeof *stats);
        stats_from_words(h, stats);
        cudaEventElapsedTime(&stats->ms_render, s->ev_part[0], s->ev_part[1]);
        stats->ms_total = stats->ms_render;
        stats->n_launches = launches;
    }
    return RT_OK;
}

// enqueue: this part's bands into the handle's packed buffer, then straight into the rows of a host frame
int enqueue_part_to_host(RtScene *s, const RtCamera *cam, int aa, int rank, int world, unsigned char *host_frame, int *launches) {
    const ItemGeom g = item_geometry(cam, aa, world);
    const int k = kControlSlots - 1;
    // Heavily supersampled frame into a page-locked frame: the kernel stores its finished pixel runs straight into the
    // frame's rows over this GPU's PCIe link, the transfer rides along with the rendering.  What decides is kernel time
    // against transfer time, i.e. rays per output byte ~ aa^2, not the size of the part: at 16x16 the kernel runs ~100x
    // longer than its 3 bytes per pixel take to cross PCIe (8 GPUs, 8K: 0.4 ms of copy after a 72 ms kernel saved).
    void *alias = nullptr;
    const bool zero_copy = g_zero_copy_min >= 0 && aa >= kZeroCopyPartAA && is_pinned(host_frame) &&
                           cudaHostGetDevicePointer(&alias, host_frame, 0) == cudaSuccess && alias;
    cudaGetLastError();
    if (zero_copy) {
        CU(cudaEventRecord(s->ev_part[0], s->stream));
        int rc = enqueue_part(s, cam, aa, rank, world, (unsigned char *) alias, kOutFrame, k, s->stream, launches);
        if (rc != RT_OK) return rc;
        CU(cudaEventRecord(s->ev_part[1], s->stream));
        return RT_OK;
    }
    const size_t bytes = (size_t) part_bands(g, rank, world) * g.Ph * g.rpb * cam->image_width * 3;
    int rc = ensure(&s->d_parts, &s->parts_cap, bytes ? bytes : 1, false);
    if (rc != RT_OK) return rc;
    CU(cudaEventRecord(s->ev_part[0], s->stream));
    rc = enqueue_part(s, cam, aa, rank, world, s->d_parts, kOutPacked, k, s->stream, launches);
    if (rc != RT_OK) return rc;
    CU(cudaEventRecord(s->ev_part[1], s->stream));
    return copy_part_to_frame(g, cam, rank, world, s->d_parts, host_frame, s->stream);
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

int rt_set_device(int device) {
    CU(cudaSetDevice(device));
    return RT_OK;
}

// Brings the CUDA context of `device` up and loads the render kernels (CUDA loads modules and functions lazily), so
// that a caller can pay for both on a helper thread while it parses its scene file.
int rt_warmup(int device) {
    CU(cudaSetDevice(device));
    CU(cudaFree(nullptr));
    int occ[2], warps;
    if (render_kernel_v2_occupancy(occ, &warps) != 0) return fail(RT_ERR_CUDA, "render kernel cannot be loaded on this device (built for sm_100a)");
    return RT_OK;
}

// tuning: smallest frame (bytes) that rt_render / rt_render_async let the kernel write straight into a page-locked
// destination; negative = always render into device memory and copy afterwards
int rt_set_block_width(int pixels) {
    g_block_width = pixels;
    return RT_OK;
}

int rt_set_zero_copy(int64_t min_frame_bytes) {
    g_zero_copy_min = min_frame_bytes;
    return RT_OK;
}

// frees the device blocks kept for reuse by later rt_scene_create calls (scene_build.h block cache)
int rt_trim(void) {
    block_cache_trim();
    return RT_OK;
}

int rt_host_alloc(int64_t bytes, void **ptr) {
    if (!ptr || bytes <= 0) return fail(RT_ERR_INVALID, "bad argument");
    CU(cudaMallocHost(ptr, (size_t) bytes));
    return RT_OK;
}

int rt_host_free(void *ptr) {
    CU(cudaFreeHost(ptr));
    return RT_OK;
}

// ---- a host frame shared by one process per GPU (POSIX shared memory, page-locked in every process) -------------

int rt_host_frame_create(const char *name, int64_t bytes, void **ptr) {
    if (!name || !ptr || bytes <= 0) return fail(RT_ERR_INVALID, "bad argument");
    shm_unlink(name);
    const int fd = shm_open(name, O_CREAT | O_EXCL | O_RDWR, 0600);
    if (fd < 0) return fail(RT_ERR_STATE, std::string("shm_open(create) failed for ") + name);
    if (ftruncate(fd, (off_t) bytes) != 0) {
        close(fd);
        shm_unlink(name);
        return fail(RT_ERR_NOMEM, "ftruncate on the shared frame failed");
    }
    void *p = mmap(nullptr, (size_t) bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (p == MAP_FAILED) {
        shm_unlink(name);
        return fail(RT_ERR_NOMEM, "mmap of the shared frame failed");
    }
    memset(p, 0, (size_t) bytes);  // touch every page before pinning
    if (cudaHostRegister(p, (size_t) bytes, cudaHostRegisterPortable | cudaHostRegisterMapped) != cudaSuccess) {
        cudaGetLastError();
        munmap(p, (size_t) bytes);
        shm_unlink(name);
        return fail(RT_ERR_CUDA, "cudaHostRegister of the shared frame failed");
    }
    *ptr = p;
    return RT_OK;
}

int rt_host_frame_open(const char *name, int64_t bytes, void **ptr) {
    if (!name || !ptr || bytes <= 0) return fail(RT_ERR_INVALID, "bad argument");
    const int fd = shm_open(name, O_RDWR, 0600);
    if (fd < 0) return fail(RT_ERR_STATE, std::string("shm_open failed for ") + name);
    void *p = mmap(nullptr, (size_t) bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (p == MAP_FAILED) return fail(RT_ERR_NOMEM, "mmap of the shared frame failed");
    if (cudaHostRegister(p, (size_t) bytes, cudaHostRegisterPortable | cudaHostRegisterMapped) != cudaSuccess) {
        cudaGetLastError();
        munmap(p, (size_t) bytes);
        return fail(RT_ERR_CUDA, "cudaHostRegister of the shared frame failed");
    }
    *ptr = p;
    return RT_OK;
}

int rt_host_frame_close(void *ptr, int64_t bytes, const char *unlink_name) {
    if (ptr) {
        cudaHostUnregister(ptr);
        cudaGetLastError();
        munmap(ptr, (size_t) bytes);
    }
    if (unlink_name) shm_unlink(unlink_name);
    return RT_OK;
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
// invariants of a host tree after the device-side re-layout and padding: every primitive in exactly one leaf and
// inside its leaf's box, every child box inside its parent's.  Returns the node count or a negative error.
static int check_host_bvh(HostBvh bvh, const std::vector<Aabb> &bounds) {
    compact_dfs(bvh, 7);  // a depth-first re-layout (here with a breadth-first prefix) must keep the tree intact
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
    size_t visited = 0;
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
        if (it.ref >= (int) bvh.nodes.size() || ++visited > bvh.nodes.size()) return fail(RT_ERR_STATE, "node reference out of bounds or a cycle");
        const HostNode &n = bvh.nodes[it.ref];
        for (int c = 0; c < 2; c++) {
            int ch = c ? n.child1 : n.child0;
            if (ch == kEmptyChild) continue;
            if (!inside(child_box(n, c), it.box)) return fail(RT_ERR_STATE, "child box outside its parent's box");
            st.push_back({ch, child_box(n, c)});
        }
    }
    for (int i = 0; i < np; i++)
        if (seen[i] != 1) return fail(RT_ERR_STATE, "primitive not in exactly one leaf");
    return (int) bvh.nodes.size();
}

int rt_host_check_bvh(const RtSceneDesc *desc, float *sah_cost, int32_t *max_depth) {
    int rc = validate(desc);
    if (rc != RT_OK) return rc;
    std::vector<Aabb> bounds;
    primitive_bounds(*desc, bounds);
    HostBvh bvh;
    build_bvh_sah_host(bounds, bvh);
    if (sah_cost) *sah_cost = bvh_sah_cost(bvh);
    if (max_depth) *max_depth = tree_depth(bvh);
    return check_host_bvh(bvh, bounds);
}

int rt_host_reinsert(const RtSceneDesc *desc, int rounds, float accept_ratio, float *cost2, int32_t *stats4) {
    int rc = validate(desc);
    if (rc != RT_OK) return rc;
    std::vector<Aabb> bounds;
    primitive_bounds(*desc, bounds);
    HostBvh bvh;
    build_bvh_sah_host_plain(bounds, bvh);
    ReinsertReport rep;
    reinsert_optimize_host(bvh, rounds, accept_ratio, &rep);
    if (cost2) cost2[0] = rep.cost_before, cost2[1] = rep.accepted ? bvh_sah_cost(bvh) : rep.cost_after;
    if (stats4) stats4[0] = rep.moves, stats4[1] = rep.rounds, stats4[2] = tree_depth(bvh), stats4[3] = rep.accepted ? 1 : 0;
    return check_host_bvh(bvh, bounds);
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

float rt_host_pow_ref(float base, float e) { return pow_ref(base, e); }
int rt_host_specular_gate(float cos_theta) { return specular_gate(cos_theta) ? 1 : 0; }

// ---- scene -------------------------------------------------------------------------------------------------------

int rt_scene_create(const RtSceneDesc *desc, const RtBuildOptions *opts, RtScene **out) {
    return scene_create_impl(desc, opts, out, nullptr);
}

// the GPU build of the reference-order tree and ranks, read back for the tests (they compare it with the host build
// of ref_order.cpp bit for bit); needs a device
int rt_device_reference_ranks(const RtSceneDesc *desc, uint32_t *ranks_out, int32_t *stats4, uint64_t *tree_hash) {
    RtScene *s = nullptr;
    SceneBuild build;
    RtBuildOptions o;
    memset(&o, 0, sizeof o);
    o.builder = RT_BUILD_SAH_GPU;
    int rc = scene_create_impl(desc, &o, &s, &build);
    if (rc != RT_OK) {
        build.release_scratch();
        return rc;
    }
    std::vector<uint32_t> ranks;
    RefTreeStats st;
    RefTree tree;
    std::string err;
    const int e = build.read_reference_tree(*desc, s->stream, s->buf, ranks, st, tree, err);
    build.release_scratch();
    if (e == 0) {
        if (ranks_out) memcpy(ranks_out, ranks.data(), ranks.size() * sizeof(uint32_t));
        if (stats4) stats4[0] = st.nodes, stats4[1] = st.leaves, stats4[2] = st.max_leaf, stats4[3] = st.max_depth;
        if (tree_hash) *tree_hash = ref_tree_hash(tree);
        // the statistics the build itself reported must agree with the tree that was read back
        if (s->info.ref_tree_nodes != st.nodes || s->info.ref_tree_leaves != st.leaves || s->info.ref_tree_max_leaf != st.max_leaf ||
            s->info.ref_tree_max_depth != st.max_depth) {
            rt_scene_destroy(s);
            return fail(RT_ERR_STATE, "device-side reference tree statistics disagree with the tree");
        }
    }
    rt_scene_destroy(s);
    return e == 0 ? RT_OK : fail(RT_ERR_CUDA, err);
}

void rt_scene_destroy(RtScene *s) {
    if (!s) return;
    DeviceGuard g;
    g.use(s->device);
    if (s->stream) cudaStreamSynchronize(s->stream);
    if (s->copy_stream) cudaStreamSynchronize(s->copy_stream);
    block_release(s->arena, s->arena_cap);  // the streams are drained: the block may be handed to the next scene
    cudaFree(s->d_parts);
    for (auto &sl: s->slot) {
        cudaFree(sl.d_frame);
        if (sl.h_pinned) cudaFreeHost(sl.h_pinned);
        if (sl.h_stats) cudaFreeHost(sl.h_stats);
        if (sl.ev_start) cudaEventDestroy(sl.ev_start);
        if (sl.ev_kernel) cudaEventDestroy(sl.ev_kernel);
        if (sl.ev_done) cudaEventDestroy(sl.ev_done);
    }
    for (auto &e: s->ev_part)
        if (e) cudaEventDestroy(e);
    if (s->stream) cudaStreamDestroy(s->stream);
    if (s->copy_stream) cudaStreamDestroy(s->copy_stream);
    cudaGetLastError();
    delete s;
}

int rt_scene_info(const RtScene *s, RtSceneInfo *info) {
    if (!s || !info) return fail(RT_ERR_INVALID, "NULL argument");
    *info = s->info;
    return RT_OK;
}

// ---- partition bookkeeping ------------------------------------------------------------------------------------

int rt_band_height(const RtCamera *cam, int aa, int part_world) {
    if (!cam || part_world < 1 || aa < 1) return -1;
    const ItemGeom g = item_geometry(cam, aa, part_world);
    return g.Ph * g.rpb;
}

// experiments: band height in pixel rows and strip width of the register-accumulator mode (0 = defaults).  Process-wide,
// because the partition has to stay a pure function of (camera, aa, world) that every rank evaluates identically.
int rt_set_partition(int band_rows, int strip_width) {
    g_band_rows = band_rows > 0 ? band_rows : 0;
    g_strip_width = strip_width > 0 ? strip_width : 0;
    return RT_OK;
}

int64_t rt_part_rows(const RtCamera *cam, int aa, int part_rank, int part_world) {
    if (!cam || part_world < 1 || part_rank < 0 || part_rank >= part_world || aa < 1) return -1;
    const ItemGeom g = item_geometry(cam, aa, part_world);
    return part_bands(g, part_rank, part_world) * g.Ph * g.rpb;  // the last band is padded to the full band height
}

int64_t rt_part_bytes(const RtCamera *cam, int aa, int part_rank, int part_world) {
    const int64_t rows = rt_part_rows(cam, aa, part_rank, part_world);
    return rows < 0 ? rows : rows * cam->image_width * 3;
}

// ---- rendering ------------------------------------------------------------------------------------------------

int rt_render_async(RtScene *s, const RtCamera *cam, int aa, unsigned char *rgb_out, int *ticket) {
    int rc = check_render_args(s, cam, aa, 0, 1);
    if (rc != RT_OK) return rc;
    if (!rgb_out || !ticket) return fail(RT_ERR_INVALID, "rgb_out or ticket is NULL");
    DeviceGuard g;
    if (g.use(s->device) != 0) return fail(RT_ERR_CUDA, "cudaSetDevice failed");
    int k = 0;  // the lowest free slot (a synchronous caller always reuses slot 0 and its buffers)
    while (k < kFrameSlots && s->slot[k].busy) k++;
    if (k == kFrameSlots) return fail(RT_ERR_STATE, "too many frames in flight on this handle (rt_wait the oldest first)");
    RenderSlot &sl = s->slot[k];
    rc = slot_prepare(sl);
    if (rc != RT_OK) return rc;
    const size_t bytes = (size_t) cam->image_width * cam->image_height * 3;
    // Large frame into a page-locked destination: the kernel stores its finished pixel runs straight into it over PCIe
    // (the memory is mapped into the device's address space), so the transfer rides along with the rendering instead of
    // following it.
    const bool pinned_dst = is_pinned(rgb_out);
    void *alias = nullptr;
    const bool zero_copy = pinned_dst && g_zero_copy_min >= 0 && (long long) bytes >= g_zero_copy_min &&
                           cudaHostGetDevicePointer(&alias, rgb_out, 0) == cudaSuccess && alias;
    cudaGetLastError();
    if (!zero_copy) {
        rc = ensure(&sl.d_frame, &sl.frame_cap, bytes, false);
        if (rc != RT_OK) return rc;
    }
    sl.launches = 0;
    CU(cudaEventRecord(sl.ev_start, s->stream));
    rc = enqueue_part(s, cam, aa, 0, 1, zero_copy ? (unsigned char *) alias : sl.d_frame, kOutFrame, k, s->stream, &sl.launches);
    if (rc != RT_OK) return rc;
    CU(cudaEventRecord(sl.ev_kernel, s->stream));
    // while the kernel runs: make sure there is page-locked memory to copy into
    if (!pinned_dst) {
        rc = ensure(&sl.h_pinned, &sl.pinned_cap, bytes, true);
        if (rc != RT_OK) return rc;
    }
    // the copies go to a second stream so that the next frame's kernel does not queue behind this frame's D2H
    CU(cudaStreamWaitEvent(s->copy_stream, sl.ev_kernel, 0));
    if (!zero_copy) CU(cudaMemcpyAsync(pinned_dst ? rgb_out : sl.h_pinned, sl.d_frame, bytes, cudaMemcpyDeviceToHost, s->copy_stream));
    CU(cudaMemcpyAsync(sl.h_stats, s->buf.control + 8 * k, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s->copy_stream));
    CU(cudaEventRecord(sl.ev_done, s->copy_stream));
    sl.busy = true;
    sl.dst = rgb_out;
    sl.bytes = bytes;
    sl.staged = !pinned_dst;
    *ticket = k;
    return RT_OK;
}

int rt_wait(RtScene *s, int ticket, RtStats *stats) {
    if (!s) return fail(RT_ERR_INVALID, "NULL scene");
    if (ticket < 0 || ticket >= kFrameSlots || !s->slot[ticket].busy) return fail(RT_ERR_STATE, "no frame in flight under this ticket");
    RenderSlot &sl = s->slot[ticket];
    DeviceGuard g;
    if (g.use(s->device) != 0) return fail(RT_ERR_CUDA, "cudaSetDevice failed");
    sl.busy = false;
    CU(cudaEventSynchronize(sl.ev_done));
    if (sl.staged) memcpy(sl.dst, sl.h_pinned, sl.bytes);
    if (stats) {
        memset(stats, 0, sizeof *stats);
        stats_from_words(sl.h_stats, stats);
        cudaEventElapsedTime(&stats->ms_render, sl.ev_start, sl.ev_kernel);
        cudaEventElapsedTime(&stats->ms_total, sl.ev_start, sl.ev_done);
        stats->ms_d2h = stats->ms_total - stats->ms_render;
        stats->n_launches = sl.launches;
    }
    return RT_OK;
}

int rt_render(RtScene *s, const RtCamera *cam, int aa, unsigned char *rgb_out, RtStats *stats) {
    int ticket = 0;
    int rc = rt_render_async(s, cam, aa, rgb_out, &ticket);
    if (rc != RT_OK) return rc;
    return rt_wait(s, ticket, stats);
}

int rt_render_part(RtScene *s, const RtCamera *cam, int aa, int rank, int world, void *d_rows, void *cuda_stream,
                   RtStats *stats) {
    return render_part_common(s, cam, aa, rank, world, d_rows, kOutPacked, cuda_stream, stats);
}

int rt_render_part_into_frame(RtScene *s, const RtCamera *cam, int aa, int rank, int world, void *d_frame,
                              void *cuda_stream, RtStats *stats) {
    return render_part_common(s, cam, aa, rank, world, d_frame, kOutFrame, cuda_stream, stats);
}

int rt_assemble_parts(const RtCamera *cam, int aa, int part_world, const void *d_parts, int64_t part_stride_bytes, void *d_frame,
                      void *cuda_stream) {
    if (!cam || !d_parts || !d_frame || part_world < 1 || aa < 1) return fail(RT_ERR_INVALID, "bad argument");
    const ItemGeom g = item_geometry(cam, aa, part_world);
    cudaError_t e = (cudaError_t) launch_assemble((const unsigned char *) d_parts, part_stride_bytes, part_world, cam->image_width,
                                                  cam->image_height, g.Ph * g.rpb, (unsigned char *) d_frame, (cudaStream_t) cuda_stream);
    if (e != cudaSuccess) return fail(RT_ERR_CUDA, std::string("assemble kernel launch: ") + cudaGetErrorString(e));
    return RT_OK;
}

int rt_render_part_to_host(RtScene *s, const RtCamera *cam, int aa, int rank, int world, unsigned char *host_frame, RtStats *stats) {
    int rc = check_render_args(s, cam, aa, rank, world);
    if (rc != RT_OK) return rc;
    if (!host_frame) return fail(RT_ERR_INVALID, "host_frame is NULL");
    DeviceGuard g;
    if (g.use(s->device) != 0) return fail(RT_ERR_CUDA, "cudaSetDevice failed");
    const double t0 = now_ms();
    int launches = 0;
    rc = enqueue_part_to_host(s, cam, aa, rank, world, host_frame, &launches);
    if (rc != RT_OK) return rc;
    unsigned long long h[8];
    CU(cudaMemcpyAsync(h, s->buf.control + 8 * (kControlSlots - 1), sizeof h, cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    if (stats) {
        memset(stats, 0, sizeof *stats);
        stats_from_words(h, stats);
        cudaEventElapsedTime(&stats->ms_render, s->ev_part[0], s->ev_part[1]);
        stats->ms_total = (float) (now_ms() - t0);
        stats->ms_d2h = stats->ms_total - stats->ms_render;
        stats->n_launches = launches;
    }
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
    DeviceGuard g;
    RtScene *root = scenes[0];
    const size_t bytes = (size_t) cam->image_width * cam->image_height * 3;
    const double t0 = now_ms();
    // every GPU copies its own bands over its own PCIe link straight into the caller's frame (page-locked), or into
    // the root handle's page-locked staging frame when the caller's memory is pageable
    unsigned char *frame = rgb_out;
    const bool pinned_dst = is_pinned(rgb_out);
    if (!pinned_dst) {
        if (g.use(root->device) != 0) return fail(RT_ERR_CUDA, "cudaSetDevice failed");
        if (root->slot[0].busy) return fail(RT_ERR_STATE, "rt_render_multi into pageable memory while an asynchronous frame is in flight");
        int rc = ensure(&root->slot[0].h_pinned, &root->slot[0].pinned_cap, bytes, true);
        if (rc != RT_OK) return rc;
        frame = root->slot[0].h_pinned;
    }
    int launches = 0;
    for (int i = 0; i < n; i++) {
        if (g.use(scenes[i]->device) != 0) return fail(RT_ERR_CUDA, "cudaSetDevice failed");
        int rc = enqueue_part_to_host(scenes[i], cam, aa, i, n, frame, &launches);
        if (rc != RT_OK) return rc;
    }
    float ms_render = 0;
    if (stats) memset(stats, 0, sizeof *stats);
    for (int i = 0; i < n; i++) {
        RtScene *s = scenes[i];
        if (g.use(s->device) != 0) return fail(RT_ERR_CUDA, "cudaSetDevice failed");
        unsigned long long h[8];
        CU(cudaMemcpyAsync(h, s->buf.control + 8 * (kControlSlots - 1), sizeof h, cudaMemcpyDeviceToHost, s->stream));
        CU(cudaStreamSynchronize(s->stream));
        if (stats) {
            RtStats part;
            memset(&part, 0, sizeof part);
            stats_from_words(h, &part);
            stats->primary_rays += part.primary_rays;
            stats->reflection_rays += part.reflection_rays;
            stats->shadow_rays += part.shadow_rays;
            stats->shadow_occluded += part.shadow_occluded;
            stats->replayed_closest += part.replayed_closest;
            stats->replayed_any += part.replayed_any;
            float ms = 0;
            cudaEventElapsedTime(&ms, s->ev_part[0], s->ev_part[1]);
            ms_render = std::max(ms_render, ms);
        }
    }
    if (!pinned_dst) memcpy(rgb_out, frame, bytes);
    if (stats) {
        stats->ms_render = ms_render;  // the slowest GPU's kernel
        stats->ms_total = (float) (now_ms() - t0);
        stats->ms_d2h = stats->ms_total - ms_render;
        stats->n_launches = launches;
    }
    return RT_OK;
}

}  // extern "C"
