// bvh_sah_device.cu — top-down binned-SAH BVH2 build on the GPU (RT_BUILD_SAH_GPU).
//
// The same algorithm as the host yardstick (bvh_host.cpp: 32 centroid bins per axis, SAH sweep, leaves of at most 8
// primitives when cheaper, median split when the SAH finds nothing), level-synchronous: one kernel launch per tree
// level, persistent CTAs striding over the open ranges of the level (their number lives in device memory: no host
// round trip between levels).  A CTA reduces the range's bounds and centroid bounds, bins the primitives with
// shared-memory atomics (order-independent: counts and min/max are exact), one thread sweeps the 3 x 31 split
// candidates in the host's order with the host's float operations (the same decisions, hence the same tree up to
// the order of primitives inside a leaf), then the CTA partitions the range with a block-wide prefix sum, reduces
// the two child boxes and opens the child ranges for the next level.  Together with the PLOC builder this lets
// RT_BUILD_AUTO choose between two GPU-built trees.
#include <cfloat>
#include <cstdint>

#include <cooperative_groups.h>
#include <cuda_runtime.h>

#include "build_device.h"

namespace cg = cooperative_groups;

namespace rtb {

namespace {

constexpr int kBins = 32;
constexpr float kCostNode = kSahCostNode;
constexpr float kCostPrim = kSahCostPrim;
constexpr int kT = 256;

__device__ __forceinline__ unsigned ord(float f) {
    unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float unord(unsigned u) { return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u); }

struct Box6 {
    float mn[3], mx[3];
};
__device__ __forceinline__ Box6 empty6() {
    Box6 b;
    for (int k = 0; k < 3; k++) b.mn[k] = FLT_MAX, b.mx[k] = -FLT_MAX;
    return b;
}
__device__ __forceinline__ void grow6(Box6 &a, const Box6 &b) {
    for (int k = 0; k < 3; k++) a.mn[k] = fminf(a.mn[k], b.mn[k]), a.mx[k] = fmaxf(a.mx[k], b.mx[k]);
}
__device__ __forceinline__ float half_area6(const Box6 &b) {  // bvh_host.cpp half_area
    const float dx = b.mx[0] - b.mn[0], dy = b.mx[1] - b.mn[1], dz = b.mx[2] - b.mn[2];
    if (dx < 0 || dy < 0 || dz < 0) return 0;
    return dx * dy + dy * dz + dz * dx;
}

// shared-memory min/max box made of ordered uints
struct SBox {
    unsigned mn[3], mx[3];
};
__device__ __forceinline__ void sbox_reset(SBox &b) {
    for (int k = 0; k < 3; k++) b.mn[k] = 0xffffffffu, b.mx[k] = 0u;
}
__device__ __forceinline__ void sbox_add(SBox &b, const float *mn, const float *mx) {
    for (int k = 0; k < 3; k++) {
        atomicMin(&b.mn[k], ord(mn[k]));
        atomicMax(&b.mx[k], ord(mx[k]));
    }
}
__device__ __forceinline__ Box6 sbox_get(const SBox &b) {
    Box6 r;
    for (int k = 0; k < 3; k++) {
        r.mn[k] = b.mn[k] == 0xffffffffu ? FLT_MAX : unord(b.mn[k]);
        r.mx[k] = b.mx[k] == 0u ? -FLT_MAX : unord(b.mx[k]);
    }
    return r;
}

__device__ __forceinline__ int bin_of(const Aabb &pb, int axis, float c0, float scale) {
    return min(kBins - 1, max(0, (int) ((0.5f * (pb.mn[axis] + pb.mx[axis]) - c0) * scale)));
}

__global__ void sah_init_kernel(int n, SahScratch s) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) s.ids[i] = i;
    if (i == 0) {
        for (int l = 0; l < kSahLevels + 2; l++) s.level_count[l] = 0;
        s.level_count[0] = n > 1 ? 1 : 0;
        s.queue[0][0] = SahTask{-1, 0, 0, n};
        *s.n_nodes = 0;
        *s.root_ref = n == 1 ? ~0 : 0;
    }
}

// All levels in ONE cooperative launch: persistent CTAs stride over the open ranges of a level, a grid barrier separates
// the levels, the loop ends when a level leaves no open range (round 1 launched one kernel per level and read the
// range count back to the host in between).
__global__ void __launch_bounds__(kT) sah_build_kernel(const Aabb *bounds, SahScratch sc, HostNode *nodes) {
    cg::grid_group grid = cg::this_grid();
  for (int level = 0; level < kSahLevels; level++) {
    if (__ldcg(&sc.level_count[level]) == 0) break;  // the same value in every CTA: written before the last grid barrier
    __shared__ SBox s_box, s_cbox, s_left, s_right;
    __shared__ SBox s_bins[3][kBins];
    __shared__ int s_cnt[3][kBins];
    __shared__ int s_scan[kT];
    __shared__ int s_axis, s_bin, s_leaf, s_node;
    __shared__ float s_c0, s_scale;

    const int t = threadIdx.x;
    const int n_tasks = __ldcg(&sc.level_count[level]);
    const SahTask *tasks = sc.queue[level & 1];
    SahTask *next = sc.queue[(level + 1) & 1];
    int *n_next = &sc.level_count[level + 1];
    int *ids = sc.ids, *tmp = sc.tmp, *n_nodes = sc.n_nodes, *root_ref = sc.root_ref;
  for (int ti = blockIdx.x; ti < n_tasks; ti += gridDim.x) {
    const SahTask task = tasks[ti];
    const int lo = task.lo, hi = task.hi, n = hi - lo;
    __syncthreads();  // the previous range's shared state is no longer read

    if (t == 0) {
        sbox_reset(s_box);
        sbox_reset(s_cbox);
        sbox_reset(s_left);
        sbox_reset(s_right);
    }
    for (int i = t; i < 3 * kBins; i += kT) {
        sbox_reset(s_bins[i / kBins][i % kBins]);
        s_cnt[i / kBins][i % kBins] = 0;
    }
    __syncthreads();
    // 1. bounds and centroid bounds of the range
    for (int i = lo + t; i < hi; i += kT) {
        const Aabb b = bounds[ids[i]];
        sbox_add(s_box, b.mn, b.mx);
        float c[3];
        for (int k = 0; k < 3; k++) c[k] = 0.5f * (b.mn[k] + b.mx[k]);
        sbox_add(s_cbox, c, c);
    }
    __syncthreads();
    const Box6 box = sbox_get(s_box), cbox = sbox_get(s_cbox);
    // 2. binning, all three axes at once
    for (int i = lo + t; i < hi; i += kT) {
        const Aabb b = bounds[ids[i]];
        for (int axis = 0; axis < 3; axis++) {
            const float c0 = cbox.mn[axis], c1 = cbox.mx[axis];
            if (!(c1 > c0)) continue;
            const int bin = bin_of(b, axis, c0, kBins / (c1 - c0));
            atomicAdd(&s_cnt[axis][bin], 1);
            sbox_add(s_bins[axis][bin], b.mn, b.mx);
        }
    }
    __syncthreads();
    // 3. SAH sweep in the host's order (bvh_host.cpp SahBuilder::build)
    if (t == 0) {
        int best_axis = -1, best_bin = -1;
        float best_cost = FLT_MAX;
        const float parent_area = half_area6(box);
        for (int axis = 0; axis < 3; axis++) {
            const float c0 = cbox.mn[axis], c1 = cbox.mx[axis];
            if (!(c1 > c0)) continue;
            float right_area[kBins];
            Box6 acc = empty6();
            for (int b = kBins - 1; b > 0; b--) {
                grow6(acc, sbox_get(s_bins[axis][b]));
                right_area[b] = half_area6(acc);
            }
            acc = empty6();
            int nl = 0;
            for (int b = 0; b < kBins - 1; b++) {
                grow6(acc, sbox_get(s_bins[axis][b]));
                nl += s_cnt[axis][b];
                if (nl == 0 || nl == n) continue;
                const float cost = half_area6(acc) * nl + right_area[b + 1] * (n - nl);
                if (cost < best_cost) best_cost = cost, best_axis = axis, best_bin = b;
            }
        }
        const float leaf_cost = kCostPrim * n;
        const float split_cost = best_axis < 0 ? FLT_MAX : kCostNode + kCostPrim * best_cost / (parent_area > 0 ? parent_area : 1e-30f);
        const int leaf = n <= kMaxLeafPrims && (best_axis < 0 || leaf_cost <= split_cost);
        s_leaf = leaf;
        s_axis = best_axis;
        s_bin = best_bin;
        if (best_axis >= 0) {
            s_c0 = cbox.mn[best_axis];
            s_scale = kBins / (cbox.mx[best_axis] - cbox.mn[best_axis]);
        }
        const int ref = leaf ? ~((lo << 3) | (n - 1)) : atomicAdd(n_nodes, 1);
        s_node = ref;
        if (task.parent < 0) *root_ref = ref;
        else if (task.side == 0) nodes[task.parent].child0 = ref;
        else nodes[task.parent].child1 = ref;
    }
    __syncthreads();
    if (s_leaf) continue;
    const int axis = s_axis, me = s_node;

    // 4. partition [lo, hi): left = bin <= best bin (order inside the halves is irrelevant to the tree)
    int mid;
    if (axis >= 0) {
        const float c0 = s_c0, scale = s_scale;
        const int best_bin = s_bin;
        int my_left = 0;
        for (int i = lo + t; i < hi; i += kT) my_left += bin_of(bounds[ids[i]], axis, c0, scale) <= best_bin;
        s_scan[t] = my_left;
        __syncthreads();
        for (int off = 1; off < kT; off <<= 1) {
            const int add = t >= off ? s_scan[t - off] : 0;
            __syncthreads();
            s_scan[t] += add;
            __syncthreads();
        }
        const int n_left = s_scan[kT - 1];
        // second pass in the same thread-strided order: thread t owns left slots [excl, excl + my_left)
        int l = lo + (s_scan[t] - my_left);
        // right slots: count of rights before this thread
        int my_cnt = 0;
        for (int i = lo + t; i < hi; i += kT) my_cnt++;
        __syncthreads();
        s_scan[t] = my_cnt - my_left;
        __syncthreads();
        for (int off = 1; off < kT; off <<= 1) {
            const int add = t >= off ? s_scan[t - off] : 0;
            __syncthreads();
            s_scan[t] += add;
            __syncthreads();
        }
        int r = lo + n_left + (s_scan[t] - (my_cnt - my_left));
        for (int i = lo + t; i < hi; i += kT) {
            const int id = ids[i];
            RT_CHECK(id >= 0 && l >= lo && l <= hi && r >= lo && r <= hi);
            if (bin_of(bounds[id], axis, c0, scale) <= best_bin) tmp[l++] = id;
            else tmp[r++] = id;
        }
        __syncthreads();
        for (int i = lo + t; i < hi; i += kT) ids[i] = tmp[i];
        mid = lo + n_left;
    } else {
        mid = lo + n / 2;  // all centroids coincide: split the list
    }
    if (mid == lo || mid == hi) mid = lo + n / 2;
    __syncthreads();

    // 5. child boxes, child ranges
    for (int i = lo + t; i < hi; i += kT) {
        const Aabb b = bounds[ids[i]];
        sbox_add(i < mid ? s_left : s_right, b.mn, b.mx);
    }
    __syncthreads();
    if (t == 0) {
        const Box6 b0 = sbox_get(s_left), b1 = sbox_get(s_right);
        HostNode &nd = nodes[me];
        for (int k = 0; k < 3; k++) {
            nd.c0mn[k] = b0.mn[k], nd.c0mx[k] = b0.mx[k];
            nd.c1mn[k] = b1.mn[k], nd.c1mx[k] = b1.mx[k];
        }
        const int los[2] = {lo, mid}, his[2] = {mid, hi};
        for (int c = 0; c < 2; c++) {
            if (his[c] - los[c] == 1) {  // a single primitive is a leaf without a task (bvh_host.cpp: n == 1)
                const int ref = ~((los[c] << 3) | 0);
                if (c == 0) nd.child0 = ref;
                else nd.child1 = ref;
            } else {
                const int slot = atomicAdd(n_next, 1);
                RT_CHECK(slot >= 0 && slot <= sc.cap_tasks && me >= 0 && me < sc.cap_nodes);
                next[slot] = SahTask{me, c, los[c], his[c]};
            }
        }
    }
  }
    grid.sync();
  }
}

// the whole scene is one leaf (or one primitive): give it a parent so that node 0 always exists
__global__ void sah_fixup_kernel(SahScratch s, HostNode *nodes, const BuildResult *res, int *status) {
    const int root_ref = *s.root_ref;
    if (root_ref < 0) {
        HostNode nd;
        for (int k = 0; k < 3; k++) {
            nd.c0mn[k] = ord2f(res->scene_bounds[k]), nd.c0mx[k] = ord2f(res->scene_bounds[3 + k]);
            nd.c1mn[k] = FLT_MAX, nd.c1mx[k] = -FLT_MAX;
        }
        nd.child0 = root_ref;
        nd.child1 = kEmptyChild;
        nodes[0] = nd;
        *s.n_nodes = 1;
    }
    // still open ranges after the last level: the tree would be deeper than the traversal stack anyway
    *status = s.level_count[kSahLevels] != 0 ? 1 : 0;
}

}  // namespace

int sah_max_grid(int n_sms) {
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, sah_build_kernel, kT, 0) != cudaSuccess || per_sm < 1) return 1;
    return n_sms * per_sm;
}

int enqueue_sah(const Aabb *bounds, int n, const SahScratch &s, DevTree &out, const BuildResult *res, int grid, cudaStream_t stream) {
    sah_init_kernel<<<(n + 255) / 256, 256, 0, stream>>>(n, s);
    long long want = n < grid ? n : grid;  // never more CTAs than primitives (a level has at most n / 2 open ranges)
    if (want < 1) want = 1;
    SahScratch sc = s;
    HostNode *nodes = out.nodes;
    void *args[] = {&bounds, &sc, &nodes};
    const cudaError_t e = cudaLaunchCooperativeKernel((void *) sah_build_kernel, dim3((unsigned) want), dim3(kT), args, 0, stream);
    if (e != cudaSuccess) return (int) e;
    sah_fixup_kernel<<<1, 1, 0, stream>>>(s, out.nodes, res, out.status);
    return (int) cudaGetLastError();
}

}  // namespace rtb
