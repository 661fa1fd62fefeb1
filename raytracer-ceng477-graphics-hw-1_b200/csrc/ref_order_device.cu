// ref_order_device.cu — the reference's own midpoint-split tree (bvh.h:48-163) and the 8 visit-rank arrays, built
// on the GPU.  Same result, bit for bit, as ref_order.cpp (checked by
// tests/test_gpu_parity.py::test_reference_tree_on_gpu_equals_host): level-synchronous, one CTA per open node —
// exact min/max bounds by shared-memory atomics, widest axis, the up-to-19 shrinking midpoint retries with a
// block-wide count each, a STABLE block-wide partition (the reference's push_back loops keep list order, and list
// order inside a leaf is visit order), then one thread per primitive walks from the root to its leaf and
// accumulates, for each of the 8 direction-sign octants, how many primitives the reference visits before it.
//
// No host round trip: ONE cooperative launch runs all levels — the number of open nodes of a level lives in device
// memory, persistent CTAs stride over however many tasks the previous level left, a grid barrier separates the levels —
// and the nodes are written straight into the replay layout (rt_internal.h `ref_nodes`).
#include <cfloat>
#include <cstdint>
#include <vector>

#include <cooperative_groups.h>
#include <cuda_runtime.h>

#include "build_device.h"

namespace cg = cooperative_groups;

namespace rtb {

namespace {

constexpr int kT = 256;

__device__ __forceinline__ int block_sum(int v, int *s) {  // inclusive scan, returns total; s[t] = inclusive prefix
    const int t = threadIdx.x;
    s[t] = v;
    __syncthreads();
    for (int off = 1; off < kT; off <<= 1) {
        const int add = t >= off ? s[t - off] : 0;
        __syncthreads();
        s[t] += add;
        __syncthreads();
    }
    return s[kT - 1];
}

__global__ void ref_init_kernel(int np, RefScratch s, BuildResult *res) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < np) s.ids[i] = i;
    if (i == 0) {
        for (int l = 0; l < kRefLevels + 2; l++) s.level_count[l] = l == 0 ? 1 : 0;
        *s.n_nodes = 1;  // node 0 = root
        s.queue[0][0] = RefTask{0, 0, np, 0};
        res->ref_leaves = 0;
        res->ref_max_leaf = 0;
        res->ref_max_depth = 0;
        res->ref_overflow = 0;
    }
}

// all levels in ONE cooperative launch (grid barrier between levels, the loop ends with the first empty level)
__global__ void __launch_bounds__(kT) ref_build_kernel(const Aabb *bounds, const float *key, int np, RefScratch s, float4 *ref_nodes,
                                                        BuildResult *res) {
    cg::grid_group grid = cg::this_grid();
  for (int level = 0; level < kRefLevels; level++) {
    if (__ldcg(&s.level_count[level]) == 0) break;
    __shared__ unsigned s_mn[3], s_mx[3];
    __shared__ int s_scan[kT];
    const int t = threadIdx.x;
    const int n_tasks = __ldcg(&s.level_count[level]);
    const RefTask *tasks = s.queue[level & 1];
    RefTask *next = s.queue[(level + 1) & 1];
    int *ids = s.ids, *tmp = s.tmp;
    for (int ti = blockIdx.x; ti < n_tasks; ti += gridDim.x) {
        const RefTask task = tasks[ti];
        const int lo = task.lo, hi = task.hi, n = hi - lo;
        __syncthreads();  // the previous task's shared state is no longer read
        if (t < 3) s_mn[t] = 0xffffffffu, s_mx[t] = 0u;
        __syncthreads();
        for (int i = lo + t; i < hi; i += kT) {  // parser.h:272-317: exact min / max over the primitives' bounds
            const Aabb b = bounds[ids[i]];
            for (int k = 0; k < 3; k++) {
                atomicMin(&s_mn[k], f2ord(b.mn[k]));
                atomicMax(&s_mx[k], f2ord(b.mx[k]));
            }
        }
        __syncthreads();
        float mn[3], mx[3];
        for (int k = 0; k < 3; k++) mn[k] = ord2f(s_mn[k]), mx[k] = ord2f(s_mx[k]);

        bool split = false;
        int axis = 0, n_left = 0;
        float mid = 0;
        if (n > 1 && task.depth < 19) {  // bvh.h:57, MAX_DEPTH bvh.h:18
            for (int a = 1; a < 3; a++)  // parser.h:227-235
                if (mx[a] - mn[a] > mx[axis] - mn[axis]) axis = a;
            const float *k = key + (size_t) axis * np;
            float start = mn[axis], end = mx[axis];
            mid = (start + end) / 2;
            for (int tries = 19; tries > 0 && !split; tries--) {  // bvh.h:117-145 (uniform across the CTA)
                int mine = 0;
                for (int i = lo + t; i < hi; i += kT) mine += k[ids[i]] < mid;
                n_left = block_sum(mine, s_scan);
                __syncthreads();
                if (n_left == 0) {
                    start = mid;
                    mid = (start + end) / 2;
                } else if (n_left == n) {
                    end = mid;
                    mid = (start + end) / 2;
                } else {
                    split = true;
                }
            }
        }
        int left = -1;
        if (split) {  // stable split, bvh.h:146-159: each thread owns a contiguous chunk
            const float *k = key + (size_t) axis * np;
            const int chunk = (n + kT - 1) / kT;
            const int b0 = min(hi, lo + t * chunk), b1 = min(hi, b0 + chunk);
            int mine = 0;
            for (int i = b0; i < b1; i++) mine += k[ids[i]] < mid;
            block_sum(mine, s_scan);
            int l = lo + (s_scan[t] - mine);
            int r = lo + n_left + ((b0 - lo) - (s_scan[t] - mine));
            __syncthreads();
            for (int i = b0; i < b1; i++) {
                const int id = ids[i];
                RT_CHECK(id >= 0 && id < np && l >= lo && l <= hi && r >= lo && r <= hi);
                if (k[id] < mid) tmp[l++] = id;
                else tmp[r++] = id;
            }
            __syncthreads();
            for (int i = lo + t; i < hi; i += kT) ids[i] = tmp[i];
            if (t == 0) {
                left = atomicAdd(s.n_nodes, 2);
                const int slot = atomicAdd(&s.level_count[level + 1], 2);
                RT_CHECK(slot >= 0 && slot + 1 <= np && left >= 1 && left + 1 <= 2 * np);
                next[slot] = RefTask{left, lo, lo + n_left, task.depth + 1};
                next[slot + 1] = RefTask{left + 1, lo + n_left, hi, task.depth + 1};
            }
        }
        if (t == 0) {
            DevRefNode nd;
            for (int k = 0; k < 3; k++) nd.mn[k] = mn[k], nd.mx[k] = mx[k];
            nd.axis = axis;
            nd.is_leaf = split ? 0 : 1;
            nd.left = left;
            nd.right = split ? left + 1 : -1;
            nd.first = lo;
            nd.count = n;
            nd.depth = task.depth;
            s.nodes[task.node] = nd;
            // replay layout (device_common.cuh ref_closest / ref_any): right child = left + 1
            ref_nodes[3 * (size_t) task.node + 0] = make_float4(mn[0], mn[1], mn[2], __int_as_float(axis | (split ? 0 : 4)));
            ref_nodes[3 * (size_t) task.node + 1] = make_float4(mx[0], mx[1], mx[2], __int_as_float(left));
            ref_nodes[3 * (size_t) task.node + 2] = make_float4(__int_as_float(lo), __int_as_float(n), 0.f, 0.f);
            atomicMax(&res->ref_max_depth, task.depth);
            if (!split) {
                atomicAdd(&res->ref_leaves, 1);
                atomicMax(&res->ref_max_leaf, n);
            }
        }
    }
    grid.sync();
  }
}

// raytracer.cpp:190-196: at an inner node the left child is visited first iff direction[axis] > 0
__global__ void ref_ranks_kernel(RefScratch s, int np, uint32_t *ranks, int *ref_leaf_prims, BuildResult *res) {
    const int pos = blockIdx.x * blockDim.x + threadIdx.x;
    if (pos == 0) res->ref_nodes = *s.n_nodes;
    if (pos >= np) return;
    const DevRefNode *nodes = s.nodes;
    uint32_t acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int node = 0;
    for (int guard = 0; guard < 64; guard++) {
        const DevRefNode nd = nodes[node];
        if (nd.is_leaf) {
            const int prim = s.ids[pos];
            ref_leaf_prims[pos] = prim;
            for (int o = 0; o < 8; o++) ranks[(size_t) o * np + prim] = acc[o] + (uint32_t) (pos - nd.first);
            return;
        }
        const int mid = nodes[nd.right].first;
        const bool in_left = pos < mid;
        const uint32_t other = in_left ? (uint32_t) (nd.first + nd.count - mid) : (uint32_t) (mid - nd.first);
        for (int o = 0; o < 8; o++) {
            const bool left_first = (o >> nd.axis) & 1;
            if (in_left != left_first) acc[o] += other;  // the other subtree is visited before this one
        }
        node = in_left ? nd.left : nd.right;
    }
    res->ref_overflow = 1;
}

}  // namespace

int ref_max_grid(int n_sms) {
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ref_build_kernel, kT, 0) != cudaSuccess || per_sm < 1) return 1;
    return n_sms * per_sm;
}

int enqueue_reference_tree(const Aabb *bounds, const float *key, int np, const RefScratch &s, uint32_t *ranks, float4 *ref_nodes,
                           int *ref_leaf_prims, BuildResult *result, int grid, cudaStream_t stream) {
    if (np <= 0) return 0;
    ref_init_kernel<<<(np + 255) / 256, 256, 0, stream>>>(np, s, result);
    long long want = np < grid ? np : grid;
    if (want < 1) want = 1;
    RefScratch sc = s;
    void *args[] = {&bounds, &key, &np, &sc, &ref_nodes, &result};
    const cudaError_t e = cudaLaunchCooperativeKernel((void *) ref_build_kernel, dim3((unsigned) want), dim3(kT), args, 0, stream);
    if (e != cudaSuccess) return (int) e;
    ref_ranks_kernel<<<(np + 255) / 256, 256, 0, stream>>>(s, np, ranks, ref_leaf_prims, result);
    return (int) cudaGetLastError();
}

}  // namespace rtb
