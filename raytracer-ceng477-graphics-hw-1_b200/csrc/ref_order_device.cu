// ref_order_device.cu — the reference's own midpoint-split tree (bvh.h:48-163) and the 8 visit-rank arrays, built
// on the GPU.  Same result, bit for bit, as ref_order.cpp (checked by
// tests/test_gpu_parity.py::test_reference_tree_on_gpu_equals_host): level-synchronous, one CTA per open node —
// exact min/max bounds by shared-memory atomics, widest axis, the up-to-19 shrinking midpoint retries with a
// block-wide count each, a STABLE block-wide partition (the reference's push_back loops keep list order, and list
// order inside a leaf is visit order), then one thread per primitive walks from the root to its leaf and
// accumulates, for each of the 8 direction-sign octants, how many primitives the reference visits before it.
#include <cfloat>
#include <cstdint>
#include <vector>

#include <cuda_runtime.h>

#include "rt_internal.h"

namespace rtb {

namespace {

constexpr int kT = 256;

struct RefTask {
    int node, lo, hi, depth;
};

struct DevRefNode {  // host-readable mirror of RefTreeNode plus the explicit left child
    float mn[3], mx[3];
    int axis, is_leaf, left, right, first, count, depth;
};

__device__ __forceinline__ unsigned ord(float f) {
    unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float unord(unsigned u) { return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u); }

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

__global__ void __launch_bounds__(kT) ref_level_kernel(const Aabb *bounds, const float *key, int np, int *ids, int *tmp,
                                                        const RefTask *tasks, RefTask *next, int *n_next, DevRefNode *nodes,
                                                        int *n_nodes) {
    __shared__ unsigned s_mn[3], s_mx[3];
    __shared__ int s_scan[kT];
    __shared__ float s_mid;
    __shared__ int s_split, s_axis, s_nleft;
    const int t = threadIdx.x;
    const RefTask task = tasks[blockIdx.x];
    const int lo = task.lo, hi = task.hi, n = hi - lo;
    if (t < 3) s_mn[t] = 0xffffffffu, s_mx[t] = 0u;
    __syncthreads();
    for (int i = lo + t; i < hi; i += kT) {  // parser.h:272-317: exact min / max over the primitives' bounds
        const Aabb b = bounds[ids[i]];
        for (int k = 0; k < 3; k++) {
            atomicMin(&s_mn[k], ord(b.mn[k]));
            atomicMax(&s_mx[k], ord(b.mx[k]));
        }
    }
    __syncthreads();
    float mn[3], mx[3];
    for (int k = 0; k < 3; k++) mn[k] = unord(s_mn[k]), mx[k] = unord(s_mx[k]);

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
    int left = -1, right = -1;
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
            if (k[id] < mid) tmp[l++] = id;
            else tmp[r++] = id;
        }
        __syncthreads();
        for (int i = lo + t; i < hi; i += kT) ids[i] = tmp[i];
        if (t == 0) {
            left = atomicAdd(n_nodes, 2);
            right = left + 1;
            const int slot = atomicAdd(n_next, 2);
            next[slot] = RefTask{left, lo, lo + n_left, task.depth + 1};
            next[slot + 1] = RefTask{right, lo + n_left, hi, task.depth + 1};
        }
    }
    if (t == 0) {
        DevRefNode nd;
        for (int k = 0; k < 3; k++) nd.mn[k] = mn[k], nd.mx[k] = mx[k];
        nd.axis = axis;
        nd.is_leaf = split ? 0 : 1;
        nd.left = left;
        nd.right = right;
        nd.first = lo;
        nd.count = n;
        nd.depth = task.depth;
        nodes[task.node] = nd;
    }
}

// raytracer.cpp:190-196: at an inner node the left child is visited first iff direction[axis] > 0
__global__ void ref_ranks_kernel(const DevRefNode *nodes, const int *ids, int np, uint32_t *ranks) {
    const int pos = blockIdx.x * blockDim.x + threadIdx.x;
    if (pos >= np) return;
    uint32_t acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int node = 0;
    for (;;) {
        const DevRefNode nd = nodes[node];
        if (nd.is_leaf) {
            const int prim = ids[pos];
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
}

}  // namespace

int build_reference_ranks_device(const RtSceneDesc &d, const std::vector<Aabb> &bounds, std::vector<uint32_t> &ranks,
                                 RefTreeStats &stats, RefTree &tree, float *ms_device) {
    const int np = (int) bounds.size(), nt = d.n_triangles;
    ranks.assign((size_t) 8 * np, 0u);
    stats = RefTreeStats();
    tree = RefTree();
    if (np == 0) return 0;
    // split keys exactly as ref_order.cpp computes them (raytracer.cpp:347 centroid; bvh.h:131 sphere centre)
    std::vector<float> key((size_t) 3 * np);
    for (int i = 0; i < nt; i++) {
        const RtTriangle &t = d.triangles[i];
        const RtVec3 &p = d.vertices[t.v0_id - 1], &q = d.vertices[t.v1_id - 1], &r = d.vertices[t.v2_id - 1];
        key[i] = ((p.x + q.x) + r.x) / 3;
        key[(size_t) np + i] = ((p.y + q.y) + r.y) / 3;
        key[(size_t) 2 * np + i] = ((p.z + q.z) + r.z) / 3;
    }
    for (int i = 0; i < d.n_spheres; i++) {
        const RtVec3 &c = d.vertices[d.spheres[i].center_vertex_id - 1];
        key[(size_t) nt + i] = c.x;
        key[(size_t) np + nt + i] = c.y;
        key[(size_t) 2 * np + nt + i] = c.z;
    }
    std::vector<int> ids((size_t) np);
    for (int i = 0; i < np; i++) ids[i] = i;

    Aabb *d_bounds = nullptr;
    float *d_key = nullptr;
    int *d_ids = nullptr, *d_tmp = nullptr, *d_counters = nullptr;
    RefTask *d_q[2] = {nullptr, nullptr};
    DevRefNode *d_nodes = nullptr;
    uint32_t *d_ranks = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    auto cleanup = [&]() {
        cudaFree(d_bounds); cudaFree(d_key); cudaFree(d_ids); cudaFree(d_tmp); cudaFree(d_counters); cudaFree(d_q[0]);
        cudaFree(d_q[1]); cudaFree(d_nodes); cudaFree(d_ranks);
        if (e0) cudaEventDestroy(e0);
        if (e1) cudaEventDestroy(e1);
    };
#define CKR(call)                    \
    do {                             \
        if ((call) != cudaSuccess) { \
            cleanup();               \
            return -1;               \
        }                            \
    } while (0)
    const size_t max_nodes = (size_t) 2 * np + 1;
    CKR(cudaMalloc(&d_bounds, sizeof(Aabb) * np));
    CKR(cudaMalloc(&d_key, sizeof(float) * 3 * np));
    CKR(cudaMalloc(&d_ids, sizeof(int) * np));
    CKR(cudaMalloc(&d_tmp, sizeof(int) * np));
    CKR(cudaMalloc(&d_counters, sizeof(int) * 2));
    CKR(cudaMalloc(&d_q[0], sizeof(RefTask) * (size_t) (np + 1)));
    CKR(cudaMalloc(&d_q[1], sizeof(RefTask) * (size_t) (np + 1)));
    CKR(cudaMalloc(&d_nodes, sizeof(DevRefNode) * max_nodes));
    CKR(cudaMalloc(&d_ranks, sizeof(uint32_t) * 8 * np));
    CKR(cudaEventCreate(&e0));
    CKR(cudaEventCreate(&e1));
    CKR(cudaMemcpy(d_bounds, bounds.data(), sizeof(Aabb) * np, cudaMemcpyHostToDevice));
    CKR(cudaMemcpy(d_key, key.data(), sizeof(float) * 3 * np, cudaMemcpyHostToDevice));
    CKR(cudaMemcpy(d_ids, ids.data(), sizeof(int) * np, cudaMemcpyHostToDevice));
    const int counters0[2] = {0, 1};  // n_next, n_nodes (node 0 = root)
    CKR(cudaMemcpy(d_counters, counters0, sizeof counters0, cudaMemcpyHostToDevice));
    const RefTask root = {0, 0, np, 0};
    CKR(cudaMemcpy(d_q[0], &root, sizeof root, cudaMemcpyHostToDevice));

    CKR(cudaEventRecord(e0));
    int n_tasks = 1, cur = 0;
    for (int level = 0; n_tasks > 0 && level <= 20; level++) {
        CKR(cudaMemsetAsync(d_counters, 0, sizeof(int)));
        ref_level_kernel<<<n_tasks, kT>>>(d_bounds, d_key, np, d_ids, d_tmp, d_q[cur], d_q[cur ^ 1], d_counters, d_nodes, d_counters + 1);
        CKR(cudaMemcpy(&n_tasks, d_counters, sizeof(int), cudaMemcpyDeviceToHost));
        cur ^= 1;
    }
    ref_ranks_kernel<<<(np + 255) / 256, 256>>>(d_nodes, d_ids, np, d_ranks);
    CKR(cudaEventRecord(e1));
    CKR(cudaEventSynchronize(e1));
    CKR(cudaGetLastError());
    if (ms_device) cudaEventElapsedTime(ms_device, e0, e1);

    int counters[2];
    CKR(cudaMemcpy(counters, d_counters, sizeof counters, cudaMemcpyDeviceToHost));
    const int n_nodes = counters[1];
    std::vector<DevRefNode> dn((size_t) n_nodes);
    CKR(cudaMemcpy(dn.data(), d_nodes, sizeof(DevRefNode) * n_nodes, cudaMemcpyDeviceToHost));
    CKR(cudaMemcpy(ranks.data(), d_ranks, sizeof(uint32_t) * 8 * np, cudaMemcpyDeviceToHost));
    tree.leaf_prims.resize((size_t) np);
    CKR(cudaMemcpy(tree.leaf_prims.data(), d_ids, sizeof(int) * np, cudaMemcpyDeviceToHost));
    cleanup();
#undef CKR

    // relabel into the reference's pre-order (left child = index + 1, bvh.h:81-105), which the replay kernels use
    std::vector<int> order;
    order.reserve((size_t) n_nodes);
    std::vector<int> new_index((size_t) n_nodes, -1), todo(1, 0);
    while (!todo.empty()) {
        const int o = todo.back();
        todo.pop_back();
        new_index[o] = (int) order.size();
        order.push_back(o);
        if (!dn[o].is_leaf) {
            todo.push_back(dn[o].right);
            todo.push_back(dn[o].left);
        }
    }
    tree.nodes.resize(order.size());
    tree.leaf_of_prim.assign((size_t) np, 0);
    for (size_t i = 0; i < order.size(); i++) {
        const DevRefNode &n = dn[order[i]];
        RefTreeNode &o = tree.nodes[i];
        for (int a = 0; a < 3; a++) o.mn[a] = n.mn[a], o.mx[a] = n.mx[a];
        o.axis = n.axis;
        o.is_leaf = n.is_leaf;
        o.right = n.is_leaf ? -1 : new_index[n.right];
        o.first = n.first;
        o.count = n.count;
        stats.nodes++;
        if (n.depth > stats.max_depth) stats.max_depth = n.depth;
        if (n.is_leaf) {
            stats.leaves++;
            if (n.count > stats.max_leaf) stats.max_leaf = n.count;
            for (int k = 0; k < n.count; k++) tree.leaf_of_prim[tree.leaf_prims[n.first + k]] = (int) i;
        }
    }
    return 0;
}

}  // namespace rtb
