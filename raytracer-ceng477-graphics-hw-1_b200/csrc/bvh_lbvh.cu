// bvh_lbvh.cu — BVH build on the GPU (replaces BVHNode::build / BVHTree::build, bvh.h:48-178).
//
//   1. centroid bounds            (block reduction + ordered-int atomics)
//   2. 30-bit Morton code per primitive, key = code << 32 | index (unique keys)
//   3. bitonic sort of the 64-bit keys (shared-memory stages fused, global stages one launch each)
//   4. Karras 2012 radix tree: one thread per internal node finds its range and split by binary
//      search on the common-prefix length of the sorted keys
//   5. bottom-up refit: one thread per leaf climbs; the second arrival at a node (atomic flag)
//      merges the children's boxes — and evaluates the SAH: a subtree whose primitives are cheaper
//      to test as one leaf (<= 8 contiguous primitives in Morton order) is collapsed ("SAH refinement")
//   6. emit 64-byte nodes (both children's boxes in the parent), leaf ranges into the sorted order
//
// The tree only has to be conservative: exact-t ties are settled by the
// reference-order ranks, so nothing here needs to mimic the reference's midpoint splits.
#include <cfloat>
#include <cstdint>
#include <cstdlib>
#include <vector>

#include <cuda_runtime.h>

#include "rt_internal.h"

namespace rtb {

namespace {

constexpr float kCostNode = 1.0f;  // keep in sync with bvh_host.cpp
constexpr float kCostPrim = 1.6f;

__device__ __forceinline__ unsigned ordered(float f) {
    unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float unordered(unsigned u) {
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

// [0..2] = min (ordered uint), [3..5] = max
__global__ void centroid_bounds_kernel(const Aabb *bounds, int n, unsigned *out) {
    __shared__ unsigned smin[3][256], smax[3][256];
    unsigned mn[3] = {0xffffffffu, 0xffffffffu, 0xffffffffu}, mx[3] = {0u, 0u, 0u};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const Aabb b = bounds[i];
        for (int k = 0; k < 3; k++) {
            unsigned c = ordered(0.5f * (b.mn[k] + b.mx[k]));
            mn[k] = min(mn[k], c);
            mx[k] = max(mx[k], c);
        }
    }
    for (int k = 0; k < 3; k++) smin[k][threadIdx.x] = mn[k], smax[k][threadIdx.x] = mx[k];
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s)
            for (int k = 0; k < 3; k++) {
                smin[k][threadIdx.x] = min(smin[k][threadIdx.x], smin[k][threadIdx.x + s]);
                smax[k][threadIdx.x] = max(smax[k][threadIdx.x], smax[k][threadIdx.x + s]);
            }
        __syncthreads();
    }
    if (threadIdx.x == 0)
        for (int k = 0; k < 3; k++) {
            atomicMin(&out[k], smin[k][0]);
            atomicMax(&out[3 + k], smax[k][0]);
        }
}

__device__ __forceinline__ unsigned expand10(unsigned v) {  // 10 bits -> every third bit
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}

__global__ void morton_kernel(const Aabb *bounds, int n, int n_pad, const unsigned *cb, unsigned long long *keys) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pad) return;
    if (i >= n) {
        keys[i] = ~0ull;
        return;
    }
    const Aabb b = bounds[i];
    unsigned code = 0;
    for (int k = 0; k < 3; k++) {
        const float lo = unordered(cb[k]), hi = unordered(cb[3 + k]);
        const float c = 0.5f * (b.mn[k] + b.mx[k]);
        float x = hi > lo ? (c - lo) / (hi - lo) : 0.0f;
        x = fminf(fmaxf(x * 1024.0f, 0.0f), 1023.0f);
        code |= expand10((unsigned) x) << (2 - k);
    }
    keys[i] = ((unsigned long long) code << 32) | (unsigned) i;
}

// ---- bitonic sort (ascending), n_pad a power of two ------------------------------------------
constexpr int kSortTile = 2048;  // keys per CTA in the shared-memory stages (1024 threads)

__device__ __forceinline__ void cmpswap(unsigned long long &a, unsigned long long &b, bool up) {
    if ((a > b) == up) {
        unsigned long long t = a;
        a = b;
        b = t;
    }
}

// all stages with k <= kSortTile, entirely in shared memory
__global__ void bitonic_local_kernel(unsigned long long *keys, int n_pad) {
    __shared__ unsigned long long s[kSortTile];
    const int base = blockIdx.x * kSortTile;
    for (int i = threadIdx.x; i < kSortTile; i += blockDim.x) s[i] = (base + i < n_pad) ? keys[base + i] : ~0ull;
    __syncthreads();
    for (int k = 2; k <= kSortTile; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = threadIdx.x; t < kSortTile / 2; t += blockDim.x) {
                const int i = 2 * t - (t & (j - 1));  // index with bit j clear
                const bool up = (((base + i) & k) == 0);
                cmpswap(s[i], s[i + j], up);
            }
            __syncthreads();
        }
    for (int i = threadIdx.x; i < kSortTile; i += blockDim.x)
        if (base + i < n_pad) keys[base + i] = s[i];
}

// one global stage (j >= kSortTile)
__global__ void bitonic_global_kernel(unsigned long long *keys, int n_pad, int j, int k) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_pad / 2) return;
    const int i = 2 * t - (t & (j - 1));
    const bool up = ((i & k) == 0);
    unsigned long long a = keys[i], b = keys[i + j];
    if ((a > b) == up) {
        keys[i] = b;
        keys[i + j] = a;
    }
}

// the stages j < kSortTile of a merge step k > kSortTile, in shared memory
__global__ void bitonic_merge_local_kernel(unsigned long long *keys, int n_pad, int k) {
    __shared__ unsigned long long s[kSortTile];
    const int base = blockIdx.x * kSortTile;
    for (int i = threadIdx.x; i < kSortTile; i += blockDim.x) s[i] = keys[base + i];
    __syncthreads();
    for (int j = kSortTile >> 1; j > 0; j >>= 1) {
        for (int t = threadIdx.x; t < kSortTile / 2; t += blockDim.x) {
            const int i = 2 * t - (t & (j - 1));
            const bool up = (((base + i) & k) == 0);
            cmpswap(s[i], s[i + j], up);
        }
        __syncthreads();
    }
    for (int i = threadIdx.x; i < kSortTile; i += blockDim.x) keys[base + i] = s[i];
}

// ---- Karras radix tree -------------------------------------------------------------------------
struct TreeNode {
    int left, right;    // >= 0 internal index, < 0: ~leaf position
    int parent;
    int first, last;    // leaf range covered
};

__device__ __forceinline__ int delta(const unsigned long long *keys, int n, int i, int j) {
    if (j < 0 || j >= n) return -1;
    return __clzll(keys[i] ^ keys[j]);
}

__global__ void radix_tree_kernel(const unsigned long long *keys, int n, TreeNode *nodes, int *leaf_parent) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    const int d = (delta(keys, n, i, i + 1) - delta(keys, n, i, i - 1)) >= 0 ? 1 : -1;
    const int dmin = delta(keys, n, i, i - d);
    int lmax = 2;
    while (delta(keys, n, i, i + lmax * d) > dmin) lmax <<= 1;
    int l = 0;
    for (int t = lmax >> 1; t >= 1; t >>= 1)
        if (delta(keys, n, i, i + (l + t) * d) > dmin) l += t;
    const int j = i + l * d;
    const int dnode = delta(keys, n, i, j);
    int s = 0;
    for (int t = (l + 1) >> 1;; t = (t + 1) >> 1) {
        if (delta(keys, n, i, i + (s + t) * d) > dnode) s += t;
        if (t == 1) break;
    }
    const int gamma = i + s * d + min(d, 0);
    const int lo = min(i, j), hi = max(i, j);
    TreeNode nd;
    nd.first = lo;
    nd.last = hi;
    nd.parent = (i == 0) ? -1 : nodes[i].parent;  // parent is written by the parent's thread (below); keep what is there
    if (lo == gamma) {
        nd.left = ~gamma;
        leaf_parent[gamma] = i;
    } else {
        nd.left = gamma;
    }
    if (hi == gamma + 1) {
        nd.right = ~(gamma + 1);
        leaf_parent[gamma + 1] = i;
    } else {
        nd.right = gamma + 1;
    }
    nodes[i].left = nd.left;
    nodes[i].right = nd.right;
    nodes[i].first = nd.first;
    nodes[i].last = nd.last;
    if (nd.left >= 0) nodes[nd.left].parent = i;
    if (nd.right >= 0) nodes[nd.right].parent = i;
    if (i == 0) nodes[0].parent = -1;
}

__device__ __forceinline__ float half_area(const Aabb &b) {
    const float dx = b.mx[0] - b.mn[0], dy = b.mx[1] - b.mn[1], dz = b.mx[2] - b.mn[2];
    return dx * dy + dy * dz + dz * dx;
}
__device__ __forceinline__ Aabb merge(const Aabb &a, const Aabb &b) {
    Aabb r;
    for (int k = 0; k < 3; k++) r.mn[k] = fminf(a.mn[k], b.mn[k]), r.mx[k] = fmaxf(a.mx[k], b.mx[k]);
    return r;
}

// One thread per leaf climbs towards the root; the second thread to arrive at a node owns it.
// cost[] holds the SAH cost of the (possibly collapsed) subtree, collapsed[] marks subtrees turned into leaves.
__global__ void refit_kernel(const unsigned long long *keys, const Aabb *bounds, int n, const TreeNode *nodes,
                             const int *leaf_parent, Aabb *node_box, float *cost, int *collapsed, unsigned *flags,
                             int do_collapse) {
    const int leaf = blockIdx.x * blockDim.x + threadIdx.x;
    if (leaf >= n) return;
    int cur = leaf_parent[leaf];
    while (cur >= 0) {
        __threadfence();
        if (atomicAdd(&flags[cur], 1u) == 0u) return;  // first arrival: the sibling subtree is not finished yet
        __threadfence();
        const TreeNode nd = nodes[cur];
        Aabb bl, br;
        float cl, cr;
        // children finished by other SMs: read through L2 (L1 is not coherent)
        auto load_box = [&](int idx) {
            Aabb b;
            const float *src = (const float *) &node_box[idx];
            for (int k = 0; k < 3; k++) b.mn[k] = __ldcg(src + k), b.mx[k] = __ldcg(src + 3 + k);
            return b;
        };
        if (nd.left < 0) {
            bl = bounds[(unsigned) keys[~nd.left]];
            cl = kCostPrim;
        } else {
            bl = load_box(nd.left);
            cl = __ldcg(&cost[nd.left]);
        }
        if (nd.right < 0) {
            br = bounds[(unsigned) keys[~nd.right]];
            cr = kCostPrim;
        } else {
            br = load_box(nd.right);
            cr = __ldcg(&cost[nd.right]);
        }
        const Aabb box = merge(bl, br);
        const float a = half_area(box);
        const int count = nd.last - nd.first + 1;
        float c_split = kCostNode + (a > 0.0f ? (half_area(bl) * cl + half_area(br) * cr) / a : cl + cr);
        const float c_leaf = kCostPrim * (float) count;
        int col = 0;
        if (do_collapse && cur != 0 && count <= kMaxLeafPrims && c_leaf <= c_split) {
            col = 1;
            c_split = c_leaf;
        }
        node_box[cur] = box;
        cost[cur] = c_split;
        collapsed[cur] = col;
        cur = nd.parent;
    }
}

__global__ void emit_kernel(const unsigned long long *keys, const Aabb *bounds, int n, const TreeNode *nodes,
                            const Aabb *node_box, const int *collapsed, HostNode *out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    const TreeNode nd = nodes[i];
    HostNode h;
    const int ch[2] = {nd.left, nd.right};
    for (int c = 0; c < 2; c++) {
        Aabb b;
        int ref;
        if (ch[c] < 0) {
            const int pos = ~ch[c];
            b = bounds[(unsigned) keys[pos]];
            ref = ~((pos << 3) | 0);
        } else {
            b = node_box[ch[c]];
            if (collapsed[ch[c]]) {
                const TreeNode cn = nodes[ch[c]];
                ref = ~((cn.first << 3) | (cn.last - cn.first));
            } else {
                ref = ch[c];
            }
        }
        for (int k = 0; k < 3; k++) {
            if (c == 0) h.c0mn[k] = b.mn[k], h.c0mx[k] = b.mx[k];
            else h.c1mn[k] = b.mn[k], h.c1mx[k] = b.mx[k];
        }
        if (c == 0) h.child0 = ref;
        else h.child1 = ref;
    }
    out[i] = h;
}

// ---- PLOC: parallel locally-ordered clustering (Meister & Bittner 2018) on the Morton-sorted primitives --------
// One CTA of 1024 threads (the shipped scenes have <= 32 K primitives; the build is a one-off per scene):
// every round, each cluster looks `kPlocRadius` neighbours left and right in the current (Morton) order for the
// partner whose merged box has the smallest area; mutual nearest neighbours merge into a new node; the cluster
// list is compacted with a block-wide prefix sum.  SAH cost / leaf collapse are evaluated when a node is
// created (children always exist already); DFS positions are then pushed down batch by batch so that every
// subtree owns a contiguous range of the final primitive order.
constexpr int kPlocRadiusDefault = 16;
constexpr float kPlocLeafCost = 1.0f;  // per-primitive cost in the collapse decision (tools/ploc_tune.py: 1.0 beats 1.6 and 2.5)
constexpr int kPlocThreads = 1024;

struct PlocNode {
    Aabb box;
    int left, right;   // node ids; leaves are ids 0..n-1 (sorted position), internal ids n..2n-2
    int size;          // primitives below
    int first;         // DFS position of the leftmost primitive
    float cost;
    int collapsed;
};

__device__ __forceinline__ int block_exclusive_scan(int v, int *smem, int &total) {
    // smem: kPlocThreads ints
    const int t = threadIdx.x;
    smem[t] = v;
    __syncthreads();
    for (int off = 1; off < kPlocThreads; off <<= 1) {
        int add = (t >= off) ? smem[t - off] : 0;
        __syncthreads();
        smem[t] += add;
        __syncthreads();
    }
    total = smem[kPlocThreads - 1];
    const int incl = smem[t];
    __syncthreads();
    return incl - v;
}

__global__ void __launch_bounds__(kPlocThreads, 1)
ploc_kernel(const unsigned long long *keys, const Aabb *bounds, int n, PlocNode *nodes, int *cl_a, int *cl_b, int *nn,
            int *batch_start, int *n_batches_out, HostNode *out, int *prim_order, int do_collapse, int kPlocRadius,
            float cost_prim) {
    __shared__ int scan[kPlocThreads];
    const int t = threadIdx.x;
    for (int i = t; i < n; i += kPlocThreads) {
        PlocNode nd;
        nd.box = bounds[(unsigned) keys[i]];
        nd.left = nd.right = -1;
        nd.size = 1;
        nd.first = 0;
        nd.cost = cost_prim;
        nd.collapsed = 0;
        nodes[i] = nd;
        cl_a[i] = i;
    }
    __syncthreads();
    int m = n, n_alloc = n, n_batches = 0;
    int *cin = cl_a, *cout = cl_b;
    while (m > 1) {
        if (t == 0) batch_start[n_batches] = n_alloc;
        // 1. nearest neighbour within the radius (merged half-area), ties to the lower index
        for (int i = t; i < m; i += kPlocThreads) {
            const Aabb bi = nodes[cin[i]].box;
            float best = FLT_MAX;
            int bj = -1;
            const int lo = max(0, i - kPlocRadius), hi = min(m - 1, i + kPlocRadius);
            for (int j = lo; j <= hi; j++) {
                if (j == i) continue;
                const float a = half_area(merge(bi, nodes[cin[j]].box));
                if (a < best) best = a, bj = j;
            }
            nn[i] = bj;
        }
        __syncthreads();
        // 2. mutual pairs merge (the lower index keeps the slot); count new nodes and surviving clusters
        const int per = (m + kPlocThreads - 1) / kPlocThreads;
        const int b0 = min(m, t * per), b1 = min(m, b0 + per);
        int my_new = 0, my_keep = 0;
        for (int i = b0; i < b1; i++) {
            const int j = nn[i];
            const bool mutual = j >= 0 && nn[j] == i;
            if (mutual && i < j) my_new++;
            if (!(mutual && j < i)) my_keep++;
        }
        int tot_new, tot_keep;
        int off_new = block_exclusive_scan(my_new, scan, tot_new);
        int off_keep = block_exclusive_scan(my_keep, scan, tot_keep);
        for (int i = b0; i < b1; i++) {
            const int j = nn[i];
            const bool mutual = j >= 0 && nn[j] == i;
            if (mutual && j < i) continue;  // absorbed by its partner
            int id = cin[i];
            if (mutual) {
                const int l = cin[i], r = cin[j];
                const PlocNode nl = nodes[l], nr = nodes[r];
                PlocNode nd;
                nd.box = merge(nl.box, nr.box);
                nd.left = l;
                nd.right = r;
                nd.size = nl.size + nr.size;
                nd.first = 0;
                const float a = half_area(nd.box);
                float c_split = kCostNode + (a > 0.0f ? (half_area(nl.box) * nl.cost + half_area(nr.box) * nr.cost) / a : nl.cost + nr.cost);
                const float c_leaf = cost_prim * (float) nd.size;
                nd.collapsed = 0;
                if (do_collapse && nd.size <= kMaxLeafPrims && c_leaf <= c_split) {
                    nd.collapsed = 1;
                    c_split = c_leaf;
                }
                nd.cost = c_split;
                id = n_alloc + off_new++;
                nodes[id] = nd;
            }
            cout[off_keep++] = id;
        }
        __syncthreads();
        n_alloc += tot_new;
        m = tot_keep;
        n_batches++;
        int *tmp = cin;
        cin = cout;
        cout = tmp;
        __syncthreads();
    }
    const int root = cin[0];  // == n_alloc - 1 == 2n - 2
    if (t == 0) {
        batch_start[n_batches] = n_alloc;
        *n_batches_out = n_batches;
        nodes[root].first = 0;
        nodes[root].collapsed = 0;  // node 0 of the traversal tree must be an inner node
    }
    __syncthreads();
    // 3. DFS positions, top-down, one creation batch at a time (children are always older than their parent)
    for (int b = n_batches - 1; b >= 0; b--) {
        for (int id = batch_start[b] + t; id < batch_start[b + 1]; id += kPlocThreads) {
            const PlocNode nd = nodes[id];
            nodes[nd.left].first = nd.first;
            nodes[nd.right].first = nd.first + nodes[nd.left].size;
        }
        __syncthreads();
    }
    for (int i = t; i < n; i += kPlocThreads) prim_order[nodes[i].first] = (int) (unsigned) keys[i];
    // 4. emit the 64-byte traversal nodes: internal id -> index (2n-2 - id), so the root lands at 0
    for (int id = n + t; id < 2 * n - 1; id += kPlocThreads) {
        const PlocNode nd = nodes[id];
        HostNode h;
        const int ch[2] = {nd.left, nd.right};
        for (int c = 0; c < 2; c++) {
            const PlocNode cn = nodes[ch[c]];
            int ref;
            if (ch[c] < n) ref = ~((cn.first << 3) | 0);
            else if (cn.collapsed) ref = ~((cn.first << 3) | (cn.size - 1));
            else ref = (2 * n - 2) - ch[c];
            for (int k = 0; k < 3; k++) {
                if (c == 0) h.c0mn[k] = cn.box.mn[k], h.c0mx[k] = cn.box.mx[k];
                else h.c1mn[k] = cn.box.mn[k], h.c1mx[k] = cn.box.mx[k];
            }
            if (c == 0) h.child0 = ref;
            else h.child1 = ref;
        }
        out[(2 * n - 2) - id] = h;
    }
}

#define CK(call)                         \
    do {                                 \
        if ((call) != cudaSuccess) {     \
            cleanup();                   \
            return -1;                   \
        }                                \
    } while (0)

}  // namespace

int build_bvh_device(const std::vector<Aabb> &bounds, HostBvh &out, float *ms_device, int use_ploc) {
    out = HostBvh();
    const int n = (int) bounds.size();
    if (n == 0) return 0;
    if (n == 1) {  // a single leaf under a root node
        HostNode nd;
        for (int k = 0; k < 3; k++) {
            nd.c0mn[k] = bounds[0].mn[k], nd.c0mx[k] = bounds[0].mx[k];
            nd.c1mn[k] = FLT_MAX, nd.c1mx[k] = -FLT_MAX;
        }
        nd.child0 = ~0;
        nd.child1 = kEmptyChild;
        out.nodes.push_back(nd);
        out.prim_order = {0};
        return 0;
    }
    int n_pad = 1;
    while (n_pad < n) n_pad <<= 1;
    if (n_pad < kSortTile) n_pad = kSortTile;

    Aabb *d_bounds = nullptr, *d_box = nullptr;
    unsigned *d_cb = nullptr, *d_flags = nullptr;
    unsigned long long *d_keys = nullptr;
    TreeNode *d_nodes = nullptr;
    int *d_leaf_parent = nullptr, *d_collapsed = nullptr;
    float *d_cost = nullptr;
    HostNode *d_out = nullptr;
    PlocNode *d_ploc = nullptr;
    int *d_cl_a = nullptr, *d_cl_b = nullptr, *d_nn = nullptr, *d_batch = nullptr, *d_order = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    auto cleanup = [&]() {
        cudaFree(d_bounds); cudaFree(d_box); cudaFree(d_cb); cudaFree(d_flags); cudaFree(d_keys); cudaFree(d_nodes);
        cudaFree(d_leaf_parent); cudaFree(d_collapsed); cudaFree(d_cost); cudaFree(d_out);
        cudaFree(d_ploc); cudaFree(d_cl_a); cudaFree(d_cl_b); cudaFree(d_nn); cudaFree(d_batch); cudaFree(d_order);
        if (e0) cudaEventDestroy(e0);
        if (e1) cudaEventDestroy(e1);
    };
    CK(cudaMalloc(&d_bounds, sizeof(Aabb) * n));
    CK(cudaMalloc(&d_box, sizeof(Aabb) * n));
    CK(cudaMalloc(&d_cb, sizeof(unsigned) * 6));
    CK(cudaMalloc(&d_flags, sizeof(unsigned) * n));
    CK(cudaMalloc(&d_keys, sizeof(unsigned long long) * n_pad));
    CK(cudaMalloc(&d_nodes, sizeof(TreeNode) * n));
    CK(cudaMalloc(&d_leaf_parent, sizeof(int) * n));
    CK(cudaMalloc(&d_collapsed, sizeof(int) * n));
    CK(cudaMalloc(&d_cost, sizeof(float) * n));
    CK(cudaMalloc(&d_out, sizeof(HostNode) * n));
    if (use_ploc) {
        CK(cudaMalloc(&d_ploc, sizeof(PlocNode) * (2 * (size_t) n)));
        CK(cudaMalloc(&d_cl_a, sizeof(int) * n));
        CK(cudaMalloc(&d_cl_b, sizeof(int) * n));
        CK(cudaMalloc(&d_nn, sizeof(int) * n));
        CK(cudaMalloc(&d_batch, sizeof(int) * 4096));
        CK(cudaMalloc(&d_order, sizeof(int) * n));
    }
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    CK(cudaMemcpy(d_bounds, bounds.data(), sizeof(Aabb) * n, cudaMemcpyHostToDevice));
    const unsigned cb_init[6] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0u, 0u, 0u};
    CK(cudaMemcpy(d_cb, cb_init, sizeof cb_init, cudaMemcpyHostToDevice));
    CK(cudaMemset(d_flags, 0, sizeof(unsigned) * n));
    CK(cudaMemset(d_collapsed, 0, sizeof(int) * n));
    CK(cudaMemset(d_nodes, 0xff, sizeof(TreeNode) * n));

    CK(cudaEventRecord(e0));
    const int T = 256;
    centroid_bounds_kernel<<<min(148 * 4, (n + T - 1) / T), T>>>(d_bounds, n, d_cb);
    morton_kernel<<<(n_pad + T - 1) / T, T>>>(d_bounds, n, n_pad, d_cb, d_keys);
    bitonic_local_kernel<<<n_pad / kSortTile, 1024>>>(d_keys, n_pad);
    for (int k = kSortTile * 2; k <= n_pad; k <<= 1) {
        for (int j = k >> 1; j >= kSortTile; j >>= 1)
            bitonic_global_kernel<<<(n_pad / 2 + T - 1) / T, T>>>(d_keys, n_pad, j, k);
        bitonic_merge_local_kernel<<<n_pad / kSortTile, 1024>>>(d_keys, n_pad, k);
    }
    if (use_ploc) {
        // tuning knobs for experiments (tools/builder_ab.py): search radius and the leaf cost of the SAH collapse
        const char *er = getenv("RT_B200_PLOC_RADIUS"), *ec = getenv("RT_B200_PLOC_LEAF_COST");
        const int radius = er ? atoi(er) : kPlocRadiusDefault;
        const float cost_prim = ec ? (float) atof(ec) : kPlocLeafCost;
        ploc_kernel<<<1, kPlocThreads>>>(d_keys, d_bounds, n, d_ploc, d_cl_a, d_cl_b, d_nn, d_batch, d_batch + 4095, d_out, d_order,
                                         cost_prim < 1e9f, radius, cost_prim);
    } else {
        radix_tree_kernel<<<(n - 1 + T - 1) / T, T>>>(d_keys, n, d_nodes, d_leaf_parent);
        refit_kernel<<<(n + T - 1) / T, T>>>(d_keys, d_bounds, n, d_nodes, d_leaf_parent, d_box, d_cost, d_collapsed, d_flags, 1);
        emit_kernel<<<(n - 1 + T - 1) / T, T>>>(d_keys, d_bounds, n, d_nodes, d_box, d_collapsed, d_out);
    }
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    CK(cudaGetLastError());
    if (ms_device) cudaEventElapsedTime(ms_device, e0, e1);

    out.nodes.resize((size_t) n - 1);
    CK(cudaMemcpy(out.nodes.data(), d_out, sizeof(HostNode) * (n - 1), cudaMemcpyDeviceToHost));
    std::vector<unsigned long long> keys((size_t) n);
    CK(cudaMemcpy(keys.data(), d_keys, sizeof(unsigned long long) * n, cudaMemcpyDeviceToHost));
    out.prim_order.resize((size_t) n);
    if (use_ploc) CK(cudaMemcpy(out.prim_order.data(), d_order, sizeof(int) * n, cudaMemcpyDeviceToHost));
    else
        for (int i = 0; i < n; i++) out.prim_order[i] = (int) (unsigned) keys[i];
    cleanup();
    out.sah_cost = bvh_sah_cost(out);
    return 0;
}

}  // namespace rtb
