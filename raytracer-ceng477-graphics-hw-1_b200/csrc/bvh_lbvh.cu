// bvh_lbvh.cu — BVH build on the GPU (replaces BVHNode::build / BVHTree::build, bvh.h:48-178).
//
//   1. centroid bounds            (ordered-int atomics in the per-primitive pre-pass, scene_build.cu)
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

#include <cooperative_groups.h>
#include <cuda_runtime.h>

#include "build_device.h"

namespace cg = cooperative_groups;

namespace rtb {

namespace {

constexpr float kCostNode = kSahCostNode;
constexpr float kCostPrim = kSahCostPrim;

__device__ __forceinline__ float unordered(unsigned u) { return ord2f(u); }

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
    RT_CHECK(gamma >= 0 && gamma + 1 < n && j >= 0 && j < n);
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
// Every round, each cluster looks `radius` neighbours left and right in the current (Morton) order for the partner
// whose merged box has the smallest area; mutual nearest neighbours merge into a new node; the cluster list is
// compacted in order.  One COOPERATIVE launch runs all rounds: the clusters are dealt in contiguous chunks to the
// CTAs of the grid, the compaction offsets come from a per-CTA total plus a block-wide scan, and the phases of a
// round are separated by grid barriers (3 per round); once at most kPlocTail clusters are left, CTA 0 finishes alone
// with block barriers.  SAH cost / leaf collapse are evaluated when a node is created (children always exist
// already).  A second, ordinary kernel then gives every leaf its depth-first position by walking up the parent
// links (subtree sizes are known), which makes every subtree own a contiguous range of the final primitive order.
// The globally closest pair is always mutual (ties go to the lower index on both sides), so every round merges at
// least one pair; a chain-like input that needs more than kPlocMaxRounds rounds raises the status flag and the caller
// falls back to the top-down builder.
constexpr int kPlocThreads = 1024;
constexpr int kPlocTail = 2048;
constexpr int kPlocMaxRounds = 4096;

struct PlocNode {
    Aabb box;
    int left, right;   // node ids; leaves are ids 0..n-1 (sorted position), internal ids n..2n-2
    int size;          // primitives below
    int parent;        // -1: root
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
            int *cta_tot, int *state, int *status, int do_collapse, int radius, float cost_prim) {
    __shared__ int scan[kPlocThreads];
    cg::grid_group grid = cg::this_grid();
    const int t = threadIdx.x, G = gridDim.x, cta = blockIdx.x;
    for (int i = cta * kPlocThreads + t; i < n; i += G * kPlocThreads) {
        PlocNode nd;
        nd.box = bounds[(unsigned) keys[i]];
        nd.left = nd.right = -1;
        nd.size = 1;
        nd.parent = -1;
        nd.cost = cost_prim;
        nd.collapsed = 0;
        nodes[i] = nd;
        cl_a[i] = i;
    }
    if (G > 1) grid.sync();
    else __syncthreads();
    int m = n, n_alloc = n, rounds = 0;
    int *cin = cl_a, *cout = cl_b;
    bool alone = G == 1;  // CTA 0 finishing the tail (or a single-CTA launch)
    while (m > 1 && rounds < kPlocMaxRounds) {
        if (!alone && m <= kPlocTail) {  // hand the tail over to CTA 0 (uniform decision: m is the same in every CTA)
            alone = true;
            if (cta != 0) break;
        }
        const int ctas = alone ? 1 : G;
        const int my_cta = alone ? 0 : cta;
        // this CTA's contiguous chunk of the cluster list, and this thread's contiguous piece of it
        const int per_cta = (m + ctas - 1) / ctas;
        const int c0 = min(m, my_cta * per_cta), c1 = min(m, c0 + per_cta);
        // 1. nearest neighbour within the radius (merged half-area), ties to the lower index
        for (int i = c0 + t; i < c1; i += kPlocThreads) {
            const Aabb bi = nodes[cin[i]].box;
            float best = FLT_MAX;
            int bj = -1;
            const int lo = max(0, i - radius), hi = min(m - 1, i + radius);
            for (int j = lo; j <= hi; j++) {
                if (j == i) continue;
                const float a = half_area(merge(bi, nodes[cin[j]].box));
                if (a < best) best = a, bj = j;
            }
            nn[i] = bj;
        }
        if (alone) __syncthreads();
        else grid.sync();
        // 2. mutual pairs merge (the lower index keeps the slot); count new nodes and surviving clusters
        const int per = (c1 - c0 + kPlocThreads - 1) / kPlocThreads;
        const int b0 = min(c1, c0 + t * per), b1 = min(c1, b0 + per);
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
        int all_new = tot_new, all_keep = tot_keep;
        if (!alone) {
            if (t == 0) cta_tot[2 * cta] = tot_new, cta_tot[2 * cta + 1] = tot_keep;
            grid.sync();
            all_new = all_keep = 0;
            int before_new = 0, before_keep = 0;
            for (int g = 0; g < G; g++) {  // G <= a few hundred: every thread sums the same values
                const int a = cta_tot[2 * g], b = cta_tot[2 * g + 1];
                if (g < cta) before_new += a, before_keep += b;
                all_new += a, all_keep += b;
            }
            off_new += before_new;
            off_keep += before_keep;
        }
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
                nd.parent = -1;
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
                RT_CHECK(id >= n && id < 2 * n - 1 && l >= 0 && l < id && r >= 0 && r < id);
                nodes[id] = nd;
                nodes[l].parent = id;
                nodes[r].parent = id;
            }
            RT_CHECK(off_keep >= 0 && off_keep < m);
            cout[off_keep++] = id;
        }
        n_alloc += all_new;
        m = all_keep;
        rounds++;
        int *tmp = cin;
        cin = cout;
        cout = tmp;
        if (alone) __syncthreads();
        else grid.sync();
    }
    if (cta == 0 && t == 0) {
        state[0] = m;
        state[1] = n_alloc;
        state[2] = rounds;
        state[3] = cin[0];  // the root when m == 1 (== 2n - 2)
        *status = m > 1 ? 1 : 0;
        if (m == 1) nodes[cin[0]].collapsed = 0;  // node 0 of the traversal tree must be an inner node
    }
}

// depth-first position of every leaf (walk up: add the left sibling's size whenever the path comes from the right),
// then the traversal nodes: internal id -> index (2n-2 - id), so the root lands at 0
__global__ void ploc_order_kernel(const unsigned long long *keys, int n, const PlocNode *nodes, int *first_of, int *prim_order,
                                  const int *status) {
    const int id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= 2 * n - 1 || *status != 0) return;
    // position of the leftmost leaf below `id`
    int pos = 0, cur = id;
    for (int guard = 0; guard < (1 << 22); guard++) {
        const int par = nodes[cur].parent;
        if (par < 0) break;
        if (nodes[par].right == cur) pos += nodes[nodes[par].left].size;
        cur = par;
    }
    RT_CHECK(pos >= 0 && pos < n);
    first_of[id] = pos;
    if (id < n) prim_order[pos] = (int) (unsigned) keys[id];
}

__global__ void ploc_emit_kernel(int n, const PlocNode *nodes, const int *first_of, HostNode *out, const int *status) {
    const int id = n + blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= 2 * n - 1 || *status != 0) return;
    const PlocNode nd = nodes[id];
    HostNode h;
    const int ch[2] = {nd.left, nd.right};
    for (int c = 0; c < 2; c++) {
        const PlocNode cn = nodes[ch[c]];
        const int first = first_of[ch[c]];
        int ref;
        if (ch[c] < n) ref = ~((first << 3) | 0);
        else if (cn.collapsed) ref = ~((first << 3) | (cn.size - 1));
        else ref = (2 * n - 2) - ch[c];
        for (int k = 0; k < 3; k++) {
            if (c == 0) h.c0mn[k] = cn.box.mn[k], h.c0mx[k] = cn.box.mx[k];
            else h.c1mn[k] = cn.box.mn[k], h.c1mx[k] = cn.box.mx[k];
        }
        if (c == 0) h.child0 = ref;
        else h.child1 = ref;
    }
    RT_CHECK((2 * n - 2) - id >= 0 && (2 * n - 2) - id < n - 1);
    out[(2 * n - 2) - id] = h;
}

__global__ void lbvh_order_kernel(const unsigned long long *keys, int n, int *prim_order, int *status) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) prim_order[i] = (int) (unsigned) keys[i];
    if (i == 0) *status = 0;
}

}  // namespace

int morton_pad(int n) {
    int n_pad = 1;
    while (n_pad < n) n_pad <<= 1;
    if (n_pad < kSortTile) n_pad = kSortTile;
    return n_pad;
}

void enqueue_morton_sort(const Aabb *bounds, int n, const unsigned *centroid_bounds, const MortonScratch &m, cudaStream_t stream) {
    const int T = 256, n_pad = m.n_pad;
    morton_kernel<<<(n_pad + T - 1) / T, T, 0, stream>>>(bounds, n, n_pad, centroid_bounds, m.keys);
    bitonic_local_kernel<<<n_pad / kSortTile, 1024, 0, stream>>>(m.keys, n_pad);
    for (int k = kSortTile * 2; k <= n_pad; k <<= 1) {
        for (int j = k >> 1; j >= kSortTile; j >>= 1)
            bitonic_global_kernel<<<(n_pad / 2 + T - 1) / T, T, 0, stream>>>(m.keys, n_pad, j, k);
        bitonic_merge_local_kernel<<<n_pad / kSortTile, 1024, 0, stream>>>(m.keys, n_pad, k);
    }
}

size_t ploc_node_bytes() { return sizeof(PlocNode); }

int ploc_max_grid(int n_sms) {
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ploc_kernel, kPlocThreads, 0) != cudaSuccess || per_sm < 1) return 1;
    return n_sms * per_sm;
}

int enqueue_ploc(const Aabb *bounds, int n, const MortonScratch &m, const PlocScratch &s, DevTree &out, int radius,
                 float leaf_cost, int grid, cudaStream_t stream) {
    // n >= 2.  One CTA per 4096 clusters, at most `grid`; small inputs run as a single CTA without grid barriers.
    int want = (n + 4095) / 4096;
    if (want > grid) want = grid;
    if (want < 1 || n <= kPlocTail) want = 1;
    const unsigned long long *keys = m.keys;
    PlocNode *nodes = (PlocNode *) s.nodes;
    int *cl_a = s.cl_a, *cl_b = s.cl_b, *nn = s.nn, *cta_tot = s.cta_tot, *state = s.state, *status = out.status;
    int do_collapse = leaf_cost < 1e9f ? 1 : 0;
    void *args[] = {&keys, &bounds, &n, &nodes, &cl_a, &cl_b, &nn, &cta_tot, &state, &status, &do_collapse, &radius, &leaf_cost};
    cudaError_t e;
    if (want > 1) e = cudaLaunchCooperativeKernel((void *) ploc_kernel, dim3(want), dim3(kPlocThreads), args, 0, stream);
    else e = cudaLaunchKernel((void *) ploc_kernel, dim3(1), dim3(kPlocThreads), args, 0, stream);
    if (e != cudaSuccess) return (int) e;
    const int T = 256;
    // first_of lives in nn's neighbour cl_b? No: both cluster lists are dead now, reuse cl_a ++ cl_b as first_of[2n]
    int *first_of = s.cl_a;  // cl_a and cl_b are adjacent (scene_build.cu allocates them as one block of 2n ints)
    ploc_order_kernel<<<(2 * n - 1 + T - 1) / T, T, 0, stream>>>(keys, n, nodes, first_of, out.prim_order, status);
    ploc_emit_kernel<<<(n - 1 + T - 1) / T, T, 0, stream>>>(n, nodes, first_of, out.nodes, status);
    return (int) cudaGetLastError();
}

size_t lbvh_node_bytes() { return sizeof(TreeNode); }

void enqueue_lbvh(const Aabb *bounds, int n, const MortonScratch &m, const LbvhScratch &s, DevTree &out, cudaStream_t stream) {
    const int T = 256;
    TreeNode *nodes = (TreeNode *) s.nodes;
    cudaMemsetAsync(s.flags, 0, sizeof(unsigned) * n, stream);
    cudaMemsetAsync(s.collapsed, 0, sizeof(int) * n, stream);
    cudaMemsetAsync(nodes, 0xff, sizeof(TreeNode) * n, stream);
    radix_tree_kernel<<<(n - 1 + T - 1) / T, T, 0, stream>>>(m.keys, n, nodes, s.leaf_parent);
    refit_kernel<<<(n + T - 1) / T, T, 0, stream>>>(m.keys, bounds, n, nodes, s.leaf_parent, s.box, s.cost, s.collapsed, s.flags, 1);
    emit_kernel<<<(n - 1 + T - 1) / T, T, 0, stream>>>(m.keys, bounds, n, nodes, s.box, s.collapsed, out.nodes);
    lbvh_order_kernel<<<(n + T - 1) / T, T, 0, stream>>>(m.keys, n, out.prim_order, out.status);
}

}  // namespace rtb
