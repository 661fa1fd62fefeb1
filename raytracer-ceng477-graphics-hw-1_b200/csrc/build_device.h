// build_device.h — the GPU-resident scene build (replaces RayTracer::RayTracer, raytracer.cpp:335-350, and
// BVHTree::build, bvh.h:48-178): everything between "flat scene arrays uploaded" and "final traversal buffers in
// HBM" runs as kernels on ONE stream inside ONE arena allocation, without a host synchronisation in between; the
// host reads a 128-byte result block once at the end.
//
//   prep_kernel           per primitive: bounds (parser.h:272-317), split keys (raytracer.cpp:347, bvh.h:131), unit
//                         normal and its re-normalisation (raytracer.cpp:346, :414), the 48-byte primitive record,
//                         scene and centroid bounds (atomics)                                        scene_build.cu
//   reference tree        level-synchronous midpoint splits + 8 visit-rank arrays                    ref_order_device.cu
//   candidate trees       Morton sort -> PLOC (cooperative, multi-CTA) or Karras LBVH; top-down binned SAH
//                                                                                                    bvh_lbvh.cu, bvh_sah_device.cu
//   reinsertion           the top-down tree optimised by moving subtrees to where the total box area grows least
//                         (kept when the SAH cost falls below 0.8x)                                  bvh_reinsert.cu
//   tree_parents/_reduce  per candidate: parent links, then a bottom-up pass (second arrival continues) giving
//                         every node its inner-node count, height and SAH area sum                   scene_build.cu
//   choose_kernel         RT_BUILD_AUTO: keeps the PLOC tree when its SAH cost is < 0.8x the top-down tree's
//   layout_kernel         per reachable node: depth-first index from the subtree counts on the way to the root,
//                         outward padding, centre / half-extent encoding, child references -> final `nodes`
//   place_prims_kernel    primitive records into leaf order, slot_of_prim
#pragma once

#include <cstdint>

#include <cuda_runtime.h>

#include "rt_internal.h"

namespace rtb {

// A builder's output, device-resident: 56-byte nodes with both children's un-padded boxes (HostNode), root at
// index 0, child >= 0 inner index, child < 0 leaf ~((first << 3) | (count - 1)) into prim_order, kEmptyChild none.
// Unreachable nodes may be left behind (the interiors of collapsed subtrees).
struct DevTree {
    HostNode *nodes = nullptr;
    int *prim_order = nullptr;
    int cap = 0;          // node slots that may be in use (upper bound of node indices + 1)
    int *status = nullptr;  // device flag, != 0: the builder gave up (PLOC round cap)
    int *n_used = nullptr;  // device: node slots in use (indices below it hold valid nodes)
    // finalisation scratch (one entry per node slot)
    int *parent = nullptr;     // parent << 1 | side, -1 none
    int *arrivals = nullptr;   // bottom-up pass: children that have reported
    int *count = nullptr;      // inner nodes in the subtree (this one included)
    int *height = nullptr;     // levels below (a node over two leaves: 1)
    float *area = nullptr;     // SAH sum below: sum over children of half_area(child box) * (inner ? Cnode : Cprim * n)
};

// what the device pipeline reports back (one D2H copy of this block at the end of rt_scene_create)
struct BuildResult {
    int chosen;          // index of the candidate tree that was laid out
    int status[2];       // candidates' builder status
    int n_nodes[2];      // reachable inner nodes
    int height[2];
    float sah_cost[2];
    int ref_nodes, ref_leaves, ref_max_leaf, ref_max_depth;
    int ref_overflow;    // reference tree deeper than the replay stack (cannot happen: depth cap 19)
    unsigned scene_bounds[6];     // ordered-uint min xyz, max xyz over all primitive bounds
    unsigned centroid_bounds[6];  // the same over box centres (Morton quantisation)
    float reinsert_cost_before, reinsert_cost_after;  // bvh_reinsert.cu: SAH cost of the top-down tree as built / optimised
    int reinsert_moves, reinsert_rounds, reinsert_accepted;
    int pad[1];
};
static_assert(sizeof(BuildResult) == 128, "BuildResult is copied back as one 128-byte block");

constexpr float kSahCostNode = 1.0f;  // one two-box node step
constexpr float kSahCostPrim = 1.6f;  // one exact (division-bearing) primitive test
constexpr int kMaxTreeHeight = 60;    // traversal stack: 64 entries

// ordered-uint encoding of floats for atomicMin/atomicMax
__host__ __device__ inline unsigned f2ord(float f) {
#if defined(__CUDA_ARCH__)
    const unsigned u = __float_as_uint(f);
#else
    union { float f; unsigned u; } c = {f};
    const unsigned u = c.u;
#endif
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ inline float ord2f(unsigned u) {
    const unsigned v = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
#if defined(__CUDA_ARCH__)
    return __uint_as_float(v);
#else
    union { unsigned u; float f; } c = {v};
    return c.f;
#endif
}

// ---- reference-order tree (ref_order_device.cu) ---------------------------------------------------------------
struct DevRefNode {  // build-time view of one node of the reference's tree
    float mn[3], mx[3];
    int axis, is_leaf, left, right, first, count, depth;
};
struct RefTask {
    int node, lo, hi, depth;
};
struct RefScratch {
    int *ids, *tmp;          // [np] primitive ids in the reference's list order / partition scratch
    RefTask *queue[2];       // [np + 1] each
    int *level_count;        // [kRefLevels + 2] open tasks per level (zeroed)
    int *n_nodes;            // allocated nodes (starts at 1: the root)
    DevRefNode *nodes;       // [2 * np + 1]
};
constexpr int kRefLevels = 21;  // depth cap 19 (bvh.h:18) -> at most 20 levels of splits
size_t ref_scratch_bytes(int np);
// enqueues the whole build on `stream`: writes ranks[8][np], ref_nodes (float4[3] per node), ref_leaf_prims and the
// statistics fields of *result
// (`grid` = co-resident CTAs of the cooperative level loop, ref_max_grid; returns a cudaError_t as int)
int ref_max_grid(int n_sms);
int enqueue_reference_tree(const Aabb *bounds, const float *key, int np, const RefScratch &s, uint32_t *ranks,
                           float4 *ref_nodes, int *ref_leaf_prims, BuildResult *result, int grid, cudaStream_t stream);

// ---- candidate trees ------------------------------------------------------------------------------------------
struct MortonScratch {
    unsigned long long *keys;  // [n_pad]
    int n_pad;
};
int morton_pad(int n);
void enqueue_morton_sort(const Aabb *bounds, int n, const unsigned *centroid_bounds, const MortonScratch &m, cudaStream_t stream);

struct PlocScratch {
    void *nodes;  // PlocNode[2n]
    int *cl_a, *cl_b, *nn;  // [n] each
    int *cta_tot;           // [2 * grid]
    int *state;             // [4] m, n_alloc, rounds, root
};
size_t ploc_node_bytes();
// returns cudaError_t as int; `grid` must not exceed the co-resident CTAs (ploc_max_grid)
int ploc_max_grid(int n_sms);
int enqueue_ploc(const Aabb *bounds, int n, const MortonScratch &m, const PlocScratch &s, DevTree &out, int radius,
                 float leaf_cost, int grid, cudaStream_t stream);

struct LbvhScratch {
    void *nodes;  // TreeNode[n]
    int *leaf_parent, *collapsed;
    unsigned *flags;
    float *cost;
    Aabb *box;
};
size_t lbvh_node_bytes();
void enqueue_lbvh(const Aabb *bounds, int n, const MortonScratch &m, const LbvhScratch &s, DevTree &out, cudaStream_t stream);

struct SahTask {
    int parent;  // node that owns the child slot to patch (-1: root)
    int side;    // 0: child0, 1: child1
    int lo, hi;
};
struct SahScratch {
    int *ids, *tmp;       // [n]
    SahTask *queue[2];    // [n + 1] each
    int *level_count;     // [kSahLevels + 2]
    int *n_nodes;         // allocated nodes
    int *root_ref;
    int cap_tasks, cap_nodes;  // extents of the queues / node array (bounds-checked builds)
};
constexpr int kSahLevels = 64;
int sah_max_grid(int n_sms);
int enqueue_sah(const Aabb *bounds, int n, const SahScratch &s, DevTree &out, const BuildResult *res, int grid, cudaStream_t stream);

// ---- insertion-based optimisation of a finished tree (bvh_reinsert.cu, reinsert_core.h) ----------------------------
struct ReinsertScratch {
    Aabb *box;                       // [ne] entity arrays: inner node i = entity i, leaf range starting at f = t.cap + f
    int *left, *right, *parent;      // [ne]
    unsigned long long *lock, *key;  // [ne]
    int *mv_y, *mv_pivot;            // [ne]
    int *canon;                      // [ne] run-independent id of an entity: depth-first index of the tree as built
    float *partial;                  // [grid] per-CTA partial sums
    int *counters;                   // [reinsert_counter_slots()]
    int ne;
};
int reinsert_max_grid(int n_sms);
int reinsert_counter_slots();
// `rounds` rounds on the tree `t` (all node slots reachable, no missing child — the top-down SAH builder's output);
// the optimised tree replaces t's nodes when its SAH cost < accept_ratio x the cost before and it is not deeper than
// the traversal stack; res->reinsert_* report.  Returns a cudaError_t as int.
int enqueue_reinsert(const DevTree &t, const ReinsertScratch &s, int rounds, float accept_ratio, BuildResult *res, int grid,
                     cudaStream_t stream);

}  // namespace rtb
