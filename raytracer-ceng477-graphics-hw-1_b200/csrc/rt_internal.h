// rt_internal.h — structures shared by the host-side builders and the CUDA kernels.
#pragma once

#include <cstdint>
#include <vector>

#include "rt_b200.h"

// Bounds-checked build (tools/build_variants.sh checked:-DRT_BOUNDS_CHECK; tests/test_gpu_parity.py::
// test_bounds_checked_build runs it in a subprocess): every computed index of the kernels is asserted against the extent
// of the array it goes into.  compute-sanitizer is closed on the GPU pool this was developed on, so this is the memory-
// safety evidence; in the product build the macro compiles to nothing.
#ifdef RT_BOUNDS_CHECK
#include <cassert>
#define RT_CHECK(cond) assert(cond)
#else
#define RT_CHECK(cond) ((void) 0)
#endif

namespace rtb {

// ---------------------------------------------------------------------------------------------
// Device data layout (everything is read-only during rendering and, at < 5 MB for the largest
// shipped scene, L2-resident; all records are 16-byte aligned so every fetch is one LDG.128).
//
//  nodes      float4[4 * n_nodes]   BVH2, both children's boxes in the parent (64 B per node):
//                                   (per axis centre c and half-extent h of the padded box)
//                                    [0] = c0.c.x c0.h.x c0.c.y c0.h.y
//                                    [1] = c1.c.x c1.h.x c1.c.y c1.h.y
//                                    [2] = c0.c.z c0.h.z c1.c.z c1.h.z
//                                    [3] = bits(child0) bits(child1) - -
//                                   child >= 0: inner node as float4 index (4 * node index: the kernel adds it to
//                                   the base pointer unscaled); child < 0: leaf, ~child = (first << 3) | (count - 1)
//                                   into `prims`; a missing child is a leaf over the never-hit dummy primitive np
//  prims      float4[3 * n_prims]   primitives in leaf order (48 B each):
//                                    triangle [0] = a.x a.y a.z bits(prim_id)      (a = vertex v0)
//                                             [1] = (a-b).xyz  bits(0)
//                                             [2] = (a-c).xyz  (a-b).y*(a-c).z - (a-c).y*(a-b).z   (ray-independent minor)
//                                    sphere   [0] = c.x c.y c.z bits(prim_id)
//                                             [1] = radius 0 0 bits(1)
//                                   prim_id: triangles 0..nt-1 in the reference's list order, spheres nt..nt+ns-1
//  tri_nm     float4[nt]            unit geometric normal (raytracer.cpp:346) + bits(material_id)
//  tri_nn     float4[nt]            the stored normal normalised once more (what raytracer.cpp:414/:432 compute per hit)
//  sph_cr     float4[ns]            centre + radius; sph_mat int[ns]
//  ranks      uint32[8 * n_prims]   visit rank of prim_id in the reference's traversal order for each
//                                   ray-direction sign octant (bit a <=> dir[a] > 0) — exact-t tie breaking
//  ref_nodes  float4[3 * n_ref]     the reference's own tree: [0] = min.xyz bits(axis | is_leaf << 2)
//                                   [1] = max.xyz bits(left child)  [2] = bits(first) bits(count) - -   (right child = left + 1)
//  ref_leaf_prims int[n_prims]      prim ids in the reference's leaf order; slot_of_prim int[n_prims]; prim_bounds float4[2 * n_prims]
//  materials  float4[4 * nm]        [0] = ka.xyz phong  [1] = kd.xyz bits(is_mirror | no_specular << 1)  [2] = ks.xyz 0  [3] = km.xyz 0
//  lights     float4[2 * nl]        [0] = position.xyz  [1] = intensity.xyz
// ---------------------------------------------------------------------------------------------

constexpr int kEmptyChild = 0x7fffffff;
constexpr int kMaxLeafPrims = 8;
constexpr int kMaxSupportedDepth = 32;  // max_recursion_depth accepted by rt_scene_create

struct Aabb {
    float mn[3], mx[3];
};

struct HostNode {  // one 64-byte device node, host view
    float c0mn[3], c0mx[3], c1mn[3], c1mx[3];
    int child0, child1;
};

struct HostBvh {
    std::vector<HostNode> nodes;      // node 0 is the root (absent when there are no primitives)
    std::vector<int> prim_order;      // leaf-ordered prim ids
    int max_depth = 0;
    float sah_cost = 0;
};

// Bounds of every primitive exactly as the reference computes them (parser.h:272-317): vertex
// min/max for triangles, centre -/+ radius (float ops) for spheres.
void primitive_bounds(const RtSceneDesc &d, std::vector<Aabb> &out);

// Reference-order ranks: rebuilds the reference's midpoint-split tree
// (bvh.h:48-163) with the same float operations and writes, for each of the 8 direction-sign
// octants, the position of every primitive in the reference's leaf visit order.
struct RefTreeStats {
    int nodes = 0, leaves = 0, max_leaf = 0, max_depth = 0;
};
// The reference's flattened tree (pre-order, left child = index + 1, bvh.h:81-105) with its un-padded boxes:
// used on the device to decide whether the reference's own box culling would have reached a hit
// (DESIGN.md "reference visibility") and, for the doubtful few rays, to replay its traversal exactly.
struct RefTreeNode {
    float mn[3], mx[3];
    int axis, is_leaf, right, first, count;
};
struct RefTree {
    std::vector<RefTreeNode> nodes;
    std::vector<int> leaf_prims;    // prim ids in the reference's leaf order (triangles before spheres inside a leaf)
    std::vector<int> leaf_of_prim;  // node index of the leaf holding each primitive
};
void build_reference_ranks(const RtSceneDesc &d, std::vector<uint32_t> &ranks /* [8][np] */, RefTreeStats &stats,
                           RefTree *tree = nullptr);

// Host binned-SAH BVH2 (quality yardstick and fallback for tiny scenes).
// build_bvh_sah_host = the plain top-down build followed by reinsert_optimize_host with the library's defaults
// (kReinsertDefaultRounds rounds, kept when the SAH cost falls below kReinsertAccept x the plain tree's).
void build_bvh_sah_host_plain(const std::vector<Aabb> &bounds, HostBvh &out);
void build_bvh_sah_host(const std::vector<Aabb> &bounds, HostBvh &out, int reinsert_rounds = -1, float reinsert_accept = 0.0f);

// Insertion-based optimisation (reinsert_core.h) of a finished tree, on the host; the GPU pipeline runs the same rounds
// in bvh_reinsert.cu.
struct ReinsertReport {
    float cost_before = 0, cost_after = 0;  // SAH cost of the tree as built / after the rounds
    int moves = 0, rounds = 0, height = 0;
    bool accepted = false;
};
int reinsert_optimize_host(HostBvh &bvh, int rounds, float accept_ratio, ReinsertReport *report = nullptr);

// Outward padding applied to every child box before upload (conservative node test; the
// primitive tests stay exact).  See DESIGN.md "why the boxes are padded".
void pad_boxes(HostBvh &bvh, const std::vector<Aabb> &bounds);

float bvh_sah_cost(const HostBvh &bvh);
void compact_dfs(HostBvh &bvh, int bfs_top = 0);

}  // namespace rtb
