// reinsert_core.h — insertion-based optimisation of a finished BVH2, the part shared by the GPU kernel
// (bvh_reinsert.cu) and the host builder (bvh_host.cpp): same code, same float operations (-fmad=false /
// -ffp-contract=off), hence the same tree on both sides.
//
// Why: the traversal kernel is issue-bound and 46 % of its instructions are node steps, so the
// number of node steps per ray is its cost.  A top-down binned-SAH tree carries a huge primitive (the floor of
// horse_and_mug) deep into the hierarchy; PLOC avoids that but clusters locally.  Removing a subtree and re-inserting
// it where the tree's total box area grows least (Bittner, Hapala, Havran 2013) repairs both: on horse_and_mug the
// optimised top-down tree needs 5.77 node steps per ray against 6.51 (PLOC) and 8.58 (top-down as built) —
// tools/tree_lab.cpp walks the renderer's rays through candidate trees on the CPU and is where this was designed.
//
// Shape (Meister & Bittner 2018, "parallel reinsertion"), one ROUND:
//   1. every node x searches its best new position on the unchanged tree: walk up x's ancestors ("pivots"); the
//      ancestors' boxes are shrunk as if x were gone (`saved` = area released so far); in each pivot's other subtree a
//      branch-and-bound descent looks for the node y next to which x costs least (direct area of y+x plus the growth
//      induced on y's ancestors below the pivot); the shrunken ancestors themselves are candidates too;
//   2. the move locks both paths x -> pivot <- y with atomicMax(key(round, gain, x));
//   3. moves that own all their locks are applied (x's parent node is re-used as the new common parent of y and x)
//      and refit the boxes of their own paths; the pivot's box does not change.
// Winning moves touch disjoint node sets, so the predicted gains are exact and the total area falls monotonically.
// A handful of rounds is enough: the big repairs (the floor moves to the root) win the early rounds.
//
// The tree is held as ENTITY arrays: inner node i of the builder's tree is entity i; a leaf (a contiguous range of the
// primitive order, referenced from its parent's child slot) is entity `cap + first primitive of the range`.
// left[e] < 0 marks a leaf entity and holds its leaf reference.
#pragma once

#include <cfloat>
#include <cstdint>

#include "rt_internal.h"

#if defined(__CUDACC__)
#define RT_HD __host__ __device__ __forceinline__
#else
#define RT_HD inline
#endif

namespace rtb {

struct ReinsertView {
    Aabb *box;
    int *left, *right, *parent;  // parent[root] = -1
};
struct ReinsertMove {
    float gain;  // area removed from the tree
    int y;       // x becomes y's sibling
    int pivot;   // lowest node whose box the move leaves unchanged: both paths end there
};

constexpr int kReinsertStack = 72;     // >= tree height + 2 (every builder's output is at most 64 levels deep)
constexpr int kReinsertMaxWalk = 160;  // hard bound on every upward walk
constexpr int kReinsertDefaultRounds = 8;
// The optimised tree replaces the builder's only when its SAH cost is clearly lower (horse_and_mug: 4.38 against 7.19).
// Where the top-down tree is already good the optimisation still lowers the SAH cost by 1-13 %, but the rays do not
// follow: secondary rays start ON surfaces, which the SAH's "rays from outside" model ignores (car: SAH 3.58 -> 3.12,
// node steps per ray 8.20 -> 8.72; bunny, dragon: +-1 %; tools/tree_lab.cpp).  The same 0.8 rule chooses between the
// PLOC and the top-down tree (scene_build.cu, choose_kernel).
constexpr float kReinsertAccept = 0.8f;

RT_HD Aabb box_merge(const Aabb &a, const Aabb &b) {
    Aabb r;
    for (int k = 0; k < 3; k++) {
        r.mn[k] = a.mn[k] < b.mn[k] ? a.mn[k] : b.mn[k];
        r.mx[k] = a.mx[k] > b.mx[k] ? a.mx[k] : b.mx[k];
    }
    return r;
}
RT_HD float box_half_area(const Aabb &b) {
    const float dx = b.mx[0] - b.mn[0], dy = b.mx[1] - b.mn[1], dz = b.mx[2] - b.mn[2];
    if (dx < 0 || dy < 0 || dz < 0) return 0;
    return dx * dy + dy * dz + dz * dx;
}

// round (6 bits) | gain (31 bits of a positive float) | entity (27 bits): later rounds beat stale locks, so the lock
// array is never cleared; within a round the largest gain wins, ties go to the larger entity id
RT_HD unsigned long long reinsert_key(int round, float gain, int x) {
    union {
        float f;
        unsigned u;
    } c;
    c.f = gain;
    return ((unsigned long long) (round + 1) << 58) | ((unsigned long long) (c.u & 0x7fffffffu) << 27) | (unsigned) x;
}
constexpr int kReinsertMaxEntities = 1 << 27;
constexpr int kReinsertMaxRounds = 62;

struct ReinsertSearch {
    Aabb xb;
    float xa;
    float best;
    int best_y, best_pivot;
};

// branch and bound below `sub` (a subtree hanging off x's path under `pivot`): `induced` is the growth of the nodes
// between a candidate and the pivot when they take x in
RT_HD void reinsert_explore(const ReinsertView &t, ReinsertSearch &q, int sub, int pivot, float saved) {
    int stack_n[kReinsertStack];
    float stack_i[kReinsertStack];
    int sp = 1;
    stack_n[0] = sub, stack_i[0] = 0.0f;
    while (sp > 0) {
        sp--;
        const int n = stack_n[sp];
        const float induced = stack_i[sp];
        if (saved - (induced + q.xa) <= q.best) continue;  // even a zero-growth position below cannot beat the best
        const Aabb nb = t.box[n];
        const float direct = box_half_area(box_merge(nb, q.xb));
        const float gain = saved - (induced + direct);
        if (gain > q.best) q.best = gain, q.best_y = n, q.best_pivot = pivot;
        const float ci = induced + direct - box_half_area(nb);
        const int l = t.left[n];
        if (l >= 0 && saved - (ci + q.xa) > q.best && sp + 2 <= kReinsertStack) {
            stack_n[sp] = l, stack_i[sp] = ci, sp++;
            stack_n[sp] = t.right[n], stack_i[sp] = ci, sp++;
        }
    }
}

// Best new position of x.  min_gain: smallest area reduction worth a move (a fraction of the root's area).
RT_HD bool reinsert_find(const ReinsertView &t, int x, float min_gain, ReinsertMove &mv) {
    const int p = t.parent[x];
    if (p < 0) return false;
    const int g = t.parent[p];
    if (g < 0) return false;  // children of the root stay: the root is never re-used
    ReinsertSearch q;
    q.xb = t.box[x];
    q.xa = box_half_area(q.xb);
    q.best = 0.0f;
    q.best_y = q.best_pivot = -1;
    const int s = t.left[p] == x ? t.right[p] : t.left[p];
    float saved = box_half_area(t.box[p]);  // p disappears ...
    Aabb shrunk = t.box[s];                 // ... and s takes its place
    reinsert_explore(t, q, s, g, saved);    // x moves down inside p's old box: nothing above p changes (pivot g)
    int cur = p;
    for (int walk = 0; walk < kReinsertMaxWalk; walk++) {
        const int a = t.parent[cur];
        if (a < 0) break;
        const int u = t.left[a] == cur ? t.right[a] : t.left[a];
        // positions below a's other child: a takes x back and keeps its box, so `saved` stops short of a
        reinsert_explore(t, q, u, a, saved);
        // a itself, shrunk, as x's new sibling: the new parent then has a's old box
        shrunk = box_merge(shrunk, t.box[u]);
        const float sa = box_half_area(shrunk);
        const int pa = t.parent[a];
        if (pa >= 0 && saved - sa > q.best) q.best = saved - sa, q.best_y = a, q.best_pivot = pa;
        saved += box_half_area(t.box[a]) - sa;
        cur = a;
    }
    if (q.best_y < 0 || q.best_y == s || !(q.best > min_gain)) return false;
    mv.gain = q.best;
    mv.y = q.best_y;
    mv.pivot = q.best_pivot;
    return true;
}

// The nodes a move locks: x's parent up to the pivot (inclusive), y up to the pivot (exclusive).  fn(node) -> bool
// (false stops the walk); returns false when fn did or a walk ran away.
template <class F>
RT_HD bool reinsert_paths(const ReinsertView &t, int x, int y, int pivot, F fn) {
    int a = t.parent[x];
    for (int i = 0;; i++) {
        if (a < 0 || i >= kReinsertMaxWalk) return false;
        if (!fn(a)) return false;
        if (a == pivot) break;
        a = t.parent[a];
    }
    a = y;
    for (int i = 0; a != pivot; i++) {
        if (a < 0 || i >= kReinsertMaxWalk) return false;
        if (!fn(a)) return false;
        a = t.parent[a];
    }
    return true;
}

// Applies a winning move.  Only the locked nodes, x's sibling's parent link and x/y's parent links are written.
RT_HD void reinsert_apply(const ReinsertView &t, int x, int y, int pivot) {
    const int p = t.parent[x], g = t.parent[p];
    const int s = t.left[p] == x ? t.right[p] : t.left[p];
    // detach: s takes p's place
    if (t.left[g] == p) t.left[g] = s;
    else t.right[g] = s;
    t.parent[s] = g;
    for (int a = g, i = 0; a != pivot && a >= 0 && i < kReinsertMaxWalk; a = t.parent[a], i++)
        t.box[a] = box_merge(t.box[t.left[a]], t.box[t.right[a]]);
    // attach: p becomes the parent of (y, x) where y was
    const int yp = t.parent[y];
    t.left[p] = y, t.right[p] = x, t.parent[p] = yp;
    if (t.left[yp] == y) t.left[yp] = p;
    else t.right[yp] = p;
    t.parent[y] = p;
    t.box[p] = box_merge(t.box[y], t.box[x]);
    for (int a = yp, i = 0; a != pivot && a >= 0 && i < kReinsertMaxWalk; a = t.parent[a], i++)
        t.box[a] = box_merge(t.box[t.left[a]], t.box[t.right[a]]);
}

// SAH cost contribution of one inner entity: its two children (kSahCostNode / kSahCostPrim as in build_device.h)
RT_HD float reinsert_node_cost(const ReinsertView &t, int e, float cost_node, float cost_prim) {
    float c = 0;
    const int ch[2] = {t.left[e], t.right[e]};
    for (int k = 0; k < 2; k++) {
        const float a = box_half_area(t.box[ch[k]]);
        const int l = t.left[ch[k]];
        c += l >= 0 ? cost_node * a : cost_prim * a * (float) (((~l) & 7) + 1);
    }
    return c;
}

}  // namespace rtb
