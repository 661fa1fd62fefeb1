// bvh_host.cpp — host binned-SAH BVH2 builder (quality yardstick / RT_BUILD_SAH_HOST), the
// outward box padding and the SAH cost metric shared with the GPU builder.
//
// The traversal BVH replaces the reference's acceleration structure (bvh.h:37-181).  It only has
// to be conservative: a node box must contain everything the exact primitive tests can report as
// a hit, so the hit SET equals the reference's and ties are settled by ref_order.cpp's ranks.
#include <algorithm>
#include <cfloat>
#include <cmath>

#include "reinsert_core.h"
#include "rt_internal.h"

namespace rtb {

namespace {

inline void grow(Aabb &a, const Aabb &b) {
    for (int k = 0; k < 3; k++) {
        a.mn[k] = std::min(a.mn[k], b.mn[k]);
        a.mx[k] = std::max(a.mx[k], b.mx[k]);
    }
}
inline Aabb empty_box() { return Aabb{{FLT_MAX, FLT_MAX, FLT_MAX}, {-FLT_MAX, -FLT_MAX, -FLT_MAX}}; }
inline float half_area(const Aabb &b) {
    float dx = b.mx[0] - b.mn[0], dy = b.mx[1] - b.mn[1], dz = b.mx[2] - b.mn[2];
    if (dx < 0 || dy < 0 || dz < 0) return 0;
    return dx * dy + dy * dz + dz * dx;
}

constexpr int kBins = 32;
constexpr float kCostNode = 1.0f;  // one two-box node step
constexpr float kCostPrim = 1.6f;  // one exact (division-bearing) primitive test

struct SahBuilder {
    const std::vector<Aabb> &bounds;
    std::vector<int> ids;
    HostBvh &out;

    SahBuilder(const std::vector<Aabb> &b, HostBvh &o) : bounds(b), out(o) {}

    int encode_leaf(int lo, int hi) { return ~((lo << 3) | (hi - lo - 1)); }

    // returns the child reference for ids[lo,hi) and writes its box
    int build(int lo, int hi, int depth, Aabb &box) {
        box = empty_box();
        Aabb cbox = empty_box();
        for (int i = lo; i < hi; i++) {
            const Aabb &b = bounds[ids[i]];
            grow(box, b);
            for (int k = 0; k < 3; k++) {
                float c = 0.5f * (b.mn[k] + b.mx[k]);
                cbox.mn[k] = std::min(cbox.mn[k], c);
                cbox.mx[k] = std::max(cbox.mx[k], c);
            }
        }
        const int n = hi - lo;
        if (depth > out.max_depth) out.max_depth = depth;
        if (n == 1) return encode_leaf(lo, hi);

        int best_axis = -1, best_bin = -1;
        float best_cost = FLT_MAX;
        const float parent_area = half_area(box);
        for (int axis = 0; axis < 3; axis++) {
            const float c0 = cbox.mn[axis], c1 = cbox.mx[axis];
            if (!(c1 > c0)) continue;
            const float scale = kBins / (c1 - c0);
            Aabb bb[kBins];
            int cnt[kBins];
            for (int b = 0; b < kBins; b++) bb[b] = empty_box(), cnt[b] = 0;
            for (int i = lo; i < hi; i++) {
                const Aabb &pb = bounds[ids[i]];
                int b = std::min(kBins - 1, std::max(0, (int) ((0.5f * (pb.mn[axis] + pb.mx[axis]) - c0) * scale)));
                cnt[b]++;
                grow(bb[b], pb);
            }
            float right_area[kBins];
            Aabb acc = empty_box();
            for (int b = kBins - 1; b > 0; b--) {
                grow(acc, bb[b]);
                right_area[b] = half_area(acc);
            }
            acc = empty_box();
            int nl = 0;
            for (int b = 0; b < kBins - 1; b++) {
                grow(acc, bb[b]);
                nl += cnt[b];
                if (nl == 0 || nl == n) continue;
                float cost = half_area(acc) * nl + right_area[b + 1] * (n - nl);
                if (cost < best_cost) best_cost = cost, best_axis = axis, best_bin = b;
            }
        }
        const float leaf_cost = kCostPrim * n;
        const float split_cost = best_axis < 0 ? FLT_MAX
                                 : kCostNode + kCostPrim * best_cost / (parent_area > 0 ? parent_area : 1e-30f);
        if (n <= kMaxLeafPrims && (best_axis < 0 || leaf_cost <= split_cost)) return encode_leaf(lo, hi);

        int mid;
        if (best_axis >= 0) {
            const float c0 = cbox.mn[best_axis], scale = kBins / (cbox.mx[best_axis] - c0);
            auto it = std::partition(ids.begin() + lo, ids.begin() + hi, [&](int id) {
                const Aabb &pb = bounds[id];
                int b = std::min(kBins - 1, std::max(0, (int) ((0.5f * (pb.mn[best_axis] + pb.mx[best_axis]) - c0) * scale)));
                return b <= best_bin;
            });
            mid = (int) (it - ids.begin());
        } else {
            mid = lo + n / 2;  // all centroids coincide: split the list
        }
        if (mid == lo || mid == hi) mid = lo + n / 2;

        const int me = (int) out.nodes.size();
        out.nodes.push_back(HostNode());
        Aabb b0, b1;
        int c0 = build(lo, mid, depth + 1, b0);
        int c1 = build(mid, hi, depth + 1, b1);
        HostNode &nd = out.nodes[me];
        for (int k = 0; k < 3; k++) {
            nd.c0mn[k] = b0.mn[k], nd.c0mx[k] = b0.mx[k];
            nd.c1mn[k] = b1.mn[k], nd.c1mx[k] = b1.mx[k];
        }
        nd.child0 = c0;
        nd.child1 = c1;
        return me;
    }
};

}  // namespace

void build_bvh_sah_host_plain(const std::vector<Aabb> &bounds, HostBvh &out) {
    out = HostBvh();
    const int np = (int) bounds.size();
    if (np == 0) return;
    SahBuilder b(bounds, out);
    b.ids.resize(np);
    for (int i = 0; i < np; i++) b.ids[i] = i;
    Aabb box;
    out.nodes.reserve((size_t) np);
    int root = b.build(0, np, 0, box);
    if (root < 0) {  // the whole scene is one leaf: give it a parent so that node 0 always exists
        HostNode nd;
        for (int k = 0; k < 3; k++) {
            nd.c0mn[k] = box.mn[k], nd.c0mx[k] = box.mx[k];
            nd.c1mn[k] = FLT_MAX, nd.c1mx[k] = -FLT_MAX;
        }
        nd.child0 = root;
        nd.child1 = kEmptyChild;
        out.nodes.push_back(nd);
    }
    out.prim_order = b.ids;
    out.sah_cost = bvh_sah_cost(out);
}

// Insertion-based optimisation of a finished host tree: the rounds of reinsert_core.h run one after the other, every
// round exactly as the GPU kernel runs it (all searches on the unchanged tree, locks by largest key, winners applied),
// so the host and the device produce the same tree from the same input.  Returns the number of applied moves; the
// tree is only replaced when the optimised SAH cost is below accept_ratio x the cost it came with and the result
// fits the traversal stack.
int reinsert_optimize_host(HostBvh &bvh, int rounds, float accept_ratio, ReinsertReport *report) {
    ReinsertReport rep;
    const int cap = (int) bvh.nodes.size(), np = (int) bvh.prim_order.size();
    rep.cost_before = rep.cost_after = bvh.sah_cost;
    if (report) *report = rep;
    if (cap < 3 || rounds <= 0) return 0;
    for (auto &n: bvh.nodes)
        if (n.child0 == kEmptyChild || n.child1 == kEmptyChild) return 0;
    const int ne = cap + np + 1;
    if (ne >= kReinsertMaxEntities) return 0;
    if (rounds > kReinsertMaxRounds) rounds = kReinsertMaxRounds;
    std::vector<Aabb> box(ne);
    std::vector<int> left(ne, -1), right(ne, -1), parent(ne, -1);
    for (int i = 0; i < cap; i++) {
        const HostNode &n = bvh.nodes[i];
        const int ch[2] = {n.child0, n.child1};
        const float *mns[2] = {n.c0mn, n.c1mn}, *mxs[2] = {n.c0mx, n.c1mx};
        int ce[2];
        for (int c = 0; c < 2; c++) {
            ce[c] = ch[c] >= 0 ? ch[c] : cap + ((~ch[c]) >> 3);
            if (ch[c] < 0) left[ce[c]] = ch[c];
            for (int k = 0; k < 3; k++) box[ce[c]].mn[k] = mns[c][k], box[ce[c]].mx[k] = mxs[c][k];
            parent[ce[c]] = i;
        }
        left[i] = ce[0], right[i] = ce[1];
    }
    box[0] = box_merge(box[left[0]], box[right[0]]);
    ReinsertView t{box.data(), left.data(), right.data(), parent.data()};
    auto cost = [&]() {
        double c = 0;
        std::vector<int> todo(1, 0);
        while (!todo.empty()) {
            const int e = todo.back();
            todo.pop_back();
            c += reinsert_node_cost(t, e, kCostNode, kCostPrim);
            if (left[left[e]] >= 0) todo.push_back(left[e]);
            if (left[right[e]] >= 0) todo.push_back(right[e]);
        }
        const float ra = box_half_area(box[0]);
        return ra > 0 ? (float) (kCostNode + c / ra) : 0.0f;
    };
    rep.cost_before = cost();
    const float min_gain = 1e-6f * box_half_area(box[0]);
    std::vector<unsigned long long> lock(ne, 0ull), key(ne);
    std::vector<ReinsertMove> mv(ne);
    std::vector<char> has(ne);
    for (int round = 0; round < rounds; round++) {
        for (int x = 0; x < ne; x++) {
            has[x] = (x == 0 || parent[x] >= 0) && reinsert_find(t, x, min_gain, mv[x]);
            if (!has[x]) continue;
            key[x] = reinsert_key(round, mv[x].gain, x);
            if (!reinsert_paths(t, x, mv[x].y, mv[x].pivot, [&](int n) { lock[n] = std::max(lock[n], key[x]); return true; })) has[x] = 0;
        }
        for (int x = 0; x < ne; x++)
            if (has[x] && !reinsert_paths(t, x, mv[x].y, mv[x].pivot, [&](int n) { return lock[n] == key[x]; })) has[x] = 0;
        int applied = 0;
        for (int x = 0; x < ne; x++)
            if (has[x]) reinsert_apply(t, x, mv[x].y, mv[x].pivot), applied++;
        rep.moves += applied;
        rep.rounds = round + 1;
        if (applied == 0) break;
    }
    rep.cost_after = cost();
    // height of the optimised tree (a node over two leaves: 1)
    int height = 0;
    {
        std::vector<std::pair<int, int>> todo(1, {0, 1});
        while (!todo.empty()) {
            auto [e, d] = todo.back();
            todo.pop_back();
            height = std::max(height, d);
            if (left[left[e]] >= 0) todo.emplace_back(left[e], d + 1);
            if (left[right[e]] >= 0) todo.emplace_back(right[e], d + 1);
        }
    }
    rep.height = height;
    rep.accepted = rep.cost_after < accept_ratio * rep.cost_before && height <= 60;
    if (rep.accepted) {
        for (int i = 0; i < cap; i++) {
            HostNode &n = bvh.nodes[i];
            const int ce[2] = {left[i], right[i]};
            for (int k = 0; k < 3; k++) {
                n.c0mn[k] = box[ce[0]].mn[k], n.c0mx[k] = box[ce[0]].mx[k];
                n.c1mn[k] = box[ce[1]].mn[k], n.c1mx[k] = box[ce[1]].mx[k];
            }
            n.child0 = ce[0] < cap ? ce[0] : left[ce[0]];
            n.child1 = ce[1] < cap ? ce[1] : left[ce[1]];
        }
        bvh.max_depth = height;
        bvh.sah_cost = bvh_sah_cost(bvh);
    }
    if (report) *report = rep;
    return rep.accepted ? rep.moves : 0;
}

void build_bvh_sah_host(const std::vector<Aabb> &bounds, HostBvh &out, int reinsert_rounds, float reinsert_accept) {
    build_bvh_sah_host_plain(bounds, out);
    if (reinsert_rounds < 0) reinsert_rounds = kReinsertDefaultRounds;
    if (!(reinsert_accept > 0)) reinsert_accept = kReinsertAccept;
    reinsert_optimize_host(out, reinsert_rounds, reinsert_accept);
}

// Padding: per axis  1e-4 * extent  +  4e-6 * (scene diagonal + largest |coordinate| of the box).
// The first term follows the box, the second keeps flat (zero-thickness) boxes and boxes far from
// the world origin a few dozen ulps thick, so that the fp32 slab test with FMA cannot reject a ray
// the exact primitive test accepts.
void pad_boxes(HostBvh &bvh, const std::vector<Aabb> &bounds) {
    Aabb scene = empty_box();
    for (auto &b: bounds) grow(scene, b);
    float diag = 0;
    if (!bounds.empty()) {
        double s = 0;
        for (int k = 0; k < 3; k++) s += (double) (scene.mx[k] - scene.mn[k]) * (scene.mx[k] - scene.mn[k]);
        diag = (float) std::sqrt(s);
    }
    auto pad = [&](float *mn, float *mx) {
        if (mn[0] > mx[0]) return;  // empty child
        for (int k = 0; k < 3; k++) {
            float ext = mx[k] - mn[k];
            float mag = std::max(std::fabs(mn[k]), std::fabs(mx[k]));
            float p = 1e-4f * ext + 4e-6f * (diag + mag);
            mn[k] -= p;
            mx[k] += p;
        }
    };
    for (auto &n: bvh.nodes) {
        pad(n.c0mn, n.c0mx);
        pad(n.c1mn, n.c1mx);
    }
}

// Re-lays the reachable nodes out in depth-first order (a node's first child right behind it) and drops the
// unreachable ones (the GPU builders leave the interior of collapsed subtrees behind): smaller array, and the
// nodes a ray touches next are the ones next in memory.
void compact_dfs(HostBvh &bvh, int bfs_top) {
    if (bvh.nodes.empty()) return;
    std::vector<HostNode> out;
    out.reserve(bvh.nodes.size());
    // (old index, slot in `out` whose child ref must be patched: -1 root, 2k+c)
    std::vector<std::pair<int, int>> todo;
    auto emit = [&](int old_idx, int patch) {
        const int me = (int) out.size();
        out.push_back(bvh.nodes[old_idx]);
        if (patch >= 0) {
            if (patch & 1) out[patch >> 1].child1 = me;
            else out[patch >> 1].child0 = me;
        }
        return me;
    };
    // optional breadth-first prefix (the top of the tree in the first `bfs_top` slots: shared-memory experiment)
    std::vector<std::pair<int, int>> queue(1, {0, -1});
    size_t head = 0;
    while (head < queue.size() && (int) out.size() < bfs_top) {
        auto [old_idx, patch] = queue[head++];
        const int me = emit(old_idx, patch);
        const HostNode &n = bvh.nodes[old_idx];
        if (n.child0 >= 0 && n.child0 != kEmptyChild) queue.emplace_back(n.child0, 2 * me);
        if (n.child1 >= 0 && n.child1 != kEmptyChild) queue.emplace_back(n.child1, 2 * me + 1);
    }
    for (size_t i = queue.size(); i-- > head;) todo.push_back(queue[i]);  // remaining subtrees, depth-first, in order
    while (!todo.empty()) {
        auto [old_idx, patch] = todo.back();
        todo.pop_back();
        const int me = emit(old_idx, patch);
        const HostNode &n = bvh.nodes[old_idx];
        // push child1 first so that child0's subtree is emitted right after this node
        if (n.child1 >= 0 && n.child1 != kEmptyChild) todo.emplace_back(n.child1, 2 * me + 1);
        if (n.child0 >= 0 && n.child0 != kEmptyChild) todo.emplace_back(n.child0, 2 * me);
    }
    bvh.nodes.swap(out);
}

float bvh_sah_cost(const HostBvh &bvh) {
    if (bvh.nodes.empty()) return 0;
    // root box = union of the root's children
    const HostNode &r = bvh.nodes[0];
    Aabb root = empty_box();
    Aabb a, b;
    for (int k = 0; k < 3; k++) a.mn[k] = r.c0mn[k], a.mx[k] = r.c0mx[k], b.mn[k] = r.c1mn[k], b.mx[k] = r.c1mx[k];
    grow(root, a);
    if (r.child1 != kEmptyChild) grow(root, b);
    const float ra = half_area(root);
    if (!(ra > 0)) return 0;
    double cost = kCostNode;  // the root step
    std::vector<int> todo(1, 0);  // reachable nodes only (the GPU builder leaves collapsed subtrees behind)
    while (!todo.empty()) {
        const HostNode &n = bvh.nodes[todo.back()];
        todo.pop_back();
        if (n.child0 >= 0 && n.child0 != kEmptyChild) todo.push_back(n.child0);
        if (n.child1 >= 0 && n.child1 != kEmptyChild) todo.push_back(n.child1);
        const float *mns[2] = {n.c0mn, n.c1mn}, *mxs[2] = {n.c0mx, n.c1mx};
        const int ch[2] = {n.child0, n.child1};
        for (int c = 0; c < 2; c++) {
            if (ch[c] == kEmptyChild) continue;
            Aabb cb;
            for (int k = 0; k < 3; k++) cb.mn[k] = mns[c][k], cb.mx[k] = mxs[c][k];
            float p = half_area(cb) / ra;
            if (ch[c] >= 0) cost += kCostNode * p;
            else cost += kCostPrim * p * (((~ch[c]) & 7) + 1);
        }
    }
    return (float) cost;
}

}  // namespace rtb
