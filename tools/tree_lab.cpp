// tree_lab.cpp — CPU laboratory for the TRAVERSAL tree's quality (no GPU needed).
//
// Builds candidate BVH2 trees for a scene (CPU restatements of the library's PLOC and binned-SAH builders, plus
// optimisation passes), then walks a sample of the rays the renderer would trace (primary, shadow, reflection of
// raytracer.cpp:385-452, plain float arithmetic — statistics, not parity) through each tree exactly the way
// render_v2.cu's loop does (two padded child boxes per step, near child first, any-hit exit) and counts node steps,
// leaf visits and primitive tests per ray kind.  The GPU kernel is issue-bound and 46 % of its instructions are node
// steps, so steps per ray IS its cost model; this tool ranks builder ideas before GPU time is
// spent on them.
//
//   g++ -O2 -std=c++17 -I include -I raytracer-ceng477-graphics-hw-1_b200/csrc -I raytracer-ceng477-graphics-hw-1_b200/csrc/host \
//       tools/tree_lab.cpp raytracer-ceng477-graphics-hw-1_b200/csrc/host/xml_scene.cpp \
//       raytracer-ceng477-graphics-hw-1_b200/csrc/bvh_host.cpp raytracer-ceng477-graphics-hw-1_b200/csrc/ref_order.cpp -o /tmp/tree_lab
//   /tmp/tree_lab scene.xml [width height]
#include <algorithm>
#include <cfloat>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <queue>
#include <string>
#include <vector>

#include "rt_internal.h"
#include "scene.h"

using namespace rtb;

namespace {

constexpr float kCostNode = 1.0f;
float kCostPrim = 1.6f;  // the library's constant; LAB_COST_PRIM overrides
int kLeafMax = kMaxLeafPrims;

struct V {
    double x, y, z;
};
V operator+(V a, V b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
V operator-(V a, V b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
V operator*(V a, double f) { return {a.x * f, a.y * f, a.z * f}; }
double dot(V a, V b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
V cross(V a, V b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
V norm(V a) { return a * (1.0 / std::sqrt(dot(a, a))); }

Aabb empty_box() { return Aabb{{FLT_MAX, FLT_MAX, FLT_MAX}, {-FLT_MAX, -FLT_MAX, -FLT_MAX}}; }
Aabb merge(const Aabb &a, const Aabb &b) {
    Aabb r;
    for (int k = 0; k < 3; k++) r.mn[k] = std::min(a.mn[k], b.mn[k]), r.mx[k] = std::max(a.mx[k], b.mx[k]);
    return r;
}
float harea(const Aabb &b) {
    float dx = b.mx[0] - b.mn[0], dy = b.mx[1] - b.mn[1], dz = b.mx[2] - b.mn[2];
    if (dx < 0 || dy < 0 || dz < 0) return 0;
    return dx * dy + dy * dz + dz * dx;
}

// ---- generic binary tree: ids 0..n-1 are the leaves (one primitive each), n..2n-2 inner ----------------------
struct BNode {
    Aabb box;
    int left = -1, right = -1, parent = -1;
};
struct BTree {
    int n = 0;
    std::vector<BNode> nd;
    int root = -1;
};

double sah_of(const BTree &t, std::vector<float> *cost_out = nullptr, std::vector<char> *col_out = nullptr) {
    // bottom-up cost with the leaf collapse the builders apply (<= 8 primitives, cheaper as a leaf); root stays inner
    std::vector<float> cost(t.nd.size(), 0.0f);
    std::vector<int> size(t.nd.size(), 0);
    std::vector<char> col(t.nd.size(), 0);
    std::vector<int> order;
    order.reserve(t.nd.size());
    std::vector<int> st(1, t.root);
    while (!st.empty()) {
        int x = st.back();
        st.pop_back();
        order.push_back(x);
        if (t.nd[x].left >= 0) st.push_back(t.nd[x].left), st.push_back(t.nd[x].right);
    }
    for (size_t i = order.size(); i-- > 0;) {
        int x = order[i];
        const BNode &b = t.nd[x];
        if (b.left < 0) {
            cost[x] = kCostPrim, size[x] = 1;
            continue;
        }
        size[x] = size[b.left] + size[b.right];
        float a = harea(b.box);
        float cs = kCostNode + (a > 0 ? (harea(t.nd[b.left].box) * cost[b.left] + harea(t.nd[b.right].box) * cost[b.right]) / a
                                      : cost[b.left] + cost[b.right]);
        float cl = kCostPrim * size[x];
        if (x != t.root && size[x] <= kLeafMax && cl <= cs) col[x] = 1, cs = cl;
        cost[x] = cs;
    }
    if (cost_out) *cost_out = cost;
    if (col_out) *col_out = col;
    return cost[t.root];
}

HostBvh to_host(const BTree &t) {
    std::vector<char> col;
    sah_of(t, nullptr, &col);
    HostBvh out;
    // DFS: inner nodes get indices in pre-order, leaves contiguous ranges of prim_order
    std::function<int(int, Aabb &, int)> rec = [&](int x, Aabb &box, int depth) -> int {
        const BNode &b = t.nd[x];
        box = b.box;
        out.max_depth = std::max(out.max_depth, depth);
        if (b.left < 0 || col[x]) {
            int first = (int) out.prim_order.size();
            std::vector<int> st(1, x);
            while (!st.empty()) {
                int y = st.back();
                st.pop_back();
                if (t.nd[y].left < 0) out.prim_order.push_back(y);
                else st.push_back(t.nd[y].right), st.push_back(t.nd[y].left);
            }
            int cnt = (int) out.prim_order.size() - first;
            return ~((first << 3) | (cnt - 1));
        }
        int me = (int) out.nodes.size();
        out.nodes.push_back(HostNode());
        Aabb b0, b1;
        int c0 = rec(b.left, b0, depth + 1), c1 = rec(b.right, b1, depth + 1);
        HostNode &h = out.nodes[me];
        for (int k = 0; k < 3; k++) h.c0mn[k] = b0.mn[k], h.c0mx[k] = b0.mx[k], h.c1mn[k] = b1.mn[k], h.c1mx[k] = b1.mx[k];
        h.child0 = c0, h.child1 = c1;
        return me;
    };
    Aabb rb;
    rec(t.root, rb, 0);
    out.sah_cost = bvh_sah_cost(out);
    return out;
}

// ---- PLOC as bvh_lbvh.cu runs it (30-bit Morton order, radius, mutual nearest neighbours) -----------------------
unsigned expand10(unsigned v) {
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}
BTree build_ploc(const std::vector<Aabb> &bounds, int radius) {
    const int n = (int) bounds.size();
    Aabb cb = empty_box();
    for (auto &b: bounds)
        for (int k = 0; k < 3; k++) {
            float c = 0.5f * (b.mn[k] + b.mx[k]);
            cb.mn[k] = std::min(cb.mn[k], c), cb.mx[k] = std::max(cb.mx[k], c);
        }
    std::vector<unsigned long long> keys(n);
    for (int i = 0; i < n; i++) {
        unsigned code = 0;
        for (int k = 0; k < 3; k++) {
            float lo = cb.mn[k], hi = cb.mx[k], c = 0.5f * (bounds[i].mn[k] + bounds[i].mx[k]);
            float x = hi > lo ? (c - lo) / (hi - lo) : 0.0f;
            x = std::min(std::max(x * 1024.0f, 0.0f), 1023.0f);
            code |= expand10((unsigned) x) << (2 - k);
        }
        keys[i] = ((unsigned long long) code << 32) | (unsigned) i;
    }
    std::sort(keys.begin(), keys.end());
    BTree t;
    t.n = n;
    t.nd.resize(2 * n - 1);
    for (int i = 0; i < n; i++) t.nd[i].box = bounds[i];
    std::vector<int> cin(n), cout, nn;
    for (int i = 0; i < n; i++) cin[i] = (int) (unsigned) keys[i];
    int alloc = n;
    while (cin.size() > 1) {
        const int m = (int) cin.size();
        nn.assign(m, -1);
        for (int i = 0; i < m; i++) {
            float best = FLT_MAX;
            int bj = -1;
            for (int j = std::max(0, i - radius); j <= std::min(m - 1, i + radius); j++) {
                if (j == i) continue;
                float a = harea(merge(t.nd[cin[i]].box, t.nd[cin[j]].box));
                if (a < best) best = a, bj = j;
            }
            nn[i] = bj;
        }
        cout.clear();
        for (int i = 0; i < m; i++) {
            int j = nn[i];
            bool mutual = j >= 0 && nn[j] == i;
            if (mutual && j < i) continue;
            int id = cin[i];
            if (mutual) {
                id = alloc++;
                BNode &b = t.nd[id];
                b.left = cin[i], b.right = cin[j];
                b.box = merge(t.nd[cin[i]].box, t.nd[cin[j]].box);
                t.nd[cin[i]].parent = id, t.nd[cin[j]].parent = id;
            }
            cout.push_back(id);
        }
        cin.swap(cout);
    }
    t.root = cin[0];
    return t;
}

// ---- top-down builders: 32-bin SAH on centroids (bvh_host.cpp) and the full sweep ------------------------------
struct TopDown {
    const std::vector<Aabb> &bounds;
    BTree &t;
    std::vector<int> ids;
    int alloc;
    bool sweep;
    TopDown(const std::vector<Aabb> &b, BTree &tr, bool sw) : bounds(b), t(tr), sweep(sw) {}
    int build(int lo, int hi) {
        const int n = hi - lo;
        if (n == 1) return ids[lo];
        Aabb box = empty_box(), cbox = empty_box();
        for (int i = lo; i < hi; i++) {
            const Aabb &b = bounds[ids[i]];
            box = merge(box, b);
            for (int k = 0; k < 3; k++) {
                float c = 0.5f * (b.mn[k] + b.mx[k]);
                cbox.mn[k] = std::min(cbox.mn[k], c), cbox.mx[k] = std::max(cbox.mx[k], c);
            }
        }
        int mid = -1;
        if (sweep) {
            float best = FLT_MAX;
            int best_axis = -1, best_i = -1;
            std::vector<float> ra(n);
            for (int axis = 0; axis < 3; axis++) {
                std::sort(ids.begin() + lo, ids.begin() + hi, [&](int a, int b) {
                    float ca = bounds[a].mn[axis] + bounds[a].mx[axis], cb2 = bounds[b].mn[axis] + bounds[b].mx[axis];
                    return ca < cb2 || (ca == cb2 && a < b);
                });
                Aabb acc = empty_box();
                for (int i = n - 1; i > 0; i--) acc = merge(acc, bounds[ids[lo + i]]), ra[i] = harea(acc);
                acc = empty_box();
                for (int i = 0; i < n - 1; i++) {
                    acc = merge(acc, bounds[ids[lo + i]]);
                    float c = harea(acc) * (i + 1) + ra[i + 1] * (n - i - 1);
                    if (c < best) best = c, best_axis = axis, best_i = i;
                }
            }
            std::sort(ids.begin() + lo, ids.begin() + hi, [&](int a, int b) {
                float ca = bounds[a].mn[best_axis] + bounds[a].mx[best_axis], cb2 = bounds[b].mn[best_axis] + bounds[b].mx[best_axis];
                return ca < cb2 || (ca == cb2 && a < b);
            });
            mid = lo + best_i + 1;
        } else {
            constexpr int kBins = 32;
            int best_axis = -1, best_bin = -1;
            float best_cost = FLT_MAX;
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
                    bb[b] = merge(bb[b], pb);
                }
                float right_area[kBins];
                Aabb acc = empty_box();
                for (int b = kBins - 1; b > 0; b--) acc = merge(acc, bb[b]), right_area[b] = harea(acc);
                acc = empty_box();
                int nl = 0;
                for (int b = 0; b < kBins - 1; b++) {
                    acc = merge(acc, bb[b]);
                    nl += cnt[b];
                    if (nl == 0 || nl == n) continue;
                    float cost = harea(acc) * nl + right_area[b + 1] * (n - nl);
                    if (cost < best_cost) best_cost = cost, best_axis = axis, best_bin = b;
                }
            }
            if (best_axis >= 0) {
                const float c0 = cbox.mn[best_axis], scale = kBins / (cbox.mx[best_axis] - c0);
                auto it = std::partition(ids.begin() + lo, ids.begin() + hi, [&](int id) {
                    const Aabb &pb = bounds[id];
                    int b = std::min(kBins - 1, std::max(0, (int) ((0.5f * (pb.mn[best_axis] + pb.mx[best_axis]) - c0) * scale)));
                    return b <= best_bin;
                });
                mid = (int) (it - ids.begin());
            }
        }
        if (mid <= lo || mid >= hi) mid = lo + n / 2;
        const int me = alloc++;
        int l = build(lo, mid), r = build(mid, hi);
        BNode &b = t.nd[me];
        b.left = l, b.right = r, b.box = box;
        t.nd[l].parent = me, t.nd[r].parent = me;
        return me;
    }
};
BTree build_topdown(const std::vector<Aabb> &bounds, bool sweep) {
    BTree t;
    t.n = (int) bounds.size();
    t.nd.resize(2 * t.n - 1);
    for (int i = 0; i < t.n; i++) t.nd[i].box = bounds[i];
    TopDown td(bounds, t, sweep);
    td.ids.resize(t.n);
    for (int i = 0; i < t.n; i++) td.ids[i] = i;
    td.alloc = t.n;
    t.root = td.build(0, t.n);
    return t;
}

// ---- insertion-based optimisation (Bittner, Hapala, Havran 2013): remove a subtree, re-insert it where the tree's
// total inner-node area grows least (branch and bound over the tree) --------------------------------------------
void refit_up(BTree &t, int x) {
    while (x >= 0) {
        BNode &b = t.nd[x];
        Aabb nb = merge(t.nd[b.left].box, t.nd[b.right].box);
        if (!memcmp(&nb, &b.box, sizeof nb)) break;
        b.box = nb;
        x = b.parent;
    }
}
int reinsert_pass(BTree &t, double frac, unsigned seed) {
    // candidates: all non-root nodes whose parent is not the root, largest parent-area first (they hurt most)
    std::vector<std::pair<float, int>> cand;
    for (int x = 0; x < (int) t.nd.size(); x++) {
        int p = t.nd[x].parent;
        if (p < 0 || t.nd[p].parent < 0) continue;
        cand.emplace_back(-harea(t.nd[p].box), x);
    }
    std::sort(cand.begin(), cand.end());
    int limit = (int) (cand.size() * frac), moved = 0;
    (void) seed;
    for (int ci = 0; ci < limit; ci++) {
        const int x = cand[ci].second;
        const int p = t.nd[x].parent;
        if (p < 0) continue;
        const int g = t.nd[p].parent;
        if (g < 0) continue;
        const int s = t.nd[p].left == x ? t.nd[p].right : t.nd[p].left;
        // remove x and p: s takes p's place
        const float area_before_p = harea(t.nd[p].box);
        if (t.nd[g].left == p) t.nd[g].left = s;
        else t.nd[g].right = s;
        t.nd[s].parent = g;
        // area released along the path (for the gain estimate): refit ancestors, remember old boxes to undo
        std::vector<std::pair<int, Aabb>> undo;
        double released = area_before_p;
        for (int a = g; a >= 0; a = t.nd[a].parent) {
            Aabb nb = merge(t.nd[t.nd[a].left].box, t.nd[t.nd[a].right].box);
            if (!memcmp(&nb, &t.nd[a].box, sizeof nb)) break;
            undo.emplace_back(a, t.nd[a].box);
            released += harea(t.nd[a].box) - harea(nb);
            t.nd[a].box = nb;
        }
        // best insertion position: minimise direct + induced area
        const Aabb xb = t.nd[x].box;
        const float xa = harea(xb);
        struct Item {
            float induced;
            int node;
            bool operator<(const Item &o) const { return induced > o.induced; }
        };
        std::priority_queue<Item> pq;
        pq.push({0.0f, t.root});
        float best = FLT_MAX;
        int best_node = -1;
        while (!pq.empty()) {
            Item it = pq.top();
            pq.pop();
            if (it.induced + xa >= best) break;
            const BNode &b = t.nd[it.node];
            const float direct = harea(merge(b.box, xb));
            const float total = it.induced + direct;
            if (total < best) best = total, best_node = it.node;
            const float ind = total - harea(b.box);
            if (b.left >= 0 && ind + xa < best) {
                pq.push({ind, b.left});
                pq.push({ind, b.right});
            }
        }
        // insert: p becomes the parent of (best_node, x) at best_node's place
        if ((double) best < released - 1e-7 * released && best_node != s) {
            const int y = best_node, yp = t.nd[y].parent;
            t.nd[p].left = y, t.nd[p].right = x;
            t.nd[p].parent = yp;
            if (yp < 0) t.root = p;
            else if (t.nd[yp].left == y) t.nd[yp].left = p;
            else t.nd[yp].right = p;
            t.nd[y].parent = p;
            t.nd[x].parent = p;
            t.nd[p].box = merge(t.nd[y].box, xb);
            for (int a = yp; a >= 0; a = t.nd[a].parent) t.nd[a].box = merge(t.nd[t.nd[a].left].box, t.nd[t.nd[a].right].box);
            moved++;
        } else {  // undo
            if (t.nd[g].left == s) t.nd[g].left = p;
            else t.nd[g].right = p;
            t.nd[s].parent = p;
            for (auto &u: undo) t.nd[u.first].box = u.second;
        }
    }
    return moved;
}

// ---- the same optimisation in the shape a GPU can run (Meister & Bittner 2018, "parallel reinsertion"): EVERY node
// searches its best new position on the unchanged tree (walk up the ancestors = "pivots", branch and bound in each
// pivot's other subtree, the ancestors' boxes shrunk as if the node were gone); the two paths node -> pivot <- target
// are locked with max(gain, id); moves that own all their locks are applied and refit their own paths.  One call =
// one round; the GPU version is three grid-wide phases per round.
struct Move {
    float gain;
    int x, y;
};
std::vector<char> g_frozen, g_inside;  // optional: collapsed subtrees move as one leaf (what the GPU builder's tree looks like)
long long g_visits = 0, g_visits_max = 0, g_searches = 0;
bool find_move(const BTree &t, int x, Move &mv) {
    long long visits = 0;
    struct VisitCount {
        long long &v;
        ~VisitCount() { g_visits += v, g_visits_max = std::max(g_visits_max, v), g_searches++; }
    } vc{visits};
    const int p = t.nd[x].parent;
    if (p < 0 || t.nd[p].parent < 0) return false;
    if (!g_inside.empty() && g_inside[x]) return false;
    const Aabb xb = t.nd[x].box;
    const float xa = harea(xb);
    float best = 0.0f;
    int best_y = -1;
    int stack_n[128];
    float stack_i[128];
    auto explore = [&](int r, float saved) {
        int sp = 0;
        stack_n[sp] = r, stack_i[sp] = 0.0f, sp++;
        while (sp > 0) {
            sp--;
            visits++;
            const int nodei = stack_n[sp];
            const float induced = stack_i[sp];
            if (saved - (induced + xa) <= best) continue;  // even a zero-growth position below cannot beat the best
            const BNode &b = t.nd[nodei];
            const float direct = harea(merge(b.box, xb));
            const float gain = saved - (induced + direct);
            if (gain > best) best = gain, best_y = nodei;
            const float ci = induced + direct - harea(b.box);
            if (b.left >= 0 && !(!g_frozen.empty() && g_frozen[nodei]) && saved - (ci + xa) > best && sp + 2 <= 128) {
                stack_n[sp] = b.left, stack_i[sp] = ci, sp++;
                stack_n[sp] = b.right, stack_i[sp] = ci, sp++;
            }
        }
    };
    const int s = t.nd[p].left == x ? t.nd[p].right : t.nd[p].left;
    float saved = harea(t.nd[p].box);
    Aabb shrunk = t.nd[s].box;
    explore(s, saved);
    int cur = p;
    for (;;) {
        const int a = t.nd[cur].parent;
        if (a < 0) break;
        const int u = t.nd[a].left == cur ? t.nd[a].right : t.nd[a].left;
        explore(u, saved);
        shrunk = merge(shrunk, t.nd[u].box);
        const float sa = harea(shrunk);
        if (t.nd[a].parent >= 0 && saved - sa > best) best = saved - sa, best_y = a;  // next to the shrunken ancestor itself
        saved += harea(t.nd[a].box) - sa;
        cur = a;
    }
    if (best_y < 0 || best_y == s) return false;
    mv = Move{best, x, best_y};
    return best > 1e-6f * harea(t.nd[t.root].box);
}
int parallel_round(BTree &t, double *gain_sum = nullptr) {
    const int N = (int) t.nd.size();
    std::vector<Move> moves;
    for (int x = 0; x < N; x++) {
        Move m;
        if (find_move(t, x, m)) moves.push_back(m);
    }
    // locks: max over (gain, x) on both paths up to (and including) the pivot = lowest common ancestor
    std::vector<unsigned long long> lock(N, 0ull);
    auto key = [](const Move &m) {
        unsigned g;
        memcpy(&g, &m.gain, 4);
        return ((unsigned long long) g << 32) | (unsigned) m.x;
    };
    std::vector<int> depth(N, 0);
    {
        std::vector<int> st(1, t.root);
        while (!st.empty()) {
            int x = st.back();
            st.pop_back();
            if (t.nd[x].left >= 0) {
                depth[t.nd[x].left] = depth[t.nd[x].right] = depth[x] + 1;
                st.push_back(t.nd[x].left), st.push_back(t.nd[x].right);
            }
        }
    }
    auto for_paths = [&](const Move &m, auto fn) {  // fn(node) -> bool continue
        int a = t.nd[m.x].parent, b = m.y;
        // m.y may be an ancestor of x (insertion next to a shrunken ancestor): then its path is part of a's
        while (a != b) {
            if (depth[a] >= depth[b]) {
                if (!fn(a)) return false;
                a = t.nd[a].parent;
            } else {
                if (!fn(b)) return false;
                b = t.nd[b].parent;
            }
        }
        if (!fn(a)) return false;  // the pivot (or y itself when y is an ancestor)
        if (a == m.y && t.nd[a].parent >= 0 && !fn(t.nd[a].parent)) return false;  // y's parent gets a new child
        return true;
    };
    for (auto &m: moves) {
        const unsigned long long k = key(m);
        for_paths(m, [&](int n) { lock[n] = std::max(lock[n], k); return true; });
        // y's parent is touched too when y is not an ancestor: it is on y's path already (y != pivot) — covered above
    }
    int applied = 0;
    double gs = 0;
    for (auto &m: moves) {
        const unsigned long long k = key(m);
        if (!for_paths(m, [&](int n) { return lock[n] == k; })) continue;
        const int x = m.x, p = t.nd[x].parent, g = t.nd[p].parent;
        const int s = t.nd[p].left == x ? t.nd[p].right : t.nd[p].left;
        if (t.nd[g].left == p) t.nd[g].left = s;
        else t.nd[g].right = s;
        t.nd[s].parent = g;
        for (int a = g; a >= 0; a = t.nd[a].parent) t.nd[a].box = merge(t.nd[t.nd[a].left].box, t.nd[t.nd[a].right].box);
        const int y = m.y, yp = t.nd[y].parent;
        t.nd[p].left = y, t.nd[p].right = x, t.nd[p].parent = yp;
        if (t.nd[yp].left == y) t.nd[yp].left = p;
        else t.nd[yp].right = p;
        t.nd[y].parent = p;
        for (int a = p; a >= 0; a = t.nd[a].parent) t.nd[a].box = merge(t.nd[t.nd[a].left].box, t.nd[t.nd[a].right].box);
        applied++;
        gs += m.gain;
    }
    if (gain_sum) *gain_sum = gs;
    return applied;
}
double inner_area(const BTree &t) {
    double a = 0;
    std::vector<int> st(1, t.root);
    while (!st.empty()) {
        int x = st.back();
        st.pop_back();
        if (t.nd[x].left >= 0) a += harea(t.nd[x].box), st.push_back(t.nd[x].left), st.push_back(t.nd[x].right);
    }
    return a;
}

// ---- ray simulation --------------------------------------------------------------------------------------
struct Tri {
    V a, b, c, n;
    int mat;
};
struct Stats {
    double rays = 0, steps = 0, leaves = 0, tests = 0;
};
struct Sim {
    const HostBvh &bvh;
    const std::vector<Tri> &tris;
    Stats st[3];  // primary, reflection, shadow
    Sim(const HostBvh &b, const std::vector<Tri> &t) : bvh(b), tris(t) {}

    static bool tri_hit(const Tri &T, V o, V d, double &t) {
        V e1 = T.b - T.a, e2 = T.c - T.a, pv = cross(d, e2);
        double det = dot(e1, pv);
        if (det == 0) return false;
        double inv = 1.0 / det;
        V tv = o - T.a;
        double u = dot(tv, pv) * inv;
        if (u < 0 || u > 1) return false;
        V qv = cross(tv, e1);
        double v = dot(d, qv) * inv;
        if (v < 0 || u + v > 1) return false;
        t = dot(e2, qv) * inv;
        return t >= 0;
    }
    static void slab(const float *mn, const float *mx, const float o[3], const float inv[3], float &tmin, float &tmax) {
        tmin = -FLT_MAX, tmax = FLT_MAX;
        for (int k = 0; k < 3; k++) {
            float t0 = (mn[k] - o[k]) * inv[k], t1 = (mx[k] - o[k]) * inv[k];
            tmin = std::max(tmin, std::min(t0, t1));
            tmax = std::min(tmax, std::max(t0, t1));
        }
    }
    // returns prim or -1; any: first hit with t < limit
    int trace(V o, V d, bool any, double limit, double &tbest, Stats &s) {
        s.rays++;
        float of[3] = {(float) o.x, (float) o.y, (float) o.z}, inv[3];
        const double dd[3] = {d.x, d.y, d.z};
        for (int k = 0; k < 3; k++) inv[k] = std::min(std::max((float) (1.0 / dd[k]), -1e18f), 1e18f);
        tbest = limit;
        int pbest = -1;
        int stack[128], sp = 0;
        int node = 0;
        stack[sp++] = INT32_MAX;
        while (node != INT32_MAX) {
            if (node >= 0) {
                s.steps++;
                const HostNode &n = bvh.nodes[node];
                float a0, b0, a1, b1;
                slab(n.c0mn, n.c0mx, of, inv, a0, b0);
                slab(n.c1mn, n.c1mx, of, inv, a1, b1);
                bool h0 = b0 >= std::max(a0, 0.0f) && a0 <= (float) tbest, h1 = b1 >= std::max(a1, 0.0f) && a1 <= (float) tbest;
                if (n.child1 == kEmptyChild) h1 = false;
                bool swap = a1 < a0;
                if (h0 && h1) {
                    stack[sp++] = swap ? n.child0 : n.child1;
                    node = swap ? n.child1 : n.child0;
                } else if (h0) node = n.child0;
                else if (h1) node = n.child1;
                else node = stack[--sp];
            } else {
                s.leaves++;
                int enc = ~node, first = enc >> 3, count = (enc & 7) + 1;
                node = stack[--sp];
                for (int i = first; i < first + count; i++) {
                    s.tests++;
                    double t;
                    int prim = bvh.prim_order[i];
                    if (tri_hit(tris[prim], o, d, t)) {
                        if (any) {
                            if (t < limit) {
                                tbest = t;
                                return prim;
                            }
                        } else if (pbest < 0 || t < tbest) tbest = t, pbest = prim;
                    }
                }
            }
        }
        return pbest;
    }
};

}  // namespace

int main(int argc, char **argv) {
    if (argc < 2) {
        fprintf(stderr, "usage: tree_lab scene.xml [width height]\n");
        return 2;
    }
    if (getenv("LAB_COST_PRIM")) kCostPrim = (float) atof(getenv("LAB_COST_PRIM"));
    if (getenv("LAB_LEAF_MAX")) kLeafMax = atoi(getenv("LAB_LEAF_MAX"));
    parser::Scene scene;
    scene.loadFromXml(argv[1]);
    parser::FlatScene flat;
    parser::flatten(scene, flat);
    const RtSceneDesc &d = flat.desc;
    std::vector<Aabb> bounds;
    primitive_bounds(d, bounds);
    const int nt = d.n_triangles;
    if (d.n_spheres) fprintf(stderr, "note: %d spheres ignored by the ray simulation (boxes only)\n", d.n_spheres);
    std::vector<Tri> tris(bounds.size());
    for (int i = 0; i < nt; i++) {
        auto vv = [&](int id) { const RtVec3 &v = d.vertices[id - 1]; return V{v.x, v.y, v.z}; };
        Tri &T = tris[i];
        T.a = vv(d.triangles[i].v0_id), T.b = vv(d.triangles[i].v1_id), T.c = vv(d.triangles[i].v2_id);
        V n = cross(T.b - T.a, T.c - T.a);
        double l = std::sqrt(dot(n, n));
        T.n = l > 0 ? n * (1.0 / l) : V{0, 1, 0};
        T.mat = d.triangles[i].material_id;
    }
    for (size_t i = nt; i < tris.size(); i++) tris[i] = Tri{{1e30, 1e30, 1e30}, {1e30, 1e30, 1e30}, {1e30, 1e30, 1e30}, {0, 1, 0}, 1};
    const parser::Camera &cam = scene.cameras[0];
    const int W = argc > 3 ? atoi(argv[2]) : 480, H = argc > 3 ? atoi(argv[3]) : 240;

    auto evaluate_host = [&](const char *name, HostBvh bvh) {
        pad_boxes(bvh, bounds);
        Sim sim(bvh, tris);
        // camera basis (raytracer.cpp:292-316)
        V e{cam.position.x, cam.position.y, cam.position.z}, g{cam.gaze.x, cam.gaze.y, cam.gaze.z}, up{cam.up.x, cam.up.y, cam.up.z};
        V w = norm(g) * -1.0, u = norm(cross(up, w)), v = cross(w, u);
        double l = cam.near_plane.x, r = cam.near_plane.y, b = cam.near_plane.z, tp = cam.near_plane.w;
        V m = e + w * (-(double) cam.near_distance), q = m + u * l + v * tp;
        double occluded = 0;
        std::function<void(V, V, int)> path = [&](V o, V dir, int depth) {
            if (depth > d.max_recursion_depth) return;
            double t;
            int prim = sim.trace(o, dir, false, DBL_MAX, t, sim.st[depth ? 1 : 0]);
            if (prim < 0) return;
            const Tri &T = tris[prim];
            V P = o + dir * t, Pe = P + T.n * (double) d.shadow_ray_epsilon;
            for (int li = 0; li < d.n_lights; li++) {
                V lp{d.lights[li].position.x, d.lights[li].position.y, d.lights[li].position.z};
                V toL = lp - Pe;
                double dist = std::sqrt(dot(toL, toL)), tt;
                if (sim.trace(Pe, toL * (1.0 / dist), true, dist, tt, sim.st[2]) >= 0) occluded++;
            }
            if (d.materials[T.mat - 1].is_mirror) {
                V dn = norm(dir);
                path(Pe, dn + T.n * (2.0 * dot(dn * -1.0, T.n)), depth + 1);
            }
        };
        auto t0 = std::chrono::steady_clock::now();
        for (int y = 0; y < H; y++)
            for (int x = 0; x < W; x++) {
                double su = (x + 0.5) * (r - l) / W, sv = (y + 0.5) * (tp - b) / H;
                V s = q + u * su - v * sv;
                path(e, s - e, 0);
            }
        double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        Stats all;
        for (int k = 0; k < 3; k++) all.rays += sim.st[k].rays, all.steps += sim.st[k].steps, all.leaves += sim.st[k].leaves, all.tests += sim.st[k].tests;
        // kernel cost model (instructions per ray): 47 per node step, ~14 per leaf visit, ~40 per primitive test
        auto model = [](const Stats &s) { return (47 * s.steps + 14 * s.leaves + 40 * s.tests) / s.rays; };
        printf("%-28s sah %6.3f nodes %6zu depth %2d | steps/ray %6.3f leaves %5.3f tests %5.3f model %6.1f |", name, bvh.sah_cost,
               bvh.nodes.size(), bvh.max_depth, all.steps / all.rays, all.leaves / all.rays, all.tests / all.rays, model(all));
        const char *kn[3] = {"pri", "refl", "shad"};
        for (int k = 0; k < 3; k++)
            printf(" %s %.3f/%.3f/%.3f", kn[k], sim.st[k].steps / sim.st[k].rays, sim.st[k].leaves / sim.st[k].rays, sim.st[k].tests / sim.st[k].rays);
        printf(" | refl/pri %.3f shad/pri %.3f occl %.3f (%.1fs)\n", sim.st[1].rays / sim.st[0].rays, sim.st[2].rays / sim.st[0].rays,
               occluded / sim.st[2].rays, secs);
        fflush(stdout);
    };

    auto evaluate = [&](const char *name, const BTree &t) { evaluate_host(name, to_host(t)); };
    auto dump_top = [&](const BTree &t, int levels) {
        std::function<int(int)> size = [&](int x) -> int { return t.nd[x].left < 0 ? 1 : size(t.nd[x].left) + size(t.nd[x].right); };
        std::function<void(int, int)> rec = [&](int x, int lv) {
            const Aabb &b = t.nd[x].box;
            fprintf(stderr, "%*s[%d prims] x %.2f..%.2f y %.2f..%.2f z %.2f..%.2f area %.1f\n", 2 * lv, "", size(x), b.mn[0], b.mx[0], b.mn[1], b.mx[1], b.mn[2], b.mx[2], harea(b));
            if (lv < levels && t.nd[x].left >= 0) rec(t.nd[x].left, lv + 1), rec(t.nd[x].right, lv + 1);
        };
        rec(t.root, 0);
    };
    auto timed = [&](const char *what, auto fn) {
        auto t0 = std::chrono::steady_clock::now();
        auto r = fn();
        fprintf(stderr, "[%s: %.1f ms]\n", what, 1e3 * std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
        return r;
    };
    BTree ploc = timed("ploc r16", [&] { return build_ploc(bounds, 16); });
    evaluate("ploc r16", ploc);
    BTree binned = timed("binned sah", [&] { return build_topdown(bounds, false); });
    evaluate("binned sah 32", binned);
    BTree sweep = timed("sweep sah", [&] { return build_topdown(bounds, true); });
    evaluate("sweep sah", sweep);
    if (getenv("DUMP")) dump_top(ploc, 4);
    for (BTree *base: {&ploc, &sweep, &binned}) {
        BTree t = *base;
        for (int pass = 1; pass <= 8; pass++) {
            int moved = timed("reinsert pass", [&] { return reinsert_pass(t, 1.0, pass); });
            char nm[64];
            snprintf(nm, sizeof nm, "%s + reinsert x%d (%d)", base == &ploc ? "ploc" : base == &sweep ? "sweep" : "binned", pass, moved);
            if (pass == 2 || pass == 8) evaluate(nm, t);
        }
        if (getenv("DUMP")) dump_top(t, 4);
    }
    {   // the library's own host implementation (bvh_host.cpp + reinsert_core.h): what RT_BUILD_SAH_HOST uploads
        HostBvh lib;
        build_bvh_sah_host(bounds, lib);
        evaluate_host("library host builder", lib);
        for (int rounds: {1, 4, 8, 16}) {
            HostBvh h;
            build_bvh_sah_host_plain(bounds, h);
            ReinsertReport rep;
            auto t0 = std::chrono::steady_clock::now();
            reinsert_optimize_host(h, rounds, 10.0f, &rep);
            double ms = 1e3 * std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            char nm[96];
            snprintf(nm, sizeof nm, "library reinsert x%d (%d)", rounds, rep.moves);
            fprintf(stderr, "[%s: %.1f ms, cost %.3f -> %.3f, height %d, accepted %d]\n", nm, ms, rep.cost_before, rep.cost_after, rep.height, (int) rep.accepted);
            evaluate_host(nm, h);
        }
    }
    if (getenv("LAB_PROBE")) {
        // Could the builder CHOOSE between candidate trees by measurement instead of by SAH cost, without knowing the camera?
        // Probe rays a scene offers at creation time: from random surface points (area-weighted) to every light (what shadow
        // rays are), one random direction off each point (reflections), and from a sphere around the scene to random surface
        // points (primary rays of some camera).  Below: node steps per probe ray next to the steps per ray of the scene's own
        // camera — if the two rank the candidates alike, a few thousand probe rays on the device can replace the 0.8x rule.
        auto probe = [&](HostBvh bvh, double *per_kind) {
            pad_boxes(bvh, bounds);
            Sim sim(bvh, tris);
            std::vector<double> cdf(nt);
            double acc = 0;
            for (int i = 0; i < nt; i++) {
                V c = cross(tris[i].b - tris[i].a, tris[i].c - tris[i].a);
                acc += 0.5 * std::sqrt(dot(c, c));
                cdf[i] = acc;
            }
            unsigned long long rng = 0x9E3779B97F4A7C15ull;
            auto uni = [&]() { rng ^= rng << 13, rng ^= rng >> 7, rng ^= rng << 17; return (double) (rng >> 11) / 9007199254740992.0; };
            Aabb sb = empty_box();
            for (auto &b: bounds) sb = merge(sb, b);
            V ctr{0.5 * (sb.mn[0] + sb.mx[0]), 0.5 * (sb.mn[1] + sb.mx[1]), 0.5 * (sb.mn[2] + sb.mx[2])};
            double rad = 0;
            for (int k = 0; k < 3; k++) rad += (double) (sb.mx[k] - sb.mn[k]) * (sb.mx[k] - sb.mn[k]);
            rad = std::sqrt(rad);  // sphere of twice the scene's radius
            Stats st[3];
            for (int it = 0; it < 4000 && nt > 0; it++) {
                int i = (int) (std::lower_bound(cdf.begin(), cdf.end(), uni() * acc) - cdf.begin());
                if (i >= nt) i = nt - 1;
                double u = uni(), v = uni();
                if (u + v > 1) u = 1 - u, v = 1 - v;
                const Tri &T = tris[i];
                V P = T.a + (T.b - T.a) * u + (T.c - T.a) * v, Pe = P + T.n * (double) d.shadow_ray_epsilon;
                double t;
                for (int li = 0; li < d.n_lights; li++) {
                    V lp{d.lights[li].position.x, d.lights[li].position.y, d.lights[li].position.z}, toL = lp - Pe;
                    double dist = std::sqrt(dot(toL, toL));
                    sim.trace(Pe, toL * (1.0 / dist), true, dist, t, st[2]);
                }
                V dir{uni() * 2 - 1, uni() * 2 - 1, uni() * 2 - 1};
                if (dot(dir, T.n) < 0) dir = dir * -1.0;
                sim.trace(Pe, norm(dir), false, DBL_MAX, t, st[1]);
                V dd = norm(V{uni() * 2 - 1, uni() * 2 - 1, uni() * 2 - 1});
                V eye = ctr + dd * rad;
                sim.trace(eye, P - eye, false, DBL_MAX, t, st[0]);
            }
            double all_steps = 0, all_rays = 0;
            for (int k = 0; k < 3; k++) per_kind[k] = st[k].steps / std::max(1.0, st[k].rays), all_steps += st[k].steps, all_rays += st[k].rays;
            return all_steps / std::max(1.0, all_rays);
        };
        HostBvh plain, opt, pl = to_host(build_ploc(bounds, 16));
        build_bvh_sah_host_plain(bounds, plain);
        build_bvh_sah_host_plain(bounds, opt);
        reinsert_optimize_host(opt, 8, 1e9f);
        const char *names[3] = {"top-down as built", "top-down + 8 rounds", "ploc r16"};
        HostBvh *cands[3] = {&plain, &opt, &pl};
        for (int c = 0; c < 3; c++) {
            double k[3];
            double all = probe(*cands[c], k);
            printf("probe rays  %-22s sah %6.3f | steps/ray %6.3f  (outside-in %.3f, off-surface %.3f, to-lights %.3f)\n", names[c], cands[c]->sah_cost, all, k[0], k[1], k[2]);
            evaluate_host(names[c], *cands[c]);
        }
        return 0;
    }
    if (getenv("LAB_LIBRARY_ONLY")) return 0;
    for (int variant = 0; variant < 3; variant++) {
        BTree *base = variant == 0 ? &ploc : &binned;
        BTree t = *base;
        g_frozen.clear(), g_inside.clear();
        if (variant == 2) {  // leaves of up to 8 primitives are atomic
            sah_of(t, nullptr, &g_frozen);
            g_inside.assign(t.nd.size(), 0);
            std::vector<int> st(1, t.root);
            while (!st.empty()) {
                int x = st.back();
                st.pop_back();
                if (t.nd[x].left < 0) continue;
                for (int c: {t.nd[x].left, t.nd[x].right}) g_inside[c] = g_inside[x] || g_frozen[x], st.push_back(c);
            }
        }
        for (int round = 1; round <= 16; round++) {
            double gs = 0, before = inner_area(t);
            int applied = timed("parallel round", [&] { return parallel_round(t, &gs); });
            double after = inner_area(t);
            fprintf(stderr, "round %d: %d moves, predicted gain %.3f, area %.3f -> %.3f; search visits mean %.1f max %lld\n", round, applied, gs, before, after, (double) g_visits / g_searches, g_visits_max);
            g_visits = g_visits_max = g_searches = 0;
            char nm[64];
            snprintf(nm, sizeof nm, "%s + parallel x%d (%d)", variant == 0 ? "ploc" : variant == 1 ? "binned" : "binned/leaves", round, applied);
            if (round == 1 || round == 2 || round == 4 || round == 8 || round == 16 || round == 32) evaluate(nm, t);
        }
    }
    return 0;
}
