// ref_order.cpp — exact-t tie breaking: ranks of the reference's primitive visit order.
//
// Why this exists: the traversal BVH of this library is not the reference's tree,
// and it does not have to be — any conservative tree finds the same hit SET.  What the tree does
// decide in the reference is which primitive wins when two hits have exactly equal float t (shared
// edges and vertices): its closest-hit loop keeps the first one visited (raytracer.cpp:202, 211), and
// its visit order is a DFS whose child order at every inner node depends only on the sign of the ray
// direction along the node's split axis (raytracer.cpp:190-196).  So the order of the leaves is one
// of 8 fixed permutations, chosen by the direction's sign octant.  Here the reference's tree
// (bvh.h:48-163: widest axis, spatial midpoint with up to 19 shrinking retries, depth cap 19,
// stable split, triangles before spheres in a leaf) is rebuilt with the same fp32 operations on one
// index array, and each primitive's position in each of the 8 leaf orders is written out.  The
// kernels compare (t, rank[octant][prim]) lexicographically.
#include <cfloat>
#include <cstring>

#include "rt_internal.h"

namespace rtb {

void primitive_bounds(const RtSceneDesc &d, std::vector<Aabb> &out) {
    const int nt = d.n_triangles, ns = d.n_spheres;
    out.resize((size_t) nt + ns);
    for (int i = 0; i < nt; i++) {
        const RtTriangle &t = d.triangles[i];
        const RtVec3 *v[3] = {&d.vertices[t.v0_id - 1], &d.vertices[t.v1_id - 1], &d.vertices[t.v2_id - 1]};
        Aabb b = {{FLT_MAX, FLT_MAX, FLT_MAX}, {-FLT_MAX, -FLT_MAX, -FLT_MAX}};
        for (int k = 0; k < 3; k++) {
            const float c[3] = {v[k]->x, v[k]->y, v[k]->z};
            for (int a = 0; a < 3; a++) {
                if (c[a] < b.mn[a]) b.mn[a] = c[a];
                if (c[a] > b.mx[a]) b.mx[a] = c[a];
            }
        }
        out[i] = b;
    }
    for (int i = 0; i < ns; i++) {
        const RtSphere &s = d.spheres[i];
        const RtVec3 &c = d.vertices[s.center_vertex_id - 1];
        const float cc[3] = {c.x, c.y, c.z};
        Aabb b;
        for (int a = 0; a < 3; a++) {
            b.mn[a] = cc[a] - s.radius;  // parser.h:307-311
            b.mx[a] = cc[a] + s.radius;
        }
        out[(size_t) nt + i] = b;
    }
}

namespace {

struct RefNode {
    int axis;
    int left, right;   // -1 for leaves
    int first, count;  // leaf range in the ordered id array
    float mn[3], mx[3];
};

struct RefBuilder {
    const std::vector<Aabb> &bounds;
    std::vector<float> key;  // [3][np] split keys: triangle centroid / sphere centre
    std::vector<int> ids, tmp;
    std::vector<RefNode> nodes;
    RefTreeStats stats;
    int np;

    explicit RefBuilder(const std::vector<Aabb> &b) : bounds(b), np((int) b.size()) {}

    int build(int lo, int hi, int depth) {
        const int me = (int) nodes.size();
        nodes.push_back(RefNode{0, -1, -1, lo, hi - lo, {0, 0, 0}, {0, 0, 0}});
        stats.nodes++;
        if (depth > stats.max_depth) stats.max_depth = depth;
        float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
        for (int i = lo; i < hi; i++) {
            const Aabb &b = bounds[ids[i]];
            for (int a = 0; a < 3; a++) {
                if (b.mn[a] < mn[a]) mn[a] = b.mn[a];
                if (b.mx[a] > mx[a]) mx[a] = b.mx[a];
            }
        }
        for (int a = 0; a < 3; a++) nodes[me].mn[a] = mn[a], nodes[me].mx[a] = mx[a];
        bool split = false;
        int n_left = 0;
        if (hi - lo > 1 && depth < 19) {            // bvh.h:57, MAX_DEPTH bvh.h:18
            int axis = 0;                           // parser.h:227-235, ties keep the lower axis
            for (int a = 1; a < 3; a++)
                if (mx[a] - mn[a] > mx[axis] - mn[axis]) axis = a;
            nodes[me].axis = axis;
            const float *k = &key[(size_t) axis * np];
            float start = mn[axis], end = mx[axis];
            float mid = (start + end) / 2;
            for (int tries = 19; tries > 0 && !split; tries--) {  // bvh.h:117-145
                n_left = 0;
                for (int i = lo; i < hi; i++) n_left += (k[ids[i]] < mid);
                if (n_left == 0) {
                    start = mid;
                    mid = (start + end) / 2;
                } else if (n_left == hi - lo) {
                    end = mid;
                    mid = (start + end) / 2;
                } else {
                    split = true;
                }
            }
            if (split) {  // stable split, bvh.h:146-159
                int l = lo, r = 0;
                for (int i = lo; i < hi; i++) {
                    if (k[ids[i]] < mid) ids[l++] = ids[i];
                    else tmp[r++] = ids[i];
                }
                memcpy(&ids[l], tmp.data(), sizeof(int) * (size_t) r);
            }
        }
        if (split) {
            int left = build(lo, lo + n_left, depth + 1);
            int right = build(lo + n_left, hi, depth + 1);
            nodes[me].left = left;
            nodes[me].right = right;
        } else {
            stats.leaves++;
            if (hi - lo > stats.max_leaf) stats.max_leaf = hi - lo;
        }
        return me;
    }
};

}  // namespace

void build_reference_ranks(const RtSceneDesc &d, std::vector<uint32_t> &ranks, RefTreeStats &stats, RefTree *tree) {
    std::vector<Aabb> bounds;
    primitive_bounds(d, bounds);
    RefBuilder b(bounds);
    const int np = b.np, nt = d.n_triangles;
    ranks.assign((size_t) 8 * np, 0u);
    stats = RefTreeStats();
    if (np == 0) return;
    b.key.resize((size_t) 3 * np);
    for (int i = 0; i < nt; i++) {
        const RtTriangle &t = d.triangles[i];
        const RtVec3 &p = d.vertices[t.v0_id - 1], &q = d.vertices[t.v1_id - 1], &r = d.vertices[t.v2_id - 1];
        // raytracer.cpp:347  center = (a + b + c) / 3
        b.key[i] = ((p.x + q.x) + r.x) / 3;
        b.key[(size_t) np + i] = ((p.y + q.y) + r.y) / 3;
        b.key[(size_t) 2 * np + i] = ((p.z + q.z) + r.z) / 3;
    }
    for (int i = 0; i < d.n_spheres; i++) {
        const RtVec3 &c = d.vertices[d.spheres[i].center_vertex_id - 1];  // bvh.h:131
        b.key[(size_t) nt + i] = c.x;
        b.key[(size_t) np + nt + i] = c.y;
        b.key[(size_t) 2 * np + nt + i] = c.z;
    }
    b.ids.resize(np);
    b.tmp.resize(np);
    for (int i = 0; i < np; i++) b.ids[i] = i;  // triangles first, then spheres: leaf order of raytracer.cpp:199-216
    b.build(0, np, 0);
    stats = b.stats;

    if (tree) {  // the reference's flattened tree itself (bvh.h:81-105 order: left child = index + 1)
        const int nn = (int) b.nodes.size();
        tree->nodes.resize((size_t) nn);
        tree->leaf_prims = b.ids;
        tree->leaf_of_prim.assign((size_t) np, 0);
        for (int i = 0; i < nn; i++) {
            const RefNode &n = b.nodes[i];
            RefTreeNode &o = tree->nodes[i];
            for (int a = 0; a < 3; a++) o.mn[a] = n.mn[a], o.mx[a] = n.mx[a];
            o.axis = n.axis;
            o.is_leaf = n.left < 0;
            o.right = n.right;
            o.first = n.first;
            o.count = n.count;
            if (n.left < 0)
                for (int k = 0; k < n.count; k++) tree->leaf_of_prim[b.ids[n.first + k]] = i;
        }
    }

    std::vector<int> stack;
    for (int oct = 0; oct < 8; oct++) {
        uint32_t next = 0;
        uint32_t *out = &ranks[(size_t) oct * np];
        stack.clear();
        stack.push_back(0);
        while (!stack.empty()) {
            const RefNode n = b.nodes[stack.back()];
            stack.pop_back();
            if (n.left < 0) {
                for (int i = 0; i < n.count; i++) out[b.ids[n.first + i]] = next++;
            } else if ((oct >> n.axis) & 1) {  // direction[axis] > 0: left child first (raytracer.cpp:190-192)
                stack.push_back(n.right);
                stack.push_back(n.left);
            } else {
                stack.push_back(n.left);
                stack.push_back(n.right);
            }
        }
    }
}

}  // namespace rtb
