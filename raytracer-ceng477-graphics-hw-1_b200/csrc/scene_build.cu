// scene_build.cu — orchestration of the GPU-resident scene build (build_device.h): one persistent arena for the
// buffers the render kernel reads, one scratch arena for the builders, every step a kernel on one stream, one
// 128-byte result block read back at the end.  Replaces RayTracer::RayTracer (raytracer.cpp:335-350: triangle
// list, normals, centroids) and BVHTree::build (bvh.h:48-178).
//
// Arithmetic: compiled with -fmad=false like the render kernel.  The per-primitive pre-pass reproduces the
// reference's fp32 operations (vertex min/max bounds parser.h:272-317, centroid ((a+b)+c)/3 raytracer.cpp:347,
// normal ((b-a)x(c-a)).normalize() raytracer.cpp:346) — the same values the host-side staging of round 1 produced
// with `volatile float`: every golden frame is byte-identical through them, and the reference-order tree built from
// these bounds and keys is hash-compared with the host build (test_reference_tree_on_gpu_equals_host).
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "build_device.h"
#include "scene_build.h"

namespace rtb {

namespace {

// ---------------------------------------------------------------------------------------------------------------
// per-primitive pre-pass
// ---------------------------------------------------------------------------------------------------------------

struct PrepArgs {
    const RtVec3 *vertices;
    const RtTriangle *triangles;
    const RtSphere *spheres;
    int nt, ns;
    Aabb *bounds;          // [np]
    float *key;            // [3][np]
    float4 *rec;           // [3 * np] primitive records by id (placed into leaf order later)
    float4 *prim_bounds;   // [2 * np]
    float4 *tri_nm, *tri_nn;
    float4 *sph_cr;
    int *sph_mat;
    BuildResult *result;
};

__global__ void init_result_kernel(BuildResult *r) {
    if (threadIdx.x == 0) {
        memset(r, 0, sizeof *r);
        for (int k = 0; k < 3; k++) {
            r->scene_bounds[k] = 0xffffffffu, r->scene_bounds[3 + k] = 0u;
            r->centroid_bounds[k] = 0xffffffffu, r->centroid_bounds[3 + k] = 0u;
        }
    }
}

__device__ __forceinline__ float3 ld_vertex(const RtVec3 *v, int id) {
    const RtVec3 p = v[id - 1];
    return make_float3(p.x, p.y, p.z);
}

__global__ void prep_kernel(PrepArgs a) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int np = a.nt + a.ns;
    Aabb b = {{FLT_MAX, FLT_MAX, FLT_MAX}, {-FLT_MAX, -FLT_MAX, -FLT_MAX}};
    const bool live = i < np;
    if (live && i < a.nt) {
        const RtTriangle t = a.triangles[i];
        RT_CHECK(t.v0_id >= 1 && t.v1_id >= 1 && t.v2_id >= 1);
        const float3 p = ld_vertex(a.vertices, t.v0_id), q = ld_vertex(a.vertices, t.v1_id), r = ld_vertex(a.vertices, t.v2_id);
        const float vx[3] = {p.x, q.x, r.x}, vy[3] = {p.y, q.y, r.y}, vz[3] = {p.z, q.z, r.z};
        for (int k = 0; k < 3; k++) {  // parser.h:272-296
            if (vx[k] < b.mn[0]) b.mn[0] = vx[k];
            if (vx[k] > b.mx[0]) b.mx[0] = vx[k];
            if (vy[k] < b.mn[1]) b.mn[1] = vy[k];
            if (vy[k] > b.mx[1]) b.mx[1] = vy[k];
            if (vz[k] < b.mn[2]) b.mn[2] = vz[k];
            if (vz[k] > b.mx[2]) b.mx[2] = vz[k];
        }
        a.key[i] = ((p.x + q.x) + r.x) / 3;  // raytracer.cpp:347
        a.key[(size_t) np + i] = ((p.y + q.y) + r.y) / 3;
        a.key[(size_t) 2 * np + i] = ((p.z + q.z) + r.z) / 3;
        // raytracer.cpp:135-138: a - b and a - c, the same fp32 subtractions done once; det()'s ray-independent minor
        const float abx = p.x - q.x, aby = p.y - q.y, abz = p.z - q.z;
        const float acx = p.x - r.x, acy = p.y - r.y, acz = p.z - r.z;
        const float mnr = aby * acz - acy * abz;
        a.rec[3 * (size_t) i + 0] = make_float4(p.x, p.y, p.z, __int_as_float(i));
        a.rec[3 * (size_t) i + 1] = make_float4(abx, aby, abz, __int_as_float(0));
        a.rec[3 * (size_t) i + 2] = make_float4(acx, acy, acz, mnr);
        // raytracer.cpp:346  ((b - a) x (c - a)).normalize()
        const float bax = q.x - p.x, bay = q.y - p.y, baz = q.z - p.z;
        const float cax = r.x - p.x, cay = r.y - p.y, caz = r.z - p.z;
        const float nx = bay * caz - baz * cay, ny = baz * cax - bax * caz, nz = bax * cay - bay * cax;
        const float len = sqrtf((nx * nx + ny * ny) + nz * nz);
        const float ux = nx / len, uy = ny / len, uz = nz / len;
        a.tri_nm[i] = make_float4(ux, uy, uz, __int_as_float(t.material_id));
        // intersection.normal.normalize() of an already unit-length normal (raytracer.cpp:414, :432): same ops, once
        const float ulen = sqrtf((ux * ux + uy * uy) + uz * uz);
        a.tri_nn[i] = make_float4(ux / ulen, uy / ulen, uz / ulen, 0.f);
    } else if (live) {
        const int s = i - a.nt;
        const RtSphere sp = a.spheres[s];
        const float3 c = ld_vertex(a.vertices, sp.center_vertex_id);
        b.mn[0] = c.x - sp.radius, b.mn[1] = c.y - sp.radius, b.mn[2] = c.z - sp.radius;  // parser.h:307-311
        b.mx[0] = c.x + sp.radius, b.mx[1] = c.y + sp.radius, b.mx[2] = c.z + sp.radius;
        a.key[i] = c.x;  // bvh.h:131
        a.key[(size_t) np + i] = c.y;
        a.key[(size_t) 2 * np + i] = c.z;
        a.rec[3 * (size_t) i + 0] = make_float4(c.x, c.y, c.z, __int_as_float(i));
        a.rec[3 * (size_t) i + 1] = make_float4(sp.radius, 0.f, 0.f, __int_as_float(1));
        a.rec[3 * (size_t) i + 2] = make_float4(0.f, 0.f, 0.f, 0.f);
        a.sph_cr[s] = make_float4(c.x, c.y, c.z, sp.radius);
        a.sph_mat[s] = sp.material_id;
    }
    if (live) {
        a.bounds[i] = b;
        a.prim_bounds[2 * (size_t) i] = make_float4(b.mn[0], b.mn[1], b.mn[2], 0.f);
        a.prim_bounds[2 * (size_t) i + 1] = make_float4(b.mx[0], b.mx[1], b.mx[2], 0.f);
    }
    // scene bounds and centroid bounds: warp reduction, then one atomic per warp and component
    for (int k = 0; k < 3; k++) {
        const unsigned lo = live ? f2ord(b.mn[k]) : 0xffffffffu, hi = live ? f2ord(b.mx[k]) : 0u;
        const unsigned c = live ? f2ord(0.5f * (b.mn[k] + b.mx[k])) : 0u;
        const unsigned wlo = __reduce_min_sync(0xffffffffu, lo), whi = __reduce_max_sync(0xffffffffu, hi);
        const unsigned clo = __reduce_min_sync(0xffffffffu, live ? c : 0xffffffffu), chi = __reduce_max_sync(0xffffffffu, c);
        if ((threadIdx.x & 31) == 0 && wlo <= whi) {
            atomicMin(&a.result->scene_bounds[k], wlo);
            atomicMax(&a.result->scene_bounds[3 + k], whi);
            atomicMin(&a.result->centroid_bounds[k], clo);
            atomicMax(&a.result->centroid_bounds[3 + k], chi);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// candidate-tree finalisation
// ---------------------------------------------------------------------------------------------------------------

__global__ void single_prim_tree_kernel(const Aabb *bounds, DevTree t) {
    HostNode nd;
    for (int k = 0; k < 3; k++) {
        nd.c0mn[k] = bounds[0].mn[k], nd.c0mx[k] = bounds[0].mx[k];
        nd.c1mn[k] = FLT_MAX, nd.c1mx[k] = -FLT_MAX;
    }
    nd.child0 = ~0;
    nd.child1 = kEmptyChild;
    t.nodes[0] = nd;
    t.prim_order[0] = 0;
    *t.n_used = 1;
    *t.status = 0;
}

__global__ void set_int_kernel(int *p, int v) { *p = v; }

__device__ __forceinline__ bool is_inner(int ref) { return ref >= 0 && ref != kEmptyChild; }

__global__ void tree_init_kernel(DevTree t) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= t.cap) return;
    t.parent[i] = -1;
    t.arrivals[i] = 0;
}

__global__ void tree_parents_kernel(DevTree t) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (*t.status != 0 || i >= *t.n_used) return;
    const int c0 = t.nodes[i].child0, c1 = t.nodes[i].child1;
    RT_CHECK((!is_inner(c0) || c0 < t.cap) && (!is_inner(c1) || c1 < t.cap));
    if (is_inner(c0)) t.parent[c0] = (i << 1);
    if (is_inner(c1)) t.parent[c1] = (i << 1) | 1;
}

__device__ __forceinline__ float half_area_of(const float *mn, const float *mx) {
    const float dx = mx[0] - mn[0], dy = mx[1] - mn[1], dz = mx[2] - mn[2];
    if (dx < 0 || dy < 0 || dz < 0) return 0;
    return dx * dy + dy * dz + dz * dx;
}

// bottom-up: a node is complete once both children have reported; leaf (and missing) children report at once, an
// inner child reports when it is complete itself; whoever completes a node carries on to its parent
__global__ void tree_reduce_kernel(DevTree t) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (*t.status != 0 || i >= *t.n_used) return;
    int cur = i;
    {
        const int c0 = t.nodes[i].child0, c1 = t.nodes[i].child1;
        const int leaf_reports = (is_inner(c0) ? 0 : 1) + (is_inner(c1) ? 0 : 1);
        if (leaf_reports == 0) return;
        if (leaf_reports == 1 && atomicAdd(&t.arrivals[i], 1) != 1) return;
    }
    for (int guard = 0; guard < (1 << 24); guard++) {
        __threadfence();
        const HostNode nd = t.nodes[cur];
        int count = 1, height = 1;
        float area = 0;
        const int ch[2] = {nd.child0, nd.child1};
        const float *mns[2] = {nd.c0mn, nd.c1mn}, *mxs[2] = {nd.c0mx, nd.c1mx};
        for (int c = 0; c < 2; c++) {
            if (ch[c] == kEmptyChild) continue;
            const float ha = half_area_of(mns[c], mxs[c]);
            if (ch[c] >= 0) {
                count += __ldcg(&t.count[ch[c]]);
                height = max(height, 1 + __ldcg(&t.height[ch[c]]));
                area += kSahCostNode * ha + __ldcg(&t.area[ch[c]]);
            } else {
                area += kSahCostPrim * ha * (float) (((~ch[c]) & 7) + 1);
            }
        }
        t.count[cur] = count;
        t.height[cur] = height;
        t.area[cur] = area;
        __threadfence();
        const int pp = t.parent[cur];
        if (pp < 0) return;
        cur = pp >> 1;
        RT_CHECK(cur >= 0 && cur < t.cap);
        if (atomicAdd(&t.arrivals[cur], 1) != 1) return;  // the sibling subtree is not finished yet
    }
}

__global__ void tree_root_kernel(DevTree t, BuildResult *res, int slot) {
    res->status[slot] = *t.status;
    if (*t.status != 0) {
        res->n_nodes[slot] = 0, res->height[slot] = 0, res->sah_cost[slot] = 0;
        return;
    }
    const HostNode r = t.nodes[0];
    float mn[3], mx[3];
    for (int k = 0; k < 3; k++) {
        mn[k] = r.c0mn[k], mx[k] = r.c0mx[k];
        if (r.child1 != kEmptyChild) mn[k] = fminf(mn[k], r.c1mn[k]), mx[k] = fmaxf(mx[k], r.c1mx[k]);
    }
    const float ra = half_area_of(mn, mx);
    res->n_nodes[slot] = t.count[0];
    res->height[slot] = t.height[0];
    res->sah_cost[slot] = ra > 0 ? kSahCostNode + t.area[0] / ra : 0.0f;
}

// RT_BUILD_AUTO: keep the PLOC tree (candidate 0) when its SAH cost is clearly lower than the top-down binned-SAH
// tree's (< 0.8x: scenes with huge primitives next to dense meshes, e.g. horse_and_mug 4.4 vs 7.2); otherwise the
// shallower top-down tree traverses 4-12 % faster (tools/ploc_tune.py)
__global__ void choose_kernel(BuildResult *res, int n_candidates) {
    int chosen = 0;
    if (n_candidates == 2) {
        const bool ok0 = res->status[0] == 0 && res->height[0] <= kMaxTreeHeight;
        const bool ok1 = res->status[1] == 0 && res->height[1] <= kMaxTreeHeight;
        chosen = (ok0 && (!ok1 || res->sah_cost[0] < 0.8f * res->sah_cost[1])) ? 0 : 1;
    }
    res->chosen = chosen;
}

struct LayoutArgs {
    DevTree tree[2];
    const BuildResult *res;
    float4 *nodes;         // final traversal nodes
    int np;
};

// Outward padding of every child box: per axis 1e-4 * extent + 4e-6 * (scene diagonal + largest |coordinate| of the
// box).  The first term follows the box, the second keeps flat (zero-thickness) boxes and boxes far from the world
// origin a few dozen ulps thick, so that the fp32 slab test with FMA cannot reject a ray the exact primitive test
// accepts.  Then centre / half-extent; the half-extent is rounded up so the stored box contains the padded one.
__device__ __forceinline__ void encode_box(const float *mn, const float *mx, float diag, float *c, float *h) {
    for (int k = 0; k < 3; k++) {
        const float ext = mx[k] - mn[k];
        const float mag = fmaxf(fabsf(mn[k]), fabsf(mx[k]));
        const float p = 1e-4f * ext + 4e-6f * (diag + mag);
        const float lo = mn[k] - p, hi = mx[k] + p;
        c[k] = 0.5f * (lo + hi);
        h[k] = fmaxf(hi - c[k], c[k] - lo) * 1.000001f + fabsf(c[k]) * 2e-7f;
    }
}

__global__ void layout_kernel(LayoutArgs a) {
    const DevTree &t = a.tree[a.res->chosen];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (*t.status != 0 || i >= *t.n_used) return;
    // depth-first (pre-order) index: on the way to the root, every step from a second child skips the first child's
    // subtree; a node that does not reach the root is unreachable (left behind by a leaf collapse) and dropped
    int idx = 0, cur = i;
    for (int guard = 0; cur != 0; guard++) {
        const int pp = t.parent[cur];
        if (pp < 0 || guard > (1 << 24)) return;
        const int par = pp >> 1;
        idx += 1;
        if (pp & 1) {
            const int c0 = t.nodes[par].child0;
            if (is_inner(c0)) idx += t.count[c0];
        }
        cur = par;
    }
    float diag;
    {
        double s = 0;
        for (int k = 0; k < 3; k++) {
            const double e = (double) ord2f(a.res->scene_bounds[3 + k]) - (double) ord2f(a.res->scene_bounds[k]);
            s += e * e;
        }
        diag = (float) sqrt(s);
    }
    HostNode nd = t.nodes[i];
    // A single-primitive scene has a root with one real child.  The missing child becomes a leaf over a dummy
    // all-zero triangle (slot np: detA = 0, every comparison on NaN fails, it can never report a hit), so the
    // traversal loop needs no "empty child" test.
    if (nd.child1 == kEmptyChild) {
        nd.child1 = ~((a.np << 3) | 0);
        for (int k = 0; k < 3; k++) nd.c1mn[k] = nd.c0mn[k], nd.c1mx[k] = nd.c0mx[k];
    }
    if (nd.child0 == kEmptyChild) {
        nd.child0 = ~((a.np << 3) | 0);
        for (int k = 0; k < 3; k++) nd.c0mn[k] = nd.c1mn[k], nd.c0mx[k] = nd.c1mx[k];
    }
    float c0[3], h0[3], c1[3], h1[3];
    encode_box(nd.c0mn, nd.c0mx, diag, c0, h0);
    encode_box(nd.c1mn, nd.c1mx, diag, c1, h1);
    // inner references are float4 indices (4 * node index), see rt_internal.h
    const int first_count = nd.child0 >= 0 ? t.count[nd.child0] : 0;
    const int r0 = nd.child0 >= 0 ? 4 * (idx + 1) : nd.child0;
    const int r1 = nd.child1 >= 0 ? 4 * (idx + 1 + first_count) : nd.child1;
    RT_CHECK(idx >= 0 && idx < a.res->n_nodes[a.res->chosen] && idx < (a.np > 1 ? a.np : 1));
    RT_CHECK(nd.child0 >= 0 || ((~nd.child0) >> 3) + ((~nd.child0) & 7) + 1 <= a.np + 1);
    RT_CHECK(nd.child1 >= 0 || ((~nd.child1) >> 3) + ((~nd.child1) & 7) + 1 <= a.np + 1);
    float4 *o = a.nodes + 4 * (size_t) idx;
    o[0] = make_float4(c0[0], h0[0], c0[1], h0[1]);
    o[1] = make_float4(c1[0], h1[0], c1[1], h1[1]);
    o[2] = make_float4(c0[2], h0[2], c1[2], h1[2]);
    o[3] = make_float4(__int_as_float(r0), __int_as_float(r1), 0.f, 0.f);
}

__global__ void place_prims_kernel(LayoutArgs a, const float4 *rec, float4 *prims, int *slot_of_prim) {
    const DevTree &t = a.tree[a.res->chosen];
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s > a.np) return;
    if (s == a.np) {  // the never-hit dummy primitive
        for (int k = 0; k < 3; k++) prims[3 * (size_t) s + k] = make_float4(0.f, 0.f, 0.f, 0.f);
        return;
    }
    if (*t.status != 0) return;
    const int id = t.prim_order[s];
    RT_CHECK(id >= 0 && id < a.np);
    for (int k = 0; k < 3; k++) prims[3 * (size_t) s + k] = rec[3 * (size_t) id + k];
    slot_of_prim[id] = s;
}

// ---------------------------------------------------------------------------------------------------------------
// arenas
// ---------------------------------------------------------------------------------------------------------------

struct Bump {
    char *base = nullptr;
    size_t off = 0;
    template <typename T>
    T *take(size_t count) {
        off = (off + 255) & ~(size_t) 255;
        T *p = base ? (T *) (base + off) : nullptr;
        off += sizeof(T) * (count ? count : 1);
        return p;
    }
};

struct Scratch {
    RtVec3 *vertices;
    RtTriangle *triangles;
    RtSphere *spheres;
    Aabb *bounds;
    float *key;
    float4 *rec;
    RefScratch ref;
    MortonScratch morton;
    PlocScratch ploc;
    LbvhScratch lbvh;
    SahScratch sah;
    ReinsertScratch reinsert;
    DevTree tree[2];
};

void carve_tree(Bump &b, DevTree &t, int cap, int np) {
    t.cap = cap;
    t.nodes = b.take<HostNode>(cap);
    t.prim_order = b.take<int>(np);
    t.status = b.take<int>(2);
    t.n_used = t.status ? t.status + 1 : nullptr;
    t.parent = b.take<int>(cap);
    t.arrivals = b.take<int>(cap);
    t.count = b.take<int>(cap);
    t.height = b.take<int>(cap);
    t.area = b.take<float>(cap);
}

void carve_scratch(Bump &b, Scratch &s, const RtSceneDesc &d, int builder, int ploc_grid, int reinsert_grid) {
    const int nt = d.n_triangles, ns = d.n_spheres, np = nt + ns;
    s.vertices = b.take<RtVec3>(d.n_vertices);
    s.triangles = b.take<RtTriangle>(nt);
    s.spheres = b.take<RtSphere>(ns);
    s.bounds = b.take<Aabb>(np);
    s.key = b.take<float>((size_t) 3 * np);
    s.rec = b.take<float4>((size_t) 3 * np);
    s.ref.ids = b.take<int>(np);
    s.ref.tmp = b.take<int>(np);
    s.ref.queue[0] = b.take<RefTask>((size_t) np + 1);
    s.ref.queue[1] = b.take<RefTask>((size_t) np + 1);
    s.ref.level_count = b.take<int>(kRefLevels + 2);
    s.ref.n_nodes = b.take<int>(1);
    s.ref.nodes = b.take<DevRefNode>((size_t) 2 * np + 1);
    const bool morton = builder == RT_BUILD_AUTO || builder == RT_BUILD_PLOC_GPU || builder == RT_BUILD_LBVH_GPU;
    const bool ploc = builder == RT_BUILD_AUTO || builder == RT_BUILD_PLOC_GPU;
    const bool sah = builder == RT_BUILD_AUTO || builder == RT_BUILD_SAH_GPU;
    const int cap = np > 1 ? np : 1;
    if (morton) {
        s.morton.n_pad = morton_pad(np);
        s.morton.keys = b.take<unsigned long long>(s.morton.n_pad);
    }
    if (ploc) {
        s.ploc.nodes = b.take<char>(ploc_node_bytes() * 2 * (size_t) cap);
        s.ploc.cl_a = b.take<int>((size_t) 2 * cap);  // cluster lists a | b; later the first-leaf positions of all 2n-1 nodes
        s.ploc.cl_b = s.ploc.cl_a ? s.ploc.cl_a + cap : nullptr;
        s.ploc.nn = b.take<int>(cap);
        s.ploc.cta_tot = b.take<int>((size_t) 2 * (ploc_grid > 0 ? ploc_grid : 1));
        s.ploc.state = b.take<int>(4);
    }
    if (builder == RT_BUILD_LBVH_GPU) {
        s.lbvh.nodes = b.take<char>(lbvh_node_bytes() * (size_t) cap);
        s.lbvh.leaf_parent = b.take<int>(cap);
        s.lbvh.collapsed = b.take<int>(cap);
        s.lbvh.flags = b.take<unsigned>(cap);
        s.lbvh.cost = b.take<float>(cap);
        s.lbvh.box = b.take<Aabb>(cap);
    }
    if (sah) {
        s.sah.ids = b.take<int>(cap);
        s.sah.tmp = b.take<int>(cap);
        s.sah.queue[0] = b.take<SahTask>((size_t) cap + 1);
        s.sah.queue[1] = b.take<SahTask>((size_t) cap + 1);
        s.sah.level_count = b.take<int>(kSahLevels + 2);
        s.sah.n_nodes = b.take<int>(1);
        s.sah.root_ref = b.take<int>(1);
        s.sah.cap_tasks = cap;
        s.sah.cap_nodes = cap + 1;
    }
    if (sah && reinsert_grid > 0) {
        const size_t ne = (size_t) (cap + 1) + cap + 1;  // node slots + leaf ranges by first primitive
        s.reinsert.ne = (int) ne;
        s.reinsert.box = b.take<Aabb>(ne);
        s.reinsert.left = b.take<int>(ne);
        s.reinsert.right = b.take<int>(ne);
        s.reinsert.parent = b.take<int>(ne);
        s.reinsert.lock = b.take<unsigned long long>(ne);
        s.reinsert.key = b.take<unsigned long long>(ne);
        s.reinsert.mv_y = b.take<int>(ne);
        s.reinsert.mv_pivot = b.take<int>(ne);
        s.reinsert.canon = b.take<int>(ne);
        s.reinsert.partial = b.take<float>(reinsert_grid);
        s.reinsert.counters = b.take<int>(reinsert_counter_slots());
    }
    carve_tree(b, s.tree[0], cap + 1, cap);
    if (builder == RT_BUILD_AUTO) carve_tree(b, s.tree[1], cap + 1, cap);
}

void carve_scene(Bump &b, SceneBuffers &o, const RtSceneDesc &d) {
    const int nt = d.n_triangles, ns = d.n_spheres, np = nt + ns;
    o.control = b.take<unsigned long long>(8 * kControlSlots);
    o.result = b.take<BuildResult>(1);
    o.nodes = b.take<float4>((size_t) 4 * (np > 1 ? np : 1));
    o.prims = b.take<float4>((size_t) 3 * (np + 1));
    o.tri_nm = b.take<float4>(nt);
    o.tri_nn = b.take<float4>(nt);
    o.sph_cr = b.take<float4>(ns);
    o.sph_mat = b.take<int>(ns);
    o.ranks = b.take<uint32_t>((size_t) 8 * np);
    o.ref_nodes = b.take<float4>((size_t) 3 * (2 * (size_t) np + 1));
    o.ref_leaf_prims = b.take<int>(np);
    o.prim_bounds = b.take<float4>((size_t) 2 * np);
    o.slot_of_prim = b.take<int>(np);
    o.materials = b.take<float4>((size_t) 4 * d.n_materials);
    o.lights = b.take<float4>((size_t) 2 * d.n_lights);
}

#define CKB(call)                                                                          \
    do {                                                                                   \
        cudaError_t e_ = (call);                                                           \
        if (e_ != cudaSuccess) {                                                           \
            err = std::string(#call) + ": " + cudaGetErrorString(e_);                      \
            return -1;                                                                     \
        }                                                                                  \
    } while (0)

void enqueue_tree_finalise(DevTree &t, BuildResult *res, int slot, cudaStream_t stream) {
    const int T = 256, g = (t.cap + T - 1) / T;
    tree_init_kernel<<<g, T, 0, stream>>>(t);
    tree_parents_kernel<<<g, T, 0, stream>>>(t);
    tree_reduce_kernel<<<g, T, 0, stream>>>(t);
    tree_root_kernel<<<1, 1, 0, stream>>>(t, res, slot);
}

double now_ms() {
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

}  // namespace

// ---- block cache ----------------------------------------------------------------------------------------------
namespace {
struct CachedBlock {
    int device;
    void *ptr;
    size_t cap;
};
std::mutex g_cache_mutex;
std::vector<CachedBlock> g_cache;
constexpr size_t kCacheBlocksPerDevice = 6;
}  // namespace

void *block_acquire(size_t bytes, size_t *cap) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
    {
        std::lock_guard<std::mutex> lock(g_cache_mutex);
        int best = -1;
        for (size_t i = 0; i < g_cache.size(); i++)
            if (g_cache[i].device == dev && g_cache[i].cap >= bytes && g_cache[i].cap <= 4 * bytes + (1 << 20) &&
                (best < 0 || g_cache[i].cap < g_cache[(size_t) best].cap))
                best = (int) i;
        if (best >= 0) {
            void *p = g_cache[(size_t) best].ptr;
            *cap = g_cache[(size_t) best].cap;
            g_cache.erase(g_cache.begin() + best);
            return p;
        }
    }
    void *p = nullptr;
    const size_t want = (bytes + (1 << 16)) & ~(size_t) ((1 << 16) - 1);  // a little headroom so that similar scenes reuse it
    if (cudaMalloc(&p, want) != cudaSuccess) {
        cudaGetLastError();
        block_cache_trim();  // give the cached blocks back and try once more
        if (cudaMalloc(&p, want) != cudaSuccess) {
            cudaGetLastError();
            return nullptr;
        }
    }
    *cap = want;
    return p;
}

void block_release(void *ptr, size_t cap) {
    if (!ptr) return;
    int dev = 0;
    cudaGetDevice(&dev);
    void *evict = nullptr;
    {
        std::lock_guard<std::mutex> lock(g_cache_mutex);
        size_t mine = 0;
        int smallest = -1;
        for (size_t i = 0; i < g_cache.size(); i++)
            if (g_cache[i].device == dev) {
                mine++;
                if (smallest < 0 || g_cache[i].cap < g_cache[(size_t) smallest].cap) smallest = (int) i;
            }
        if (mine >= kCacheBlocksPerDevice) {  // full: keep the larger blocks
            if (g_cache[(size_t) smallest].cap < cap) {
                evict = g_cache[(size_t) smallest].ptr;
                g_cache[(size_t) smallest] = CachedBlock{dev, ptr, cap};
            } else {
                evict = ptr;
            }
        } else {
            g_cache.push_back(CachedBlock{dev, ptr, cap});
        }
    }
    if (evict) cudaFree(evict);
}

void block_cache_trim() {
    std::vector<CachedBlock> all;
    {
        std::lock_guard<std::mutex> lock(g_cache_mutex);
        all.swap(g_cache);
    }
    int prev = 0;
    cudaGetDevice(&prev);
    for (auto &b: all) {
        cudaSetDevice(b.device);
        cudaFree(b.ptr);
    }
    cudaSetDevice(prev);
}

void SceneBuild::release_scratch() {
    if (scratch) block_release(scratch, scratch_bytes);
    scratch = nullptr;
}

int SceneBuild::run(const RtSceneDesc &d, int builder, int ploc_radius, float ploc_leaf_cost, int reinsert_rounds, float reinsert_accept,
                    int n_sms, cudaStream_t stream, SceneBuffers &out, std::string &err) {
    const int nt = d.n_triangles, ns = d.n_spheres, np = nt + ns;
    const double t_begin = now_ms();
    int ploc_grid = 1;
    if (builder == RT_BUILD_AUTO || builder == RT_BUILD_PLOC_GPU) ploc_grid = ploc_max_grid(n_sms);
    last_builder = builder;
    last_ploc_grid = ploc_grid;
    const bool sah_device = builder == RT_BUILD_AUTO || builder == RT_BUILD_SAH_GPU;
    const int reinsert_grid = (sah_device && reinsert_rounds > 0 && np > 2) ? reinsert_max_grid(n_sms) : 0;
    last_reinsert_grid = reinsert_grid;

    // ---- the two arenas -----------------------------------------------------------------------------------
    Bump size_scene, size_scratch;
    SceneBuffers dummy_o;
    Scratch dummy_s = Scratch();
    carve_scene(size_scene, dummy_o, d);
    carve_scratch(size_scratch, dummy_s, d, builder, ploc_grid, reinsert_grid);
    arena = block_acquire(size_scene.off + 256, &arena_bytes);
    scratch = block_acquire(size_scratch.off + 256, &scratch_bytes);
    if (!arena || !scratch) {
        err = "out of device memory";
        return -1;
    }
    Bump bs, bx;
    bs.base = (char *) arena;
    bx.base = (char *) scratch;
    Scratch s = Scratch();
    carve_scene(bs, out, d);
    carve_scratch(bx, s, d, builder, ploc_grid, reinsert_grid);

    cudaEvent_t e0 = nullptr, e1 = nullptr;
    CKB(cudaEventCreate(&e0));
    CKB(cudaEventCreate(&e1));
    struct EventGuard {
        cudaEvent_t &a, &b;
        ~EventGuard() {
            if (a) cudaEventDestroy(a);
            if (b) cudaEventDestroy(b);
        }
    } guard{e0, e1};

    // ---- uploads (pageable sources are staged by the runtime before the call returns) --------------------------
    std::vector<float4> mats((size_t) d.n_materials * 4), lights((size_t) d.n_lights * 2);
    auto bits = [](int v) { float f; memcpy(&f, &v, 4); return f; };
    for (int i = 0; i < d.n_materials; i++) {
        const RtMaterial &m = d.materials[i];
        mats[4 * (size_t) i + 0] = make_float4(m.ambient.x, m.ambient.y, m.ambient.z, m.phong_exponent);
        // bit 0: mirror; bit 1: the specular term of this material is exactly +-0 whatever the geometry (ks == 0 and an
        // exponent for which pow() of a value in [0, 1 + 3e-7] stays finite) — render_v2.cu then skips computing it
        const bool no_specular = m.specular.x == 0.0f && m.specular.y == 0.0f && m.specular.z == 0.0f && m.phong_exponent >= 0.0f &&
                                 m.phong_exponent <= 1e6f;
        mats[4 * (size_t) i + 1] = make_float4(m.diffuse.x, m.diffuse.y, m.diffuse.z, bits((m.is_mirror ? 1 : 0) | (no_specular ? 2 : 0)));
        mats[4 * (size_t) i + 2] = make_float4(m.specular.x, m.specular.y, m.specular.z, 0.f);
        mats[4 * (size_t) i + 3] = make_float4(m.mirror.x, m.mirror.y, m.mirror.z, 0.f);
    }
    for (int i = 0; i < d.n_lights; i++) {
        const RtPointLight &l = d.lights[i];
        lights[2 * (size_t) i + 0] = make_float4(l.position.x, l.position.y, l.position.z, 0.f);
        lights[2 * (size_t) i + 1] = make_float4(l.intensity.x, l.intensity.y, l.intensity.z, 0.f);
    }
    CKB(cudaEventRecord(e0, stream));
    CKB(cudaMemsetAsync(out.control, 0, sizeof(unsigned long long) * 8 * kControlSlots, stream));
    if (d.n_vertices) CKB(cudaMemcpyAsync(s.vertices, d.vertices, sizeof(RtVec3) * d.n_vertices, cudaMemcpyHostToDevice, stream));
    if (nt) CKB(cudaMemcpyAsync(s.triangles, d.triangles, sizeof(RtTriangle) * nt, cudaMemcpyHostToDevice, stream));
    if (ns) CKB(cudaMemcpyAsync(s.spheres, d.spheres, sizeof(RtSphere) * ns, cudaMemcpyHostToDevice, stream));
    if (!mats.empty()) CKB(cudaMemcpyAsync(out.materials, mats.data(), sizeof(float4) * mats.size(), cudaMemcpyHostToDevice, stream));
    if (!lights.empty()) CKB(cudaMemcpyAsync(out.lights, lights.data(), sizeof(float4) * lights.size(), cudaMemcpyHostToDevice, stream));
    const double t_uploaded = now_ms();

    // ---- device pipeline ----------------------------------------------------------------------------------
    const int T = 256;
    init_result_kernel<<<1, 32, 0, stream>>>(out.result);
    int n_candidates = 0;
    bool host_tree = false;
    if (np > 0) {
        PrepArgs pa = {s.vertices, s.triangles, s.spheres, nt, ns, s.bounds, s.key, s.rec, out.prim_bounds, out.tri_nm, out.tri_nn,
                       out.sph_cr, out.sph_mat, out.result};
        prep_kernel<<<(np + T - 1) / T, T, 0, stream>>>(pa);
        {
            const int e = enqueue_reference_tree(s.bounds, s.key, np, s.ref, out.ranks, out.ref_nodes, out.ref_leaf_prims, out.result,
                                                 ref_max_grid(n_sms), stream);
            if (e != 0) {
                err = std::string("reference tree launch: ") + cudaGetErrorString((cudaError_t) e);
                return -1;
            }
        }
        if (np == 1) {
            single_prim_tree_kernel<<<1, 1, 0, stream>>>(s.bounds, s.tree[0]);
            n_candidates = 1;
        } else if (builder == RT_BUILD_SAH_HOST) {
            host_tree = true;
        } else {
            if (builder == RT_BUILD_AUTO || builder == RT_BUILD_PLOC_GPU || builder == RT_BUILD_LBVH_GPU)
                enqueue_morton_sort(s.bounds, np, out.result->centroid_bounds, s.morton, stream);
            if (builder == RT_BUILD_AUTO || builder == RT_BUILD_PLOC_GPU) {
                set_int_kernel<<<1, 1, 0, stream>>>(s.tree[0].n_used, np - 1);
                const int e = enqueue_ploc(s.bounds, np, s.morton, s.ploc, s.tree[0], ploc_radius, ploc_leaf_cost, ploc_grid, stream);
                if (e != 0) {
                    err = std::string("PLOC launch: ") + cudaGetErrorString((cudaError_t) e);
                    return -1;
                }
                n_candidates = 1;
            }
            if (builder == RT_BUILD_LBVH_GPU) {
                set_int_kernel<<<1, 1, 0, stream>>>(s.tree[0].n_used, np - 1);
                enqueue_lbvh(s.bounds, np, s.morton, s.lbvh, s.tree[0], stream);
                n_candidates = 1;
            }
            if (builder == RT_BUILD_AUTO || builder == RT_BUILD_SAH_GPU) {
                DevTree &t = s.tree[builder == RT_BUILD_AUTO ? 1 : 0];
                s.sah.n_nodes = t.n_used;  // the builder's node counter IS the tree's used-slot count
                t.prim_order = s.sah.ids;  // ... and the index array it partitions IS the leaf order
                const int e = enqueue_sah(s.bounds, np, s.sah, t, out.result, sah_max_grid(n_sms), stream);
                if (e != 0) {
                    err = std::string("SAH builder launch: ") + cudaGetErrorString((cudaError_t) e);
                    return -1;
                }
                if (reinsert_grid > 0) {
                    const int e2 = enqueue_reinsert(t, s.reinsert, reinsert_rounds, reinsert_accept, out.result, reinsert_grid, stream);
                    if (e2 != 0) {
                        err = std::string("reinsertion launch: ") + cudaGetErrorString((cudaError_t) e2);
                        return -1;
                    }
                }
                n_candidates = builder == RT_BUILD_AUTO ? 2 : 1;
            }
        }
    }
    LayoutArgs la;
    la.tree[0] = s.tree[0];
    la.tree[1] = n_candidates == 2 ? s.tree[1] : s.tree[0];
    la.res = out.result;
    la.nodes = out.nodes;
    la.np = np;
    auto finalise = [&](int n_cand) {
        for (int c = 0; c < n_cand; c++) enqueue_tree_finalise(s.tree[c], out.result, c, stream);
        choose_kernel<<<1, 1, 0, stream>>>(out.result, n_cand);
        const int cap = s.tree[0].cap > la.tree[1].cap ? s.tree[0].cap : la.tree[1].cap;
        layout_kernel<<<(cap + T - 1) / T, T, 0, stream>>>(la);
        place_prims_kernel<<<(np + 1 + T - 1) / T, T, 0, stream>>>(la, s.rec, out.prims, out.slot_of_prim);
    };
    if (n_candidates > 0) finalise(n_candidates);
    else if (np == 0) place_prims_kernel<<<1, 1, 0, stream>>>(la, s.rec, out.prims, out.slot_of_prim);  // just the dummy primitive

    auto fetch = [&]() -> int {
        CKB(cudaMemcpyAsync(&result, out.result, sizeof result, cudaMemcpyDeviceToHost, stream));
        CKB(cudaStreamSynchronize(stream));
        CKB(cudaGetLastError());
        return 0;
    };
    if (!host_tree && fetch() != 0) return -1;

    // ---- host fallback: RT_BUILD_SAH_HOST, or every device candidate unusable (a pathological primitive order made
    // the clustered tree too deep for the traversal stack, or the clustering gave up): the host builder splits at the
    // median when the SAH finds nothing and stays O(log n) deep
    used_host_builder = false;
    if (np > 1 && (host_tree || result.status[result.chosen] != 0 || result.height[result.chosen] > kMaxTreeHeight)) {
        std::vector<Aabb> hb;
        primitive_bounds(d, hb);
        HostBvh bvh;
        build_bvh_sah_host(hb, bvh, reinsert_rounds, reinsert_accept);
        if ((int) bvh.nodes.size() > s.tree[0].cap) {
            err = "host BVH larger than the node arena";
            return -1;
        }
        const int n_used = (int) bvh.nodes.size(), zero = 0;
        CKB(cudaMemcpyAsync(s.tree[0].nodes, bvh.nodes.data(), sizeof(HostNode) * bvh.nodes.size(), cudaMemcpyHostToDevice, stream));
        CKB(cudaMemcpyAsync(s.tree[0].prim_order, bvh.prim_order.data(), sizeof(int) * np, cudaMemcpyHostToDevice, stream));
        CKB(cudaMemcpyAsync(s.tree[0].n_used, &n_used, sizeof(int), cudaMemcpyHostToDevice, stream));
        CKB(cudaMemcpyAsync(s.tree[0].status, &zero, sizeof(int), cudaMemcpyHostToDevice, stream));
        la.tree[1] = s.tree[0];
        finalise(1);
        if (fetch() != 0) return -1;
        used_host_builder = true;
    }
    CKB(cudaEventRecord(e1, stream));
    CKB(cudaEventSynchronize(e1));
    cudaEventElapsedTime(&ms_device, e0, e1);
    ms_host_before_sync = (float) (t_uploaded - t_begin);
    ms_wall = (float) (now_ms() - t_begin);
    if (np > 0 && (result.status[result.chosen] != 0 || result.height[result.chosen] > kMaxTreeHeight)) {
        err = "BVH deeper than the traversal stack";
        return -2;
    }
    return 0;
}

// ---- validators used by the test hooks (api.cu): read intermediate device state back ---------------------------

int SceneBuild::read_reference_tree(const RtSceneDesc &d, cudaStream_t stream, const SceneBuffers &out, std::vector<uint32_t> &ranks,
                                    RefTreeStats &stats, RefTree &tree, std::string &err) {
    const int np = d.n_triangles + d.n_spheres;
    ranks.assign((size_t) 8 * np, 0u);
    stats = RefTreeStats();
    tree = RefTree();
    if (np == 0) return 0;
    if (!scratch) {
        err = "scratch already released";
        return -1;
    }
    // re-derive the scratch layout (same carve order as run())
    Bump bx;
    bx.base = (char *) scratch;
    Scratch s = Scratch();
    carve_scratch(bx, s, d, last_builder, last_ploc_grid, last_reinsert_grid);
    const int n_nodes = result.ref_nodes;
    std::vector<DevRefNode> dn((size_t) n_nodes);
    CKB(cudaMemcpyAsync(dn.data(), s.ref.nodes, sizeof(DevRefNode) * n_nodes, cudaMemcpyDeviceToHost, stream));
    CKB(cudaMemcpyAsync(ranks.data(), out.ranks, sizeof(uint32_t) * 8 * np, cudaMemcpyDeviceToHost, stream));
    tree.leaf_prims.resize((size_t) np);
    CKB(cudaMemcpyAsync(tree.leaf_prims.data(), out.ref_leaf_prims, sizeof(int) * np, cudaMemcpyDeviceToHost, stream));
    CKB(cudaStreamSynchronize(stream));
    // relabel into the reference's pre-order (left child = index + 1, bvh.h:81-105), the numbering ref_order.cpp uses
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
