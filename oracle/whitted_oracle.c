/*
 * oracle/whitted_oracle.c — TEST INFRASTRUCTURE, not product code.
 *
 * A plain-C, CPU restatement of the reference's per-pixel Whitted path, written against
 * the flat scene description of include/rt_b200.h.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it; the product (the CUDA
 * library) never does.  Every function cites the reference lines it restates.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks this file byte-for-byte against
 * images rendered by the unmodified reference (oracle/_ref/libref.so, built from
 * /root/reference by oracle/Makefile) for all 16 cameras of the 13 shipped scenes; the
 * committed fixtures under tests/golden/ are those reference renders.
 *
 * Arithmetic contract: IEEE fp32, no FMA contraction (build with
 * -ffp-contract=off), source operation order, true division; double precision exactly
 * where the reference promotes (sphere roots, acos, pow, the (col+0.5)*mul product).
 */
#include "rt_b200.h"

#include <float.h>
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct { float x, y, z; } V3;

static inline V3 v3(float x, float y, float z) { V3 r = {x, y, z}; return r; }
/* parser.h:22-48 */
static inline V3 vadd(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline V3 vsub(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline V3 vmulf(V3 a, float f) { return v3(a.x * f, a.y * f, a.z * f); }
static inline V3 vdivf(V3 a, float f) { return v3(a.x / f, a.y / f, a.z / f); }
static inline V3 vneg(V3 a) { return v3(-a.x, -a.y, -a.z); }
static inline float vdot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline V3 vcross(V3 a, V3 v) { return v3(a.y * v.z - a.z * v.y, a.z * v.x - a.x * v.z, a.x * v.y - a.y * v.x); }
static inline V3 vmulv(V3 a, V3 b) { return v3(a.x * b.x, a.y * b.y, a.z * b.z); }
/* parser.h:77-79: ::sqrt(double) of a float sum, result narrowed to float */
static inline float vlen(V3 a) { return (float) sqrt((double) (a.x * a.x + a.y * a.y + a.z * a.z)); }
/* parser.h:72-75 */
static inline V3 vnorm(V3 a) { float l = vlen(a); return v3(a.x / l, a.y / l, a.z / l); }
static inline float vget(V3 a, int i) { return i == 1 ? a.y : (i == 2 ? a.z : a.x); } /* parser.h:55-66 */
/* std::min / std::max exactly as libstdc++ defines them (NaN behaviour matters) */
static inline float stdmin(float a, float b) { return (b < a) ? b : a; }
static inline float stdmax(float a, float b) { return (a < b) ? b : a; }
/* parser.h:81-86 Vec3f::clamp(a,b) = max(a, min(x, b)) */
static inline float clampf(float x, float a, float b) { return stdmax(a, stdmin(x, b)); }

typedef struct { V3 min, max; } Box;

typedef struct {
    int32_t v0, v1, v2, material_id; /* 1-based */
    V3 normal, center;
} OTri;

typedef struct {
    Box box;
    int axis;
    int right;              /* index of the right child (left = self + 1), bvh.h:81-105 */
    int is_leaf;
    int tri_begin, tri_count; /* into leaf_tris */
    int sph_begin, sph_count; /* into leaf_sphs */
    int depth;
} ONode;

typedef struct OrScene {
    RtSceneDesc d; /* deep copy */
    OTri *tris;
    int n_tris;
    ONode *nodes;
    int n_nodes, cap_nodes;
    int *leaf_tris; int n_leaf_tris;
    int *leaf_sphs; int n_leaf_sphs;
} OrScene;

typedef struct OrStats {
    uint64_t primary_rays, reflection_rays, shadow_rays, shadow_occluded;
    uint64_t box_tests, tri_tests, sphere_tests;
} OrStats;

/* ------------------------------------------------------------------ BVH (bvh.h:48-163) */

static V3 vtx(const OrScene *s, int id1) { const RtVec3 *p = &s->d.vertices[id1 - 1]; return v3(p->x, p->y, p->z); }

/* parser.h:272-317 getBoundingBox + extendBoundingBox */
static Box bounds(const OrScene *s, const int *tris, int nt, const int *sphs, int ns) {
    Box b = {{FLT_MAX, FLT_MAX, FLT_MAX}, {-FLT_MAX, -FLT_MAX, -FLT_MAX}};
    for (int i = 0; i < nt; i++) {
        const OTri *t = &s->tris[tris[i]];
        int ids[3] = {t->v0, t->v1, t->v2};
        for (int k = 0; k < 3; k++) {
            V3 p = vtx(s, ids[k]);
            if (p.x < b.min.x) b.min.x = p.x;
            if (p.y < b.min.y) b.min.y = p.y;
            if (p.z < b.min.z) b.min.z = p.z;
            if (p.x > b.max.x) b.max.x = p.x;
            if (p.y > b.max.y) b.max.y = p.y;
            if (p.z > b.max.z) b.max.z = p.z;
        }
    }
    for (int i = 0; i < ns; i++) {
        const RtSphere *sp = &s->d.spheres[sphs[i]];
        V3 c = vtx(s, sp->center_vertex_id);
        float cc[3] = {c.x, c.y, c.z};
        float *mn = &b.min.x, *mx = &b.max.x;
        for (int a = 0; a < 3; a++) {
            if (cc[a] - sp->radius < mn[a]) mn[a] = cc[a] - sp->radius;
            if (cc[a] + sp->radius > mx[a]) mx[a] = cc[a] + sp->radius;
        }
    }
    return b;
}

/* parser.h:227-235 */
static int widest_axis(Box b) {
    const float *mn = &b.min.x, *mx = &b.max.x;
    int w = 0;
    for (int a = 1; a < 3; a++)
        if (mx[a] - mn[a] > mx[w] - mn[w]) w = a;
    return w;
}

static int new_node(OrScene *s) {
    if (s->n_nodes == s->cap_nodes) {
        s->cap_nodes = s->cap_nodes ? s->cap_nodes * 2 : 1024;
        s->nodes = (ONode *) realloc(s->nodes, sizeof(ONode) * (size_t) s->cap_nodes);
    }
    memset(&s->nodes[s->n_nodes], 0, sizeof(ONode));
    return s->n_nodes++;
}

#define OR_MAX_DEPTH 19 /* bvh.h:18 */

/* bvh.h:48-79 build + bvh.h:111-163 partition; emitted directly in the pre-order of
 * bvh.h:81-105 (left subtree right after its parent).  tris/sphs are index lists whose
 * order is preserved by the stable split, as the reference's push_back loops do. */
static int build_node(OrScene *s, int *tris, int nt, int *sphs, int ns, int depth) {
    int me = new_node(s);
    Box box = bounds(s, tris, nt, sphs, ns);
    s->nodes[me].box = box;
    s->nodes[me].depth = depth;
    int leaf = (nt + ns <= 1) || depth >= OR_MAX_DEPTH;
    int *lt = NULL, *rt = NULL, *ls = NULL, *rs = NULL;
    int nlt = 0, nrt = 0, nls = 0, nrs = 0;
    if (!leaf) {
        int axis = widest_axis(box);
        s->nodes[me].axis = axis;
        float start = vget(box.min, axis), end = vget(box.max, axis);
        float mid = (start + end) / 2;
        int maxTries = 19, leftCount = 0, rightCount = 0;
        lt = (int *) malloc(sizeof(int) * (size_t) (nt + 1)); rt = (int *) malloc(sizeof(int) * (size_t) (nt + 1));
        ls = (int *) malloc(sizeof(int) * (size_t) (ns + 1)); rs = (int *) malloc(sizeof(int) * (size_t) (ns + 1));
        while (maxTries-- && (leftCount == 0 || rightCount == 0)) {
            leftCount = rightCount = 0;
            for (int i = 0; i < nt; i++) {
                if (vget(s->tris[tris[i]].center, axis) < mid) leftCount++; else rightCount++;
            }
            for (int i = 0; i < ns; i++) {
                if (vget(vtx(s, s->d.spheres[sphs[i]].center_vertex_id), axis) < mid) leftCount++; else rightCount++;
            }
            if (leftCount == 0) {
                start = mid;
                mid = (start + end) / 2;
            } else if (rightCount == 0) {
                end = mid;
                mid = (start + end) / 2;
            } else {
                for (int i = 0; i < nt; i++) {
                    if (vget(s->tris[tris[i]].center, axis) < mid) lt[nlt++] = tris[i]; else rt[nrt++] = tris[i];
                }
                for (int i = 0; i < ns; i++) {
                    if (vget(vtx(s, s->d.spheres[sphs[i]].center_vertex_id), axis) < mid) ls[nls++] = sphs[i]; else rs[nrs++] = sphs[i];
                }
            }
        }
        if ((nlt == 0 && nls == 0) || (nrt == 0 && nrs == 0)) leaf = 1; /* bvh.h:162, 71-74 */
    }
    if (leaf) {
        s->nodes[me].is_leaf = 1;
        s->nodes[me].tri_begin = s->n_leaf_tris;
        s->nodes[me].tri_count = nt;
        memcpy(s->leaf_tris + s->n_leaf_tris, tris, sizeof(int) * (size_t) nt);
        s->n_leaf_tris += nt;
        s->nodes[me].sph_begin = s->n_leaf_sphs;
        s->nodes[me].sph_count = ns;
        memcpy(s->leaf_sphs + s->n_leaf_sphs, sphs, sizeof(int) * (size_t) ns);
        s->n_leaf_sphs += ns;
    } else {
        build_node(s, lt, nlt, ls, nls, depth + 1); /* lands at me + 1 */
        int r = build_node(s, rt, nrt, rs, nrs, depth + 1);
        s->nodes[me].right = r;
    }
    free(lt); free(rt); free(ls); free(rs);
    return me;
}

void or_scene_destroy(OrScene *s) {
    if (!s) return;
    free((void *) s->d.vertices); free((void *) s->d.triangles); free((void *) s->d.spheres);
    free((void *) s->d.materials); free((void *) s->d.lights);
    free(s->tris); free(s->nodes); free(s->leaf_tris); free(s->leaf_sphs);
    free(s);
}

static void *dup_mem(const void *p, size_t n) {
    void *q = malloc(n ? n : 1);
    if (n) memcpy(q, p, n);
    return q;
}

/* raytracer.cpp:335-350: triangle list (already flattened by the caller in that order),
 * unit geometric normals, centroids, then the tree */
OrScene *or_scene_create(const RtSceneDesc *desc) {
    OrScene *s = (OrScene *) calloc(1, sizeof(OrScene));
    s->d = *desc;
    s->d.vertices = (const RtVec3 *) dup_mem(desc->vertices, sizeof(RtVec3) * (size_t) desc->n_vertices);
    s->d.triangles = (const RtTriangle *) dup_mem(desc->triangles, sizeof(RtTriangle) * (size_t) desc->n_triangles);
    s->d.spheres = (const RtSphere *) dup_mem(desc->spheres, sizeof(RtSphere) * (size_t) desc->n_spheres);
    s->d.materials = (const RtMaterial *) dup_mem(desc->materials, sizeof(RtMaterial) * (size_t) desc->n_materials);
    s->d.lights = (const RtPointLight *) dup_mem(desc->lights, sizeof(RtPointLight) * (size_t) desc->n_lights);
    s->n_tris = desc->n_triangles;
    s->tris = (OTri *) malloc(sizeof(OTri) * (size_t) (s->n_tris + 1));
    for (int i = 0; i < s->n_tris; i++) {
        const RtTriangle *t = &desc->triangles[i];
        OTri *o = &s->tris[i];
        o->v0 = t->v0_id; o->v1 = t->v1_id; o->v2 = t->v2_id; o->material_id = t->material_id;
        V3 a = vtx(s, o->v0), b = vtx(s, o->v1), c = vtx(s, o->v2);
        o->normal = vnorm(vcross(vsub(b, a), vsub(c, a)));
        o->center = vdivf(vadd(vadd(a, b), c), 3);
    }
    s->leaf_tris = (int *) malloc(sizeof(int) * (size_t) (s->n_tris + 1));
    s->leaf_sphs = (int *) malloc(sizeof(int) * (size_t) (desc->n_spheres + 1));
    int *ti = (int *) malloc(sizeof(int) * (size_t) (s->n_tris + 1));
    int *si = (int *) malloc(sizeof(int) * (size_t) (desc->n_spheres + 1));
    for (int i = 0; i < s->n_tris; i++) ti[i] = i;
    for (int i = 0; i < desc->n_spheres; i++) si[i] = i;
    if (s->n_tris + desc->n_spheres > 0) build_node(s, ti, s->n_tris, si, desc->n_spheres, 0); /* bvh.h:49-51 */
    free(ti); free(si);
    return s;
}

/* nodes, leaves, max leaf size, max depth — known answers in SURVEY.md appendix A */
void or_bvh_stats(const OrScene *s, int *out4) {
    int leaves = 0, maxleaf = 0, maxdepth = 0;
    for (int i = 0; i < s->n_nodes; i++) {
        const ONode *n = &s->nodes[i];
        if (n->is_leaf) {
            leaves++;
            if (n->tri_count + n->sph_count > maxleaf) maxleaf = n->tri_count + n->sph_count;
        }
        if (n->depth > maxdepth) maxdepth = n->depth;
    }
    out4[0] = s->n_nodes; out4[1] = leaves; out4[2] = maxleaf; out4[3] = maxdepth;
}

/* Rank of every primitive (triangles 0..nt-1 then spheres nt..nt+ns-1) in the leaf visit
 * order of each of the 8 ray-direction sign octants (bit a set <=> direction[a] > 0,
 * raytracer.cpp:190-196).  out is [8][nt+ns].  Used by tests to cross-check the product's
 * own rank builder. */
void or_visit_ranks(const OrScene *s, uint32_t *out) {
    int np = s->n_tris + s->d.n_spheres;
    int *stack = (int *) malloc(sizeof(int) * 64);
    for (int oct = 0; oct < 8; oct++) {
        uint32_t next = 0;
        int sp = 0;
        if (s->n_nodes) stack[sp++] = 0;
        while (sp) {
            int n = stack[--sp];
            const ONode *nd = &s->nodes[n];
            if (!nd->is_leaf) {
                if ((oct >> nd->axis) & 1) { stack[sp++] = nd->right; stack[sp++] = n + 1; }
                else { stack[sp++] = n + 1; stack[sp++] = nd->right; }
            } else {
                for (int i = 0; i < nd->tri_count; i++) out[(size_t) oct * np + s->leaf_tris[nd->tri_begin + i]] = next++;
                for (int i = 0; i < nd->sph_count; i++) out[(size_t) oct * np + s->n_tris + s->leaf_sphs[nd->sph_begin + i]] = next++;
            }
        }
    }
    free(stack);
}

/* ------------------------------------------------------------------ rays (raytracer.cpp:47-282) */

typedef struct {
    V3 o, d, inv; /* raytracer.cpp:61-67: direction as passed (the normalize() there is a dead store) */
} Ray;

static Ray make_ray(V3 o, V3 d) {
    Ray r; r.o = o; r.d = d;
    r.inv = v3(1 / d.x, 1 / d.y, 1 / d.z);
    return r;
}

typedef struct {
    float tSmall;
    V3 normal;
    int material_id;
    int exists;
    int prim; /* debugging only: triangle index, or n_tris + sphere index */
} Hit;

/* raytracer.cpp:15-19 */
static float det3(float m[3][3]) {
    return m[0][0] * (m[1][1] * m[2][2] - m[1][2] * m[2][1]) -
           m[0][1] * (m[1][0] * m[2][2] - m[1][2] * m[2][0]) +
           m[0][2] * (m[1][0] * m[2][1] - m[1][1] * m[2][0]);
}

/* raytracer.cpp:70-96 */
static Hit hit_sphere(const OrScene *s, const Ray *ray, const RtSphere *sp) {
    Hit h = {-1, {0, 0, 0}, sp->material_id, 0};
    V3 c = vtx(s, sp->center_vertex_id);
    float r = sp->radius;
    V3 d = ray->d, o = ray->o;
    float B = 2 * vdot(d, vsub(o, c));
    float A = vdot(d, d);
    float C = vdot(vsub(o, c), vsub(o, c)) - r * r;
    float disc = B * B - 4 * A * C;
    if (disc >= 0) {
        float t1 = (float) ((-B - sqrt((double) disc)) / (2 * A));
        float t2 = (float) ((-B + sqrt((double) disc)) / (2 * A));
        if (t1 < 0 && t2 < 0) return h;
        h.exists = 1;
        h.tSmall = t1;
        h.normal = vnorm(vdivf(vsub(vadd(o, vmulf(d, t1)), c), r));
    }
    return h;
}

/* raytracer.cpp:101-126 */
static int hit_box(const Ray *ray, const Box *b, float *t) {
    float tx1 = (b->min.x - ray->o.x) * ray->inv.x;
    float tx2 = (b->max.x - ray->o.x) * ray->inv.x;
    float tmin = stdmin(tx1, tx2);
    float tmax = stdmax(tx1, tx2);
    float ty1 = (b->min.y - ray->o.y) * ray->inv.y;
    float ty2 = (b->max.y - ray->o.y) * ray->inv.y;
    tmin = stdmax(tmin, stdmin(ty1, ty2));
    tmax = stdmin(tmax, stdmax(ty1, ty2));
    float tz1 = (b->min.z - ray->o.z) * ray->inv.z;
    float tz2 = (b->max.z - ray->o.z) * ray->inv.z;
    tmin = stdmax(tmin, stdmin(tz1, tz2));
    tmax = stdmin(tmax, stdmax(tz1, tz2));
    if (tmax >= stdmax(0.0f, tmin)) { *t = tmin; return 1; }
    *t = -1;
    return 0;
}

/* raytracer.cpp:129-175 */
static Hit hit_triangle(const OrScene *s, const Ray *ray, const OTri *tr) {
    Hit p = {-1, tr->normal, tr->material_id, 0};
    V3 a = vtx(s, tr->v0), b = vtx(s, tr->v1), c = vtx(s, tr->v2);
    V3 d = ray->d, o = ray->o;
    float A[3][3] = {{a.x - b.x, a.x - c.x, d.x}, {a.y - b.y, a.y - c.y, d.y}, {a.z - b.z, a.z - c.z, d.z}};
    float detA = det3(A);
    float Bm[3][3] = {{a.x - o.x, a.x - c.x, d.x}, {a.y - o.y, a.y - c.y, d.y}, {a.z - o.z, a.z - c.z, d.z}};
    float beta = det3(Bm) / detA;
    float Gm[3][3] = {{a.x - b.x, a.x - o.x, d.x}, {a.y - b.y, a.y - o.y, d.y}, {a.z - b.z, a.z - o.z, d.z}};
    float gamma = det3(Gm) / detA;
    float Tm[3][3] = {{a.x - b.x, a.x - c.x, a.x - o.x}, {a.y - b.y, a.y - c.y, a.y - o.y}, {a.z - b.z, a.z - c.z, a.z - o.z}};
    float t = det3(Tm) / detA;
    float alpha = 1 - beta - gamma;
    if (alpha >= 0 && beta >= 0 && gamma >= 0 && t >= 0) {
        p.exists = 1;
        p.tSmall = t;
    }
    return p;
}

/* Debug switch (tests only): 1 makes every box test pass, i.e. every primitive is tested in the reference's
 * visit order — the "no culling" variant of SURVEY.md 7.3 used to locate rays whose hit the reference's own
 * slab test culls. */
static volatile int g_no_culling = 0;
void or_set_no_culling(int on) { g_no_culling = on; }

/* raytracer.cpp:177-225 */
static Hit first_hit(const OrScene *s, const Ray *ray, OrStats *st) {
    Hit best = {-1, {-1, -1, 0}, -1, 0};
    int stack[64]; int sp = 0;
    if (s->n_nodes) stack[sp++] = 0;
    float tMax = FLT_MAX;
    while (sp) {
        int n = stack[--sp];
        const ONode *nd = &s->nodes[n];
        float tb;
        st->box_tests++;
        int ok = hit_box(ray, &nd->box, &tb);
        if (g_no_culling) { ok = 1; tb = -FLT_MAX; }
        if (ok && tb <= tMax) {
            if (!nd->is_leaf) {
                if (vget(ray->d, nd->axis) > 0) { stack[sp++] = nd->right; stack[sp++] = n + 1; }
                else { stack[sp++] = n + 1; stack[sp++] = nd->right; }
            } else {
                for (int i = 0; i < nd->tri_count; i++) {
                    st->tri_tests++;
                    Hit h = hit_triangle(s, ray, &s->tris[s->leaf_tris[nd->tri_begin + i]]);
                    h.prim = s->leaf_tris[nd->tri_begin + i];
                    if (h.exists && (h.tSmall < best.tSmall || best.tSmall == -1)) { best = h; tMax = best.tSmall; }
                }
                for (int i = 0; i < nd->sph_count; i++) {
                    st->sphere_tests++;
                    Hit h = hit_sphere(s, ray, &s->d.spheres[s->leaf_sphs[nd->sph_begin + i]]);
                    h.prim = s->n_tris + s->leaf_sphs[nd->sph_begin + i];
                    if (h.exists && (h.tSmall < best.tSmall || best.tSmall == -1)) { best = h; tMax = best.tSmall; }
                }
            }
        }
    }
    return best;
}

/* raytracer.cpp:227-280 */
static int any_hit_until(const OrScene *s, const Ray *ray, float t, OrStats *st) {
    int stack[64]; int sp = 0;
    if (s->n_nodes) stack[sp++] = 0;
    while (sp) {
        int n = stack[--sp];
        const ONode *nd = &s->nodes[n];
        float tb;
        st->box_tests++;
        if (!hit_box(ray, &nd->box, &tb) && !g_no_culling) continue;
        if (!nd->is_leaf) {
            if (vget(ray->d, nd->axis) > 0) { stack[sp++] = nd->right; stack[sp++] = n + 1; }
            else { stack[sp++] = n + 1; stack[sp++] = nd->right; }
        } else {
            for (int i = 0; i < nd->tri_count; i++) {
                st->tri_tests++;
                Hit h = hit_triangle(s, ray, &s->tris[s->leaf_tris[nd->tri_begin + i]]);
                if (h.exists && h.tSmall < t) return 1;
            }
            for (int i = 0; i < nd->sph_count; i++) {
                st->sphere_tests++;
                Hit h = hit_sphere(s, ray, &s->d.spheres[s->leaf_sphs[nd->sph_begin + i]]);
                if (h.exists && h.tSmall < t) return 1;
            }
        }
    }
    return 0;
}

/* ------------------------------------------------------------------ shading (raytracer.cpp:385-452) */

static V3 ray_trace(const OrScene *s, Ray *ray, int depth, OrStats *st) {
    V3 color = {0, 0, 0};
    if (depth > s->d.max_recursion_depth) return color;
    if (depth == 0) st->primary_rays++; else st->reflection_rays++;
    Hit hit = first_hit(s, ray, st);
    if (hit.exists) {
        const RtMaterial *m = &s->d.materials[hit.material_id - 1];
        V3 ka = v3(m->ambient.x, m->ambient.y, m->ambient.z), kd = v3(m->diffuse.x, m->diffuse.y, m->diffuse.z);
        V3 ks = v3(m->specular.x, m->specular.y, m->specular.z), km = v3(m->mirror.x, m->mirror.y, m->mirror.z);
        V3 Ia = v3(s->d.ambient_light.x, s->d.ambient_light.y, s->d.ambient_light.z);
        color = vadd(color, vmulv(ka, Ia));
        V3 P = vadd(ray->o, vmulf(ray->d, hit.tSmall));
        V3 Pe = vadd(P, vmulf(hit.normal, s->d.shadow_ray_epsilon));
        for (int li = 0; li < s->d.n_lights; li++) {
            const RtPointLight *L = &s->d.lights[li];
            V3 lp = v3(L->position.x, L->position.y, L->position.z), I = v3(L->intensity.x, L->intensity.y, L->intensity.z);
            float dist = vlen(vsub(lp, Pe));
            V3 wi = vnorm(vsub(lp, Pe));
            V3 wiReal = vnorm(vsub(lp, vadd(ray->o, vmulf(ray->d, hit.tSmall))));
            Ray lray = make_ray(Pe, wi);
            st->shadow_rays++;
            int occluded = any_hit_until(s, &lray, dist, st);
            if (occluded) { st->shadow_occluded++; continue; }
            float cosTheta = vdot(wiReal, hit.normal);
            V3 E = vdivf(I, dist * dist);
            float theta = (float) (acos((double) cosTheta) * 180 / 3.1415);
            if (theta <= 90.01) {
                V3 h = vnorm(vadd(lray.d, vneg(vnorm(ray->d))));
                float c = (float) pow((double) stdmax(0.0f, vdot(vnorm(hit.normal), h)), (double) m->phong_exponent);
                color = vadd(color, vmulv(vmulf(ks, c), E));
            }
            /* clampFloat(x,0,1) = std::max(0, std::min(1, x)), raytracer.cpp:21-23 */
            float cd = stdmax(0.0f, stdmin(1.0f, cosTheta));
            color = vadd(color, vmulv(vmulf(kd, cd), E));
        }
        if (m->is_mirror) {
            V3 dn = vnorm(ray->d);
            V3 nn = vnorm(hit.normal);
            float rc = vdot(vneg(dn), nn);
            Ray rr = make_ray(Pe, vadd(dn, vmulf(vmulf(nn, 2), rc)));
            V3 refl = ray_trace(s, &rr, depth + 1, st);
            color = vadd(color, vmulv(refl, km));
        }
    } else {
        if (depth > 0) return v3(0, 0, 0);
        return v3((float) s->d.background[0], (float) s->d.background[1], (float) s->d.background[2]);
    }
    return v3(clampf(color.x, 0, FLT_MAX), clampf(color.y, 0, FLT_MAX), clampf(color.z, 0, FLT_MAX));
}

/* ------------------------------------------------------------------ camera (raytracer.cpp:284-325) */

typedef struct {
    V3 q, u, v, e;
    float suMul, svMul;
} EyeGen;

static EyeGen eye_init(const RtCamera *c, int width, int height) {
    EyeGen g;
    g.e = v3(c->position.x, c->position.y, c->position.z);
    V3 w = vneg(v3(c->gaze.x, c->gaze.y, c->gaze.z));
    g.v = v3(c->up.x, c->up.y, c->up.z);
    g.u = vcross(g.v, w);
    V3 m = vadd(g.e, vmulf(vneg(w), c->near_distance));
    g.q = vadd(vadd(m, vmulf(g.u, c->l)), vmulf(g.v, c->t));
    g.suMul = (c->r - c->l) / (float) width;
    g.svMul = (c->t - c->b) / (float) height;
    return g;
}

static Ray eye_ray(const EyeGen *g, long long row, long long col) {
    float su = (float) ((col + 0.5) * (double) g->suMul);
    float sv = (float) ((row + 0.5) * (double) g->svMul);
    V3 s = vsub(vadd(g->q, vmulf(g->u, su)), vmulf(g->v, sv));
    return make_ray(g->e, vsub(s, g->e));
}

/* parser.h:88-93: clamp(0,255), ::round (half away from zero), narrowing to unsigned char */
static void to_pixel(V3 c, unsigned char *px) {
    px[0] = (unsigned char) round((double) clampf(c.x, 0, 255));
    px[1] = (unsigned char) round((double) clampf(c.y, 0, 255));
    px[2] = (unsigned char) round((double) clampf(c.z, 0, 255));
}

/* ------------------------------------------------------------------ frame drivers */

typedef struct {
    const OrScene *s;
    EyeGen g;
    int aa, out_w, out_h;
    int tid, nthreads;
    unsigned char *out;
    /* row-sample mode */
    long long row0, row_stride;
    int n_rows;
    unsigned char *rows_out;
    OrStats st;
} Job;

/* raytracer.cpp:352-360 (rows dealt round-robin) fused with downSample raytracer.cpp:459-484:
 * each output pixel integer-sums its aa*aa quantised sub-samples and divides, truncating. */
static void *frame_worker(void *p) {
    Job *j = (Job *) p;
    const int f = j->aa;
    for (int row = j->tid; row < j->out_h; row += j->nthreads) {
        for (int col = 0; col < j->out_w; col++) {
            int sum[3] = {0, 0, 0};
            for (int k = 0; k < f; k++)
                for (int l = 0; l < f; l++) {
                    Ray r = eye_ray(&j->g, (long long) row * f + k, (long long) col * f + l);
                    V3 c = ray_trace(j->s, &r, 0, &j->st);
                    unsigned char px[3];
                    to_pixel(c, px);
                    sum[0] += px[0]; sum[1] += px[1]; sum[2] += px[2];
                }
            unsigned char *o = j->out + ((size_t) row * j->out_w + col) * 3;
            o[0] = (unsigned char) (sum[0] / (f * f));
            o[1] = (unsigned char) (sum[1] / (f * f));
            o[2] = (unsigned char) (sum[2] / (f * f));
        }
    }
    return NULL;
}

static void *rows_worker(void *p) {
    Job *j = (Job *) p;
    long long width = (long long) j->out_w * j->aa;
    for (int k = j->tid; k < j->n_rows; k += j->nthreads) {
        long long row = j->row0 + j->row_stride * k;
        for (long long col = 0; col < width; col++) {
            Ray r = eye_ray(&j->g, row, col);
            V3 c = ray_trace(j->s, &r, 0, &j->st);
            unsigned char px[3];
            to_pixel(c, px);
            if (j->rows_out) memcpy(j->rows_out + ((size_t) k * (size_t) width + (size_t) col) * 3, px, 3);
        }
    }
    return NULL;
}

static void run_jobs(Job *proto, int nthreads, void *(*fn)(void *), OrStats *stats) {
    if (nthreads < 1) nthreads = 1;
    Job *jobs = (Job *) calloc((size_t) nthreads, sizeof(Job));
    pthread_t *th = (pthread_t *) calloc((size_t) nthreads, sizeof(pthread_t));
    for (int i = 0; i < nthreads; i++) {
        jobs[i] = *proto;
        jobs[i].tid = i;
        jobs[i].nthreads = nthreads;
        memset(&jobs[i].st, 0, sizeof(OrStats));
        pthread_create(&th[i], NULL, fn, &jobs[i]);
    }
    OrStats tot; memset(&tot, 0, sizeof(tot));
    for (int i = 0; i < nthreads; i++) {
        pthread_join(th[i], NULL);
        uint64_t *a = (uint64_t *) &tot, *b = (uint64_t *) &jobs[i].st;
        for (size_t k = 0; k < sizeof(OrStats) / sizeof(uint64_t); k++) a[k] += b[k];
    }
    if (stats) *stats = tot;
    free(jobs); free(th);
}

/* whole frame: cam->image_width x image_height OUTPUT pixels, aa x aa sub-samples each */
int or_render(const OrScene *s, const RtCamera *cam, int aa, unsigned char *rgb_out, OrStats *stats, int nthreads) {
    if (!s || !cam || !rgb_out || aa < 1) return -1;
    Job j; memset(&j, 0, sizeof(j));
    j.s = s; j.aa = aa; j.out_w = cam->image_width; j.out_h = cam->image_height; j.out = rgb_out;
    j.g = eye_init(cam, cam->image_width * aa, cam->image_height * aa);
    run_jobs(&j, nthreads, frame_worker, stats);
    return 0;
}

/* bounded sample: sub-sample rows row0 + k*row_stride (k < n_rows) of the (w*aa) x (h*aa) grid,
 * quantised sub-samples to rows_out (may be NULL); same sample definition as ref_time_rows */
int or_render_rows(const OrScene *s, const RtCamera *cam, int aa, long long row0, long long row_stride, int n_rows,
                   unsigned char *rows_out, OrStats *stats, int nthreads) {
    if (!s || !cam || aa < 1) return -1;
    Job j; memset(&j, 0, sizeof(j));
    j.s = s; j.aa = aa; j.out_w = cam->image_width; j.out_h = cam->image_height;
    j.row0 = row0; j.row_stride = row_stride; j.n_rows = n_rows; j.rows_out = rows_out;
    j.g = eye_init(cam, cam->image_width * aa, cam->image_height * aa);
    run_jobs(&j, nthreads, rows_worker, stats);
    return 0;
}

/* the specular gate of raytracer.cpp:411-412 as a predicate on cos(theta), exported so that
 * tests can pin the product's closed-form threshold against libm's acos */
int or_specular_gate(float cosTheta) {
    float theta = (float) (acos((double) cosTheta) * 180 / 3.1415);
    return theta <= 90.01;
}

/* a block of sub-samples (debugging aid for tests: single thread, works with or_set_no_culling) */
int or_render_block(const OrScene *s, const RtCamera *cam, int aa, long long row0, long long col0, int n_rows, int n_cols,
                    unsigned char *out, OrStats *stats) {
    if (!s || !cam || aa < 1 || !out) return -1;
    EyeGen g = eye_init(cam, cam->image_width * aa, cam->image_height * aa);
    OrStats st; memset(&st, 0, sizeof st);
    for (int r = 0; r < n_rows; r++)
        for (int c = 0; c < n_cols; c++) {
            Ray ray = eye_ray(&g, row0 + r, col0 + c);
            V3 col = ray_trace(s, &ray, 0, &st);
            to_pixel(col, out + ((size_t) r * n_cols + c) * 3);
        }
    if (stats) *stats = st;
    return 0;
}

/* debugging aid: the closest hit of an arbitrary ray; out8 = exists, prim, t, then the slab test of every node on
 * the path from the root to that primitive's leaf is printed to stderr when verbose */
int or_debug_first_hit(const OrScene *s, const float *o3, const float *d3, float *t_out, int *prim_out) {
    Ray r = make_ray(v3(o3[0], o3[1], o3[2]), v3(d3[0], d3[1], d3[2]));
    OrStats st; memset(&st, 0, sizeof st);
    Hit h = first_hit(s, &r, &st);
    *t_out = h.tSmall;
    *prim_out = h.exists ? h.prim : -1;
    return h.exists;
}
/* eye ray of a sub-sample: o3, d3 */
void or_debug_eye_ray(const RtCamera *cam, int aa, long long row, long long col, float *o3, float *d3) {
    EyeGen g = eye_init(cam, cam->image_width * aa, cam->image_height * aa);
    Ray r = eye_ray(&g, row, col);
    o3[0] = r.o.x; o3[1] = r.o.y; o3[2] = r.o.z; d3[0] = r.d.x; d3[1] = r.d.y; d3[2] = r.d.z;
}
/* slab test (raytracer.cpp:101-126) of node i for a ray: returns exists, *t = entry; box6 receives the box */
int or_debug_node_box(const OrScene *s, int node, const float *o3, const float *d3, float *t, float *box6, int *meta4) {
    Ray r = make_ray(v3(o3[0], o3[1], o3[2]), v3(d3[0], d3[1], d3[2]));
    const ONode *nd = &s->nodes[node];
    box6[0] = nd->box.min.x; box6[1] = nd->box.min.y; box6[2] = nd->box.min.z;
    box6[3] = nd->box.max.x; box6[4] = nd->box.max.y; box6[5] = nd->box.max.z;
    meta4[0] = nd->is_leaf; meta4[1] = nd->axis; meta4[2] = nd->right; meta4[3] = nd->tri_count + nd->sph_count;
    return hit_box(&r, &nd->box, t);
}
int or_debug_leaf_has(const OrScene *s, int node, int prim) {
    const ONode *nd = &s->nodes[node];
    for (int i = 0; i < nd->tri_count; i++) if (s->leaf_tris[nd->tri_begin + i] == prim) return 1;
    for (int i = 0; i < nd->sph_count; i++) if (s->n_tris + s->leaf_sphs[nd->sph_begin + i] == prim) return 1;
    return 0;
}
int or_num_nodes(const OrScene *s) { return s->n_nodes; }
