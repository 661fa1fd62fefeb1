// render_v2.cu — the hot path as a warp-granular persistent kernel with a per-lane ray state machine.
//
// Replaces, per sub-sample (reference lines in brackets):
//   eye ray generation                [raytracer.cpp:319-324]
//   closest-hit BVH traversal         [raytracer.cpp:177-225]  triangle test [:129-175], sphere test [:70-96]
//   any-hit shadow traversal          [raytracer.cpp:227-280]
//   Blinn-Phong shading               [raytracer.cpp:392-427]
//   mirror recursion (made iterative) [raytracer.cpp:386-389, 430-451]
//   8-bit quantisation                [parser.h:88-93]
//   SSAA box filter on quantised data [raytracer.cpp:459-484]   (fused: no sub-sample image exists)
//
// Execution shape:
//  * the unit of work is a WARP ITEM (a few output pixels) that one warp claims from a global atomic counter —
//    no CTA barrier anywhere, warps never wait for each other;
//  * every lane runs a small state machine {IDLE, CLOSEST, SHADOW}.  One loop iteration = one ray per lane: ALL
//    lanes walk the BVH in one unified closest-hit / any-hit traversal loop, then each lane consumes its result: a
//    closest hit sets up shading and issues the first shadow ray, a shadow result adds that light's Blinn-Phong
//    terms and issues the next shadow ray, the reflection ray or the final colour.  The recursion of
//    raytracer.cpp:385-452 thus becomes iterative ray generations without materialising queues in memory: the
//    "queue entry" of a path is its lane's registers;
//  * two instantiations, chosen per launch by the host (api.cu):
//      kAccShared  any factor f: an item is P x Ph pixels, idle lanes are refilled from the item's sub-samples in
//                  8x4 blocks (ballot + popc prefix), SSAA sums live in warp-private shared memory (atomics);
//      kAccRegs    f a multiple of 8 (the 16x16 headline configuration): an item is a strip of up to 32 pixels of
//                  one row; a round hands each lane one sub-sample of ONE pixel (8x4 block inside the pixel), the
//                  lane keeps its quantised sums in registers, the warp adds them up with three REDUX when the
//                  pixel's f*f/32 rounds are done, and the finished strip leaves as 32-bit words built with
//                  shuffles (24 lanes x 4 B = 96 contiguous bytes, whole 32-byte sectors): no shared memory, no
//                  atomics (ncu: the shared-memory atomics of the first mode were 19 % of all L1 wavefronts and
//                  97 % of them bank-conflict replays, since the 32 lanes of a round add into the same pixel).
#include "device_common.cuh"

namespace rtb {

namespace {

constexpr int kWarps2 = 4;
constexpr int kThreads2 = kWarps2 * 32;
constexpr int kMaxP2 = 16;  // kAccShared: item side in pixels when f > 1 (acc size); f == 1 writes pixels directly

#ifndef RT_MIN_CTAS2
#define RT_MIN_CTAS2 7
#endif
#ifndef RT_SKIP_ZERO_SPECULAR
#define RT_SKIP_ZERO_SPECULAR 1  // A/B: 0 computes the specular term of materials without specular reflectance too
#endif
#ifndef RT_ROUND_REDUX
#define RT_ROUND_REDUX 1  // A/B: 0 keeps three per-lane colour sums across the rounds of a pixel
#endif
#ifndef RT_RELOAD_HIT
#define RT_RELOAD_HIT 1  // the hit's normal and material are re-read from the primitive id (one 128-bit load) whenever a lane
                         // consumes a ray, instead of living in four registers across the traversals: car -4 %, the other
                         // scenes -0.3..-0.7 % (A/B: 0)
#endif
#ifndef RT_FORCE_EAGER_LOOP
#define RT_FORCE_EAGER_LOOP 0  // A/B: run the refill loop (round 1's only shape) even at threshold 0
#endif

// Lane::state: < 0 idle, 0 a closest-hit ray in flight, k > 0 the shadow ray towards light k - 1 (phase and light index
// in one register)
constexpr int kIdle = -1, kClosest = 0;
enum AccMode : int { kAccShared = 0, kAccRegs = 1, kAccEager = 2 };

RT_DEV V3 clamp3(V3 c) {  // Vec3f::clamp(0, FLT_MAX), raytracer.cpp:451
    return mk(clamp_ref(c.x, 0.0f, FLT_MAX), clamp_ref(c.y, 0.0f, FLT_MAX), clamp_ref(c.z, 0.0f, FLT_MAX));
}

// per-lane path state (scalars only: the compiler keeps a struct in registers only if nothing in it is indexed
// dynamically, so the three stacks live in separate local arrays — see LaneStacks)
struct Lane {
    int state;
    int depth, hitprim;  // depth: reflection levels above this ray == levels pushed on the fold stacks
    V3 color, Pt, dn;
#if !RT_RELOAD_HIT
    int mat;
    V3 n;
#endif
    Ray ray;
    float limit;
};
struct LaneStacks {
    // reflection levels of the current path (folded back to front at the end of the path)
    V3 *local_stack;  // [kMaxSupportedDepth + 1]
    int *mat_stack;   // [kMaxSupportedDepth + 1]
    int *stack;       // [kStackSize] traversal stack
    float *sph_n;     // [3] RT_RELOAD_HIT: the normal of a hit sphere (a triangle's is re-read from tri_nm)
};

RT_DEV void start_primary(const RenderParams &p, Lane &L, V3 E0, V3 Q, V3 U, V3 Vv, int sx, int sy) {
    // raytracer.cpp:319-324 on the (nx*f) x (ny*f) sub-sample grid
    const float su = ((float) sx + 0.5f) * p.su_mul;
    const float sv = ((float) sy + 0.5f) * p.sv_mul;
    const V3 s = (Q + U * su) - Vv * sv;
    L.ray = make_ray(E0, s - E0);
    L.limit = FLT_MAX;
    L.depth = 0;
    L.state = kClosest;
}

// One ray per active lane through the BVH (closest-hit and any-hit share the loop), then the lane consumes its
// result.  Returns true when the lane's path is finished; rgb then holds the quantised sample.
template <bool FAR>
RT_DEV bool trace_step(const RenderParams &p, Lane &L, const LaneStacks &S, V3 Ia, Counters &cnt, unsigned &r8, unsigned &g8, unsigned &b8) {
    const int state = L.state;
    const bool any = state > 0;
    float tbest = L.limit;
    int pbest = -1;
    float tsecond = FLT_MAX;
    bool occluded = false;
    if (state >= 0 && p.n_nodes > 0) {
        if (p.brute_force) {
            for (int s = 0; s < p.n_prims && !occluded; s++) {
                float t;
                int prim;
                if (hit_prim(p, L.ray, s, t, prim)) {
                    if (any) {
                        if (t < L.limit) {
                            occluded = true;
                            tbest = t;
                            pbest = prim;
                        }
                    } else {
                        closest_update(p, L.ray, t, prim, tbest, pbest, tsecond);
                    }
                }
            }
        } else {
            int *sp = S.stack;  // pointer, not index: saves the index scaling on every push/pop
            *sp++ = kSentinel;
            int node = 0;  // inner references are float4 indices (4 * node), leaves negative, see rt_internal.h
            const float4 *__restrict__ nodes = p.nodes;
            while (node != kSentinel) {
                // "while-while": lanes keep descending until every lane of the warp holds a leaf (or is done), then
                // the warp runs the primitive tests together
                while ((unsigned) node < (unsigned) kSentinel) {
                    RT_CHECK((node & 3) == 0 && (node >> 2) < p.n_nodes && sp >= S.stack && sp < S.stack + kStackSize);
                    const float4 *nd = nodes + (unsigned) node;
                    float4 n0, n1, n2, n3;
                    ld256(nd, n0, n1);
                    ld256(nd + 2, n2, n3);
                    float tmin0, tmax0, tmin1, tmax1;
                    slab(L.ray, n0.x, n0.y, n0.z, n0.w, n2.x, n2.y, tmin0, tmax0);
                    slab(L.ray, n1.x, n1.y, n1.z, n1.w, n2.z, n2.w, tmin1, tmax1);
                    // visit iff the ray overlaps the box for t >= 0 (raytracer.cpp:120) and the entry is not beyond
                    // the current limit (raytracer.cpp:188; non-strict so equal-t ties are still seen)
                    bool h0, h1;
                    if (FAR) {
                        // camera far outside the scene: c*inv - o*inv cancels catastrophically (error ~ |o*inv| 2^-23,
                        // which outgrows the boxes' padding beyond ~30 scene diagonals): widen the interval by it
                        const float er = ray_slack(L.ray);
                        h0 = tmax0 + er >= fmaxf(tmin0 - er, 0.0f) && tmin0 - er <= tbest;
                        h1 = tmax1 + er >= fmaxf(tmin1 - er, 0.0f) && tmin1 - er <= tbest;
                    } else {
                        h0 = tmax0 >= fmaxf(tmin0, 0.0f) && tmin0 <= tbest;
                        h1 = tmax1 >= fmaxf(tmin1, 0.0f) && tmin1 <= tbest;
                    }
                    const int c0 = __float_as_int(n3.x), c1 = __float_as_int(n3.y);
                    // select-based step: one predicated push, one predicated pop, no 4-way branch
                    const bool swap = tmin1 < tmin0;
                    const bool take1 = h1 && (!h0 || swap);
                    if (h0 && h1) *sp++ = swap ? c0 : c1;
                    node = take1 ? c1 : c0;
                    if (!(h0 || h1)) node = *--sp;
                }
                if (node < 0) {
                    const int enc = ~node;
                    const int first = enc >> 3, count = (enc & 7) + 1;
                    RT_CHECK(first >= 0 && first + count <= p.n_prims + 1 && sp > S.stack);
                    node = *--sp;
                    for (int s = first; s < first + count; s++) {
                        float t;
                        int prim;
                        if (hit_prim(p, L.ray, s, t, prim)) {
                            if (any) {
                                if (t < L.limit) {  // raytracer.cpp:237, 245
                                    occluded = true;
                                    tbest = t;
                                    pbest = prim;
                                    node = kSentinel;
                                    break;
                                }
                            } else {
                                closest_update(p, L.ray, t, prim, tbest, pbest, tsecond);
                            }
                        }
                    }
                }
            }
        }
    }

    // reference visibility: a doubtful hit is replayed on the reference's own tree (device_common.cuh)
    if (p.exact_culling && state >= 0 && pbest >= 0 && (p.exact_culling == 2 || !robust_visible(p, L.ray, pbest, tbest, any ? FLT_MAX : tsecond))) {
        if (any) {
            cnt.replay_any++;
            occluded = ref_any(p, L.ray, L.limit);
        } else {
            cnt.replay_closest++;
            ref_closest(p, L.ray, tbest, pbest);
        }
    }

    // ---- consume the result ----------------------------------------------------------------------
    bool lights_done = false;
    bool finish = false;
    V3 result = mk(0.0f, 0.0f, 0.0f);

    int next_light = 0;
    V3 n = mk(0.0f, 0.0f, 0.0f);  // normal and material of the surface this lane is shading
    int mat = 1;
#if RT_RELOAD_HIT
    if (state > 0) {
        if (L.hitprim < p.n_tris) {
            const float4 nm = __ldg(&p.tri_nm[L.hitprim]);
            n = xyz(nm);
            mat = __float_as_int(nm.w);
        } else {
            n = mk(S.sph_n[0], S.sph_n[1], S.sph_n[2]);
            mat = __ldg(&p.sph_mat[L.hitprim - p.n_tris]);
        }
    }
#else
    if (state > 0) n = L.n, mat = L.mat;
#endif
    if (state == kClosest) {
        if (pbest < 0) {  // raytracer.cpp:442-449
            result = L.depth > 0 ? mk(0.0f, 0.0f, 0.0f) : ld3(p.background);
            finish = true;
        } else {
            if (pbest < p.n_tris) {
                const float4 nm = __ldg(&p.tri_nm[pbest]);
                n = xyz(nm);
                mat = __float_as_int(nm.w);
            } else {
                const float4 cr = __ldg(&p.sph_cr[pbest - p.n_tris]);
                mat = __ldg(&p.sph_mat[pbest - p.n_tris]);
                n = normalize((((L.ray.o + L.ray.d * tbest) - xyz(cr)) / cr.w));  // raytracer.cpp:91
#if RT_RELOAD_HIT
                S.sph_n[0] = n.x, S.sph_n[1] = n.y, S.sph_n[2] = n.z;
#endif
            }
#if !RT_RELOAD_HIT
            L.n = n, L.mat = mat;
#endif
            const float4 m0 = __ldg(&p.materials[4 * (mat - 1)]);
            L.color = mk(0.0f, 0.0f, 0.0f) + mulv(xyz(m0), Ia);  // raytracer.cpp:394-395
            L.Pt = L.ray.o + L.ray.d * tbest;
            L.hitprim = pbest;
            L.dn = normalize(L.ray.d);
            lights_done = p.n_lights == 0;
        }
    } else if (state > 0) {
        const int light = state - 1;
        if (occluded) {
            cnt.occluded++;
        } else {  // raytracer.cpp:406-423
            const V3 lpos = xyz(__ldg(&p.lights[2 * light]));
            const V3 I = xyz(__ldg(&p.lights[2 * light + 1]));
            const float4 m0 = __ldg(&p.materials[4 * (mat - 1)]);
            const float4 m1 = __ldg(&p.materials[4 * (mat - 1) + 1]);
            const V3 wi = L.ray.d;
            const float dist = L.limit;
            const V3 wiReal = normalize(lpos - L.Pt);
            const float cosTheta = dot(wiReal, n);
            const V3 E = I / (dist * dist);
#if RT_SKIP_ZERO_SPECULAR
            // raytracer.cpp:411-418 with ks == (0, 0, 0): the term is (ks * pow(..)) (.) E = (+-0) (.) E, which is +-0 in every
            // component when E is finite, and colour + (+-0) == colour bit for bit (colour is never -0: it starts as
            // +0 + ambient).  pow() cannot make it NaN: its base is max(0, .) of two unit vectors' dot product (NaN -> 0 by
            // std::max's operand order) and the exponent is checked on the host (scene_build.cu).  With a non-finite E
            // (a light ON the surface point) the full path below runs.  On horse_and_mug this skips one normalisation, one
            // pow and two loads for every unoccluded shadow ray from the floor.
            const bool no_spec = (__float_as_int(m1.w) & 2) && (fabsf(E.x) + fabsf(E.y) + fabsf(E.z) <= FLT_MAX);
            if (!no_spec && specular_gate(cosTheta)) {
#else
            if (specular_gate(cosTheta)) {
#endif
                const float4 m2 = __ldg(&p.materials[4 * (mat - 1) + 2]);
                const V3 h = normalize(wi + (-L.dn));
                const float c = pow_ref(std_max(0.0f, dot(L.hitprim < p.n_tris ? xyz(__ldg(&p.tri_nn[L.hitprim])) : normalize(n), h)), m0.w);
                L.color = L.color + mulv(xyz(m2) * c, E);
            }
            const float cd = std_max(0.0f, std_min(1.0f, cosTheta));
            L.color = L.color + mulv(xyz(m1) * cd, E);
        }
        next_light = state;
        lights_done = next_light >= p.n_lights;
    }

    if (state >= 0 && !finish) {
        const V3 Pe = L.Pt + n * p.eps;  // raytracer.cpp:397
        if (!lights_done) {                // raytracer.cpp:399-404: shadow ray towards light `light`
            const V3 lpos = xyz(__ldg(&p.lights[2 * next_light]));
            const V3 toL = lpos - Pe;
            const float dist = length(toL);
            L.ray = make_ray(Pe, toL / dist);
            L.limit = dist;
            L.state = next_light + 1;
            cnt.shadow++;
        } else {
            const float4 m1 = __ldg(&p.materials[4 * (mat - 1) + 1]);
            if (__float_as_int(m1.w) & 1) {  // mirror: raytracer.cpp:430-439
                RT_CHECK(L.depth >= 0 && L.depth <= kMaxSupportedDepth);
                S.local_stack[L.depth] = L.color;
                S.mat_stack[L.depth] = mat;
                const V3 nn = L.hitprim < p.n_tris ? xyz(__ldg(&p.tri_nn[L.hitprim])) : normalize(n);
                const float rc = dot(-L.dn, nn);
                L.depth++;
                if (L.depth > p.max_depth) {  // raytracer.cpp:387-389
                    result = mk(0.0f, 0.0f, 0.0f);
                    finish = true;
                } else {
                    L.ray = make_ray(Pe, L.dn + (nn * 2.0f) * rc);
                    L.limit = FLT_MAX;
                    L.state = kClosest;
                    cnt.reflection++;
                }
            } else {
                result = clamp3(L.color);
                finish = true;
            }
        }
    }

    if (finish) {
        int npush = L.depth;  // (a path cut off by the depth limit has pushed its last level too: depth was incremented)
        while (npush > 0) {  // fold the mirror levels back to front
            npush--;
            const V3 km = xyz(__ldg(&p.materials[4 * (S.mat_stack[npush] - 1) + 3]));
            result = clamp3(S.local_stack[npush] + mulv(result, km));
        }
        r8 = quantise(result.x), g8 = quantise(result.y), b8 = quantise(result.z);
        L.state = kIdle;
    }
    return finish;
}

RT_DEV unsigned char *pixel_ptr(const RenderParams &p, int local_band, int y_in_band, int px, int py) {
    RT_CHECK(px >= 0 && px < p.nx && py >= 0 && py < p.ny && local_band >= 0 && local_band < p.n_bands && y_in_band >= 0 && y_in_band < p.Ph);
    if (p.out_mode == kOutFrame) return p.out + ((size_t) py * p.nx + px) * 3;
    return p.out + (((size_t) local_band * p.Ph + y_in_band) * p.nx + px) * 3;
}

}  // namespace

template <int ACC, bool FAR>
__global__ void __launch_bounds__(kThreads2, RT_MIN_CTAS2) render_kernel_v2(const __grid_constant__ RenderParams p) {
    // kAccShared: warp-private SSAA accumulators, sized per launch (P*Ph*3 words per warp; nothing when f == 1):
    // whatever shared memory the kernel does not need stays L1 cache for the BVH
    extern __shared__ unsigned acc_all[];

    const int lane = threadIdx.x & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    (void) lt_mask;
    const int f = p.f, P = p.P;
    const unsigned per_tile = (unsigned) (p.tile_items * p.group_bands), per_group = per_tile * (unsigned) p.tiles_per_group;
    unsigned *work_counter = (unsigned *) p.control;
    const V3 E0 = ld3(p.e), Q = ld3(p.q), U = ld3(p.u), Vv = ld3(p.v);
    const V3 Ia = ld3(p.ambient);
    Counters cnt = {0u, 0u, 0u, 0u, 0u, 0u};
    Lane L;
    V3 local_stack[kMaxSupportedDepth + 1];
    int mat_stack[kMaxSupportedDepth + 1];
    int stack[kStackSize];
    float sph_n[3];
    const LaneStacks S = {local_stack, mat_stack, stack, sph_n};

    if (ACC == kAccRegs) {
        // ---- f = 8 * bpx = 4 * bpy: strips of up to 32 pixels of one row, GUIDED self-scheduling ---------------------
        // The part's pixels are numbered in block order — index = ((group of 32 rows * tiles_x + 32-pixel column) * 32 + row
        // in group) * 32 + pixel in column — and a warp claims the next `n` of them with one atomic, n = 32 while plenty
        // is left, then 16, 8, ... 1 as the part runs out (n ~ remaining / (4 x warps in flight)): big items while
        // throughput matters, single pixels when only the tail is left.  (Fixed 32-pixel strips cost 0.6 % on the full 8K
        // frame and 9 % on a 1/8 part; fixed 8-pixel strips 1.8 % on the part: tools/partition_experiment.py.)
        const int bpx = f >> 3, bpy = f >> 2;
        const unsigned ff = (unsigned) (f * f);
        const unsigned total = p.n_items, tiles_x = (unsigned) p.tiles_per_group;
        const unsigned bw = (unsigned) p.blk_w_log2, bh = 10u - bw;  // block of 1024 pixel slots: 2^bw wide, 2^bh (local) rows high
        for (;;) {
            unsigned start = 0, n = 0;
            if (lane == 0) {
                const unsigned cur = *(volatile unsigned *) work_counter;
                const unsigned c = (cur < total ? total - cur : 0u) / p.guide;
                n = c >= (unsigned) p.P ? (unsigned) p.P : (c >= 1u ? 1u << (31 - __clz(c)) : 1u);
                start = atomicAdd(work_counter, n);
            }
            start = __shfl_sync(0xffffffffu, start, 0);
            n = __shfl_sync(0xffffffffu, n, 0);
            if (start >= total) break;
            n = min(n, total - start);
            // lane i < n owns pixel start + i
            const unsigned idx = start + (unsigned) lane;
            const unsigned t = idx >> 10;
            const int local_row = (int) (((t / tiles_x) << bh) + ((idx >> bw) & ((1u << bh) - 1u)));
            const int x = (int) (((t % tiles_x) << bw) + (idx & ((1u << bw) - 1u)));
            // local row -> global row: bands of rows_per_band rows, band b of this part = part_rank + b * part_world
            const int y = (p.part_rank + (local_row / p.rows_per_band) * p.part_world) * p.rows_per_band + local_row % p.rows_per_band;
            const bool valid = lane < (int) n && local_row < p.n_bands && x < p.nx && y < p.ny;
            L.state = kIdle;
            unsigned mine = 0u;  // B << 16 | G << 8 | R of this lane's pixel
            for (int pix = 0; pix < (int) n; pix++) {
                if (!__shfl_sync(0xffffffffu, (int) valid, pix)) continue;
                const int sx0 = __shfl_sync(0xffffffffu, x, pix) * f + (lane & 7), sy0 = __shfl_sync(0xffffffffu, y, pix) * f + (lane >> 3);
#if RT_ROUND_REDUX
                // one packed register per lane (the quantised sample of this round) and per-pixel sums the warp adds up after
                // every round, instead of three per-lane sums alive across every traversal: ncu showed those spilled and
                // re-loaded around each ray (6 of ~25 local/global memory instructions per ray); -1.9 % on the bench frame
                unsigned tr = 0u, tg = 0u, tb = 0u;
                for (int by = 0; by < bpy; by++) {
                    for (int bx = 0; bx < bpx; bx++) {
                        start_primary(p, L, E0, Q, U, Vv, sx0 + bx * 8, sy0 + by * 4);
                        cnt.primary += p.max_depth >= 0;  // a ray is a closest-hit query (raytracer.cpp:387: none when the depth limit is negative)
                        unsigned rgb = 0u;
                        do {
                            unsigned r8, g8, b8;
                            if (trace_step<FAR>(p, L, S, Ia, cnt, r8, g8, b8)) rgb = r8 | (g8 << 8) | (b8 << 16);
                        } while (__any_sync(0xffffffffu, L.state >= 0));
                        tr += __reduce_add_sync(0xffffffffu, rgb & 0xffu);
                        tg += __reduce_add_sync(0xffffffffu, (rgb >> 8) & 0xffu);
                        tb += __reduce_add_sync(0xffffffffu, rgb >> 16);
                    }
                }
                // raytracer.cpp:475-477: truncating integer average of the quantised sub-samples
                const unsigned R = tr / ff, G = tg / ff, B = tb / ff;
#else
                unsigned sr = 0u, sg = 0u, sb = 0u;
                for (int by = 0; by < bpy; by++) {
                    for (int bx = 0; bx < bpx; bx++) {
                        start_primary(p, L, E0, Q, U, Vv, sx0 + bx * 8, sy0 + by * 4);
                        cnt.primary += p.max_depth >= 0;  // a ray is a closest-hit query (raytracer.cpp:387: none when the depth limit is negative)
                        do {
                            unsigned r8, g8, b8;
                            if (trace_step<FAR>(p, L, S, Ia, cnt, r8, g8, b8)) sr += r8, sg += g8, sb += b8;
                        } while (__any_sync(0xffffffffu, L.state >= 0));
                    }
                }
                // raytracer.cpp:475-477: truncating integer average of the quantised sub-samples
                const unsigned R = __reduce_add_sync(0xffffffffu, sr) / ff, G = __reduce_add_sync(0xffffffffu, sg) / ff,
                               B = __reduce_add_sync(0xffffffffu, sb) / ff;
#endif
                if (lane == pix) mine = R | (G << 8) | (B << 16);
            }
            unsigned char *o = p.out_mode == kOutFrame ? p.out + ((size_t) y * p.nx + x) * 3 : p.out + ((size_t) local_row * p.nx + x) * 3;
            RT_CHECK(!valid || (x >= 0 && x < p.nx && y >= 0 && y < p.ny && local_row >= 0 && local_row < p.n_bands));
            // an aligned run of 4k pixels of one row leaves as 3k 32-bit words built with two shuffles each
            // (word w = bytes 4w .. 4w+3 = pixels (4w)/3 and (4w)/3 + 1, shifted by w % 3 bytes)
            const bool whole = (n & 3u) == 0u && (start & (n - 1u)) == 0u && __all_sync(0xffffffffu, valid || lane >= (int) n);
            unsigned char *row = (unsigned char *) __shfl_sync(0xffffffffu, (unsigned long long) o, 0);
            if (whole && (((size_t) row) & 3) == 0) {
                const int p0 = (4 * lane) / 3;
                const unsigned lo = __shfl_sync(0xffffffffu, mine, p0 & 31), hi = __shfl_sync(0xffffffffu, mine, (p0 + 1) & 31);
                const unsigned long long two = (unsigned long long) lo | ((unsigned long long) hi << 24);
                if (lane < (int) (n * 3u / 4u)) ((unsigned *) row)[lane] = (unsigned) (two >> (8 * (lane % 3)));
            } else if (valid) {
                o[0] = (unsigned char) mine;
                o[1] = (unsigned char) (mine >> 8);
                o[2] = (unsigned char) (mine >> 16);
            }
        }
    } else {
      for (;;) {
        unsigned item = 0;
        if (lane == 0) item = atomicAdd(work_counter, 1u);
        item = __shfl_sync(0xffffffffu, item, 0);
        if (item >= p.n_items) break;

        // item -> (band of this part, item within the band): consecutive items cover a compact block (render_params.h)
        const unsigned g = item / per_group, r = item % per_group;
        const unsigned tc = r / per_tile, q = r % per_tile;
        const int local_band = (int) (g * (unsigned) p.group_bands + q / (unsigned) p.tile_items);
        const int ix = (int) (tc * (unsigned) p.tile_items + q % (unsigned) p.tile_items);
        if (local_band >= p.n_bands || ix >= p.items_x) continue;
        // local item row -> global item row: bands of p.rows_per_band item rows, band b of this part = part_rank + b * part_world
        const int band = local_band / p.rows_per_band, in_band = local_band % p.rows_per_band;
        const int px0 = ix * P, py0 = ((p.part_rank + band * p.part_world) * p.rows_per_band + in_band) * p.Ph;
        const int pw = min(P, p.nx - px0);
        const int ph = min(p.Ph, p.ny - py0);
        if (pw <= 0 || ph <= 0) continue;
        L.state = kIdle;
        {
            // ---- P x Ph pixels, sub-samples in 8x4 blocks, lanes refilled from the item ------------------------
            unsigned *acc = acc_all + (threadIdx.x >> 5) * (P * p.Ph * 3);
            const int sw = pw * f, sh = ph * f;
            const int nbx = (sw + 7) >> 3, nby = (sh + 3) >> 2;
            const int total = nbx * nby * 32;  // sub-sample slots, 8x4 blocks in row-major block order
            if (f > 1) {
                for (int i = lane; i < pw * ph * 3; i += 32) acc[i] = 0u;
                __syncwarp();
            }
            auto sample_done = [&](int lx, int ly, unsigned r8, unsigned g8, unsigned b8) {
                if (f == 1) {
                    unsigned char *o = pixel_ptr(p, local_band, ly, px0 + lx, py0 + ly);
                    o[0] = (unsigned char) r8;
                    o[1] = (unsigned char) g8;
                    o[2] = (unsigned char) b8;
                } else {
                    RT_CHECK(((ly / f) * pw + (lx / f)) < P * p.Ph);
                    unsigned *a = &acc[((ly / f) * pw + (lx / f)) * 3];
                    atomicAdd(a, r8);
                    atomicAdd(a + 1, g8);
                    atomicAdd(a + 2, b8);
                }
            };
            if (ACC == kAccShared) {
                // rounds in lock-step: every lane takes one sub-sample of the next 8x4 block, the warp traces until all of
                // its paths are finished, then moves on (the measured best: rays of one kind and neighbouring stay together)
                for (int b = 0; b < nbx * nby; b++) {
                    const int lx = (b % nbx) * 8 + (lane & 7), ly = (b / nbx) * 4 + (lane >> 3);
                    if (lx < sw && ly < sh) {
                        start_primary(p, L, E0, Q, U, Vv, px0 * f + lx, py0 * f + ly);
                        cnt.primary += p.max_depth >= 0;
                    }
                    do {
                        unsigned r8, g8, b8;
                        if (trace_step<FAR>(p, L, S, Ia, cnt, r8, g8, b8)) sample_done(lx, ly, r8, g8, b8);
                    } while (__any_sync(0xffffffffu, L.state >= 0));
                }
            } else {
                // kAccEager (experiments, RtBuildOptions.refill_threshold > 0): idle lanes are refilled with the item's next
                // sub-samples as soon as at most `refill_threshold` lanes are still busy (ballot + popc prefix = warp-level
                // work stealing; 31 = refill eagerly).  Measured 17 % slower than lock-step rounds on large frames (ray kinds
                // mix, coherence is lost) and within 1 % on small ones.
                int next = 0;        // warp-uniform: next unassigned slot of this item
                int lx = 0, ly = 0;  // the lane's sub-sample within the item
                for (;;) {
                    const unsigned idle_mask = __ballot_sync(0xffffffffu, L.state < 0);
                    if (idle_mask != 0u && next < total && 32 - __popc(idle_mask) <= p.refill_threshold) {
                        const int my = next + __popc(idle_mask & lt_mask);
                        next += __popc(idle_mask);
                        if (L.state < 0 && my < total) {
                            const int b = my >> 5, l = my & 31;
                            lx = (b % nbx) * 8 + (l & 7);
                            ly = (b / nbx) * 4 + (l >> 3);
                            if (lx < sw && ly < sh) {
                                start_primary(p, L, E0, Q, U, Vv, px0 * f + lx, py0 * f + ly);
                                cnt.primary += p.max_depth >= 0;
                            }
                        }
                    }
                    if (idle_mask == 0xffffffffu && __ballot_sync(0xffffffffu, L.state >= 0) == 0u) {
                        if (next >= total) break;
                        continue;
                    }
                    unsigned r8, g8, b8;
                    if (trace_step<FAR>(p, L, S, Ia, cnt, r8, g8, b8)) sample_done(lx, ly, r8, g8, b8);
                }
            }
            if (f > 1) {  // raytracer.cpp:475-477: truncating integer average of the quantised sub-samples
                __syncwarp();
                const unsigned ff = (unsigned) (f * f);
                for (int i = lane; i < pw * ph; i += 32) {
                    const int x = i % pw, y = i / pw;
                    const unsigned *a = &acc[i * 3];
                    unsigned char *o = pixel_ptr(p, local_band, y, px0 + x, py0 + y);
                    o[0] = (unsigned char) (a[0] / ff);
                    o[1] = (unsigned char) (a[1] / ff);
                    o[2] = (unsigned char) (a[2] / ff);
                }
                __syncwarp();
            }
        }
      }
    }

    const unsigned v0 = __reduce_add_sync(0xffffffffu, cnt.primary);
    const unsigned v1 = __reduce_add_sync(0xffffffffu, cnt.reflection);
    const unsigned v2 = __reduce_add_sync(0xffffffffu, cnt.shadow);
    const unsigned v3 = __reduce_add_sync(0xffffffffu, cnt.occluded);
    const unsigned v4 = __reduce_add_sync(0xffffffffu, cnt.replay_closest);
    const unsigned v5 = __reduce_add_sync(0xffffffffu, cnt.replay_any);
    if (lane == 0) {
        atomicAdd(&p.control[1], (unsigned long long) v0);
        atomicAdd(&p.control[2], (unsigned long long) v1);
        atomicAdd(&p.control[3], (unsigned long long) v2);
        atomicAdd(&p.control[4], (unsigned long long) v3);
        if (v4) atomicAdd(&p.control[5], (unsigned long long) v4);
        if (v5) atomicAdd(&p.control[6], (unsigned long long) v5);
    }
}

int launch_render_v2(const RenderParams &p, int n_ctas, cudaStream_t stream) {
    const size_t smem = (p.acc_mode == kAccShared && p.f > 1) ? (size_t) kWarps2 * p.P * p.Ph * 3 * sizeof(unsigned) : 0;
    if (p.acc_mode == kAccRegs) {
        if (p.far_camera) render_kernel_v2<kAccRegs, true><<<n_ctas, kThreads2, 0, stream>>>(p);
        else render_kernel_v2<kAccRegs, false><<<n_ctas, kThreads2, 0, stream>>>(p);
    } else if (p.refill_threshold > 0 || RT_FORCE_EAGER_LOOP) {
        if (p.far_camera) render_kernel_v2<kAccEager, true><<<n_ctas, kThreads2, smem, stream>>>(p);
        else render_kernel_v2<kAccEager, false><<<n_ctas, kThreads2, smem, stream>>>(p);
    } else {
        if (p.far_camera) render_kernel_v2<kAccShared, true><<<n_ctas, kThreads2, smem, stream>>>(p);
        else render_kernel_v2<kAccShared, false><<<n_ctas, kThreads2, smem, stream>>>(p);
    }
    return (int) cudaGetLastError();
}

// resident CTAs per SM of the two accumulator modes (the shared-memory one at its largest accumulator); querying
// the attributes also makes the runtime load the kernels (CUDA loads functions lazily on first use)
int render_kernel_v2_occupancy(int *ctas_per_sm, int *warps_per_cta) {
    *warps_per_cta = kWarps2;
    int a = 0, b = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a, render_kernel_v2<kAccShared, false>, kThreads2,
                                                                  (size_t) kWarps2 * kMaxP2 * kMaxP2 * 3 * sizeof(unsigned));
    if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, render_kernel_v2<kAccRegs, false>, kThreads2, 0);
    ctas_per_sm[0] = a;
    ctas_per_sm[1] = b;
    return (int) e;
}

}  // namespace rtb
