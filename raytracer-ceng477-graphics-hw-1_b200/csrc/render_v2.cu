// render_v2.cu — the hot path as a warp-granular persistent kernel with a per-lane ray state machine.
//
// Same arithmetic as render.cu (the exact tests and shading of device_common.cuh; reference lines
// cited there), different execution shape:
//
//  * the unit of work is a WARP TILE (P x P output pixels, P*f ~ 16..32 sub-samples a side) that one
//    warp claims from a global atomic counter — no CTA barrier anywhere, warps never wait for each other;
//  * every lane runs a small state machine {IDLE, CLOSEST, SHADOW}.  One loop iteration = one ray per
//    lane: idle lanes are refilled with the tile's next sub-samples (ballot + popc prefix = warp-level
//    work stealing, the warp stays full until the tile runs dry), then ALL lanes walk the BVH in one
//    unified closest-hit / any-hit traversal loop, then each lane consumes its result: a closest hit
//    sets up shading and issues the first shadow ray, a shadow result adds that light's Blinn-Phong
//    terms and issues the next shadow ray, the reflection ray or the final colour.  The recursion of
//    raytracer.cpp:385-452 thus becomes iterative ray generations without materialising queues in
//    memory: the "queue entry" of a path is its lane's registers;
//  * SSAA sums live in warp-private shared memory (one warp owns all sub-samples of its pixels), are
//    divided with truncation and stored as RGB8 by the same warp (raytracer.cpp:466-477, fused).
#include "device_common.cuh"

namespace rtb {

namespace {

constexpr int kWarps2 = 4;
constexpr int kThreads2 = kWarps2 * 32;
constexpr int kMaxP2 = 16;  // warp tile side in pixels when f > 1 (acc size); f == 1 writes pixels directly

#ifndef RT_MIN_CTAS2
#define RT_MIN_CTAS2 7
#endif
#ifndef RT_SMEM_TOP
#define RT_SMEM_TOP 0  // experiment: keep the first RT_SMEM_TOP nodes (breadth-first top of the tree) in shared memory
#endif
#ifndef RT_BRANCHLESS_STEP
#define RT_BRANCHLESS_STEP 1
#endif
#ifndef RT_WHILE_WHILE
#define RT_WHILE_WHILE 1
#endif

enum Phase : int { kIdle = 0, kClosest = 1, kShadow = 2 };

RT_DEV V3 clamp3(V3 c) {  // Vec3f::clamp(0, FLT_MAX), raytracer.cpp:451
    return mk(clamp_ref(c.x, 0.0f, FLT_MAX), clamp_ref(c.y, 0.0f, FLT_MAX), clamp_ref(c.z, 0.0f, FLT_MAX));
}

}  // namespace

__global__ void __launch_bounds__(kThreads2, RT_MIN_CTAS2) render_kernel_v2(const __grid_constant__ RenderParams p) {
    // warp-private SSAA accumulators, sized per launch (P*P*3 words per warp; nothing when f == 1): whatever shared
    // memory the kernel does not need stays L1 cache for the BVH
    extern __shared__ unsigned acc_all[];
#if RT_SMEM_TOP
    __shared__ float4 s_top[4 * RT_SMEM_TOP];
    for (int i = threadIdx.x; i < 4 * min(RT_SMEM_TOP, p.n_nodes); i += kThreads2) s_top[i] = __ldg(&p.nodes[i]);
    __syncthreads();
#endif

    const int lane = threadIdx.x & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    const int f = p.f, P = p.P;
    unsigned *acc = acc_all + (threadIdx.x >> 5) * (P * P * 3);
    const int items_per_tile = p.items_x * p.items_y;
    const V3 E0 = ld3(p.e), Q = ld3(p.q), U = ld3(p.u), Vv = ld3(p.v);
    const V3 Ia = ld3(p.ambient);
    Counters cnt = {0u, 0u, 0u, 0u, 0u, 0u};

    // reflection levels of the lane's current path (folded back to front at the end of the path)
    V3 local_stack[kMaxSupportedDepth + 1];
    int mat_stack[kMaxSupportedDepth + 1];
    int stack[kStackSize];

    for (;;) {
        unsigned item = 0;
        if (lane == 0) item = atomicAdd(p.work_counter, 1u);
        item = __shfl_sync(0xffffffffu, item, 0);
        if (item >= p.n_items) break;

        const int local_tile = (int) (item / (unsigned) items_per_tile);
        const int sub = (int) (item % (unsigned) items_per_tile);
        const int tile = p.part_rank + local_tile * p.part_world;
        const int tx0 = (tile % p.tiles_x) * RT_TILE, ty0 = (tile / p.tiles_x) * RT_TILE;
        const int ix0 = (sub % p.items_x) * P, iy0 = (sub / p.items_x) * p.Ph;
        const int px0 = tx0 + ix0, py0 = ty0 + iy0;
        const int pw = max(0, min(min(P, RT_TILE - ix0), p.nx - px0));
        const int ph = max(0, min(min(p.Ph, RT_TILE - iy0), p.ny - py0));
        if (pw == 0 || ph == 0) continue;
        const int sw = pw * f, sh = ph * f;
        const int nbx = (sw + 7) >> 3, nby = (sh + 3) >> 2;
        const int total = nbx * nby * 32;  // sub-sample slots, 8x4 blocks in row-major block order

        if (f > 1) {
            for (int i = lane; i < pw * ph * 3; i += 32) acc[i] = 0u;
            __syncwarp();
        }

        // ---- per-lane path state -----------------------------------------------------------------
        int phase = kIdle;
        int next = 0;        // warp-uniform: next unassigned slot of this tile
        int lx = 0, ly = 0;  // the lane's sub-sample within the tile
        int depth = 0, npush = 0, light = 0, mat = 0, hitprim = 0;
        V3 color = mk(0.f, 0.f, 0.f), Pt = color, n = color, dn = color;
        Ray ray = make_ray(E0, mk(0.f, 0.f, -1.f));
        float limit = FLT_MAX;

        for (;;) {
            // ---- 1. refill idle lanes with the tile's next sub-samples (warp-level work stealing) ----
            // Policy: refill only once at most `refill_threshold` lanes are still busy, so that the rays a
            // warp traces together stay of one kind and neighbouring (coherent BVH walks, few divergent
            // branches); 31 = refill eagerly, 0 = only when the whole warp has drained.
            const unsigned idle_mask = __ballot_sync(0xffffffffu, phase == kIdle);
            if (idle_mask != 0u && next < total && 32 - __popc(idle_mask) <= p.refill_threshold) {
                const int my = next + __popc(idle_mask & lt_mask);
                next += __popc(idle_mask);
                if (phase == kIdle && my < total) {
                    const int b = my >> 5, l = my & 31;
                    lx = (b % nbx) * 8 + (l & 7);
                    ly = (b / nbx) * 4 + (l >> 3);
                    if (lx < sw && ly < sh) {
                        // raytracer.cpp:319-324 on the (nx*f) x (ny*f) sub-sample grid
                        const float su = ((float) (px0 * f + lx) + 0.5f) * p.su_mul;
                        const float sv = ((float) (py0 * f + ly) + 0.5f) * p.sv_mul;
                        const V3 s = (Q + U * su) - Vv * sv;
                        ray = make_ray(E0, s - E0);
                        limit = FLT_MAX;
                        depth = 0;
                        npush = 0;
                        phase = kClosest;
                        cnt.primary++;
                    }
                }
            }
            if (idle_mask == 0xffffffffu && __ballot_sync(0xffffffffu, phase != kIdle) == 0u) {
                if (next >= total) break;
                continue;
            }

            // ---- 2. one ray per active lane through the BVH (closest-hit and any-hit share the loop) ----
            const bool any = phase == kShadow;
            float tbest = limit;
            int pbest = -1;
            float tsecond = FLT_MAX;
            bool occluded = false;
            if (phase != kIdle && p.n_nodes > 0) {
                if (p.brute_force) {
                    for (int s = 0; s < p.n_prims && !occluded; s++) {
                        float t;
                        int prim;
                        if (hit_prim(p, ray, s, t, prim)) {
                            if (any) {
                                if (t < limit) {
                                    occluded = true;
                                    tbest = t;
                                    pbest = prim;
                                }
                            } else {
                                closest_update(p, ray, t, prim, tbest, pbest, tsecond);
                            }
                        }
                    }
                } else {
                    int *sp = stack;  // pointer, not index: saves the index scaling on every push/pop
                    *sp++ = kSentinel;
                    int node = 0;
                    while (node != kSentinel) {
#if RT_WHILE_WHILE
                        // "while-while": lanes keep descending until every lane of the warp holds a leaf (or is
                        // done), then the warp runs the primitive tests together
                        while ((unsigned) node < (unsigned) kSentinel) {
#else
                        if (node >= 0) {
#endif
#if RT_SMEM_TOP
                            float4 n0, n1, n2, n3;
                            if (node < RT_SMEM_TOP) {
                                n0 = s_top[4 * node], n1 = s_top[4 * node + 1], n2 = s_top[4 * node + 2], n3 = s_top[4 * node + 3];
                            } else {
                                n0 = __ldg(&p.nodes[4 * node]), n1 = __ldg(&p.nodes[4 * node + 1]);
                                n2 = __ldg(&p.nodes[4 * node + 2]), n3 = __ldg(&p.nodes[4 * node + 3]);
                            }
#else
                            const float4 n0 = __ldg(&p.nodes[4 * node]);
                            const float4 n1 = __ldg(&p.nodes[4 * node + 1]);
                            const float4 n2 = __ldg(&p.nodes[4 * node + 2]);
                            const float4 n3 = __ldg(&p.nodes[4 * node + 3]);
#endif
                            float tmin0, tmax0, tmin1, tmax1;
                            slab(ray, n0.x, n0.y, n0.z, n0.w, n2.x, n2.y, tmin0, tmax0);
                            slab(ray, n1.x, n1.y, n1.z, n1.w, n2.z, n2.w, tmin1, tmax1);
                            const bool h0 = tmax0 >= fmaxf(tmin0, 0.0f) && tmin0 <= tbest;
                            const bool h1 = tmax1 >= fmaxf(tmin1, 0.0f) && tmin1 <= tbest;
                            const int c0 = __float_as_int(n3.x), c1 = __float_as_int(n3.y);
#if RT_BRANCHLESS_STEP
                            // select-based step: one predicated push, one predicated pop, no 4-way branch
                            const bool swap = tmin1 < tmin0;
                            const bool take1 = h1 && (!h0 || swap);
                            if (h0 && h1) *sp++ = swap ? c0 : c1;
                            node = take1 ? c1 : c0;
                            if (!(h0 || h1)) node = *--sp;
#else
                            if (h0 && h1) {
                                const bool swap = tmin1 < tmin0;
                                node = swap ? c1 : c0;
                                *sp++ = swap ? c0 : c1;
                            } else if (h0) {
                                node = c0;
                            } else if (h1) {
                                node = c1;
                            } else {
                                node = *--sp;
                            }
#endif
#if RT_WHILE_WHILE
                        }
                        if (node < 0) {
#else
                        } else {
#endif
                            const int enc = ~node;
                            const int first = enc >> 3, count = (enc & 7) + 1;
                            node = *--sp;
                            for (int s = first; s < first + count; s++) {
                                float t;
                                int prim;
                                if (hit_prim(p, ray, s, t, prim)) {
                                    if (any) {
                                        if (t < limit) {  // raytracer.cpp:237, 245
                                            occluded = true;
                                            tbest = t;
                                            pbest = prim;
                                            node = kSentinel;
                                            break;
                                        }
                                    } else {
                                        closest_update(p, ray, t, prim, tbest, pbest, tsecond);
                                    }
                                }
                            }
                        }
                    }
                }
            }

            // reference visibility: a doubtful hit is replayed on the reference's own tree (device_common.cuh)
            if (p.exact_culling && phase != kIdle && pbest >= 0 && !robust_visible(p, ray, pbest, tbest, any ? FLT_MAX : tsecond)) {
                if (any) {
                    cnt.replay_any++;
                    occluded = ref_any(p, ray, limit);
                } else {
                    cnt.replay_closest++;
                    ref_closest(p, ray, tbest, pbest);
                }
            }

            // ---- 3. consume the result ---------------------------------------------------------------
            bool lights_done = false;
            bool finish = false;
            V3 result = mk(0.0f, 0.0f, 0.0f);

            if (phase == kClosest) {
                if (pbest < 0) {  // raytracer.cpp:442-449
                    result = depth > 0 ? mk(0.0f, 0.0f, 0.0f) : ld3(p.background);
                    finish = true;
                } else {
                    if (pbest < p.n_tris) {
                        const float4 nm = __ldg(&p.tri_nm[pbest]);
                        n = xyz(nm);
                        mat = __float_as_int(nm.w);
                    } else {
                        const float4 cr = __ldg(&p.sph_cr[pbest - p.n_tris]);
                        mat = __ldg(&p.sph_mat[pbest - p.n_tris]);
                        n = normalize((((ray.o + ray.d * tbest) - xyz(cr)) / cr.w));  // raytracer.cpp:91
                    }
                    const float4 m0 = __ldg(&p.materials[4 * (mat - 1)]);
                    color = mk(0.0f, 0.0f, 0.0f) + mulv(xyz(m0), Ia);  // raytracer.cpp:394-395
                    Pt = ray.o + ray.d * tbest;
                    hitprim = pbest;
                    dn = normalize(ray.d);
                    light = 0;
                    lights_done = p.n_lights == 0;
                }
            } else if (phase == kShadow) {
                if (occluded) {
                    cnt.occluded++;
                } else {  // raytracer.cpp:406-423
                    const V3 lpos = xyz(__ldg(&p.lights[2 * light]));
                    const V3 I = xyz(__ldg(&p.lights[2 * light + 1]));
                    const float4 m0 = __ldg(&p.materials[4 * (mat - 1)]);
                    const float4 m1 = __ldg(&p.materials[4 * (mat - 1) + 1]);
                    const V3 wi = ray.d;
                    const float dist = limit;
                    const V3 wiReal = normalize(lpos - Pt);
                    const float cosTheta = dot(wiReal, n);
                    const float d2 = dist * dist;
                    const V3 E = mk(I.x / d2, I.y / d2, I.z / d2);
                    if (specular_gate(cosTheta)) {
                        const float4 m2 = __ldg(&p.materials[4 * (mat - 1) + 2]);
                        const V3 h = normalize(wi + (-dn));
                        const float c = pow_ref(std_max(0.0f, dot(hitprim < p.n_tris ? xyz(__ldg(&p.tri_nn[hitprim])) : normalize(n), h)), m0.w);
                        color = color + mulv(xyz(m2) * c, E);
                    }
                    const float cd = std_max(0.0f, std_min(1.0f, cosTheta));
                    color = color + mulv(xyz(m1) * cd, E);
                }
                light++;
                lights_done = light >= p.n_lights;
            }

            if (phase != kIdle && !finish) {
                const V3 Pe = Pt + n * p.eps;  // raytracer.cpp:397
                if (!lights_done) {            // raytracer.cpp:399-404: shadow ray towards light `light`
                    const V3 lpos = xyz(__ldg(&p.lights[2 * light]));
                    const V3 toL = lpos - Pe;
                    const float dist = length(toL);
                    ray = make_ray(Pe, mk(toL.x / dist, toL.y / dist, toL.z / dist));
                    limit = dist;
                    phase = kShadow;
                    cnt.shadow++;
                } else {
                    const float4 m1 = __ldg(&p.materials[4 * (mat - 1) + 1]);
                    if (__float_as_int(m1.w) != 0) {  // mirror: raytracer.cpp:430-439
                        local_stack[npush] = color;
                        mat_stack[npush] = mat;
                        npush++;
                        const V3 nn = hitprim < p.n_tris ? xyz(__ldg(&p.tri_nn[hitprim])) : normalize(n);
                        const float rc = dot(-dn, nn);
                        depth++;
                        if (depth > p.max_depth) {  // raytracer.cpp:387-389
                            result = mk(0.0f, 0.0f, 0.0f);
                            finish = true;
                        } else {
                            ray = make_ray(Pe, dn + (nn * 2.0f) * rc);
                            limit = FLT_MAX;
                            phase = kClosest;
                            cnt.reflection++;
                        }
                    } else {
                        result = clamp3(color);
                        finish = true;
                    }
                }
            }

            if (finish) {
                while (npush > 0) {  // fold the mirror levels back to front
                    npush--;
                    const V3 km = xyz(__ldg(&p.materials[4 * (mat_stack[npush] - 1) + 3]));
                    result = clamp3(local_stack[npush] + mulv(result, km));
                }
                const unsigned r8 = quantise(result.x), g8 = quantise(result.y), b8 = quantise(result.z);
                if (f == 1) {
                    unsigned char *o;
                    if (p.out_mode == kOutFrame) o = p.out + ((size_t) (py0 + ly) * p.nx + (px0 + lx)) * 3;
                    else o = p.out + (((size_t) local_tile * RT_TILE + (iy0 + ly)) * RT_TILE + (ix0 + lx)) * 3;
                    o[0] = (unsigned char) r8;
                    o[1] = (unsigned char) g8;
                    o[2] = (unsigned char) b8;
                } else {
                    unsigned *a = &acc[((ly / f) * pw + (lx / f)) * 3];
                    atomicAdd(a, r8);
                    atomicAdd(a + 1, g8);
                    atomicAdd(a + 2, b8);
                }
                phase = kIdle;
            }
        }

        if (f > 1) {  // raytracer.cpp:475-477: truncating integer average of the quantised sub-samples
            __syncwarp();
            const unsigned ff = (unsigned) (f * f);
            for (int i = lane; i < pw * ph; i += 32) {
                const int x = i % pw, y = i / pw;
                const unsigned *a = &acc[i * 3];
                unsigned char *o;
                if (p.out_mode == kOutFrame) o = p.out + ((size_t) (py0 + y) * p.nx + (px0 + x)) * 3;
                else o = p.out + (((size_t) local_tile * RT_TILE + (iy0 + y)) * RT_TILE + (ix0 + x)) * 3;
                o[0] = (unsigned char) (a[0] / ff);
                o[1] = (unsigned char) (a[1] / ff);
                o[2] = (unsigned char) (a[2] / ff);
            }
            __syncwarp();
        }
    }

    unsigned v0 = __reduce_add_sync(0xffffffffu, cnt.primary);
    unsigned v1 = __reduce_add_sync(0xffffffffu, cnt.reflection);
    unsigned v2 = __reduce_add_sync(0xffffffffu, cnt.shadow);
    unsigned v3 = __reduce_add_sync(0xffffffffu, cnt.occluded);
    unsigned v4 = __reduce_add_sync(0xffffffffu, cnt.replay_closest);
    unsigned v5 = __reduce_add_sync(0xffffffffu, cnt.replay_any);
    if (lane == 0) {
        atomicAdd(&p.stats[0], (unsigned long long) v0);
        atomicAdd(&p.stats[1], (unsigned long long) v1);
        atomicAdd(&p.stats[2], (unsigned long long) v2);
        atomicAdd(&p.stats[3], (unsigned long long) v3);
        atomicAdd(&p.stats[4], (unsigned long long) v4);
        atomicAdd(&p.stats[5], (unsigned long long) v5);
    }
}

int launch_render_v2(const RenderParams &p, int n_ctas, cudaStream_t stream) {
    const size_t smem = p.f > 1 ? (size_t) kWarps2 * p.P * p.P * 3 * sizeof(unsigned) : 0;
    render_kernel_v2<<<n_ctas, kThreads2, smem, stream>>>(p);
    return (int) cudaGetLastError();
}

int render_kernel_v2_occupancy(int *ctas_per_sm, int *warps_per_cta) {
    *warps_per_cta = kWarps2;
    return (int) cudaOccupancyMaxActiveBlocksPerMultiprocessor(ctas_per_sm, render_kernel_v2, kThreads2,
                                                               (size_t) kWarps2 * kMaxP2 * kMaxP2 * 3 * sizeof(unsigned));
}

}  // namespace rtb
