// render.cu — the A/B baselines of the render kernel, plus the multi-GPU tile scatter.
//
//   render_kernel     (RT_B200_KERNEL=1)  the first shape of the path: a CTA of 256 threads owns a tile, each thread
//                                         runs whole paths (trace_path), SSAA sums in CTA-shared memory, barriers
//   render_kernel_v3  (RT_B200_KERNEL=3)  the same per-thread code with warp-granular tiles, no barriers
// The product default is render_kernel_v2 (render_v2.cu); these two stay selectable because the ncu comparison that
// motivated it (profiles/r1_kernel_variants_compared.txt) is between the three, and every GPU
// parity test on seeded scenes runs all of them.
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
// Arithmetic contract: this file is compiled with -fmad=false, IEEE division and square root.
// Everything that feeds a hit/miss decision or a colour is written in the reference's operation
// order; the only FMAs are the explicit __fmaf_rn of the (conservative, padded) box test.
#include <cfloat>
#include <cstdint>

#include "device_common.cuh"

namespace rtb {

// Closest hit: argmin over all reported intersections of (t, reference visit rank).
// ANY: true as soon as some primitive reports t < limit (raytracer.cpp:237, 245).
template <bool ANY>
RT_DEV bool traverse(const RenderParams &p, const Ray &r, float limit, float &tbest, int &pbest, float &tsecond) {
    tbest = limit;
    pbest = -1;
    tsecond = FLT_MAX;
    if (p.n_nodes == 0) return false;

    if (p.brute_force) {
        for (int s = 0; s < p.n_prims; s++) {
            float t;
            int prim;
            if (hit_prim(p, r, s, t, prim)) {
                if (ANY) {
                    if (t < limit) {
                        tbest = t;
                        pbest = prim;
                        return true;
                    }
                } else {
                    closest_update(p, r, t, prim, tbest, pbest, tsecond);
                }
            }
        }
        return pbest >= 0;
    }

    int stack[kStackSize];
    int sp = 0;
    stack[sp++] = kSentinel;
    int node = 0;
    while (node != kSentinel) {
        if (node >= 0) {
            const float4 n0 = __ldg(&p.nodes[4 * node]);
            const float4 n1 = __ldg(&p.nodes[4 * node + 1]);
            const float4 n2 = __ldg(&p.nodes[4 * node + 2]);
            const float4 n3 = __ldg(&p.nodes[4 * node + 3]);
            float tmin0, tmax0, tmin1, tmax1;
            slab(r, n0.x, n0.y, n0.z, n0.w, n2.x, n2.y, tmin0, tmax0);
            slab(r, n1.x, n1.y, n1.z, n1.w, n2.z, n2.w, tmin1, tmax1);
            // visit iff the ray overlaps the box for t >= 0 (raytracer.cpp:120) and the entry is not
            // beyond the current limit (raytracer.cpp:188; non-strict so equal-t ties are still seen)
            const bool h0 = tmax0 >= fmaxf(tmin0, 0.0f) && tmin0 <= tbest;
            const bool h1 = tmax1 >= fmaxf(tmin1, 0.0f) && tmin1 <= tbest;
            const int c0 = __float_as_int(n3.x), c1 = __float_as_int(n3.y);
            if (h0 && h1) {
                const bool swap = tmin1 < tmin0;
                node = swap ? c1 : c0;
                stack[sp++] = swap ? c0 : c1;
            } else if (h0) {
                node = c0;
            } else if (h1) {
                node = c1;
            } else {
                node = stack[--sp];
            }
        } else {
            const int enc = ~node;
            const int first = enc >> 3, count = (enc & 7) + 1;
            for (int s = first; s < first + count; s++) {
                float t;
                int prim;
                if (hit_prim(p, r, s, t, prim)) {
                    if (ANY) {
                        if (t < limit) {
                            tbest = t;
                            pbest = prim;
                            return true;
                        }
                    } else {
                        closest_update(p, r, t, prim, tbest, pbest, tsecond);
                    }
                }
            }
            node = stack[--sp];
        }
    }
    return pbest >= 0;
}

// rayTrace (raytracer.cpp:385-452) with the recursion unrolled into a loop: every mirror level
// pushes its local colour and material; the result is folded back to front so that each level's
// `clamp(local + reflected * km)` rounds exactly like the recursive original.
RT_DEV V3 trace_path(const RenderParams &p, V3 o, V3 d, Counters &cnt) {
    V3 local_stack[kMaxSupportedDepth + 1];
    int mat_stack[kMaxSupportedDepth + 1];
    int npush = 0;
    V3 result = mk(0.0f, 0.0f, 0.0f);
    const V3 Ia = ld3(p.ambient);

    for (int depth = 0;; depth++) {
        if (depth > p.max_depth) {  // raytracer.cpp:387-389
            result = mk(0.0f, 0.0f, 0.0f);
            break;
        }
        if (depth == 0) cnt.primary++;
        else cnt.reflection++;
        const Ray ray = make_ray(o, d);
        float t;
        int prim;
        float t2;
        bool hit = traverse<false>(p, ray, FLT_MAX, t, prim, t2);
        if (hit && p.exact_culling && !robust_visible(p, ray, prim, t, t2)) {  // doubtful: replay the reference's traversal
            cnt.replay_closest++;
            ref_closest(p, ray, t, prim);
            hit = prim >= 0;
        }
        if (!hit) {  // raytracer.cpp:442-449
            result = depth > 0 ? mk(0.0f, 0.0f, 0.0f) : ld3(p.background);
            break;
        }
        V3 n;
        int mat;
        if (prim < p.n_tris) {
            const float4 nm = __ldg(&p.tri_nm[prim]);
            n = xyz(nm);
            mat = __float_as_int(nm.w);
        } else {
            const float4 cr = __ldg(&p.sph_cr[prim - p.n_tris]);
            mat = __ldg(&p.sph_mat[prim - p.n_tris]);
            n = normalize((((o + d * t) - xyz(cr)) / cr.w));  // raytracer.cpp:91
        }
        const float4 m0 = __ldg(&p.materials[4 * (mat - 1)]);      // ka, phong
        const float4 m1 = __ldg(&p.materials[4 * (mat - 1) + 1]);  // kd, is_mirror
        const float4 m2 = __ldg(&p.materials[4 * (mat - 1) + 2]);  // ks

        V3 color = mk(0.0f, 0.0f, 0.0f) + mulv(xyz(m0), Ia);  // raytracer.cpp:394-395
        const V3 P = o + d * t;
        const V3 Pe = P + n * p.eps;  // raytracer.cpp:397
        const V3 dn = normalize(d);
        const V3 nn = prim < p.n_tris ? xyz(__ldg(&p.tri_nn[prim])) : normalize(n);

        for (int li = 0; li < p.n_lights; li++) {  // raytracer.cpp:399-427
            const V3 lpos = xyz(__ldg(&p.lights[2 * li]));
            const V3 toL = lpos - Pe;
            const float dist = length(toL);
            const V3 wi = mk(toL.x / dist, toL.y / dist, toL.z / dist);
            cnt.shadow++;
            const Ray sray = make_ray(Pe, wi);
            float ts;
            int ps;
            float ts2;
            bool occluded = traverse<true>(p, sray, dist, ts, ps, ts2);
            if (occluded && p.exact_culling && !robust_visible(p, sray, ps, ts, FLT_MAX)) {
                cnt.replay_any++;
                occluded = ref_any(p, sray, dist);
            }
            if (occluded) {
                cnt.occluded++;
                continue;
            }
            const V3 wiReal = normalize(lpos - P);
            const float cosTheta = dot(wiReal, n);
            const V3 I = xyz(__ldg(&p.lights[2 * li + 1]));
            const float d2 = dist * dist;
            const V3 E = mk(I.x / d2, I.y / d2, I.z / d2);
            if (specular_gate(cosTheta)) {
                const V3 h = normalize(wi + (-dn));
                const float c = pow_ref(std_max(0.0f, dot(nn, h)), m0.w);
                color = color + mulv(xyz(m2) * c, E);
            }
            const float cd = std_max(0.0f, std_min(1.0f, cosTheta));  // clampFloat(cos, 0, 1), raytracer.cpp:21-23
            color = color + mulv(xyz(m1) * cd, E);
        }

        if (__float_as_int(m1.w) != 0) {  // raytracer.cpp:430-439
            local_stack[npush] = color;
            mat_stack[npush] = mat;
            npush++;
            const float rc = dot(-dn, nn);
            o = Pe;
            d = dn + (nn * 2.0f) * rc;
            continue;
        }
        result = mk(clamp_ref(color.x, 0.0f, FLT_MAX), clamp_ref(color.y, 0.0f, FLT_MAX), clamp_ref(color.z, 0.0f, FLT_MAX));
        break;
    }
    while (npush > 0) {
        npush--;
        const V3 km = xyz(__ldg(&p.materials[4 * (mat_stack[npush] - 1) + 3]));
        const V3 c = local_stack[npush] + mulv(result, km);
        result = mk(clamp_ref(c.x, 0.0f, FLT_MAX), clamp_ref(c.y, 0.0f, FLT_MAX), clamp_ref(c.z, 0.0f, FLT_MAX));
    }
    return result;
}

constexpr int kThreads = 256;
constexpr int kMaxP = RT_TILE;  // output pixels per work-item side never exceed one tile

#ifndef RT_MIN_CTAS
#define RT_MIN_CTAS 3
#endif
__global__ void __launch_bounds__(kThreads, RT_MIN_CTAS) render_kernel(const __grid_constant__ RenderParams p) {
    __shared__ unsigned acc[kMaxP * kMaxP * 3];
    __shared__ unsigned s_item;

    const int tid = threadIdx.x;
    const int f = p.f, P = p.P, S = P * f;
    const int Sw = (S + 7) & ~7, Sh = (S + 3) & ~3;
    const int wbx = Sw >> 3;
    const int n_slots = Sw * Sh;
    const int items_per_tile = p.items_x * p.items_x;
    const bool warp_in_one_pixel = (f % 8) == 0;
    Counters cnt = {0u, 0u, 0u, 0u, 0u, 0u};
    const V3 E0 = ld3(p.e), Q = ld3(p.q), U = ld3(p.u), Vv = ld3(p.v);

    for (;;) {
        __syncthreads();  // previous item's pixels are written, acc and s_item are free
        if (tid == 0) s_item = atomicAdd(p.work_counter, 1u);
        if (f > 1)
            for (int i = tid; i < P * P * 3; i += kThreads) acc[i] = 0u;
        __syncthreads();
        const unsigned item = s_item;
        if (item >= p.n_items) break;

        const int local_tile = (int) (item / (unsigned) items_per_tile);
        const int sub = (int) (item % (unsigned) items_per_tile);
        const int tile = p.part_rank + local_tile * p.part_world;
        const int tx0 = (tile % p.tiles_x) * RT_TILE, ty0 = (tile / p.tiles_x) * RT_TILE;
        const int ix0 = (sub % p.items_x) * P, iy0 = (sub / p.items_x) * P;  // within the tile
        const int px0 = tx0 + ix0, py0 = ty0 + iy0;
        const int pw = max(0, min(min(P, RT_TILE - ix0), p.nx - px0));
        const int ph = max(0, min(min(P, RT_TILE - iy0), p.ny - py0));
        const int sw = pw * f, sh = ph * f;

        if (pw > 0 && ph > 0) {
            for (int slot = tid; slot < n_slots; slot += kThreads) {
                const int wb = slot >> 5, lane = slot & 31;
                const int lx = (wb % wbx) * 8 + (lane & 7);
                const int ly = (wb / wbx) * 4 + (lane >> 3);
                const bool valid = lx < sw && ly < sh;
                unsigned r8 = 0, g8 = 0, b8 = 0;
                if (valid) {
                    // raytracer.cpp:319-324 on the (nx*f) x (ny*f) sub-sample grid
                    const float su = ((float) (px0 * f + lx) + 0.5f) * p.su_mul;
                    const float sv = ((float) (py0 * f + ly) + 0.5f) * p.sv_mul;
                    const V3 s = (Q + U * su) - Vv * sv;
                    const V3 c = trace_path(p, E0, s - E0, cnt);
                    r8 = quantise(c.x);
                    g8 = quantise(c.y);
                    b8 = quantise(c.z);
                }
                if (f == 1) {
                    if (valid) {
                        unsigned char *o;
                        if (p.out_mode == kOutFrame) o = p.out + ((size_t) (py0 + ly) * p.nx + (px0 + lx)) * 3;
                        else o = p.out + (((size_t) local_tile * RT_TILE + (iy0 + ly)) * RT_TILE + (ix0 + lx)) * 3;
                        o[0] = (unsigned char) r8;
                        o[1] = (unsigned char) g8;
                        o[2] = (unsigned char) b8;
                    }
                } else if (warp_in_one_pixel) {
                    // the warp's 8x4 block lies inside one output pixel: one shared atomic per channel
                    r8 = __reduce_add_sync(0xffffffffu, r8);
                    g8 = __reduce_add_sync(0xffffffffu, g8);
                    b8 = __reduce_add_sync(0xffffffffu, b8);
                    if (lane == 0 && lx < sw && ly < sh) {
                        unsigned *a = &acc[((ly / f) * P + (lx / f)) * 3];
                        atomicAdd(a, r8);
                        atomicAdd(a + 1, g8);
                        atomicAdd(a + 2, b8);
                    }
                } else if (valid) {
                    unsigned *a = &acc[((ly / f) * P + (lx / f)) * 3];
                    atomicAdd(a, r8);
                    atomicAdd(a + 1, g8);
                    atomicAdd(a + 2, b8);
                }
            }
        }
        if (f > 1) {
            __syncthreads();
            const unsigned ff = (unsigned) (f * f);
            for (int i = tid; i < pw * ph; i += kThreads) {  // raytracer.cpp:475-477, truncating division
                const int x = i % pw, y = i / pw;
                const unsigned *a = &acc[(y * P + x) * 3];
                unsigned char *o;
                if (p.out_mode == kOutFrame) o = p.out + ((size_t) (py0 + y) * p.nx + (px0 + x)) * 3;
                else o = p.out + (((size_t) local_tile * RT_TILE + (iy0 + y)) * RT_TILE + (ix0 + x)) * 3;
                o[0] = (unsigned char) (a[0] / ff);
                o[1] = (unsigned char) (a[1] / ff);
                o[2] = (unsigned char) (a[2] / ff);
            }
        }
    }

    // exact ray counters: warp-reduce, one 64-bit atomic per warp and counter
    unsigned v0 = __reduce_add_sync(0xffffffffu, cnt.primary);
    unsigned v1 = __reduce_add_sync(0xffffffffu, cnt.reflection);
    unsigned v2 = __reduce_add_sync(0xffffffffu, cnt.shadow);
    unsigned v3 = __reduce_add_sync(0xffffffffu, cnt.occluded);
    unsigned v4 = __reduce_add_sync(0xffffffffu, cnt.replay_closest);
    unsigned v5 = __reduce_add_sync(0xffffffffu, cnt.replay_any);
    if ((tid & 31) == 0) {
        atomicAdd(&p.stats[0], (unsigned long long) v0);
        atomicAdd(&p.stats[1], (unsigned long long) v1);
        atomicAdd(&p.stats[2], (unsigned long long) v2);
        atomicAdd(&p.stats[3], (unsigned long long) v3);
        atomicAdd(&p.stats[4], (unsigned long long) v4);
        atomicAdd(&p.stats[5], (unsigned long long) v5);
    }
}

// ---- kernel 3: the same per-lane code (trace_path), but WARP-granular ---------------------------
// One warp claims a warp tile (P x P output pixels) from the global counter and walks its sub-samples
// 32 at a time in 8x4 blocks; SSAA sums live in warp-private shared memory.  No CTA barrier: warps
// never wait for each other, and small frames still produce enough independent work items.
constexpr int kWarps3 = 4;
constexpr int kThreads3 = kWarps3 * 32;
constexpr int kMaxP3 = 16;
#ifndef RT_MIN_CTAS3
#define RT_MIN_CTAS3 6
#endif

__global__ void __launch_bounds__(kThreads3, RT_MIN_CTAS3) render_kernel_v3(const __grid_constant__ RenderParams p) {
    __shared__ unsigned acc_all[kWarps3][kMaxP3 * kMaxP3 * 3];
    const int lane = threadIdx.x & 31;
    unsigned *acc = acc_all[threadIdx.x >> 5];
    const int f = p.f, P = p.P;
    const int items_per_tile = p.items_x * p.items_y;
    const bool warp_in_one_pixel = (f % 8) == 0;
    Counters cnt = {0u, 0u, 0u, 0u, 0u, 0u};
    const V3 E0 = ld3(p.e), Q = ld3(p.q), U = ld3(p.u), Vv = ld3(p.v);

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
        const int total = nbx * nby * 32;
        if (f > 1) {
            for (int i = lane; i < pw * ph * 3; i += 32) acc[i] = 0u;
            __syncwarp();
        }
        for (int slot = lane; slot < total; slot += 32) {
            const int b = slot >> 5;
            const int lx = (b % nbx) * 8 + (lane & 7);
            const int ly = (b / nbx) * 4 + (lane >> 3);
            const bool valid = lx < sw && ly < sh;
            unsigned r8 = 0, g8 = 0, b8 = 0;
            if (valid) {
                const float su = ((float) (px0 * f + lx) + 0.5f) * p.su_mul;
                const float sv = ((float) (py0 * f + ly) + 0.5f) * p.sv_mul;
                const V3 s = (Q + U * su) - Vv * sv;
                const V3 c = trace_path(p, E0, s - E0, cnt);
                r8 = quantise(c.x);
                g8 = quantise(c.y);
                b8 = quantise(c.z);
            }
            if (f == 1) {
                if (valid) {
                    unsigned char *o;
                    if (p.out_mode == kOutFrame) o = p.out + ((size_t) (py0 + ly) * p.nx + (px0 + lx)) * 3;
                    else o = p.out + (((size_t) local_tile * RT_TILE + (iy0 + ly)) * RT_TILE + (ix0 + lx)) * 3;
                    o[0] = (unsigned char) r8;
                    o[1] = (unsigned char) g8;
                    o[2] = (unsigned char) b8;
                }
            } else if (warp_in_one_pixel) {
                r8 = __reduce_add_sync(0xffffffffu, r8);
                g8 = __reduce_add_sync(0xffffffffu, g8);
                b8 = __reduce_add_sync(0xffffffffu, b8);
                if (lane == 0 && valid) {
                    unsigned *a = &acc[((ly / f) * pw + (lx / f)) * 3];
                    a[0] += r8;
                    a[1] += g8;
                    a[2] += b8;
                }
            } else if (valid) {
                unsigned *a = &acc[((ly / f) * pw + (lx / f)) * 3];
                atomicAdd(a, r8);
                atomicAdd(a + 1, g8);
                atomicAdd(a + 2, b8);
            }
        }
        if (f > 1) {
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

int launch_render_v3(const RenderParams &p, int n_ctas, cudaStream_t stream) {
    render_kernel_v3<<<n_ctas, kThreads3, 0, stream>>>(p);
    return (int) cudaGetLastError();
}

int render_kernel_v3_occupancy(int *ctas_per_sm, int *warps_per_cta) {
    *warps_per_cta = kWarps3;
    return (int) cudaOccupancyMaxActiveBlocksPerMultiprocessor(ctas_per_sm, render_kernel_v3, kThreads3, 0);
}

// Gathering GPU: scatter `part_world` packed tile buffers into the row-major frame.
__global__ void assemble_tiles_kernel(const unsigned char *parts, long long part_stride, int part_world, int nx, int ny,
                                      int tiles_x, int n_tiles, unsigned char *frame) {
    // one thread per output byte-triple
    const long long idx = (long long) blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long) nx * ny;
    if (idx >= total) return;
    const int x = (int) (idx % nx), y = (int) (idx / nx);
    const int tile = (y / RT_TILE) * tiles_x + (x / RT_TILE);
    const int part = tile % part_world, local_tile = tile / part_world;
    const unsigned char *src = parts + (long long) part * part_stride +
                               (((long long) local_tile * RT_TILE + (y % RT_TILE)) * RT_TILE + (x % RT_TILE)) * 3;
    unsigned char *dst = frame + idx * 3;
    dst[0] = src[0];
    dst[1] = src[1];
    dst[2] = src[2];
}

// ---- launch wrappers (called from api.cu) -----------------------------------------------------

int launch_render(const RenderParams &p, int n_ctas, cudaStream_t stream) {
    render_kernel<<<n_ctas, kThreads, 0, stream>>>(p);
    return (int) cudaGetLastError();
}

int launch_assemble(const unsigned char *parts, long long part_stride, int part_world, int nx, int ny, int tiles_x,
                    int n_tiles, unsigned char *frame, cudaStream_t stream) {
    const long long total = (long long) nx * ny;
    const int threads = 256;
    const long long blocks = (total + threads - 1) / threads;
    assemble_tiles_kernel<<<(unsigned) blocks, threads, 0, stream>>>(parts, part_stride, part_world, nx, ny, tiles_x, n_tiles, frame);
    return (int) cudaGetLastError();
}

int render_kernel_occupancy(int *ctas_per_sm) {
    return (int) cudaOccupancyMaxActiveBlocksPerMultiprocessor(ctas_per_sm, render_kernel, kThreads, 0);
}

}  // namespace rtb
