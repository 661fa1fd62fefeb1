// render_params.h — the argument block of the render kernels (passed by value, __grid_constant__).
#pragma once

#include <cstdint>

#include <cuda_runtime.h>

namespace rtb {

enum OutMode : int {
    kOutFrame = 0,   // row-major RGB8 frame [ny][nx][3] (possibly peer memory)
    kOutPacked = 1,  // this part's row bands back to back, [local_band][band_h][nx][3]
};

struct RenderParams {
    // scene (device pointers; layouts in rt_internal.h)
    const float4 *nodes;
    const float4 *prims;
    const float4 *tri_nm;
    const float4 *tri_nn;  // normalize(normal) per triangle (raytracer.cpp:414, 432 re-normalise the stored normal)
    const float4 *sph_cr;
    const int *sph_mat;
    const uint32_t *ranks;
    const float4 *ref_nodes;        // reference tree (exact culling replay)
    const int *ref_leaf_prims;
    const float4 *prim_bounds;      // [2 * n_prims] un-padded bounds of each primitive as the reference computes them
    const int *slot_of_prim;        // prim id -> slot in `prims`
    int exact_culling;              // 1: reproduce the reference's box-culling decisions (default); 2: replay every hit (tests)
    const float4 *materials;
    const float4 *lights;
    int n_nodes, n_tris, n_prims, n_lights;
    int max_depth;       // Scene::max_recursion_depth
    int brute_force;
    float eps;           // Scene::shadow_ray_epsilon
    float ambient[3];    // Scene::ambient_light
    float background[3]; // (float) Scene::background_color
    // camera, precomputed on the host exactly as EyeRayGenerator::init does (raytracer.cpp:292-314)
    float e[3], q[3], u[3], v[3];
    float su_mul, sv_mul;
    // frame and work decomposition (api.cu item_geometry()): the frame is cut into row bands of band_h = Ph pixel
    // rows, band b belongs to part b % part_world (the reference deals rows to its threads the same way,
    // raytracer.cpp:353); a band is cut into items of P x Ph pixels; items are numbered so that consecutive ones
    // cover a compact ~32 x 32 pixel block (group of `group_bands` bands x column of `tile_items` items)
    int nx, ny;          // output resolution
    int f;               // supersampling factor (sub-sample grid is nx*f by ny*f)
    int P, Ph;           // item width / height in pixels
    int items_x;         // items per band = ceil(nx / P)
    int n_bands;         // item rows of this part (bands x rows_per_band; rows beyond the image are skipped)
    int rows_per_band;   // item rows (of Ph pixel rows each) per band
    int group_bands, tile_items, tiles_per_group;
    int part_rank, part_world;
    unsigned int n_items;  // work items of this part (incl. empty padding items); register-accumulator mode: pixel slots
    int blk_w_log2;        // register-accumulator mode: pixel slots are numbered in blocks of 2^blk_w_log2 x 2^(10 - blk_w_log2) pixels
    unsigned int guide;    // register-accumulator mode: a warp claims ~ remaining / guide pixels at a time (4 x warps in flight)
    int out_mode;
    int refill_threshold;  // shared-accumulator mode: refill idle lanes once <= this many lanes are busy
    int acc_mode;          // 0: SSAA sums in warp-private shared memory (any f); 1: in registers (f % 8 == 0), see render_v2.cu
    int far_camera;        // 1: widen the box test by the ray origin's rounding error (camera far outside the scene)
    unsigned char *out;
    unsigned long long *control;     // [8]: [0] work counter (zeroed before launch), [1..6] primary, reflection, shadow,
                                     // occluded, replayed closest, replayed any
};

}  // namespace rtb
