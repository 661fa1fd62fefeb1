/*
 * rt_b200.h — C-ABI of the B200-native Whitted hot path.
 *
 * This is the drop-in boundary for the per-pixel path of the CENG477 ray tracer
 * (reference: raytracer.cpp:335 `RayTracer::RayTracer(parser::Scene&)` and
 * raytracer.cpp:362 `Image RayTracer::render(Camera&)`; the post step
 * raytracer.cpp:459 `ImageProcessor::downSample` is fused into rt_render).
 * Plain pointers and sizes only; nothing throws across this boundary.
 *
 * All arrays are caller-owned, read-only and may be freed as soon as the call
 * returns.  Ids are 1-based exactly as in the reference's scene files
 * (parser.h:194-204).  There is no CPU fallback: every entry point that needs a
 * GPU returns RT_ERR_CUDA when none is usable.
 */
#ifndef RT_B200_H
#define RT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RT_B200_ABI_VERSION 3

/* error codes (0 = success, negative = failure; text via rt_last_error()) */
#define RT_OK 0
#define RT_ERR_INVALID (-1) /* null pointer, bad id, bad size */
#define RT_ERR_CUDA (-2)    /* CUDA runtime error or no device */
#define RT_ERR_NOMEM (-3)
#define RT_ERR_STATE (-4) /* handle used on the wrong device / busy */

/* parser.h:18-20 Vec3f */
typedef struct RtVec3 {
  float x, y, z;
} RtVec3;

/* parser.h:185-192 Material (is_mirror <- `type="mirror"`, parser.cpp:119) */
typedef struct RtMaterial {
  RtVec3 ambient;
  RtVec3 diffuse;
  RtVec3 specular;
  RtVec3 mirror;
  float phong_exponent;
  int32_t is_mirror;
} RtMaterial;

/* parser.h:180-183 PointLight */
typedef struct RtPointLight {
  RtVec3 position;
  RtVec3 intensity;
} RtPointLight;

/* parser.h:243-251 Triangle: the flat list the reference builds at
 * raytracer.cpp:336-341 — every <Triangle> first, then each mesh's faces in
 * file order, the mesh's material copied onto each face. */
typedef struct RtTriangle {
  int32_t v0_id, v1_id, v2_id; /* 1-based into vertices */
  int32_t material_id;         /* 1-based into materials */
} RtTriangle;

/* parser.h:200-204 Sphere */
typedef struct RtSphere {
  int32_t material_id;
  int32_t center_vertex_id;
  float radius;
} RtSphere;

/* parser.h:254-266 Scene, flattened */
typedef struct RtSceneDesc {
  const RtVec3 *vertices;
  int32_t n_vertices;
  const RtTriangle *triangles;
  int32_t n_triangles;
  const RtSphere *spheres;
  int32_t n_spheres;
  const RtMaterial *materials;
  int32_t n_materials;
  const RtPointLight *lights;
  int32_t n_lights;
  RtVec3 ambient_light;
  int32_t background[3]; /* parser.cpp:33 — parsed as ints */
  float shadow_ray_epsilon;
  int32_t max_recursion_depth;
} RtSceneDesc;

/* parser.h:170-178 Camera (near_plane = l r b t); image_width/height are the
 * OUTPUT resolution — the supersampling factor is an argument of rt_render,
 * not pre-multiplied into the camera as raytracer.cpp:506-509 does. */
typedef struct RtCamera {
  RtVec3 position;
  RtVec3 gaze;
  RtVec3 up;
  float l, r, b, t;
  float near_distance;
  int32_t image_width, image_height;
} RtCamera;

/* acceleration-structure builders (rt_scene_create) */
#define RT_BUILD_DEFAULT 0 /* = RT_BUILD_AUTO */
#define RT_BUILD_LBVH_GPU 1 /* Morton + radix sort + Karras on the GPU */
#define RT_BUILD_SAH_HOST 2 /* binned SAH on the host (quality yardstick) */
#define RT_BUILD_PLOC_GPU 3 /* Morton sort + PLOC agglomerative clustering + SAH leaf collapse on the GPU */
#define RT_BUILD_AUTO 4     /* PLOC and top-down binned SAH, both on the GPU; keeps the PLOC tree when its SAH cost is
                               < 0.8x the other's (RtSceneInfo.builder tells which one was kept) */
#define RT_BUILD_SAH_GPU 5  /* top-down binned SAH on the GPU (same algorithm as RT_BUILD_SAH_HOST) */

typedef struct RtBuildOptions {
  int32_t builder;    /* RT_BUILD_* */
  int32_t brute_force; /* 1: ignore the BVH and test every primitive (parity debugging) */
  int32_t no_exact_culling; /* 1: skip the reference-visibility check (a hit the reference's own box test would
                               have culled is then reported; a few rays per 10^8 on the shipped scenes) */
  int32_t refill_threshold; /* experiments: refill idle lanes of a warp once at most this many are busy (0 = default:
                               only when the warp has drained; needs an AA factor that is not a multiple of 8) */
  int32_t ploc_radius;      /* experiments: PLOC neighbour search radius (0 = default 16) */
  float ploc_leaf_cost;     /* experiments: per-primitive cost in PLOC's leaf-collapse decision (0 = default 1.0) */
  int32_t force_replay;     /* tests: every ray that reports a hit is re-decided by the exact replay of the reference's
                               traversal (slow; must give the same frame as the default path) */
  int32_t max_ctas_per_sm;  /* experiments: cap on resident CTAs per SM for the render kernel (0 = occupancy limit) */
  int32_t reinsert_rounds;  /* rounds of insertion-based optimisation of the top-down SAH tree (subtrees move to where the
                               tree's total box area grows least): 0 = default (8), < 0 = off */
  float reinsert_accept;    /* the optimised tree is kept when its SAH cost < this x the cost as built (0 = default 0.8;
                               a large value keeps it always) */
} RtBuildOptions;

/* counters are exact (device atomics); "ray" = one closest-hit query
 * (raytracer.cpp:177) or one any-hit query (raytracer.cpp:227) */
typedef struct RtStats {
  uint64_t primary_rays;
  uint64_t reflection_rays; /* reflection rays actually traced */
  uint64_t shadow_rays;
  uint64_t shadow_occluded;
  uint64_t replayed_closest; /* closest-hit rays recomputed by the exact replay of the reference's traversal */
  uint64_t replayed_any;     /* shadow rays recomputed likewise */
  float ms_render; /* device time of the render kernel(s), CUDA events */
  float ms_d2h;    /* device-to-host copy of the RGB8 frame */
  float ms_total;  /* camera known -> RGB8 on host */
  int32_t n_launches; /* kernels launched by this call */
  int32_t reserved[3];
} RtStats;

typedef struct RtSceneInfo {
  int32_t n_triangles, n_spheres;
  int32_t bvh_nodes;      /* nodes of the traversal BVH */
  int32_t bvh_max_depth;
  int32_t ref_tree_nodes; /* nodes of the reference-order tree (bvh.h:48-105) used for tie ranks */
  int32_t ref_tree_leaves;
  int32_t ref_tree_max_leaf;
  int32_t ref_tree_max_depth;
  float ms_build_host;   /* host time until the whole build was enqueued (validation, uploads, launches) */
  float ms_build_device; /* uploads + every build kernel on the device (CUDA events) */
  float bvh_sah_cost;
  int32_t builder;       /* the builder whose tree was kept (RT_BUILD_AUTO reports PLOC or SAH_GPU) */
  int32_t device;
  float sah_cost_ploc, sah_cost_sah; /* RT_BUILD_AUTO: the two candidates' SAH costs (computed on the device) */
  float ms_create_wall;  /* wall time of the whole rt_scene_create call */
  float reinsert_cost_before, reinsert_cost_after; /* SAH cost of the top-down tree as built / after the optimisation rounds */
  int32_t reinsert_moves, reinsert_rounds;         /* subtrees moved, rounds run (stops early when a round moves nothing) */
  int32_t reinsert_accepted;                       /* 1: the optimised tree replaced the builder's */
} RtSceneInfo;

typedef struct RtScene RtScene; /* opaque */

/* Replaces RayTracer::RayTracer (raytracer.cpp:335-350): stages the scene as
 * SoA buffers in HBM on the CURRENT CUDA device, builds the BVH there and the
 * reference-order tie ranks.  opts may be NULL. */
int rt_scene_create(const RtSceneDesc *desc, const RtBuildOptions *opts, RtScene **out);
void rt_scene_destroy(RtScene *scene);
/* The library keeps a few device blocks of destroyed scenes and finished builds per GPU and reuses them, so that creating
 * scene after scene makes no cudaMalloc / cudaFree call; rt_trim() gives them back to the driver. */
int rt_trim(void);
int rt_scene_info(const RtScene *scene, RtSceneInfo *info);

/* Replaces RayTracer::render + ImageProcessor::downSample
 * (raytracer.cpp:362-383, 459-484): renders `cam` at aa_factor x aa_factor
 * regular-grid supersampling, each sub-sample quantised to 8 bits and the
 * f*f samples averaged with truncating integer division, into caller-allocated
 * HOST memory rgb_out[image_height][image_width][3] (top row first).
 * Synchronous; one call at a time per handle.  stats may be NULL. */
int rt_render(RtScene *scene, const RtCamera *cam, int aa_factor, unsigned char *rgb_out,
              RtStats *stats);

/* Asynchronous variant for rendering several cameras back to back on one resident scene
 * (raytracer.cpp:505-519): rt_render_async enqueues the frame and returns a ticket; rt_wait blocks until that
 * frame is in rgb_out.  Up to 3 frames may be in flight per handle; the device-to-host copy of frame i runs on
 * a second stream while the kernel of frame i+1 executes.  rgb_out must stay valid until rt_wait; page-locked
 * memory (rt_host_alloc) avoids a staging copy. */
int rt_render_async(RtScene *scene, const RtCamera *cam, int aa_factor, unsigned char *rgb_out, int *ticket);
int rt_wait(RtScene *scene, int ticket, RtStats *stats);

/* When rgb_out is page-locked (rt_host_alloc, cudaHostAlloc, cudaHostRegister) and the frame is large (>= 32 MB by
 * default), rt_render / rt_render_async let the kernel store finished pixels straight into it over PCIe instead of rendering
 * into device memory and copying afterwards.  rt_set_zero_copy(min_frame_bytes) moves that threshold for the whole
 * process (negative = never). */
int rt_set_zero_copy(int64_t min_frame_bytes);

/* Page-locked host memory for frames (cudaMallocHost / cudaFreeHost). */
int rt_host_alloc(int64_t bytes, void **ptr);
int rt_host_free(void *ptr);

/* Creates the CUDA context of `device` and loads the render kernels; callable from a helper thread so that both
 * overlap the caller's scene parsing. */
int rt_warmup(int device);

/* ---- multi-GPU: interleaved row bands, scene replicated per GPU -----------
 * The output image is cut into bands of rt_band_height(cam, aa, world) pixel rows; band b belongs to part
 * (b % part_world).  This mirrors the reference's interleaved rows (raytracer.cpp:353: row i goes to thread
 * i % cores) for load balance; at the 16x16 headline configuration a band is ONE pixel row.  The band height is a
 * pure function of (camera, aa_factor, part_world), so every rank and the gathering side agree without talking. */
int rt_band_height(const RtCamera *cam, int aa_factor, int part_world);
/* Tuning / experiments: overrides the band height (pixel rows) and the strip width of the 16x16-type kernel for the whole
 * process (0 = built-in defaults).  Every rank of a multi-process run must set the same values. */
int rt_set_partition(int band_rows, int strip_width);
/* Tuning / experiments: width in pixels (a power of two, 4..512) of the 1024-pixel blocks in which the 16x16-type kernel
 * numbers a part's pixels (0 = built-in: 32). */
int rt_set_block_width(int pixels);

/* pixel rows part `part_rank` owns (its last band padded to the full band height), and bytes of its packed buffer
 * (rows * image_width * 3). */
int64_t rt_part_rows(const RtCamera *cam, int aa_factor, int part_rank, int part_world);
int64_t rt_part_bytes(const RtCamera *cam, int aa_factor, int part_rank, int part_world);

/* Renders only this part's bands into DEVICE memory d_rows (packed: the part's bands back to back,
 * [local_band][band_h][image_width][3], rt_part_bytes long) on `cuda_stream` (a cudaStream_t, may be NULL for the
 * default stream).  Asynchronous w.r.t. the host unless stats != NULL. */
int rt_render_part(RtScene *scene, const RtCamera *cam, int aa_factor, int part_rank,
                   int part_world, void *d_rows, void *cuda_stream, RtStats *stats);

/* As rt_render_part, but each finished pixel is stored straight into a
 * row-major RGB8 frame at d_frame (image_height*image_width*3 bytes), which
 * may be peer memory of another GPU mapped into this process (NVLink P2P):
 * the gather is fused into the render kernel's epilogue. */
int rt_render_part_into_frame(RtScene *scene, const RtCamera *cam, int aa_factor, int part_rank,
                              int part_world, void *d_frame, void *cuda_stream, RtStats *stats);

/* On the gathering GPU: scatters `part_world` packed band buffers laid out back
 * to back with stride `part_stride_bytes` (>= the largest rt_part_bytes) into a
 * row-major RGB8 frame d_frame. */
int rt_assemble_parts(const RtCamera *cam, int aa_factor, int part_world, const void *d_parts,
                      int64_t part_stride_bytes, void *d_frame, void *cuda_stream);

/* End to end without a gather: renders this part's bands and copies them device-to-host over THIS GPU's own PCIe
 * link straight into their rows of `host_frame` (row-major RGB8, image_height*image_width*3 bytes; one strided
 * copy).  Synchronous.  With one process per GPU, host_frame is a frame shared by all ranks (rt_host_frame_*);
 * when every rank has returned, the frame is complete. */
int rt_render_part_to_host(RtScene *scene, const RtCamera *cam, int aa_factor, int part_rank, int part_world,
                           unsigned char *host_frame, RtStats *stats);

/* A host frame shared between the processes of one node: POSIX shared memory `name` ("/something"), mapped and
 * page-locked in the calling process.  One rank creates, the others open; close unmaps (and unlinks when
 * unlink_name != NULL). */
int rt_host_frame_create(const char *name, int64_t bytes, void **ptr);
int rt_host_frame_open(const char *name, int64_t bytes, void **ptr);
int rt_host_frame_close(void *ptr, int64_t bytes, const char *unlink_name);

/* Single-process multi-GPU (the `raytracer --gpus N` path): one handle per device in scenes[0..n_gpus); every GPU
 * renders its bands and copies them over its own PCIe link into rgb_out (directly when rgb_out is page-locked). */
int rt_render_multi(RtScene *const *scenes, int n_gpus, const RtCamera *cam, int aa_factor,
                    unsigned char *rgb_out, RtStats *stats);

/* ---- peer memory for the fused gather (rt_render_part_into_frame) ------------
 * One process per GPU: the gathering rank allocates the frame with
 * rt_device_alloc and exports it; every other rank maps it with rt_ipc_open
 * (CUDA IPC, peer access over NVLink) and passes the mapped pointer as d_frame,
 * so its render kernel stores finished pixels straight into GPU 0's memory. */
#define RT_IPC_HANDLE_BYTES 64
int rt_device_alloc(int64_t bytes, void **d_ptr);
int rt_device_free(void *d_ptr);
int rt_ipc_export(void *d_ptr, unsigned char handle[RT_IPC_HANDLE_BYTES]);
int rt_ipc_open(const unsigned char handle[RT_IPC_HANDLE_BYTES], void **d_ptr);
int rt_ipc_close(void *d_ptr);

/* ---- host-only hooks: no CUDA call inside, usable on a machine without a GPU.
 * They expose the host-side logic of rt_scene_create to CPU tests. */

/* The reference-order tie ranks (bvh.h:48-163 rebuilt, raytracer.cpp:190-196
 * visit order): ranks_out[8][n_triangles + n_spheres]; stats4 = nodes, leaves,
 * max leaf size, max depth of the reference's tree. Either may be NULL. */
int rt_host_reference_ranks(const RtSceneDesc *desc, uint32_t *ranks_out, int32_t *stats4);
/* FNV-1a hash of the rebuilt reference tree (host build). */
int rt_host_reference_tree_hash(const RtSceneDesc *desc, uint64_t *hash);
/* The same ranks, statistics and tree hash from the GPU build that rt_scene_create uses
 * (needs a device; tests compare it with the host build bit for bit). */
int rt_device_reference_ranks(const RtSceneDesc *desc, uint32_t *ranks_out, int32_t *stats4, uint64_t *tree_hash);
/* Builds the host SAH BVH, pads it and checks its invariants (every primitive
 * in exactly one leaf, every box contains what is below it). Returns the node
 * count (>= 0) or a negative error. */
int rt_host_check_bvh(const RtSceneDesc *desc, float *sah_cost, int32_t *max_depth);
/* The plain top-down host tree put through `rounds` rounds of the insertion-based optimisation (the code the GPU build
 * runs, on the host), then the same invariant check.  cost2 = SAH cost before / after; stats4 = subtrees moved, rounds
 * run, depth of the resulting tree, 1 if the optimised tree was kept (cost after < accept_ratio x cost before). */
int rt_host_reinsert(const RtSceneDesc *desc, int rounds, float accept_ratio, float *cost2, int32_t *stats4);
/* The closed forms the kernels use in place of libm's pow and acos
 * (raytracer.cpp:411-414), compiled for the host. */
float rt_host_pow_ref(float base, float exponent);
int rt_host_specular_gate(float cos_theta);

/* Device self check (needs a GPU): the kernels' shared-reciprocal division against the IEEE `/` operator on n
 * pseudo-random operand quadruples, bit for bit; reports the number of mismatches (must be 0) and how many
 * quadruples took the fast path. */
int rt_selftest_div3(uint64_t n, uint32_t seed, uint64_t *mismatches, uint64_t *fast_path);

const char *rt_last_error(void);
int rt_abi_version(void);
int rt_device_count(void);
/* cudaSetDevice for callers that do not link the CUDA runtime themselves (rt_scene_create builds on
 * the CURRENT device). */
int rt_set_device(int device);

#ifdef __cplusplus
}
#endif
#endif /* RT_B200_H */
