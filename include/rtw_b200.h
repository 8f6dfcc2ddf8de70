/* rtw_b200.h -- C ABI of the B200-native path tracer (librtw_b200.so).
 *
 * This is the device boundary that the host-side `rtweekend::render(const Scene&, const Config&)`
 * (reference: src/render.h:35,42; body src/render.cpp:135-191) crosses.  Everything below the boundary is
 * hand-written sm_100a CUDA; there is no CPU fallback: every entry point fails with a non-zero status and a
 * message in rtw_last_error() when no CUDA device / kernel image is available.
 *
 * Plain C types only (pointers + sizes + PODs); caller owns every host buffer; the library owns device memory.
 * All calls are blocking unless stated otherwise; the host-buffer entry points are serialised per device, rtw_render_device is
 * re-entrant (see there).
 *
 * Reference interface replaced by each entry point:
 *   rtw_render / rtw_render_multi_gpu  do_work lambda + thread fan-out + sum   render.cpp:150-180
 *                                      (ray_color :112-129, BVHNode::hit :52-71, Camera::get_ray
 *                                      common-model.cpp:156-167, *::hit :64-125, *::scatter :13-62,
 *                                      random-utils.cpp:6-41)
 *   rtw_scene_upload                   Scene::get_root_bvh / BVHNode ctor       render.cpp:73-110,131-133
 *   rtw_primary_hits                   BVHNode::hit on camera rays (parity mode; no reference entry point)
 *   rtw_render_rgb8 / ..._multi_gpu_rgb8  the same + write_color on the device   render.cpp:150-186
 *   rtw_finalize_rgb8[_device]         write_color                             render.cpp:11-20
 *   rtw_scene_hash                     (identity of a Scene for checkpoints; the reference has no resumable render)
 *   rtw_release_cached_buffers         (the reference frees its per-thread images at scope exit, render.cpp:151,176)
 */
#ifndef RTW_B200_H
#define RTW_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RTW_ABI_VERSION 3

#if defined(__GNUC__)
#define RTW_API __attribute__((visibility("default")))
#else
#define RTW_API
#endif

enum rtw_prim_kind { RTW_SPHERE = 0, RTW_MOVING_SPHERE = 1, RTW_TRIANGLE = 2 };
enum rtw_mat_kind { RTW_LAMBERTIAN = 0, RTW_METAL = 1, RTW_DIELECTRIC = 2 };
/* AUTO = BVH (the faster choice at every scene size measured).  SPHERES_SMEM: the brute-force shared-memory sphere sweep, the kernel
 * the FP32-FMA roofline is defined on (sphere-only scenes whose tables fit in shared memory).  BVH picks between its two kernels: the wavefront-per-warp kernel when the
 * scene tables and the per-warp path records fit in shared memory (up to ~1150 spheres) and for every larger sphere-only scene,
 * for scenes with triangles a compressed 8-wide BVH (80-byte nodes, 8-bit child boxes) walked by the per-lane state machine
 * (RTW_BVH_CWIDE).  BVH_PERLANE forces the per-lane state machine on sphere scenes (A/B measurements). */
enum rtw_kernel { RTW_KERNEL_AUTO = 0, RTW_KERNEL_SPHERES_SMEM = 1, RTW_KERNEL_BVH = 2, RTW_KERNEL_BVH_PERLANE = 3 };
enum rtw_bvh_variant { RTW_BVH_NONE = 0, RTW_BVH_PERLANE = 1, RTW_BVH_WAVEFRONT = 2, RTW_BVH_CWIDE = 3 };
/* STATS: count tests / node visits (slower).  SPLIT_ROWS: multi-GPU row-tile split instead of the sample split.
 * NO_SCENE_CACHE: the host-buffer entry points re-flatten, re-build and re-upload the scene even when it is the one the device
 * already holds from the previous call (what the first call of a process pays; bench.py's end-to-end leg uses it).
 * BVH_BUILD_GPU / BVH_BUILD_HOST: where the host-buffer entry points (and rtw_scene_upload_ex) build the BVH: on the device (radix
 * tree, its top and its subtrees rebuilt with SAH: ~15 ms for a million triangles, ~3 % slower to trace) or binned SAH on the host
 * cores (~80 ms).  Neither: the device for scenes of >= 200 000 primitives rendered with < 6e9 paths, the host otherwise. */
enum rtw_flags { RTW_FLAG_STATS = 1, RTW_FLAG_SPLIT_ROWS = 2, RTW_FLAG_NO_SCENE_CACHE = 4, RTW_FLAG_BVH_BUILD_GPU = 16, RTW_FLAG_BVH_BUILD_HOST = 32 };

/* One primitive, in scene insertion order (index in the array == primitive id used for parity).
 * Mirrors the constructor arguments of Sphere / MovingSphere / Triangle (oo-primitives.h:28,49,76). */
typedef struct rtw_primitive {
  int32_t kind;     /* rtw_prim_kind */
  int32_t material; /* index into rtw_scene_desc.mats */
  double a[3];      /* sphere: centre (at time 0); triangle: vertex a */
  double b[3];      /* moving sphere: centre at time 1; triangle: vertex b */
  double c[3];      /* triangle: vertex c */
  double radius;    /* spheres; may be negative (hollow sphere trick) */
} rtw_primitive;

/* Mirrors Lambertian{albedo}, Metal{albedo,fuzz}, Dielectric{ior,fuzz} (common-model.h:123-150). */
typedef struct rtw_material {
  int32_t kind; /* rtw_mat_kind */
  int32_t reserved;
  double albedo[3];
  double fuzz; /* clamped to [0,1] by the library, like the reference constructors */
  double ior;
} rtw_material;

/* The derived camera block computed by Camera's constructor (common-model.cpp:136-154). */
typedef struct rtw_camera {
  double origin[3], lower_left[3], horizontal[3], vertical[3], u[3], v[3];
  double lens_radius, t0, t1;
} rtw_camera;

typedef struct rtw_scene_desc {
  const rtw_primitive* prims;
  int64_t nprims;
  const rtw_material* mats;
  int64_t nmats;
  rtw_camera camera;
} rtw_scene_desc;

typedef struct rtw_render_cfg {
  int32_t width, height;          /* height = int(width / aspect_ratio), render.cpp:137 */
  int32_t sample_begin, sample_end; /* global sample indices [begin,end) rendered by this call (spp shard) */
  int32_t max_child_rays;         /* Config::max_child_rays: up to this many scatters per path */
  int32_t kernel;                 /* rtw_kernel */
  uint64_t seed;                  /* Philox key */
  int32_t device;                 /* CUDA device ordinal */
  int32_t flags;                  /* rtw_flags */
  int32_t rays_per_lane;          /* 0 = default; sphere kernel variant (1, 2 or 4 paths in flight per lane) */
  /* Row-tile split (the alternative to the sample split): the image is cut into tiles of row_tile_rows rows and
   * this call renders tiles row_tile_index, row_tile_index + row_tile_count, ... at the full sample range.  The accumulation
   * buffer then holds only those tiles, packed: rtw_row_tile_local_rows() rows of `width` pixels.  row_tile_count <= 1: whole image. */
  int32_t row_tile_rows, row_tile_count, row_tile_index;
} rtw_render_cfg;

typedef struct rtw_stats {
  uint64_t paths, rays;                       /* always filled by the blocking entry points */
  uint64_t sphere_tests, sphere_candidates;   /* RTW_FLAG_STATS only */
  uint64_t tri_tests, node_visits;            /* RTW_FLAG_STATS only */
  double kernel_ms;                           /* CUDA-event time of the render kernel(s) */
  double h2d_ms, d2h_ms, total_ms;            /* host-buffer entry points */
  double bvh_build_gpu_ms;                    /* > 0: the BVH was built on the device, time of the build kernels (inside h2d_ms) */
  int32_t kernel_used;                        /* rtw_kernel actually launched (SPHERES_SMEM or BVH) */
  int32_t launches;                           /* kernels of this library launched by the call */
  int32_t bvh_variant;                        /* rtw_bvh_variant when kernel_used == RTW_KERNEL_BVH */
  int32_t scene_cache_hit;                    /* host-buffer entry points: 1 = the device(s) already held this scene (no flatten / upload) */
} rtw_stats;

typedef struct rtw_scene rtw_scene; /* device-resident flattened scene (SoA tables, BVH, materials, camera) */

RTW_API int rtw_abi_version(void);
RTW_API const char* rtw_last_error(void);
RTW_API int rtw_device_count(int* count);

/* Flatten + upload (sphere tables, SAH BVH for meshes/mixed scenes, materials, camera) to `device`. */
RTW_API int rtw_scene_upload(const rtw_scene_desc* desc, int32_t device, rtw_scene** out);
/* The same with flags: RTW_FLAG_BVH_BUILD_GPU builds the BVH on the device (default: host SAH). */
RTW_API int rtw_scene_upload_ex(const rtw_scene_desc* desc, int32_t device, int32_t flags, rtw_scene** out);
RTW_API void rtw_scene_free(rtw_scene* scene);
/* Replace the contents of an uploaded scene by a new description (re-flatten, re-build, one H2D copy into the same allocation when
 * it fits).  No render of `scene` may be in flight. */
RTW_API int rtw_scene_update(rtw_scene* scene, const rtw_scene_desc* desc);
/* Number of CUDA kernels this library has launched in this process so far (every <<<>>> is counted where it is issued). */
RTW_API unsigned long long rtw_kernel_launches(void);

/* One-shot render with HOST buffers: upload, render samples [sample_begin,sample_end), download.
 * accum_rgba: width*height*4 floats, (sum r, sum g, sum b, number of samples) per pixel, row 0 = top.
 * The device keeps the flattened scene and its BVH between calls, keyed on a 64-bit hash of the caller's arrays (rtw_scene_hash):
 * a repeated call with an unchanged scene (progressive slices, animation of the sample range) skips flatten, build and upload. */
RTW_API int rtw_render(const rtw_scene_desc* desc, const rtw_render_cfg* cfg, float* accum_rgba, rtw_stats* stats);
/* The same render, finished on the device: rgb8[3 * pixel + c] = int(256 * clamp(sqrt(sum_c / spp), 0, 0.999)) (write_color,
 * render.cpp:11-20, in double from the exact integer sums) with spp = sample_end - sample_begin; 3 bytes per pixel come back
 * instead of 16. */
RTW_API int rtw_render_rgb8(const rtw_scene_desc* desc, const rtw_render_cfg* cfg, uint8_t* rgb8, rtw_stats* stats);

/* Starts creating the CUDA contexts of devices first_device .. first_device + ngpus - 1 on background threads and returns at
 * once; the host-buffer entry points wait for them.  Lets a caller overlap the 0.3-1 s per GPU a cold process pays for its
 * contexts with its own scene construction (and the GPUs with each other).  Optional. */
RTW_API int rtw_prewarm(int32_t first_device, int32_t ngpus);
/* The host-buffer entry points keep per-device state (scene, accumulation buffers, stream) between calls; this frees it. */
RTW_API void rtw_release_cached_buffers(void);
/* 64-bit hash of a scene description (primitives, materials, camera): the key of the scene cache, also stored in checkpoints. */
RTW_API int rtw_scene_hash(const rtw_scene_desc* desc, uint64_t* hash);

/* Device-resident render.  accum_fx: width*height*4 int64 on the scene's device; the kernel ADDS
 * fixed-point radiance (1 unit = 2^-32) per channel and 1 per finished path to channel 3, so shards rendered by
 * different calls / GPUs combine with an exact integer sum (ncclSum on int64) independent of order.
 * Work is enqueued on `cuda_stream` (a cudaStream_t, may be NULL); the call returns without synchronising unless
 * `stats` is non-NULL, in which case it synchronises the stream and fills it.  Every launch gets its own work-queue / statistics
 * block: up to 64 renders of the same rtw_scene may be in flight at once, from any threads and on any streams. */
RTW_API int rtw_render_device(const rtw_scene* scene, const rtw_render_cfg* cfg, int64_t* accum_fx, void* cuda_stream,
                      rtw_stats* stats);
/* accum_fx (int64 x4 per pixel, device) -> accum_rgba (float x4 per pixel, device), on `cuda_stream`. */
RTW_API int rtw_accum_to_float(const int64_t* accum_fx, float* accum_rgba, int64_t npixels, int32_t device,
                       void* cuda_stream);

/* Rows of the packed accumulation buffer of ONE participant of a row-tile split (the same for every participant, so that
 * the buffers can be gathered with equal counts): ceil(ceil(height / tile_rows) / count) * tile_rows. */
RTW_API int32_t rtw_row_tile_local_rows(int32_t height, int32_t tile_rows, int32_t count);
/* gathered: `count` packed buffers back to back ([count][local_rows][width][4] int64, device) -> accum_fx ([height][width][4]). */
RTW_API int rtw_untile_accum(const int64_t* gathered, int64_t* accum_fx, int32_t width, int32_t height, int32_t tile_rows,
                     int32_t count, int32_t device, void* cuda_stream);

/* Single-process multi-GPU render: samples split as evenly as possible over devices 0..ngpus-1 (the first spp % ngpus devices
 * render one sample more; sample indices stay global, so the image does not depend on ngpus).  The int64 buffers are combined
 * over NVLink peer memory: device g sums the g-th slice of every device's buffer with peer loads, converts and downloads that
 * slice itself (needs peer access between the devices).  Bit-identical to the one-GPU result.
 * With cfg->flags & RTW_FLAG_SPLIT_ROWS every GPU renders ALL samples of its interleaved row tiles (cfg->row_tile_rows rows each,
 * default 8) and downloads them straight into their rows of the host image: no exchange between GPUs at all. */
RTW_API int rtw_render_multi_gpu(const rtw_scene_desc* desc, const rtw_render_cfg* cfg, int32_t ngpus, float* accum_rgba,
                         rtw_stats* stats);
RTW_API int rtw_render_multi_gpu_rgb8(const rtw_scene_desc* desc, const rtw_render_cfg* cfg, int32_t ngpus, uint8_t* rgb8,
                              rtw_stats* stats);

/* write_color (render.cpp:11-20) on the device: rgb8 = int(256 * clamp(sqrt(sum / spp), 0, 0.999)).
 * accum_rgba and rgb8 are HOST buffers (npixels*4 floats in, npixels*3 bytes out). */
RTW_API int rtw_finalize_rgb8(const float* accum_rgba, int64_t npixels, int32_t spp, int32_t device, uint8_t* rgb8);
/* The same from a DEVICE-resident int64 accumulation buffer into a DEVICE rgb8 buffer, on `cuda_stream`, without the float rounding
 * of the sums in between: what rtw_render_rgb8 runs. */
RTW_API int rtw_finalize_rgb8_device(const int64_t* accum_fx, int64_t npixels, int32_t spp, int32_t device, void* cuda_stream,
                             uint8_t* rgb8);

/* Deterministic primary-ray mode: aperture 0, shutter [time,time], rays through pixel centres.
 * precision 32: the production fp32 intersection routines (kernel = SPHERES_SMEM or BVH as in rtw_render);
 * precision 64: the reference formulas evaluated in double on the device (brute force over all primitives).
 * Outputs (HOST, per pixel): prim_id (-1 = miss), t, normal[3], front_facing. */
RTW_API int rtw_primary_hits(const rtw_scene_desc* desc, int32_t width, int32_t height, double time, int32_t precision,
                     int32_t kernel, int32_t device, int32_t* prim_id, double* t, double* normal, uint8_t* front);

/* Unit-level hooks used by the parity tests (HOST buffers, n items each).
 * rtw_debug_scatter: the device scatter routines on explicit inputs. ball = the "random_unit_vector" sample,
 *   coin = the Schlick random number.  out_dir/out_att: 3 floats per item; scattered: 0/1 per item.
 * rtw_debug_samples: the device samplers driven by Philox (key=seed, counter=(i,0,dim,0)):
 *   ball[3n] (octant-ball sample), disk[2n] (unit-disk sample), u01[4n] (raw uniforms of dimension 0). */
RTW_API int rtw_debug_scatter(int32_t device, int64_t n, const rtw_material* mats, const float* dir_in, const float* normal,
                      const uint8_t* front, const float* ball, const float* coin, float* out_dir, float* out_att,
                      uint8_t* scattered);
RTW_API int rtw_debug_samples(int32_t device, int64_t n, uint64_t seed, float* ball, float* disk, float* u01);

/* Host-only view of what rtw_scene_upload builds (no CUDA call): table sizes, BVH shape and build time.  Lets CPU-only
 * tests check the flattening of the primitive list (north_star item 1) and the tree invariants. */
typedef struct rtw_flatten_report {
  int32_t n_static_spheres, n_moving_spheres, n_big_spheres, n_triangles;
  int32_t n_bvh_nodes, bvh_max_depth, leaf_direct, reserved;
  int64_t arena_bytes, bvh_errors; /* bvh_errors: primitives missing from / duplicated in the tree (must be 0) */
  double flatten_ms, bvh_build_ms;
  /* the kernel RTW_KERNEL_BVH / AUTO would launch for this scene: rtw_bvh_variant, warps per CTA, whether the scene tables are
   * staged in shared memory, and the dynamic shared memory of the launch */
  int32_t bvh_variant, bvh_warps_per_cta, bvh_tables_in_smem, reserved2;
  int64_t bvh_smem_bytes;
} rtw_flatten_report;
RTW_API int rtw_flatten_info(const rtw_scene_desc* desc, rtw_flatten_report* out);
/* The same structural check on an UPLOADED scene with a binary tree, whichever builder made it (the arena is read back): every
 * primitive referenced exactly once, every stored child box contains the exact bounds of everything below it.  Fills n_bvh_nodes,
 * bvh_max_depth, bvh_errors and bvh_build_ms (time of the device build, 0 for a host-built tree). */
RTW_API int rtw_scene_check(const rtw_scene* scene, rtw_flatten_report* out);

/* FP32 FMA micro-benchmark (scalar FFMA and packed FFMA2 chains, the higher rate): sustained TFLOP/s (2 flop per FMA) and the SM clock seen, the denominator of the
 * sphere-scene roofline (MEASURED_PEAKS.json carries only HBM and bf16 figures). */
RTW_API int rtw_fp32_peak(int32_t device, double seconds, double* tflops, double* sm_mhz);

#ifdef __cplusplus
}
#endif
#endif /* RTW_B200_H */
