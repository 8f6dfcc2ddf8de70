// rtw_internal.h -- shared between the kernels (rtw_kernels.cu) and the C-ABI host code (rtw_abi.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/rtw_b200.h"
#include "rtw_device.cuh"

namespace rtw {

constexpr int kRenderThreads = 256;
constexpr uint32_t kGroupPixels = 128;  // pixels per work unit: a 16 x 8 tile of the image
constexpr double kBigRadius = 100.0;    // |r| >= this: sphere leaves the fp32 tables and is tested in fp64

enum : int { kCtrWork = 0, kCtrRays = 1, kCtrPaths = 2, kCtrSphereTests = 3, kCtrCandidates = 4, kCtrNodes = 5, kCtrTriTests = 6, kCtrCount = 8 };
constexpr int kCtrSlots = 64;  // counter blocks per uploaded scene: one per launch in flight (rtw_host.h)

struct SmemLayout {  // byte offsets into dynamic shared memory and sizes of the staged tables (filled by launch_render)
  uint32_t nodes, refs, sa, sb, tri, records;
  uint32_t b_nodes, b_refs, b_sph, b_tri;
};

struct RenderParams {
  DevScene sc;
  unsigned long long* accum;     // [npix][4] int64 fixed point (2^-32) + finished-path count
  unsigned long long* counters;  // kCtr*
  uint32_t width, height, npix;
  uint32_t tiles_x;              // 16-pixel tile columns: ceil(width / 16)
  uint32_t s_begin, s_end;       // global sample range of this launch
  uint32_t su;                   // samples per unit
  uint32_t n_chunks;             // ceil((s_end - s_begin) / su)
  unsigned long long n_units;    // n_groups * n_chunks
  int32_t max_depth;
  float inv_wm1, inv_hm1;
  uint64_t seed;
  uint32_t n_leaf_refs;
  SmemLayout so;
  // row-tile split: this launch renders tiles tile_index, tile_index + tile_count, ... of tile_rows rows; npix counts the pixels of
  // the packed local buffer (accum); Philox and the camera are keyed on the GLOBAL pixel.  tile_count <= 1: whole image.
  uint32_t tile_rows, tile_count, tile_index;
  uint32_t leaf_min;   // per-lane BVH kernel: lanes that must stand on a leaf before leaves are tested (set by launch_render)
  uint32_t cw_steps, cw_service_min;   // compressed-wide-BVH kernel: steps per traversal phase, lanes that must need service before a service phase
};

struct PrimaryParams {
  DevScene sc;
  uint32_t width, height, npix;
  float time;
  int32_t* prim_id;
  double* t;
  double* normal;
  uint8_t* front;
};

// Which BVH kernel a scene gets (pure host logic, also reported by rtw_flatten_info so that it can be tested without a GPU).
//   table_bytes: 16 + nodes + leaf refs + sphere tables + triangles, the shared-memory footprint of the staged tables.
// Wavefront-per-warp kernel (K2w) while the tables leave room for the per-warp path records: 32 warps per SM (64 registers: the lane
// keeps only what the node visits need, 92 records per warp) up to ~46 KB of tables -- the cover scene fits with 176 bytes to spare --,
// 28 warps (72 registers, 96 records) up to ~58 KB, and for sphere-only scenes 24 warps up to ~87 KB and 20 warps up to ~111 KB (cover
// scene with -n 13..17, 678 to 1159 spheres: +12..29 % over the per-lane kernel; 16 warps at 1296 spheres: -2 %, not instantiated).
// Sphere-only scenes beyond that: the same kernel with the tables read through L1/L2, three 256-thread CTAs per SM (-n 40, 6 402 spheres:
// 4 284 against 3 912 Mpaths/s; -n 120, 57 603 spheres: 3 307 against 3 267).  Meshes beyond the 28-warp tier, multi-primitive leaves, or
// on request: the per-lane state machine (K2), its tables in shared memory up to 72 KB.
// Measured on the cover scene (1080p x 128 spp, one box, scripts/r2_gpu10.sh / r2_gpu12.sh): 20 / 24 / 28 warps 39.46 / 37.97 / 35.48 ms,
// 32 warps with 88 records and batches of 28 / 88, 32 / 92, 30 / 92, 32: 34.91 / 34.96 / 34.84 / 34.68 ms.
constexpr int kWfRecords = 96;                      // path records per warp of K2w
constexpr int kWfRecords32 = 92;                    // ... of its 32-warp tier (>= 32 in flight + 2 x 30: a batch of >= 30 records of one kind always exists)
constexpr size_t kSmemCap = 227u * 1024u;           // dynamic shared memory one CTA may ask for on sm_100
constexpr size_t kPerLaneSmemTables = 72u * 1024u;  // K2 stages its tables in shared memory up to this size (4 CTAs per SM for the cover scene)
__host__ __device__ constexpr uint32_t wf_warp_bytes(int P) { return static_cast<uint32_t>((P * (48 + 8 + 4) + 3 * P + 15) & ~15); }
struct BvhPlan {
  int variant;          // rtw_bvh_variant
  int warps;            // warps per CTA (K2w: one CTA per SM when the tables are in shared memory, three otherwise; K2: 8, four CTAs)
  bool tables_in_smem;
  size_t smem_bytes;    // dynamic shared memory of the launch
};
// (Measured and not kept: the nodes at an 80-byte stride in shared memory, one spare quad per node, so that the same quad of
// different nodes spreads over all eight bank groups instead of two.  Bank conflicts 149 M -> 96 M per 8-spp frame, LSU pipe
// 55 -> 48 %, run time unchanged -- 35.73 against 35.58-35.76 ms at 128 spp -- because the kernel is bound by issue, not by that pipe.)
inline BvhPlan plan_bvh(size_t table_bytes, int n_tri, bool leaf_direct, bool force_perlane, bool cw = false) {
  if (cw) {  // scenes with triangles: compressed wide BVH, per-lane state machine, three CTAs per SM (the 8-child slab test needs registers)
    if (table_bytes <= kPerLaneSmemTables) return {RTW_BVH_CWIDE, 8, true, table_bytes};
    return {RTW_BVH_CWIDE, 8, false, 0};
  }
  const size_t wf_warp = wf_warp_bytes(kWfRecords);
  if (leaf_direct && !force_perlane) {
    if (table_bytes + 32 * wf_warp_bytes(kWfRecords32) <= kSmemCap) return {RTW_BVH_WAVEFRONT, 32, true, table_bytes + 32 * wf_warp_bytes(kWfRecords32)};
    if (table_bytes + 28 * wf_warp <= kSmemCap) return {RTW_BVH_WAVEFRONT, 28, true, table_bytes + 28 * wf_warp};
    if (n_tri == 0) {   // (suzanne, 108 KB of tables, on the 20-warp tier: 17.44 against 16.84 ms for the per-lane kernel reading them through L1)
      if (table_bytes + 24 * wf_warp <= kSmemCap) return {RTW_BVH_WAVEFRONT, 24, true, table_bytes + 24 * wf_warp};
      if (table_bytes + 20 * wf_warp <= kSmemCap) return {RTW_BVH_WAVEFRONT, 20, true, table_bytes + 20 * wf_warp};
      return {RTW_BVH_WAVEFRONT, 8, false, 16 + 8 * wf_warp};
    }
  }
  if (table_bytes <= kPerLaneSmemTables) return {RTW_BVH_PERLANE, 8, true, table_bytes};
  return {RTW_BVH_PERLANE, 8, false, 0};
}

// mode 0: sphere sweep (rays_per_lane 1/2/4); mode 1: BVH (force_perlane: never the wavefront kernel).  *variant <- rtw_bvh_variant launched.
cudaError_t launch_render(const RenderParams& p, int mode, int rays_per_lane, bool force_perlane, bool stats, int sm_count, cudaStream_t stream,
                          int* variant);
cudaError_t launch_primary_f32(const PrimaryParams& p, int mode, cudaStream_t stream);
cudaError_t launch_primary_f64(const rtw_primitive* prims, int nprims, const rtw_camera& cam, uint32_t width, uint32_t height, double time,
                               int32_t* prim_id, double* t, double* normal, uint8_t* front, cudaStream_t stream);
cudaError_t launch_untile(const long long* gathered, long long* full, uint32_t width, uint32_t height, uint32_t tile_rows, uint32_t count,
                          uint32_t local_rows, cudaStream_t stream);
cudaError_t launch_accum_to_float(const long long* fx, float* out, long long npix, cudaStream_t stream);
cudaError_t launch_finalize_rgb8(const float* acc, uint8_t* rgb, long long npix, int spp, cudaStream_t stream);
cudaError_t launch_finalize_rgb8_fx(const long long* fx, uint8_t* rgb, long long npix, int spp, cudaStream_t stream);
cudaError_t launch_debug_scatter(long long n, const int* kind, const float* fuzz, const float* ior, const float* dir_in, const float* normal,
                                 const uint8_t* front, const float* ball, const float* coin, float* out_dir, float* out_att,
                                 const float* albedo, uint8_t* scattered, cudaStream_t stream);
cudaError_t launch_debug_samples(long long n, uint64_t seed, float* ball, float* disk, float* uni, cudaStream_t stream);
cudaError_t launch_ffma_peak(float* out, int blocks, int iters, bool packed, cudaStream_t stream);   // packed: FFMA2, two FMAs per lane and instruction
void count_launch();                 // other translation units report their own kernels
unsigned long long launch_count();

}  // namespace rtw
