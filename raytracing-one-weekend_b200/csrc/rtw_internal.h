// rtw_internal.h -- shared between the kernels (rtw_kernels.cu) and the C-ABI host code (rtw_abi.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/rtw_b200.h"
#include "rtw_device.cuh"

namespace rtw {

constexpr int kRenderThreads = 256;
constexpr uint32_t kGroupPixels = 128;  // pixels per work unit (consecutive in row-major order)
constexpr double kBigRadius = 100.0;    // |r| >= this: sphere leaves the fp32 tables and is tested in fp64

enum : int { kCtrWork = 0, kCtrRays = 1, kCtrPaths = 2, kCtrSphereTests = 3, kCtrCandidates = 4, kCtrNodes = 5, kCtrTriTests = 6, kCtrCount = 8 };

struct SmemLayout {  // byte offsets into dynamic shared memory and sizes of the staged tables (filled by launch_render)
  uint32_t nodes, refs, sa, sb, tri, records;
  uint32_t b_nodes, b_refs, b_sph, b_tri;
};

struct RenderParams {
  DevScene sc;
  unsigned long long* accum;     // [npix][4] int64 fixed point (2^-32) + finished-path count
  unsigned long long* counters;  // kCtr*
  uint32_t width, height, npix;
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
cudaError_t launch_debug_scatter(long long n, const int* kind, const float* fuzz, const float* ior, const float* dir_in, const float* normal,
                                 const uint8_t* front, const float* ball, const float* coin, float* out_dir, float* out_att,
                                 const float* albedo, uint8_t* scattered, cudaStream_t stream);
cudaError_t launch_debug_samples(long long n, uint64_t seed, float* ball, float* disk, float* uni, cudaStream_t stream);
cudaError_t launch_ffma_peak(float* out, int blocks, int iters, cudaStream_t stream);

}  // namespace rtw
