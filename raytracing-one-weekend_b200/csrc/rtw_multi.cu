// rtw_multi.cu -- single-process multi-GPU driver behind rtw_render_multi_gpu (include/rtw_b200.h).
// Replaces the reference's thread fan-out + image sum (render.cpp:169-180): samples-per-pixel are
// split over the GPUs (global sample index keys the Philox stream, so the union of samples does not depend on the
// split), each GPU renders into its own int64 fixed-point accumulation buffer, and ONE ncclReduce(sum, int64)
// over NVLink combines them on device 0.  Integer sums are exact, so the result is bit-identical to one GPU.
// NCCL is loaded with dlopen at first use so that processes which already carry their own NCCL (PyTorch) never
// see a second copy just because they loaded this library.
#include <dlfcn.h>
#include <unistd.h>

#include <algorithm>
#include <cstdio>

#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "rtw_internal.h"

namespace {

typedef struct ncclComm* ncclComm_t;
enum { kNcclInt64 = 4, kNcclSum = 0 };

struct NcclApi {
  void* lib = nullptr;
  int (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  int (*Reduce)(const void*, void*, size_t, int, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  std::string error;
  bool load() {
    if (lib) return true;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) { lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (lib) break; }
    if (!lib) { error = std::string("dlopen(libnccl.so.2) failed: ") + dlerror(); return false; }
#define RTW_SYM(field, name)                                           \
  field = reinterpret_cast<decltype(field)>(dlsym(lib, name));         \
  if (!field) { error = std::string("NCCL symbol missing: ") + name; return false; }
    RTW_SYM(CommInitAll, "ncclCommInitAll");
    RTW_SYM(CommDestroy, "ncclCommDestroy");
    RTW_SYM(GroupStart, "ncclGroupStart");
    RTW_SYM(GroupEnd, "ncclGroupEnd");
    RTW_SYM(Reduce, "ncclReduce");
    RTW_SYM(Send, "ncclSend");
    RTW_SYM(Recv, "ncclRecv");
    RTW_SYM(GetErrorString, "ncclGetErrorString");
#undef RTW_SYM
    return true;
  }
};
NcclApi g_nccl;

}  // namespace

extern "C" int rtw_set_error_(const char* msg);  // defined below (thread-local error lives in rtw_abi.cu)

extern "C" int rtw_render_multi_gpu(const rtw_scene_desc* desc, const rtw_render_cfg* cfg, int32_t ngpus, float* accum_rgba,
                                    rtw_stats* stats) {
  if (!desc || !cfg || !accum_rgba) return rtw_set_error_("rtw_render_multi_gpu: null argument");
  if (ngpus < 1) return rtw_set_error_("rtw_render_multi_gpu: ngpus must be >= 1");
  if (ngpus == 1) return rtw_render(desc, cfg, accum_rgba, stats);
  int ndev = 0;
  if (rtw_device_count(&ndev) != 0) return 2;
  if (ngpus > ndev) return rtw_set_error_("rtw_render_multi_gpu: more GPUs requested than present");
  const int S = cfg->sample_end - cfg->sample_begin;
  const bool rows = (cfg->flags & RTW_FLAG_SPLIT_ROWS) != 0;
  const int tile_rows = cfg->row_tile_rows > 0 ? cfg->row_tile_rows : 8;
  if (S <= 0 || (!rows && S % ngpus != 0)) return rtw_set_error_("rtw_render_multi_gpu: samples must split evenly over the GPUs (reference analogue: render.cpp:174)");
  if (cfg->width < 2 || cfg->height < 2) return rtw_set_error_("render: width and height must be >= 2");
  if (!g_nccl.load()) return rtw_set_error_(g_nccl.error.c_str());

  const size_t npix = static_cast<size_t>(cfg->width) * static_cast<size_t>(cfg->height);
  // row-tile split: every GPU fills a packed buffer of local_rows rows; device 0 additionally holds the gathered buffers
  const size_t local_pix = rows ? static_cast<size_t>(rtw_row_tile_local_rows(cfg->height, tile_rows, ngpus)) * static_cast<size_t>(cfg->width) : npix;
  long long* gathered = nullptr;
  std::vector<rtw_scene*> scenes(ngpus, nullptr);
  std::vector<long long*> fx(ngpus, nullptr);
  std::vector<cudaStream_t> streams(ngpus, nullptr);
  std::vector<ncclComm_t> comms(ngpus, nullptr);
  std::vector<int> devs(ngpus);
  std::vector<int> rcs(ngpus, 0);
  std::vector<std::string> errs(ngpus);
  std::vector<rtw_stats> sts(ngpus);
  for (int g = 0; g < ngpus; ++g) devs[g] = g;

  auto cleanup = [&]() {
    for (int g = 0; g < ngpus; ++g) {
      cudaSetDevice(g);
      if (comms[g]) g_nccl.CommDestroy(comms[g]);
      if (fx[g]) cudaFree(fx[g]);
      if (streams[g]) cudaStreamDestroy(streams[g]);
      if (scenes[g]) rtw_scene_free(scenes[g]);
    }
  };

  // NCCL prints its version banner on stdout when NCCL_DEBUG=VERSION/INFO is set; stdout is the image (P3 text) for the
  // drop-in render(), so route fd 1 to stderr while the communicators are created.
  std::fflush(stdout);
  const int saved_stdout = dup(1);
  if (saved_stdout >= 0) dup2(2, 1);
  int nrc = g_nccl.CommInitAll(comms.data(), ngpus, devs.data());
  if (saved_stdout >= 0) { std::fflush(stdout); dup2(saved_stdout, 1); close(saved_stdout); }
  if (nrc != 0) { cleanup(); return rtw_set_error_((std::string("ncclCommInitAll: ") + g_nccl.GetErrorString(nrc)).c_str()); }

  // one host thread per GPU: upload, zero, render its sample shard
  std::vector<std::thread> th;
  for (int g = 0; g < ngpus; ++g) {
    th.emplace_back([&, g]() {
      rtw_render_cfg c = *cfg;
      c.device = g;
      if (rows) {
        c.row_tile_rows = tile_rows; c.row_tile_count = ngpus; c.row_tile_index = g;
      } else {
        c.sample_begin = cfg->sample_begin + g * (S / ngpus);
        c.sample_end = c.sample_begin + S / ngpus;
      }
      rcs[g] = rtw_scene_upload(desc, g, &scenes[g]);
      if (rcs[g]) { errs[g] = rtw_last_error(); return; }
      if (cudaStreamCreate(&streams[g]) != cudaSuccess || cudaMalloc(reinterpret_cast<void**>(&fx[g]), std::max(npix, local_pix) * 4 * sizeof(long long)) != cudaSuccess ||
          cudaMemsetAsync(fx[g], 0, local_pix * 4 * sizeof(long long), streams[g]) != cudaSuccess) {
        rcs[g] = 2; errs[g] = "device allocation failed"; return;
      }
      rcs[g] = rtw_render_device(scenes[g], &c, reinterpret_cast<int64_t*>(fx[g]), streams[g], &sts[g]);
      if (rcs[g]) errs[g] = rtw_last_error();
    });
  }
  for (auto& t : th) t.join();
  for (int g = 0; g < ngpus; ++g)
    if (rcs[g]) { cleanup(); return rtw_set_error_(("GPU " + std::to_string(g) + ": " + errs[g]).c_str()); }

  if (!rows) {
    // the one collective of the path: sum of the accumulation buffers onto device 0 (in place on the root)
    g_nccl.GroupStart();
    for (int g = 0; g < ngpus; ++g) {
      cudaSetDevice(g);
      nrc = g_nccl.Reduce(fx[g], fx[g], npix * 4, kNcclInt64, kNcclSum, 0, comms[g], streams[g]);
      if (nrc != 0) break;
    }
    const int nrc2 = g_nccl.GroupEnd();
    if (nrc != 0 || nrc2 != 0) { cleanup(); return rtw_set_error_((std::string("ncclReduce: ") + g_nccl.GetErrorString(nrc ? nrc : nrc2)).c_str()); }
  } else {
    // row-tile alternative: gather the packed buffers on device 0 (send/recv pairs in one group), then put the tiles in place
    cudaSetDevice(0);
    if (cudaMalloc(reinterpret_cast<void**>(&gathered), static_cast<size_t>(ngpus) * local_pix * 4 * sizeof(long long)) != cudaSuccess) {
      cleanup(); return rtw_set_error_("cudaMalloc (gather buffer) failed");
    }
    g_nccl.GroupStart();
    for (int g = 0; g < ngpus && nrc == 0; ++g) {
      cudaSetDevice(g);
      nrc = g_nccl.Send(fx[g], local_pix * 4, kNcclInt64, 0, comms[g], streams[g]);
      if (nrc == 0) { cudaSetDevice(0); nrc = g_nccl.Recv(gathered + static_cast<size_t>(g) * local_pix * 4, local_pix * 4, kNcclInt64, g, comms[0], streams[0]); }
    }
    const int nrc2 = g_nccl.GroupEnd();
    if (nrc != 0 || nrc2 != 0) { cudaSetDevice(0); cudaFree(gathered); cleanup(); return rtw_set_error_((std::string("ncclSend/Recv: ") + g_nccl.GetErrorString(nrc ? nrc : nrc2)).c_str()); }
    cudaSetDevice(0);
    if (rtw_untile_accum(reinterpret_cast<const int64_t*>(gathered), reinterpret_cast<int64_t*>(fx[0]), cfg->width, cfg->height, tile_rows, ngpus, 0, streams[0]) != 0) {
      cudaFree(gathered); cleanup(); return 2;
    }
  }
  for (int g = 0; g < ngpus; ++g) { cudaSetDevice(g); cudaStreamSynchronize(streams[g]); }

  cudaSetDevice(0);
  float* out = nullptr;
  int rc = 0;
  if (cudaMalloc(reinterpret_cast<void**>(&out), npix * 4 * sizeof(float)) != cudaSuccess) rc = rtw_set_error_("cudaMalloc failed");
  if (!rc) rc = rtw_accum_to_float(reinterpret_cast<const int64_t*>(fx[0]), out, static_cast<int64_t>(npix), 0, streams[0]);
  if (!rc && cudaMemcpyAsync(accum_rgba, out, npix * 4 * sizeof(float), cudaMemcpyDeviceToHost, streams[0]) != cudaSuccess) rc = rtw_set_error_("D2H failed");
  if (!rc && cudaStreamSynchronize(streams[0]) != cudaSuccess) rc = rtw_set_error_("stream sync failed");
  if (out) cudaFree(out);
  if (gathered) cudaFree(gathered);
  if (stats && !rc) {
    std::memset(stats, 0, sizeof *stats);
    for (int g = 0; g < ngpus; ++g) {
      stats->paths += sts[g].paths; stats->rays += sts[g].rays;
      stats->sphere_tests += sts[g].sphere_tests; stats->sphere_candidates += sts[g].sphere_candidates;
      stats->tri_tests += sts[g].tri_tests; stats->node_visits += sts[g].node_visits;
      if (sts[g].kernel_ms > stats->kernel_ms) stats->kernel_ms = sts[g].kernel_ms;  // max over GPUs
    }
    stats->kernel_used = sts[0].kernel_used;
    stats->bvh_variant = sts[0].bvh_variant;
    stats->launches = ngpus + 1 + (rows ? 1 : 0);
  }
  cleanup();
  return rc;
}
