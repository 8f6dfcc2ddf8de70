// rtw_multi.cu -- single-process multi-GPU driver behind rtw_render_multi_gpu / rtw_render_multi_gpu_rgb8 (include/rtw_b200.h).
// Replaces the reference's thread fan-out + image sum (render.cpp:169-180): samples-per-pixel are split over the GPUs
// (the global sample index keys the Philox stream, so the union of samples does not depend on the split), each GPU renders into
// its own int64 fixed-point accumulation buffer, and the buffers are combined over NVLink peer memory by ONE kernel per GPU:
// GPU g owns the g-th slice of the image, reads that slice of every GPU's buffer with peer loads (a reduce-scatter by direct
// loads through NVSwitch), sums the integers, converts (float sums, or write_color -> rgb8) and downloads its slice into the
// caller's host buffer over its own PCIe link.  Integer sums are exact, so the result is bit-identical to one GPU.
// No communicator, no library: the only set-up is cudaDeviceEnablePeerAccess, done once per device pair.  (The one-process-per-GPU
// launch of bench.py / torchrun combines the same buffers with one NCCL reduce instead.)
// The row-tile alternative needs no exchange at all: every GPU converts and downloads its own tiles.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "rtw_host.h"

namespace {

constexpr int kMaxGpus = 16;
struct PeerBufs {
  const longlong4* p[kMaxGpus];
  int n;
};

// One thread per pixel of this GPU's slice: sum the n accumulation buffers (n - 1 of them in peer memory), then either the float
// accumulation pixel or write_color (render.cpp:11-20) of the exact sum.  out index = pixel - pix0 (slice-local).
__global__ void __launch_bounds__(256) k_combine_peers(const PeerBufs b, long long pix0, long long count, float4* __restrict__ out_f32,
                                                      uint8_t* __restrict__ out_u8, double spp) {
  const long long k = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (k >= count) return;
  long long r = 0, g = 0, bl = 0, w = 0;
  for (int i = 0; i < b.n; ++i) {
    const longlong4 v = b.p[i][pix0 + k];
    r += v.x; g += v.y; bl += v.z; w += v.w;
  }
  const double s = 1.0 / 4294967296.0;
  if (out_f32) out_f32[k] = make_float4(static_cast<float>(r * s), static_cast<float>(g * s), static_cast<float>(bl * s), static_cast<float>(w));
  if (out_u8) {
    const double c[3] = {sqrt(r * s / spp), sqrt(g * s / spp), sqrt(bl * s / spp)};
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
      const double cl = c[ch] < 0.0 ? 0.0 : (c[ch] > 0.999 ? 0.999 : c[ch]);
      out_u8[3 * k + ch] = static_cast<uint8_t>(static_cast<int>(256 * cl));
    }
  }
}

double now_ms() {
  using namespace std::chrono;
  return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}

int multi_host(const rtw_scene_desc* desc, const rtw_render_cfg* cfg, int32_t ngpus, float* out_f32, uint8_t* out_u8, rtw_stats* stats) {
  using rtw::fail;
  if (!desc || !cfg || (!out_f32 && !out_u8)) return fail("rtw_render_multi_gpu: null argument");
  if (ngpus < 1) return fail("rtw_render_multi_gpu: ngpus must be >= 1");
  if (ngpus == 1) return out_f32 ? rtw_render(desc, cfg, out_f32, stats) : rtw_render_rgb8(desc, cfg, out_u8, stats);
  int ndev = 0;
  if (rtw_device_count(&ndev) != 0) return 2;
  if (ngpus > ndev) return fail("rtw_render_multi_gpu: more GPUs requested than present");
  if (ngpus > kMaxGpus) return fail("rtw_render_multi_gpu: at most 16 GPUs");
  const int S = cfg->sample_end - cfg->sample_begin;
  const bool rows = (cfg->flags & RTW_FLAG_SPLIT_ROWS) != 0;
  const int tile_rows = cfg->row_tile_rows > 0 ? cfg->row_tile_rows : 8;
  if (S <= 0 || cfg->sample_begin < 0) return fail("render: empty sample range");
  if (cfg->width < 2 || cfg->height < 2) return fail("render: width and height must be >= 2");
  if (desc->nprims < 0 || desc->nmats < 0 || (desc->nprims > 0 && !desc->prims) || (desc->nmats > 0 && !desc->mats)) return fail("rtw_scene_upload: invalid scene description");
  const double t_start = now_ms();
  rtw::prewarm_join();

  const size_t npix = static_cast<size_t>(cfg->width) * static_cast<size_t>(cfg->height);
  const int local_rows = rows ? rtw_row_tile_local_rows(cfg->height, tile_rows, ngpus) : cfg->height;
  const size_t local_pix = static_cast<size_t>(local_rows) * static_cast<size_t>(cfg->width);
  const bool gpu_build = rtw::choose_gpu_build(desc, cfg);
  const uint64_t key = rtw::scene_key(desc) ^ (gpu_build ? 0x6b9d0f5a1c2e3d47ull : 0ull);
  const bool use_cache = (cfg->flags & RTW_FLAG_NO_SCENE_CACHE) == 0;

  std::vector<rtw::DeviceSlot*> slots(ngpus, nullptr);
  for (int g = 0; g < ngpus; ++g) {
    slots[g] = rtw::device_slot(g);
    if (!slots[g]) return 1;
  }
  // every GPU's slot stays locked for the whole call (in device order: no deadlock with a concurrent multi-GPU call)
  std::vector<std::unique_lock<std::mutex>> locks;
  for (int g = 0; g < ngpus; ++g) locks.emplace_back(slots[g]->m);

  rtw::HostFlat flat;       // flattened at most once, by whichever GPU thread misses its cache first
  flat.gpu_build = gpu_build;
  std::mutex flat_mutex;
  std::vector<int> rcs(ngpus, 0);
  std::vector<std::string> errs(ngpus);
  std::vector<rtw_stats> sts(ngpus);
  std::vector<char> hits(ngpus, 0);
  std::vector<double> t_scene(ngpus, 0.0);
  const bool trace = std::getenv("RTW_TRACE") != nullptr;   // per-GPU timeline of the call on stderr
  struct Mark { double ctx = 0, peer = 0, scene = 0, alloc = 0, done = 0; };
  std::vector<Mark> marks(ngpus);
  const double t_joined = now_ms();

  // one host thread per GPU: context, peer access, scene (cached), zero, render its shard (waits for its kernel)
  auto work = [&](int g) {
    rtw::DeviceSlot* s = slots[g];
    auto bail = [&](int rc) { rcs[g] = rc ? rc : 1; errs[g] = rtw_last_error(); };
    if (int rc = rtw::slot_prepare(s)) return bail(rc);
    marks[g].ctx = now_ms();
    if (!rows) {
      for (int q = 0; q < ngpus; ++q) {
        if (q == g || (s->peer_enabled_mask >> q & 1)) continue;
        int can = 0;
        cudaDeviceCanAccessPeer(&can, g, q);
        if (!can) { fail("rtw_render_multi_gpu: GPUs " + std::to_string(g) + " and " + std::to_string(q) + " have no peer access (NVLink / NVSwitch needed for the sample split; use the row split)"); return bail(1); }
        const cudaError_t e = cudaDeviceEnablePeerAccess(q, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { rtw::fail_cuda("cudaDeviceEnablePeerAccess", e); return bail(2); }
        cudaGetLastError();
        s->peer_enabled_mask |= 1 << q;
      }
    }
    marks[g].peer = now_ms();
    bool hit = false;
    if (int rc = rtw::slot_set_scene(s, desc, key, use_cache, &flat, &flat_mutex, &hit)) return bail(rc);
    hits[g] = hit ? 1 : 0;
    t_scene[g] = marks[g].scene = now_ms();
    if (cudaSuccess != s->fx.reserve(std::max(npix, local_pix) * 4)) { fail("rtw_render_multi_gpu: device allocation failed"); return bail(2); }
    if (cudaSuccess != cudaMemsetAsync(s->fx.p, 0, (rows ? local_pix : npix) * 4 * sizeof(long long), s->stream)) { fail("cudaMemsetAsync failed"); return bail(2); }
    marks[g].alloc = now_ms();
    rtw_render_cfg c = *cfg;
    c.device = g;
    if (rows) {
      c.row_tile_rows = tile_rows; c.row_tile_count = ngpus; c.row_tile_index = g;
    } else {
      // samples split as evenly as possible: the first S % ngpus GPUs render one more (the global sample index keys Philox)
      const int q = S / ngpus, r = S % ngpus;
      c.sample_begin = cfg->sample_begin + g * q + std::min(g, r);
      c.sample_end = c.sample_begin + q + (g < r ? 1 : 0);
    }
    std::memset(&sts[g], 0, sizeof(rtw_stats));
    if (c.sample_end > c.sample_begin) {
      if (int rc = rtw_render_device(&s->scene, &c, reinterpret_cast<int64_t*>(s->fx.p), s->stream, &sts[g])) return bail(rc);
    } else if (cudaStreamSynchronize(s->stream) != cudaSuccess) { fail("stream sync failed"); return bail(2); }
    marks[g].done = now_ms();
  };
  {
    std::vector<std::thread> th;
    for (int g = 1; g < ngpus; ++g) th.emplace_back(work, g);
    work(0);
    for (auto& t : th) t.join();
  }
  for (int g = 0; g < ngpus; ++g)
    if (rcs[g]) return fail("GPU " + std::to_string(g) + ": " + errs[g]);
  const double t_rendered = now_ms();

  // ---- combine + convert + download: every GPU handles its own part of the image ------------------------------------------------------
  const double spp = static_cast<double>(S);
  int launches = 0;
  for (int g = 0; g < ngpus; ++g) {
    rtw::DeviceSlot* s = slots[g];
    RTW_CUDA(cudaSetDevice(g));
    if (!rows) {
      const long long pix0 = static_cast<long long>(npix * g / ngpus), pix1 = static_cast<long long>(npix * (g + 1) / ngpus), cnt = pix1 - pix0;
      if (cnt <= 0) continue;
      PeerBufs pb{};
      pb.n = ngpus;
      for (int q = 0; q < ngpus; ++q) pb.p[q] = reinterpret_cast<const longlong4*>(slots[(g + q) % ngpus]->fx.p);  // own buffer first, peers staggered
      if (out_f32) RTW_CUDA(s->out_f32.reserve(static_cast<size_t>(cnt) * 4));
      if (out_u8) RTW_CUDA(s->out_u8.reserve(static_cast<size_t>(cnt) * 3));
      k_combine_peers<<<static_cast<unsigned>((cnt + 255) / 256), 256, 0, s->stream>>>(pb, pix0, cnt, out_f32 ? reinterpret_cast<float4*>(s->out_f32.p) : nullptr,
                                                                                      out_u8 ? s->out_u8.p : nullptr, spp);
      RTW_CUDA(cudaGetLastError());
      rtw::count_launch();
      ++launches;
      if (out_f32) RTW_CUDA(cudaMemcpyAsync(out_f32 + 4 * pix0, s->out_f32.p, static_cast<size_t>(cnt) * 4 * sizeof(float), cudaMemcpyDeviceToHost, s->stream));
      if (out_u8) RTW_CUDA(cudaMemcpyAsync(out_u8 + 3 * pix0, s->out_u8.p, static_cast<size_t>(cnt) * 3, cudaMemcpyDeviceToHost, s->stream));
    } else {
      // packed local tiles -> converted in place order, then one copy per tile into its rows of the host image
      if (out_f32) {
        RTW_CUDA(s->out_f32.reserve(local_pix * 4));
        if (int rc = rtw_accum_to_float(reinterpret_cast<const int64_t*>(s->fx.p), s->out_f32.p, static_cast<int64_t>(local_pix), g, s->stream)) return rc;
      }
      if (out_u8) {
        RTW_CUDA(s->out_u8.reserve(local_pix * 3));
        if (int rc = rtw_finalize_rgb8_device(reinterpret_cast<const int64_t*>(s->fx.p), static_cast<int64_t>(local_pix), S, g, s->stream, s->out_u8.p)) return rc;
      }
      ++launches;
      const int tiles_local = local_rows / tile_rows;
      for (int k = 0; k < tiles_local; ++k) {
        const long long row0 = (static_cast<long long>(k) * ngpus + g) * tile_rows;
        if (row0 >= cfg->height) break;
        const size_t nrows = static_cast<size_t>(std::min<long long>(tile_rows, cfg->height - row0));
        const size_t src = static_cast<size_t>(k) * tile_rows * cfg->width, dst = static_cast<size_t>(row0) * cfg->width, n = nrows * cfg->width;
        if (out_f32) RTW_CUDA(cudaMemcpyAsync(out_f32 + 4 * dst, s->out_f32.p + 4 * src, n * 4 * sizeof(float), cudaMemcpyDeviceToHost, s->stream));
        if (out_u8) RTW_CUDA(cudaMemcpyAsync(out_u8 + 3 * dst, s->out_u8.p + 3 * src, n * 3, cudaMemcpyDeviceToHost, s->stream));
      }
    }
  }
  for (int g = 0; g < ngpus; ++g) {
    RTW_CUDA(cudaSetDevice(g));
    RTW_CUDA(cudaStreamSynchronize(slots[g]->stream));
  }
  const double t_end = now_ms();
  if (trace) {
    std::fprintf(stderr, "rtw trace: prewarm join %.1f ms; per GPU (ms since call): context+stream, peer access, scene, buffers, kernel done\n", t_joined - t_start);
    for (int g = 0; g < ngpus; ++g)
      std::fprintf(stderr, "rtw trace: gpu %d  %.1f  %.1f  %.1f  %.1f  %.1f\n", g, marks[g].ctx - t_start, marks[g].peer - t_start, marks[g].scene - t_start,
                   marks[g].alloc - t_start, marks[g].done - t_start);
    std::fprintf(stderr, "rtw trace: rendered %.1f, combined + downloaded %.1f\n", t_rendered - t_start, t_end - t_start);
  }
  if (stats) {
    std::memset(stats, 0, sizeof *stats);
    double t_up = t_start;
    int all_hit = 1;
    for (int g = 0; g < ngpus; ++g) {
      stats->paths += sts[g].paths; stats->rays += sts[g].rays;
      stats->sphere_tests += sts[g].sphere_tests; stats->sphere_candidates += sts[g].sphere_candidates;
      stats->tri_tests += sts[g].tri_tests; stats->node_visits += sts[g].node_visits;
      if (sts[g].kernel_ms > stats->kernel_ms) stats->kernel_ms = sts[g].kernel_ms;  // max over GPUs
      t_up = std::max(t_up, t_scene[g]);
      all_hit &= hits[g];
    }
    stats->kernel_used = sts[0].kernel_used;
    stats->bvh_variant = sts[0].bvh_variant;
    stats->launches = ngpus + launches;
    stats->h2d_ms = t_up - t_start;          // contexts + peer access + flatten + upload on the slowest GPU
    stats->d2h_ms = t_end - t_rendered;      // combine + convert + download
    stats->total_ms = t_end - t_start;
    stats->scene_cache_hit = all_hit;
    stats->bvh_build_gpu_ms = slots[0]->scene.gpu_build_ms;
  }
  return 0;
}

}  // namespace

extern "C" int rtw_render_multi_gpu(const rtw_scene_desc* desc, const rtw_render_cfg* cfg, int32_t ngpus, float* accum_rgba, rtw_stats* stats) {
  if (!accum_rgba) return rtw::fail("rtw_render_multi_gpu: null argument");
  return multi_host(desc, cfg, ngpus, accum_rgba, nullptr, stats);
}

extern "C" int rtw_render_multi_gpu_rgb8(const rtw_scene_desc* desc, const rtw_render_cfg* cfg, int32_t ngpus, uint8_t* rgb8, rtw_stats* stats) {
  if (!rgb8) return rtw::fail("rtw_render_multi_gpu_rgb8: null argument");
  return multi_host(desc, cfg, ngpus, nullptr, rgb8, stats);
}
