// rtw_host.h -- host-side internals shared by the translation units behind the C ABI (rtw_abi.cu, rtw_multi.cu):
// the flattened host arena, the device-resident scene, and the per-device slots that keep a scene and the accumulation
// buffers alive between calls (so that rtw_render / rtw_render_multi_gpu cost one hash of the caller's arrays when the
// scene has not changed, instead of a flatten + BVH build + upload: render.cpp:146 builds its BVH once per render() too).
#pragma once
#include <atomic>
#include <cstdint>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "rtw_internal.h"

namespace rtw {

template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  ~DevBuf() { if (p) cudaFree(p); }
  cudaError_t alloc(size_t count) {
    if (p) { cudaFree(p); p = nullptr; }
    n = 0;
    if (count == 0) return cudaSuccess;
    const cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&p), count * sizeof(T));
    if (e == cudaSuccess) n = count;
    return e;
  }
  cudaError_t reserve(size_t count) { return count <= n ? cudaSuccess : alloc(count + count / 8); }  // grow-only
  cudaError_t upload(const std::vector<T>& v, cudaStream_t s = nullptr) {
    cudaError_t e = alloc(v.size());
    if (e != cudaSuccess || v.empty()) return e;
    return cudaMemcpyAsync(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, s);
  }
};

struct EventPair {  // destroyed on every exit path
  cudaEvent_t a = nullptr, b = nullptr;
  cudaError_t create() {
    cudaError_t e = cudaEventCreate(&a);
    return e != cudaSuccess ? e : cudaEventCreate(&b);
  }
  ~EventPair() { if (a) cudaEventDestroy(a); if (b) cudaEventDestroy(b); }
};

// Flatten (north_star item 1): variant/virtual primitive list -> SoA tables, BVH, materials, all in one host arena whose
// layout is the device layout.  Pure host code.
struct HostFlat {
  std::unique_ptr<unsigned char[]> host;
  size_t bytes = 0;          // size of the device arena
  size_t upload_bytes = 0;   // leading part of it that `host` holds and that is copied (all of it unless the device builds the BVH)
  size_t o_sA = 0, o_sB = 0, o_sId = 0, o_big = 0, o_tri = 0, o_triId = 0, o_nodes = 0, o_refs = 0, o_matA = 0, o_matB = 0, o_ctr = 0, o_cw = 0;
  int32_t n_static = 0, n_moving = 0, n_big = 0, n_tri = 0, n_nodes = 0, leaf_direct = 0, bvh_depth = 0;
  int32_t n_cw = 0, n_records = 0, cw_has_spheres = 0;   // compressed wide BVH (scenes with triangles): nodes, leaf-ordered records
  size_t n_leaf_refs = 0;
  double bvh_ms = 0.0;
  // build on the device (rtw_build.cu): flatten_host leaves the node table empty and keeps its build records for upload_flat
  bool gpu_build = false;
  std::shared_ptr<void> gpu_items;   // std::vector<BvhBuilder::Item>
  size_t n_gpu_items = 0;
};
int flatten_host(const rtw_scene_desc* desc, HostFlat* hf);          // 0 or an error code with rtw_last_error() set
bool mesh_bvh_is_cw8();                                               // RTW_MESH_BVH=cw8: compressed 8-wide BVH for scenes with triangles
uint64_t scene_key(const rtw_scene_desc* desc);                       // 64-bit hash of prims, mats and camera (never 0)

}  // namespace rtw

// One launch of a render kernel uses one of kCtrSlots counter blocks of its scene (work-queue cursor + statistics), handed out
// round robin: up to kCtrSlots renders of the same uploaded scene may be in flight at once, on any streams or threads.
struct rtw_scene {
  int device = 0;
  int sm_count = 0;
  rtw::DevBuf<unsigned char> arena;    // every table of the scene in ONE allocation, filled by ONE host-to-device copy
  unsigned char* arena_ptr = nullptr;  // = arena.p, or memory borrowed from a device slot
  unsigned long long* counters = nullptr;  // [kCtrSlots][kCtrCount]
  mutable std::atomic<uint32_t> launch_seq{0};
  size_t n_leaf_refs = 0;
  rtw::DevScene dev{};
  int64_t nprims = 0;
  bool has_triangles = false;
  size_t smem_bytes = 0;
  size_t arena_bytes = 0;
  int bvh_depth = 0;
  double gpu_build_ms = 0.0;   // > 0: the BVH was built on the device (rtw_build.cu), time of the build kernels
};

namespace rtw {

// rtw_build.cu: linear BVH on the device.  items_host: n >= 2 BvhBuilder::Item records (rtw_bvh.h) in host memory; nodes_out: device
// memory for n - 1 PackedNodes (root = node 0).  Blocks until the tree is built; *depth_out <- deepest root-to-leaf path.
void release_build_scratch();   // frees the per-device scratch allocation of gpu_build_bvh
int gpu_build_bvh(const void* items_host, size_t n, void* nodes_out, cudaStream_t stream, bool sah_top, bool sah_clusters, int* depth_out, double* build_ms,
                  int* top_nodes_out, int* clusters_rebuilt_out);

int upload_flat(const HostFlat& hf, const rtw_camera& cam, int64_t nprims, int device, rtw_scene* sc, DevBuf<unsigned char>* borrowed,
                cudaStream_t stream);

// Per-device state kept between host-buffer renders (grow-only; rtw_release_cached_buffers frees it).
struct DeviceSlot {
  std::mutex m;                 // one host-buffer render per device at a time
  int device = -1;
  cudaStream_t stream = nullptr;
  cudaEvent_t rendered = nullptr;   // recorded after the render kernel (multi-GPU combine waits on it)
  DevBuf<unsigned char> arena;  // scene tables
  rtw_scene scene;              // borrows `arena`
  uint64_t key = 0;             // scene_key of what `scene` holds; 0 = nothing
  DevBuf<long long> fx;         // int64 accumulation buffer [npix][4]
  DevBuf<float> out_f32;        // float accumulation buffer [npix][4]
  DevBuf<uint8_t> out_u8;       // rgb8 [npix][3]
  int peer_enabled_mask = 0;    // bit g: cudaDeviceEnablePeerAccess(g) done from this device
};
DeviceSlot* device_slot(int device);   // nullptr + error when the ordinal is out of range
int slot_prepare(DeviceSlot* s);        // cudaSetDevice + stream/event creation on first use
// Makes s->scene hold `desc`: a hit costs nothing; a miss uploads `*flat` (flattening `desc` into it first when it is empty).
int slot_set_scene(DeviceSlot* s, const rtw_scene_desc* desc, uint64_t key, bool use_cache, HostFlat* flat, std::mutex* flat_mutex, bool* hit);

bool choose_gpu_build(const rtw_scene_desc* desc, const rtw_render_cfg* cfg);   // host or device BVH build for a host-buffer render
void prewarm_join();                    // waits for the context-creation threads of rtw_prewarm
int fail(const std::string& msg);
int fail_cuda(const char* what, cudaError_t e);

}  // namespace rtw

#define RTW_CUDA(call)                                          \
  do {                                                          \
    cudaError_t e__ = (call);                                   \
    if (e__ != cudaSuccess) return rtw::fail_cuda(#call, e__);  \
  } while (0)
