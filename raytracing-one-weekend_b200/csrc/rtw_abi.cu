// rtw_abi.cu -- host side of the C ABI declared in include/rtw_b200.h: flatten the caller's scene description
// into the device SoA layout, build the BVH, manage device memory, launch the kernels of rtw_kernels.cu.
// There is deliberately no CPU implementation behind any entry point: without a CUDA device every call fails.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "rtw_bvh.h"
#include "rtw_internal.h"

namespace {

thread_local std::string g_last_error;

int fail(const std::string& msg) {
  g_last_error = msg;
  return 1;
}
int fail_cuda(const char* what, cudaError_t e) {
  g_last_error = std::string(what) + ": " + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ")";
  return 2;
}
#define RTW_CUDA(call)                                     \
  do {                                                     \
    cudaError_t e__ = (call);                              \
    if (e__ != cudaSuccess) return fail_cuda(#call, e__);  \
  } while (0)

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
    n = count;
    if (count == 0) return cudaSuccess;
    return cudaMalloc(reinterpret_cast<void**>(&p), count * sizeof(T));
  }
  cudaError_t upload(const std::vector<T>& v, cudaStream_t s = nullptr) {
    cudaError_t e = alloc(v.size());
    if (e != cudaSuccess || v.empty()) return e;
    return cudaMemcpyAsync(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, s);
  }
};

float __int_as_float_host(int v) {
  float f;
  std::memcpy(&f, &v, sizeof f);
  return f;
}

double now_ms() {
  using namespace std::chrono;
  return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}

}  // namespace

struct rtw_scene {
  int device = 0;
  int sm_count = 0;
  DevBuf<unsigned char> arena;  // every table of the scene in ONE allocation, filled by ONE host-to-device copy
  unsigned char* arena_ptr = nullptr;  // = arena.p, or memory borrowed from rtw_render's per-device cache
  unsigned long long* counters = nullptr;
  size_t n_leaf_refs = 0;
  rtw::DevScene dev{};
  int64_t nprims = 0;
  bool has_triangles = false;
  size_t smem_bytes = 0;
};

namespace {

// Flatten (north_star item 1): variant/virtual primitive list -> SoA tables, BVH, materials, all in one host arena whose
// layout is the device layout.  Pure host code: no CUDA call in here (rtw_flatten_info exposes it to CPU-only tests).
struct HostFlat {
  std::vector<unsigned char> host;
  size_t o_sA = 0, o_sB = 0, o_sId = 0, o_big = 0, o_tri = 0, o_triId = 0, o_nodes = 0, o_refs = 0, o_matA = 0, o_matB = 0, o_ctr = 0;
  int32_t n_static = 0, n_moving = 0, n_big = 0, n_tri = 0, n_nodes = 0, leaf_direct = 0;
  size_t n_leaf_refs = 0;
  double bvh_ms = 0.0;
};

int flatten_host(const rtw_scene_desc* desc, HostFlat* hf) {
  if (!desc || desc->nprims < 0 || desc->nmats < 0 || (desc->nprims > 0 && !desc->prims) || (desc->nmats > 0 && !desc->mats))
    return fail("rtw_scene_upload: invalid scene description");
  if (desc->nprims >= (1ll << 28)) return fail("rtw_scene_upload: too many primitives");
  std::vector<float4> sA_static, sB_static, sA_moving, sB_moving;
  std::vector<int2> id_static, id_moving;
  std::vector<rtw::BigSphere> big;
  std::vector<float4> tri;
  std::vector<int2> triId;

  // scene bound for the conservative slack of the reject test (see rtw_kernels.cu trace_spheres)
  double bound = 0.0;
  for (int k = 0; k < 3; ++k) bound = std::max(bound, std::fabs(desc->camera.origin[k]));
  for (int64_t i = 0; i < desc->nprims; ++i) {
    const rtw_primitive& P = desc->prims[i];
    if (P.material < 0 || P.material >= desc->nmats) return fail("rtw_scene_upload: primitive references a missing material");
    if (P.kind == RTW_TRIANGLE) {
      for (int k = 0; k < 3; ++k) bound = std::max({bound, std::fabs(P.a[k]), std::fabs(P.b[k]), std::fabs(P.c[k])});
    } else if (P.kind == RTW_SPHERE || P.kind == RTW_MOVING_SPHERE) {
      if (std::fabs(P.radius) >= rtw::kBigRadius) continue;
      for (int k = 0; k < 3; ++k) {
        bound = std::max(bound, std::fabs(P.a[k]) + std::fabs(P.radius));
        if (P.kind == RTW_MOVING_SPHERE) bound = std::max(bound, std::fabs(P.b[k]) + std::fabs(P.radius));
      }
    } else {
      return fail("rtw_scene_upload: unknown primitive kind");
    }
  }
  const double E = 24.0 * 1.1920929e-7 * bound;

  for (int64_t i = 0; i < desc->nprims; ++i) {
    const rtw_primitive& P = desc->prims[i];
    const int id = static_cast<int>(i);
    if (P.kind == RTW_TRIANGLE) {
      const double e1[3] = {P.b[0] - P.a[0], P.b[1] - P.a[1], P.b[2] - P.a[2]};
      const double e2[3] = {P.c[0] - P.a[0], P.c[1] - P.a[1], P.c[2] - P.a[2]};
      const double n[3] = {e1[1] * e2[2] - e2[1] * e1[2], e1[2] * e2[0] - e2[2] * e1[0], e1[0] * e2[1] - e2[0] * e1[1]};
      tri.push_back(make_float4((float)P.a[0], (float)P.a[1], (float)P.a[2], (float)n[0]));
      tri.push_back(make_float4((float)e1[0], (float)e1[1], (float)e1[2], (float)n[1]));
      tri.push_back(make_float4((float)e2[0], (float)e2[1], (float)e2[2], (float)n[2]));
      triId.push_back(make_int2(id, P.material));
      continue;
    }
    const bool moving = P.kind == RTW_MOVING_SPHERE && (P.a[0] != P.b[0] || P.a[1] != P.b[1] || P.a[2] != P.b[2]);
    const double dc[3] = {moving ? P.b[0] - P.a[0] : 0.0, moving ? P.b[1] - P.a[1] : 0.0, moving ? P.b[2] - P.a[2] : 0.0};
    if (std::fabs(P.radius) >= rtw::kBigRadius) {
      rtw::BigSphere b{};
      for (int k = 0; k < 3; ++k) { b.c0[k] = P.a[k]; b.dc[k] = dc[k]; }
      b.r = P.radius; b.prim_id = id; b.material = P.material;
      big.push_back(b);
      continue;
    }
    const double r = std::fabs(P.radius);
    const float r2c = static_cast<float>((r * r + 3.0 * r * E + E * E) * (1.0 + 4e-7));
    const float4 A = make_float4((float)P.a[0], (float)P.a[1], (float)P.a[2], r2c);
    const float4 B = make_float4((float)dc[0], (float)dc[1], (float)dc[2], (float)P.radius);
    if (moving) { sA_moving.push_back(A); sB_moving.push_back(B); id_moving.push_back(make_int2(id, P.material)); }
    else { sA_static.push_back(A); sB_static.push_back(B); id_static.push_back(make_int2(id, P.material)); }
  }
  std::vector<float4> sA(sA_static), sB(sB_static);
  std::vector<int2> sId(id_static);
  sA.insert(sA.end(), sA_moving.begin(), sA_moving.end());
  sB.insert(sB.end(), sB_moving.begin(), sB_moving.end());
  sId.insert(sId.end(), id_moving.begin(), id_moving.end());

  // BVH over small spheres (swept bounds for moving ones, common-model.cpp:197-207) and triangles
  std::vector<rtw::Box3> boxes;
  std::vector<uint32_t> refs;
  boxes.reserve(sA.size() + triId.size());
  for (size_t i = 0; i < sA.size(); ++i) {
    const float r = std::fabs(sB[i].w);
    const float c0[3] = {sA[i].x, sA[i].y, sA[i].z};
    const float c1[3] = {sA[i].x + sB[i].x, sA[i].y + sB[i].y, sA[i].z + sB[i].z};
    rtw::Box3 b;
    for (int k = 0; k < 3; ++k) { b.lo[k] = std::min(c0[k], c1[k]) - r; b.hi[k] = std::max(c0[k], c1[k]) + r; }
    boxes.push_back(b);
    refs.push_back(static_cast<uint32_t>(i));
  }
  for (size_t i = 0; i < triId.size(); ++i) {
    const float4 q0 = tri[3 * i], q1 = tri[3 * i + 1], q2 = tri[3 * i + 2];
    const float a[3] = {q0.x, q0.y, q0.z};
    const float b1[3] = {q0.x + q1.x, q0.y + q1.y, q0.z + q1.z};
    const float c1[3] = {q0.x + q2.x, q0.y + q2.y, q0.z + q2.z};
    rtw::Box3 b; b.reset(); b.grow(a); b.grow(b1); b.grow(c1);
    boxes.push_back(b);
    refs.push_back((1u << 30) | static_cast<uint32_t>(i));
  }
  rtw::BvhBuilder builder;
  if (const char* e = std::getenv("RTW_BVH_LEAF")) builder.kMaxLeaf = std::min(std::max(std::atoi(e), 1), 31);  // tuning knob
  const double t_bvh = now_ms();
  builder.build(boxes, refs);
  hf->bvh_ms = now_ms() - t_bvh;
  std::vector<float4> nodes(builder.nodes().size() * 4);
  if (!nodes.empty()) std::memcpy(nodes.data(), builder.nodes().data(), nodes.size() * sizeof(float4));

  std::vector<float4> matA(static_cast<size_t>(desc->nmats));
  std::vector<float2> matB(static_cast<size_t>(desc->nmats));
  for (int64_t i = 0; i < desc->nmats; ++i) {
    const rtw_material& m = desc->mats[i];
    if (m.kind < RTW_LAMBERTIAN || m.kind > RTW_DIELECTRIC) return fail("rtw_scene_upload: unknown material kind");
    const double fuzz = std::min(std::max(m.fuzz, 0.0), 1.0);  // common-model.h:132-133,143-144
    matA[i] = make_float4((float)m.albedo[0], (float)m.albedo[1], (float)m.albedo[2], (float)fuzz);
    matB[i] = make_float2((float)m.ior, __int_as_float_host(m.kind));
  }

  // ---- one arena, one copy ---------------------------------------------------------------------------------------------
  std::vector<unsigned char>& host = hf->host;
  host.clear();
  auto put = [&host](const void* src, size_t bytes) {
    const size_t off = (host.size() + 255) & ~size_t(255);
    host.resize(off + bytes);
    if (bytes) std::memcpy(host.data() + off, src, bytes);
    return off;
  };
  const unsigned long long zero_counters[rtw::kCtrCount] = {};
  const size_t o_sA = put(sA.data(), sA.size() * sizeof(float4)), o_sB = put(sB.data(), sB.size() * sizeof(float4)),
               o_sId = put(sId.data(), sId.size() * sizeof(int2)), o_big = put(big.data(), big.size() * sizeof(rtw::BigSphere)),
               o_tri = put(tri.data(), tri.size() * sizeof(float4)), o_triId = put(triId.data(), triId.size() * sizeof(int2)),
               o_nodes = put(nodes.data(), nodes.size() * sizeof(float4)),
               o_refs = put(builder.leaf_refs().data(), builder.leaf_refs().size() * sizeof(uint32_t)),
               o_matA = put(matA.data(), matA.size() * sizeof(float4)), o_matB = put(matB.data(), matB.size() * sizeof(float2)),
               o_ctr = put(zero_counters, sizeof zero_counters);
  host.resize((host.size() + 255) & ~size_t(255));
  hf->o_sA = o_sA; hf->o_sB = o_sB; hf->o_sId = o_sId; hf->o_big = o_big; hf->o_tri = o_tri; hf->o_triId = o_triId; hf->o_nodes = o_nodes;
  hf->o_refs = o_refs; hf->o_matA = o_matA; hf->o_matB = o_matB; hf->o_ctr = o_ctr;
  hf->n_static = static_cast<int32_t>(sA_static.size()); hf->n_moving = static_cast<int32_t>(sA_moving.size());
  hf->n_big = static_cast<int32_t>(big.size()); hf->n_tri = static_cast<int32_t>(triId.size());
  hf->n_nodes = static_cast<int32_t>(builder.nodes().size()); hf->leaf_direct = builder.kMaxLeaf == 1 ? 1 : 0;
  hf->n_leaf_refs = builder.leaf_refs().size();
  return 0;
}

int flatten_and_upload(const rtw_scene_desc* desc, int device, rtw_scene* sc, DevBuf<unsigned char>* borrowed = nullptr) {
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess) return fail_cuda("cudaGetDeviceCount (no CUDA device: this library has no CPU fallback)", e);
  if (device < 0 || device >= ndev) return fail("rtw_scene_upload: device ordinal out of range");
  RTW_CUDA(cudaSetDevice(device));
  int cc_major = 0, sm_count = 0;
  RTW_CUDA(cudaDeviceGetAttribute(&cc_major, cudaDevAttrComputeCapabilityMajor, device));
  RTW_CUDA(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, device));
  if (cc_major < 10) return fail("rtw_b200 kernels are built for sm_100a only; this device has an older compute capability");
  sc->device = device;
  sc->sm_count = sm_count;

  sc->nprims = desc ? desc->nprims : 0;
  HostFlat hf;
  if (int rc = flatten_host(desc, &hf)) return rc;
  const std::vector<unsigned char>& host = hf.host;
  const size_t o_sA = hf.o_sA, o_sB = hf.o_sB, o_sId = hf.o_sId, o_big = hf.o_big, o_tri = hf.o_tri, o_triId = hf.o_triId, o_nodes = hf.o_nodes,
               o_refs = hf.o_refs, o_matA = hf.o_matA, o_matB = hf.o_matB, o_ctr = hf.o_ctr;
  if (borrowed) {  // grow-only buffer owned by the caller (rtw_render's cache): no allocation in the steady state
    if (borrowed->n < host.size()) RTW_CUDA(borrowed->alloc(host.size() + host.size() / 8));
    sc->arena_ptr = borrowed->p;
  } else {
    RTW_CUDA(sc->arena.alloc(host.size()));
    sc->arena_ptr = sc->arena.p;
  }
  RTW_CUDA(cudaMemcpy(sc->arena_ptr, host.data(), host.size(), cudaMemcpyHostToDevice));
  unsigned char* base = sc->arena_ptr;
  sc->counters = reinterpret_cast<unsigned long long*>(base + o_ctr);
  sc->n_leaf_refs = hf.n_leaf_refs;

  rtw::DevScene& d = sc->dev;
  d.sphA = reinterpret_cast<const float4*>(base + o_sA); d.sphB = reinterpret_cast<const float4*>(base + o_sB);
  d.sphId = reinterpret_cast<const int2*>(base + o_sId);
  d.n_static = hf.n_static; d.n_moving = hf.n_moving;
  d.big = reinterpret_cast<const rtw::BigSphere*>(base + o_big); d.n_big = hf.n_big;
  d.tri = reinterpret_cast<const float4*>(base + o_tri); d.triId = reinterpret_cast<const int2*>(base + o_triId);
  d.n_tri = hf.n_tri;
  d.nodes = reinterpret_cast<const float4*>(base + o_nodes); d.leafRefs = reinterpret_cast<const uint32_t*>(base + o_refs);
  d.n_nodes = hf.n_nodes;
  d.leaf_direct = hf.leaf_direct;
  d.matA = reinterpret_cast<const float4*>(base + o_matA); d.matB = reinterpret_cast<const float2*>(base + o_matB);
  const rtw_camera& c = desc->camera;
  for (int k = 0; k < 3; ++k) {
    d.cam.origin[k] = (float)c.origin[k]; d.cam.lower_left[k] = (float)c.lower_left[k];
    d.cam.horizontal[k] = (float)c.horizontal[k]; d.cam.vertical[k] = (float)c.vertical[k];
    d.cam.u[k] = (float)c.u[k]; d.cam.v[k] = (float)c.v[k];
  }
  d.cam.lens_radius = (float)c.lens_radius; d.cam.t0 = (float)c.t0; d.cam.t1 = (float)c.t1;
  sc->has_triangles = hf.n_tri > 0;
  sc->smem_bytes = 16 + (static_cast<size_t>(hf.n_static + hf.n_moving) + 1) * 32;
  return 0;
}

// Which kernel a scene gets.  The shared-memory sphere sweep (K1) needs a sphere-only scene whose tables fit in shared
// memory; it is the faster kernel only for small tables (its cost is linear in the sphere count, the BVH's is
// logarithmic: measured crossover well below the cover scene's 484 spheres), so AUTO picks it up to kSweepAutoMax.
constexpr size_t kSweepAutoMax = 64;
int choose_mode(const rtw_scene* sc, int requested, int* mode) {
  const bool smem_ok = !sc->has_triangles && sc->smem_bytes <= 100 * 1024;
  const size_t nspheres = static_cast<size_t>(sc->dev.n_static + sc->dev.n_moving);
  if (requested == RTW_KERNEL_SPHERES_SMEM) {
    if (!smem_ok) return fail("RTW_KERNEL_SPHERES_SMEM needs a sphere-only scene whose tables fit in shared memory");
    *mode = 0;
  } else if (requested == RTW_KERNEL_BVH) {
    *mode = 1;
  } else if (requested == RTW_KERNEL_AUTO) {
    *mode = (smem_ok && nspheres <= kSweepAutoMax) ? 0 : 1;
  } else {
    return fail("unknown kernel selector");
  }
  return 0;
}

int fill_params(const rtw_scene* sc, const rtw_render_cfg* cfg, int mode, unsigned long long* accum, rtw::RenderParams* p) {
  if (cfg->width < 2 || cfg->height < 2) return fail("render: width and height must be >= 2 (pixel mapping divides by W-1, H-1)");
  if (cfg->sample_begin < 0 || cfg->sample_end <= cfg->sample_begin) return fail("render: empty sample range");
  if (cfg->max_child_rays < 0) return fail("render: max_child_rays must be >= 0");
  if (static_cast<long long>(cfg->width) * cfg->height >= (1ll << 31)) return fail("render: image too large");
  p->sc = sc->dev;
  p->accum = accum;
  p->counters = sc->counters;
  p->width = static_cast<uint32_t>(cfg->width); p->height = static_cast<uint32_t>(cfg->height);
  p->npix = p->width * p->height;
  p->s_begin = static_cast<uint32_t>(cfg->sample_begin); p->s_end = static_cast<uint32_t>(cfg->sample_end);
  const uint32_t S = p->s_end - p->s_begin;
  const unsigned long long n_groups = (p->npix + rtw::kGroupPixels - 1) / rtw::kGroupPixels;
  // enough units to keep every resident warp busy and the tail short: aim for >= 16 units per warp
  const unsigned long long warps = static_cast<unsigned long long>(sc->sm_count) * 4ull * (rtw::kRenderThreads / 32);
  unsigned long long su = (static_cast<unsigned long long>(S) * n_groups) / (16ull * warps);
  su = std::min<unsigned long long>(std::max<unsigned long long>(su, 1ull), 8ull);
  su = std::min<unsigned long long>(su, S);
  p->su = static_cast<uint32_t>(su);
  p->n_chunks = (S + p->su - 1) / p->su;
  p->n_units = n_groups * p->n_chunks;
  p->max_depth = cfg->max_child_rays;
  p->inv_wm1 = 1.0f / static_cast<float>(cfg->width - 1);
  p->inv_hm1 = 1.0f / static_cast<float>(cfg->height - 1);
  p->seed = cfg->seed;
  p->n_leaf_refs = static_cast<uint32_t>(sc->n_leaf_refs);
  (void)mode;
  return 0;
}

void read_counters(const unsigned long long* h, rtw_stats* st) {
  st->rays = h[rtw::kCtrRays]; st->paths = h[rtw::kCtrPaths];
  st->sphere_tests = h[rtw::kCtrSphereTests]; st->sphere_candidates = h[rtw::kCtrCandidates];
  st->node_visits = h[rtw::kCtrNodes]; st->tri_tests = h[rtw::kCtrTriTests];
}

}  // namespace

extern "C" {

int rtw_abi_version(void) { return RTW_ABI_VERSION; }
const char* rtw_last_error(void) { return g_last_error.c_str(); }

int rtw_device_count(int* count) {
  if (!count) return fail("rtw_device_count: null argument");
  cudaError_t e = cudaGetDeviceCount(count);
  if (e != cudaSuccess) { *count = 0; return fail_cuda("cudaGetDeviceCount", e); }
  return 0;
}

int rtw_flatten_info(const rtw_scene_desc* desc, rtw_flatten_report* out) {
  if (!out) return fail("rtw_flatten_info: null output");
  HostFlat hf;
  const double t0 = now_ms();
  if (int rc = flatten_host(desc, &hf)) return rc;
  out->n_static_spheres = hf.n_static; out->n_moving_spheres = hf.n_moving; out->n_big_spheres = hf.n_big; out->n_triangles = hf.n_tri;
  out->n_bvh_nodes = hf.n_nodes; out->leaf_direct = hf.leaf_direct; out->arena_bytes = static_cast<int64_t>(hf.host.size());
  out->flatten_ms = now_ms() - t0; out->bvh_build_ms = hf.bvh_ms;
  // structural self-check of the tree: every primitive referenced exactly once, child boxes inside the parent box
  const rtw::PackedNode* nodes = reinterpret_cast<const rtw::PackedNode*>(hf.host.data() + hf.o_nodes);
  const uint32_t* refs = reinterpret_cast<const uint32_t*>(hf.host.data() + hf.o_refs);
  std::vector<uint8_t> seen(static_cast<size_t>(hf.n_static + hf.n_moving) + static_cast<size_t>(hf.n_tri), 0);
  int64_t dup = 0, depth_max = 0;
  std::vector<std::pair<int32_t, int>> todo;
  if (hf.n_nodes > 0) todo.push_back({0, 1});
  auto visit_ref = [&](uint32_t ref) {
    const size_t idx = (ref >> 30) ? static_cast<size_t>(hf.n_static + hf.n_moving) + (ref & 0x1fffffffu) : (ref & 0x1fffffffu);
    if (idx >= seen.size() || seen[idx]++) ++dup;
  };
  while (!todo.empty()) {
    auto [n, dpt] = todo.back(); todo.pop_back();
    depth_max = std::max<int64_t>(depth_max, dpt);
    for (int which = 0; which < 2; ++which) {
      const int32_t code = which == 0 ? nodes[n].left : nodes[n].right;
      if (code >= 0) { todo.push_back({code, dpt + 1}); continue; }
      if (which == 1 && seen.size() <= static_cast<size_t>(hf.leaf_direct ? 1 : 31) && hf.n_nodes == 1 && nodes[n].rmin_z > nodes[n].rmax[2])
        continue;  // the never-entered filler child of a single-leaf tree (inverted box)
      const uint32_t v = static_cast<uint32_t>(~code);
      if (hf.leaf_direct) visit_ref(v);
      else for (uint32_t k = 0; k < (v & 31u); ++k) visit_ref(refs[(v >> 5) + k]);
    }
  }
  int64_t missing = 0;
  for (uint8_t c : seen) if (!c) ++missing;
  out->bvh_max_depth = static_cast<int32_t>(depth_max);
  out->bvh_errors = dup + missing;
  return 0;
}

int rtw_scene_upload(const rtw_scene_desc* desc, int32_t device, rtw_scene** out) {
  if (!out) return fail("rtw_scene_upload: null output");
  *out = nullptr;
  rtw_scene* sc = new rtw_scene();
  const int rc = flatten_and_upload(desc, device, sc);
  if (rc != 0) { delete sc; return rc; }
  *out = sc;
  return 0;
}

void rtw_scene_free(rtw_scene* scene) {
  if (!scene) return;
  cudaSetDevice(scene->device);
  delete scene;
}

int rtw_render_device(const rtw_scene* scene, const rtw_render_cfg* cfg, int64_t* accum_fx, void* cuda_stream, rtw_stats* stats) {
  if (!scene || !cfg || !accum_fx) return fail("rtw_render_device: null argument");
  RTW_CUDA(cudaSetDevice(scene->device));
  cudaStream_t stream = static_cast<cudaStream_t>(cuda_stream);
  int mode = 0;
  if (int rc = choose_mode(scene, cfg->kernel, &mode)) return rc;
  rtw::RenderParams p{};
  if (int rc = fill_params(scene, cfg, mode, reinterpret_cast<unsigned long long*>(accum_fx), &p)) return rc;
  const int rpl = cfg->rays_per_lane;
  const bool want_stats = (cfg->flags & RTW_FLAG_STATS) != 0;
  RTW_CUDA(cudaMemsetAsync(scene->counters, 0, rtw::kCtrCount * sizeof(unsigned long long), stream));
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (stats) {
    RTW_CUDA(cudaEventCreate(&e0));
    RTW_CUDA(cudaEventCreate(&e1));
    RTW_CUDA(cudaEventRecord(e0, stream));
  }
  cudaError_t le = rtw::launch_render(p, mode, rpl, want_stats, scene->sm_count, stream);
  if (le != cudaSuccess) return fail_cuda("launch k_render", le);
  if (stats) {
    RTW_CUDA(cudaEventRecord(e1, stream));
    unsigned long long h[rtw::kCtrCount];
    RTW_CUDA(cudaMemcpyAsync(h, scene->counters, sizeof h, cudaMemcpyDeviceToHost, stream));
    RTW_CUDA(cudaStreamSynchronize(stream));
    float ms = 0.f;
    RTW_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    std::memset(stats, 0, sizeof *stats);
    read_counters(h, stats);
    stats->kernel_ms = ms;
    stats->kernel_used = mode == 0 ? RTW_KERNEL_SPHERES_SMEM : RTW_KERNEL_BVH;
    stats->launches = 1;
  }
  return 0;
}

int rtw_accum_to_float(const int64_t* accum_fx, float* accum_rgba, int64_t npixels, int32_t device, void* cuda_stream) {
  if (!accum_fx || !accum_rgba || npixels <= 0) return fail("rtw_accum_to_float: invalid argument");
  RTW_CUDA(cudaSetDevice(device));
  cudaError_t e = rtw::launch_accum_to_float(reinterpret_cast<const long long*>(accum_fx), accum_rgba, npixels,
                                             static_cast<cudaStream_t>(cuda_stream));
  if (e != cudaSuccess) return fail_cuda("launch k_accum_to_float", e);
  return 0;
}

// Per-device accumulation buffers of the host-buffer entry point, kept across calls (grow-only) so that a render costs one
// small arena allocation, one H2D copy, the kernels and one D2H copy.  Released by rtw_release_cached_buffers().
namespace {
struct RenderCache {
  DevBuf<long long> fx;
  DevBuf<float> out;
  DevBuf<unsigned char> arena;
  size_t npix = 0;
};
std::mutex g_cache_mutex;
RenderCache* g_cache[64] = {};
}  // namespace

void rtw_release_cached_buffers(void) {
  std::lock_guard<std::mutex> lock(g_cache_mutex);
  for (int d = 0; d < 64; ++d) {
    if (!g_cache[d]) continue;
    cudaSetDevice(d);
    delete g_cache[d];
    g_cache[d] = nullptr;
  }
}

int rtw_render(const rtw_scene_desc* desc, const rtw_render_cfg* cfg, float* accum_rgba, rtw_stats* stats) {
  if (!desc || !cfg || !accum_rgba) return fail("rtw_render: null argument");
  if (cfg->width < 2 || cfg->height < 2) return fail("render: width and height must be >= 2 (pixel mapping divides by W-1, H-1)");
  const double t_start = now_ms();
  std::lock_guard<std::mutex> lock(g_cache_mutex);  // also serialises host-buffer renders per process (not re-entrant per device)
  if (cfg->device < 0 || cfg->device >= 64) return fail("rtw_render: device ordinal out of range");
  if (!g_cache[cfg->device]) g_cache[cfg->device] = new RenderCache();
  RenderCache& rc_ = *g_cache[cfg->device];
  rtw_scene scene_obj;
  rtw_scene* sc = &scene_obj;
  if (int rc = flatten_and_upload(desc, cfg->device, sc, &rc_.arena)) return rc;
  const double t_up = now_ms();
  const size_t npix = static_cast<size_t>(cfg->width) * static_cast<size_t>(cfg->height);
  if (rc_.npix < npix) {
    RTW_CUDA(rc_.fx.alloc(npix * 4));
    RTW_CUDA(rc_.out.alloc(npix * 4));
    rc_.npix = npix;
  }
  RTW_CUDA(cudaMemsetAsync(rc_.fx.p, 0, npix * 4 * sizeof(long long)));
  rtw_stats st{};
  if (int rc = rtw_render_device(sc, cfg, reinterpret_cast<int64_t*>(rc_.fx.p), nullptr, &st)) return rc;
  if (int rc = rtw_accum_to_float(reinterpret_cast<const int64_t*>(rc_.fx.p), rc_.out.p, static_cast<int64_t>(npix), cfg->device, nullptr)) return rc;
  const double t_d0 = now_ms();
  RTW_CUDA(cudaMemcpy(accum_rgba, rc_.out.p, npix * 4 * sizeof(float), cudaMemcpyDeviceToHost));
  const double t_end = now_ms();
  if (stats) {
    *stats = st;
    stats->h2d_ms = t_up - t_start;
    stats->d2h_ms = t_end - t_d0;
    stats->total_ms = t_end - t_start;
    stats->launches = 2;
  }
  return 0;
}

int rtw_finalize_rgb8(const float* accum_rgba, int64_t npixels, int32_t spp, int32_t device, uint8_t* rgb8) {
  if (!accum_rgba || !rgb8 || npixels <= 0 || spp <= 0) return fail("rtw_finalize_rgb8: invalid argument");
  RTW_CUDA(cudaSetDevice(device));
  DevBuf<float> acc;
  DevBuf<uint8_t> out;
  RTW_CUDA(acc.alloc(static_cast<size_t>(npixels) * 4));
  RTW_CUDA(out.alloc(static_cast<size_t>(npixels) * 3));
  RTW_CUDA(cudaMemcpy(acc.p, accum_rgba, static_cast<size_t>(npixels) * 4 * sizeof(float), cudaMemcpyHostToDevice));
  cudaError_t e = rtw::launch_finalize_rgb8(acc.p, out.p, npixels, spp, nullptr);
  if (e != cudaSuccess) return fail_cuda("launch k_finalize_rgb8", e);
  RTW_CUDA(cudaMemcpy(rgb8, out.p, static_cast<size_t>(npixels) * 3, cudaMemcpyDeviceToHost));
  return 0;
}

int rtw_primary_hits(const rtw_scene_desc* desc, int32_t width, int32_t height, double time, int32_t precision, int32_t kernel,
                     int32_t device, int32_t* prim_id, double* t, double* normal, uint8_t* front) {
  if (!desc || !prim_id || !t || !normal || !front) return fail("rtw_primary_hits: null argument");
  if (width < 2 || height < 2) return fail("rtw_primary_hits: width and height must be >= 2");
  if (precision != 32 && precision != 64) return fail("rtw_primary_hits: precision must be 32 or 64");
  const size_t npix = static_cast<size_t>(width) * static_cast<size_t>(height);
  rtw_scene* sc = nullptr;
  // upload also validates the description and the device
  rtw_scene_desc d0 = *desc;
  d0.camera.lens_radius = 0.0; d0.camera.t0 = time; d0.camera.t1 = time;
  if (int rc = rtw_scene_upload(&d0, device, &sc)) return rc;
  struct Guard { rtw_scene* s; ~Guard() { rtw_scene_free(s); } } guard{sc};
  DevBuf<int32_t> d_id; DevBuf<double> d_t, d_n; DevBuf<uint8_t> d_f;
  RTW_CUDA(d_id.alloc(npix)); RTW_CUDA(d_t.alloc(npix)); RTW_CUDA(d_n.alloc(npix * 3)); RTW_CUDA(d_f.alloc(npix));
  if (precision == 64) {
    DevBuf<rtw_primitive> d_prims;
    RTW_CUDA(d_prims.alloc(static_cast<size_t>(desc->nprims)));
    if (desc->nprims > 0)
      RTW_CUDA(cudaMemcpy(d_prims.p, desc->prims, static_cast<size_t>(desc->nprims) * sizeof(rtw_primitive), cudaMemcpyHostToDevice));
    cudaError_t e = rtw::launch_primary_f64(d_prims.p, static_cast<int>(desc->nprims), d0.camera, static_cast<uint32_t>(width),
                                            static_cast<uint32_t>(height), time, d_id.p, d_t.p, d_n.p, d_f.p, nullptr);
    if (e != cudaSuccess) return fail_cuda("launch k_primary_f64", e);
    RTW_CUDA(cudaDeviceSynchronize());
  } else {
    int mode = 0;
    if (int rc = choose_mode(sc, kernel, &mode)) return rc;
    rtw::PrimaryParams p{};
    p.sc = sc->dev; p.width = static_cast<uint32_t>(width); p.height = static_cast<uint32_t>(height);
    p.npix = static_cast<uint32_t>(npix); p.time = static_cast<float>(time);
    p.prim_id = d_id.p; p.t = d_t.p; p.normal = d_n.p; p.front = d_f.p;
    cudaError_t e = rtw::launch_primary_f32(p, mode, nullptr);
    if (e != cudaSuccess) return fail_cuda("launch k_primary_f32", e);
    RTW_CUDA(cudaDeviceSynchronize());
  }
  RTW_CUDA(cudaMemcpy(prim_id, d_id.p, npix * sizeof(int32_t), cudaMemcpyDeviceToHost));
  RTW_CUDA(cudaMemcpy(t, d_t.p, npix * sizeof(double), cudaMemcpyDeviceToHost));
  RTW_CUDA(cudaMemcpy(normal, d_n.p, npix * 3 * sizeof(double), cudaMemcpyDeviceToHost));
  RTW_CUDA(cudaMemcpy(front, d_f.p, npix, cudaMemcpyDeviceToHost));
  return 0;
}

int rtw_debug_scatter(int32_t device, int64_t n, const rtw_material* mats, const float* dir_in, const float* normal,
                      const uint8_t* front, const float* ball, const float* coin, float* out_dir, float* out_att,
                      uint8_t* scattered) {
  if (n <= 0 || !mats || !dir_in || !normal || !front || !ball || !coin || !out_dir || !out_att || !scattered)
    return fail("rtw_debug_scatter: invalid argument");
  RTW_CUDA(cudaSetDevice(device));
  const size_t N = static_cast<size_t>(n);
  std::vector<int> kind(N); std::vector<float> fuzz(N), ior(N), alb(3 * N);
  for (size_t i = 0; i < N; ++i) {
    kind[i] = mats[i].kind; fuzz[i] = (float)std::min(std::max(mats[i].fuzz, 0.0), 1.0); ior[i] = (float)mats[i].ior;
    for (int k = 0; k < 3; ++k) alb[3 * i + k] = (float)mats[i].albedo[k];
  }
  DevBuf<int> d_kind; DevBuf<float> d_fuzz, d_ior, d_alb, d_din, d_n, d_ball, d_coin, d_out, d_att; DevBuf<uint8_t> d_front, d_sc;
  RTW_CUDA(d_kind.upload(kind)); RTW_CUDA(d_fuzz.upload(fuzz)); RTW_CUDA(d_ior.upload(ior)); RTW_CUDA(d_alb.upload(alb));
  RTW_CUDA(d_din.upload(std::vector<float>(dir_in, dir_in + 3 * N))); RTW_CUDA(d_n.upload(std::vector<float>(normal, normal + 3 * N)));
  RTW_CUDA(d_ball.upload(std::vector<float>(ball, ball + 3 * N))); RTW_CUDA(d_coin.upload(std::vector<float>(coin, coin + N)));
  RTW_CUDA(d_front.upload(std::vector<uint8_t>(front, front + N)));
  RTW_CUDA(d_out.alloc(3 * N)); RTW_CUDA(d_att.alloc(3 * N)); RTW_CUDA(d_sc.alloc(N));
  cudaError_t e = rtw::launch_debug_scatter(n, d_kind.p, d_fuzz.p, d_ior.p, d_din.p, d_n.p, d_front.p, d_ball.p, d_coin.p, d_out.p, d_att.p,
                                            d_alb.p, d_sc.p, nullptr);
  if (e != cudaSuccess) return fail_cuda("launch k_debug_scatter", e);
  RTW_CUDA(cudaMemcpy(out_dir, d_out.p, 3 * N * sizeof(float), cudaMemcpyDeviceToHost));
  RTW_CUDA(cudaMemcpy(out_att, d_att.p, 3 * N * sizeof(float), cudaMemcpyDeviceToHost));
  RTW_CUDA(cudaMemcpy(scattered, d_sc.p, N, cudaMemcpyDeviceToHost));
  return 0;
}

int rtw_debug_samples(int32_t device, int64_t n, uint64_t seed, float* ball, float* disk, float* u01) {
  if (n <= 0 || !ball || !disk || !u01) return fail("rtw_debug_samples: invalid argument");
  RTW_CUDA(cudaSetDevice(device));
  const size_t N = static_cast<size_t>(n);
  DevBuf<float> d_ball, d_disk, d_u;
  RTW_CUDA(d_ball.alloc(3 * N)); RTW_CUDA(d_disk.alloc(2 * N)); RTW_CUDA(d_u.alloc(4 * N));
  cudaError_t e = rtw::launch_debug_samples(n, seed, d_ball.p, d_disk.p, d_u.p, nullptr);
  if (e != cudaSuccess) return fail_cuda("launch k_debug_samples", e);
  RTW_CUDA(cudaMemcpy(ball, d_ball.p, 3 * N * sizeof(float), cudaMemcpyDeviceToHost));
  RTW_CUDA(cudaMemcpy(disk, d_disk.p, 2 * N * sizeof(float), cudaMemcpyDeviceToHost));
  RTW_CUDA(cudaMemcpy(u01, d_u.p, 4 * N * sizeof(float), cudaMemcpyDeviceToHost));
  return 0;
}

int rtw_fp32_peak(int32_t device, double seconds, double* tflops, double* sm_mhz) {
  if (!tflops) return fail("rtw_fp32_peak: null argument");
  RTW_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop{};
  RTW_CUDA(cudaGetDeviceProperties(&prop, device));
  const int blocks = prop.multiProcessorCount * 8;
  DevBuf<float> out;
  RTW_CUDA(out.alloc(static_cast<size_t>(blocks) * 256));
  cudaEvent_t e0, e1;
  RTW_CUDA(cudaEventCreate(&e0)); RTW_CUDA(cudaEventCreate(&e1));
  const int iters = 4096;
  const double flop_per_launch = static_cast<double>(blocks) * 256.0 * iters * 16.0 * 8.0 * 2.0;
  // warm up, then repeat launches for ~`seconds` and keep the best and the mean rate
  for (int w = 0; w < 3; ++w) { cudaError_t e = rtw::launch_ffma_peak(out.p, blocks, iters, nullptr); if (e != cudaSuccess) return fail_cuda("launch k_ffma_peak", e); }
  RTW_CUDA(cudaDeviceSynchronize());
  double total_ms = 0.0; int launches = 0;
  const double t_begin = now_ms();
  do {
    RTW_CUDA(cudaEventRecord(e0));
    for (int k = 0; k < 8; ++k) { cudaError_t e = rtw::launch_ffma_peak(out.p, blocks, iters, nullptr); if (e != cudaSuccess) return fail_cuda("launch k_ffma_peak", e); }
    RTW_CUDA(cudaEventRecord(e1));
    RTW_CUDA(cudaEventSynchronize(e1));
    float ms = 0.f; RTW_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    total_ms += ms; launches += 8;
  } while (now_ms() - t_begin < seconds * 1000.0);
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  *tflops = flop_per_launch * launches / (total_ms * 1e-3) / 1e12;
  if (sm_mhz) {
    // clock implied by the measured rate if every SM issued 128 FMA lanes per cycle
    *sm_mhz = (*tflops * 1e12) / (2.0 * 128.0 * prop.multiProcessorCount) / 1e6;
  }
  return 0;
}

}  // extern "C"

// error hook for the other translation units of the library
extern "C" int rtw_set_error_(const char* msg) { return fail(msg ? msg : "unknown error"); }
