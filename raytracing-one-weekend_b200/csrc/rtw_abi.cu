// rtw_abi.cu -- host side of the C ABI declared in include/rtw_b200.h: flatten the caller's scene description
// into the device SoA layout, build the BVH, manage device memory, launch the kernels of rtw_kernels.cu.
// There is deliberately no CPU implementation behind any entry point: without a CUDA device every call fails.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <array>
#include <cstring>
#include <future>
#include <memory>
#include <mutex>
#include <thread>
#include <string>
#include <vector>

#include "rtw_bvh.h"
#include "rtw_host.h"

namespace {
thread_local std::string g_last_error;

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

namespace rtw {
int fail(const std::string& msg) {
  g_last_error = msg;
  return 1;
}
int fail_cuda(const char* what, cudaError_t e) {
  g_last_error = std::string(what) + ": " + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ")";
  return 2;
}
}  // namespace rtw
using rtw::fail;
using rtw::fail_cuda;
using rtw::DevBuf;
using rtw::HostFlat;

namespace {

// Flatten (north_star item 1): variant/virtual primitive list -> SoA tables, BVH, materials, all in one host arena whose
// layout is the device layout (HostFlat, rtw_host.h).  Pure host code: no CUDA call in here (rtw_flatten_info exposes it to
// CPU-only tests).  Two passes over the primitive list, both split over host threads for large scenes: (1) classify + count +
// scene bound, (2) write every table entry and BVH build record straight into its final place (insertion order is kept per table).
enum : uint8_t { kClsStatic = 0, kClsMoving = 1, kClsBig = 2, kClsTri = 3, kClsBadKind = 4, kClsBadMaterial = 5 };

template <typename F>
void parallel_chunks(int64_t n, int nchunks, F&& fn) {  // fn(chunk, begin, end)
  if (nchunks <= 1) { fn(0, int64_t(0), n); return; }
  std::vector<std::future<void>> tasks;
  for (int c = 1; c < nchunks; ++c) tasks.push_back(std::async(std::launch::async, [&fn, c, n, nchunks] { fn(c, n * c / nchunks, n * (c + 1) / nchunks); }));
  fn(0, int64_t(0), n / nchunks);
  for (auto& t : tasks) t.get();
}

}  // namespace

int rtw::flatten_host(const rtw_scene_desc* desc, HostFlat* hf) {
  if (!desc || desc->nprims < 0 || desc->nmats < 0 || (desc->nprims > 0 && !desc->prims) || (desc->nmats > 0 && !desc->mats))
    return fail("rtw_scene_upload: invalid scene description");
  if (desc->nprims >= (1ll << 28)) return fail("rtw_scene_upload: too many primitives");
  const int64_t n = desc->nprims;
  const int hw = static_cast<int>(std::max(1u, std::thread::hardware_concurrency()));
  const int nchunks = n >= 65536 ? std::min(hw, 32) : 1;

  // ---- pass 1: classify, count per chunk, bound of the small primitives (slack of the sweep's reject test) -------------------
  std::vector<uint8_t> cls(static_cast<size_t>(n));
  struct ChunkInfo { int64_t cnt[4] = {0, 0, 0, 0}; double bound = 0.0; int bad = 0; };
  std::vector<ChunkInfo> info(static_cast<size_t>(nchunks));
  parallel_chunks(n, nchunks, [&](int c, int64_t begin, int64_t end) {
    ChunkInfo ci;
    for (int64_t i = begin; i < end; ++i) {
      const rtw_primitive& P = desc->prims[i];
      uint8_t k;
      if (P.material < 0 || P.material >= desc->nmats) { k = kClsBadMaterial; ci.bad = std::max(ci.bad, 2); }
      else if (P.kind == RTW_TRIANGLE) {
        k = kClsTri;
        for (int q = 0; q < 3; ++q) ci.bound = std::max({ci.bound, std::fabs(P.a[q]), std::fabs(P.b[q]), std::fabs(P.c[q])});
      } else if (P.kind == RTW_SPHERE || P.kind == RTW_MOVING_SPHERE) {
        if (std::fabs(P.radius) >= rtw::kBigRadius) k = kClsBig;
        else {
          const bool moving = P.kind == RTW_MOVING_SPHERE && (P.a[0] != P.b[0] || P.a[1] != P.b[1] || P.a[2] != P.b[2]);
          k = moving ? kClsMoving : kClsStatic;
          for (int q = 0; q < 3; ++q) {
            ci.bound = std::max(ci.bound, std::fabs(P.a[q]) + std::fabs(P.radius));
            if (moving) ci.bound = std::max(ci.bound, std::fabs(P.b[q]) + std::fabs(P.radius));
          }
        }
      } else { k = kClsBadKind; ci.bad = std::max(ci.bad, 1); }
      cls[static_cast<size_t>(i)] = k;
      if (k < 4) ++ci.cnt[k];
    }
    info[static_cast<size_t>(c)] = ci;
  });
  double bound = 0.0;
  for (int k = 0; k < 3; ++k) bound = std::max(bound, std::fabs(desc->camera.origin[k]));
  int64_t total[4] = {0, 0, 0, 0};
  std::vector<std::array<int64_t, 4>> start(static_cast<size_t>(nchunks));
  for (int c = 0; c < nchunks; ++c) {
    const ChunkInfo& ci = info[static_cast<size_t>(c)];
    if (ci.bad == 2) return fail("rtw_scene_upload: primitive references a missing material");
    if (ci.bad == 1) return fail("rtw_scene_upload: unknown primitive kind");
    bound = std::max(bound, ci.bound);
    for (int k = 0; k < 4; ++k) { start[static_cast<size_t>(c)][static_cast<size_t>(k)] = total[k]; total[k] += ci.cnt[k]; }
  }
  for (int64_t i = 0; i < desc->nmats; ++i)
    if (desc->mats[i].kind < RTW_LAMBERTIAN || desc->mats[i].kind > RTW_DIELECTRIC) return fail("rtw_scene_upload: unknown material kind");
  const double E = 24.0 * 1.1920929e-7 * bound;
  const size_t n_static = static_cast<size_t>(total[kClsStatic]), n_moving = static_cast<size_t>(total[kClsMoving]),
               n_big = static_cast<size_t>(total[kClsBig]), n_tri = static_cast<size_t>(total[kClsTri]);
  const size_t n_small = n_static + n_moving, n_items = n_small + n_tri;

  rtw::BvhBuilder builder;
  if (const char* e = std::getenv("RTW_SAH_SINGLE_AXIS_BELOW")) builder.single_axis_below = static_cast<size_t>(std::max(0, std::atoi(e)));  // tuning knob
  if (const char* e = std::getenv("RTW_BVH_LEAF")) builder.kMaxLeaf = std::min(std::max(std::atoi(e), 1), 31);  // tuning knob
  // Scenes with triangles can be given a compressed 8-wide BVH (rtw_bvh.h CwBuilder) with leaf-ordered 48-byte primitive records
  // instead of the binary tree with 64-byte nodes: RTW_MESH_BVH=cw8.  Measured on B200: a third of the node fetches,
  // L1 hit rate 46 -> 69 %, but 31 % more warp instructions on the ALU pipe; the binary tree renders 8-23 % faster and is the default.
  const bool cw = n_tri > 0 && rtw::mesh_bvh_is_cw8();
  const bool direct = builder.kMaxLeaf == 1 || cw;
  const size_t node_cap = cw ? 0 : std::max<size_t>(n_items, 1), ref_cap = direct ? 0 : n_items;
  const size_t n_records = cw ? n_items : n_tri;   // cw: one record per leaf primitive (triangles and small spheres)
  // build on the device (rtw_build.cu): binary single-primitive-leaf tree only, at least two primitives
  const bool gpu_build = hf->gpu_build && !cw && builder.kMaxLeaf == 1 && n_items >= 2;
  hf->gpu_build = gpu_build;

  // ---- arena layout ----------------------------------------------------------------------------------------------------------------
  size_t cursor = 0;
  auto reserve = [&cursor](size_t bytes) { const size_t off = (cursor + 255) & ~size_t(255); cursor = off + bytes; return off; };
  hf->o_sA = reserve(n_small * sizeof(float4)); hf->o_sB = reserve(n_small * sizeof(float4)); hf->o_sId = reserve(n_small * sizeof(int2));
  hf->o_big = reserve(n_big * sizeof(rtw::BigSphere));
  hf->o_tri = reserve(n_records * 3 * sizeof(float4)); hf->o_triId = reserve(n_records * sizeof(int2));
  hf->o_nodes = reserve(gpu_build ? 0 : node_cap * sizeof(rtw::PackedNode)); hf->o_refs = reserve(ref_cap * sizeof(uint32_t));
  hf->o_matA = reserve(static_cast<size_t>(desc->nmats) * sizeof(float4)); hf->o_matB = reserve(static_cast<size_t>(desc->nmats) * sizeof(float2));
  hf->o_ctr = reserve(static_cast<size_t>(rtw::kCtrSlots) * rtw::kCtrCount * sizeof(unsigned long long));
  // the wide nodes come last: their number is only known after the collapse (at most one per primitive; typically a sixth), so the
  // arena is sized for the bound and trimmed afterwards (untouched pages of the host allocation are never committed)
  hf->o_cw = reserve(cw ? std::max<size_t>(n_items, 1) * sizeof(rtw::CwNode) : 0);
  hf->upload_bytes = 0;
  if (gpu_build) {   // the device fills the node table: it sits behind everything that is uploaded
    hf->upload_bytes = (cursor + 255) & ~size_t(255);
    hf->o_nodes = reserve(node_cap * sizeof(rtw::PackedNode));
  }
  hf->bytes = (cursor + 255) & ~size_t(255);
  if (!gpu_build) hf->upload_bytes = hf->bytes;
  hf->host.reset(new unsigned char[hf->upload_bytes]);
  unsigned char* base = hf->host.get();
  float4* sA = reinterpret_cast<float4*>(base + hf->o_sA);
  float4* sB = reinterpret_cast<float4*>(base + hf->o_sB);
  int2* sId = reinterpret_cast<int2*>(base + hf->o_sId);
  rtw::BigSphere* big = reinterpret_cast<rtw::BigSphere*>(base + hf->o_big);
  float4* tri = reinterpret_cast<float4*>(base + hf->o_tri);
  int2* triId = reinterpret_cast<int2*>(base + hf->o_triId);
  std::memset(base + hf->o_ctr, 0, static_cast<size_t>(rtw::kCtrSlots) * rtw::kCtrCount * sizeof(unsigned long long));

  // ---- pass 2: table entries and BVH build records (small spheres: swept bounds, common-model.cpp:197-207) ----------------------
  std::vector<rtw::BvhBuilder::Item> items(n_items);
  std::vector<uint32_t> tri_src(cw ? n_tri : 0);   // cw: triangle rank -> primitive index (the records are written in leaf order after the build)
  auto tri_record = [](const rtw_primitive& P, float4& q0, float4& q1, float4& q2) {
    const double e1[3] = {P.b[0] - P.a[0], P.b[1] - P.a[1], P.b[2] - P.a[2]};
    const double e2[3] = {P.c[0] - P.a[0], P.c[1] - P.a[1], P.c[2] - P.a[2]};
    const double nn[3] = {e1[1] * e2[2] - e2[1] * e1[2], e1[2] * e2[0] - e2[2] * e1[0], e1[0] * e2[1] - e2[0] * e1[1]};
    q0 = make_float4((float)P.a[0], (float)P.a[1], (float)P.a[2], (float)nn[0]);
    q1 = make_float4((float)e1[0], (float)e1[1], (float)e1[2], (float)nn[1]);
    q2 = make_float4((float)e2[0], (float)e2[1], (float)e2[2], (float)nn[2]);
  };
  parallel_chunks(n, nchunks, [&](int c, int64_t begin, int64_t end) {
    std::array<int64_t, 4> at = start[static_cast<size_t>(c)];
    for (int64_t i = begin; i < end; ++i) {
      const rtw_primitive& P = desc->prims[i];
      const int id = static_cast<int>(i);
      const uint8_t k = cls[static_cast<size_t>(i)];
      if (k == kClsTri) {
        const size_t t = static_cast<size_t>(at[kClsTri]++);
        float4 q0, q1, q2;
        tri_record(P, q0, q1, q2);
        if (cw) tri_src[t] = static_cast<uint32_t>(i);
        else { tri[3 * t] = q0; tri[3 * t + 1] = q1; tri[3 * t + 2] = q2; triId[t] = make_int2(id, P.material); }
        rtw::BvhBuilder::Item& it = items[n_small + t];
        const float va[3] = {q0.x, q0.y, q0.z}, vb[3] = {q0.x + q1.x, q0.y + q1.y, q0.z + q1.z}, vc[3] = {q0.x + q2.x, q0.y + q2.y, q0.z + q2.z};
        it.box.reset(); it.box.grow(va); it.box.grow(vb); it.box.grow(vc);
        it.ref = (1u << 30) | static_cast<uint32_t>(t);
        continue;
      }
      const bool moving = k == kClsMoving;
      const double dc[3] = {P.kind == RTW_MOVING_SPHERE ? P.b[0] - P.a[0] : 0.0, P.kind == RTW_MOVING_SPHERE ? P.b[1] - P.a[1] : 0.0,
                            P.kind == RTW_MOVING_SPHERE ? P.b[2] - P.a[2] : 0.0};
      if (k == kClsBig) {
        rtw::BigSphere bsp{};
        for (int q = 0; q < 3; ++q) { bsp.c0[q] = P.a[q]; bsp.dc[q] = dc[q]; }
        bsp.r = P.radius; bsp.prim_id = id; bsp.material = P.material;
        big[static_cast<size_t>(at[kClsBig]++)] = bsp;
        continue;
      }
      const size_t t = moving ? n_static + static_cast<size_t>(at[kClsMoving]++) : static_cast<size_t>(at[kClsStatic]++);
      const double r = std::fabs(P.radius);
      const float r2c = static_cast<float>((r * r + 3.0 * r * E + E * E) * (1.0 + 4e-7));
      const float4 A = make_float4((float)P.a[0], (float)P.a[1], (float)P.a[2], r2c);
      const float4 B = make_float4(moving ? (float)dc[0] : 0.0f, moving ? (float)dc[1] : 0.0f, moving ? (float)dc[2] : 0.0f, (float)P.radius);
      sA[t] = A; sB[t] = B; sId[t] = make_int2(id, P.material);
      rtw::BvhBuilder::Item& it = items[t];
      const float rf = std::fabs(B.w);
      const float c0[3] = {A.x, A.y, A.z}, c1[3] = {A.x + B.x, A.y + B.y, A.z + B.z};
      for (int q = 0; q < 3; ++q) { it.box.lo[q] = std::min(c0[q], c1[q]) - rf; it.box.hi[q] = std::max(c0[q], c1[q]) + rf; }
      it.ref = static_cast<uint32_t>(t);
    }
  });

  // ---- BVH ---------------------------------------------------------------------------------------------------------------------------
  const double t_bvh = now_ms();
  rtw::PackedNode* nodes_out = gpu_build ? nullptr : reinterpret_cast<rtw::PackedNode*>(base + hf->o_nodes);
  size_t n_nodes = 0, n_refs = 0;
  size_t n_cw = 0;
  int cw_depth = 0;
  if (gpu_build) {
    auto keep = std::make_shared<std::vector<rtw::BvhBuilder::Item>>();
    keep->swap(items);
    hf->n_gpu_items = keep->size();
    hf->gpu_items = keep;
    n_nodes = n_items - 1;
  } else if (cw) {
    std::vector<rtw::BinNode> bin(n_items >= 2 ? n_items - 1 : 0);
    builder.build_items_binary(items, bin.data());
    rtw::CwBuilder wide;
    if (const char* e = std::getenv("RTW_CW_LEAF")) wide.max_leaf = std::min(std::max(std::atoi(e), 1), 3);  // tuning knob
    wide.margin = 4.0 * 1.1920929e-7 * bound;
    wide.build(bin.data(), bin.size(), n_items == 1 ? &items[0].box : nullptr, n_items == 1 ? items[0].ref : 0u);
    n_cw = wide.nodes().size();
    cw_depth = wide.depth();
    if (n_cw) std::memcpy(base + hf->o_cw, wide.nodes().data(), n_cw * sizeof(rtw::CwNode));
    hf->bytes = (hf->o_cw + std::max<size_t>(n_cw, 1) * sizeof(rtw::CwNode) + 255) & ~size_t(255);
    hf->upload_bytes = hf->bytes;
    // leaf-ordered records: the primitives of one leaf (and of neighbouring leaves) next to each other in memory
    const std::vector<uint32_t>& order = wide.leaf_order();
    const float nan = std::numeric_limits<float>::quiet_NaN();
    parallel_chunks(static_cast<int64_t>(order.size()), nchunks, [&](int, int64_t begin, int64_t end) {
      for (int64_t P = begin; P < end; ++P) {
        const uint32_t ref = order[static_cast<size_t>(P)], t = ref & 0x1fffffffu;
        if (ref >> 30) {
          const uint32_t i = tri_src[t];
          tri_record(desc->prims[i], tri[3 * P], tri[3 * P + 1], tri[3 * P + 2]);
          triId[P] = make_int2(static_cast<int>(i), desc->prims[i].material);
        } else {
          tri[3 * P] = sA[t]; tri[3 * P + 1] = sB[t];
          tri[3 * P + 2] = make_float4(nan, __int_as_float_host(static_cast<int>(t)), 0.0f, 0.0f);
          triId[P] = sId[t];
        }
      }
    });
    if (cw_depth > rtw::kCwStack)
      return fail("rtw_scene_upload: the wide BVH is deeper than the kernels' traversal stack (" + std::to_string(cw_depth) + " > " +
                  std::to_string(rtw::kCwStack) + " levels): degenerate primitive distribution");
  } else if (direct) {
    n_nodes = builder.build_items_direct(items, nodes_out);
  } else {
    std::vector<rtw::Box3> boxes(n_items);
    std::vector<uint32_t> refs(n_items);
    for (size_t i = 0; i < n_items; ++i) { boxes[i] = items[i].box; refs[i] = items[i].ref; }
    builder.build(boxes, refs);
    n_nodes = builder.nodes().size(); n_refs = builder.leaf_refs().size();
    if (n_nodes) std::memcpy(nodes_out, builder.nodes().data(), n_nodes * sizeof(rtw::PackedNode));
    if (n_refs) std::memcpy(base + hf->o_refs, builder.leaf_refs().data(), n_refs * sizeof(uint32_t));
  }
  // the kernels drop a push when their traversal stack is full, which would lose a subtree: refuse such a tree instead (both
  // builders switch to median splits below kSahDepthLimit, so this only triggers on a builder bug)
  if (builder.max_depth() > rtw::kBvhStack)
    return fail("rtw_scene_upload: the BVH is deeper than the kernels' traversal stack (" + std::to_string(builder.max_depth()) + " > " +
                std::to_string(rtw::kBvhStack) + " levels): degenerate primitive distribution");
  hf->bvh_depth = cw ? cw_depth : builder.max_depth();
  hf->bvh_ms = now_ms() - t_bvh;
  hf->n_cw = static_cast<int32_t>(n_cw); hf->n_records = static_cast<int32_t>(n_records); hf->cw_has_spheres = cw && n_small > 0 ? 1 : 0;

  float4* matA = reinterpret_cast<float4*>(base + hf->o_matA);
  float2* matB = reinterpret_cast<float2*>(base + hf->o_matB);
  for (int64_t i = 0; i < desc->nmats; ++i) {
    const rtw_material& m = desc->mats[i];
    const double fuzz = std::min(std::max(m.fuzz, 0.0), 1.0);  // common-model.h:132-133,143-144
    matA[i] = make_float4((float)m.albedo[0], (float)m.albedo[1], (float)m.albedo[2], (float)fuzz);
    matB[i] = make_float2((float)m.ior, __int_as_float_host(m.kind));
  }
  hf->n_static = static_cast<int32_t>(n_static); hf->n_moving = static_cast<int32_t>(n_moving);
  hf->n_big = static_cast<int32_t>(n_big); hf->n_tri = static_cast<int32_t>(n_tri);
  hf->n_nodes = static_cast<int32_t>(n_nodes); hf->leaf_direct = direct ? 1 : 0;
  hf->n_leaf_refs = n_refs;
  return 0;
}

// 64-bit key of a scene description: every byte of the primitive and material arrays and of the camera block.  Word-wise
// multiply-xorshift mixing, four independent lanes per chunk, chunks hashed by separate host threads for big scenes and combined in
// order (1 M primitives = 88 MB: ~3 ms on 16 threads).
namespace {
uint64_t mix64(uint64_t h, uint64_t w) {
  h = (h ^ w) * 0x9E3779B97F4A7C15ull;
  return h ^ (h >> 29);
}
uint64_t hash_bytes(const void* data, size_t bytes, uint64_t seed) {
  const unsigned char* p = static_cast<const unsigned char*>(data);
  uint64_t h0 = seed ^ 0x243F6A8885A308D3ull, h1 = seed ^ 0x13198A2E03707344ull, h2 = seed ^ 0xA4093822299F31D0ull, h3 = seed ^ 0x082EFA98EC4E6C89ull;
  size_t i = 0;
  for (; i + 32 <= bytes; i += 32) {
    uint64_t w[4];
    std::memcpy(w, p + i, 32);
    h0 = mix64(h0, w[0]); h1 = mix64(h1, w[1]); h2 = mix64(h2, w[2]); h3 = mix64(h3, w[3]);
  }
  uint64_t tail[4] = {0, 0, 0, 0};
  std::memcpy(tail, p + i, bytes - i);
  h0 = mix64(h0, tail[0]); h1 = mix64(h1, tail[1]); h2 = mix64(h2, tail[2]); h3 = mix64(h3, tail[3] ^ bytes);
  return mix64(mix64(mix64(h0, h1), h2), h3);
}
}  // namespace

bool rtw::mesh_bvh_is_cw8() {
  const char* e = std::getenv("RTW_MESH_BVH");
  return e && std::string(e) == "cw8";
}

uint64_t rtw::scene_key(const rtw_scene_desc* desc) {
  if (!desc) return 1;
  const size_t pb = desc->nprims > 0 && desc->prims ? static_cast<size_t>(desc->nprims) * sizeof(rtw_primitive) : 0;
  const size_t mb = desc->nmats > 0 && desc->mats ? static_cast<size_t>(desc->nmats) * sizeof(rtw_material) : 0;
  uint64_t h = hash_bytes(&desc->camera, sizeof desc->camera, static_cast<uint64_t>(desc->nprims) * 0x100000001B3ull + static_cast<uint64_t>(desc->nmats));
  h = mix64(h, hash_bytes(desc->mats, mb, 2));
  const int hw = static_cast<int>(std::max(1u, std::thread::hardware_concurrency()));
  const int nchunks = pb >= (size_t(4) << 20) ? std::min(hw, 32) : 1;
  std::vector<uint64_t> part(static_cast<size_t>(nchunks), 0);
  const unsigned char* base = reinterpret_cast<const unsigned char*>(desc->prims);
  parallel_chunks(static_cast<int64_t>(desc->nprims > 0 ? desc->nprims : 0), nchunks, [&](int c, int64_t begin, int64_t end) {
    part[static_cast<size_t>(c)] = hash_bytes(base + static_cast<size_t>(begin) * sizeof(rtw_primitive), static_cast<size_t>(end - begin) * sizeof(rtw_primitive), 3 + static_cast<uint64_t>(c));
  });
  for (uint64_t v : part) h = mix64(h, v);
  if (mesh_bvh_is_cw8()) h = mix64(h, 0xC8);   // the device tables depend on the tree format
  return h ? h : 1;
}

int rtw::upload_flat(const HostFlat& hf, const rtw_camera& c, int64_t nprims, int device, rtw_scene* sc, DevBuf<unsigned char>* borrowed, cudaStream_t stream) {
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
  sc->nprims = nprims;
  if (borrowed) {  // grow-only buffer owned by the caller (a device slot): no allocation in the steady state
    RTW_CUDA(borrowed->reserve(hf.bytes));
    sc->arena_ptr = borrowed->p;
  } else {
    RTW_CUDA(sc->arena.alloc(hf.bytes));
    sc->arena_ptr = sc->arena.p;
  }
  RTW_CUDA(cudaMemcpyAsync(sc->arena_ptr, hf.host.get(), hf.upload_bytes, cudaMemcpyHostToDevice, stream));
  sc->gpu_build_ms = 0.0;
  if (hf.gpu_build) {
    int depth = 0;
    const auto* items = static_cast<const std::vector<rtw::BvhBuilder::Item>*>(hf.gpu_items.get());
    const bool sah_top = !(std::getenv("RTW_LBVH_SAH_TOP") && std::atoi(std::getenv("RTW_LBVH_SAH_TOP")) == 0);   // A/B knob
    const bool sah_clusters = !(std::getenv("RTW_LBVH_SAH_CLUSTERS") && std::atoi(std::getenv("RTW_LBVH_SAH_CLUSTERS")) == 0);   // A/B knob
    int top_nodes = 0, clusters_rebuilt = 0;
    if (int rc = gpu_build_bvh(items->data(), items->size(), sc->arena_ptr + hf.o_nodes, stream, sah_top, sah_clusters, &depth, &sc->gpu_build_ms, &top_nodes, &clusters_rebuilt)) return rc;
    if (std::getenv("RTW_TRACE"))
      std::fprintf(stderr, "rtw trace: device BVH build %.2f ms, %d top nodes rebuilt with SAH on the host, %d subtrees below them with SAH on the device, depth bound %d\n",
                   sc->gpu_build_ms, top_nodes, clusters_rebuilt, depth);
    if (depth > rtw::kBvhStack)
      return fail("rtw_scene_upload: the device-built BVH is deeper than the kernels' traversal stack (" + std::to_string(depth) + " > " +
                  std::to_string(rtw::kBvhStack) + " levels): use the host builder (RTW_FLAG_BVH_BUILD_HOST)");
    sc->bvh_depth = depth;
  } else {
    sc->bvh_depth = hf.bvh_depth;
  }
  RTW_CUDA(cudaStreamSynchronize(stream));  // the host arena may go away when the caller returns
  unsigned char* base = sc->arena_ptr;
  sc->counters = reinterpret_cast<unsigned long long*>(base + hf.o_ctr);
  sc->n_leaf_refs = hf.n_leaf_refs;
  sc->arena_bytes = hf.bytes;

  rtw::DevScene& d = sc->dev;
  d.sphA = reinterpret_cast<const float4*>(base + hf.o_sA); d.sphB = reinterpret_cast<const float4*>(base + hf.o_sB);
  d.sphId = reinterpret_cast<const int2*>(base + hf.o_sId);
  d.n_static = hf.n_static; d.n_moving = hf.n_moving;
  d.big = reinterpret_cast<const rtw::BigSphere*>(base + hf.o_big); d.n_big = hf.n_big;
  d.tri = reinterpret_cast<const float4*>(base + hf.o_tri); d.triId = reinterpret_cast<const int2*>(base + hf.o_triId);
  d.n_tri = hf.n_cw > 0 ? hf.n_records : hf.n_tri;
  d.cwNodes = reinterpret_cast<const uint4*>(base + hf.o_cw); d.n_cw_nodes = hf.n_cw; d.cw_has_spheres = hf.cw_has_spheres;
  d.nodes = reinterpret_cast<const float4*>(base + hf.o_nodes); d.leafRefs = reinterpret_cast<const uint32_t*>(base + hf.o_refs);
  d.n_nodes = hf.n_nodes;
  d.leaf_direct = hf.leaf_direct;
  d.matA = reinterpret_cast<const float4*>(base + hf.o_matA); d.matB = reinterpret_cast<const float2*>(base + hf.o_matB);
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

// Context creation takes 0.3-1 s per GPU on a cold process: rtw_prewarm starts it on background threads (one per device) so that it
// overlaps the caller's scene construction and, for several GPUs, each other; the host-buffer entry points join them first.
namespace {
std::mutex g_warm_mutex;
std::vector<std::thread> g_warm_threads;
uint64_t g_warm_started = 0;   // bit d: device d has been (or is being) warmed
}  // namespace
void rtw::prewarm_join() {
  std::lock_guard<std::mutex> lock(g_warm_mutex);
  for (auto& t : g_warm_threads) t.join();
  g_warm_threads.clear();
}

// ---- per-device slots -----------------------------------------------------------------------------------------------------------
namespace {
std::mutex g_slots_mutex;
rtw::DeviceSlot* g_slots[64] = {};
}  // namespace

rtw::DeviceSlot* rtw::device_slot(int device) {
  if (device < 0 || device >= 64) { fail("device ordinal out of range"); return nullptr; }
  std::lock_guard<std::mutex> lock(g_slots_mutex);
  if (!g_slots[device]) { g_slots[device] = new DeviceSlot(); g_slots[device]->device = device; }
  return g_slots[device];
}

int rtw::slot_prepare(DeviceSlot* s) {
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess) return fail_cuda("cudaGetDeviceCount (no CUDA device: this library has no CPU fallback)", e);
  if (s->device >= ndev) return fail("render: device ordinal out of range");
  RTW_CUDA(cudaSetDevice(s->device));
  if (!s->stream) RTW_CUDA(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
  if (!s->rendered) RTW_CUDA(cudaEventCreateWithFlags(&s->rendered, cudaEventDisableTiming));
  return 0;
}

// Where the BVH of a host-buffer render is built.  Explicit flags win; otherwise the device builds it for big scenes unless the render
// is very long: the device build (radix tree, SAH top on the host, SAH inside the subtrees on the device) takes ~15 ms for a million
// triangles against ~80 ms for the SAH tree on 16 host cores and traces 3.3 % slower (991k-triangle mesh: 36.8 against 35.6 ms per
// 133 M paths), so it pays while the render itself is shorter than about two seconds.
bool rtw::choose_gpu_build(const rtw_scene_desc* desc, const rtw_render_cfg* cfg) {
  if (cfg->flags & RTW_FLAG_BVH_BUILD_HOST) return false;
  if (cfg->flags & RTW_FLAG_BVH_BUILD_GPU) return true;
  if (mesh_bvh_is_cw8()) return false;
  const double paths = static_cast<double>(cfg->width) * cfg->height * (cfg->sample_end - cfg->sample_begin);
  return desc->nprims >= 200000 && paths < 6.0e9;
}

int rtw::slot_set_scene(DeviceSlot* s, const rtw_scene_desc* desc, uint64_t key, bool use_cache, HostFlat* flat, std::mutex* flat_mutex, bool* hit) {
  if (hit) *hit = false;
  if (use_cache && s->key != 0 && s->key == key) { if (hit) *hit = true; return 0; }
  s->key = 0;
  {
    std::unique_lock<std::mutex> lock;
    if (flat_mutex) lock = std::unique_lock<std::mutex>(*flat_mutex);   // several GPUs of one call share ONE flatten
    if (!flat->host) {
      if (int rc = flatten_host(desc, flat)) return rc;
    }
  }
  if (int rc = upload_flat(*flat, desc->camera, desc->nprims, s->device, &s->scene, &s->arena, s->stream)) return rc;
  s->key = key;
  return 0;
}

namespace {

// Which kernel a scene gets.  AUTO is always a BVH kernel: since the wavefront kernel (K2w) it is the faster one at every table
// size (1080p, 64 spp, Mpaths/s, K1 vs K2w: 8 primitives 11 803 / 15 088, 20: 10 042 / 14 068, 40: 8 077 / 13 398, 66: 6 111 / 11 680,
// 145: 3 870 / 9 504; scripts/kernel_crossover.py).  The shared-memory sphere sweep (K1) stays available on request
// (RTW_KERNEL_SPHERES_SMEM) as the kernel the FP32-FMA roofline of SURVEY 8(d) is defined on; it needs a sphere-only scene
// whose tables fit in shared memory.
int choose_mode(const rtw_scene* sc, int requested, int* mode) {
  const bool smem_ok = !sc->has_triangles && sc->smem_bytes <= 100 * 1024;
  if (requested == RTW_KERNEL_SPHERES_SMEM) {
    if (!smem_ok) return fail("RTW_KERNEL_SPHERES_SMEM needs a sphere-only scene whose tables fit in shared memory");
    *mode = 0;
  } else if (requested == RTW_KERNEL_BVH || requested == RTW_KERNEL_BVH_PERLANE) {
    *mode = 1;
  } else if (requested == RTW_KERNEL_AUTO) {
    *mode = 1;
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
  p->counters = sc->counters;   // rtw_render_device picks the launch's own block
  p->width = static_cast<uint32_t>(cfg->width); p->height = static_cast<uint32_t>(cfg->height);
  p->npix = p->width * p->height;
  p->tile_rows = 1; p->tile_count = 1; p->tile_index = 0;
  if (cfg->row_tile_count > 1) {
    if (cfg->row_tile_rows < 1 || cfg->row_tile_index < 0 || cfg->row_tile_index >= cfg->row_tile_count)
      return fail("render: row-tile split needs row_tile_rows >= 1 and 0 <= row_tile_index < row_tile_count");
    p->tile_rows = static_cast<uint32_t>(cfg->row_tile_rows); p->tile_count = static_cast<uint32_t>(cfg->row_tile_count);
    p->tile_index = static_cast<uint32_t>(cfg->row_tile_index);
    p->npix = p->width * static_cast<uint32_t>(rtw_row_tile_local_rows(cfg->height, cfg->row_tile_rows, cfg->row_tile_count));
  }
  p->s_begin = static_cast<uint32_t>(cfg->sample_begin); p->s_end = static_cast<uint32_t>(cfg->sample_end);
  const uint32_t S = p->s_end - p->s_begin;
  // work groups: 16 x 8 pixel tiles of the (local) image
  p->tiles_x = (p->width + 15u) / 16u;
  const unsigned long long n_groups = static_cast<unsigned long long>(p->tiles_x) * ((p->npix / p->width + 7u) / 8u);
  // enough units to keep every resident warp busy and the tail short: aim for >= 64 units per warp, at most 16 samples per unit
  // (measured on the cover scene at 1080p, kernel ms: 128 spp su 1/2/4/8 = 36.25/35.73/35.58/35.76, 1024 spp su 4/8/16 = 291.2/282.8/281.6)
  const unsigned long long warps = static_cast<unsigned long long>(sc->sm_count) * 4ull * (rtw::kRenderThreads / 32);
  unsigned long long su = (static_cast<unsigned long long>(S) * n_groups) / (64ull * warps);
  su = std::min<unsigned long long>(std::max<unsigned long long>(su, 1ull), 16ull);
  if (const char* e = std::getenv("RTW_SU")) su = std::max(1, std::atoi(e));   // tuning knob (samples per work unit)
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
  if (int rc = rtw::flatten_host(desc, &hf)) return rc;
  out->n_static_spheres = hf.n_static; out->n_moving_spheres = hf.n_moving; out->n_big_spheres = hf.n_big; out->n_triangles = hf.n_tri;
  out->n_bvh_nodes = hf.n_nodes; out->leaf_direct = hf.leaf_direct; out->arena_bytes = static_cast<int64_t>(hf.bytes);
  out->flatten_ms = now_ms() - t0; out->bvh_build_ms = hf.bvh_ms;
  {
    const size_t tables = 16 + (hf.n_cw > 0 ? static_cast<size_t>(hf.n_cw) * 80 : static_cast<size_t>(hf.n_nodes) * 64) + ((hf.n_leaf_refs * 4 + 15) & ~size_t(15)) +
                          static_cast<size_t>(hf.n_static + hf.n_moving) * 32 + static_cast<size_t>(hf.n_cw > 0 ? hf.n_records : hf.n_tri) * 48;
    const rtw::BvhPlan plan = rtw::plan_bvh(tables, hf.n_tri, hf.leaf_direct != 0, false, hf.n_cw > 0);
    out->bvh_variant = plan.variant; out->bvh_warps_per_cta = plan.warps; out->bvh_tables_in_smem = plan.tables_in_smem ? 1 : 0;
    out->reserved2 = 0; out->bvh_smem_bytes = static_cast<int64_t>(plan.smem_bytes);
  }
  if (hf.n_cw > 0) {
    // compressed wide BVH: every leaf record referenced exactly once, and every dequantised child box contains the exact bounds of
    // everything below it (triangle vertices, swept sphere bounds), checked bottom-up
    const rtw::CwNode* cwn = reinterpret_cast<const rtw::CwNode*>(hf.host.get() + hf.o_cw);
    const float4* rec = reinterpret_cast<const float4*>(hf.host.get() + hf.o_tri);
    std::vector<uint8_t> seen(static_cast<size_t>(hf.n_records), 0);
    int64_t errors = 0;
    int depth_max = 0;
    struct Walk {
      const rtw::CwNode* n; const float4* rec; std::vector<uint8_t>& seen; int64_t& errors; int& depth_max; int32_t n_cw;
      rtw::Box3 visit(uint32_t idx, int depth) {
        rtw::Box3 all; all.reset();
        if (idx >= static_cast<uint32_t>(n_cw) || depth > 64) { ++errors; return all; }
        depth_max = std::max(depth_max, depth);
        const rtw::CwNode& nd = n[idx];
        float p[3]; std::memcpy(p, nd.w, 12);
        const uint32_t e = nd.w[3], imask = e >> 24;
        uint8_t meta[8], q[6][8];
        std::memcpy(meta, &nd.w[6], 8);
        for (int a = 0; a < 6; ++a) std::memcpy(q[a], &nd.w[8 + 2 * a], 8);
        uint32_t rank = 0;
        for (int s = 0; s < 8; ++s) {
          if (meta[s] == 0) { if (imask >> s & 1) ++errors; continue; }
          rtw::Box3 below; below.reset();
          const bool inner = (meta[s] & 0x18) == 0x18;
          if (inner != ((imask >> s & 1) != 0)) ++errors;
          if (inner) {
            if ((meta[s] & 0x1f) != 24 + s || (meta[s] >> 5) != 1) ++errors;
            below = visit(nd.w[4] + rank++, depth + 1);
          } else {
            const uint32_t unary = meta[s] >> 5, cnt = unary == 1 ? 1 : (unary == 3 ? 2 : (unary == 7 ? 3 : 0)), first = nd.w[5] + (meta[s] & 0x1f);
            if (cnt == 0) ++errors;
            for (uint32_t k = 0; k < cnt; ++k) {
              const size_t P = first + k;
              if (P >= seen.size() || seen[P]++) { ++errors; continue; }
              const float4 q0 = rec[3 * P], q1 = rec[3 * P + 1], q2 = rec[3 * P + 2];
              if (q2.x != q2.x) {   // sphere record: swept bounds
                const float r = std::fabs(q1.w);
                const float c0[3] = {q0.x, q0.y, q0.z}, c1[3] = {q0.x + q1.x, q0.y + q1.y, q0.z + q1.z};
                for (int a = 0; a < 3; ++a) { below.lo[a] = std::min(below.lo[a], std::min(c0[a], c1[a]) - r); below.hi[a] = std::max(below.hi[a], std::max(c0[a], c1[a]) + r); }
              } else {
                const float va[3] = {q0.x, q0.y, q0.z}, vb[3] = {q0.x + q1.x, q0.y + q1.y, q0.z + q1.z}, vc[3] = {q0.x + q2.x, q0.y + q2.y, q0.z + q2.z};
                below.grow(va); below.grow(vb); below.grow(vc);
              }
            }
          }
          for (int a = 0; a < 3; ++a) {
            const double step = std::ldexp(1.0, static_cast<int>(e >> (8 * a) & 0xff) - 127);
            const double lo = p[a] + q[a][s] * step, hi = p[a] + q[3 + a][s] * step;
            if (!(lo <= below.lo[a] && below.hi[a] <= hi)) ++errors;
          }
          all.grow(below);
        }
        return all;
      }
    } walk{cwn, rec, seen, errors, depth_max, hf.n_cw};
    walk.visit(0, 1);
    for (uint8_t c : seen) if (c != 1) ++errors;
    out->n_bvh_nodes = hf.n_cw;
    out->bvh_max_depth = depth_max;
    out->bvh_errors = errors;
    return 0;
  }
  // structural self-check of the tree: every primitive referenced exactly once, child boxes inside the parent box
  const rtw::PackedNode* nodes = reinterpret_cast<const rtw::PackedNode*>(hf.host.get() + hf.o_nodes);
  const uint32_t* refs = reinterpret_cast<const uint32_t*>(hf.host.get() + hf.o_refs);
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
      if (which == 1 && seen.size() <= static_cast<size_t>(hf.leaf_direct ? 1 : 31) && hf.n_nodes == 1 && nodes[n].c[2][1] >= 1.0e38f)
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
  HostFlat hf;
  if (int rc = rtw::flatten_host(desc, &hf)) return rc;
  rtw_scene* sc = new rtw_scene();
  const int rc = rtw::upload_flat(hf, desc->camera, desc->nprims, device, sc, nullptr, nullptr);
  if (rc != 0) { delete sc; return rc; }
  *out = sc;
  return 0;
}

int rtw_scene_upload_ex(const rtw_scene_desc* desc, int32_t device, int32_t flags, rtw_scene** out) {
  if (!out) return fail("rtw_scene_upload_ex: null output");
  *out = nullptr;
  HostFlat hf;
  hf.gpu_build = (flags & RTW_FLAG_BVH_BUILD_GPU) != 0;
  if (int rc = rtw::flatten_host(desc, &hf)) return rc;
  rtw_scene* sc = new rtw_scene();
  const int rc = rtw::upload_flat(hf, desc->camera, desc->nprims, device, sc, nullptr, nullptr);
  if (rc != 0) { delete sc; return rc; }
  *out = sc;
  return 0;
}

// Structural check of a device-resident BINARY tree (whichever builder made it): the arena comes back to the host, every primitive
// must be referenced exactly once and every stored child box (centre +- half-extent) must contain the exact bounds of everything
// below it.  out: n_bvh_nodes, bvh_max_depth, bvh_errors, bvh_build_ms (device build time, 0 for a host-built tree).
int rtw_scene_check(const rtw_scene* scene, rtw_flatten_report* out) {
  if (!scene || !out) return fail("rtw_scene_check: null argument");
  std::memset(out, 0, sizeof *out);
  const rtw::DevScene& d = scene->dev;
  if (d.n_cw_nodes > 0) return fail("rtw_scene_check: binary trees only (the compressed wide tree is checked by rtw_flatten_info)");
  RTW_CUDA(cudaSetDevice(scene->device));
  std::vector<unsigned char> host(scene->arena_bytes);
  RTW_CUDA(cudaMemcpy(host.data(), scene->arena_ptr, scene->arena_bytes, cudaMemcpyDeviceToHost));
  auto at = [&](const void* dev_ptr) { return host.data() + (static_cast<const unsigned char*>(dev_ptr) - scene->arena_ptr); };
  const rtw::PackedNode* nodes = reinterpret_cast<const rtw::PackedNode*>(at(d.nodes));
  const float4* sA = reinterpret_cast<const float4*>(at(d.sphA));
  const float4* sB = reinterpret_cast<const float4*>(at(d.sphB));
  const float4* tri = reinterpret_cast<const float4*>(at(d.tri));
  const size_t n_small = static_cast<size_t>(d.n_static + d.n_moving);
  std::vector<uint8_t> seen(n_small + static_cast<size_t>(d.n_tri), 0);
  int64_t errors = 0;
  int depth_max = 0;
  struct Walk {
    const rtw::PackedNode* nodes; const float4 *sA, *sB, *tri; std::vector<uint8_t>& seen; size_t n_small; int64_t& errors; int& depth_max; int n_nodes;
    rtw::Box3 prim_box(uint32_t ref) {
      rtw::Box3 b; b.reset();
      const uint32_t i = ref & 0x1fffffffu;
      const size_t slot = (ref >> 30) ? n_small + i : i;
      if (slot >= seen.size() || seen[slot]++) { ++errors; return b; }
      if (ref >> 30) {
        const float4 q0 = tri[3 * i], q1 = tri[3 * i + 1], q2 = tri[3 * i + 2];
        const float va[3] = {q0.x, q0.y, q0.z}, vb[3] = {q0.x + q1.x, q0.y + q1.y, q0.z + q1.z}, vc[3] = {q0.x + q2.x, q0.y + q2.y, q0.z + q2.z};
        b.grow(va); b.grow(vb); b.grow(vc);
      } else {
        const float4 A = sA[i], B = sB[i];
        const float r = std::fabs(B.w);
        const float c0[3] = {A.x, A.y, A.z}, c1[3] = {A.x + B.x, A.y + B.y, A.z + B.z};
        for (int k = 0; k < 3; ++k) { b.lo[k] = std::min(c0[k], c1[k]) - r; b.hi[k] = std::max(c0[k], c1[k]) + r; }
      }
      return b;
    }
    rtw::Box3 visit(int32_t n, int depth) {
      rtw::Box3 all; all.reset();
      if (n < 0 || n >= n_nodes || depth > 200) { ++errors; return all; }
      depth_max = std::max(depth_max, depth);
      const rtw::PackedNode& nd = nodes[n];
      for (int side = 0; side < 2; ++side) {
        const int32_t code = side ? nd.right : nd.left;
        const float c[3] = {nd.c[0][side], nd.c[1][side], nd.c[2][side]};
        const float e[3] = {nd.e[0][side], nd.e[1][side], nd.e[2][side]};
        if (side == 1 && n_nodes == 1 && c[2] >= 1.0e38f) continue;   // the never-entered filler child of a single-leaf tree
        const rtw::Box3 below = code >= 0 ? visit(code, depth + 1) : prim_box(static_cast<uint32_t>(~code));
        for (int k = 0; k < 3; ++k)
          if (!(static_cast<double>(c[k]) - e[k] <= below.lo[k] && below.hi[k] <= static_cast<double>(c[k]) + e[k])) { ++errors; break; }
        all.grow(below);
      }
      return all;
    }
  } walk{nodes, sA, sB, tri, seen, n_small, errors, depth_max, d.n_nodes};
  if (d.n_nodes > 0) walk.visit(0, 1);
  for (uint8_t c : seen) if (c != 1) ++errors;
  out->n_static_spheres = d.n_static; out->n_moving_spheres = d.n_moving; out->n_big_spheres = d.n_big; out->n_triangles = d.n_tri;
  out->n_bvh_nodes = d.n_nodes; out->bvh_max_depth = depth_max; out->leaf_direct = d.leaf_direct; out->arena_bytes = static_cast<int64_t>(scene->arena_bytes);
  out->bvh_errors = errors; out->bvh_build_ms = scene->gpu_build_ms;
  return 0;
}

int rtw_scene_hash(const rtw_scene_desc* desc, uint64_t* out) {
  if (!desc || !out) return fail("rtw_scene_hash: null argument");
  if (desc->nprims < 0 || desc->nmats < 0 || (desc->nprims > 0 && !desc->prims) || (desc->nmats > 0 && !desc->mats)) return fail("rtw_scene_hash: invalid scene description");
  *out = rtw::scene_key(desc);
  return 0;
}

int rtw_scene_update(rtw_scene* scene, const rtw_scene_desc* desc) {
  if (!scene || !desc) return fail("rtw_scene_update: null argument");
  HostFlat hf;
  if (int rc = rtw::flatten_host(desc, &hf)) return rc;
  // same device, same allocation when the new tables fit (grow-only): no cudaMalloc in the steady state
  return rtw::upload_flat(hf, desc->camera, desc->nprims, scene->device, scene, &scene->arena, nullptr);
}

unsigned long long rtw_kernel_launches(void) { return rtw::launch_count(); }

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
  // this launch's own counter block (work-queue cursor + statistics): renders of one scene may overlap on other streams / threads
  p.counters = scene->counters + static_cast<size_t>(scene->launch_seq.fetch_add(1, std::memory_order_relaxed) % rtw::kCtrSlots) * rtw::kCtrCount;
  const int rpl = cfg->rays_per_lane;
  const bool want_stats = (cfg->flags & RTW_FLAG_STATS) != 0;
  RTW_CUDA(cudaMemsetAsync(p.counters, 0, rtw::kCtrCount * sizeof(unsigned long long), stream));
  rtw::EventPair ev;
  if (stats) {
    RTW_CUDA(ev.create());
    RTW_CUDA(cudaEventRecord(ev.a, stream));
  }
  int variant = RTW_BVH_NONE;
  cudaError_t le = rtw::launch_render(p, mode, rpl, cfg->kernel == RTW_KERNEL_BVH_PERLANE, want_stats, scene->sm_count, stream, &variant);
  if (le != cudaSuccess) return fail_cuda("launch k_render", le);
  if (stats) {
    RTW_CUDA(cudaEventRecord(ev.b, stream));
    unsigned long long h[rtw::kCtrCount];
    RTW_CUDA(cudaMemcpyAsync(h, p.counters, sizeof h, cudaMemcpyDeviceToHost, stream));
    RTW_CUDA(cudaStreamSynchronize(stream));
    float ms = 0.f;
    RTW_CUDA(cudaEventElapsedTime(&ms, ev.a, ev.b));
    std::memset(stats, 0, sizeof *stats);
    read_counters(h, stats);
    stats->kernel_ms = ms;
    stats->kernel_used = mode == 0 ? RTW_KERNEL_SPHERES_SMEM : RTW_KERNEL_BVH;
    stats->bvh_variant = variant;
    stats->launches = 1;
  }
  return 0;
}

int32_t rtw_row_tile_local_rows(int32_t height, int32_t tile_rows, int32_t count) {
  if (height < 1 || tile_rows < 1 || count < 1) return 0;
  const int32_t tiles = (height + tile_rows - 1) / tile_rows;
  return (tiles + count - 1) / count * tile_rows;
}

int rtw_untile_accum(const int64_t* gathered, int64_t* accum_fx, int32_t width, int32_t height, int32_t tile_rows, int32_t count,
                     int32_t device, void* cuda_stream) {
  if (!gathered || !accum_fx || width < 1 || height < 1 || tile_rows < 1 || count < 1) return fail("rtw_untile_accum: invalid argument");
  RTW_CUDA(cudaSetDevice(device));
  cudaError_t e = rtw::launch_untile(reinterpret_cast<const long long*>(gathered), reinterpret_cast<long long*>(accum_fx), static_cast<uint32_t>(width),
                                     static_cast<uint32_t>(height), static_cast<uint32_t>(tile_rows), static_cast<uint32_t>(count),
                                     static_cast<uint32_t>(rtw_row_tile_local_rows(height, tile_rows, count)), static_cast<cudaStream_t>(cuda_stream));
  if (e != cudaSuccess) return fail_cuda("launch k_untile", e);
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

int rtw_finalize_rgb8_device(const int64_t* accum_fx, int64_t npixels, int32_t spp, int32_t device, void* cuda_stream, uint8_t* rgb8) {
  if (!accum_fx || !rgb8 || npixels <= 0 || spp <= 0) return fail("rtw_finalize_rgb8_device: invalid argument");
  RTW_CUDA(cudaSetDevice(device));
  cudaError_t e = rtw::launch_finalize_rgb8_fx(reinterpret_cast<const long long*>(accum_fx), rgb8, npixels, spp, static_cast<cudaStream_t>(cuda_stream));
  if (e != cudaSuccess) return fail_cuda("launch k_finalize_rgb8_fx", e);
  return 0;
}

int rtw_prewarm(int32_t first_device, int32_t ngpus) {
  if (first_device < 0 || ngpus < 1 || first_device + ngpus > 64) return fail("rtw_prewarm: bad device range");
  std::lock_guard<std::mutex> lock(g_warm_mutex);
  for (int d = first_device; d < first_device + ngpus; ++d) {
    if (g_warm_started >> d & 1) continue;
    g_warm_started |= 1ull << d;
    g_warm_threads.emplace_back([d] {
      const double t0 = now_ms();
      if (cudaSetDevice(d) == cudaSuccess) cudaFree(nullptr);   // creates the primary context; errors surface in the render call
      cudaGetLastError();
      if (std::getenv("RTW_TRACE")) std::fprintf(stderr, "rtw trace: context of gpu %d up after %.1f ms\n", d, now_ms() - t0);
    });
  }
  return 0;
}

void rtw_release_cached_buffers(void) {
  rtw::release_build_scratch();
  std::lock_guard<std::mutex> lock(g_slots_mutex);
  for (int d = 0; d < 64; ++d) {
    rtw::DeviceSlot* s = g_slots[d];
    if (!s) continue;
    std::lock_guard<std::mutex> busy(s->m);
    cudaSetDevice(d);
    if (s->stream) cudaStreamDestroy(s->stream);
    if (s->rendered) cudaEventDestroy(s->rendered);
    s->arena.alloc(0); s->fx.alloc(0); s->out_f32.alloc(0); s->out_u8.alloc(0);
    s->stream = nullptr; s->rendered = nullptr; s->key = 0;
  }
}

// Host-buffer render on one GPU.  The device slot keeps the uploaded scene (keyed on a hash of the caller's arrays) and the
// accumulation buffers between calls: a repeated call with the same scene costs the hash, the kernels and one download.
// out_f32 != nullptr: (sum r, sum g, sum b, samples) as floats, 16 bytes per pixel; out_u8 != nullptr: write_color on the device
// from the exact int64 sums, 3 bytes per pixel.
static int render_host(const rtw_scene_desc* desc, const rtw_render_cfg* cfg, float* out_f32, uint8_t* out_u8, rtw_stats* stats) {
  if (!desc || !cfg || (!out_f32 && !out_u8)) return fail("rtw_render: null argument");
  if (cfg->width < 2 || cfg->height < 2) return fail("render: width and height must be >= 2 (pixel mapping divides by W-1, H-1)");
  if (cfg->row_tile_count > 1) return fail("rtw_render: the row-tile split goes through rtw_render_device or rtw_render_multi_gpu");
  const double t_start = now_ms();
  rtw::prewarm_join();
  rtw::DeviceSlot* slot = rtw::device_slot(cfg->device);
  if (!slot) return fail("rtw_render: device ordinal out of range");
  std::lock_guard<std::mutex> lock(slot->m);   // host-buffer renders are serialised per device
  if (int rc = rtw::slot_prepare(slot)) return rc;
  HostFlat hf;
  bool hit = false;
  const bool use_cache = (cfg->flags & RTW_FLAG_NO_SCENE_CACHE) == 0;
  if (desc->nprims < 0 || desc->nmats < 0 || (desc->nprims > 0 && !desc->prims) || (desc->nmats > 0 && !desc->mats)) return fail("rtw_scene_upload: invalid scene description");
  hf.gpu_build = rtw::choose_gpu_build(desc, cfg);
  const uint64_t key = rtw::scene_key(desc) ^ (hf.gpu_build ? 0x6b9d0f5a1c2e3d47ull : 0ull);   // the device tables depend on the builder
  if (int rc = rtw::slot_set_scene(slot, desc, key, use_cache, &hf, nullptr, &hit)) return rc;
  const double t_up = now_ms();
  const size_t npix = static_cast<size_t>(cfg->width) * static_cast<size_t>(cfg->height);
  RTW_CUDA(slot->fx.reserve(npix * 4));
  cudaStream_t stream = slot->stream;
  RTW_CUDA(cudaMemsetAsync(slot->fx.p, 0, npix * 4 * sizeof(long long), stream));
  rtw_stats st{};
  if (int rc = rtw_render_device(&slot->scene, cfg, reinterpret_cast<int64_t*>(slot->fx.p), stream, &st)) return rc;
  const double t_d0 = now_ms();
  int launches = 1;
  if (out_f32) {
    RTW_CUDA(slot->out_f32.reserve(npix * 4));
    if (int rc = rtw_accum_to_float(reinterpret_cast<const int64_t*>(slot->fx.p), slot->out_f32.p, static_cast<int64_t>(npix), cfg->device, stream)) return rc;
    RTW_CUDA(cudaMemcpyAsync(out_f32, slot->out_f32.p, npix * 4 * sizeof(float), cudaMemcpyDeviceToHost, stream));
    ++launches;
  }
  if (out_u8) {
    RTW_CUDA(slot->out_u8.reserve(npix * 3));
    if (int rc = rtw_finalize_rgb8_device(reinterpret_cast<const int64_t*>(slot->fx.p), static_cast<int64_t>(npix), cfg->sample_end - cfg->sample_begin,
                                          cfg->device, stream, slot->out_u8.p)) return rc;
    RTW_CUDA(cudaMemcpyAsync(out_u8, slot->out_u8.p, npix * 3, cudaMemcpyDeviceToHost, stream));
    ++launches;
  }
  RTW_CUDA(cudaStreamSynchronize(stream));
  const double t_end = now_ms();
  if (stats) {
    *stats = st;
    stats->h2d_ms = t_up - t_start;
    stats->d2h_ms = t_end - t_d0;
    stats->total_ms = t_end - t_start;
    stats->launches = launches;
    stats->scene_cache_hit = hit ? 1 : 0;
    stats->bvh_build_gpu_ms = slot->scene.gpu_build_ms;
  }
  return 0;
}

int rtw_render(const rtw_scene_desc* desc, const rtw_render_cfg* cfg, float* accum_rgba, rtw_stats* stats) {
  if (!accum_rgba) return fail("rtw_render: null argument");
  return render_host(desc, cfg, accum_rgba, nullptr, stats);
}

int rtw_render_rgb8(const rtw_scene_desc* desc, const rtw_render_cfg* cfg, uint8_t* rgb8, rtw_stats* stats) {
  if (!rgb8) return fail("rtw_render_rgb8: null argument");
  return render_host(desc, cfg, nullptr, rgb8, stats);
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
  // scalar FFMA chains, then packed FFMA2 chains (twice the FMAs per instruction): half of `seconds` each, the higher rate is the peak
  double best = 0.0;
  for (int packed = 0; packed < 2; ++packed) {
    const double flop_per_launch = static_cast<double>(blocks) * 256.0 * iters * 16.0 * 8.0 * 2.0 * (packed ? 2.0 : 1.0);
    for (int w = 0; w < 3; ++w) { cudaError_t e = rtw::launch_ffma_peak(out.p, blocks, iters, packed != 0, nullptr); if (e != cudaSuccess) return fail_cuda("launch k_ffma_peak", e); }
    RTW_CUDA(cudaDeviceSynchronize());
    double total_ms = 0.0; int launches = 0;
    const double t_begin = now_ms();
    do {
      RTW_CUDA(cudaEventRecord(e0));
      for (int k = 0; k < 8; ++k) { cudaError_t e = rtw::launch_ffma_peak(out.p, blocks, iters, packed != 0, nullptr); if (e != cudaSuccess) return fail_cuda("launch k_ffma_peak", e); }
      RTW_CUDA(cudaEventRecord(e1));
      RTW_CUDA(cudaEventSynchronize(e1));
      float ms = 0.f; RTW_CUDA(cudaEventElapsedTime(&ms, e0, e1));
      total_ms += ms; launches += 8;
    } while (now_ms() - t_begin < seconds * 500.0);
    best = std::max(best, flop_per_launch * launches / (total_ms * 1e-3) / 1e12);
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  *tflops = best;
  if (sm_mhz) {
    // clock implied by the measured rate if every SM issued 128 FMA lanes per cycle
    *sm_mhz = (*tflops * 1e12) / (2.0 * 128.0 * prop.multiProcessorCount) / 1e6;
  }
  return 0;
}

}  // extern "C"
