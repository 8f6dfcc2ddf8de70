// rtw_kernels.cu -- sm_100a kernels of the B200 path tracer and their launchers.
//
//   The three render kernels are the per-pixel / per-sample loop of render.cpp:150-167 with ray_color (render.cpp:112-129) turned
//   into an iterative bounce loop.  Persistent warps pull units of (16 x 8 pixel tile x SU samples) from a global atomic counter; a
//   finished path is replaced at once from the warp's pool (ballot + popc ranking).
//   k_render_wf (K2w)   sphere scenes whose tables leave room in shared memory: SAH BVH (64-byte two-child nodes, single-primitive
//                       leaves), tables staged with TMA bulk copies, the warp's paths kept as records in shared memory, shading in
//                       batches of 32 records of one kind ("wavefront per warp").
//   k_render_bvh (K2)   meshes and big sphere scenes: the same tree walked by a resumable per-lane state machine, tables in
//                       shared memory when they fit, else read through L1/L2 (nodes with two 256-bit loads).
//   k_render_sweep<R> (K1)  small sphere scenes: the whole sphere table staged in shared memory (cp.async.bulk + mbarrier)
//                       and swept brute force, R paths per lane: per (ray, sphere) 8 FMA for the line-distance reject test
//                       (+3 for the centre lerp of moving spheres); survivors go through the exact reference-rule root selection.
//   k_primary_f32       K3: deterministic primary hits through the SAME tracing routines (parity mode)
//   k_primary_f64       K3 in double: reference formulas verbatim, brute force over the raw primitive list
//   k_accum_to_float, k_finalize_rgb8, k_untile   render.cpp:11-20,176-186 on the device; row-tile split reassembly
//   k_debug_*, k_ffma_peak              unit hooks and the FP32 roofline denominator
//
// Radiance is accumulated as int64 fixed point (2^-32) with RED.ADD.64: integer sums are exact, so the image
// does not depend on scheduling order or on how samples are sharded over GPUs.
#include <atomic>
#include <cstdlib>
#include <algorithm>

#include "rtw_internal.h"

namespace rtw {

// ---------------------------------------------------------------------------------------------------------
// TMA bulk copy + mbarrier (PTX; SASS: UBLKCP / SYNCS)
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t phase) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(phase)
      : "memory");
  return ok != 0;
}

// Stage the sphere tables into shared memory.  sA/sB: n float4 each.  One thread issues the copies.
__device__ __forceinline__ void stage_spheres(const DevScene& sc, float4* sA, float4* sB, uint64_t* bar) {
  const int n = sc.n_static + sc.n_moving;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0 && n > 0) {
    const uint32_t bytes = static_cast<uint32_t>(n) * 16u;
    mbar_expect_tx(bar, 2u * bytes);
    constexpr uint32_t kChunk = 32768u;
    for (uint32_t off = 0; off < bytes; off += kChunk) {
      const uint32_t sz = min(kChunk, bytes - off);
      tma_bulk_g2s(reinterpret_cast<char*>(sA) + off, reinterpret_cast<const char*>(sc.sphA) + off, sz, bar);
      tma_bulk_g2s(reinterpret_cast<char*>(sB) + off, reinterpret_cast<const char*>(sc.sphB) + off, sz, bar);
    }
  }
  if (n > 0) {
    // bounded spin: a copy that never lands must surface as a launch failure, not as a hung GPU
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, 0)) {
      if (++spins > (1u << 26)) __trap();
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// K1 tracer: brute-force sweep of the shared-memory sphere table for R rays per lane.
// Reject test per (ray, sphere): with (u, v) an orthonormal basis of the plane perpendicular to the ray,
// |((c - o).u, (c - o).v)|^2 <= r^2  <=>  the ray's LINE passes within r of the centre.  8 FMA for a static
// sphere, 11 for a moving one; r2c carries a conservative slack so fp32 rounding can only add candidates.
// The sweep is branch-free: the sign bit of each test is funnel-shifted into a per-ray candidate mask (one SHF per
// test); after every kMaskWords x 32 spheres the ~1 candidate per ray goes through the exact root selection.
// ---------------------------------------------------------------------------------------------------------
template <int R>
struct RaySet {
  float ox[R], oy[R], oz[R], dx[R], dy[R], dz[R], tm[R];
};

// 64 spheres between candidate flushes, sweep unrolled by 4: measured best of {1,2,4} words x {2,4,8,16} (1627 Mpaths/s against
// 1286-1611 for the others on the cover scene)
constexpr int kMaskWords = 2;
constexpr int kSweepUnroll = 4;

template <int R, bool STATS>
__device__ __forceinline__ void trace_spheres(const DevScene& sc, const float4* __restrict__ sA, const float4* __restrict__ sB,
                                              const RaySet<R>& ray, const bool (&alive)[R], float (&best_t)[R], int (&best_i)[R],
                                              unsigned long long& n_cand) {
  float ux[R], uy[R], uz[R], vx[R], vy[R], vz[R], ou[R], ov[R], aa[R], ia[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const float a = ray.dx[r] * ray.dx[r] + ray.dy[r] * ray.dy[r] + ray.dz[r] * ray.dz[r];
    const float il = rsqrtf(a);
    const float nx = ray.dx[r] * il, ny = ray.dy[r] * il, nz = ray.dz[r] * il;
    // branchless orthonormal basis (Duff et al. 2017)
    const float sg = copysignf(1.0f, nz);
    const float p = -1.0f / (sg + nz);
    const float q = nx * ny * p;
    ux[r] = 1.0f + sg * nx * nx * p; uy[r] = sg * q; uz[r] = -sg * nx;
    vx[r] = q; vy[r] = sg + ny * ny * p; vz[r] = -ny;
    ou[r] = ray.ox[r] * ux[r] + ray.oy[r] * uy[r] + ray.oz[r] * uz[r];
    ov[r] = ray.ox[r] * vx[r] + ray.oy[r] * vy[r] + ray.oz[r] * vz[r];
    aa[r] = a; ia[r] = 1.0f / a;
    best_t[r] = kInf; best_i[r] = kMiss;
    if (!alive[r]) { ux[r] = uy[r] = uz[r] = vx[r] = vy[r] = vz[r] = 0.0f; ou[r] = 1e18f; ov[r] = 0.0f; }
  }
  const int ns = sc.n_static, nt = sc.n_static + sc.n_moving;

  uint32_t mm[R][kMaskWords];
  int word_base[kMaskWords];
#pragma unroll
  for (int r = 0; r < R; ++r)
#pragma unroll
    for (int w = 0; w < kMaskWords; ++w) mm[r][w] = 0u;

  // exact test of the flagged spheres of the words gathered so far
  auto flush = [&](int nwords) {
    uint32_t any = 0u;
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int w = 0; w < kMaskWords; ++w) any |= mm[r][w];
    if (any != 0u) {
#pragma unroll
      for (int r = 0; r < R; ++r) {
#pragma unroll
        for (int w = 0; w < kMaskWords; ++w) {
          if (w >= nwords) continue;
          uint32_t m = mm[r][w];
          while (m != 0u) {
            const int b = __clz(m);
            m &= ~(0x80000000u >> b);
            const int i = word_base[w] + b;
            if (STATS) ++n_cand;
            const float4 A = sA[i], B = sB[i];
            const float t = sphere_hit_t<float>(mk<float>(ray.ox[r], ray.oy[r], ray.oz[r]), mk<float>(ray.dx[r], ray.dy[r], ray.dz[r]), aa[r],
                                                ia[r], mk<float>(fmaf(ray.tm[r], B.x, A.x), fmaf(ray.tm[r], B.y, A.y), fmaf(ray.tm[r], B.z, A.z)),
                                                B.w, kTMin, best_t[r]);
            if (t >= 0.0f) { best_t[r] = t; best_i[r] = i; }
          }
        }
      }
    }
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int w = 0; w < kMaskWords; ++w) mm[r][w] = 0u;
  };

  // one word = up to 32 consecutive table entries; bit 31-j of the word's mask <- sign of the test of entry base+j
  // R even: the rays of a lane are handled in PAIRS by the packed FP32 FMA of sm_100 (PTX fma.rn.f32x2, SASS FFMA2): the per-ray
  // constants sit in 64-bit register pairs, the table entry enters as the instruction's broadcast operand, and the 16 (22) FFMA
  // per static (moving) sphere and ray pair become 8 (11) FFMA2.  Same fma.rn per half: the masks are bit-identical to the scalar
  // form.  FFMA2 issues every other cycle at the same FP32 peak (scripts/ffma2_probe.cu), so the loop's LDS / SHF / control now
  // fit into the free issue slots: this kernel is bound by the FMA pipe, not by latency like the tree walks.
#ifdef RTW_SWEEP_SCALAR   // A/B build
  constexpr bool PAIRS = false;
#else
  constexpr bool PAIRS = (R % 2 == 0);
#endif
  constexpr int H = PAIRS ? R / 2 : 1;
  uint64_t ux2[H], uy2[H], uz2[H], vx2[H], vy2[H], vz2[H], nou2[H], nov2[H], tm2[H];
  if (PAIRS) {
#pragma unroll
    for (int h = 0; h < H; ++h) {
      const int a = 2 * h, b = PAIRS ? 2 * h + 1 : 0;
      ux2[h] = pk2(ux[a], ux[b]); uy2[h] = pk2(uy[a], uy[b]); uz2[h] = pk2(uz[a], uz[b]);
      vx2[h] = pk2(vx[a], vx[b]); vy2[h] = pk2(vy[a], vy[b]); vz2[h] = pk2(vz[a], vz[b]);
      nou2[h] = pk2(-ou[a], -ou[b]); nov2[h] = pk2(-ov[a], -ov[b]); tm2[h] = pk2(ray.tm[a], ray.tm[b]);
    }
  }
  auto sweep_word = [&](int base, int cnt, bool moving, uint32_t (&m)[R]) {
#pragma unroll
    for (int r = 0; r < R; ++r) m[r] = 0u;
    // software pipeline: the next table entry is in flight while the current one is tested (the tables carry one
    // spare entry so the last prefetch stays in bounds)
    if (!moving) {
      float4 A = sA[base];
#pragma unroll kSweepUnroll
      for (int j = 0; j < cnt; ++j) {
        const float4 An = sA[base + j + 1];
        if (PAIRS) {
          const uint64_t ax = pk2(A.x, A.x), ay = pk2(A.y, A.y), az = pk2(A.z, A.z), nw = pk2(-A.w, -A.w);
#pragma unroll
          for (int h = 0; h < H; ++h) {
            const uint64_t pu = fma2(ax, ux2[h], fma2(ay, uy2[h], fma2(az, uz2[h], nou2[h])));
            const uint64_t pv = fma2(ax, vx2[h], fma2(ay, vy2[h], fma2(az, vz2[h], nov2[h])));
            float s0, s1;
            up2(fma2(pu, pu, fma2(pv, pv, nw)), s0, s1);
            m[2 * h] = __funnelshift_l(__float_as_uint(s0), m[2 * h], 1);
            m[PAIRS ? 2 * h + 1 : 0] = __funnelshift_l(__float_as_uint(s1), m[PAIRS ? 2 * h + 1 : 0], 1);
          }
        } else {
#pragma unroll
          for (int r = 0; r < R; ++r) {
            const float pu = fmaf(A.x, ux[r], fmaf(A.y, uy[r], fmaf(A.z, uz[r], -ou[r])));
            const float pv = fmaf(A.x, vx[r], fmaf(A.y, vy[r], fmaf(A.z, vz[r], -ov[r])));
            const float sgn = fmaf(pu, pu, fmaf(pv, pv, -A.w));
            m[r] = __funnelshift_l(__float_as_uint(sgn), m[r], 1);
          }
        }
        A = An;
      }
    } else {
      float4 A = sA[base], B = sB[base];
#pragma unroll kSweepUnroll
      for (int j = 0; j < cnt; ++j) {
        const float4 An = sA[base + j + 1];
        const float4 Bn = sB[base + j + 1];
        if (PAIRS) {
          const uint64_t ax = pk2(A.x, A.x), ay = pk2(A.y, A.y), az = pk2(A.z, A.z), nw = pk2(-A.w, -A.w);
          const uint64_t bx = pk2(B.x, B.x), by = pk2(B.y, B.y), bz = pk2(B.z, B.z);
#pragma unroll
          for (int h = 0; h < H; ++h) {
            const uint64_t cx = fma2(tm2[h], bx, ax), cy = fma2(tm2[h], by, ay), cz = fma2(tm2[h], bz, az);
            const uint64_t pu = fma2(cx, ux2[h], fma2(cy, uy2[h], fma2(cz, uz2[h], nou2[h])));
            const uint64_t pv = fma2(cx, vx2[h], fma2(cy, vy2[h], fma2(cz, vz2[h], nov2[h])));
            float s0, s1;
            up2(fma2(pu, pu, fma2(pv, pv, nw)), s0, s1);
            m[2 * h] = __funnelshift_l(__float_as_uint(s0), m[2 * h], 1);
            m[PAIRS ? 2 * h + 1 : 0] = __funnelshift_l(__float_as_uint(s1), m[PAIRS ? 2 * h + 1 : 0], 1);
          }
        } else {
#pragma unroll
          for (int r = 0; r < R; ++r) {
            const float cx = fmaf(ray.tm[r], B.x, A.x), cy = fmaf(ray.tm[r], B.y, A.y), cz = fmaf(ray.tm[r], B.z, A.z);
            const float pu = fmaf(cx, ux[r], fmaf(cy, uy[r], fmaf(cz, uz[r], -ou[r])));
            const float pv = fmaf(cx, vx[r], fmaf(cy, vy[r], fmaf(cz, vz[r], -ov[r])));
            const float sgn = fmaf(pu, pu, fmaf(pv, pv, -A.w));
            m[r] = __funnelshift_l(__float_as_uint(sgn), m[r], 1);
          }
        }
        A = An; B = Bn;
      }
    }
    if (cnt < 32) {
#pragma unroll
      for (int r = 0; r < R; ++r) m[r] <<= (32 - cnt);
    }
  };

  int w = 0;
#pragma unroll 1
  for (int seg = 0; seg < 2; ++seg) {
    const int seg_begin = seg == 0 ? 0 : ns, seg_end = seg == 0 ? ns : nt;
#pragma unroll 1
    for (int base = seg_begin; base < seg_end; base += 32) {
      uint32_t m[R];
      sweep_word(base, min(32, seg_end - base), seg == 1, m);
#pragma unroll
      for (int k = 0; k < kMaskWords; ++k)
        if (k == w) {
          word_base[k] = base;
#pragma unroll
          for (int r = 0; r < R; ++r) mm[r][k] = m[r];
        }
      if (++w == kMaskWords) { flush(kMaskWords); w = 0; }
    }
  }
  if (w > 0) flush(w);
#pragma unroll
  for (int r = 0; r < R; ++r)
    if (alive[r]) trace_big_spheres(sc, mk<float>(ray.ox[r], ray.oy[r], ray.oz[r]), mk<float>(ray.dx[r], ray.dy[r], ray.dz[r]),
                                    ray.tm[r], best_t[r], best_i[r]);
}

template <bool SMEM>
__device__ __forceinline__ float4 ld4(const float4* p, int i) { return SMEM ? p[i] : __ldg(p + i); }
template <bool SMEM>
__device__ __forceinline__ uint32_t ld1(const uint32_t* p, int i) { return SMEM ? p[i] : __ldg(p + i); }

// One leaf primitive of a compressed-wide-BVH scene: record P of the leaf-ordered table (DevScene::tri).
template <bool SMEM, bool STATS>
__device__ __forceinline__ void cw_test_prim(const DevScene& sc, const float4* __restrict__ prims, uint32_t P, F3 o, F3 d, float qa, float qia, float tm,
                                             float& best_t, int& best_i, unsigned long long& n_sph, unsigned long long& n_tri) {
  const int b = 3 * static_cast<int>(P);
  const float4 q0 = ld4<SMEM>(prims, b), q1 = ld4<SMEM>(prims, b + 1), q2 = ld4<SMEM>(prims, b + 2);
  if (sc.cw_has_spheres && q2.x != q2.x) {   // NaN tag: a small sphere (sphA entry, sphB entry, table index)
    if (STATS) ++n_sph;
    const float t = sphere_hit_fast(o, d, qa, qia, mk<float>(fmaf(tm, q1.x, q0.x), fmaf(tm, q1.y, q0.y), fmaf(tm, q1.z, q0.z)), q1.w, kTMin, best_t);
    if (t >= 0.0f) { best_t = t; best_i = __float_as_int(q2.y); }
  } else {
    if (STATS) ++n_tri;
    const float t = triangle_hit_fast(o, d, mk<float>(q0.x, q0.y, q0.z), mk<float>(q1.x, q1.y, q1.z), mk<float>(q2.x, q2.y, q2.z),
                                        mk<float>(q0.w, q1.w, q2.w), kTMin, best_t);
    if (t >= 0.0f) { best_t = t; best_i = kHitTri | static_cast<int>(P); }
  }
}

// Compressed wide BVH, one ray per lane, run to completion (K3 primary-hit mode; the render kernel runs the same steps as a
// resumable state machine).
template <bool STATS>
__device__ __forceinline__ void trace_cw(const DevScene& sc, F3 o, F3 d, float tm, float& best_t, int& best_i,
                                         unsigned long long& n_nodes, unsigned long long& n_sph, unsigned long long& n_tri) {
  best_t = kInf; best_i = kMiss;
  const float qa = dot(d, d), qia = fast_rcp(qa);
  const float idx = fast_rcp(d.x), idy = fast_rcp(d.y), idz = fast_rcp(d.z);
  const uint32_t oct_inv4 = cw_oct_inv4(d);
  uint2 stack[kCwStack];
  int sp = 0;
  uint2 ngroup = make_uint2(0u, sc.n_cw_nodes > 0 ? 0x80000000u : 0u), tgroup = make_uint2(0u, 0u);
  for (;;) {
    if (ngroup.y & 0xff000000u) {
      const uint32_t child = cw_pop_child(ngroup, oct_inv4, stack, sp);
      if (STATS) ++n_nodes;
      cw_intersect_node<false>(sc.cwNodes, child, o, idx, idy, idz, oct_inv4, kTMin, best_t, ngroup, tgroup);
    }
    while (tgroup.y != 0u) {
      const uint32_t k = 31u - __clz(tgroup.y);
      tgroup.y &= ~(1u << k);
      cw_test_prim<false, STATS>(sc, sc.tri, tgroup.x + k, o, d, qa, qia, tm, best_t, best_i, n_sph, n_tri);
    }
    if ((ngroup.y & 0xff000000u) == 0u) {
      if (sp == 0) break;
      ngroup = stack[--sp];
    }
  }
  trace_big_spheres(sc, o, d, tm, best_t, best_i);
}

// ---------------------------------------------------------------------------------------------------------
// K2 tracer: BVH traversal (one ray per lane), closest hit, near child first.
// ---------------------------------------------------------------------------------------------------------
template <bool STATS>
__device__ __forceinline__ void trace_bvh(const DevScene& sc, F3 o, F3 d, float tm, float& best_t, int& best_i,
                                          unsigned long long& n_nodes, unsigned long long& n_sph, unsigned long long& n_tri) {
  best_t = kInf; best_i = kMiss;
  const float a = dot(d, d), ia = 1.0f / a;
  const float idx = 1.0f / d.x, idy = 1.0f / d.y, idz = 1.0f / d.z;
  const float odx = o.x * idx, ody = o.y * idy, odz = o.z * idz;
  int stack[kBvhStack];
  int sp = 0;
  int node = sc.n_nodes > 0 ? 0 : kMiss;

  auto leaf = [&](int code) {
    const uint32_t v = static_cast<uint32_t>(~code);
    const uint32_t first = v >> 5, cnt = sc.leaf_direct ? 1u : (v & 31u);
    for (uint32_t k = 0; k < cnt; ++k) {
      const uint32_t ref = sc.leaf_direct ? v : __ldg(&sc.leafRefs[first + k]);
      const int i = static_cast<int>(ref & 0x1fffffffu);
      if (ref >> 30) {
        if (STATS) ++n_tri;
        const float4 q0 = __ldg(&sc.tri[3 * i]), q1 = __ldg(&sc.tri[3 * i + 1]), q2 = __ldg(&sc.tri[3 * i + 2]);
        const float t = triangle_hit_fast(o, d, mk<float>(q0.x, q0.y, q0.z), mk<float>(q1.x, q1.y, q1.z), mk<float>(q2.x, q2.y, q2.z),
                                            mk<float>(q0.w, q1.w, q2.w), kTMin, best_t);
        if (t >= 0.0f) { best_t = t; best_i = kHitTri | i; }
      } else {
        if (STATS) ++n_sph;
        const float4 A = __ldg(&sc.sphA[i]), B = __ldg(&sc.sphB[i]);
        const float t = sphere_hit_t<float>(o, d, a, ia, mk<float>(fmaf(tm, B.x, A.x), fmaf(tm, B.y, A.y), fmaf(tm, B.z, A.z)), B.w,
                                            kTMin, best_t);
        if (t >= 0.0f) { best_t = t; best_i = i; }
      }
    }
  };

  while (node != kMiss) {
    if (node < 0) {
      leaf(node);
      node = sp > 0 ? stack[--sp] : kMiss;
      continue;
    }
    if (STATS) ++n_nodes;
    float4 q0, q1, q2, q3;
    load_node<false>(sc.nodes, node, q0, q1, q2, q3);
    // slab test of both children against [tmin, best_t]
    float ln, lf, rn, rf;
    node_slabs(q0, q1, q2, idx, idy, idz, odx, ody, odz, best_t, ln, lf, rn, rf);
    const bool hl = ln <= lf, hr = rn <= rf;
    const int left = __float_as_int(q3.x), right = __float_as_int(q3.y);
    if (hl && hr) {
      const bool lfirst = ln <= rn;
      node = lfirst ? left : right;
      if (sp < kBvhStack) stack[sp++] = lfirst ? right : left;
    } else if (hl) {
      node = left;
    } else if (hr) {
      node = right;
    } else {
      node = sp > 0 ? stack[--sp] : kMiss;
    }
  }
  trace_big_spheres(sc, o, d, tm, best_t, best_i);
}

// ---------------------------------------------------------------------------------------------------------
// The render kernel
// ---------------------------------------------------------------------------------------------------------
// Row-tile split: pixel lp of the packed local buffer -> global image row i, column j and global pixel index (false: outside the image)
__device__ __forceinline__ bool local_to_global(const RenderParams& p, uint32_t lp, uint32_t& gp, uint32_t& i, uint32_t& j) {
  // a work group is a 16 x 8 pixel tile of the (local) image and 32 consecutive indices inside it form an 8 x 4 block, so that
  // the primary rays a warp starts together are neighbours in both directions (they share the top of the tree walk)
  const uint32_t g = lp / kGroupPixels, w = lp - g * kGroupPixels;
  const uint32_t ty = g / p.tiles_x, tx = g - ty * p.tiles_x;
  j = tx * 16u + ((w & 7u) | ((w >> 2) & 8u));
  i = ty * 8u + (((w >> 3) & 3u) | ((w >> 4) & 4u));
  if (j >= p.width || i * p.width >= p.npix) return false;
  if (p.tile_count > 1u) {
    const uint32_t t = i / p.tile_rows, r = i - t * p.tile_rows;
    i = (t * p.tile_count + p.tile_index) * p.tile_rows + r;
    if (i >= p.height) return false;
  }
  gp = i * p.width + j;
  return true;
}
// ... and back: where global pixel gp lives in the packed local accumulation buffer
__device__ __forceinline__ uint32_t accum_index(const RenderParams& p, uint32_t gp) {
  if (p.tile_count <= 1u) return gp;
  const uint32_t i = gp / p.width, j = gp - i * p.width;
  const uint32_t t = i / p.tile_rows, w = i - t * p.tile_rows;
  return ((t / p.tile_count) * p.tile_rows + w) * p.width + j;
}

__device__ __forceinline__ void accum_add(unsigned long long* accum, uint32_t pix, float r, float g, float b) {
  // NaN / negative / absurd contributions are dropped (the reference would print garbage for them)
  const float lim = 1.0e9f;
  r = (r >= 0.0f && r < lim) ? r : 0.0f;
  g = (g >= 0.0f && g < lim) ? g : 0.0f;
  b = (b >= 0.0f && b < lim) ? b : 0.0f;
  unsigned long long* px = accum + 4ull * pix;
  atomicAdd(px + 0, static_cast<unsigned long long>(__float2ll_rn(r * 4294967296.0f)));
  atomicAdd(px + 1, static_cast<unsigned long long>(__float2ll_rn(g * 4294967296.0f)));
  atomicAdd(px + 2, static_cast<unsigned long long>(__float2ll_rn(b * 4294967296.0f)));
}

template <int R, bool STATS>
__global__ void __launch_bounds__(kRenderThreads, (R >= 4 ? 2 : (R == 2 ? 3 : 4))) k_render_sweep(const __grid_constant__ RenderParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const DevScene& sc = p.sc;
  const int n_table = sc.n_static + sc.n_moving;
  float4* sA = reinterpret_cast<float4*>(smem_raw + 16);
  float4* sB = sA + n_table + 1;
  stage_spheres(sc, sA, sB, reinterpret_cast<uint64_t*>(smem_raw));

  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t lt = (1u << lane) - 1u;
  const uint2 key = make_uint2(static_cast<uint32_t>(p.seed), static_cast<uint32_t>(p.seed >> 32));

  RaySet<R> ray;
  float tr[R], tg[R], tb[R];
  uint32_t pix[R], smp[R];
  int depth[R];
  bool alive[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    alive[r] = false; pix[r] = 0; smp[r] = 0; depth[r] = 0;
    ray.ox[r] = ray.oy[r] = ray.oz[r] = 0.f; ray.dx[r] = ray.dy[r] = 0.f; ray.dz[r] = 1.f; ray.tm[r] = 0.f;
    tr[r] = tg[r] = tb[r] = 0.f;
  }
  uint32_t pool_next = 0, pool_end = 0, grp = 0, s0 = 0;
  bool exhausted = false;
  unsigned long long n_rays = 0, n_paths = 0, n_tests = 0, n_cand = 0;

  for (;;) {
    // ---- refill: every dead slot takes the next (pixel, sample) of the warp's pool --------------------------
#pragma unroll
    for (int r = 0; r < R; ++r) {
#pragma unroll 1
      for (int pass = 0; pass < 2; ++pass) {
        const uint32_t need = __ballot_sync(0xffffffffu, !alive[r]);
        if (need == 0u) break;
        if (pool_next == pool_end) {
          if (exhausted) break;
          unsigned long long u = 0;
          if (lane == 0) u = atomicAdd(p.counters + kCtrWork, 1ull);
          u = __shfl_sync(0xffffffffu, u, 0);
          if (u >= p.n_units) { exhausted = true; break; }
          grp = static_cast<uint32_t>(u / p.n_chunks);
          const uint32_t chunk = static_cast<uint32_t>(u - static_cast<unsigned long long>(grp) * p.n_chunks);
          s0 = p.s_begin + chunk * p.su;
          const uint32_t ns = min(p.su, p.s_end - s0);
          pool_next = 0; pool_end = ns * kGroupPixels;
        }
        const uint32_t avail = pool_end - pool_next;
        const uint32_t rank = __popc(need & lt);
        if (!alive[r] && rank < avail) {
          const uint32_t within = pool_next + rank;
          const uint32_t lp = grp * kGroupPixels + (within & (kGroupPixels - 1u));
          const uint32_t sm = s0 + within / kGroupPixels;
          uint32_t px, i, j;
          if (local_to_global(p, lp, px, i, j)) {
            // primary ray: render.cpp:158-160 with the pixel mapping of SURVEY Q12
            const uint4 x0 = philox4x32_10(make_uint4(px, sm, 0u, 0u), key);
            const float u = (static_cast<float>(j) + u01(x0.x)) * p.inv_wm1;
            const float v = (static_cast<float>(p.height - 1u - i) + u01(x0.y)) * p.inv_hm1;
            const float2 dk = sample_disk(u01(x0.z), u01(x0.w));
            F3 o, d;
            camera_ray<float>(sc.cam, u, v, dk.x, dk.y, o, d);
            ray.ox[r] = o.x; ray.oy[r] = o.y; ray.oz[r] = o.z;
            ray.dx[r] = d.x; ray.dy[r] = d.y; ray.dz[r] = d.z;
            ray.tm[r] = fmaf(u01_low(x0), sc.cam.t1 - sc.cam.t0, sc.cam.t0);
            tr[r] = tg[r] = tb[r] = 1.0f;
            pix[r] = px; smp[r] = sm; depth[r] = 0;
            alive[r] = true;
          }
        }
        pool_next += min(static_cast<uint32_t>(__popc(need)), avail);
      }
    }
    bool any_alive = false;
#pragma unroll
    for (int r = 0; r < R; ++r) any_alive |= alive[r];
    if (!__any_sync(0xffffffffu, any_alive)) {
      if (exhausted) break;
      continue;
    }

    // ---- trace ---------------------------------------------------------------------------------------------
    float best_t[R];
    int best_i[R];
    trace_spheres<R, STATS>(sc, sA, sB, ray, alive, best_t, best_i, n_cand);
    if (STATS) {
#pragma unroll
      for (int r = 0; r < R; ++r) if (alive[r]) n_tests += static_cast<unsigned long long>(sc.n_static + sc.n_moving);
    }

    // ---- shade: ray_color, render.cpp:112-129, one bounce ------------------------------------------------
#pragma unroll
    for (int r = 0; r < R; ++r) {
      if (!alive[r]) continue;
      ++n_rays;
      const F3 o = mk<float>(ray.ox[r], ray.oy[r], ray.oz[r]), d = mk<float>(ray.dx[r], ray.dy[r], ray.dz[r]);
      if (best_i[r] == kMiss) {
        const F3 c = sky_color(d);
        accum_add(p.accum, accum_index(p, pix[r]), tr[r] * c.x, tg[r] * c.y, tb[r] * c.z);
        alive[r] = false;
      } else if (depth[r] >= p.max_depth) {
        alive[r] = false;  // hit at depth 0 of the recursion: black
      } else {
        const HitGeom g = hit_geometry(sc, sA, sB, o, d, ray.tm[r], best_t[r], best_i[r]);
        const float4 mA = __ldg(&sc.matA[g.material]);
        const float2 mB = __ldg(&sc.matB[g.material]);
        const int kind = __float_as_int(mB.y);
        const uint4 x = philox4x32_10(make_uint4(pix[r], smp[r], 2u + static_cast<uint32_t>(depth[r]), 0u), key);
        const F3 ball = sample_octant_ball(u01(x.x), u01(x.y), u01(x.z));
        F3 dn;
        if (scatter_dir(kind, mA.w, mB.x, d, g.n, g.front, ball, u01(x.w), dn)) {
          ray.ox[r] = g.p.x; ray.oy[r] = g.p.y; ray.oz[r] = g.p.z;
          ray.dx[r] = dn.x; ray.dy[r] = dn.y; ray.dz[r] = dn.z;
          if (kind != kDielectric) { tr[r] *= mA.x; tg[r] *= mA.y; tb[r] *= mA.z; }
          ++depth[r];
        } else {
          alive[r] = false;
        }
      }
      if (!alive[r]) {
        ++n_paths;
        atomicAdd(p.accum + 4ull * accum_index(p, pix[r]) + 3, 1ull);
      }
    }
  }

  // ---- counters: one atomic per warp -------------------------------------------------------------------------
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    n_rays += __shfl_xor_sync(0xffffffffu, n_rays, off);
    n_paths += __shfl_xor_sync(0xffffffffu, n_paths, off);
    if (STATS) {
      n_tests += __shfl_xor_sync(0xffffffffu, n_tests, off);
      n_cand += __shfl_xor_sync(0xffffffffu, n_cand, off);
    }
  }
  if (lane == 0) {
    atomicAdd(p.counters + kCtrRays, n_rays);
    atomicAdd(p.counters + kCtrPaths, n_paths);
    if (STATS) {
      atomicAdd(p.counters + kCtrSphereTests, n_tests);
      atomicAdd(p.counters + kCtrCandidates, n_cand);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// K2 render kernel: BVH traversal as a resumable per-lane state machine.
//   A lane is DEAD (needs a path), TRAV (walking the tree; node, stack and closest hit live across outer iterations)
//   or DONE (traversal finished, waiting to be shaded).  The outer loop alternates
//     service phase  -- only when at least SERVICE_MIN lanes are DONE/DEAD (or nobody is traversing): big spheres in fp64,
//                       shade DONE lanes (scatter -> next ray, or terminate), refill DEAD lanes from the warp's pool
//                       (ballot + popc ranking, units fetched by lane 0 from the global counter);
//     traversal phase -- STEPS node/leaf steps for every TRAV lane.
//   A short traversal therefore never waits for the longest one in its warp, which is what held the first version of
//   this kernel at 15 of 32 active lanes per instruction.  With SMEM the nodes, leaf references, sphere tables and
//   triangles are staged into shared memory with TMA bulk copies when they fit.
// ---------------------------------------------------------------------------------------------------------
struct BvhTables {
  const float4* nodes;
  const uint32_t* leafRefs;
  const float4* sphA;
  const float4* sphB;
  const float4* tri;
};

template <bool SMEM, bool STATS, int STEPS, int SERVICE_MIN, int MINB, bool CW = false>
__global__ void __launch_bounds__(kRenderThreads, MINB) k_render_bvh(const __grid_constant__ RenderParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const DevScene& sc = p.sc;
  BvhTables tb{CW ? reinterpret_cast<const float4*>(sc.cwNodes) : sc.nodes, sc.leafRefs, sc.sphA, sc.sphB, sc.tri};
  if (SMEM) {
    // nodes, leafRefs, sphA, sphB, tri back to back (every size a multiple of 16 bytes), one mbarrier for all copies
    const uint32_t nspheres = static_cast<uint32_t>(sc.n_static + sc.n_moving);
    const uint32_t b_nodes = CW ? static_cast<uint32_t>(sc.n_cw_nodes) * 80u : static_cast<uint32_t>(sc.n_nodes) * 64u,
                   b_refs = (p.n_leaf_refs * 4u + 15u) & ~15u, b_sph = nspheres * 16u, b_tri = static_cast<uint32_t>(sc.n_tri) * 48u;
    unsigned char* d_nodes = smem_raw + 16;
    unsigned char* d_refs = d_nodes + b_nodes;
    unsigned char* d_sa = d_refs + b_refs;
    unsigned char* d_sb = d_sa + b_sph;
    unsigned char* d_tri = d_sb + b_sph;
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
    if (threadIdx.x == 0) {
      mbar_init(bar, 1);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      mbar_expect_tx(bar, b_nodes + b_refs + 2u * b_sph + b_tri);
      auto copy = [&](unsigned char* dst, const void* src, uint32_t bytes) {
        constexpr uint32_t kChunk = 32768u;
        for (uint32_t off = 0; off < bytes; off += kChunk)
          tma_bulk_g2s(dst + off, static_cast<const unsigned char*>(src) + off, min(kChunk, bytes - off), bar);
      };
      copy(d_nodes, CW ? static_cast<const void*>(sc.cwNodes) : static_cast<const void*>(sc.nodes), b_nodes);
      copy(d_refs, sc.leafRefs, b_refs); copy(d_sa, sc.sphA, b_sph); copy(d_sb, sc.sphB, b_sph);
      copy(d_tri, sc.tri, b_tri);
    }
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, 0)) {
      if (++spins > (1u << 26)) __trap();
    }
    tb.nodes = reinterpret_cast<const float4*>(d_nodes);
    tb.leafRefs = reinterpret_cast<const uint32_t*>(d_refs);
    tb.sphA = reinterpret_cast<const float4*>(d_sa);
    tb.sphB = reinterpret_cast<const float4*>(d_sb);
    tb.tri = reinterpret_cast<const float4*>(d_tri);
  }

  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t lt = (1u << lane) - 1u;
  const uint2 key = make_uint2(static_cast<uint32_t>(p.seed), static_cast<uint32_t>(p.seed >> 32));
  enum : int { DEAD = 0, TRAV = 1, DONE = 2 };
  // compressed wide BVH (CW): node group / primitive group / octant of the ray in flight, and the stack of node groups
  uint2 ngroup = make_uint2(0u, 0u), tgroup = make_uint2(0u, 0u);
  uint32_t oct_inv4 = 0u;
  uint2 cw_stack[CW ? kCwStack : 1];

  int state = DEAD;
  F3 o = mk<float>(0.f, 0.f, 0.f), d = mk<float>(0.f, 0.f, 1.f);
  float tm = 0.f, tr = 0.f, tg = 0.f, tbl = 0.f;
  uint32_t pix = 0, smp = 0;
  int depth = 0;
  // traversal state
  float idx = 0.f, idy = 0.f, idz = 0.f, odx = 0.f, ody = 0.f, odz = 0.f, qa = 1.f, qia = 1.f, best_t = kInf;
  int best_i = kMiss, node = kMiss, sp = 0;
  int stack[CW ? 1 : kBvhStack];

  uint32_t pool_next = 0, pool_end = 0, grp = 0, s0 = 0;
  bool exhausted = false;
  unsigned long long n_rays = 0, n_paths = 0, n_tests = 0, n_nodes = 0, n_tri = 0;

  auto start_traversal = [&]() {
    qa = dot(d, d); qia = fast_rcp(qa);
    idx = fast_rcp(d.x); idy = fast_rcp(d.y); idz = fast_rcp(d.z);
    best_t = kInf; best_i = kMiss; sp = 0;
    if (CW) {
      oct_inv4 = cw_oct_inv4(d);
      ngroup = make_uint2(0u, 0x80000000u); tgroup = make_uint2(0u, 0u);   // the root as the one hit child of a virtual node
      state = sc.n_cw_nodes > 0 ? TRAV : DONE;
    } else {
      odx = o.x * idx; ody = o.y * idy; odz = o.z * idz;
      node = sc.n_nodes > 0 ? 0 : kMiss;
      state = node == kMiss ? DONE : TRAV;
    }
  };

  for (;;) {
    const uint32_t trav_mask = __ballot_sync(0xffffffffu, state == TRAV);
    const int n_service = 32 - __popc(trav_mask);
    if (n_service >= (CW ? static_cast<int>(p.cw_service_min) : SERVICE_MIN) || trav_mask == 0u) {
      // ---- shade lanes whose traversal finished ------------------------------------------------------------
      if (state == DONE) {
        trace_big_spheres(sc, o, d, tm, best_t, best_i);
        ++n_rays;
        bool ended = false;
        if (best_i == kMiss) {
          const F3 c = sky_color(d);
          accum_add(p.accum, accum_index(p, pix), tr * c.x, tg * c.y, tbl * c.z);
          ended = true;
        } else if (depth >= p.max_depth) {
          ended = true;
        } else {
          const HitGeom g = hit_geometry(sc, tb.sphA, tb.sphB, o, d, tm, best_t, best_i);
          const float4 mA = __ldg(&sc.matA[g.material]);
          const float2 mB = __ldg(&sc.matB[g.material]);
          const int kind = __float_as_int(mB.y);
          const uint4 x = philox4x32_10(make_uint4(pix, smp, 2u + static_cast<uint32_t>(depth), 0u), key);
          const F3 ball = sample_octant_ball(u01(x.x), u01(x.y), u01(x.z));
          F3 dn;
          if (scatter_dir(kind, mA.w, mB.x, d, g.n, g.front, ball, u01(x.w), dn)) {
            o = g.p; d = dn;
            if (kind != kDielectric) { tr *= mA.x; tg *= mA.y; tbl *= mA.z; }
            ++depth;
            start_traversal();
          } else {
            ended = true;
          }
        }
        if (ended) {
          ++n_paths;
          atomicAdd(p.accum + 4ull * accum_index(p, pix) + 3, 1ull);
          state = DEAD;
        }
      }
      // ---- refill dead lanes from the warp's pool ------------------------------------------------------------------
#pragma unroll 1
      for (int pass = 0; pass < 2; ++pass) {
        const uint32_t need = __ballot_sync(0xffffffffu, state == DEAD);
        if (need == 0u) break;
        if (pool_next == pool_end) {
          if (exhausted) break;
          unsigned long long u = 0;
          if (lane == 0) u = atomicAdd(p.counters + kCtrWork, 1ull);
          u = __shfl_sync(0xffffffffu, u, 0);
          if (u >= p.n_units) { exhausted = true; break; }
          grp = static_cast<uint32_t>(u / p.n_chunks);
          const uint32_t chunk = static_cast<uint32_t>(u - static_cast<unsigned long long>(grp) * p.n_chunks);
          s0 = p.s_begin + chunk * p.su;
          const uint32_t ns = min(p.su, p.s_end - s0);
          pool_next = 0; pool_end = ns * kGroupPixels;
        }
        const uint32_t avail = pool_end - pool_next;
        const uint32_t rank = __popc(need & lt);
        if (state == DEAD && rank < avail) {
          const uint32_t within = pool_next + rank;
          const uint32_t lp = grp * kGroupPixels + (within & (kGroupPixels - 1u));
          const uint32_t sm = s0 + within / kGroupPixels;
          uint32_t px, i, j;
          if (local_to_global(p, lp, px, i, j)) {
            const uint4 x0 = philox4x32_10(make_uint4(px, sm, 0u, 0u), key);
            const float u = (static_cast<float>(j) + u01(x0.x)) * p.inv_wm1;
            const float v = (static_cast<float>(p.height - 1u - i) + u01(x0.y)) * p.inv_hm1;
            const float2 dk = sample_disk(u01(x0.z), u01(x0.w));
            camera_ray<float>(sc.cam, u, v, dk.x, dk.y, o, d);
            tm = fmaf(u01_low(x0), sc.cam.t1 - sc.cam.t0, sc.cam.t0);
            tr = tg = tbl = 1.0f;
            pix = px; smp = sm; depth = 0;
            start_traversal();
          }
        }
        pool_next += min(static_cast<uint32_t>(__popc(need)), avail);
      }
      if (!__any_sync(0xffffffffu, state != DEAD)) {
        if (exhausted) break;
        continue;
      }
    }

    // ---- traversal phase: STEPS steps; in each, lanes on an inner node visit it, then lanes on a leaf test its primitive(s):
    // at once in sphere scenes (holding sphere leaves back until 4..16 lanes stand on one lost 3-8 % on the cover scene), once
    // p.leaf_min lanes stand on one in scenes with triangles (measured on the 991k-triangle mesh, Mpaths/s: 1/2/3/4/5/6/8/12 lanes =
    // 3034/3094/3176/3211/3205/3163/3017/2573; suzanne 7290 -> 7345).
    if constexpr (CW) {
      // one step = (a) lanes with an unvisited inner child fetch and test that node's 8 children, then (b) lanes holding leaf
      // primitives test ONE of them -- once p.leaf_min lanes hold some or no lane has a node left to visit (the triangle test runs
      // at a handful of lanes otherwise), then (c) lanes with neither pop a node group or finish
#pragma unroll 1
      for (uint32_t step = 0; step < p.cw_steps; ++step) {
        if (state == TRAV && tgroup.y == 0u && (ngroup.y & 0xff000000u)) {
          const uint32_t child = cw_pop_child(ngroup, oct_inv4, cw_stack, sp);
          if (STATS) ++n_nodes;
          cw_intersect_node<SMEM>(reinterpret_cast<const uint4*>(tb.nodes), child, o, idx, idy, idz, oct_inv4, kTMin, best_t, ngroup, tgroup);
        }
        bool do_leaf = true;
        if (p.leaf_min > 1u) {
          const uint32_t lm = __ballot_sync(0xffffffffu, state == TRAV && tgroup.y != 0u);
          const uint32_t im = __ballot_sync(0xffffffffu, state == TRAV && tgroup.y == 0u && (ngroup.y & 0xff000000u));
          do_leaf = static_cast<uint32_t>(__popc(lm)) >= p.leaf_min || im == 0u;
        }
        if (do_leaf && state == TRAV && tgroup.y != 0u) {
          const uint32_t k = 31u - __clz(tgroup.y);
          tgroup.y &= ~(1u << k);
          cw_test_prim<SMEM, STATS>(sc, tb.tri, tgroup.x + k, o, d, qa, qia, tm, best_t, best_i, n_tests, n_tri);
        }
        if (state == TRAV && tgroup.y == 0u && (ngroup.y & 0xff000000u) == 0u) {
          if (sp > 0) ngroup = cw_stack[--sp];
          else state = DONE;
        }
      }
    } else {
    // RTW_K2_V inner-node visits per leaf phase (as in K2w: one convergence point less per extra visit).  Same box, 1080p x 64 spp,
    // (steps per traversal phase, V) on suzanne / the 991k-triangle mesh, ms per frame (scripts/r2_gpu27.sh): (4, 1) 16.84 / 35.51
    // (4, 2) 16.13 / 34.28  (4, 4) 16.37 / 35.37  (6, 2) 15.95 / 34.02  (6, 3) 16.15 / 34.16  (8, 2) 15.98 / 34.06  (8, 4) 16.29 / 34.30
#ifndef RTW_K2_V
#define RTW_K2_V 2
#endif
    static_assert(STEPS % RTW_K2_V == 0, "steps per traversal phase must be a multiple of the visits per leaf phase");
#pragma unroll 1
    for (int step = 0; step < STEPS / RTW_K2_V; ++step) {
#pragma unroll
      for (int vv = 0; vv < RTW_K2_V; ++vv) {
        if (state == TRAV && node >= 0) {
          if (STATS) ++n_nodes;
          float4 q0, q1, q2, q3;
          load_node<SMEM>(tb.nodes, node, q0, q1, q2, q3);
          float ln, lf, rn, rf;
          node_slabs(q0, q1, q2, idx, idy, idz, odx, ody, odz, best_t, ln, lf, rn, rf);
          const bool hl = ln <= lf, hr = rn <= rf;
          const int left = __float_as_int(q3.x), right = __float_as_int(q3.y);
          if (hl && hr) {
            const bool lfirst = ln <= rn;
            node = lfirst ? left : right;
            if (sp < kBvhStack) stack[sp++] = lfirst ? right : left;
          } else if (hl) {
            node = left;
          } else if (hr) {
            node = right;
          } else if (sp > 0) {
            node = stack[--sp];
          } else {
            node = kMiss; state = DONE;
          }
        }
      }
      bool do_leaf = true;
      if (!SMEM && p.leaf_min > 1u) {  // (tables in L1/L2 only) hold leaves back until enough lanes stand on one, or no lane has an inner node left to visit
        const uint32_t lm = __ballot_sync(0xffffffffu, state == TRAV && node < 0), im = __ballot_sync(0xffffffffu, state == TRAV && node >= 0);
        do_leaf = static_cast<uint32_t>(__popc(lm)) >= p.leaf_min || im == 0u;
      }
      if (do_leaf && state == TRAV && node < 0) {
        const uint32_t v = static_cast<uint32_t>(~node);
        const uint32_t first = v >> 5, cnt = sc.leaf_direct ? 1u : (v & 31u);
        for (uint32_t k = 0; k < cnt; ++k) {
          const uint32_t ref = sc.leaf_direct ? v : ld1<SMEM>(tb.leafRefs, static_cast<int>(first + k));
          const int i = static_cast<int>(ref & 0x1fffffffu);
          if (ref >> 30) {
            if (STATS) ++n_tri;
            const float4 q0 = ld4<SMEM>(tb.tri, 3 * i), q1 = ld4<SMEM>(tb.tri, 3 * i + 1), q2 = ld4<SMEM>(tb.tri, 3 * i + 2);
            const float t = triangle_hit_fast(o, d, mk<float>(q0.x, q0.y, q0.z), mk<float>(q1.x, q1.y, q1.z), mk<float>(q2.x, q2.y, q2.z),
                                                mk<float>(q0.w, q1.w, q2.w), kTMin, best_t);
            if (t >= 0.0f) { best_t = t; best_i = kHitTri | i; }
          } else {
            if (STATS) ++n_tests;
            const float4 A = ld4<SMEM>(tb.sphA, i), B = ld4<SMEM>(tb.sphB, i);
            const float t = sphere_hit_fast(o, d, qa, qia, mk<float>(fmaf(tm, B.x, A.x), fmaf(tm, B.y, A.y), fmaf(tm, B.z, A.z)), B.w, kTMin, best_t);
            if (t >= 0.0f) { best_t = t; best_i = i; }
          }
        }
        if (sp > 0) node = stack[--sp];
        else { node = kMiss; state = DONE; }
      }
    }
    }   // binary tree
  }

#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    n_rays += __shfl_xor_sync(0xffffffffu, n_rays, off);
    n_paths += __shfl_xor_sync(0xffffffffu, n_paths, off);
    if (STATS) {
      n_tests += __shfl_xor_sync(0xffffffffu, n_tests, off);
      n_nodes += __shfl_xor_sync(0xffffffffu, n_nodes, off);
      n_tri += __shfl_xor_sync(0xffffffffu, n_tri, off);
    }
  }
  if (lane == 0) {
    atomicAdd(p.counters + kCtrRays, n_rays);
    atomicAdd(p.counters + kCtrPaths, n_paths);
    if (STATS) {
      atomicAdd(p.counters + kCtrSphereTests, n_tests);
      atomicAdd(p.counters + kCtrNodes, n_nodes);
      atomicAdd(p.counters + kCtrTriTests, n_tri);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// K2w render kernel: the same BVH traversal, with the paths of a warp kept as RECORDS in shared memory instead of in
// the lanes' registers ("wavefront per warp").
//   Every warp owns P path records (ray, closest hit so far, throughput, pixel, sample, depth) and three index stacks:
//     ready -- rays waiting to be traversed          hit -- traversal found a hit: scatter next
//     miss  -- traversal missed / path ended / record is fresh: add the sky term, start a new path in the record
//   A lane only ever holds the traversal state of ONE ray.  When its traversal finishes it files the record under hit or
//   miss and takes the next ready ray, so the traversal phase runs with (nearly) all lanes busy.  Shading runs as BATCHES
//   of up to 32 records of ONE kind, lane l shading the l-th record of the stack: Philox, sampling, the big-sphere test of
//   the new ray and the record traffic are then convergent across the warp.  With P >= 32 + 2 * BATCH a full batch of one
//   kind is always available when the ready stack runs dry.
//   The big spheres (fp64) are tested when a ray is CREATED (inside the batch), and the result seeds best_t of the tree
//   walk, which also prunes it.
// ---------------------------------------------------------------------------------------------------------
constexpr int kFresh = -2;   // record holds no path yet
constexpr int kEnded = -3;   // path ended without a sky term (depth cut / absorbed)

template <bool SMEM, bool STATS, int NW, int P, int STEPS, int BATCH, int MINB = 1>
__global__ void __launch_bounds__(NW * 32, MINB) k_render_wf(const __grid_constant__ RenderParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  // LEAN (32 warps per SM: 64 registers): the lane keeps only what the node visits need (1/d, o/d, closest hit); origin, direction
  // and time of its ray are read back from the record for the ~1 leaf test per ray.  More warps is what this kernel responds to
  // (20 / 24 / 28 / 32 warps: 39.5 / 38.0 / 35.5 / 34.7 ms), not fewer instructions (rtw_internal.h)
  constexpr bool LEAN = NW > 28;
  const DevScene& sc = p.sc;
  BvhTables tb{sc.nodes, sc.leafRefs, sc.sphA, sc.sphB, sc.tri};
  if (SMEM) {
    // shared-memory offsets of the tables come precomputed from the host (p.so): the compiler re-materialises these
    // addresses inside the loops instead of holding them in registers, so they must be one constant-bank load away
    const SmemLayout& so = p.so;
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
    if (threadIdx.x == 0) {
      mbar_init(bar, 1);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      mbar_expect_tx(bar, so.b_nodes + so.b_refs + 2u * so.b_sph + so.b_tri);
      auto copy = [&](unsigned char* dst, const void* src, uint32_t bytes) {
        constexpr uint32_t kChunk = 32768u;
        for (uint32_t off = 0; off < bytes; off += kChunk)
          tma_bulk_g2s(dst + off, static_cast<const unsigned char*>(src) + off, min(kChunk, bytes - off), bar);
      };
      copy(smem_raw + so.nodes, sc.nodes, so.b_nodes); copy(smem_raw + so.refs, sc.leafRefs, so.b_refs);
      copy(smem_raw + so.sa, sc.sphA, so.b_sph); copy(smem_raw + so.sb, sc.sphB, so.b_sph);
      copy(smem_raw + so.tri, sc.tri, so.b_tri);
    }
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, 0)) {
      if (++spins > (1u << 26)) __trap();
    }
    tb.nodes = reinterpret_cast<const float4*>(smem_raw + so.nodes);
    tb.leafRefs = reinterpret_cast<const uint32_t*>(smem_raw + so.refs);
    tb.sphA = reinterpret_cast<const float4*>(smem_raw + so.sa);
    tb.sphB = reinterpret_cast<const float4*>(smem_raw + so.sb);
    tb.tri = reinterpret_cast<const float4*>(smem_raw + so.tri);
  }
  unsigned char* wbase = smem_raw + p.so.records;

  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const uint32_t lt = (1u << lane) - 1u;
  const uint2 key = make_uint2(static_cast<uint32_t>(p.seed), static_cast<uint32_t>(p.seed >> 32));

  // ---- this warp's records and index stacks ------------------------------------------------------------------------
  constexpr uint32_t kWarpBytes = wf_warp_bytes(P);
  unsigned char* wb = wbase + warp * kWarpBytes;
  float4* recO = reinterpret_cast<float4*>(wb);      // (o.xyz, time)
  float4* recD = recO + P;                           // (d.xyz, best_t)
  float4* recT = recD + P;                           // (throughput.rgb, best_i as int bits)
  uint2* recI = reinterpret_cast<uint2*>(recT + P);  // (pixel, sample)
  int* recZ = reinterpret_cast<int*>(recI + P);      // depth
  uint8_t* q_ready = reinterpret_cast<uint8_t*>(recZ + P);
  uint8_t* q_hit = q_ready + P;
  uint8_t* q_miss = q_hit + P;
  for (uint32_t i = lane; i < static_cast<uint32_t>(P); i += 32u) {
    recT[i] = make_float4(0.f, 0.f, 0.f, __int_as_float(kFresh));
    q_miss[i] = static_cast<uint8_t>(i);
  }
  __syncwarp();
  uint32_t n_ready = 0, n_hit = 0, n_miss = P;   // warp-uniform

  enum : int { IDLE = 0, TRAV = 1, FIN = 2 };
  int state = IDLE;
  uint32_t rec = 0;
  F3 o = mk<float>(0.f, 0.f, 0.f), d = mk<float>(0.f, 0.f, 1.f);
  float tm = 0.f;
  float idx = 0.f, idy = 0.f, idz = 0.f, odx = 0.f, ody = 0.f, odz = 0.f, qa = 1.f, qia = 1.f, best_t = kInf;
  int best_i = kMiss, node = kMiss, sp = 0;
  int stack[kBvhStack];

  uint32_t pool_next = 0, pool_end = 0, grp = 0, s0 = 0;
  bool exhausted = false;
  uint32_t n_rays = 0, n_paths = 0, n_tests = 0, n_nodes = 0, n_tri = 0;

  for (;;) {
    // ---- (1) lanes whose traversal finished file their record under hit / miss --------------------------------------
    const uint32_t fin = __ballot_sync(0xffffffffu, state == FIN);
    if (fin != 0u) {
      const bool is_fin = state == FIN;
      const bool is_hit = is_fin && best_i != kMiss;
      const uint32_t mh = __ballot_sync(0xffffffffu, is_hit), mm = fin & ~mh;
      if (is_fin) {
        recD[rec].w = best_t;
        recT[rec].w = __int_as_float(best_i);
        if (is_hit) q_hit[n_hit + __popc(mh & lt)] = static_cast<uint8_t>(rec);
        else q_miss[n_miss + __popc(mm & lt)] = static_cast<uint8_t>(rec);
        ++n_rays;
        state = IDLE;
      }
      n_hit += __popc(mh); n_miss += __popc(mm);
      __syncwarp();
    }
    // ---- (2) idle lanes take ready rays ---------------------------------------------------------------------------------
    uint32_t idle = __ballot_sync(0xffffffffu, state == IDLE);
    if (idle != 0u && n_ready != 0u) {
      const uint32_t rank = __popc(idle & lt);
      if (state == IDLE && rank < n_ready) {
        rec = q_ready[n_ready - 1u - rank];
        const float4 a = recO[rec], b = recD[rec];
        best_t = b.w;
        best_i = __float_as_int(recT[rec].w);
        idx = fast_rcp(b.x); idy = fast_rcp(b.y); idz = fast_rcp(b.z);
        odx = a.x * idx; ody = a.y * idy; odz = a.z * idz;
        if (!LEAN) {
          o = mk<float>(a.x, a.y, a.z); tm = a.w;
          d = mk<float>(b.x, b.y, b.z);
          qa = dot(d, d); qia = fast_rcp(qa);
        }
        sp = 0;
        node = sc.n_nodes > 0 ? 0 : kMiss;
        state = node == kMiss ? FIN : TRAV;
      }
      const uint32_t took = min(static_cast<uint32_t>(__popc(idle)), n_ready);
      n_ready -= took;
      idle = __ballot_sync(0xffffffffu, state == IDLE);
      __syncwarp();
    }
    // ---- (3) shading batches ----------------------------------------------------------------------------------------------
    const uint32_t n_idle = __popc(idle);
    const bool starving = n_ready == 0u && n_idle >= 8u;  // cannot happen before the work runs out when P >= 32 + 2 * BATCH
    const bool run_hit = n_hit >= static_cast<uint32_t>(BATCH) || (starving && n_hit > 0u && n_hit >= n_miss);
    const bool run_miss = !run_hit && (n_miss >= static_cast<uint32_t>(BATCH) || (starving && n_miss > 0u));
    if (run_hit) {
      // scatter: hit geometry, material, next ray (common-model.cpp:13-62), big-sphere pre-test of the new ray
      const uint32_t take = min(32u, n_hit);
      const bool act = lane < take;
      uint32_t r2 = 0;
      if (act) r2 = q_hit[n_hit - 1u - lane];
      n_hit -= take;
      __syncwarp();
      bool to_ready = false, to_miss = false;
      if (act) {
        const float4 a = recO[r2], b = recD[r2], c = recT[r2];
        const int dep = recZ[r2];
        const F3 o2 = mk<float>(a.x, a.y, a.z), d2 = mk<float>(b.x, b.y, b.z);
        if (dep >= p.max_depth) {
          recT[r2].w = __int_as_float(kEnded);   // hit at depth 0 of the recursion: black
          to_miss = true;
        } else {
          const HitGeom g = hit_geometry(sc, tb.sphA, tb.sphB, o2, d2, a.w, b.w, __float_as_int(c.w));
          const float4 mA = __ldg(&sc.matA[g.material]);
          const float2 mB = __ldg(&sc.matB[g.material]);
          const int kind = __float_as_int(mB.y);
          const uint2 ps = recI[r2];
          const uint4 x = philox4x32_10(make_uint4(ps.x, ps.y, 2u + static_cast<uint32_t>(dep), 0u), key);
          const F3 ball = sample_octant_ball(u01(x.x), u01(x.y), u01(x.z));
          F3 dn;
          if (scatter_dir(kind, mA.w, mB.x, d2, g.n, g.front, ball, u01(x.w), dn)) {
            float bt = kInf; int bi = kMiss;
            trace_big_spheres(sc, g.p, dn, a.w, bt, bi);
            recO[r2] = make_float4(g.p.x, g.p.y, g.p.z, a.w);
            recD[r2] = make_float4(dn.x, dn.y, dn.z, bt);
            const bool die = kind == kDielectric;
            recT[r2] = make_float4(die ? c.x : c.x * mA.x, die ? c.y : c.y * mA.y, die ? c.z : c.z * mA.z, __int_as_float(bi));
            recZ[r2] = dep + 1;
            to_ready = true;
          } else {
            recT[r2].w = __int_as_float(kEnded);
            to_miss = true;
          }
        }
      }
      const uint32_t mr = __ballot_sync(0xffffffffu, to_ready), mm = __ballot_sync(0xffffffffu, to_miss);
      if (to_ready) q_ready[n_ready + __popc(mr & lt)] = static_cast<uint8_t>(r2);
      if (to_miss) q_miss[n_miss + __popc(mm & lt)] = static_cast<uint8_t>(r2);
      n_ready += __popc(mr); n_miss += __popc(mm);
      __syncwarp();
      continue;
    }
    if (run_miss) {
      // end of path: sky term (render.cpp:125-128), then a new path in the same record (render.cpp:158-160)
      const uint32_t take = min(32u, n_miss);
      const bool act = lane < take;
      uint32_t r2 = 0;
      if (act) r2 = q_miss[n_miss - 1u - lane];
      n_miss -= take;
      __syncwarp();
      if (act) {
        const float4 c = recT[r2];
        const int code = __float_as_int(c.w);
        if (code != kFresh) {
          const uint32_t ax = accum_index(p, recI[r2].x);
          if (code == kMiss) {
            const float4 b = recD[r2];
            const F3 s = sky_color(mk<float>(b.x, b.y, b.z));
            accum_add(p.accum, ax, c.x * s.x, c.y * s.y, c.z * s.z);
          }
          ++n_paths;
          atomicAdd(p.accum + 4ull * ax + 3, 1ull);
        }
      }
      bool got = false, retry = false;
#pragma unroll 1
      for (int pass = 0; pass < 2; ++pass) {
        const uint32_t need = __ballot_sync(0xffffffffu, act && !got);
        if (need == 0u) break;
        if (pool_next == pool_end) {
          if (exhausted) break;
          unsigned long long u = 0;
          if (lane == 0) u = atomicAdd(p.counters + kCtrWork, 1ull);
          u = __shfl_sync(0xffffffffu, u, 0);
          if (u >= p.n_units) { exhausted = true; break; }
          grp = static_cast<uint32_t>(u / p.n_chunks);
          const uint32_t chunk = static_cast<uint32_t>(u - static_cast<unsigned long long>(grp) * p.n_chunks);
          s0 = p.s_begin + chunk * p.su;
          const uint32_t ns = min(p.su, p.s_end - s0);
          pool_next = 0; pool_end = ns * kGroupPixels;
        }
        const uint32_t avail = pool_end - pool_next;
        const uint32_t rank = __popc(need & lt);
        if (act && !got && rank < avail) {
          const uint32_t within = pool_next + rank;
          const uint32_t lp = grp * kGroupPixels + (within & (kGroupPixels - 1u));
          const uint32_t sm = s0 + within / kGroupPixels;
          uint32_t px, i, j;
          if (local_to_global(p, lp, px, i, j)) {
            const uint4 x0 = philox4x32_10(make_uint4(px, sm, 0u, 0u), key);
            const float u = (static_cast<float>(j) + u01(x0.x)) * p.inv_wm1;
            const float v = (static_cast<float>(p.height - 1u - i) + u01(x0.y)) * p.inv_hm1;
            const float2 dk = sample_disk(u01(x0.z), u01(x0.w));
            F3 o2, d2;
            camera_ray<float>(sc.cam, u, v, dk.x, dk.y, o2, d2);
            const float t2 = fmaf(u01_low(x0), sc.cam.t1 - sc.cam.t0, sc.cam.t0);
            float bt = kInf; int bi = kMiss;
            trace_big_spheres(sc, o2, d2, t2, bt, bi);
            recO[r2] = make_float4(o2.x, o2.y, o2.z, t2);
            recD[r2] = make_float4(d2.x, d2.y, d2.z, bt);
            recT[r2] = make_float4(1.0f, 1.0f, 1.0f, __int_as_float(bi));
            recI[r2] = make_uint2(px, sm);
            recZ[r2] = 0;
            got = true;
          }
        }
        pool_next += min(static_cast<uint32_t>(__popc(need)), avail);
      }
      // a record that found no pixel this time (tail of the last group) is tried again unless the work is exhausted
      if (act && !got && !exhausted) { recT[r2].w = __int_as_float(kFresh); retry = true; }
      const uint32_t mr = __ballot_sync(0xffffffffu, got), mm = __ballot_sync(0xffffffffu, retry);
      if (got) q_ready[n_ready + __popc(mr & lt)] = static_cast<uint8_t>(r2);
      if (retry) q_miss[n_miss + __popc(mm & lt)] = static_cast<uint8_t>(r2);
      n_ready += __popc(mr); n_miss += __popc(mm);
      __syncwarp();
      continue;
    }
    if (idle == 0xffffffffu && n_ready == 0u) break;   // nothing in flight, nothing queued (hit/miss would have run)

    // ---- (4) traversal: STEPS node-or-leaf steps for every lane holding a ray --------------------------------------------
    // (Measured and not kept, profiles/r02_ffma2.txt: a lane parking its leaf and walking on, the warp's parked leaves tested together
    // every 2 / 4 / 8 / 16 steps -- 39.2 / 39.2 / 41.0 / 44.1 ms against 34.7, node visits per ray 11.45 -> 11.52 .. 12.09.)
    // (A lane that is not traversing always holds node == kMiss, and leaf codes are < -1, so `node >= 0` / `node < kMiss` alone would
    // do as conditions: measured 33.54 against 33.22 ms -- the extra compare buys the compiler a better branch layout.  Not used.)
    auto visit = [&]() {
      if (state == TRAV && node >= 0) {
        if (STATS) ++n_nodes;
        float4 q0, q1, q2, q3;
        load_node<SMEM>(tb.nodes, node, q0, q1, q2, q3);
        float ln, lf, rn, rf;
        node_slabs(q0, q1, q2, idx, idy, idz, odx, ody, odz, best_t, ln, lf, rn, rf);
        const bool hl = ln <= lf, hr = rn <= rf;
        const int left = __float_as_int(q3.x), right = __float_as_int(q3.y);
        // (a branch-free form of this -- predicated push / pop, selects -- was measured: 36.54 against 35.48 ms, not kept)
        // (so were a two-way branch -- some child hit / none -- with a conditional push: 34.7 against 33.2 ms, and dropping the stack
        // guard below, which cannot fire because rtw_scene_upload rejects deeper trees: 33.3 against 33.2 ms)
        if (hl && hr) {
          const bool lfirst = ln <= rn;
          node = lfirst ? left : right;
          if (sp < kBvhStack) stack[sp++] = lfirst ? right : left;
        } else if (hl) {
          node = left;
        } else if (hr) {
          node = right;
        } else if (sp > 0) {
          node = stack[--sp];
        } else {
          node = kMiss; state = FIN;
        }
      }
    };
    auto leaf = [&]() {
      if (state == TRAV && node < 0) {
        const uint32_t v = static_cast<uint32_t>(~node);
        if (LEAN) {
          const float4 a = recO[rec], b = recD[rec];
          o = mk<float>(a.x, a.y, a.z); tm = a.w;
          d = mk<float>(b.x, b.y, b.z);
          qa = dot(d, d); qia = fast_rcp(qa);
        }
        auto test_ref = [&](uint32_t ref) {
          const int i = static_cast<int>(ref & 0x1fffffffu);
          if (ref >> 30) {
            if (STATS) ++n_tri;
            const float4 q0 = ld4<SMEM>(tb.tri, 3 * i), q1 = ld4<SMEM>(tb.tri, 3 * i + 1), q2 = ld4<SMEM>(tb.tri, 3 * i + 2);
            const float t = triangle_hit_fast(o, d, mk<float>(q0.x, q0.y, q0.z), mk<float>(q1.x, q1.y, q1.z), mk<float>(q2.x, q2.y, q2.z),
                                                mk<float>(q0.w, q1.w, q2.w), kTMin, best_t);
            if (t >= 0.0f) { best_t = t; best_i = kHitTri | i; }
          } else {
            if (STATS) ++n_tests;
            const float4 A = ld4<SMEM>(tb.sphA, i), B = ld4<SMEM>(tb.sphB, i);
            const float t = sphere_hit_fast(o, d, qa, qia, mk<float>(fmaf(tm, B.x, A.x), fmaf(tm, B.y, A.y), fmaf(tm, B.z, A.z)), B.w, kTMin, best_t);
            if (t >= 0.0f) { best_t = t; best_i = i; }
          }
        };
        test_ref(v);   // single-primitive leaves only: the child code is the primitive reference (launch_render checks leaf_direct)
        if (sp > 0) node = stack[--sp];
        else { node = kMiss; state = FIN; }
      }
    };
    // Loop shape: RTW_WF_V inner-node visits, then ONE leaf check, RTW_WF_R times per phase.  Checking for a leaf after every visit
    // costs a convergence point (BSSY / two ISETP / BRA / BSYNC at 32 lanes) per visit for a test that 1 visit in 12 needs; a lane
    // that reaches its leaf early idles for at most V - 1 visits.  Cover 1080p x 128 spp, 32 warps (scripts/r2_gpu24.sh, r2_gpu25.sh):
    // (V, R) = (1, 16) 34.69 ms  (2, 6) 35.00  (2, 8) 33.84  (2, 10) 33.78  (3, 5) 33.67  (3, 6) 33.27  (3, 7) 33.25  (4, 4) 33.52
    // (4, 5) 33.22  (4, 6) 33.30  (5, 4) 33.17  (6, 3) 34.00; the same step unrolled twice WITH both leaf checks: 36.1 (spills).
#ifndef RTW_WF_V
#define RTW_WF_V 4
#define RTW_WF_R 5
#endif
    static_assert(STEPS == 16, "the loop shape below was tuned together with 16-step phases; STEPS only names the tuning generation");
#pragma unroll 1
    for (int step = 0; step < RTW_WF_R; ++step) {
#pragma unroll
      for (int v = 0; v < RTW_WF_V; ++v) visit();
      leaf();
    }
  }

  unsigned long long c_rays = n_rays, c_paths = n_paths, c_tests = n_tests, c_nodes = n_nodes, c_tri = n_tri;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    c_rays += __shfl_xor_sync(0xffffffffu, c_rays, off);
    c_paths += __shfl_xor_sync(0xffffffffu, c_paths, off);
    if (STATS) {
      c_tests += __shfl_xor_sync(0xffffffffu, c_tests, off);
      c_nodes += __shfl_xor_sync(0xffffffffu, c_nodes, off);
      c_tri += __shfl_xor_sync(0xffffffffu, c_tri, off);
    }
  }
  if (lane == 0) {
    atomicAdd(p.counters + kCtrRays, c_rays);
    atomicAdd(p.counters + kCtrPaths, c_paths);
    if (STATS) {
      atomicAdd(p.counters + kCtrSphereTests, c_tests);
      atomicAdd(p.counters + kCtrNodes, c_nodes);
      atomicAdd(p.counters + kCtrTriTests, c_tri);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// K3: deterministic primary hits
// ---------------------------------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(kRenderThreads) k_primary_f32(const __grid_constant__ PrimaryParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const DevScene& sc = p.sc;
  const float4* sA = sc.sphA;
  const float4* sB = sc.sphB;
  if (MODE == 0) {
    const int n = sc.n_static + sc.n_moving;
    float4* a = reinterpret_cast<float4*>(smem_raw + 16);
    float4* b = a + n + 1;
    stage_spheres(sc, a, b, reinterpret_cast<uint64_t*>(smem_raw));
    sA = a; sB = b;
  }
  const uint32_t px = blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = px < p.npix;
  const uint32_t pc = valid ? px : 0u;
  const uint32_t i = pc / p.width, j = pc - i * p.width;
  const float u = (static_cast<float>(j) + 0.5f) / static_cast<float>(p.width - 1);
  const float v = (static_cast<float>(p.height - 1u - i) + 0.5f) / static_cast<float>(p.height - 1);
  F3 o, d;
  camera_ray<float>(sc.cam, u, v, 0.0f, 0.0f, o, d);
  const float tm = p.time;
  float best_t;
  int best_i;
  unsigned long long c0 = 0, c1 = 0, c2 = 0;
  if (MODE == 0) {
    RaySet<1> ray;
    ray.ox[0] = o.x; ray.oy[0] = o.y; ray.oz[0] = o.z; ray.dx[0] = d.x; ray.dy[0] = d.y; ray.dz[0] = d.z; ray.tm[0] = tm;
    const bool alive[1] = {true};
    float bt[1]; int bi[1];
    trace_spheres<1, false>(sc, sA, sB, ray, alive, bt, bi, c0);
    best_t = bt[0]; best_i = bi[0];
  } else if (sc.n_cw_nodes > 0) {
    trace_cw<false>(sc, o, d, tm, best_t, best_i, c0, c1, c2);
  } else {
    trace_bvh<false>(sc, o, d, tm, best_t, best_i, c0, c1, c2);
  }
  if (!valid) return;
  if (best_i == kMiss) {
    p.prim_id[px] = -1; p.t[px] = 0.0; p.normal[3 * px] = p.normal[3 * px + 1] = p.normal[3 * px + 2] = 0.0; p.front[px] = 0;
  } else {
    const HitGeom g = hit_geometry(sc, sA, sB, o, d, tm, best_t, best_i);
    p.prim_id[px] = g.prim_id; p.t[px] = g.t;
    p.normal[3 * px] = g.n.x; p.normal[3 * px + 1] = g.n.y; p.normal[3 * px + 2] = g.n.z;
    p.front[px] = g.front ? 1 : 0;
  }
}

struct CamD {
  double origin[3], lower_left[3], horizontal[3], vertical[3], u[3], v[3];
  double lens_radius, t0, t1;
};

// Reference formulas in double, brute force in insertion order, "later primitive wins exact ties" accept rule
// (render.cpp:57-64): sphere_hit_helper common-model.cpp:64-91, MovingSphere::center oo-primitives.h:64-66,
// Triangle::hit common-model.cpp:103-125.
__global__ void __launch_bounds__(256) k_primary_f64(const rtw_primitive* __restrict__ prims, int nprims, rtw_camera cam, uint32_t width,
                                                     uint32_t height, double time, int32_t* prim_id, double* tout, double* normal,
                                                     uint8_t* front) {
  const uint32_t px = blockIdx.x * blockDim.x + threadIdx.x;
  if (px >= width * height) return;
  const uint32_t i = px / width, j = px - i * width;
  const double u = (j + 0.5) / (width - 1.0), v = ((height - 1u - i) + 0.5) / (height - 1.0);
  D3 o, d;
  camera_ray<double>(cam, u, v, 0.0, 0.0, o, d);
  double upper = __longlong_as_double(0x7ff0000000000000ll);
  int best = -1; D3 bn = mk<double>(0, 0, 0); bool bfront = false;
  const double tmin = 0.001;
  for (int k = 0; k < nprims; ++k) {
    const rtw_primitive& P = prims[k];
    if (P.kind == RTW_TRIANGLE) {
      const D3 A = mk<double>(P.a[0], P.a[1], P.a[2]), B = mk<double>(P.b[0], P.b[1], P.b[2]), C = mk<double>(P.c[0], P.c[1], P.c[2]);
      const D3 e1 = B - A, e2 = C - A, n = cross(e1, e2);
      const double t = triangle_hit<double>(o, d, A, e1, e2, n, tmin, upper);
      if (t >= 0.0) { upper = t; best = k; bn = n; bfront = true; }
    } else {
      D3 c = mk<double>(P.a[0], P.a[1], P.a[2]);
      if (P.kind == RTW_MOVING_SPHERE) c = c + time * (mk<double>(P.b[0], P.b[1], P.b[2]) - c);
      const D3 oc = o - c;
      const double a = dot(d, d), h = dot(oc, d), cc = dot(oc, oc) - P.radius * P.radius;
      const double disc = h * h - a * cc;
      if (disc < 0.0) continue;
      double root = (-h - sqrt(disc)) / a;
      if (root < tmin || root > upper) {
        root = (-h + sqrt(disc)) / a;
        if (root < tmin || root > upper) continue;
      }
      const D3 hp = o + d * root;
      D3 n = normalize(hp - c);
      const bool ff = (dot(d, n) < 0.0) != (P.radius < 0.0);
      upper = root; best = k; bn = ff ? n : -n; bfront = ff;
    }
  }
  prim_id[px] = best; tout[px] = best >= 0 ? upper : 0.0;
  normal[3 * px] = bn.x; normal[3 * px + 1] = bn.y; normal[3 * px + 2] = bn.z;
  front[px] = bfront ? 1 : 0;
}

// ---------------------------------------------------------------------------------------------------------
// Finalize (render.cpp:11-20, 176-186)
// ---------------------------------------------------------------------------------------------------------
// Row-tile split, after the gather: [count][local_rows][width] packed pixels -> [height][width]
__global__ void k_untile(const longlong4* __restrict__ gathered, longlong4* __restrict__ full, uint32_t width, uint32_t height, uint32_t tile_rows,
                         uint32_t count, uint32_t local_rows) {
  const uint32_t gp = blockIdx.x * blockDim.x + threadIdx.x;
  if (gp >= width * height) return;
  const uint32_t i = gp / width, j = gp - i * width;
  const uint32_t t = i / tile_rows, w = i - t * tile_rows;
  const uint32_t owner = t % count, lrow = (t / count) * tile_rows + w;
  full[gp] = gathered[(static_cast<size_t>(owner) * local_rows + lrow) * width + j];
}
__global__ void k_accum_to_float(const long long* __restrict__ fx, float4* __restrict__ out, long long npix) {
  const long long k = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (k >= npix) return;
  const double s = 1.0 / 4294967296.0;
  out[k] = make_float4(static_cast<float>(fx[4 * k] * s), static_cast<float>(fx[4 * k + 1] * s), static_cast<float>(fx[4 * k + 2] * s),
                       static_cast<float>(fx[4 * k + 3]));
}
__global__ void k_finalize_rgb8(const float4* __restrict__ acc, uint8_t* __restrict__ rgb, long long npix, double spp) {
  // write_color (render.cpp:11-20) in double, exactly the host formula: c = sqrt(sum / spp); int(256 * clamp(c, 0, 0.999))
  const long long k = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (k >= npix) return;
  const float4 a = acc[k];
  const double c[3] = {sqrt(static_cast<double>(a.x) / spp), sqrt(static_cast<double>(a.y) / spp), sqrt(static_cast<double>(a.z) / spp)};
#pragma unroll
  for (int ch = 0; ch < 3; ++ch) {
    const double cl = c[ch] < 0.0 ? 0.0 : (c[ch] > 0.999 ? 0.999 : c[ch]);
    rgb[3 * k + ch] = static_cast<uint8_t>(static_cast<int>(256 * cl));
  }
}

// write_color straight from the exact int64 fixed-point sums (no float rounding of the sum in between): what the drop-in prints
__global__ void k_finalize_rgb8_fx(const long long* __restrict__ fx, uint8_t* __restrict__ rgb, long long npix, double spp) {
  const long long k = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (k >= npix) return;
#pragma unroll
  for (int ch = 0; ch < 3; ++ch) {
    // sum * 2^-32 is exact in double (sums stay below 2^53 units up to 2 M samples per pixel); then c / spp and sqrt as render.cpp:14
    const double c = sqrt(static_cast<double>(fx[4 * k + ch]) * (1.0 / 4294967296.0) / spp);
    const double cl = c < 0.0 ? 0.0 : (c > 0.999 ? 0.999 : c);
    rgb[3 * k + ch] = static_cast<uint8_t>(static_cast<int>(256 * cl));
  }
}

// ---------------------------------------------------------------------------------------------------------
// Unit hooks
// ---------------------------------------------------------------------------------------------------------
__global__ void k_debug_scatter(long long n, const int* kind, const float* fuzz, const float* ior, const float* dir_in, const float* normal,
                                const uint8_t* front, const float* ball, const float* coin, float* out_dir, float* out_att,
                                const float* albedo, uint8_t* scattered) {
  const long long k = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (k >= n) return;
  F3 dn = mk<float>(0.f, 0.f, 0.f);
  const bool ok = scatter_dir(kind[k], fuzz[k], ior[k], mk<float>(dir_in[3 * k], dir_in[3 * k + 1], dir_in[3 * k + 2]),
                              mk<float>(normal[3 * k], normal[3 * k + 1], normal[3 * k + 2]), front[k] != 0,
                              mk<float>(ball[3 * k], ball[3 * k + 1], ball[3 * k + 2]), coin[k], dn);
  scattered[k] = ok ? 1 : 0;
  out_dir[3 * k] = dn.x; out_dir[3 * k + 1] = dn.y; out_dir[3 * k + 2] = dn.z;
  const bool die = kind[k] == kDielectric;
  for (int c = 0; c < 3; ++c) out_att[3 * k + c] = die ? 1.0f : albedo[3 * k + c];
}
__global__ void k_debug_samples(long long n, uint2 key, float* ball, float* disk, float* uni) {
  const long long k = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (k >= n) return;
  const uint4 x = philox4x32_10(make_uint4(static_cast<uint32_t>(k), 0u, 0u, 0u), key);
  const uint4 y = philox4x32_10(make_uint4(static_cast<uint32_t>(k), 0u, 2u, 0u), key);
  const F3 b = sample_octant_ball(u01(y.x), u01(y.y), u01(y.z));
  const float2 dk = sample_disk(u01(x.z), u01(x.w));
  ball[3 * k] = b.x; ball[3 * k + 1] = b.y; ball[3 * k + 2] = b.z;
  disk[2 * k] = dk.x; disk[2 * k + 1] = dk.y;
  uni[4 * k] = u01(x.x); uni[4 * k + 1] = u01(x.y); uni[4 * k + 2] = u01(x.z); uni[4 * k + 3] = u01(x.w);
}

// 8 independent FMA chains per thread: the sustained FP32 FMA rate that bounds K1.  PACKED: the chains are FFMA2 (fma.rn.f32x2, two
// FMAs per lane and instruction, one instruction every other cycle): the same pipe, but the issue slots no longer limit it, and it
// measures ~4 % above the scalar form (scripts/ffma2_probe.cu: 74.2 against 71.2 TFLOP/s).  rtw_fp32_peak reports the higher one.
template <bool PACKED>
__global__ void __launch_bounds__(256) k_ffma_peak(float* out, int iters, float a, float b) {
  if (PACKED) {
    uint64_t x[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] = pk2(static_cast<float>(threadIdx.x + j), static_cast<float>(threadIdx.x + j) + 0.5f);
    const uint64_t a2 = pk2(a, a), b2 = pk2(b, b);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int k = 0; k < 16; ++k) {
#pragma unroll
        for (int j = 0; j < 8; ++j) x[j] = fma2(x[j], a2, b2);
      }
    }
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) { float lo, hi; up2(x[j], lo, hi); s += lo + hi; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    return;
  }
  float x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
      x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

// ---------------------------------------------------------------------------------------------------------
// Launchers (host)
// ---------------------------------------------------------------------------------------------------------
std::atomic<unsigned long long> g_kernel_launches{0};   // every <<<>>> of this library (rtw_kernel_launches)
#define RTW_COUNT_LAUNCH() g_kernel_launches.fetch_add(1, std::memory_order_relaxed)
template <int R, bool STATS>
static cudaError_t launch_sweep_t(const RenderParams& p, int sm_count, size_t smem, cudaStream_t stream, int* blocks_out) {
  auto kern = k_render_sweep<R, STATS>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  if (e != cudaSuccess) return e;
  int per_sm = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kRenderThreads, smem);
  if (e != cudaSuccess) return e;
  if (per_sm < 1) return cudaErrorLaunchOutOfResources;
  // persistent grid: a whole number of CTAs per SM, never more warps than units of work
  unsigned long long want = (p.n_units + (kRenderThreads / 32) - 1) / (kRenderThreads / 32);
  unsigned long long blocks = static_cast<unsigned long long>(sm_count) * per_sm;
  if (want < blocks) blocks = want < 1 ? 1 : want;
  if (blocks_out) *blocks_out = static_cast<int>(blocks);
  kern<<<static_cast<unsigned>(blocks), kRenderThreads, smem, stream>>>(p);
  RTW_COUNT_LAUNCH();
  return cudaGetLastError();
}

template <bool SMEM, bool STATS, int STEPS, int SERVICE_MIN, int MINB, bool CW = false>
static cudaError_t launch_bvh_t(const RenderParams& p, int sm_count, size_t smem, cudaStream_t stream) {
  auto kern = k_render_bvh<SMEM, STATS, STEPS, SERVICE_MIN, MINB, CW>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  if (e != cudaSuccess) return e;
  int per_sm = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kRenderThreads, smem);
  if (e != cudaSuccess) return e;
  if (per_sm < 1) return cudaErrorLaunchOutOfResources;
  unsigned long long want = (p.n_units + (kRenderThreads / 32) - 1) / (kRenderThreads / 32);
  unsigned long long blocks = static_cast<unsigned long long>(sm_count) * per_sm;
  if (want < blocks) blocks = want < 1 ? 1 : want;
  kern<<<static_cast<unsigned>(blocks), kRenderThreads, smem, stream>>>(p);
  RTW_COUNT_LAUNCH();
  return cudaGetLastError();
}

template <bool SMEM, bool STATS, int NW, int P, int STEPS, int BATCH, int MINB = 1>
static cudaError_t launch_wf_t(const RenderParams& p, int sm_count, size_t smem, cudaStream_t stream) {
  auto kern = k_render_wf<SMEM, STATS, NW, P, STEPS, BATCH, MINB>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  if (e != cudaSuccess) return e;
  int per_sm = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, NW * 32, smem);
  if (e != cudaSuccess) return e;
  if (per_sm < 1) return cudaErrorLaunchOutOfResources;
  unsigned long long want = (p.n_units + NW - 1) / NW;
  unsigned long long blocks = static_cast<unsigned long long>(sm_count) * per_sm;
  if (want < blocks) blocks = want < 1 ? 1 : want;
  kern<<<static_cast<unsigned>(blocks), NW * 32, smem, stream>>>(p);
  RTW_COUNT_LAUNCH();
  return cudaGetLastError();
}

// shared-memory layout of the staged BVH tables: [16 B mbarrier][nodes][leaf refs][sphA][sphB][triangles][per-warp records]
static SmemLayout smem_layout(const RenderParams& p) {
  SmemLayout so{};
  const uint32_t nspheres = static_cast<uint32_t>(p.sc.n_static + p.sc.n_moving);
  so.b_nodes = p.sc.n_cw_nodes > 0 ? static_cast<uint32_t>(p.sc.n_cw_nodes) * 80u : static_cast<uint32_t>(p.sc.n_nodes) * 64u; so.b_refs = (p.n_leaf_refs * 4u + 15u) & ~15u; so.b_sph = nspheres * 16u;
  so.b_tri = static_cast<uint32_t>(p.sc.n_tri) * 48u;
  so.nodes = 16u; so.refs = so.nodes + so.b_nodes; so.sa = so.refs + so.b_refs; so.sb = so.sa + so.b_sph; so.tri = so.sb + so.b_sph;
  so.records = so.tri + so.b_tri;
  return so;
}

cudaError_t launch_render(const RenderParams& p_in, int mode, int rays_per_lane, bool force_perlane, bool stats, int sm_count, cudaStream_t stream,
                          int* variant) {
  RenderParams p = p_in;
  if (variant) *variant = RTW_BVH_NONE;
  // triangle leaves (a ~50-instruction test run by ~6 of 32 lanes when tested at once) are held back until 4 lanes stand on one;
  // sphere leaves are tested at once (holding them back lost 3-8 % on the cover scene)
  p.leaf_min = p.sc.n_tri > 0 ? 4u : 1u;
  if (const char* e = std::getenv("RTW_LEAF_MIN")) p.leaf_min = static_cast<uint32_t>(std::max(1, std::atoi(e)));   // tuning knob
  p.cw_steps = 4u; p.cw_service_min = 20u;
  if (const char* e = std::getenv("RTW_CW_STEPS")) p.cw_steps = static_cast<uint32_t>(std::max(1, std::atoi(e)));
  if (const char* e = std::getenv("RTW_CW_SERVICE")) p.cw_service_min = static_cast<uint32_t>(std::min(32, std::max(1, std::atoi(e))));
  const int cw_minb = std::getenv("RTW_CW_MINB") ? std::atoi(std::getenv("RTW_CW_MINB")) : 3;
  p.so = smem_layout(p);
  const size_t smem = mode == 0 ? 16 + static_cast<size_t>(p.sc.n_static + p.sc.n_moving + 1) * 32 : 0;
  if (mode == 0) {
    // rays_per_lane: paths in flight per lane (1, 2 or 4); 2 measured fastest
    if (rays_per_lane == 1) return stats ? launch_sweep_t<1, true>(p, sm_count, smem, stream, nullptr) : launch_sweep_t<1, false>(p, sm_count, smem, stream, nullptr);
    if (rays_per_lane == 4) return stats ? launch_sweep_t<4, true>(p, sm_count, smem, stream, nullptr) : launch_sweep_t<4, false>(p, sm_count, smem, stream, nullptr);
    return stats ? launch_sweep_t<2, true>(p, sm_count, smem, stream, nullptr) : launch_sweep_t<2, false>(p, sm_count, smem, stream, nullptr);
  }
  const bool cw = p.sc.n_cw_nodes > 0;
  const BvhPlan plan = plan_bvh(p.so.records, p.sc.n_tri, p.sc.leaf_direct != 0, force_perlane, cw);   // rtw_internal.h: who gets which kernel
  if (variant) *variant = plan.variant;
  if (plan.variant == RTW_BVH_CWIDE) {
    // scenes with triangles: compressed 8-wide BVH walked by the per-lane state machine; tables in shared memory when they fit
    // (steps per traversal phase / lanes that must need service: tuned on the 991k-triangle mesh and on suzanne)
    if (plan.tables_in_smem)
      return stats ? launch_bvh_t<true, true, 4, 20, 3, true>(p, sm_count, plan.smem_bytes, stream) : launch_bvh_t<true, false, 4, 20, 3, true>(p, sm_count, plan.smem_bytes, stream);
    if (cw_minb == 4 && !stats) return launch_bvh_t<false, false, 4, 20, 4, true>(p, sm_count, 0, stream);
    if (cw_minb == 2 && !stats) return launch_bvh_t<false, false, 4, 20, 2, true>(p, sm_count, 0, stream);
    return stats ? launch_bvh_t<false, true, 4, 20, 3, true>(p, sm_count, 0, stream) : launch_bvh_t<false, false, 4, 20, 3, true>(p, sm_count, 0, stream);
  }
  if (plan.variant == RTW_BVH_WAVEFRONT) {
    // 96 records per warp, 16 traversal steps between exchanges, shading batches of 32 (tuning record in DESIGN.md)
#define RTW_WF_LAUNCH(SM, NW, MINB)                                                                                          \
  return stats ? launch_wf_t<SM, true, NW, kWfRecords, 16, 32, MINB>(p, sm_count, smem_bytes, stream)                   \
               : launch_wf_t<SM, false, NW, kWfRecords, 16, 32, MINB>(p, sm_count, smem_bytes, stream)
    int warps = plan.warps;
    size_t smem_bytes = plan.smem_bytes;
    if (!plan.tables_in_smem) {
      p.so.records = 16u;   // no staged tables in front of the records
      RTW_WF_LAUNCH(false, 8, 3);
    }
    if (const char* e = std::getenv("RTW_WF_WARPS")) {   // tuning knob: fewer warps per SM than the plan (20 / 24 / 28)
      const int w = std::atoi(e);
      if ((w == 20 || w == 24 || w == 28) && w <= plan.warps) { warps = w; smem_bytes = p.so.records + static_cast<size_t>(w) * wf_warp_bytes(kWfRecords); }
    }
    if (warps == 32)   // the lean tier: 64 registers, 92 records per warp
      return stats ? launch_wf_t<true, true, 32, kWfRecords32, 16, 32, 1>(p, sm_count, smem_bytes, stream)
                   : launch_wf_t<true, false, 32, kWfRecords32, 16, 32, 1>(p, sm_count, smem_bytes, stream);
    switch (warps) {
      case 28: RTW_WF_LAUNCH(true, 28, 1);
      case 24: RTW_WF_LAUNCH(true, 24, 1);
      default: RTW_WF_LAUNCH(true, 20, 1);
    }
#undef RTW_WF_LAUNCH
  }
  // <steps per traversal phase, lanes that must need service before the service phase runs, CTAs per SM>: tables in shared memory
  // (sphere scenes) 8 / 24; tables in L1/L2 6 / 20 with two visits per leaf phase (the (steps, threshold) landscape is within +-4 %: DESIGN.md)
  if (plan.tables_in_smem)
    return stats ? launch_bvh_t<true, true, 8, 24, 4>(p, sm_count, plan.smem_bytes, stream) : launch_bvh_t<true, false, 8, 24, 4>(p, sm_count, plan.smem_bytes, stream);
#ifndef RTW_K2_STEPS
#define RTW_K2_STEPS 6
#endif
  return stats ? launch_bvh_t<false, true, RTW_K2_STEPS, 20, 4>(p, sm_count, 0, stream) : launch_bvh_t<false, false, RTW_K2_STEPS, 20, 4>(p, sm_count, 0, stream);
}

cudaError_t launch_primary_f32(const PrimaryParams& p, int mode, cudaStream_t stream) {
  const size_t smem = mode == 0 ? 16 + static_cast<size_t>(p.sc.n_static + p.sc.n_moving + 1) * 32 : 0;
  const unsigned blocks = (p.npix + kRenderThreads - 1) / kRenderThreads;
  if (mode == 0) {
    cudaError_t e = cudaFuncSetAttribute(k_primary_f32<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return e;
    k_primary_f32<0><<<blocks, kRenderThreads, smem, stream>>>(p);
  } else {
    k_primary_f32<1><<<blocks, kRenderThreads, 0, stream>>>(p);
  }
  RTW_COUNT_LAUNCH();
  return cudaGetLastError();
}

cudaError_t launch_primary_f64(const rtw_primitive* prims, int nprims, const rtw_camera& cam, uint32_t width, uint32_t height, double time,
                               int32_t* prim_id, double* t, double* normal, uint8_t* front, cudaStream_t stream) {
  const unsigned npix = width * height;
  k_primary_f64<<<(npix + 255) / 256, 256, 0, stream>>>(prims, nprims, cam, width, height, time, prim_id, t, normal, front);
  RTW_COUNT_LAUNCH();
  return cudaGetLastError();
}

cudaError_t launch_untile(const long long* gathered, long long* full, uint32_t width, uint32_t height, uint32_t tile_rows, uint32_t count,
                          uint32_t local_rows, cudaStream_t stream) {
  const uint32_t npix = width * height;
  k_untile<<<(npix + 255) / 256, 256, 0, stream>>>(reinterpret_cast<const longlong4*>(gathered), reinterpret_cast<longlong4*>(full), width, height,
                                                   tile_rows, count, local_rows);
  RTW_COUNT_LAUNCH();
  return cudaGetLastError();
}
cudaError_t launch_accum_to_float(const long long* fx, float* out, long long npix, cudaStream_t stream) {
  k_accum_to_float<<<static_cast<unsigned>((npix + 255) / 256), 256, 0, stream>>>(fx, reinterpret_cast<float4*>(out), npix);
  RTW_COUNT_LAUNCH();
  return cudaGetLastError();
}
cudaError_t launch_finalize_rgb8(const float* acc, uint8_t* rgb, long long npix, int spp, cudaStream_t stream) {
  k_finalize_rgb8<<<static_cast<unsigned>((npix + 255) / 256), 256, 0, stream>>>(reinterpret_cast<const float4*>(acc), rgb, npix,
                                                                                  static_cast<double>(spp));
  RTW_COUNT_LAUNCH();
  return cudaGetLastError();
}
cudaError_t launch_finalize_rgb8_fx(const long long* fx, uint8_t* rgb, long long npix, int spp, cudaStream_t stream) {
  k_finalize_rgb8_fx<<<static_cast<unsigned>((npix + 255) / 256), 256, 0, stream>>>(fx, rgb, npix, static_cast<double>(spp));
  RTW_COUNT_LAUNCH();
  return cudaGetLastError();
}
cudaError_t launch_debug_scatter(long long n, const int* kind, const float* fuzz, const float* ior, const float* dir_in, const float* normal,
                                 const uint8_t* front, const float* ball, const float* coin, float* out_dir, float* out_att,
                                 const float* albedo, uint8_t* scattered, cudaStream_t stream) {
  k_debug_scatter<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(n, kind, fuzz, ior, dir_in, normal, front, ball, coin, out_dir,
                                                                               out_att, albedo, scattered);
  RTW_COUNT_LAUNCH();
  return cudaGetLastError();
}
cudaError_t launch_debug_samples(long long n, uint64_t seed, float* ball, float* disk, float* uni, cudaStream_t stream) {
  k_debug_samples<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(
      n, make_uint2(static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32)), ball, disk, uni);
  RTW_COUNT_LAUNCH();
  return cudaGetLastError();
}
cudaError_t launch_ffma_peak(float* out, int blocks, int iters, bool packed, cudaStream_t stream) {
  if (packed) k_ffma_peak<true><<<blocks, 256, 0, stream>>>(out, iters, 0.999999f, 1e-7f);
  else k_ffma_peak<false><<<blocks, 256, 0, stream>>>(out, iters, 0.999999f, 1e-7f);
  RTW_COUNT_LAUNCH();
  return cudaGetLastError();
}

void count_launch() { RTW_COUNT_LAUNCH(); }
unsigned long long launch_count() { return g_kernel_launches.load(std::memory_order_relaxed); }

}  // namespace rtw
