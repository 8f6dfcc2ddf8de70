// rtw_device.cuh -- device-side building blocks of the B200 path tracer (sm_100a).
//
// What each block replaces in the reference (/root/reference/src):
//   philox4x32_10 / u01 / sample_*     random-utils.cpp:6-41 (global mt19937 + rejection loops) -> counter-based
//                                      stream keyed on (pixel, sample, dimension), direct inversion
//   camera_ray                         Camera::get_ray, common-model.cpp:156-167
//   sphere_hit_t<T>                    sphere_hit_helper, common-model.cpp:64-91 (root selection, tmin/tmax rule)
//   triangle_hit<T>                    Triangle::hit, common-model.cpp:103-125
//   scatter_dir                        Lambertian/Metal/Dielectric::scatter, common-model.cpp:13-62
//   sky_color                          ray_color miss branch, render.cpp:125-128
// The reference computes in double; the render kernels compute in float (fp64 only for spheres flagged "big",
// where |f|^2 - r^2 cancels catastrophically in fp32).  Quirks Q1-Q7 of SURVEY.md section 0 are kept on purpose.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace rtw {

constexpr float kTMin = 0.001f;  // render.cpp:33 default tmin of BVHNode::hit
constexpr float kInf = __builtin_huge_valf();
constexpr int kBvhStack = 64;  // the host builder bounds the tree depth by 32 + log2(n) (rtw_bvh.h) and rtw_scene_upload checks it
constexpr int kCwStack = 32;   // compressed wide BVH: at most one stack entry per level, rtw_scene_upload checks the depth

// ---------------------------------------------------------------------------------------------------------
// Device scene (flattened SoA; built by rtw_scene_upload)
// ---------------------------------------------------------------------------------------------------------
struct DevCamera {
  float origin[3], lower_left[3], horizontal[3], vertical[3], u[3], v[3];
  float lens_radius, t0, t1;
};

struct BigSphere {  // spheres with |r| >= kBigRadius: tested in fp64 for every ray, outside tables and BVH
  double c0[3];
  double dc[3];
  double r;
  int32_t prim_id;
  int32_t material;
};

struct DevScene {
  // small spheres, static ones first then moving ones
  const float4* sphA;  // (c0.x, c0.y, c0.z, r2c)   r2c = conservative r^2 for the line-distance reject test
  const float4* sphB;  // (dc.x, dc.y, dc.z, r)     dc = c1 - c0 (0 for static), r = true (signed) radius
  const int2* sphId;   // (primitive id, material)
  int32_t n_static, n_moving;
  const BigSphere* big;
  int32_t n_big;
  // triangles: 3 float4 each: (a.xyz, n.x) (e1.xyz, n.y) (e2.xyz, n.z), n = cross(e1, e2) un-normalised
  const float4* tri;
  const int2* triId;
  int32_t n_tri;
  // BVH over small spheres + triangles: 4 float4 per node, each child box as centre c and half-extent e
  //   (left, right) pairs per axis, so that one packed FFMA2 handles both children:
  //   q0 = (lc.x, rc.x, lc.y, rc.y) q1 = (lc.z, rc.z, le.x, re.x)
  //   q2 = (le.y, re.y, le.z, re.z) q3 = (left, right, -, -) as int bits
  //   child >= 0: inner node index; child < 0: leaf, ~child = (first << 5) | count  into leafRefs,
  //   or (leaf_direct) ~child = the primitive reference itself with bit 29 set (so that the code is never -1 = "no node")
  const float4* nodes;
  const uint32_t* leafRefs;  // (kind << 30) | index, kind 0 = sphere table index, 1 = triangle index
  int32_t n_nodes;
  int32_t leaf_direct;       // 1: single-primitive leaves, child code = ~reference (leafRefs unused)
  // Scenes with triangles use a compressed 8-wide BVH instead (rtw_bvh.h: CwNode, 80 bytes = 5 x uint4 per node).  Then `tri` / `triId`
  // hold the LEAF PRIMITIVES in leaf order (48-byte records, the up-to-3 primitives of a leaf next to each other): a triangle as
  // above, a small sphere as (sphA entry) (sphB entry) (NaN, sphere table index as int bits, 0, 0); n_tri counts these records.
  const uint4* cwNodes;
  int32_t n_cw_nodes;
  int32_t cw_has_spheres;    // 1: some leaf records are spheres (the NaN tag has to be looked at)
  // materials
  const float4* matA;  // (albedo.r, albedo.g, albedo.b, fuzz)
  const float2* matB;  // (ior, kind as int bits)
  DevCamera cam;
};

// ---------------------------------------------------------------------------------------------------------
// small vector helpers
// ---------------------------------------------------------------------------------------------------------
template <typename T> struct V3 { T x, y, z; };
using F3 = V3<float>;
using D3 = V3<double>;

template <typename T> __host__ __device__ __forceinline__ V3<T> mk(T x, T y, T z) { V3<T> r; r.x = x; r.y = y; r.z = z; return r; }
template <typename T> __host__ __device__ __forceinline__ V3<T> operator+(V3<T> a, V3<T> b) { return mk<T>(a.x + b.x, a.y + b.y, a.z + b.z); }
template <typename T> __host__ __device__ __forceinline__ V3<T> operator-(V3<T> a, V3<T> b) { return mk<T>(a.x - b.x, a.y - b.y, a.z - b.z); }
template <typename T> __host__ __device__ __forceinline__ V3<T> operator-(V3<T> a) { return mk<T>(-a.x, -a.y, -a.z); }
template <typename T> __host__ __device__ __forceinline__ V3<T> operator*(V3<T> a, T s) { return mk<T>(a.x * s, a.y * s, a.z * s); }
template <typename T> __host__ __device__ __forceinline__ V3<T> operator*(T s, V3<T> a) { return mk<T>(a.x * s, a.y * s, a.z * s); }
template <typename T> __host__ __device__ __forceinline__ T dot(V3<T> a, V3<T> b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
template <typename T> __host__ __device__ __forceinline__ V3<T> cross(V3<T> a, V3<T> b) {
  return mk<T>(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y);
}
__device__ __forceinline__ float rsqrt_(float x) { return rsqrtf(x); }
__device__ __forceinline__ double rsqrt_(double x) { return 1.0 / sqrt(x); }
__device__ __forceinline__ float sqrt_(float x) { return sqrtf(x); }
// single-MUFU approximations (<= 2 ulp) for the traversal: box slabs are padded and accepted hits are re-evaluated in fp64
__device__ __forceinline__ float fast_rcp(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float fast_sqrt(float x) { float y; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ double sqrt_(double x) { return sqrt(x); }
template <typename T> __device__ __forceinline__ V3<T> normalize(V3<T> a) { return a * rsqrt_(dot(a, a)); }

// Double-precision reciprocal / reciprocal square root without the software div/sqrt sequences: fp32 seed (MUFU) +
// two Newton steps in DFMA.  Relative error ~1e-14 for arguments inside the float range, which is all the renderer
// feeds them (squared lengths and discriminants of scene-scale quantities).
__device__ __forceinline__ double rcp_nr(double x) {
  double y = static_cast<double>(__frcp_rn(static_cast<float>(x)));
  y = y * fma(-x, y, 2.0);
  y = y * fma(-x, y, 2.0);
  return y;
}
__device__ __forceinline__ double rsqrt_nr(double x) {
  double y = static_cast<double>(rsqrtf(static_cast<float>(x)));
  y = y * fma(-0.5 * x, y * y, 1.5);
  y = y * fma(-0.5 * x, y * y, 1.5);
  return y;
}
__device__ __forceinline__ double sqrt_nr(double x) { return x > 1e-35 ? x * rsqrt_nr(x) : 0.0; }

// ---------------------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al., SC'11): counter = (pixel, sample, dimension, 0), key = seed.
//   dimension 0: (pixel jitter x, pixel jitter y, lens radius, lens azimuth) from the top 24 bits of the four words, and the
//                shutter time from the low bytes of words 0..2 (u01_low): one block per primary ray
//   dimension 1: unused
//   dimension 2+k: k-th scatter of the path: (ball z, ball azimuth, ball radius, Schlick coin)
// ---------------------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint32_t mulhi32(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
  return __umulhi(a, b);
#else
  return static_cast<uint32_t>((static_cast<uint64_t>(a) * b) >> 32);
#endif
}
__host__ __device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const uint32_t hi0 = mulhi32(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const uint32_t hi1 = mulhi32(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u;
    k.y += 0xBB67AE85u;
  }
  return c;
}
__host__ __device__ __forceinline__ float u01(uint32_t x) { return static_cast<float>(x >> 8) * 5.9604644775390625e-8f; }
// A fifth 24-bit uniform from the low bytes u01() leaves unused in the x, y, z words of one Philox block.
__host__ __device__ __forceinline__ float u01_low(uint4 x) {
  return static_cast<float>(((x.x & 0xffu) << 16) | ((x.y & 0xffu) << 8) | (x.z & 0xffu)) * 5.9604644775390625e-8f;
}

// Uniform point of the unit ball restricted to the positive octant, NOT normalised: the distribution the
// reference's random_unit_vector() actually has (random-utils.cpp:23-33), by direct inversion:
// z uniform in [0,1), azimuth uniform in [0,pi/2), radius = cbrt(xi).
__device__ __forceinline__ F3 sample_octant_ball(float xz, float xphi, float xrho) {
  // radius = cbrt(xi) as exp2(log2(xi) / 3) and the azimuth through the MUFU sine/cosine (argument in [0, pi/2)): absolute
  // errors ~1e-6, far below what the statistical image comparison or the same-stream oracle comparison can see
  const float rho = xrho > 0.0f ? exp2f(__log2f(xrho) * 0.33333334f) : 0.0f;
  const float sn = sqrtf(fmaf(-xz, xz, 1.0f));
  float s, c;
  __sincosf(1.5707963267948966f * xphi, &s, &c);
  const float k = rho * sn;
  return mk<float>(k * c, k * s, rho * xz);
}
// Uniform point of the unit disk (random-utils.cpp:34-41) by direct inversion.
__device__ __forceinline__ float2 sample_disk(float xr, float xphi) {
  const float r = sqrtf(xr);
  float s, c;
  sincospif(2.0f * xphi, &s, &c);
  return make_float2(r * c, r * s);
}

// Camera::get_ray (common-model.cpp:156-167): s,t in viewport units, disk sample and shutter time explicit.
template <typename T, typename Cam>
__device__ __forceinline__ void camera_ray(const Cam& cam, T s, T t, T diskx, T disky, V3<T>& org, V3<T>& dir) {
  const T rx = static_cast<T>(cam.lens_radius) * diskx, ry = static_cast<T>(cam.lens_radius) * disky;
  org = mk<T>(cam.origin[0] + cam.u[0] * rx + cam.v[0] * ry, cam.origin[1] + cam.u[1] * rx + cam.v[1] * ry,
              cam.origin[2] + cam.u[2] * rx + cam.v[2] * ry);
  dir = mk<T>(cam.lower_left[0] + s * cam.horizontal[0] + t * cam.vertical[0] - org.x,
              cam.lower_left[1] + s * cam.horizontal[1] + t * cam.vertical[1] - org.y,
              cam.lower_left[2] + s * cam.horizontal[2] + t * cam.vertical[2] - org.z);
}

// ---------------------------------------------------------------------------------------------------------
// Intersection
// ---------------------------------------------------------------------------------------------------------
// One 64-byte BVH node.  From global memory: two 256-bit loads (LDG.E.256, sm_100) instead of four 128-bit ones -- the traversal of
// a big mesh is bound by L1 wavefronts (one per distinct line per load instruction), and every lane reads a different node.
__device__ __forceinline__ void ldg256(const void* p, float4& a, float4& b) {
  asm("ld.global.nc.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
      : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w)
      : "l"(p));
}
template <bool SMEM>
__device__ __forceinline__ void load_node(const float4* nodes, int node, float4& q0, float4& q1, float4& q2, float4& q3) {
  if (SMEM) {
    q0 = nodes[4 * node]; q1 = nodes[4 * node + 1]; q2 = nodes[4 * node + 2]; q3 = nodes[4 * node + 3];
  } else {
    ldg256(nodes + 4 * node, q0, q1);
    ldg256(nodes + 4 * node + 2, q2, q3);
  }
}

// Slab test of both children of a BVH node (Aabb::hit, common-model.h:71-84) against [kTMin, tmax].  With the box stored as
// centre c and half-extent e >= 0 the entry/exit parameters along one axis are (c -+ e) * (1/d) - o/d = tc -+ e * |1/d|:
// three FMAs per axis and box on the FMA pipe and no per-axis min/max on the ALU pipe (measured on the cover scene in round 1:
// min/max form 6670, FMUL + FADD|.| form 6850, FFMA form 7040 Mpaths/s).
// The node stores the (left, right) values of an axis side by side.  That layout also feeds the packed FP32 FMA of sm_100 (PTX
// fma.rn.f32x2, SASS FFMA2: an LDS.128 / LDG.256 drops the pairs into aligned registers, 1/d, -o/d, |1/d| enter as broadcast operands),
// which issues the 18 FMAs of a node as 9 instructions, bit-identical results.  Built, measured on one box against this scalar form
// and NOT used (-DRTW_FFMA2_SLABS builds it; scripts/ffma2_probe.cu, profiles/r02_ffma2.txt): FFMA2 runs at half the issue rate of
// FFMA (same FP32 peak), K2w executes 3.7 % fewer warp instructions and is 0.5 % SLOWER (issue-active 79 -> 76 %), suzanne -1.5 %,
// the 991k-triangle mesh -2.2 %: these kernels wait on dependent-instruction latency, and the packed form has less ILP per node.
__device__ __forceinline__ uint64_t pk2(float lo, float hi) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void up2(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ void node_slabs(const float4 q0, const float4 q1, const float4 q2, float idx, float idy, float idz, float odx,
                                           float ody, float odz, float tmax, float& ln, float& lf, float& rn, float& rf) {
  const float ax = fabsf(idx), ay = fabsf(idy), az = fabsf(idz);
#ifndef RTW_FFMA2_SLABS   // the product: 18 scalar FFMA (the packed form below was measured slower, see the comment above)
  float cx = fmaf(q0.x, idx, -odx), cy = fmaf(q0.z, idy, -ody), cz = fmaf(q1.x, idz, -odz);
  ln = fmaxf(fmaxf(fmaf(-q1.z, ax, cx), fmaf(-q2.x, ay, cy)), fmaxf(fmaf(-q2.z, az, cz), kTMin));
  lf = fminf(fminf(fmaf(q1.z, ax, cx), fmaf(q2.x, ay, cy)), fminf(fmaf(q2.z, az, cz), tmax));
  cx = fmaf(q0.y, idx, -odx); cy = fmaf(q0.w, idy, -ody); cz = fmaf(q1.y, idz, -odz);
  rn = fmaxf(fmaxf(fmaf(-q1.w, ax, cx), fmaf(-q2.y, ay, cy)), fmaxf(fmaf(-q2.w, az, cz), kTMin));
  rf = fminf(fminf(fmaf(q1.w, ax, cx), fmaf(q2.y, ay, cy)), fminf(fmaf(q2.w, az, cz), tmax));
#else
  const uint64_t tcx = fma2(pk2(q0.x, q0.y), pk2(idx, idx), pk2(-odx, -odx));
  const uint64_t tcy = fma2(pk2(q0.z, q0.w), pk2(idy, idy), pk2(-ody, -ody));
  const uint64_t tcz = fma2(pk2(q1.x, q1.y), pk2(idz, idz), pk2(-odz, -odz));
  const uint64_t ex = pk2(q1.z, q1.w), ey = pk2(q2.x, q2.y), ez = pk2(q2.z, q2.w);
  float lnx, rnx, lny, rny, lnz, rnz, lfx, rfx, lfy, rfy, lfz, rfz;
  up2(fma2(ex, pk2(-ax, -ax), tcx), lnx, rnx); up2(fma2(ex, pk2(ax, ax), tcx), lfx, rfx);
  up2(fma2(ey, pk2(-ay, -ay), tcy), lny, rny); up2(fma2(ey, pk2(ay, ay), tcy), lfy, rfy);
  up2(fma2(ez, pk2(-az, -az), tcz), lnz, rnz); up2(fma2(ez, pk2(az, az), tcz), lfz, rfz);
  ln = fmaxf(fmaxf(lnx, lny), fmaxf(lnz, kTMin));
  lf = fminf(fminf(lfx, lfy), fminf(lfz, tmax));
  rn = fmaxf(fmaxf(rnx, rny), fmaxf(rnz, kTMin));
  rf = fminf(fminf(rfx, rfy), fminf(rfz, tmax));
#endif
}


// ---------------------------------------------------------------------------------------------------------
// Compressed wide BVH traversal (node layout and references: rtw_bvh.h, CwNode).  Replaces BVHNode::hit (render.cpp:52-71) for
// scenes with triangles.  A traversal state is two 64-bit groups and a stack of node groups:
//   ngroup = (child_base, hit bits of the inner children in bits 24..31 | imask in bits 0..7)
//   tgroup = (prim_base, hit bits of this node's leaf primitives in bits 0..23)
// The hit bit of the inner child in slot s sits at 24 + (s ^ oct_inv): taking the highest set bit first visits the children front to
// back for the ray's octant without comparing distances.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t sign_extend_s8x4(uint32_t x) {  // every byte -> 0xff if its top bit is set, else 0x00
  uint32_t v;
  asm("prmt.b32 %0, %1, %1, 0xba98;" : "=r"(v) : "r"(x));
  return v;
}
// byte j of x -> the float 256 + b / 128 (bits 0x4380bb00): one PRMT instead of a shift, a mask and an integer-to-float conversion.
// With A = 128 * step * (1/d) and C = (p - o) * (1/d) - 256 * A the slab parameter of plane b is fma(f, A, C).
template <int J>
__device__ __forceinline__ float q8_to_float(uint32_t x) {
  return __uint_as_float(__byte_perm(x, 0x43800000u, 0x7604u | (J << 4)));
}
__device__ __forceinline__ uint4 ldg128(const uint4* p) {
  uint4 v;
  asm("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ uint32_t cw_oct_inv4(F3 d) {
  const uint32_t oct = (d.x < 0.0f ? 1u : 0u) | (d.y < 0.0f ? 2u : 0u) | (d.z < 0.0f ? 4u : 0u);
  return (7u - oct) * 0x01010101u;
}

// Slab test of the (up to) 8 children of node `index` against [tmin, tmax] (Aabb::hit, common-model.h:71-84, on dequantised boxes).
template <bool SMEM>
__device__ __forceinline__ void cw_intersect_node(const uint4* __restrict__ nodes, uint32_t index, F3 o, float idx, float idy, float idz,
                                                  uint32_t oct_inv4, float tmin, float tmax, uint2& ngroup, uint2& tgroup) {
  const uint4* np = nodes + 5u * index;
  uint4 n0, n1, n2, n3, n4;
  if (SMEM) { n0 = np[0]; n1 = np[1]; n2 = np[2]; n3 = np[3]; n4 = np[4]; }
  else { n0 = ldg128(np); n1 = ldg128(np + 1); n2 = ldg128(np + 2); n3 = ldg128(np + 3); n4 = ldg128(np + 4); }
  const uint32_t e = n0.w;
  const float ax = __uint_as_float(((e & 0xffu) + 7u) << 23) * idx;
  const float ay = __uint_as_float(((e >> 8 & 0xffu) + 7u) << 23) * idy;
  const float az = __uint_as_float(((e >> 16 & 0xffu) + 7u) << 23) * idz;
  const float cx = fmaf(__uint_as_float(n0.x) - o.x, idx, -256.0f * ax);
  const float cy = fmaf(__uint_as_float(n0.y) - o.y, idy, -256.0f * ay);
  const float cz = fmaf(__uint_as_float(n0.z) - o.z, idz, -256.0f * az);
  const bool nx = idx < 0.0f, ny = idy < 0.0f, nz = idz < 0.0f;
  uint32_t hits = 0u;
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const uint32_t meta4 = half ? n1.w : n1.z;
    const uint32_t is_inner4 = (meta4 & (meta4 << 1)) & 0x10101010u;
    const uint32_t inner_mask4 = sign_extend_s8x4(is_inner4 << 3);
    const uint32_t bit_index4 = (meta4 ^ (oct_inv4 & inner_mask4)) & 0x1f1f1f1fu;
    const uint32_t child_bits4 = (meta4 >> 5) & 0x07070707u;
    const uint32_t lox = half ? n2.y : n2.x, loy = half ? n2.w : n2.z, loz = half ? n3.y : n3.x;
    const uint32_t hix = half ? n3.w : n3.z, hiy = half ? n4.y : n4.x, hiz = half ? n4.w : n4.z;
    const uint32_t nearx = nx ? hix : lox, farx = nx ? lox : hix;
    const uint32_t neary = ny ? hiy : loy, fary = ny ? loy : hiy;
    const uint32_t nearz = nz ? hiz : loz, farz = nz ? loz : hiz;
#define RTW_CW_CHILD(J)                                                                                                   \
    {                                                                                                                       \
      const float t0 = fmaxf(fmaxf(fmaf(q8_to_float<J>(nearx), ax, cx), fmaf(q8_to_float<J>(neary), ay, cy)),              \
                             fmaxf(fmaf(q8_to_float<J>(nearz), az, cz), tmin));                                             \
      const float t1 = fminf(fminf(fmaf(q8_to_float<J>(farx), ax, cx), fmaf(q8_to_float<J>(fary), ay, cy)),                \
                             fminf(fmaf(q8_to_float<J>(farz), az, cz), tmax));                                              \
      if (t0 <= t1) hits |= ((child_bits4 >> (8 * J)) & 0xffu) << ((bit_index4 >> (8 * J)) & 0xffu);                        \
    }
    RTW_CW_CHILD(0) RTW_CW_CHILD(1) RTW_CW_CHILD(2) RTW_CW_CHILD(3)
#undef RTW_CW_CHILD
  }
  ngroup = make_uint2(n1.x, (hits & 0xff000000u) | (e >> 24));
  tgroup = make_uint2(n1.y, hits & 0x00ffffffu);
}

// Takes the nearest not yet visited inner child out of ngroup (pushing what remains) and returns its node index.
__device__ __forceinline__ uint32_t cw_pop_child(uint2& ngroup, uint32_t oct_inv4, uint2* stack, int& sp) {
  const uint32_t bit = 31u - __clz(ngroup.y);             // highest hit bit: in 24..31
  ngroup.y &= ~(1u << bit);
  const uint32_t imask = ngroup.y & 0xffu, base = ngroup.x;
  if (ngroup.y & 0xff000000u) { if (sp < kCwStack) stack[sp++] = ngroup; }
  const uint32_t slot = (bit - 24u) ^ (oct_inv4 & 0xffu);
  return base + __popc(imask & ~(0xffffffffu << slot));
}

// sphere_hit_helper (common-model.cpp:64-91).  Same roots and the same accept rule (nearer root if inside
// [tmin,tmax], else the farther one), but the discriminant is evaluated as a*(r^2 - |oc - (h/a) d|^2), which is
// algebraically h^2 - a*c and does not cancel in fp32 (Haines et al., "Precision improvements for ray/sphere
// intersection", Ray Tracing Gems ch. 7, adapted to un-normalised d).  Returns t or -1.
template <typename T>
__device__ __forceinline__ T sphere_hit_t(V3<T> o, V3<T> d, T a, T inv_a, V3<T> c, T r, T tmin, T tmax) {
  const V3<T> oc = o - c;
  const T h = dot(oc, d);
  const T k = h * inv_a;
  const V3<T> l = mk<T>(oc.x - k * d.x, oc.y - k * d.y, oc.z - k * d.z);
  const T disc = r * r - dot(l, l);
  if (!(disc >= T(0))) return T(-1);
  const T sq = sqrt_(a * disc);
  T root = (-h - sq) * inv_a;
  if (root < tmin || root > tmax) {
    root = (-h + sq) * inv_a;
    if (root < tmin || root > tmax) return T(-1);
  }
  return root;
}

// Same test with the approximate square root, for traversal kernels that refine the accepted hit afterwards.
__device__ __forceinline__ float sphere_hit_fast(F3 o, F3 d, float a, float inv_a, F3 c, float r, float tmin, float tmax) {
  const F3 oc = o - c;
  const float h = dot(oc, d);
  const float k = h * inv_a;
  const F3 l = mk<float>(fmaf(-k, d.x, oc.x), fmaf(-k, d.y, oc.y), fmaf(-k, d.z, oc.z));
  const float disc = fmaf(r, r, -dot(l, l));
  if (!(disc >= 0.0f)) return -1.0f;
  const float sq = fast_sqrt(a * disc);
  float root = (-h - sq) * inv_a;
  if (root < tmin || root > tmax) {
    root = (-h + sq) * inv_a;
    if (root < tmin || root > tmax) return -1.0f;
  }
  return root;
}

// Triangle::hit (common-model.cpp:103-125) with e1, e2, n = e1 x e2 precomputed per triangle.  Un-normalised
// n and d, back faces culled by det >= 1e-6 exactly as the reference.  Returns t or -1.
template <typename T>
__device__ __forceinline__ T triangle_hit(V3<T> o, V3<T> d, V3<T> a, V3<T> e1, V3<T> e2, V3<T> n, T tmin, T tmax) {
  const T det = -dot(d, n);
  const T invdet = T(1) / det;
  const V3<T> ao = o - a;
  const V3<T> dao = cross(ao, d);
  const T u = dot(e2, dao) * invdet;
  const T v = -dot(e1, dao) * invdet;
  const T t = dot(ao, n) * invdet;
  if (det >= T(1e-6) && t >= tmin && t <= tmax && u >= T(0) && v >= T(0) && (u + v) <= T(1)) return t;
  return T(-1);
}

// The same test for the fp32 tracers, without the division on the reject path: with det > 0 the reference's conditions
// u >= 0, v >= 0, u + v <= 1, tmin <= t <= tmax on u = un/det, v = vn/det, t = tn/det are the same conditions on the numerators
// scaled by det; only an accepted hit (one in three or four tests on the meshes) pays for t = tn / det.  The leaf test runs at
// ~6 of 32 lanes (profiles/r01_prof_k2_suzanne_hotlines.txt), so every instruction removed from it counts five-fold.
__device__ __forceinline__ float triangle_hit_fast(F3 o, F3 d, F3 a, F3 e1, F3 e2, F3 n, float tmin, float tmax) {
  const float det = -dot(d, n);
  const F3 ao = o - a;
  const F3 dao = cross(ao, d);
  const float un = dot(e2, dao);
  const float vn = -dot(e1, dao);
  const float tn = dot(ao, n);
  if (det >= 1e-6f && un >= 0.0f && vn >= 0.0f && (un + vn) <= det && tn >= tmin * det && tn <= tmax * det) return tn / det;
  return -1.0f;
}

// ---------------------------------------------------------------------------------------------------------
// Materials (common-model.cpp:13-62).  Returns false when the path is absorbed.
// ---------------------------------------------------------------------------------------------------------
enum : int { kLambertian = 0, kMetal = 1, kDielectric = 2 };

__device__ __forceinline__ F3 reflect3(F3 I, F3 N) { return I - N * (2.0f * dot(N, I)); }

__device__ __forceinline__ bool scatter_dir(int kind, float fuzz, float ior, F3 d_in, F3 n, bool front, F3 ball,
                                            float coin, F3& d_out) {
  if (kind == kLambertian) {
    // common-model.cpp:15-18: absorbed only if normal == ball component-wise within 1e-8
    if (fabsf(n.x - ball.x) < 1e-8f && fabsf(n.y - ball.y) < 1e-8f && fabsf(n.z - ball.z) < 1e-8f) return false;
    d_out = n + ball;
    return true;
  }
  if (kind == kMetal) {  // always scatters, keeps |d_in|
    d_out = reflect3(d_in, n) + ball * fuzz;
    return true;
  }
  // Dielectric: Schlick on the normalised incoming direction; NaN from sqrt(1-cos^2) compares false exactly
  // like the reference's doubles do (common-model.cpp:45-54).
  const F3 unit = normalize(d_in);
  const float cos_theta = dot(-unit, n);
  const float sin_theta = sqrtf(1.0f - cos_theta * cos_theta);
  const float ratio = front ? (1.0f / ior) : ior;
  const bool cannot_refract = ratio * sin_theta > 1.0f;
  float r0 = (1.0f - ratio) / (1.0f + ratio);
  r0 = r0 * r0;
  const float m = 1.0f - cos_theta;
  const float m2 = m * m;
  const float refl = r0 + (1.0f - r0) * (m2 * m2 * m);
  F3 dir;
  if (cannot_refract || refl > coin) {
    dir = reflect3(unit, n);
  } else {  // glm::refract
    const float dn = dot(n, unit);
    const float k = 1.0f - ratio * ratio * (1.0f - dn * dn);
    dir = (k >= 0.0f) ? (unit * ratio - n * (ratio * dn + sqrtf(k))) : mk<float>(0.f, 0.f, 0.f);
  }
  d_out = dir + ball * fuzz;
  return true;
}

// render.cpp:125-128
__device__ __forceinline__ F3 sky_color(F3 d) {
  const float uy = d.y * rsqrtf(dot(d, d));
  const float t = 0.5f * (uy + 1.0f);
  return mk<float>((1.0f - t) + t * 0.5f, (1.0f - t) + t * 0.7f, (1.0f - t) + t * 1.0f);
}

// ---------------------------------------------------------------------------------------------------------
// Hit encoding shared by the two tracers: >= 0 small-sphere table index; kHitTri | triangle index;
// kHitBig | big-sphere index; kMiss.
// ---------------------------------------------------------------------------------------------------------
constexpr int kMiss = -1;
constexpr int kHitTri = 0x40000000;
constexpr int kHitBig = 0x20000000;

// fp64 test of the big spheres (reference formula verbatim in double: common-model.cpp:70-82).
__device__ __forceinline__ void trace_big_spheres(const DevScene& sc, F3 o, F3 d, float tm, float& best_t, int& best_i) {
  for (int k = 0; k < sc.n_big; ++k) {
    const BigSphere& b = sc.big[k];
    {
      // fp32 early out: origin clearly outside (|oc|^2 - r^2 far beyond its fp32 rounding error) and the ray pointing
      // away from the centre can only miss; everything else takes the exact fp64 test below
      const float ocx = o.x - static_cast<float>(b.c0[0] + tm * b.dc[0]), ocy = o.y - static_cast<float>(b.c0[1] + tm * b.dc[1]),
                  ocz = o.z - static_cast<float>(b.c0[2] + tm * b.dc[2]);
      const float rf = static_cast<float>(b.r);
      const float q = ocx * ocx + ocy * ocy + ocz * ocz;
      const float r2 = rf * rf;
      if (q - r2 > 1e-4f * (q + r2) && ocx * d.x + ocy * d.y + ocz * d.z > 0.0f) continue;
    }
    const double time = tm;
    const D3 c = mk<double>(b.c0[0] + time * b.dc[0], b.c0[1] + time * b.dc[1], b.c0[2] + time * b.dc[2]);
    const D3 od = mk<double>(o.x, o.y, o.z), dd = mk<double>(d.x, d.y, d.z);
    const D3 oc = od - c;
    const double a = dot(dd, dd), h = dot(oc, dd), cc = dot(oc, oc) - b.r * b.r;
    const double disc = h * h - a * cc;
    if (disc < 0.0) continue;
    const double sq = sqrt_nr(disc), ia = rcp_nr(a);
    const double tmin = static_cast<double>(kTMin), tmax = static_cast<double>(best_t);
    double root = (-h - sq) * ia;
    if (root < tmin || root > tmax) {
      root = (-h + sq) * ia;
      if (root < tmin || root > tmax) continue;
    }
    best_t = static_cast<float>(root);
    best_i = kHitBig | k;
  }
}

// Geometry of the accepted hit: point, shading normal (as the reference defines it), front flag, material.
struct HitGeom {
  F3 p, n;
  float t;  // refined parameter of the hit
  bool front;
  int material, prim_id;
};

// Accepted sphere hit: the reference formula (common-model.cpp:70-88) is re-evaluated once in double from the fp32 ray and the
// table entry, keeping the root the fp32 tracer selected.  The tracer only has to FIND the hit; point and normal then carry
// fp32 rounding of the result instead of fp32 cancellation (|o-c| ~ 13 against r = 0.2 in the cover scene).  Small and big
// spheres share ONE double-precision path (only the table reads differ) so that ground hits and small-sphere hits sitting in
// neighbouring lanes of a shading batch do not diverge.
__device__ __forceinline__ HitGeom hit_geometry(const DevScene& sc, const float4* sphA, const float4* sphB, F3 o, F3 d,
                                                  float tm, float t, int hit) {
  HitGeom g;
  if (hit & kHitTri) {
    const int i = hit & ~kHitTri;
    const float4 q0 = __ldg(&sc.tri[3 * i]), q1 = __ldg(&sc.tri[3 * i + 1]), q2 = __ldg(&sc.tri[3 * i + 2]);
    g.p = o + d * t;
    g.t = t;
    g.n = mk<float>(q0.w, q1.w, q2.w);
    g.front = true;
    const int2 id = __ldg(&sc.triId[i]);
    g.prim_id = id.x; g.material = id.y;
    return g;
  }
  const double time = tm;
  D3 c;
  double rr;
  if (hit & kHitBig) {
    const BigSphere& b = sc.big[hit & ~kHitBig];
    c = mk<double>(b.c0[0] + time * b.dc[0], b.c0[1] + time * b.dc[1], b.c0[2] + time * b.dc[2]);
    rr = b.r;
    g.prim_id = b.prim_id; g.material = b.material;
  } else {
    const float4 A = sphA[hit], B = sphB[hit];
    c = mk<double>(A.x + time * B.x, A.y + time * B.y, A.z + time * B.z);
    rr = B.w;
    const int2 id = __ldg(&sc.sphId[hit]);
    g.prim_id = id.x; g.material = id.y;
  }
  const D3 od = mk<double>(o.x, o.y, o.z), dd = mk<double>(d.x, d.y, d.z);
  const D3 oc = od - c;
  const double a = dot(dd, dd), h = dot(oc, dd), cc = dot(oc, oc) - rr * rr;
  const double disc = h * h - a * cc;
  const double sq = sqrt_nr(disc), ia = rcp_nr(a);
  const double r1 = (-h - sq) * ia, r2 = (-h + sq) * ia;
  const double td = (fabs(r1 - static_cast<double>(t)) <= fabs(r2 - static_cast<double>(t))) ? r1 : r2;
  const D3 p = od + dd * td;
  const D3 pc = p - c;
  const D3 n = pc * rsqrt_nr(dot(pc, pc));
  const bool front = (dot(dd, n) < 0.0) != (rr < 0.0);
  g.p = mk<float>(static_cast<float>(p.x), static_cast<float>(p.y), static_cast<float>(p.z));
  g.n = front ? mk<float>(static_cast<float>(n.x), static_cast<float>(n.y), static_cast<float>(n.z))
              : mk<float>(static_cast<float>(-n.x), static_cast<float>(-n.y), static_cast<float>(-n.z));
  g.front = front;
  g.t = static_cast<float>(td);
  return g;
}


}  // namespace rtw
