// rtw_build.cu -- BVH construction on the GPU (replaces the per-render host build of render.cpp:73-110 for big
// scenes when build time matters more than the last 10-20 % of tree quality).
//
// Linear BVH: 63-bit Morton codes of the primitive centroids, one radix sort, the binary radix tree of Karras ("Maximizing Parallelism
// in the Construction of BVHs, Octrees, and k-d Trees", HPG 2012) built in one pass, bounds fitted bottom-up with one atomic flag per
// node, and the result packed straight into the 64-byte two-child nodes the traversal kernels walk (child boxes as centre + half-extent
// rounded up exactly as the host builder does, rtw_bvh.h).  A million triangles take a few milliseconds instead of the ~80 ms of the
// binned-SAH build on 16 host cores.  The radix tree is a plain spatial-median tree and traces 14 % slower than the SAH tree, so its top
// (host SAH over a few thousand subtree boxes) and the subtrees themselves (binned SAH, one warp per subtree) are rebuilt: 3.3 % slower.
// Any correct BVH returns the same closest hit: parity (primitive ids) does not depend on which builder ran.
#include <cub/device/device_radix_sort.cuh>

#include <algorithm>
#include <cfloat>
#include <mutex>
#include <type_traits>
#include <cstdlib>
#include <vector>

#include "rtw_bvh.h"
#include "rtw_host.h"

namespace rtw {

using GpuBuildItem = BvhBuilder::Item;   // {Box3 box; float c[3]; uint32_t ref}: the host flattener's build records, uploaded as they are

namespace {

__device__ __forceinline__ int float_as_ordered(float f) {  // monotone float -> int map (atomicMin / atomicMax on floats)
  const int i = __float_as_int(f);
  return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float ordered_as_float(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

// bounds of the centroids: out[0..2] = min, out[3..5] = max (ordered-int encoding)
__global__ void __launch_bounds__(256) k_centroid_bounds(const GpuBuildItem* __restrict__ items, int n, int* __restrict__ out) {
  float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const Box3 b = items[i].box;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const float c = 0.5f * (b.lo[k] + b.hi[k]);
      lo[k] = fminf(lo[k], c); hi[k] = fmaxf(hi[k], c);
    }
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      lo[k] = fminf(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], off));
      hi[k] = fmaxf(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], off));
    }
  }
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int k = 0; k < 3; ++k) { atomicMin(out + k, float_as_ordered(lo[k])); atomicMax(out + 3 + k, float_as_ordered(hi[k])); }
  }
}

__device__ __forceinline__ unsigned long long spread21(unsigned long long x) {  // 21 bits -> every third bit
  x &= 0x1fffffull;
  x = (x | x << 32) & 0x1f00000000ffffull;
  x = (x | x << 16) & 0x1f0000ff0000ffull;
  x = (x | x << 8) & 0x100f00f00f00f00full;
  x = (x | x << 4) & 0x10c30c30c30c30c3ull;
  x = (x | x << 2) & 0x1249249249249249ull;
  return x;
}

__global__ void __launch_bounds__(256) k_morton(const GpuBuildItem* __restrict__ items, int n, const int* __restrict__ bounds,
                                                unsigned long long* __restrict__ keys, int* __restrict__ vals) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const Box3 b = items[i].box;
  unsigned long long code = 0;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const float lo = ordered_as_float(bounds[k]), hi = ordered_as_float(bounds[3 + k]);
    const float ext = hi - lo;
    const float c = 0.5f * (b.lo[k] + b.hi[k]);
    float u = ext > 0.0f ? (c - lo) / ext : 0.0f;
    u = fminf(fmaxf(u, 0.0f), 1.0f);
    const unsigned long long q = static_cast<unsigned long long>(fminf(u * 2097152.0f, 2097151.0f));
    code |= spread21(q) << k;
  }
  keys[i] = code;
  vals[i] = i;
}

// common prefix of the keys at sorted positions i and j (-1 outside the array); equal keys are told apart by their positions
__device__ __forceinline__ int delta(const unsigned long long* __restrict__ keys, int n, int i, int j) {
  if (j < 0 || j >= n) return -1;
  const unsigned long long a = keys[i], b = keys[j];
  if (a == b) return 64 + __clz(i ^ j);
  return __clzll(a ^ b);
}

// Karras 2012, one thread per internal node: children codes (>= 0: internal index, < 0: ~sorted leaf position) and parent links
__global__ void __launch_bounds__(256) k_hierarchy(const unsigned long long* __restrict__ keys, int n, int2* __restrict__ children,
                                                   int* __restrict__ parent_inner, int* __restrict__ parent_leaf, int* __restrict__ prefix,
                                                   int* __restrict__ range_other) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n - 1) return;
  const int d = delta(keys, n, i, i + 1) - delta(keys, n, i, i - 1) >= 0 ? 1 : -1;
  const int dmin = delta(keys, n, i, i - d);
  int lmax = 2;
  while (delta(keys, n, i, i + lmax * d) > dmin) lmax <<= 1;
  int l = 0;
  for (int t = lmax >> 1; t >= 1; t >>= 1)
    if (delta(keys, n, i, i + (l + t) * d) > dmin) l += t;
  const int j = i + l * d;
  const int dnode = delta(keys, n, i, j);
  int s = 0;
  for (int t = (l + 1) >> 1;; t = (t + 1) >> 1) {
    if (delta(keys, n, i, i + (s + t) * d) > dnode) s += t;
    if (t == 1) break;
  }
  const int gamma = i + s * d + min(d, 0);
  const int lo = min(i, j), hi = max(i, j);
  const int left = lo == gamma ? ~gamma : gamma;
  const int right = hi == gamma + 1 ? ~(gamma + 1) : gamma + 1;
  children[i] = make_int2(left, right);
  prefix[i] = dnode;   // leading bits shared by every key below this node
  range_other[i] = j;  // the node covers the sorted positions [min(i, j), max(i, j)]
  if (left >= 0) parent_inner[left] = i; else parent_leaf[gamma] = i;
  if (right >= 0) parent_inner[right] = i; else parent_leaf[gamma + 1] = i;
  if (i == 0) parent_inner[0] = -1;
}

struct NodeBox { float lo[3], hi[3]; };

// bottom-up bounds: the second thread to reach a node has both children's boxes (read past L1: another SM may have written them)
__global__ void __launch_bounds__(256) k_refit(const GpuBuildItem* __restrict__ items, const int* __restrict__ order, int n, const int2* __restrict__ children,
                                               const int* __restrict__ parent_inner, const int* __restrict__ parent_leaf, NodeBox* boxes, int* visits) {
  const int leaf = blockIdx.x * blockDim.x + threadIdx.x;
  if (leaf >= n) return;
  int node = parent_leaf[leaf];
  while (node >= 0) {
    if (atomicAdd(visits + node, 1) == 0) break;   // first arrival: the sibling subtree is not done yet
    __threadfence();
    const int2 ch = children[node];
    NodeBox b;
#pragma unroll
    for (int side = 0; side < 2; ++side) {
      const int c = side ? ch.y : ch.x;
      float lo[3], hi[3];
      if (c < 0) {
        const Box3 ib = items[order[~c]].box;
        for (int k = 0; k < 3; ++k) { lo[k] = ib.lo[k]; hi[k] = ib.hi[k]; }
      } else {
        for (int k = 0; k < 3; ++k) { lo[k] = __ldcg(&boxes[c].lo[k]); hi[k] = __ldcg(&boxes[c].hi[k]); }
      }
      for (int k = 0; k < 3; ++k) {
        b.lo[k] = side ? fminf(b.lo[k], lo[k]) : lo[k];
        b.hi[k] = side ? fmaxf(b.hi[k], hi[k]) : hi[k];
      }
    }
    for (int k = 0; k < 3; ++k) { __stcg(&boxes[node].lo[k], b.lo[k]); __stcg(&boxes[node].hi[k], b.hi[k]); }
    __threadfence();
    node = parent_inner[node];
  }
}

// depth of every leaf (walk to the root): the traversal stack of the kernels must hold the deepest path
__global__ void __launch_bounds__(256) k_depth(int n, const int* __restrict__ parent_inner, const int* __restrict__ parent_leaf, int* __restrict__ max_depth) {
  const int leaf = blockIdx.x * blockDim.x + threadIdx.x;
  int depth = 0;
  if (leaf < n) {
    for (int node = parent_leaf[leaf]; node >= 0; node = parent_inner[node]) ++depth;
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) depth = max(depth, __shfl_xor_sync(0xffffffffu, depth, off));
  if ((threadIdx.x & 31) == 0 && depth > 0) atomicMax(max_depth, depth);
}


// ---- SAH over the top of the tree ---------------------------------------------------------------------------------------------------
// The radix tree splits space at the Morton midpoints, which is what makes it 10-20 % slower to trace than a SAH tree; most of that
// loss sits in the top levels, which every ray walks.  So the top is rebuilt: the nodes whose keys share fewer than P bits ("top
// nodes") are discarded, the subtrees hanging below them ("clusters": a few thousand) keep their radix structure, and the host builds a
// binned-SAH tree over the cluster boxes (a millisecond) whose nodes go back into the slots of the discarded top nodes.
constexpr int kTopLevels = 4;   // candidate cut depths: 3 * (level + 3) + 1 key bits, i.e. octree levels 3..6
// counts[l] <- number of top nodes for cut l (clusters = top nodes + 1)
__global__ void __launch_bounds__(256) k_count_top(const int* __restrict__ prefix, int n, int* __restrict__ counts) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  int c[kTopLevels];
#pragma unroll
  for (int l = 0; l < kTopLevels; ++l) c[l] = (i < n - 1 && prefix[i] < 3 * (l + 3) + 1) ? 1 : 0;
#pragma unroll
  for (int l = 0; l < kTopLevels; ++l) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) c[l] += __shfl_xor_sync(0xffffffffu, c[l], off);
    if ((threadIdx.x & 31) == 0 && c[l]) atomicAdd(counts + l, c[l]);
  }
}
struct Cluster { float lo[3], hi[3]; int code; };   // code: >= 0 internal node of the radix tree, < 0: ~(primitive reference | direct mark)
// top nodes (prefix < P) -> top[], their non-top children -> clusters[]; counters[0] = #top, counters[1] = #clusters
__global__ void __launch_bounds__(256) k_collect_top(const GpuBuildItem* __restrict__ items, const int* __restrict__ order, const int* __restrict__ prefix, int n, int P,
                                                     const int2* __restrict__ children, const NodeBox* __restrict__ boxes, int* __restrict__ top,
                                                     Cluster* __restrict__ clusters, int* __restrict__ counters, uint8_t* __restrict__ is_top) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n - 1) return;
  const bool t = prefix[i] < P;
  is_top[i] = t ? 1 : 0;
  if (!t) return;
  top[atomicAdd(counters, 1)] = i;
  const int2 ch = children[i];
#pragma unroll
  for (int side = 0; side < 2; ++side) {
    const int c = side ? ch.y : ch.x;
    if (c >= 0 && prefix[c] < P) continue;   // another top node
    Cluster cl;
    if (c < 0) {
      const GpuBuildItem it = items[order[~c]];
      for (int k = 0; k < 3; ++k) { cl.lo[k] = it.box.lo[k]; cl.hi[k] = it.box.hi[k]; }
      cl.code = static_cast<int>(~(it.ref | BvhBuilder::kDirectMark));
    } else {
      const NodeBox b = boxes[c];
      for (int k = 0; k < 3; ++k) { cl.lo[k] = b.lo[k]; cl.hi[k] = b.hi[k]; }
      cl.code = c;
    }
    clusters[atomicAdd(counters + 1, 1)] = cl;
  }
}
// deepest path from a cluster root down to a leaf (inner nodes), over all clusters
__global__ void __launch_bounds__(256) k_depth_below_top(int n, const int* __restrict__ parent_inner, const int* __restrict__ parent_leaf,
                                                         const uint8_t* __restrict__ is_top, int* __restrict__ max_depth) {
  const int leaf = blockIdx.x * blockDim.x + threadIdx.x;
  int depth = 0;
  if (leaf < n) {
    for (int node = parent_leaf[leaf]; node >= 0 && !is_top[node]; node = parent_inner[node]) ++depth;
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) depth = max(depth, __shfl_xor_sync(0xffffffffu, depth, off));
  if ((threadIdx.x & 31) == 0 && depth > 0) atomicMax(max_depth, depth);
}
struct SlotNode { int slot; int pad[3]; PackedNode node; };
__global__ void __launch_bounds__(256) k_scatter_nodes(const SlotNode* __restrict__ src, int count, PackedNode* __restrict__ nodes) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) nodes[src[i].slot] = src[i].node;
}

// centre + half-extent of a box exactly as BvhBuilder::centre_extent does on the host (half-extent rounded up: conservative)
__device__ __forceinline__ void centre_extent_dev(const float lo[3], const float hi[3], float c[3], float e[3]) {
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const double l = lo[k], h = hi[k];
    c[k] = static_cast<float>(0.5 * (l + h));
    const double m = fmax(fabs(l), fabs(h));
    const double need = fmax(h - static_cast<double>(c[k]), static_cast<double>(c[k]) - l) + 4.0 * 1.1920929e-7 * m + 1e-30;
    float ef = static_cast<float>(need);
    if (static_cast<double>(ef) < need) ef = __int_as_float(__float_as_int(ef) + 1);   // next float up (ef > 0)
    e[k] = ef;
  }
}

__global__ void __launch_bounds__(256) k_pack(const GpuBuildItem* __restrict__ items, const int* __restrict__ order, int n, const int2* __restrict__ children,
                                              const NodeBox* __restrict__ boxes, PackedNode* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n - 1) return;
  const int2 ch = children[i];
  PackedNode nd;
  float c[3], e[3];
  {
    int code;
    if (ch.x < 0) { const GpuBuildItem it = items[order[~ch.x]]; centre_extent_dev(it.box.lo, it.box.hi, c, e); code = static_cast<int>(~(it.ref | BvhBuilder::kDirectMark)); }
    else { const NodeBox b = boxes[ch.x]; centre_extent_dev(b.lo, b.hi, c, e); code = ch.x; }
    for (int k = 0; k < 3; ++k) { nd.c[k][0] = c[k]; nd.e[k][0] = e[k]; } nd.left = code;
  }
  {
    int code;
    if (ch.y < 0) { const GpuBuildItem it = items[order[~ch.y]]; centre_extent_dev(it.box.lo, it.box.hi, c, e); code = static_cast<int>(~(it.ref | BvhBuilder::kDirectMark)); }
    else { const NodeBox b = boxes[ch.y]; centre_extent_dev(b.lo, b.hi, c, e); code = ch.y; }
    for (int k = 0; k < 3; ++k) { nd.c[k][1] = c[k]; nd.e[k][1] = e[k]; } nd.right = code;
  }
  nd.pad0 = 0; nd.pad1 = 0;
  out[i] = nd;
}


// ---- SAH inside the clusters ---------------------------------------------------------------------------------------------------------
// What is left of the radix tree after the top has been rebuilt are the clusters: a few thousand subtrees of ~60 primitives each, still
// split at Morton midpoints.  One WARP per cluster rebuilds its subtree with binned SAH (16 bins on each of the three axes, all in
// shared memory) and writes it over the cluster's own node slots: in Karras' numbering a node i covering the sorted positions [a, b] is
// i = a or i = b, and the b - a inner nodes below and including it are exactly the slots [a, b - 1] (i = a) or [a + 1, b] (i = b).  The
// cluster root keeps its slot, so nothing above it changes.  Clusters beyond kSahMax primitives keep their radix structure.
constexpr int kSahWarps = 4;     // warps per CTA (42 KB of static shared memory)
constexpr int kSahMax = 1024;    // largest cluster that is rebuilt
constexpr int kSahBins = 16;
struct SahWarp {
  int idx[kSahMax], tmp[kSahMax];    // item indices of the cluster, permuted in place range by range
  int bins[3][kSahBins][7];          // per axis and bin: box lo[3], hi[3] (ordered ints), count
  int4 stack[48];                    // pending ranges: (lo, hi, local node index, depth)
};
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, off));
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, off));
  return v;
}
__device__ __forceinline__ float half_area3(const float lo[3], const float hi[3]) {
  const float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
  return dx * dy + dy * dz + dz * dx;
}
// median_from: ranges at this depth or deeper are split in the middle of their (Morton-ordered) list, which bounds the depth of a
// cluster's subtree by median_from + log2(kSahMax) whatever the geometry looks like
__global__ void __launch_bounds__(kSahWarps * 32) k_sah_clusters(const GpuBuildItem* __restrict__ items, const int* __restrict__ order,
                                                                const int* __restrict__ range_other, const Cluster* __restrict__ clusters, int n_clusters,
                                                                int median_from, PackedNode* __restrict__ nodes, int* __restrict__ stats) {
  __shared__ SahWarp sm[kSahWarps];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int cid = blockIdx.x * kSahWarps + w;
  if (cid >= n_clusters) return;   // whole warps leave; only __syncwarp below
  const int c = clusters[cid].code;
  if (c < 0) return;               // a single primitive
  const int j = range_other[c];
  const int a = min(c, j), b = max(c, j), m = b - a + 1;
  if (m < 3 || m > kSahMax) { if (lane == 0 && m > kSahMax) atomicAdd(stats + 1, 1); return; }
  SahWarp& S = sm[w];
  const int s0 = c == a ? a : a + 1;
  auto slot = [&](int local) { return c == a ? s0 + local : (local == 0 ? c : s0 + local - 1); };
  for (int i = lane; i < m; i += 32) S.idx[i] = order[a + i];
  int sp = 0, deepest = 1;
  if (lane == 0) S.stack[0] = make_int4(0, m, 0, 1);
  sp = 1;
  __syncwarp();
  while (sp > 0) {
    const int4 top = S.stack[--sp];
    __syncwarp();
    const int lo = top.x, hi = top.y, L = top.z, depth = top.w, k = hi - lo;
    deepest = max(deepest, depth);
    float llo[3], lhi[3], rlo[3], rhi[3];
    int nl = 0;
    bool split_done = false;
    if (k > 2 && depth < median_from) {
      // centroid bounds of the range
      float cl[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, ch[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
      for (int i = lo + lane; i < hi; i += 32) {
        const Box3 bx = items[S.idx[i]].box;
#pragma unroll
        for (int q = 0; q < 3; ++q) { const float cc = 0.5f * (bx.lo[q] + bx.hi[q]); cl[q] = fminf(cl[q], cc); ch[q] = fmaxf(ch[q], cc); }
      }
      float scale[3];
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        cl[q] = warp_min(cl[q]); ch[q] = warp_max(ch[q]);
        scale[q] = ch[q] > cl[q] ? static_cast<float>(kSahBins) * 0.999f / (ch[q] - cl[q]) : 0.0f;
      }
      for (int i = lane; i < 3 * kSahBins * 7; i += 32) {
        const int f = i % 7;
        (&S.bins[0][0][0])[i] = f < 3 ? 0x7f7fffff : (f < 6 ? static_cast<int>(0xff7fffffu ^ 0x7fffffffu) : 0);   // +FLT_MAX, ordered(-FLT_MAX), 0
      }
      __syncwarp();
      for (int i = lo + lane; i < hi; i += 32) {
        const Box3 bx = items[S.idx[i]].box;
#pragma unroll
        for (int q = 0; q < 3; ++q) {
          const int bn = min(kSahBins - 1, static_cast<int>((0.5f * (bx.lo[q] + bx.hi[q]) - cl[q]) * scale[q]));
          int* B = S.bins[q][bn];
#pragma unroll
          for (int r = 0; r < 3; ++r) { atomicMin(B + r, float_as_ordered(bx.lo[r])); atomicMax(B + 3 + r, float_as_ordered(bx.hi[r])); }
          atomicAdd(B + 6, 1);
        }
      }
      __syncwarp();
      // lane = (axis, split position): cost of putting bins [0, s] left and (s, kSahBins) right
      float cost = FLT_MAX;
      int my_nl = 0;
      float xl[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, xh[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX}, yl[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, yh[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
      for (int cand = lane; cand < 3 * (kSahBins - 1); cand += 32) {
        const int ax = cand / (kSahBins - 1), s = cand - ax * (kSahBins - 1);
        float tl[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, th[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX}, ul[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, uh[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
        int cntl = 0, cntr = 0;
        for (int bn = 0; bn < kSahBins; ++bn) {
          const int* B = S.bins[ax][bn];
          const int cnt = B[6];
          if (cnt == 0) continue;
          if (bn <= s) { cntl += cnt; for (int r = 0; r < 3; ++r) { tl[r] = fminf(tl[r], ordered_as_float(B[r])); th[r] = fmaxf(th[r], ordered_as_float(B[3 + r])); } }
          else { cntr += cnt; for (int r = 0; r < 3; ++r) { ul[r] = fminf(ul[r], ordered_as_float(B[r])); uh[r] = fmaxf(uh[r], ordered_as_float(B[3 + r])); } }
        }
        if (cntl > 0 && cntr > 0) {
          const float cst = half_area3(tl, th) * cntl + half_area3(ul, uh) * cntr;
          if (cst < cost) {
            cost = cst; my_nl = cntl | (cand << 16);
            for (int r = 0; r < 3; ++r) { xl[r] = tl[r]; xh[r] = th[r]; yl[r] = ul[r]; yh[r] = uh[r]; }
          }
        }
      }
      const float best = warp_min(cost);
      if (best < FLT_MAX) {
        const unsigned who = __ballot_sync(0xffffffffu, cost == best);
        const int src = __ffs(who) - 1;
        const int packed = __shfl_sync(0xffffffffu, my_nl, src);
        nl = packed & 0xffff;
        const int cand = packed >> 16, ax = cand / (kSahBins - 1), s = cand - ax * (kSahBins - 1);
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          llo[r] = __shfl_sync(0xffffffffu, xl[r], src); lhi[r] = __shfl_sync(0xffffffffu, xh[r], src);
          rlo[r] = __shfl_sync(0xffffffffu, yl[r], src); rhi[r] = __shfl_sync(0xffffffffu, yh[r], src);
        }
        // stable partition of idx[lo, hi) by "bin on the chosen axis <= s" through tmp
        int nleft = 0, nright = 0;
        for (int base = lo; base < hi; base += 32) {
          const int i = base + lane;
          bool valid = i < hi, left = false;
          int id = 0;
          if (valid) {
            id = S.idx[i];
            const Box3 bx = items[id].box;
            const float cc = 0.5f * (bx.lo[ax] + bx.hi[ax]);
            const float lo_ax = ax == 0 ? cl[0] : (ax == 1 ? cl[1] : cl[2]), sc_ax = ax == 0 ? scale[0] : (ax == 1 ? scale[1] : scale[2]);
            left = min(kSahBins - 1, static_cast<int>((cc - lo_ax) * sc_ax)) <= s;
          }
          const unsigned ml = __ballot_sync(0xffffffffu, valid && left), mr = __ballot_sync(0xffffffffu, valid && !left);
          const unsigned below = (1u << lane) - 1u;
          if (valid && left) S.tmp[lo + nleft + __popc(ml & below)] = id;
          if (valid && !left) S.tmp[lo + nl + nright + __popc(mr & below)] = id;
          nleft += __popc(ml); nright += __popc(mr);
        }
        __syncwarp();
        for (int i = lo + lane; i < hi; i += 32) S.idx[i] = S.tmp[i];
        __syncwarp();
        split_done = true;
      }
    }
    if (!split_done) {   // two primitives, coincident centroids, or past the depth guard: split the list in the middle
      nl = k / 2;
      float tl[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, th[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX}, ul[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, uh[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
      for (int i = lo + lane; i < hi; i += 32) {
        const Box3 bx = items[S.idx[i]].box;
        if (i < lo + nl) { for (int r = 0; r < 3; ++r) { tl[r] = fminf(tl[r], bx.lo[r]); th[r] = fmaxf(th[r], bx.hi[r]); } }
        else { for (int r = 0; r < 3; ++r) { ul[r] = fminf(ul[r], bx.lo[r]); uh[r] = fmaxf(uh[r], bx.hi[r]); } }
      }
#pragma unroll
      for (int r = 0; r < 3; ++r) { llo[r] = warp_min(tl[r]); lhi[r] = warp_max(th[r]); rlo[r] = warp_min(ul[r]); rhi[r] = warp_max(uh[r]); }
    }
    const int nr = k - nl;
    const int lL = L + 1, lR = L + nl;   // local preorder numbering: the left subtree owns nl - 1 nodes
    if (lane == 0) {
      PackedNode nd;
      float cc[3], ee[3];
      centre_extent_dev(llo, lhi, cc, ee);
      for (int r = 0; r < 3; ++r) { nd.c[r][0] = cc[r]; nd.e[r][0] = ee[r]; }
      centre_extent_dev(rlo, rhi, cc, ee);
      for (int r = 0; r < 3; ++r) { nd.c[r][1] = cc[r]; nd.e[r][1] = ee[r]; }
      nd.left = nl > 1 ? slot(lL) : static_cast<int>(~(items[S.idx[lo]].ref | BvhBuilder::kDirectMark));
      nd.right = nr > 1 ? slot(lR) : static_cast<int>(~(items[S.idx[lo + nl]].ref | BvhBuilder::kDirectMark));
      nd.pad0 = 0; nd.pad1 = 0;
      nodes[slot(L)] = nd;
      int q = sp;
      if (nr > 1) S.stack[q++] = make_int4(lo + nl, hi, lR, depth + 1);
      if (nl > 1) S.stack[q++] = make_int4(lo, lo + nl, lL, depth + 1);
    }
    sp += (nr > 1 ? 1 : 0) + (nl > 1 ? 1 : 0);
    __syncwarp();
  }
  if (lane == 0) { atomicMax(stats, deepest); atomicAdd(stats + 2, 1); }
}

}  // namespace

namespace {
template <typename T> struct Carved { T* p = nullptr; };   // a slice of the scratch allocation
struct BuildScratch { std::mutex m; DevBuf<unsigned char> pool; };
BuildScratch g_scratch[64];   // per device; builds on one device take turns
}  // namespace
void release_build_scratch() {   // rtw_release_cached_buffers
  int cur = 0;
  cudaGetDevice(&cur);
  for (int d = 0; d < 64; ++d) {
    std::lock_guard<std::mutex> lock(g_scratch[d].m);
    if (g_scratch[d].pool.p) { cudaSetDevice(d); g_scratch[d].pool.alloc(0); }
  }
  cudaSetDevice(cur);
}

// items: n >= 2 build records in host memory; nodes_out: device memory for n - 1 PackedNodes.  Runs on `stream`, returns after the
// build has finished.  sah_top: rebuild the top of the radix tree with the host's SAH builder (a few thousand clusters).
// sah_clusters: then rebuild every subtree below that top with binned SAH on the device, one warp each (k_sah_clusters).
// *depth_out <- bound on the deepest root-to-leaf path (inner nodes); *top_nodes_out <- nodes replaced by the SAH top (0: none);
// *clusters_rebuilt_out <- subtrees rebuilt on the device.
int gpu_build_bvh(const void* items_host_v, size_t n, void* nodes_out_v, cudaStream_t stream, bool sah_top, bool sah_clusters, int* depth_out, double* build_ms,
                  int* top_nodes_out, int* clusters_rebuilt_out) {
  const GpuBuildItem* items_host = static_cast<const GpuBuildItem*>(items_host_v);
  PackedNode* nodes_out = static_cast<PackedNode*>(nodes_out_v);
  if (n < 2 || n >= (size_t(1) << 30)) return fail("gpu_build_bvh: primitive count out of range");
  const int ni = static_cast<int>(n);
  // every array of the build comes out of ONE grow-only allocation per device (BuildScratch): a build needs fifteen of them, and
  // allocating and freeing ~400 MB per call cost several times the build itself (47 against 80-130 ms per cold call of the 991k-triangle
  // mesh while the allocator was still settling)
  int dev = 0;
  RTW_CUDA(cudaGetDevice(&dev));
  BuildScratch& scratch = g_scratch[dev & 63];
  std::lock_guard<std::mutex> scratch_lock(scratch.m);
  size_t temp_bytes = 0;
  RTW_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, temp_bytes, static_cast<const unsigned long long*>(nullptr), static_cast<unsigned long long*>(nullptr),
                                           static_cast<const int*>(nullptr), static_cast<int*>(nullptr), ni, 0, 63, stream));
  Carved<GpuBuildItem> d_items;
  Carved<unsigned long long> d_keys, d_keys_sorted;
  Carved<int> d_vals, d_order, d_parent_inner, d_parent_leaf, d_visits, d_misc, d_prefix, d_range_other, d_counts, d_top;
  Carved<int2> d_children;
  Carved<NodeBox> d_boxes;
  Carved<unsigned char> d_temp;
  Carved<Cluster> d_clusters;
  Carved<uint8_t> d_is_top;
  Carved<SlotNode> d_slot_nodes;
  constexpr size_t kMaxTop = 16384;   // the SAH top is cut where at most this many subtrees hang below it
  auto layout = [&](unsigned char* base) {
    size_t off = 0;
    auto take = [&](auto& buf, size_t count) {
      using T = std::remove_pointer_t<decltype(buf.p)>;
      off = (off + 255) & ~size_t(255);
      buf.p = base ? reinterpret_cast<T*>(base + off) : nullptr;   // first pass (base == nullptr) only sizes the allocation
      off += count * sizeof(T);
    };
    take(d_items, n); take(d_keys, n); take(d_keys_sorted, n); take(d_vals, n); take(d_order, n);
    take(d_parent_inner, n); take(d_parent_leaf, n); take(d_visits, n); take(d_misc, 8);
    take(d_children, n); take(d_boxes, n); take(d_prefix, n); take(d_range_other, n); take(d_temp, temp_bytes);
    take(d_counts, kTopLevels + 8); take(d_top, kMaxTop); take(d_clusters, kMaxTop + 1); take(d_is_top, n); take(d_slot_nodes, kMaxTop);
    return off;
  };
  const size_t scratch_bytes = layout(nullptr);
  RTW_CUDA(scratch.pool.reserve(scratch_bytes));
  layout(scratch.pool.p);
  RTW_CUDA(cudaMemcpyAsync(d_items.p, items_host, n * sizeof(GpuBuildItem), cudaMemcpyHostToDevice, stream));
  EventPair ev;
  RTW_CUDA(ev.create());
  RTW_CUDA(cudaEventRecord(ev.a, stream));
  const int init[8] = {0x7f7fffff, 0x7f7fffff, 0x7f7fffff, static_cast<int>(0xff7fffffu ^ 0x7fffffffu), static_cast<int>(0xff7fffffu ^ 0x7fffffffu),
                       static_cast<int>(0xff7fffffu ^ 0x7fffffffu), 0, 0};   // +FLT_MAX x3, ordered(-FLT_MAX) x3, max depth, spare
  RTW_CUDA(cudaMemcpyAsync(d_misc.p, init, sizeof init, cudaMemcpyHostToDevice, stream));
  RTW_CUDA(cudaMemsetAsync(d_visits.p, 0, n * sizeof(int), stream));
  const unsigned blocks = static_cast<unsigned>((n + 255) / 256);
  k_centroid_bounds<<<std::min(blocks, 1184u), 256, 0, stream>>>(d_items.p, ni, d_misc.p);
  count_launch();
  k_morton<<<blocks, 256, 0, stream>>>(d_items.p, ni, d_misc.p, d_keys.p, d_vals.p);
  count_launch();
  RTW_CUDA(cub::DeviceRadixSort::SortPairs(d_temp.p, temp_bytes, d_keys.p, d_keys_sorted.p, d_vals.p, d_order.p, ni, 0, 63, stream));
  k_hierarchy<<<blocks, 256, 0, stream>>>(d_keys_sorted.p, ni, d_children.p, d_parent_inner.p, d_parent_leaf.p, d_prefix.p, d_range_other.p);
  count_launch();
  k_refit<<<blocks, 256, 0, stream>>>(d_items.p, d_order.p, ni, d_children.p, d_parent_inner.p, d_parent_leaf.p, d_boxes.p, d_visits.p);
  count_launch();
  k_depth<<<blocks, 256, 0, stream>>>(ni, d_parent_inner.p, d_parent_leaf.p, d_misc.p + 6);
  count_launch();
  k_pack<<<blocks, 256, 0, stream>>>(d_items.p, d_order.p, ni, d_children.p, d_boxes.p, nodes_out);
  count_launch();
  RTW_CUDA(cudaGetLastError());
  int depth = 0;
  RTW_CUDA(cudaMemcpyAsync(&depth, d_misc.p + 6, sizeof depth, cudaMemcpyDeviceToHost, stream));

  // ---- SAH over the top of the tree (see k_count_top) ------------------------------------------------------------------------------
  int top_nodes = 0;
  if (sah_top && n >= 4096) {
    RTW_CUDA(cudaMemsetAsync(d_counts.p, 0, (kTopLevels + 8) * sizeof(int), stream));
    k_count_top<<<blocks, 256, 0, stream>>>(d_prefix.p, ni, d_counts.p);
    count_launch();
    int counts[kTopLevels];
    RTW_CUDA(cudaMemcpyAsync(counts, d_counts.p, sizeof counts, cudaMemcpyDeviceToHost, stream));
    RTW_CUDA(cudaStreamSynchronize(stream));
    int level = -1;   // the deepest cut with at most 16 384 clusters (and at least 64, else the radix tree is degenerate up there)
    int max_clusters = 16384;
    if (const char* e = std::getenv("RTW_LBVH_MAX_CLUSTERS")) max_clusters = std::max(64, std::atoi(e));   // tuning knob
    for (int l = 0; l < kTopLevels; ++l)
      if (counts[l] + 1 <= std::min(max_clusters, static_cast<int>(kMaxTop)) && counts[l] + 1 >= 64) level = l;
    if (level >= 0) {
      const int P = 3 * (level + 3) + 1, T = counts[level], Cn = T + 1;
      int* ctr = d_counts.p + kTopLevels;   // {#top, #clusters, depth below the top}
      k_collect_top<<<blocks, 256, 0, stream>>>(d_items.p, d_order.p, d_prefix.p, ni, P, d_children.p, d_boxes.p, d_top.p, d_clusters.p, ctr, d_is_top.p);
      count_launch();
      k_depth_below_top<<<blocks, 256, 0, stream>>>(ni, d_parent_inner.p, d_parent_leaf.p, d_is_top.p, ctr + 2);
      count_launch();
      std::vector<int> top(static_cast<size_t>(T));
      std::vector<Cluster> clusters(static_cast<size_t>(Cn));
      int got[3] = {0, 0, 0};
      RTW_CUDA(cudaMemcpyAsync(top.data(), d_top.p, top.size() * sizeof(int), cudaMemcpyDeviceToHost, stream));
      RTW_CUDA(cudaMemcpyAsync(clusters.data(), d_clusters.p, clusters.size() * sizeof(Cluster), cudaMemcpyDeviceToHost, stream));
      RTW_CUDA(cudaMemcpyAsync(got, ctr, sizeof got, cudaMemcpyDeviceToHost, stream));
      RTW_CUDA(cudaStreamSynchronize(stream));
      if (got[0] == T && got[1] == Cn) {
        std::sort(top.begin(), top.end());      // the root of the radix tree (node 0) stays the root: slot 0
        std::vector<BvhBuilder::Item> citems(clusters.size());
        for (size_t k = 0; k < clusters.size(); ++k) {
          for (int a = 0; a < 3; ++a) { citems[k].box.lo[a] = clusters[k].lo[a]; citems[k].box.hi[a] = clusters[k].hi[a]; }
          citems[k].ref = static_cast<uint32_t>(k);
        }
        std::vector<int> code_of(clusters.size());
        for (size_t k = 0; k < clusters.size(); ++k) code_of[k] = clusters[k].code;
        std::vector<PackedNode> tnodes(static_cast<size_t>(T));
        BvhBuilder tb;
        tb.build_items_direct(citems, tnodes.data());
        if (tb.max_depth() + got[2] <= kBvhStack) {
          std::vector<SlotNode> sn(static_cast<size_t>(T));
          auto remap = [&](int32_t code) -> int32_t {
            if (code >= 0) return top[static_cast<size_t>(code)];                                   // inner node of the new top -> its slot
            return code_of[(static_cast<uint32_t>(~code) & ~BvhBuilder::kDirectMark) & 0x1fffffffu];   // leaf of the new top -> the cluster
          };
          for (int j = 0; j < T; ++j) {
            sn[static_cast<size_t>(j)].slot = top[static_cast<size_t>(j)];
            sn[static_cast<size_t>(j)].node = tnodes[static_cast<size_t>(j)];
            sn[static_cast<size_t>(j)].node.left = remap(tnodes[static_cast<size_t>(j)].left);
            sn[static_cast<size_t>(j)].node.right = remap(tnodes[static_cast<size_t>(j)].right);
          }
          RTW_CUDA(cudaMemcpyAsync(d_slot_nodes.p, sn.data(), sn.size() * sizeof(SlotNode), cudaMemcpyHostToDevice, stream));
          k_scatter_nodes<<<static_cast<unsigned>((T + 255) / 256), 256, 0, stream>>>(d_slot_nodes.p, T, nodes_out);
          count_launch();
          RTW_CUDA(cudaGetLastError());
          depth = tb.max_depth() + got[2];
          top_nodes = T;
          // SAH inside the clusters (k_sah_clusters); the depth guard leaves room for the top above and log2(kSahMax) levels below it
          const int median_from = std::min(32, kBvhStack - tb.max_depth() - 11);
          if (sah_clusters && median_from >= 4) {
            int* st = ctr + 3;   // {deepest rebuilt cluster, clusters too big to rebuild, clusters rebuilt}: zeroed with d_counts... the three ints after ctr[2]
            k_sah_clusters<<<static_cast<unsigned>((Cn + kSahWarps - 1) / kSahWarps), kSahWarps * 32, 0, stream>>>(d_items.p, d_order.p, d_range_other.p, d_clusters.p, Cn,
                                                                                                                 median_from, nodes_out, st);
            count_launch();
            RTW_CUDA(cudaGetLastError());
            int cs[3] = {0, 0, 0};
            RTW_CUDA(cudaMemcpyAsync(cs, st, sizeof cs, cudaMemcpyDeviceToHost, stream));
            RTW_CUDA(cudaStreamSynchronize(stream));
            // clusters that kept their radix structure (too big) are bounded by got[2], the rebuilt ones by what the kernel saw
            depth = tb.max_depth() + (cs[1] > 0 ? std::max(got[2], cs[0]) : cs[0]);
            if (clusters_rebuilt_out) *clusters_rebuilt_out = cs[2];
          }
          RTW_CUDA(cudaStreamSynchronize(stream));
        }
      }
    }
  }
  RTW_CUDA(cudaEventRecord(ev.b, stream));
  RTW_CUDA(cudaStreamSynchronize(stream));
  float ms = 0.f;
  RTW_CUDA(cudaEventElapsedTime(&ms, ev.a, ev.b));
  if (build_ms) *build_ms = ms;
  if (depth_out) *depth_out = depth;
  if (top_nodes_out) *top_nodes_out = top_nodes;
  return 0;
}

}  // namespace rtw
