// rtw_build.cu -- BVH construction on the GPU (replaces the per-render host build of render.cpp:73-110 for big
// scenes when build time matters more than the last 10-20 % of tree quality).
//
// Linear BVH: 63-bit Morton codes of the primitive centroids, one radix sort, the binary radix tree of Karras ("Maximizing Parallelism
// in the Construction of BVHs, Octrees, and k-d Trees", HPG 2012) built in one pass, bounds fitted bottom-up with one atomic flag per
// node, and the result packed straight into the 64-byte two-child nodes the traversal kernels walk (child boxes as centre + half-extent
// rounded up exactly as the host builder does, rtw_bvh.h).  A million triangles take a few milliseconds instead of the ~80 ms of the
// binned-SAH build on 16 host cores; the tree is a plain spatial-median tree, so rays visit more nodes (measured in DESIGN.md).
// Any correct BVH returns the same closest hit: parity (primitive ids) does not depend on which builder ran.
#include <cub/device/device_radix_sort.cuh>

#include <cfloat>

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
                                                   int* __restrict__ parent_inner, int* __restrict__ parent_leaf) {
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
    nd.lc[0] = c[0]; nd.lc[1] = c[1]; nd.lc[2] = c[2]; nd.le_x = e[0]; nd.le_yz[0] = e[1]; nd.le_yz[1] = e[2]; nd.left = code;
  }
  {
    int code;
    if (ch.y < 0) { const GpuBuildItem it = items[order[~ch.y]]; centre_extent_dev(it.box.lo, it.box.hi, c, e); code = static_cast<int>(~(it.ref | BvhBuilder::kDirectMark)); }
    else { const NodeBox b = boxes[ch.y]; centre_extent_dev(b.lo, b.hi, c, e); code = ch.y; }
    nd.rc_xy[0] = c[0]; nd.rc_xy[1] = c[1]; nd.rc_z = c[2]; nd.re[0] = e[0]; nd.re[1] = e[1]; nd.re[2] = e[2]; nd.right = code;
  }
  nd.pad0 = 0; nd.pad1 = 0;
  out[i] = nd;
}

}  // namespace

// items: n >= 2 build records in host memory; nodes_out: device memory for n - 1 PackedNodes.  Runs on `stream`, returns after the
// build has finished.  *depth_out <- deepest root-to-leaf path (inner nodes).
int gpu_build_bvh(const void* items_host_v, size_t n, void* nodes_out_v, cudaStream_t stream, int* depth_out, double* build_ms) {
  const GpuBuildItem* items_host = static_cast<const GpuBuildItem*>(items_host_v);
  PackedNode* nodes_out = static_cast<PackedNode*>(nodes_out_v);
  if (n < 2 || n >= (size_t(1) << 30)) return fail("gpu_build_bvh: primitive count out of range");
  const int ni = static_cast<int>(n);
  DevBuf<GpuBuildItem> d_items;
  DevBuf<unsigned long long> d_keys, d_keys_sorted;
  DevBuf<int> d_vals, d_order, d_parent_inner, d_parent_leaf, d_visits, d_misc;
  DevBuf<int2> d_children;
  DevBuf<NodeBox> d_boxes;
  DevBuf<unsigned char> d_temp;
  RTW_CUDA(d_items.alloc(n)); RTW_CUDA(d_keys.alloc(n)); RTW_CUDA(d_keys_sorted.alloc(n)); RTW_CUDA(d_vals.alloc(n)); RTW_CUDA(d_order.alloc(n));
  RTW_CUDA(d_parent_inner.alloc(n)); RTW_CUDA(d_parent_leaf.alloc(n)); RTW_CUDA(d_visits.alloc(n)); RTW_CUDA(d_misc.alloc(8));
  RTW_CUDA(d_children.alloc(n)); RTW_CUDA(d_boxes.alloc(n));
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
  size_t temp_bytes = 0;
  RTW_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, temp_bytes, d_keys.p, d_keys_sorted.p, d_vals.p, d_order.p, ni, 0, 63, stream));
  RTW_CUDA(d_temp.alloc(temp_bytes));
  RTW_CUDA(cub::DeviceRadixSort::SortPairs(d_temp.p, temp_bytes, d_keys.p, d_keys_sorted.p, d_vals.p, d_order.p, ni, 0, 63, stream));
  k_hierarchy<<<blocks, 256, 0, stream>>>(d_keys_sorted.p, ni, d_children.p, d_parent_inner.p, d_parent_leaf.p);
  count_launch();
  k_refit<<<blocks, 256, 0, stream>>>(d_items.p, d_order.p, ni, d_children.p, d_parent_inner.p, d_parent_leaf.p, d_boxes.p, d_visits.p);
  count_launch();
  k_depth<<<blocks, 256, 0, stream>>>(ni, d_parent_inner.p, d_parent_leaf.p, d_misc.p + 6);
  count_launch();
  k_pack<<<blocks, 256, 0, stream>>>(d_items.p, d_order.p, ni, d_children.p, d_boxes.p, nodes_out);
  count_launch();
  RTW_CUDA(cudaGetLastError());
  RTW_CUDA(cudaEventRecord(ev.b, stream));
  int misc[8];
  RTW_CUDA(cudaMemcpyAsync(misc, d_misc.p, sizeof misc, cudaMemcpyDeviceToHost, stream));
  RTW_CUDA(cudaStreamSynchronize(stream));
  float ms = 0.f;
  RTW_CUDA(cudaEventElapsedTime(&ms, ev.a, ev.b));
  if (build_ms) *build_ms = ms;
  if (depth_out) *depth_out = misc[6];
  return 0;
}

}  // namespace rtw
