// rtw_bvh.h -- host-side binned-SAH BVH builder producing the compact 64-byte two-child nodes the K2 kernel
// walks.  Replaces the reference's median-split BVHNode constructor (render.cpp:73-110), whose quality the
// survey measured at 63-108 node visits per ray; only the closest-hit semantics are kept.
#pragma once
#include <algorithm>
#include <array>
#include <atomic>
#include <cmath>
#include <future>
#include <cstdint>
#include <cstring>
#include <limits>
#include <thread>
#include <vector>

namespace rtw {

struct Box3 {
  float lo[3], hi[3];
  void reset() { for (int k = 0; k < 3; ++k) { lo[k] = std::numeric_limits<float>::infinity(); hi[k] = -std::numeric_limits<float>::infinity(); } }
  void grow(const Box3& b) { for (int k = 0; k < 3; ++k) { lo[k] = std::min(lo[k], b.lo[k]); hi[k] = std::max(hi[k], b.hi[k]); } }
  void grow(const float p[3]) { for (int k = 0; k < 3; ++k) { lo[k] = std::min(lo[k], p[k]); hi[k] = std::max(hi[k], p[k]); } }
  float half_area() const {
    const float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
    return (dx < 0 || dy < 0 || dz < 0) ? 0.0f : dx * dy + dy * dz + dz * dx;
  }
};

struct PackedNode {  // 4 x float4, see DevScene::nodes: each child box as centre + half-extent, the two children of an axis side by side
  float c[3][2];   // c[axis][child]: (left, right) pairs, so that one packed FFMA2 (fma.rn.f32x2, sm_100) works on both children
  float e[3][2];   // half-extents, same pairing
  int32_t left, right, pad0, pad1;
};
constexpr float kEmptyChildCentre = 3.0e38f;  // centre of a never-entered filler child (half-extent 0)
static_assert(sizeof(PackedNode) == 64, "node must be 64 bytes");

// Binary tree with exact child boxes and subtree sizes: the input of the wide-tree collapse (CwBuilder below).  Host only.
struct BinNode {
  Box3 box[2];
  int32_t child[2];   // >= 0: inner node index; < 0: ~(reference | kDirectMark) of a single primitive
  uint32_t count[2];  // primitives below each child
};

class BvhBuilder {
 public:
  struct Item { Box3 box; float c[3]; uint32_t ref; };  // one primitive: bounds, centroid, encoded reference
  int kMaxLeaf = 1;  // primitives per leaf (<= 31); 1 = primitive reference stored in the child code
  size_t single_axis_below = 0;  // ranges with fewer primitives than this bin only their widest centroid axis (0: always all three)
  static constexpr int kBins = 32;   // 16 -> 32: cover scene +2 %, 991k-triangle mesh +1 % (node visits), build time unchanged
  static constexpr int kSahDepthLimit = 32;  // see direct_node
  static constexpr size_t kTaskRange = 16384; // subtrees over more primitives than this are built by their own task
  static constexpr size_t kParallelRange = 1u << 17;  // ranges this big split their own passes over host threads
  template <typename F>
  static void for_chunks(size_t begin, size_t end, int nt, F&& fn) {  // fn(chunk, b, e) on nt threads
    const size_t n = end - begin;
    std::vector<std::future<void>> tasks;
    for (int t = 1; t < nt; ++t) tasks.push_back(std::async(std::launch::async, [&fn, t, begin, n, nt] { fn(t, begin + n * t / nt, begin + n * (t + 1) / nt); }));
    fn(0, begin, begin + n / nt);
    for (auto& t : tasks) t.get();
  }
  // references are (kind << 30) | index with index < 2^28; bit 29 marks a direct leaf so that its code ~ref is never -1
  static constexpr uint32_t kDirectMark = 1u << 29;

  // boxes[i] / refs[i]: bounds and encoded reference ((kind << 30) | index) of primitive i
  void build(const std::vector<Box3>& boxes, const std::vector<uint32_t>& refs) {
    nodes_.clear(); leaf_refs_.clear();
    const size_t n = boxes.size();
    if (n == 0) return;
    boxes_ = &boxes; refs_ = &refs;
    if (kMaxLeaf == 1 && n >= 2) { build_direct(boxes, refs); return; }
    order_.resize(n);
    cent_.resize(3 * n);
    for (size_t i = 0; i < n; ++i) {
      order_[i] = static_cast<uint32_t>(i);
      for (int k = 0; k < 3; ++k) cent_[3 * i + k] = 0.5f * (boxes[i].lo[k] + boxes[i].hi[k]);
    }
    nodes_.reserve(2 * n / kMaxLeaf + 16);
    leaf_refs_.reserve(n);
    Box3 rb; rb.reset();
    for (size_t i = 0; i < n; ++i) rb.grow(boxes[i]);
    if (n <= static_cast<size_t>(kMaxLeaf)) {
      // a single leaf still needs a root node: left = the leaf, right = empty leaf with an inverted box
      PackedNode nd{};
      set_child(nd, 0, rb, make_leaf(0, n));
      Box3 e; e.reset();
      set_child(nd, 1, e, empty_leaf());
      nodes_.push_back(nd);
      return;
    }
    nodes_.push_back(PackedNode{});
    build_node(0, 0, n, rb, 0);
  }

  // a child that is never entered (inverted box); its code must still decode harmlessly
  int32_t empty_leaf() const { return kMaxLeaf == 1 ? static_cast<int32_t>(~kDirectMark) : ~0; }
  // Same, for a caller that already holds the records and the destination of the nodes (max(n - 1, 1) entries):
  // single-primitive leaves only.  `items` is reordered.  Returns the number of nodes written.
  // The same tree as BinNodes (exact boxes, subtree sizes) for the wide-tree collapse: n - 1 entries for n >= 2 items.
  size_t build_items_binary(std::vector<Item>& items, BinNode* out) {
    const size_t n = items.size();
    if (n < 2) return 0;
    for (Item& it : items)
      for (int k = 0; k < 3; ++k) it.c[k] = 0.5f * (it.box.lo[k] + it.box.hi[k]);
    items_.swap(items);
    bin_nodes_ = out;
    direct_node(0, 0, n, 0);
    bin_nodes_ = nullptr;
    items_.swap(items);
    return n - 1;
  }
  size_t build_items_direct(std::vector<Item>& items, PackedNode* out) {
    const size_t n = items.size();
    if (n == 0) return 0;
    for (Item& it : items)
      for (int k = 0; k < 3; ++k) it.c[k] = 0.5f * (it.box.lo[k] + it.box.hi[k]);
    if (n == 1) {
      PackedNode nd{};
      set_child(nd, 0, items[0].box, static_cast<int32_t>(~(items[0].ref | kDirectMark)));
      Box3 e; e.reset();
      set_child(nd, 1, e, static_cast<int32_t>(~kDirectMark));
      out[0] = nd;
      return 1;
    }
    items_.swap(items);
    ext_nodes_ = out;
    direct_node(0, 0, n, 0);
    ext_nodes_ = nullptr;
    items_.swap(items);
    return n - 1;
  }

  const std::vector<PackedNode>& nodes() const { return nodes_; }
  // inner-node levels of the direct (single-primitive-leaf) tree: the traversal stack of the kernels must hold this many entries
  int max_depth() const { return max_depth_.load(); }
  const std::vector<uint32_t>& leaf_refs() const { return leaf_refs_; }

 private:
  const std::vector<Box3>* boxes_ = nullptr;
  const std::vector<uint32_t>* refs_ = nullptr;
  std::vector<uint32_t> order_;
  std::vector<float> cent_;
  std::vector<PackedNode> nodes_;
  std::vector<uint32_t> leaf_refs_;


  // ---- fast path for single-primitive leaves (the default) -------------------------------------------------------------
  // A subtree over m primitives has exactly m-1 inner nodes, so node indices can be assigned in preorder up front
  // (root = base, left subtree = base+1 ..., right subtree after it): deterministic layout, children next to their
  // parent, and disjoint index ranges that independent threads can fill without synchronisation.
  void build_direct(const std::vector<Box3>& boxes, const std::vector<uint32_t>& refs) {
    const size_t n = boxes.size();
    items_.resize(n);
    for (size_t i = 0; i < n; ++i) {
      items_[i].box = boxes[i];
      for (int k = 0; k < 3; ++k) items_[i].c[k] = 0.5f * (boxes[i].lo[k] + boxes[i].hi[k]);
      items_[i].ref = refs[i];
    }
    nodes_.assign(n - 1, PackedNode{});
    leaf_refs_.clear();
    direct_node(0, 0, n, 0);
    items_.clear(); items_.shrink_to_fit();
  }

  int32_t direct_child_code(size_t node_index, size_t begin, size_t end) const {
    return end - begin == 1 ? static_cast<int32_t>(~(items_[begin].ref | kDirectMark)) : static_cast<int32_t>(node_index);
  }

  static float area_of(const Box3& b) { return b.half_area(); }

  void direct_node(size_t idx, size_t begin, size_t end, int depth) {
    note_depth(depth + 1);
    const size_t n = end - begin;
    Item* it = items_.data();
    size_t mid = begin;
    Box3 lb, rb;
    if (n == 2) {
      mid = begin + 1; lb = it[begin].box; rb = it[begin + 1].box;
    } else if (depth >= kSahDepthLimit) {
      // SAH on a pathological distribution (centroids spread over dozens of orders of magnitude) peels a few primitives per
      // level; below this depth the range is split at the object median instead, which bounds the tree depth by
      // kSahDepthLimit + log2(n) < kBvhStack (the kernels' traversal stack)
      float clo[3] = {1e30f, 1e30f, 1e30f}, chi[3] = {-1e30f, -1e30f, -1e30f};
      for (size_t i = begin; i < end; ++i)
        for (int k = 0; k < 3; ++k) { clo[k] = std::min(clo[k], it[i].c[k]); chi[k] = std::max(chi[k], it[i].c[k]); }
      int ax = 0;
      if (chi[1] - clo[1] > chi[ax] - clo[ax]) ax = 1;
      if (chi[2] - clo[2] > chi[ax] - clo[ax]) ax = 2;
      mid = begin + n / 2;
      std::nth_element(it + begin, it + mid, it + end, [ax](const Item& a, const Item& b) { return a.c[ax] < b.c[ax]; });
      lb.reset(); rb.reset();
      for (size_t i = begin; i < mid; ++i) lb.grow(it[i].box);
      for (size_t i = mid; i < end; ++i) rb.grow(it[i].box);
    } else if (n <= 8) {
      // exact SAH over the sorted order of the widest centroid axis
      float clo[3] = {1e30f, 1e30f, 1e30f}, chi[3] = {-1e30f, -1e30f, -1e30f};
      for (size_t i = begin; i < end; ++i)
        for (int k = 0; k < 3; ++k) { clo[k] = std::min(clo[k], it[i].c[k]); chi[k] = std::max(chi[k], it[i].c[k]); }
      int ax = 0;
      if (chi[1] - clo[1] > chi[ax] - clo[ax]) ax = 1;
      if (chi[2] - clo[2] > chi[ax] - clo[ax]) ax = 2;
      std::sort(it + begin, it + end, [ax](const Item& a, const Item& b) { return a.c[ax] < b.c[ax]; });
      Box3 suffix[8];
      Box3 acc; acc.reset();
      for (size_t k = n; k-- > 1;) { acc.grow(it[begin + k].box); suffix[k] = acc; }
      acc.reset();
      float best = std::numeric_limits<float>::infinity();
      size_t best_k = 1;
      for (size_t k = 1; k < n; ++k) {
        acc.grow(it[begin + k - 1].box);
        const float cost = area_of(acc) * static_cast<float>(k) + area_of(suffix[k]) * static_cast<float>(n - k);
        if (cost < best) { best = cost; best_k = k; lb = acc; rb = suffix[k]; }
      }
      mid = begin + best_k;
    } else {
      // binned SAH, all three axes in one pass over the range.  The few ranges at the top of a big tree (more than kParallelRange
      // primitives) are the serial critical path of the build, so their passes (centroid bounds, binning, partition) are split over
      // host threads; everything below runs as independent subtree tasks.
      // 32 bins per axis for big ranges, 16 below 256 primitives, 8 below 64: the per-node cost of resetting and sweeping the bins
      // dominated the build of a million-triangle mesh (a quarter of a million ranges of 9..64 primitives)
      const int nb = n >= 256 ? kBins : (n >= 64 ? 16 : 8);
      const int nt = n >= kParallelRange ? static_cast<int>(std::min<size_t>(std::max<size_t>(1, std::thread::hardware_concurrency()), 16)) : 1;
      float clo[3] = {1e30f, 1e30f, 1e30f}, chi[3] = {-1e30f, -1e30f, -1e30f};
      if (nt > 1) {
        std::vector<std::array<float, 6>> part(static_cast<size_t>(nt));
        for_chunks(begin, end, nt, [&](int t, size_t b, size_t e) {
          std::array<float, 6> a = {1e30f, 1e30f, 1e30f, -1e30f, -1e30f, -1e30f};
          for (size_t i = b; i < e; ++i)
            for (int k = 0; k < 3; ++k) { a[k] = std::min(a[k], it[i].c[k]); a[3 + k] = std::max(a[3 + k], it[i].c[k]); }
          part[static_cast<size_t>(t)] = a;
        });
        for (const auto& a : part)
          for (int k = 0; k < 3; ++k) { clo[k] = std::min(clo[k], a[k]); chi[k] = std::max(chi[k], a[3 + k]); }
      } else {
        for (size_t i = begin; i < end; ++i)
          for (int k = 0; k < 3; ++k) { clo[k] = std::min(clo[k], it[i].c[k]); chi[k] = std::max(chi[k], it[i].c[k]); }
      }
      Box3 bb[3][kBins]; uint32_t cnt[3][kBins];
      float scale[3];
      for (int ax = 0; ax < 3; ++ax) {
        const float ext = chi[ax] - clo[ax];
        scale[ax] = ext > 0.0f ? nb / ext : 0.0f;
        for (int b = 0; b < nb; ++b) { bb[ax][b].reset(); cnt[ax][b] = 0; }
      }
      if (n < single_axis_below) {   // small ranges: only the widest centroid axis is binned (a third of the work; see DESIGN.md for what it costs)
        int wide = 0;
        if (chi[1] - clo[1] > chi[wide] - clo[wide]) wide = 1;
        if (chi[2] - clo[2] > chi[wide] - clo[wide]) wide = 2;
        for (int ax = 0; ax < 3; ++ax) if (ax != wide) scale[ax] = 0.0f;
      }
      if (nt > 1) {
        struct Bins { Box3 bb[3][kBins]; uint32_t cnt[3][kBins]; };
        std::vector<Bins> part(static_cast<size_t>(nt));
        for_chunks(begin, end, nt, [&](int t, size_t b0, size_t e0) {
          Bins& B = part[static_cast<size_t>(t)];
          for (int ax = 0; ax < 3; ++ax) for (int b = 0; b < nb; ++b) { B.bb[ax][b].reset(); B.cnt[ax][b] = 0; }
          for (size_t i = b0; i < e0; ++i)
            for (int ax = 0; ax < 3; ++ax) {
              if (!(scale[ax] > 0.0f)) continue;
              int b = static_cast<int>((it[i].c[ax] - clo[ax]) * scale[ax]);
              b = std::min(std::max(b, 0), nb - 1);
              B.bb[ax][b].grow(it[i].box); ++B.cnt[ax][b];
            }
        });
        for (const Bins& B : part)
          for (int ax = 0; ax < 3; ++ax) for (int b = 0; b < nb; ++b) { bb[ax][b].grow(B.bb[ax][b]); cnt[ax][b] += B.cnt[ax][b]; }
      } else {
        for (size_t i = begin; i < end; ++i) {
          for (int ax = 0; ax < 3; ++ax) {
            if (!(scale[ax] > 0.0f)) continue;
            int b = static_cast<int>((it[i].c[ax] - clo[ax]) * scale[ax]);
            b = std::min(std::max(b, 0), nb - 1);
            bb[ax][b].grow(it[i].box); ++cnt[ax][b];
          }
        }
      }
      int best_axis = -1, best_split = -1; float best_cost = std::numeric_limits<float>::infinity();
      for (int ax = 0; ax < 3; ++ax) {
        if (!(scale[ax] > 0.0f)) continue;
        Box3 rbox[kBins]; uint32_t rcnt[kBins];
        Box3 acc; acc.reset(); uint32_t c = 0;
        for (int b = nb - 1; b > 0; --b) { acc.grow(bb[ax][b]); c += cnt[ax][b]; rbox[b] = acc; rcnt[b] = c; }
        acc.reset(); c = 0;
        for (int b = 0; b < nb - 1; ++b) {
          acc.grow(bb[ax][b]); c += cnt[ax][b];
          if (c == 0 || rcnt[b + 1] == 0) continue;
          const float cost = area_of(acc) * static_cast<float>(c) + area_of(rbox[b + 1]) * static_cast<float>(rcnt[b + 1]);
          if (cost < best_cost) { best_cost = cost; best_axis = ax; best_split = b; lb = acc; rb = rbox[b + 1]; }
        }
      }
      if (best_axis >= 0) {
        const int ax = best_axis, sp = best_split;
        const float lo = clo[ax], sc = scale[ax];
        auto goes_left = [=](const Item& a) {
          int b = static_cast<int>((a.c[ax] - lo) * sc);
          b = std::min(std::max(b, 0), nb - 1);
          return b <= sp;
        };
        if (nt > 1) {
          // parallel partition: count per chunk, then every chunk scatters its items to their final places in a scratch copy
          std::vector<size_t> nleft(static_cast<size_t>(nt), 0), nall(static_cast<size_t>(nt), 0);
          for_chunks(begin, end, nt, [&](int t, size_t b0, size_t e0) {
            size_t c = 0;
            for (size_t i = b0; i < e0; ++i) c += goes_left(it[i]) ? 1 : 0;
            nleft[static_cast<size_t>(t)] = c; nall[static_cast<size_t>(t)] = e0 - b0;
          });
          size_t total_left = 0;
          for (size_t c : nleft) total_left += c;
          std::vector<size_t> loff(static_cast<size_t>(nt)), roff(static_cast<size_t>(nt));
          size_t l = 0, r = total_left;
          for (int t = 0; t < nt; ++t) { loff[static_cast<size_t>(t)] = l; roff[static_cast<size_t>(t)] = r; l += nleft[static_cast<size_t>(t)]; r += nall[static_cast<size_t>(t)] - nleft[static_cast<size_t>(t)]; }
          std::vector<Item> scratch(n);
          for_chunks(begin, end, nt, [&](int t, size_t b0, size_t e0) {
            size_t lo_at = loff[static_cast<size_t>(t)], hi_at = roff[static_cast<size_t>(t)];
            for (size_t i = b0; i < e0; ++i) { if (goes_left(it[i])) scratch[lo_at++] = it[i]; else scratch[hi_at++] = it[i]; }
          });
          for_chunks(begin, end, nt, [&](int, size_t b0, size_t e0) { std::copy(scratch.begin() + (b0 - begin), scratch.begin() + (e0 - begin), it + b0); });
          mid = begin + total_left;
        } else {
          mid = static_cast<size_t>(std::partition(it + begin, it + end, goes_left) - it);
        }
      }
      if (best_axis < 0 || mid == begin || mid == end) {  // coincident centroids: split the list in half
        mid = begin + n / 2;
        lb.reset(); rb.reset();
        for (size_t i = begin; i < mid; ++i) lb.grow(it[i].box);
        for (size_t i = mid; i < end; ++i) rb.grow(it[i].box);
      }
    }
    const size_t nl = mid - begin;
    const size_t left_idx = idx + 1, right_idx = idx + 1 + (nl > 1 ? nl - 1 : 0);
    if (bin_nodes_) {
      BinNode bn;
      bn.box[0] = lb; bn.box[1] = rb;
      bn.child[0] = direct_child_code(left_idx, begin, mid); bn.child[1] = direct_child_code(right_idx, mid, end);
      bn.count[0] = static_cast<uint32_t>(nl); bn.count[1] = static_cast<uint32_t>(n - nl);
      bin_nodes_[idx] = bn;
    } else {
      PackedNode nd{};
      set_child(nd, 0, lb, direct_child_code(left_idx, begin, mid));
      set_child(nd, 1, rb, direct_child_code(right_idx, mid, end));
      (ext_nodes_ ? ext_nodes_ : nodes_.data())[idx] = nd;
    }
    // big subtrees go to other threads (disjoint item and node ranges)
    std::future<void> task;
    if (nl > 1) {
      if (n > kTaskRange && depth < 12) task = std::async(std::launch::async, [=] { direct_node(left_idx, begin, mid, depth + 1); });
      else direct_node(left_idx, begin, mid, depth + 1);
    }
    if (end - mid > 1) direct_node(right_idx, mid, end, depth + 1);
    if (task.valid()) task.get();
  }
  std::vector<Item> items_;
  PackedNode* ext_nodes_ = nullptr;
  BinNode* bin_nodes_ = nullptr;
  std::atomic<int> max_depth_{0};
  void note_depth(int d) { int cur = max_depth_.load(std::memory_order_relaxed); while (d > cur && !max_depth_.compare_exchange_weak(cur, d, std::memory_order_relaxed)) {} }

  // Child boxes are stored as centre c and half-extent e (node_slabs on the device: t = (c -+ e) / d - o / d costs FMA-pipe
  // operations instead of min/max ALU operations).  e is rounded up so that [c - e, c + e] contains the exact box plus four ulps of
  // its largest coordinate: conservative against the fp32 slab arithmetic.
  static void centre_extent(const Box3& b, float c[3], float e[3]) {
    for (int k = 0; k < 3; ++k) {
      if (!(b.lo[k] <= b.hi[k])) { c[k] = kEmptyChildCentre; e[k] = 0.0f; continue; }
      const double lo = b.lo[k], hi = b.hi[k];
      c[k] = static_cast<float>(0.5 * (lo + hi));
      const double m = std::max(std::fabs(lo), std::fabs(hi));
      const double need = std::max(hi - static_cast<double>(c[k]), static_cast<double>(c[k]) - lo) + 4.0 * 1.1920929e-7 * m + 1e-30;
      float ef = static_cast<float>(need);
      if (static_cast<double>(ef) < need) ef = std::nextafter(ef, std::numeric_limits<float>::infinity());
      e[k] = ef;
    }
  }
  static void set_child(PackedNode& nd, int which, const Box3& b, int32_t code) {
    float c[3], e[3];
    centre_extent(b, c, e);
    for (int k = 0; k < 3; ++k) { nd.c[k][which] = c[k]; nd.e[k][which] = e[k]; }
    (which == 0 ? nd.left : nd.right) = code;
  }
  // leaf child code: ~((first << 5) | count) into leaf_refs, or, with single-primitive leaves (kMaxLeaf == 1),
  // ~reference itself: no indirection left on the device
  int32_t make_leaf(size_t begin, size_t end) {
    if (kMaxLeaf == 1 && end - begin == 1) return static_cast<int32_t>(~((*refs_)[order_[begin]] | kDirectMark));
    const uint32_t first = static_cast<uint32_t>(leaf_refs_.size());
    for (size_t i = begin; i < end; ++i) leaf_refs_.push_back((*refs_)[order_[i]]);
    const uint32_t v = (first << 5) | static_cast<uint32_t>(end - begin);
    return static_cast<int32_t>(~v);
  }
  Box3 range_box(size_t begin, size_t end) const {
    Box3 b; b.reset();
    for (size_t i = begin; i < end; ++i) b.grow((*boxes_)[order_[i]]);
    return b;
  }

  // returns child code for [begin,end): either a leaf code or the index of a freshly built inner node
  int32_t build_child(size_t begin, size_t end, const Box3& box, int depth) {
    if (end - begin <= static_cast<size_t>(kMaxLeaf)) return make_leaf(begin, end);
    const int32_t idx = static_cast<int32_t>(nodes_.size());
    nodes_.push_back(PackedNode{});
    build_node(idx, begin, end, box, depth);
    return idx;
  }

  void build_node(int32_t idx, size_t begin, size_t end, const Box3& box, int depth) {
    (void)box;
    note_depth(depth + 1);
    // centroid bounds
    float clo[3], chi[3];
    for (int k = 0; k < 3; ++k) { clo[k] = std::numeric_limits<float>::infinity(); chi[k] = -clo[k]; }
    for (size_t i = begin; i < end; ++i)
      for (int k = 0; k < 3; ++k) { const float c = cent_[3 * order_[i] + k]; clo[k] = std::min(clo[k], c); chi[k] = std::max(chi[k], c); }
    int best_axis = -1, best_split = -1; float best_cost = std::numeric_limits<float>::infinity();
    for (int ax = 0; ax < 3; ++ax) {
      const float ext = chi[ax] - clo[ax];
      if (!(ext > 0.0f)) continue;
      Box3 bb[kBins]; uint32_t cnt[kBins] = {0};
      for (int b = 0; b < kBins; ++b) bb[b].reset();
      const float scale = kBins / ext;
      for (size_t i = begin; i < end; ++i) {
        int b = static_cast<int>((cent_[3 * order_[i] + ax] - clo[ax]) * scale);
        b = std::min(std::max(b, 0), kBins - 1);
        bb[b].grow((*boxes_)[order_[i]]); ++cnt[b];
      }
      float right_area[kBins]; uint32_t right_cnt[kBins];
      Box3 acc; acc.reset(); uint32_t c = 0;
      for (int b = kBins - 1; b > 0; --b) { acc.grow(bb[b]); c += cnt[b]; right_area[b] = acc.half_area(); right_cnt[b] = c; }
      acc.reset(); c = 0;
      for (int b = 0; b < kBins - 1; ++b) {
        acc.grow(bb[b]); c += cnt[b];
        if (c == 0 || right_cnt[b + 1] == 0) continue;
        const float cost = acc.half_area() * c + right_area[b + 1] * right_cnt[b + 1];
        if (cost < best_cost) { best_cost = cost; best_axis = ax; best_split = b; }
      }
    }
    size_t mid;
    if (best_axis < 0 || depth >= kSahDepthLimit) {
      // all centroids coincide, or SAH keeps peeling a few primitives per level on a pathological distribution: split the list
      // in half, which bounds the depth by kSahDepthLimit + log2(n) < kBvhStack like the single-primitive-leaf builder does
      mid = begin + (end - begin) / 2;
      if (best_axis >= 0) {
        int ax = 0;
        if (chi[1] - clo[1] > chi[ax] - clo[ax]) ax = 1;
        if (chi[2] - clo[2] > chi[ax] - clo[ax]) ax = 2;
        std::nth_element(order_.begin() + begin, order_.begin() + mid, order_.begin() + end,
                         [&](uint32_t a, uint32_t b) { return cent_[3 * a + ax] < cent_[3 * b + ax]; });
      }
    } else {
      const float ext = chi[best_axis] - clo[best_axis];
      const float scale = kBins / ext;
      const float lo = clo[best_axis];
      const int ax = best_axis, sp = best_split;
      auto it = std::partition(order_.begin() + begin, order_.begin() + end, [&](uint32_t id) {
        int b = static_cast<int>((cent_[3 * id + ax] - lo) * scale);
        b = std::min(std::max(b, 0), kBins - 1);
        return b <= sp;
      });
      mid = static_cast<size_t>(it - order_.begin());
      if (mid == begin || mid == end) mid = begin + (end - begin) / 2;
    }
    const Box3 lb = range_box(begin, mid), rb = range_box(mid, end);
    const int32_t lc = build_child(begin, mid, lb, depth + 1);
    const int32_t rc = build_child(mid, end, rb, depth + 1);
    PackedNode nd{};
    set_child(nd, 0, lb, lc);
    set_child(nd, 1, rb, rc);
    nodes_[idx] = nd;
  }
};


// ---------------------------------------------------------------------------------------------------------------------------------
// Compressed wide BVH (after Ylitie, Karras, Laine: "Efficient Incoherent Ray Traversal on GPUs Through Compressed Wide BVHs", HPG 2017):
// 8-wide nodes of 80 bytes holding the children's boxes quantised to 8 bits per plane on a per-node power-of-two grid, children
// placed in slots by octant so that traversal order is a bit trick instead of a sort.  Built by collapsing the binary SAH tree above.
// Used for every scene with triangles (the mesh path): a 991k-triangle mesh needs 63 MB of 64-byte binary nodes, 5.6 MB of these,
// and a ray makes a third of the dependent node fetches.
//
// Node layout (5 x 16 bytes; word = little-endian uint32):
//   w0..w2  p.x p.y p.z   origin of the quantisation grid (float bits)
//   w3      e.x | e.y << 8 | e.z << 16 | imask << 24      biased exponents (grid step 2^(e-127) per axis), bit i of imask: slot i is an inner node
//   w4      child_base    index of the first inner child (inner children are contiguous, in slot order)
//   w5      prim_base     index of the first leaf primitive of this node (leaf primitives are contiguous, in slot order)
//   w6,w7   meta[8]       per slot: 0 = empty; inner: 0x20 | (24 + slot); leaf: (unary count 1/3/7) << 5 | offset of its first primitive from prim_base
//   w8,w9   qlo.x[8]   w10,w11 qlo.y[8]   w12,w13 qlo.z[8]   w14,w15 qhi.x[8]   w16,w17 qhi.y[8]   w18,w19 qhi.z[8]
// Slot s sits on the + side of axis k iff bit k of s is set.  Child boxes are dequantised as p + q * 2^(e-127) and contain the exact
// boxes widened by `margin` (fp32 slack of the traversal arithmetic) plus an eighth of a grid step.
// ---------------------------------------------------------------------------------------------------------------------------------
struct CwNode { uint32_t w[20]; };
static_assert(sizeof(CwNode) == 80, "wide node must be 80 bytes");

class CwBuilder {
 public:
  int max_leaf = 3;       // primitives per leaf child (1..3), same kind only
  double margin = 0.0;    // absolute widening of every child box (4 ulp of the scene's largest coordinate)

  // bin: n_items - 1 binary nodes (n_items >= 2), or none with single_ref set (n_items == 1).  Outputs the nodes in breadth-first order
  // and the primitive references in leaf order.
  void build(const BinNode* bin, size_t n_bin, const Box3* single_box, uint32_t single_ref) {
    nodes_.clear(); leaf_order_.clear(); depth_ = 0;
    bin_ = bin;
    if (n_bin == 0) {
      if (!single_box) return;
      Plan pl; pl.n = 1; pl.bin[0] = -1; pl.box[0] = *single_box; pl.nprim[0] = 1; pl.prims[0][0] = single_ref;
      for (int s = 0; s < 8; ++s) pl.slot_child[s] = s == 0 ? 0 : -1;
      CwNode nd; encode(pl, 0, 0, nd);
      nodes_.push_back(nd); leaf_order_.push_back(single_ref); depth_ = 1;
      return;
    }
    std::vector<int32_t> level{0};          // binary roots of the wide nodes of the current level; wide index = level_base + position
    size_t level_base = 0, prim_cursor = 0;
    while (!level.empty()) {
      ++depth_;
      const size_t m = level.size();
      std::vector<Plan> plans(m);
      parallel_for(m, [&](size_t i) { plan_node(level[i], plans[i]); });
      // serial prefix sums: children of one node are contiguous (next level, in node then slot order), and so are its leaf primitives
      std::vector<uint32_t> child_base(m), prim_base(m);
      size_t next_count = 0;
      for (size_t i = 0; i < m; ++i) {
        child_base[i] = static_cast<uint32_t>(level_base + m + next_count);
        prim_base[i] = static_cast<uint32_t>(prim_cursor);
        for (int s = 0; s < 8; ++s) {
          if (plans[i].slot_child[s] < 0) continue;
          const int c = plans[i].slot_child[s];
          if (plans[i].bin[c] >= 0) ++next_count; else prim_cursor += static_cast<size_t>(plans[i].nprim[c]);
        }
      }
      nodes_.resize(level_base + m);
      leaf_order_.resize(prim_cursor);
      std::vector<int32_t> next(next_count);
      parallel_for(m, [&](size_t i) {
        encode(plans[i], child_base[i], prim_base[i], nodes_[level_base + i]);
        size_t at = child_base[i] - (level_base + m), pp = prim_base[i];
        for (int s = 0; s < 8; ++s) {
          const int c = plans[i].slot_child[s];
          if (c < 0) continue;
          if (plans[i].bin[c] >= 0) next[at++] = plans[i].bin[c];
          else for (int k = 0; k < plans[i].nprim[c]; ++k) leaf_order_[pp++] = plans[i].prims[c][k];
        }
      });
      level_base += m;
      level.swap(next);
    }
  }
  const std::vector<CwNode>& nodes() const { return nodes_; }
  const std::vector<uint32_t>& leaf_order() const { return leaf_order_; }   // (kind << 30) | table index, in leaf order
  int depth() const { return depth_; }

 private:
  struct Plan {
    int n = 0;
    int32_t bin[8];          // >= 0: binary node that roots the inner child; -1: leaf child
    Box3 box[8];
    int nprim[8];
    uint32_t prims[8][3];
    int slot_child[8];       // slot -> child index or -1
  };
  const BinNode* bin_ = nullptr;
  std::vector<CwNode> nodes_;
  std::vector<uint32_t> leaf_order_;
  int depth_ = 0;

  template <typename F>
  static void parallel_for(size_t n, F&& fn) {
    const size_t hw = std::max<size_t>(1, std::thread::hardware_concurrency());
    const size_t nt = n >= 4096 ? std::min<size_t>(hw, 32) : 1;
    if (nt <= 1) { for (size_t i = 0; i < n; ++i) fn(i); return; }
    std::vector<std::future<void>> tasks;
    for (size_t t = 1; t < nt; ++t) tasks.push_back(std::async(std::launch::async, [&fn, t, n, nt] { for (size_t i = n * t / nt; i < n * (t + 1) / nt; ++i) fn(i); }));
    for (size_t i = 0; i < n / nt; ++i) fn(i);
    for (auto& t : tasks) t.get();
  }

  static uint32_t ref_of(int32_t code) { return static_cast<uint32_t>(~code) & ~BvhBuilder::kDirectMark; }
  // primitives below a binary child, if it can be ONE leaf (<= max_leaf primitives of one kind)
  bool leafable(int32_t code, uint32_t count, uint32_t* out, int* n) const {
    if (count > static_cast<uint32_t>(max_leaf)) return false;
    *n = 0;
    gather(code, out, n);
    for (int k = 1; k < *n; ++k)
      if ((out[k] >> 30) != (out[0] >> 30)) return false;
    return true;
  }
  void gather(int32_t code, uint32_t* out, int* n) const {
    if (code < 0) { out[(*n)++] = ref_of(code); return; }
    gather(bin_[code].child[0], out, n);
    gather(bin_[code].child[1], out, n);
  }

  void plan_node(int32_t root, Plan& pl) const {
    struct Open { int32_t code; uint32_t count; Box3 box; bool leaf; int np; uint32_t prims[3]; };
    Open ch[8];
    int n = 0;
    auto add = [&](int32_t code, uint32_t count, const Box3& box) {
      Open& o = ch[n++];
      o.code = code; o.count = count; o.box = box; o.np = 0;
      o.leaf = leafable(code, count, o.prims, &o.np);
    };
    add(bin_[root].child[0], bin_[root].count[0], bin_[root].box[0]);
    add(bin_[root].child[1], bin_[root].count[1], bin_[root].box[1]);
    while (n < 8) {   // open the inner child with the largest surface area
      int best = -1; float best_area = -1.0f;
      for (int i = 0; i < n; ++i)
        if (!ch[i].leaf && ch[i].box.half_area() > best_area) { best_area = ch[i].box.half_area(); best = i; }
      if (best < 0) break;
      const BinNode& b = bin_[ch[best].code];
      ch[best] = ch[n - 1]; --n;
      add(b.child[0], b.count[0], b.box[0]);
      add(b.child[1], b.count[1], b.box[1]);
    }
    pl.n = n;
    Box3 nb; nb.reset();
    for (int i = 0; i < n; ++i) {
      pl.bin[i] = ch[i].leaf ? -1 : ch[i].code;
      pl.box[i] = ch[i].box;
      pl.nprim[i] = ch[i].leaf ? ch[i].np : 0;
      for (int k = 0; k < 3; ++k) pl.prims[i][k] = ch[i].prims[k];
      nb.grow(ch[i].box);
    }
    // children -> slots: greedily the (child, slot) pair whose centroid offset points most along the slot's octant direction
    float cost[8][8];
    for (int i = 0; i < n; ++i)
      for (int s = 0; s < 8; ++s) {
        float c = 0.0f;
        for (int k = 0; k < 3; ++k) {
          const float off = 0.5f * (pl.box[i].lo[k] + pl.box[i].hi[k]) - 0.5f * (nb.lo[k] + nb.hi[k]);
          c += (s >> k & 1) ? off : -off;
        }
        cost[i][s] = c;
      }
    bool child_done[8] = {false}, slot_done[8] = {false};
    for (int s = 0; s < 8; ++s) pl.slot_child[s] = -1;
    for (int round = 0; round < n; ++round) {
      int bi = -1, bs = -1; float bc = -std::numeric_limits<float>::infinity();
      for (int i = 0; i < n; ++i) {
        if (child_done[i]) continue;
        for (int s = 0; s < 8; ++s)
          if (!slot_done[s] && cost[i][s] > bc) { bc = cost[i][s]; bi = i; bs = s; }
      }
      child_done[bi] = true; slot_done[bs] = true; pl.slot_child[bs] = bi;
    }
  }

  void encode(const Plan& pl, uint32_t child_base, uint32_t prim_base, CwNode& nd) const {
    std::memset(&nd, 0, sizeof nd);
    Box3 nb; nb.reset();
    for (int i = 0; i < pl.n; ++i) nb.grow(pl.box[i]);
    float p[3]; int eb[3]; double step[3];
    for (int k = 0; k < 3; ++k) {
      const double lo = static_cast<double>(nb.lo[k]) - margin, hi = static_cast<double>(nb.hi[k]) + margin;
      // step = smallest power of two with 249 steps covering the extent (the rest of the 255 is slack for the outward rounding), not
      // below 8 * margin so that an eighth of a step still covers the traversal's fp32 slack; grid origin a quarter step below lo
      const double need = std::max((hi - lo) / 249.0, std::max(8.0 * margin, 1e-30));
      int e = static_cast<int>(std::ceil(std::log2(need)));
      if (std::ldexp(1.0, e) < need) ++e;
      e = std::min(std::max(e, -120), 120);
      const double origin = lo - 0.25 * std::ldexp(1.0, e);
      float pf = static_cast<float>(origin);
      if (static_cast<double>(pf) > origin) pf = std::nextafter(pf, -std::numeric_limits<float>::infinity());
      p[k] = pf; eb[k] = e + 127; step[k] = std::ldexp(1.0, e);
    }
    uint32_t imask = 0;
    uint8_t meta[8] = {0}, q[6][8];
    for (int s = 0; s < 8; ++s) { for (int a = 0; a < 3; ++a) { q[a][s] = 255; q[3 + a][s] = 0; } }
    uint32_t inner_rank = 0, prim_off = 0;
    for (int s = 0; s < 8; ++s) {
      const int c = pl.slot_child[s];
      if (c < 0) continue;
      if (pl.bin[c] >= 0) { imask |= 1u << s; meta[s] = static_cast<uint8_t>(0x20 | (24 + s)); ++inner_rank; }
      else {
        const uint32_t unary = pl.nprim[c] == 1 ? 1u : (pl.nprim[c] == 2 ? 3u : 7u);
        meta[s] = static_cast<uint8_t>((unary << 5) | prim_off);
        prim_off += static_cast<uint32_t>(pl.nprim[c]);
      }
      for (int k = 0; k < 3; ++k) {
        const double lo = (static_cast<double>(pl.box[c].lo[k]) - margin - static_cast<double>(p[k])) / step[k] - 0.125;
        const double hi = (static_cast<double>(pl.box[c].hi[k]) + margin - static_cast<double>(p[k])) / step[k] + 0.125;
        q[k][s] = static_cast<uint8_t>(std::min(std::max(std::floor(lo), 0.0), 255.0));
        q[3 + k][s] = static_cast<uint8_t>(std::min(std::max(std::ceil(hi), 0.0), 255.0));
      }
    }
    (void)inner_rank;
    std::memcpy(&nd.w[0], &p[0], 4); std::memcpy(&nd.w[1], &p[1], 4); std::memcpy(&nd.w[2], &p[2], 4);
    nd.w[3] = static_cast<uint32_t>(eb[0]) | static_cast<uint32_t>(eb[1]) << 8 | static_cast<uint32_t>(eb[2]) << 16 | imask << 24;
    nd.w[4] = child_base; nd.w[5] = prim_base;
    std::memcpy(&nd.w[6], meta, 8);
    for (int a = 0; a < 6; ++a) std::memcpy(&nd.w[8 + 2 * a], q[a], 8);
  }
};

}  // namespace rtw
