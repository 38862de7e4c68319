// bvh_build.cpp - deterministic binned-SAH builder (host) producing the paired-box binary BVH the
// GPU traversal kernel reads with 16-byte loads.  See bvh.h.  (The reference's own build,
// Accelerator/BVHAccel.cpp:97-258, is a recursive 12-bucket SAH over shared_ptr<Primitive>; this one
// works on index arrays, builds leaves of up to `maxLeaf` triangles and is free to differ: only the
// answers of ray queries must match.)
#include "bvh.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>

namespace pbr {
namespace {

struct TmpNode {
  Aabb box;
  int32_t left = -1, right = -1;  // children (TmpNode indices) or -1
  int32_t first = 0, count = 0;   // leaf range in the order array
};

inline Aabb empty_box() {
  Aabb b;
  for (int k = 0; k < 3; ++k) {
    b.lo[k] = std::numeric_limits<float>::infinity();
    b.hi[k] = -std::numeric_limits<float>::infinity();
  }
  return b;
}
inline void grow(Aabb& b, const Aabb& o) {
  for (int k = 0; k < 3; ++k) {
    b.lo[k] = std::min(b.lo[k], o.lo[k]);
    b.hi[k] = std::max(b.hi[k], o.hi[k]);
  }
}
inline float half_area(const Aabb& b) {
  float dx = b.hi[0] - b.lo[0], dy = b.hi[1] - b.lo[1], dz = b.hi[2] - b.lo[2];
  if (dx < 0 || dy < 0 || dz < 0) return 0.f;
  return dx * dy + dy * dz + dz * dx;
}

constexpr int kBins = 16;
constexpr float kTravCost = 1.0f;
constexpr float kIsectCost = 1.2f;

}  // namespace

void build_bvh2(const float* verts, const int32_t* indices, int64_t nTris, int maxLeaf, BuiltBvh* out) {
  maxLeaf = std::max(1, std::min(4, maxLeaf));
  std::vector<Aabb> tb((size_t)nTris);
  std::vector<float> cen((size_t)nTris * 3);
  for (int64_t i = 0; i < nTris; ++i) {
    Aabb b = empty_box();
    for (int v = 0; v < 3; ++v) {
      const float* p = verts + 3 * (int64_t)indices[3 * i + v];
      for (int k = 0; k < 3; ++k) {
        b.lo[k] = std::min(b.lo[k], p[k]);
        b.hi[k] = std::max(b.hi[k], p[k]);
      }
    }
    tb[i] = b;
    for (int k = 0; k < 3; ++k) cen[3 * i + k] = 0.5f * b.lo[k] + 0.5f * b.hi[k];
  }
  std::vector<int32_t>& order = out->triOrder;
  order.resize((size_t)nTris);
  for (int64_t i = 0; i < nTris; ++i) order[i] = (int32_t)i;

  std::vector<TmpNode> tmp;
  tmp.reserve((size_t)std::max<int64_t>(2 * nTris, 2));
  struct Task { int32_t node, first, count, depth; };
  std::vector<Task> stack;
  tmp.emplace_back();
  stack.push_back({0, 0, (int32_t)nTris, 0});
  out->maxLeafSize = 0;
  out->nLeaves = 0;
  out->maxDepth = 0;
  while (!stack.empty()) {
    Task t = stack.back();
    stack.pop_back();
    Aabb box = empty_box(), cb = empty_box();
    for (int32_t i = t.first; i < t.first + t.count; ++i) {
      int32_t id = order[i];
      grow(box, tb[id]);
      for (int k = 0; k < 3; ++k) {
        cb.lo[k] = std::min(cb.lo[k], cen[3 * id + k]);
        cb.hi[k] = std::max(cb.hi[k], cen[3 * id + k]);
      }
    }
    tmp[t.node].box = box;
    out->maxDepth = std::max(out->maxDepth, (int)t.depth);
    auto make_leaf = [&]() {
      tmp[t.node].first = t.first;
      tmp[t.node].count = t.count;
      out->maxLeafSize = std::max(out->maxLeafSize, (int)t.count);
      out->nLeaves++;
    };
    if (t.count <= 1) { make_leaf(); continue; }
    // pick split: binned SAH over the widest centroid axis first, all three axes evaluated
    int bestAxis = -1, bestBin = -1;
    float bestCost = std::numeric_limits<float>::infinity();
    float parentArea = half_area(box);
    for (int axis = 0; axis < 3; ++axis) {
      float lo = cb.lo[axis], hi = cb.hi[axis];
      if (!(hi > lo)) continue;
      float scale = kBins / (hi - lo);
      Aabb bb[kBins];
      int bc[kBins];
      for (int b = 0; b < kBins; ++b) { bb[b] = empty_box(); bc[b] = 0; }
      for (int32_t i = t.first; i < t.first + t.count; ++i) {
        int32_t id = order[i];
        int b = (int)((cen[3 * id + axis] - lo) * scale);
        b = std::max(0, std::min(kBins - 1, b));
        bc[b]++;
        grow(bb[b], tb[id]);
      }
      float rightArea[kBins];
      int rightCount[kBins];
      Aabb acc = empty_box();
      int cnt = 0;
      for (int b = kBins - 1; b > 0; --b) {
        grow(acc, bb[b]);
        cnt += bc[b];
        rightArea[b] = half_area(acc);
        rightCount[b] = cnt;
      }
      acc = empty_box();
      cnt = 0;
      for (int b = 0; b < kBins - 1; ++b) {
        grow(acc, bb[b]);
        cnt += bc[b];
        if (cnt == 0 || rightCount[b + 1] == 0) continue;
        float cost = kTravCost + kIsectCost * (cnt * half_area(acc) + rightCount[b + 1] * rightArea[b + 1]) /
                                     std::max(parentArea, 1e-30f);
        if (cost < bestCost) { bestCost = cost; bestAxis = axis; bestBin = b; }
      }
    }
    float leafCost = kIsectCost * t.count;
    if (t.count <= maxLeaf && (bestAxis < 0 || leafCost <= bestCost)) { make_leaf(); continue; }
    int32_t mid;
    if (bestAxis >= 0) {
      float lo = cb.lo[bestAxis], hi = cb.hi[bestAxis];
      float scale = kBins / (hi - lo);
      int32_t* b0 = order.data() + t.first;
      int32_t* m = std::stable_partition(b0, b0 + t.count, [&](int32_t id) {
        int b = (int)((cen[3 * id + bestAxis] - lo) * scale);
        b = std::max(0, std::min(kBins - 1, b));
        return b <= bestBin;
      });
      mid = (int32_t)(m - order.data());
    } else {
      mid = t.first + t.count / 2;  // all centroids coincide: split by index
    }
    if (mid == t.first || mid == t.first + t.count) mid = t.first + t.count / 2;
    int32_t l = (int32_t)tmp.size();
    tmp.emplace_back();
    tmp.emplace_back();
    tmp[t.node].left = l;
    tmp[t.node].right = l + 1;
    stack.push_back({l + 1, mid, t.first + t.count - mid, t.depth + 1});
    stack.push_back({l, t.first, mid - t.first, t.depth + 1});
  }
  out->bounds = tmp[0].box;

  // flatten: interior TmpNodes -> Bvh2Node (DFS order, children of a node adjacent when both interior)
  auto encode_leaf = [](const TmpNode& n) { return ~(int32_t)((n.first << 2) | (n.count - 1)); };
  std::vector<int32_t> remap(tmp.size(), -1);
  std::vector<Bvh2Node>& nodes = out->nodes;
  nodes.clear();
  auto put_child = [&](Bvh2Node& dst, int which, const Aabb& b) {
    if (which == 0) {
      dst.a[0] = b.lo[0]; dst.a[1] = b.hi[0]; dst.a[2] = b.lo[1]; dst.a[3] = b.hi[1];
      dst.c[0] = b.lo[2]; dst.c[1] = b.hi[2];
    } else {
      dst.b[0] = b.lo[0]; dst.b[1] = b.hi[0]; dst.b[2] = b.lo[1]; dst.b[3] = b.hi[1];
      dst.c[2] = b.lo[2]; dst.c[3] = b.hi[2];
    }
  };
  if (tmp[0].left < 0) {
    // single-leaf scene: synthesize a root whose second child is an empty box
    Bvh2Node r;
    std::memset(&r, 0, sizeof(r));
    put_child(r, 0, tmp[0].box);
    put_child(r, 1, empty_box());
    r.child[0] = tmp[0].count > 0 ? encode_leaf(tmp[0]) : ~0;
    r.child[1] = r.child[0];
    if (tmp[0].count == 0) { put_child(r, 0, empty_box()); }
    nodes.push_back(r);
    return;
  }
  std::vector<int32_t> dfs;
  dfs.push_back(0);
  // first pass: assign indices in DFS order
  std::vector<int32_t> orderIdx;
  while (!dfs.empty()) {
    int32_t n = dfs.back();
    dfs.pop_back();
    remap[n] = (int32_t)orderIdx.size();
    orderIdx.push_back(n);
    const TmpNode& tn = tmp[n];
    if (tmp[tn.right].left >= 0) dfs.push_back(tn.right);
    if (tmp[tn.left].left >= 0) dfs.push_back(tn.left);
  }
  nodes.resize(orderIdx.size());
  for (size_t i = 0; i < orderIdx.size(); ++i) {
    const TmpNode& tn = tmp[orderIdx[i]];
    Bvh2Node r;
    std::memset(&r, 0, sizeof(r));
    const TmpNode& L = tmp[tn.left];
    const TmpNode& R = tmp[tn.right];
    put_child(r, 0, L.box);
    put_child(r, 1, R.box);
    r.child[0] = L.left >= 0 ? remap[tn.left] : encode_leaf(L);
    r.child[1] = R.left >= 0 ? remap[tn.right] : encode_leaf(R);
    nodes[i] = r;
  }
}

}  // namespace pbr
