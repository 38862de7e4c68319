// bvh_build.cpp - deterministic host builder of the compressed 8-wide BVH the GPU traversal kernel reads with
// 16-byte loads (see bvh.h): binned-SAH binary tree -> greedy collapse to 8 children -> quantisation.
// (The reference's own build, Accelerator/BVHAccel.cpp:97-258, is a recursive 12-bucket SAH over
// shared_ptr<Primitive>; this one works on index arrays and is free to differ: only the answers of ray
// queries must match.)
#include "bvh.h"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdlib>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <limits>
#include <thread>
#include <vector>

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

constexpr int kBins = 32;
constexpr float kTravCost = 1.0f;
constexpr float kIsectCost = 1.2f;

}  // namespace

namespace {
struct BinaryStats { int maxLeafSize = 0; int64_t nLeaves = 0; int maxDepth = 0; };
struct BuildTask { int32_t node, first, count, depth; };

// Binned-SAH top-down build of the subtree `root` into `tmp` (root.node must exist).  A node's split depends only on its
// primitive range, so subtrees are independent: with `deferred` given, every subtree of at most `cutoff` primitives is NOT
// descended into but recorded there (its node exists, empty) for another thread to build.  All three axes are binned in one
// pass over the range (same per-bin boxes and counts as three passes).
void build_range(const std::vector<Aabb>& tb, const std::vector<float>& cen, int maxLeaf, std::vector<TmpNode>& tmp, int32_t* order,
                 BuildTask root, int64_t cutoff, std::vector<BuildTask>* deferred, BinaryStats* out) {
  std::vector<BuildTask> stack;
  stack.push_back(root);
  while (!stack.empty()) {
    BuildTask t = stack.back();
    stack.pop_back();
    if (deferred && t.count <= cutoff && t.node != root.node) { deferred->push_back(t); continue; }
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
    tmp[t.node].first = t.first;   // every node records its (contiguous) triangle range
    tmp[t.node].count = t.count;
    out->maxDepth = std::max(out->maxDepth, (int)t.depth);
    auto make_leaf = [&]() {
      tmp[t.node].first = t.first;
      tmp[t.node].count = t.count;
      out->maxLeafSize = std::max(out->maxLeafSize, (int)t.count);
      out->nLeaves++;
    };
    if (t.count <= 1) { make_leaf(); continue; }
    // pick split: binned SAH, all three axes evaluated
    int bestAxis = -1, bestBin = -1;
    float bestCost = std::numeric_limits<float>::infinity();
    float parentArea = half_area(box);
    Aabb bb[3][kBins];
    int bc[3][kBins];
    float scale[3];
    bool valid[3];
    for (int axis = 0; axis < 3; ++axis) {
      valid[axis] = cb.hi[axis] > cb.lo[axis];
      scale[axis] = valid[axis] ? kBins / (cb.hi[axis] - cb.lo[axis]) : 0.f;
      for (int b = 0; b < kBins; ++b) { bb[axis][b] = empty_box(); bc[axis][b] = 0; }
    }
    for (int32_t i = t.first; i < t.first + t.count; ++i) {
      int32_t id = order[i];
      for (int axis = 0; axis < 3; ++axis) {
        if (!valid[axis]) continue;
        int b = (int)((cen[3 * id + axis] - cb.lo[axis]) * scale[axis]);
        b = std::max(0, std::min(kBins - 1, b));
        bc[axis][b]++;
        grow(bb[axis][b], tb[id]);
      }
    }
    for (int axis = 0; axis < 3; ++axis) {
      if (!valid[axis]) continue;
      float rightArea[kBins];
      int rightCount[kBins];
      Aabb acc = empty_box();
      int cnt = 0;
      for (int b = kBins - 1; b > 0; --b) {
        grow(acc, bb[axis][b]);
        cnt += bc[axis][b];
        rightArea[b] = half_area(acc);
        rightCount[b] = cnt;
      }
      acc = empty_box();
      cnt = 0;
      for (int b = 0; b < kBins - 1; ++b) {
        grow(acc, bb[axis][b]);
        cnt += bc[axis][b];
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
      float lo = cb.lo[bestAxis];
      float sc = scale[bestAxis];
      int32_t* b0 = order + t.first;
      int32_t* m = std::stable_partition(b0, b0 + t.count, [&](int32_t id) {
        int b = (int)((cen[3 * id + bestAxis] - lo) * sc);
        b = std::max(0, std::min(kBins - 1, b));
        return b <= bestBin;
      });
      mid = (int32_t)(m - order);
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
}

// Binned-SAH binary tree over triangle bounds; `order` = leaf-order permutation of triangle ids.  The top of the tree is built
// by the calling thread down to subtrees of <= nTris / (8 * threads) primitives; those are built by a pool of threads into
// private node arrays and appended in a fixed order, so the tree (and everything derived from it) does not depend on the
// number of threads or on their timing.  Children always get larger indices than their parent (compute_widening relies on it).
void build_binary(const float* verts, const int32_t* indices, int64_t nTriangles, const Aabb* extra, int64_t nExtra, int maxLeaf,
                  std::vector<TmpNode>& tmp, std::vector<int32_t>& order, BinaryStats* out) {
  const int64_t nTris = nTriangles + nExtra;  // primitives: triangles, then the caller's extra boxes
  std::vector<Aabb> tb((size_t)nTris);
  std::vector<float> cen((size_t)nTris * 3);
  for (int64_t i = 0; i < nTris; ++i) {
    Aabb b = empty_box();
    if (i < nTriangles) {
      for (int v = 0; v < 3; ++v) {
        const float* p = verts + 3 * (int64_t)indices[3 * i + v];
        for (int k = 0; k < 3; ++k) {
          b.lo[k] = std::min(b.lo[k], p[k]);
          b.hi[k] = std::max(b.hi[k], p[k]);
        }
      }
    } else {
      b = extra[i - nTriangles];
    }
    tb[i] = b;
    for (int k = 0; k < 3; ++k) cen[3 * i + k] = 0.5f * b.lo[k] + 0.5f * b.hi[k];
  }
  order.resize((size_t)nTris);
  for (int64_t i = 0; i < nTris; ++i) order[i] = (int32_t)i;

  int nThreads = (int)std::min(16u, std::max(1u, std::thread::hardware_concurrency()));
  if (const char* e = std::getenv("PBRGPU_BUILD_THREADS")) nThreads = std::max(1, std::min(64, std::atoi(e)));
  if (nTris < (1 << 15)) nThreads = 1;

  tmp.clear();
  tmp.reserve((size_t)std::max<int64_t>(2 * nTris, 2));
  tmp.emplace_back();
  *out = BinaryStats{};
  std::vector<BuildTask> deferred;
  build_range(tb, cen, maxLeaf, tmp, order.data(), {0, 0, (int32_t)nTris, 0}, nThreads > 1 ? nTris / (8 * (int64_t)nThreads) : 0,
              nThreads > 1 ? &deferred : nullptr, out);
  if (deferred.empty()) return;

  std::vector<std::vector<TmpNode>> local(deferred.size());
  std::vector<BinaryStats> stats(deferred.size());
  std::vector<size_t> bySize(deferred.size());
  for (size_t k = 0; k < bySize.size(); ++k) bySize[k] = k;
  std::stable_sort(bySize.begin(), bySize.end(), [&](size_t a, size_t b) { return deferred[a].count > deferred[b].count; });  // big subtrees first
  std::atomic<size_t> next{0};
  auto worker = [&]() {
    for (size_t q = next.fetch_add(1); q < bySize.size(); q = next.fetch_add(1)) {
      const size_t k = bySize[q];
      const BuildTask& d = deferred[k];
      local[k].reserve((size_t)2 * d.count);
      local[k].emplace_back();
      build_range(tb, cen, maxLeaf, local[k], order.data(), {0, d.first, d.count, d.depth}, 0, nullptr, &stats[k]);
    }
  };
  std::vector<std::thread> pool;
  for (int th = 1; th < nThreads; ++th) pool.emplace_back(worker);
  worker();
  for (std::thread& th : pool) th.join();
  // splice: subtree k's root is the node the top build left empty, its other nodes follow in the order of `deferred`
  for (size_t k = 0; k < deferred.size(); ++k) {
    const int32_t off = (int32_t)tmp.size() - 1;  // local id j >= 1 -> tmp id off + j
    std::vector<TmpNode>& loc = local[k];
    for (TmpNode& nd : loc)
      if (nd.left >= 0) { nd.left += off; nd.right += off; }
    tmp[deferred[k].node] = loc[0];
    tmp.insert(tmp.end(), loc.begin() + 1, loc.end());
    out->maxLeafSize = std::max(out->maxLeafSize, stats[k].maxLeafSize);
    out->nLeaves += stats[k].nLeaves;
    out->maxDepth = std::max(out->maxDepth, stats[k].maxDepth);
    std::vector<TmpNode>().swap(loc);
  }
}
}  // namespace

}  // namespace pbr

// ---------------------------------------------------------------------------------------------
// 8-wide compressed BVH: binary SAH tree (leaves <= 3 triangles) -> greedy collapse (open the
// child with the largest surface area until 8 children) -> octant-ordered slot assignment ->
// 8-bit quantised child boxes.  Triangles are re-ordered so that the leaf children of one node are
// contiguous from triBase (offsets < 24).
// ---------------------------------------------------------------------------------------------
namespace pbr {

// SAH-optimal widening of the binary tree (Ylitie et al. 2017, section 4.1): for every binary node n,
//   C(n,1) = min( C_leaf(n), C_distribute(n,8) + A_n c_node ),   C_leaf(n) = A_n P_n c_prim  if P_n <= maxLeaf
//   C(n,i) = min( C_distribute(n,i), C(n,i-1) )                  C_distribute(n,j) = min_k C(left,k) + C(right,j-k)
// i.e. the cheapest way to represent the subtree as a forest of at most i roots; the wide node of n takes the forest
// that C_distribute(n,8) found as its children, and a subtree of <= maxLeaf triangles may become one leaf child.
namespace {
constexpr float kNodeCost = 1.0f;
float kPrimCost = 0.7f;  // paper: 0.3; 0.45-1.0 measured equal here (fewer triangle tests per ray). Override: PBRGPU_PRIM_COST
struct Widening {
  std::vector<float> C;        // [node][7]: C(n, i+1)
  std::vector<uint8_t> split;  // [node][8]: k of C_distribute(n, j) for j = 2..8 at [j-1]; 0 = "use C(n, j-1)"
  std::vector<uint8_t> leaf;   // C(n,1) was realised as a leaf
};

void compute_widening(const std::vector<TmpNode>& tmp, int maxLeaf, Widening* w) {
  const size_t n = tmp.size();
  w->C.assign(n * 7, 0.f);
  w->split.assign(n * 8, 0);
  w->leaf.assign(n, 0);
  const float inf = std::numeric_limits<float>::infinity();
  for (size_t idx = n; idx-- > 0;) {  // children always have larger indices than their parent
    const TmpNode& t = tmp[idx];
    const float A = half_area(t.box);
    float* C = &w->C[idx * 7];
    uint8_t* sp = &w->split[idx * 8];
    const float cLeaf = t.count <= maxLeaf ? A * t.count * kPrimCost : inf;
    if (t.left < 0) {
      for (int i = 0; i < 7; ++i) C[i] = cLeaf;
      w->leaf[idx] = 1;
      continue;
    }
    const float* L = &w->C[(size_t)t.left * 7];
    const float* R = &w->C[(size_t)t.right * 7];
    float dist[9];
    for (int j = 2; j <= 8; ++j) {
      float best = inf;
      int bk = 1;
      for (int k = 1; k < j; ++k) {
        if (k > 7 || j - k > 7) continue;
        float c = L[k - 1] + R[j - k - 1];
        if (c < best) { best = c; bk = k; }
      }
      dist[j] = best;
      sp[j - 1] = (uint8_t)bk;
    }
    const float cInternal = dist[8] + A * kNodeCost;
    if (cLeaf <= cInternal) { C[0] = cLeaf; w->leaf[idx] = 1; }
    else C[0] = cInternal;
    for (int i = 2; i <= 7; ++i) {
      if (dist[i] < C[i - 2]) C[i - 1] = dist[i];
      else { C[i - 1] = C[i - 2]; sp[i - 1] = 0; }
    }
  }
}

// roots of the cheapest forest of at most `i` trees that represents subtree m
void collect_forest(const std::vector<TmpNode>& tmp, const Widening& w, int32_t m, int i, int32_t* kids, int* nk) {
  while (i > 1 && (tmp[m].left < 0 || w.split[(size_t)m * 8 + i - 1] == 0)) --i;  // the extra roots were not worth it
  if (i == 1) { kids[(*nk)++] = m; return; }
  const int k = w.split[(size_t)m * 8 + i - 1];
  collect_forest(tmp, w, tmp[m].left, k, kids, nk);
  collect_forest(tmp, w, tmp[m].right, i - k, kids, nk);
}
}  // namespace

void build_bvh8(const float* verts, const int32_t* indices, int64_t nTriangles, int maxLeaf, BuiltBvh8* out, const Aabb* extra,
                int64_t nExtra) {
  maxLeaf = std::max(1, std::min(3, maxLeaf));
  std::vector<TmpNode> tmp;
  std::vector<int32_t> order;
  BinaryStats bs;
  const int64_t nTris = nTriangles + nExtra;
  const bool buildLog = std::getenv("PBRGPU_BUILD_LOG") != nullptr;  // development aid: phase times on stderr
  auto now = [] { return std::chrono::steady_clock::now(); };
  auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
  const auto t0 = now();
  build_binary(verts, indices, nTriangles, extra, nExtra, 1, tmp, order, &bs);  // one triangle per binary leaf: the widening forms the leaves
  const auto t1 = now();
  Widening wd;
  if (const char* e = std::getenv("PBRGPU_PRIM_COST")) kPrimCost = (float)std::atof(e);
  compute_widening(tmp, maxLeaf, &wd);
  const auto t2 = now();
  out->bounds = tmp[0].box;
  out->maxLeafSize = 0;
  out->nLeaves = 0;
  out->maxDepth = 0;
  out->nodes.clear();
  out->triOrder.clear();
  out->triOrder.reserve((size_t)nTris);
  out->nodes.reserve((size_t)std::max<int64_t>(nTris / 4, 1));

  struct Work { int32_t bin; int32_t wide; int32_t depth; };
  std::vector<Work> queue;
  out->nodes.emplace_back();
  queue.push_back({0, 0, 1});
  size_t head = 0;
  while (head < queue.size()) {
    Work w = queue[head++];
    out->maxDepth = std::max(out->maxDepth, (int)w.depth);
    // children of this wide node = the cheapest forest of <= 8 roots of the binary subtree
    int32_t kids[8];
    int nk = 0;
    if (tmp[w.bin].left < 0 || (w.bin == 0 && wd.leaf[0])) {
      kids[nk++] = w.bin;  // degenerate: the whole scene is one leaf
    } else {
      const int k = wd.split[(size_t)w.bin * 8 + 7];
      collect_forest(tmp, wd, tmp[w.bin].left, k, kids, &nk);
      collect_forest(tmp, wd, tmp[w.bin].right, 8 - k, kids, &nk);
    }
    auto is_leaf_kid = [&](int32_t m) { return tmp[m].left < 0 || wd.leaf[m] != 0; };
    const Aabb& box = tmp[w.bin].box;
    // slot assignment: slot s is visited first by rays with direction sign (s&4 ? - : +, s&2 ? - : +, s&1 ? - : +);
    // put there the child that lies most upstream along that direction (greedy min-cost matching).
    float cost[8][8];
    float pc[3] = {0.5f * (box.lo[0] + box.hi[0]), 0.5f * (box.lo[1] + box.hi[1]), 0.5f * (box.lo[2] + box.hi[2])};
    for (int i = 0; i < nk; ++i) {
      const Aabb& cb = tmp[kids[i]].box;
      float c[3] = {0.5f * (cb.lo[0] + cb.hi[0]) - pc[0], 0.5f * (cb.lo[1] + cb.hi[1]) - pc[1], 0.5f * (cb.lo[2] + cb.hi[2]) - pc[2]};
      for (int s = 0; s < 8; ++s) {
        float dx = (s & 4) ? -1.f : 1.f, dy = (s & 2) ? -1.f : 1.f, dz = (s & 1) ? -1.f : 1.f;
        cost[i][s] = c[0] * dx + c[1] * dy + c[2] * dz;
      }
    }
    int slotOf[8], childAt[8];
    for (int i = 0; i < 8; ++i) { slotOf[i] = -1; childAt[i] = -1; }
    for (int round = 0; round < nk; ++round) {
      int bi = -1, bsl = -1;
      float bc = std::numeric_limits<float>::infinity();
      for (int i = 0; i < nk; ++i) {
        if (slotOf[i] >= 0) continue;
        for (int s = 0; s < 8; ++s) {
          if (childAt[s] >= 0) continue;
          if (cost[i][s] < bc) { bc = cost[i][s]; bi = i; bsl = s; }
        }
      }
      slotOf[bi] = bsl;
      childAt[bsl] = bi;
    }
    // quantisation frame
    Bvh8Node nd;
    std::memset(&nd, 0, sizeof(nd));
    double scale[3];
    for (int a = 0; a < 3; ++a) {
      nd.origin[a] = box.lo[a];
      double ext = (double)box.hi[a] - (double)box.lo[a];
      int e = -60;
      if (ext > 0) {
        e = (int)std::ceil(std::log2(ext / 255.0));
        while (std::ldexp(255.0, e) < ext) ++e;
      }
      e = std::max(-100, std::min(100, e));
      nd.e[a] = (uint8_t)(e + 127);
      scale[a] = std::ldexp(1.0, e);
    }
    int nInner = 0;
    for (int s = 0; s < 8; ++s)
      if (childAt[s] >= 0 && !is_leaf_kid(kids[childAt[s]])) ++nInner;
    nd.childBase = (uint32_t)out->nodes.size();
    nd.triBase = (uint32_t)out->triOrder.size();
    out->nodes.resize(out->nodes.size() + (size_t)nInner);
    int innerRank = 0, triOff = 0;
    uint8_t* qlo[3] = {nd.qlox, nd.qloy, nd.qloz};
    uint8_t* qhi[3] = {nd.qhix, nd.qhiy, nd.qhiz};
    for (int s = 0; s < 8; ++s) {
      if (childAt[s] < 0) {
        for (int a = 0; a < 3; ++a) { qlo[a][s] = 255; qhi[a][s] = 0; }
        continue;
      }
      const TmpNode& ch = tmp[kids[childAt[s]]];
      for (int a = 0; a < 3; ++a) {
        double lo = std::floor(((double)ch.box.lo[a] - (double)nd.origin[a]) / scale[a]);
        double hi = std::ceil(((double)ch.box.hi[a] - (double)nd.origin[a]) / scale[a]);
        qlo[a][s] = (uint8_t)std::max(0.0, std::min(255.0, lo));
        qhi[a][s] = (uint8_t)std::max(0.0, std::min(255.0, hi));
      }
      if (!is_leaf_kid(kids[childAt[s]])) {
        nd.imask |= (uint8_t)(1u << s);
        nd.meta[s] = (uint8_t)((1u << 5) | (24u + (unsigned)s));
        queue.push_back({kids[childAt[s]], (int32_t)nd.childBase + innerRank, w.depth + 1});
        ++innerRank;
      } else {
        unsigned unary = ch.count == 1 ? 1u : (ch.count == 2 ? 3u : 7u);
        nd.meta[s] = (uint8_t)((unary << 5) | (unsigned)triOff);
        for (int i = 0; i < ch.count; ++i) out->triOrder.push_back(order[ch.first + i]);
        triOff += ch.count;
        out->nLeaves++;
        out->maxLeafSize = std::max(out->maxLeafSize, (int)ch.count);
      }
    }
    out->nodes[w.wide] = nd;
  }
  if (buildLog)
    fprintf(stderr, "build_bvh8: binary SAH tree %.1f ms (%zu nodes), widening %.1f ms, 8-wide layout + quantisation %.1f ms\n", ms(t0, t1), tmp.size(),
            ms(t1, t2), ms(t2, now()));
}

}  // namespace pbr
