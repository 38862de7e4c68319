// bvh.h - host-side acceleration-structure build for the GPU traversal kernels.
// Replaces the reference's BVHAccel build (Accelerator/BVHAccel.cpp:57-283) with a layout designed
// for 16-byte vector loads on the GPU.  Only the *result of ray queries* has to agree with the
// reference (closest triangle / any occlusion); the tree itself is ours.
#pragma once
#include <cstdint>
#include <vector>

namespace pbr {

struct Aabb {
  float lo[3], hi[3];
};

// Binary node holding BOTH children's boxes (one 64-byte record = four 16-byte loads).
//   a = c0.lo.x c0.hi.x c0.lo.y c0.hi.y     b = c1.lo.x c1.hi.x c1.lo.y c1.hi.y
//   c = c0.lo.z c0.hi.z c1.lo.z c1.hi.z     d = child0, child1, 0, 0
// child >= 0: index of an interior node; child < 0: leaf, ~child = (firstTri << 2) | (count-1), count in 1..4.
struct alignas(16) Bvh2Node {
  float a[4], b[4], c[4];
  int32_t child[4];
};

struct BuiltBvh {
  std::vector<Bvh2Node> nodes;      // nodes[0] is the root
  std::vector<int32_t> triOrder;    // leaf-order position -> caller's triangle index
  Aabb bounds;
  int maxLeafSize = 0;
  int64_t nLeaves = 0;
  int maxDepth = 0;
};

// verts: float[nVerts*3], indices: int[nTris*3].  Deterministic binned-SAH build.
void build_bvh2(const float* verts, const int32_t* indices, int64_t nTris, int maxLeaf, BuiltBvh* out);

}  // namespace pbr
