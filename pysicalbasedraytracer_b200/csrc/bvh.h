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

// 8-wide compressed node (80 bytes = five 16-byte loads), after Ylitie, Karras & Laine,
// "Efficient Incoherent Ray Traversal on GPUs Through Compressed Wide BVHs" (HPG 2017):
//   n0 = origin.xyz | ex ey ez imask          n1 = childBase, triBase, meta[0..3], meta[4..7]
//   n2 = qlo.x[0..7], qlo.y[0..7]             n3 = qlo.z[0..7], qhi.x[0..7]      n4 = qhi.y[0..7], qhi.z[0..7]
// Child box k = origin + 2^(e-127) * [qlo_k, qhi_k] per axis (conservative: floor / ceil).
// meta[k]: empty 0 | inner (1<<5)|(24+k) | leaf (unary count 1/3/7 <<5)|(offset of first triangle from triBase).
// Inner children are stored contiguously from childBase in slot order; slot k is visited first by
// rays whose direction signs match the slot (traversal order = descending k ^ octant).
struct alignas(16) Bvh8Node {
  float origin[3];
  uint8_t e[3];
  uint8_t imask;
  uint32_t childBase;
  uint32_t triBase;
  uint8_t meta[8];
  uint8_t qlox[8], qloy[8], qloz[8], qhix[8], qhiy[8], qhiz[8];
};
static_assert(sizeof(Bvh8Node) == 80, "Bvh8Node must be 80 bytes");

struct BuiltBvh8 {
  std::vector<Bvh8Node> nodes;    // nodes[0] is the root
  std::vector<int32_t> triOrder;  // leaf-order position -> caller's triangle index
  Aabb bounds;
  int maxLeafSize = 0;
  int64_t nLeaves = 0;
  int maxDepth = 0;               // wide-tree depth = upper bound on traversal stack entries
};

// verts: float[nVerts*3], indices: int[nTris*3].  Deterministic binned-SAH build.  `extra` / nExtra: bounds of further
// primitives (analytic spheres), which get the primitive ids nTris .. nTris + nExtra - 1.
void build_bvh8(const float* verts, const int32_t* indices, int64_t nTris, int maxLeaf, BuiltBvh8* out,
                const Aabb* extra = nullptr, int64_t nExtra = 0);

}  // namespace pbr
