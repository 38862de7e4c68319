// trace.cuh - ray / triangle and ray / BVH queries.
//   Triangle::Intersect hit test   Shape/Triangle.cpp:66-145   (identical arithmetic, bit-exact t,b0,b1,b2)
//   Triangle::IntersectP           Shape/Triangle.cpp:262-350
//   BVHAccel::Intersect/IntersectP Accelerator/BVHAccel.cpp:285-366 (replaced: our own compressed 8-wide
//   BVH with a conservative slab test; the reference's slab test Geometry.h:1437-1468 has no gamma
//   widening, ours is widened, so the set of triangles we test is a superset of what the reference tests).
#pragma once
#include <cuda_fp16.h>
#include "device_types.cuh"
#include "sphere.cuh"

namespace pbr {

// Ray-dependent part of the watertight test, hoisted out of the per-triangle loop.  The reference
// recomputes these per triangle from the same inputs (Triangle.cpp:74-86) - same values.
struct RayPrep {
  V3 o;
  int kz;  // (kx, ky, kz) is the cyclic rotation of (0, 1, 2) that ends in kz
  float Sx, Sy, Sz;
};

// Permute(v, kx, ky, kz) (Geometry.h) for the cyclic (kx, ky, kz): two selects per component instead of a
// run-time indexed component fetch (which cost ~10 % of the traversal kernel's instructions)
PBR_D V3 permute_cyclic(V3 v, int kz) {
  const bool r0 = kz == 0, r1 = kz == 1;  // kz == 0: (y,z,x)   kz == 1: (z,x,y)   kz == 2: (x,y,z)
  return V3(r0 ? v.y : (r1 ? v.z : v.x), r0 ? v.z : (r1 ? v.x : v.y), r0 ? v.x : (r1 ? v.y : v.z));
}

PBR_D RayPrep prep_ray(V3 o, V3 d) {
  RayPrep rp;
  rp.o = o;
  int kz = max_dimension(vabs(d));
  // kx = kz + 1 (mod 3), ky = kx + 1 (mod 3)   (Triangle.cpp:74-78)
  V3 dp = permute_cyclic(d, kz);
  float dx = dp.x, dy = dp.y, dz = dp.z;
  rp.kz = kz;
  rp.Sx = -dx / dz;
  rp.Sy = -dy / dz;
  rp.Sz = 1.f / dz;
  return rp;
}

// Returns true on hit and writes t, b0, b1, b2.  tMax is ray.tMax at the time of the call.
PBR_D bool intersect_triangle(const RayPrep& rp, float tMax, V3 p0, V3 p1, V3 p2, float* tOut, float* b0Out,
                              float* b1Out, float* b2Out) {
  V3 q0 = p0 - rp.o, q1 = p1 - rp.o, q2 = p2 - rp.o;
  V3 p0t = permute_cyclic(q0, rp.kz);
  V3 p1t = permute_cyclic(q1, rp.kz);
  V3 p2t = permute_cyclic(q2, rp.kz);
  p0t.x += rp.Sx * p0t.z;
  p0t.y += rp.Sy * p0t.z;
  p1t.x += rp.Sx * p1t.z;
  p1t.y += rp.Sy * p1t.z;
  p2t.x += rp.Sx * p2t.z;
  p2t.y += rp.Sy * p2t.z;
  float e0 = p1t.x * p2t.y - p1t.y * p2t.x;
  float e1 = p2t.x * p0t.y - p2t.y * p0t.x;
  float e2 = p0t.x * p1t.y - p0t.y * p1t.x;
  if (e0 == 0.0f || e1 == 0.0f || e2 == 0.0f) {  // fp64 fallback (Triangle.cpp:98-109)
    double p2txp1ty = (double)p2t.x * (double)p1t.y;
    double p2typ1tx = (double)p2t.y * (double)p1t.x;
    e0 = (float)(p2typ1tx - p2txp1ty);
    double p0txp2ty = (double)p0t.x * (double)p2t.y;
    double p0typ2tx = (double)p0t.y * (double)p2t.x;
    e1 = (float)(p0typ2tx - p0txp2ty);
    double p1txp0ty = (double)p1t.x * (double)p0t.y;
    double p1typ0tx = (double)p1t.y * (double)p0t.x;
    e2 = (float)(p1typ0tx - p1txp0ty);
  }
  if ((e0 < 0 || e1 < 0 || e2 < 0) && (e0 > 0 || e1 > 0 || e2 > 0)) return false;
  float det = e0 + e1 + e2;
  if (det == 0) return false;
  p0t.z *= rp.Sz;
  p1t.z *= rp.Sz;
  p2t.z *= rp.Sz;
  float tScaled = e0 * p0t.z + e1 * p1t.z + e2 * p2t.z;
  if (det < 0 && (tScaled >= 0 || tScaled < tMax * det)) return false;
  else if (det > 0 && (tScaled <= 0 || tScaled > tMax * det)) return false;
  float invDet = 1 / det;
  float b0 = e0 * invDet, b1 = e1 * invDet, b2 = e2 * invDet;
  float t = tScaled * invDet;
  // conservative bound on the error of t (Triangle.cpp:133-145).  MaxComponent(Abs(v)) as hardware max of magnitudes: the same
  // value as the reference's compare-and-select for every non-NaN input (the arguments are magnitudes, so no -0 either), and a
  // NaN here means t is NaN, which fails `t <= deltaT` whatever deltaT is
  float maxZt = fmaxf(fmaxf(fabsf(p0t.z), fabsf(p1t.z)), fabsf(p2t.z));
  float deltaZ = gamma_f(3) * maxZt;
  float maxXt = fmaxf(fmaxf(fabsf(p0t.x), fabsf(p1t.x)), fabsf(p2t.x));
  float maxYt = fmaxf(fmaxf(fabsf(p0t.y), fabsf(p1t.y)), fabsf(p2t.y));
  float deltaX = gamma_f(5) * (maxXt + maxZt);
  float deltaY = gamma_f(5) * (maxYt + maxZt);
  float deltaE = 2 * (gamma_f(2) * maxXt * maxYt + deltaY * maxXt + deltaX * maxYt);
  float maxE = fmaxf(fmaxf(fabsf(e0), fabsf(e1)), fabsf(e2));
  float deltaT = 3 * (gamma_f(3) * maxE * maxZt + deltaE * maxZt + deltaZ * maxE) * fabsf(invDet);
  if (t <= deltaT) return false;
  *tOut = t; *b0Out = b0; *b1Out = b1; *b2Out = b2;
  return true;
}

struct alignas(8) HitRec {
  int tri;  // leaf-order triangle index, -1 = miss
  float t, b0, b1, b2;
  uint32_t w1;  // word 1 .w of the hit triangle's record (material id | type | flags), free with the vertex load
};

PBR_D V3 ld3(const float4* p, int i) {
  float4 v = __ldg(p + i);
  return V3(v.x, v.y, v.z);
}

// ---------------------------------------------------------------------------------------------
// Compressed 8-wide BVH traversal (node layout: bvh.h).  Per-lane state machine after Ylitie et al.
// 2017: a *node group* (childBase, hit bits of inner children in [31:24] | imask in [7:0]) and a
// *triangle group* (triBase, hit bits of up to 24 triangles).  The traversal stack holds node groups
// (and postponed triangle groups); its first kSmemStack levels live in shared memory, one column per
// lane ([level][lane] -> conflict-free), deeper levels spill to a small local array.
// ---------------------------------------------------------------------------------------------
#ifndef PBR_NODE_HALF2
#define PBR_NODE_HALF2 1
#endif
#ifndef PBR_H2_RATIO
#define PBR_H2_RATIO 2048.f
#endif
#ifndef PBR_SMEM_STACK
#define PBR_SMEM_STACK 8
#endif
constexpr int kSmemStack = PBR_SMEM_STACK;
constexpr int kLocalStack = 36;

constexpr unsigned kTravWideRange = 0x80000000u;  // flag in Trav::octinv4 (bit 31 is masked out wherever the octant bytes are used)

struct Trav {
  V3 dir;                     // ray direction: read by the analytic-sphere test only (dead in every other instantiation)
  V3 o;
  float idx, idy, idz;        // reciprocal direction (|d| clamped away from 0)
  unsigned octinv4;           // inverse octant replicated in 4 bytes
  unsigned f32one;            // 0x3F800000, from RenderParams::f32one (see q2f)
  RayPrep rp;
  float tMax;
  uint2 ngroup, tgroup;
  int sp;
  HitRec hit;
};
// deep levels of the traversal stack (beyond kSmemStack): a separate local array so that the rest of
// the traversal state stays in registers
struct TravSpill {
  uint2 lstack[kLocalStack];
};

PBR_D float safe_rcp(float d) {
  // |1/d| capped at 1e24: an axis-parallel ray gets +-huge slab distances instead of inf/NaN.  The reciprocal only feeds the
  // conservative box test, so the one-instruction approximation (relative error <= 2^-23, PTX ISA rcp.approx.f32) replaces the
  // IEEE division (14 instructions each, three per ray); kSlabErr below accounts for it.
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(fmaxf(fabsf(d), 1e-24f)));
  return d < 0.f ? -r : r;
}
// Rounding-error bound of the slab distances, relative to |B| + |A| (cw_node_step).  With u = 2^-24: the reciprocal carries
// 2u, B = fl(fl(origin - o) * idx) two more roundings (4u on |B| in all), fl(B - A) and fl(. -+ err) one each on |B| + |A|,
// the final FMA u on |t| <= |B| + |A|: at most 7u|B| + 3u|A| = 4.2e-7 |B| + 1.8e-7 |A|.
constexpr float kSlabErr = 5e-7f;

PBR_D void trav_init(Trav& T, V3 o, V3 d, float tMax, unsigned f32one) {
  T.o = o;
  T.dir = d;
  T.f32one = f32one;
  T.idx = safe_rcp(d.x); T.idy = safe_rcp(d.y); T.idz = safe_rcp(d.z);
  T.octinv4 = ((d.x < 0.f ? 0u : 4u) | (d.y < 0.f ? 0u : 2u) | (d.z < 0.f ? 0u : 1u)) * 0x01010101u;
#if PBR_NODE_HALF2
  {  // kTravWideRange: a direction with a component below 2^-11 of its largest one takes the fp32 box test (cw_node_test_h2)
    const float ax = fabsf(d.x), ay = fabsf(d.y), az = fabsf(d.z);
    if (!(fminf(fminf(ax, ay), az) * PBR_H2_RATIO >= fmaxf(fmaxf(ax, ay), az))) T.octinv4 |= kTravWideRange;
  }
#endif
  T.rp = prep_ray(o, d);
  T.tMax = tMax;
  T.ngroup = make_uint2(0u, 0x80000000u);
  T.tgroup = make_uint2(0u, 0u);
  T.sp = 0;
  T.hit.tri = -1;
  T.hit.t = tMax;
  T.hit.b0 = T.hit.b1 = T.hit.b2 = 0.f;
  T.hit.w1 = 0u;
}

// STRIDE: distance (in uint2) between two levels of a lane's column: 32 for a [warp][level][lane] array, the block size for
// a [level][thread] one
template <int STRIDE = 32>
PBR_D void trav_push(Trav& T, uint2* sstack, TravSpill& S, uint2 v) {
  if (T.sp < kSmemStack) sstack[T.sp * STRIDE] = v;
  else S.lstack[T.sp - kSmemStack] = v;
  T.sp++;
}
template <int STRIDE = 32>
PBR_D uint2 trav_pop(Trav& T, const uint2* sstack, const TravSpill& S) {
  T.sp--;
  return (T.sp < kSmemStack) ? sstack[T.sp * STRIDE] : S.lstack[T.sp - kSmemStack];
}

// byte J of `w` -> float in [1,2): 1 + byte / 32768   (one PRMT, no int->float conversion).  `one` holds
// 0x3F800000 in a REGISTER (made opaque by the caller): PRMT takes one immediate, and with the constant as
// the immediate ptxas re-materialised the selector in a register before almost every PRMT.
template <int J>
PBR_D float q2f(unsigned w, unsigned one) { return __uint_as_float(__byte_perm(w, one, 0x7604u | (J << 4))); }

// The slab test of one node's 8 quantised child boxes in fp32: returns the hit mask (inner children in [31:24] in traversal
// order, triangles in [23:0]).  Conservative: the rounding error of the fused evaluation is bounded per axis and folded into the
// offsets (Bn = B - err, Bf = B + err).
PBR_D unsigned cw_node_test_f32(const float4 n0, const float4 n1, const float4 n2, const float4 n3, const float4 n4, const Trav& T) {
  const unsigned ei = __float_as_uint(n0.w);
  // A = 32768 * 2^e / d (exact scalings of 1/d: the factor 2^15 is added to the exponent byte), B = (origin - o) / d
  const float Ax = __uint_as_float(((ei & 0xffu) + 15u) << 23) * T.idx;
  const float Ay = __uint_as_float((((ei >> 8) & 0xffu) + 15u) << 23) * T.idy;
  const float Az = __uint_as_float((((ei >> 16) & 0xffu) + 15u) << 23) * T.idz;
  const float Bx = (n0.x - T.o.x) * T.idx, By = (n0.y - T.o.y) * T.idy, Bz = (n0.z - T.o.z) * T.idz;
  const float ex = kSlabErr * (fabsf(Bx) + fabsf(Ax)), ey = kSlabErr * (fabsf(By) + fabsf(Ay)), ez = kSlabErr * (fabsf(Bz) + fabsf(Az));
  const float Bnx = (Bx - Ax) - ex, Bfx = (Bx - Ax) + ex;
  const float Bny = (By - Ay) - ey, Bfy = (By - Ay) + ey;
  const float Bnz = (Bz - Az) - ez, Bfz = (Bz - Az) + ez;
  const bool px = T.idx >= 0.f, py = T.idy >= 0.f, pz = T.idz >= 0.f;
  unsigned hitmask = 0;
  const unsigned one = T.f32one;  // 0x3F800000 read from a kernel parameter, so ptxas cannot fold it (see q2f)
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const unsigned meta4 = __float_as_uint(h ? n1.w : n1.z);
    const unsigned isInner4 = (meta4 & (meta4 << 1)) & 0x10101010u;
    const unsigned innerMask4 = (isInner4 >> 4) * 0xffu;
    const unsigned bitIndex4 = (meta4 ^ (T.octinv4 & innerMask4)) & 0x1f1f1f1fu;
    const unsigned childBits4 = (meta4 >> 5) & 0x07070707u;
    const unsigned lox = __float_as_uint(h ? n2.y : n2.x), loy = __float_as_uint(h ? n2.w : n2.z);
    const unsigned loz = __float_as_uint(h ? n3.y : n3.x), hix = __float_as_uint(h ? n3.w : n3.z);
    const unsigned hiy = __float_as_uint(h ? n4.y : n4.x), hiz = __float_as_uint(h ? n4.w : n4.z);
    const unsigned nx = px ? lox : hix, fx = px ? hix : lox;
    const unsigned ny = py ? loy : hiy, fy = py ? hiy : loy;
    const unsigned nz = pz ? loz : hiz, fz = pz ? hiz : loz;
    // branch-free mask update: (hit ? childBits : 0) << bitIndex, bytes extracted with one PRMT each
#define PBR_CW_CHILD(J)                                                                               \
    {                                                                                                 \
      const float tn = fmaxf(fmaxf(fmaf(q2f<J>(nx, one), Ax, Bnx), fmaf(q2f<J>(ny, one), Ay, Bny)),             \
                             fmaxf(fmaf(q2f<J>(nz, one), Az, Bnz), 0.f));                                  \
      const float tf = fminf(fminf(fmaf(q2f<J>(fx, one), Ax, Bfx), fmaf(q2f<J>(fy, one), Ay, Bfy)),             \
                             fminf(fmaf(q2f<J>(fz, one), Az, Bfz), T.tMax));                               \
      const unsigned cb = __byte_perm(childBits4, 0u, 0x4440u | J);                                   \
      const unsigned bi = __byte_perm(bitIndex4, 0u, 0x4440u | J);                                    \
      hitmask |= (tn <= tf ? cb : 0u) << bi;                                                          \
    }
    PBR_CW_CHILD(0) PBR_CW_CHILD(1) PBR_CW_CHILD(2) PBR_CW_CHILD(3)
#undef PBR_CW_CHILD
  }
  return hitmask;
}

#if PBR_NODE_HALF2
// The same test with the 48 slab distances evaluated two children at a time in half precision (HFMA2 / HMNMX2 / HSET2.LE):
// k_trace is bound by the ALU pipe (PRMT, FMNMX, SEL, LOP3, SHF: one warp instruction per two cycles - scripts/ubench/pipes.cu),
// and the fp32 test spends 88 ALU instructions per node on byte unpacking and min / max; this one 52.
//   * Pivot.  Distances are taken relative to tref, the distance at which the ray enters the node's own box (fp32), and scaled
//     by a power of two sg chosen per node so that the steepest axis' distance per grid unit is in [2^-10, 2^-9): every value
//     that decides the test is then O(1) or smaller, and the values of the other axes keep 11 significant bits down to 2^-14.
//   * A byte pair becomes two half DENORMALS q * 2^-24 with ONE PRMT (bytes b0,0,b1,0); the factor 2^24 goes into the slope.
//   * Conservative.  With u = 2^-11 the half stage commits at most 2u(1+u)(|q a| + |b|) + 3 * 2^-25 (rounding of the slope, of
//     the offset and of the fused result, each relative or - in the denormal range - absolute); kH = 2.3u and kAbs = 2^-23
//     cover it (and the fp32 roundings of this function, relative 2^-24 each), on top of the fp32 bound kSlabErr on the
//     un-pivoted distances.  Offsets that overflow become +-inf (never NaN: slopes and bytes are finite), which compare correctly.
//     A false positive costs a node visit; a false negative cannot happen.
//   * Range.  An axis whose distances are more than ~2^12 times smaller than the steepest one's sinks into the absolute error:
//     the test stays conservative but stops culling along that axis.  Rays with such directions (kTravWideRange, set by
//     trav_init: 0.15 % of uniformly distributed directions; every ray of an orthographic camera) take the fp32 test.
// Node exponents are e in [-100, 100] (bvh_build.cpp) and 1/d in [~0.01, 1e24], so the exponent arithmetic below stays in range.
PBR_D unsigned h2_as_u(__half2 h) { return *reinterpret_cast<unsigned*>(&h); }
PBR_D __half2 u_as_h2(unsigned u) { return *reinterpret_cast<__half2*>(&u); }
PBR_D unsigned h2_le_mask(__half2 a, __half2 b) {  // 0xFFFF per half where a <= b
  unsigned m;
  asm("set.le.u32.f16x2 %0, %1, %2;" : "=r"(m) : "r"(h2_as_u(a)), "r"(h2_as_u(b)));
  return m;
}
PBR_D unsigned cw_node_test_h2(const float4 n0, const float4 n1, const float4 n2, const float4 n3, const float4 n4, const Trav& T) {
  constexpr float kH = 2.3f * 4.8828125e-4f;   // 2.3 * 2^-11
  constexpr float kAbs = 1.1920929e-7f;        // 2^-23
  constexpr float kQ = 255.f * 5.9604645e-8f;  // largest byte as a half denormal: 255 * 2^-24
  const unsigned ei = __float_as_uint(n0.w);
  // a = 2^e / d: distance per grid unit;  b = (origin - o) / d: distance of the grid plane 0
  const float ax = __uint_as_float((ei & 0xffu) << 23) * T.idx;
  const float ay = __uint_as_float(((ei >> 8) & 0xffu) << 23) * T.idy;
  const float az = __uint_as_float(((ei >> 16) & 0xffu) << 23) * T.idz;
  const float bx = (n0.x - T.o.x) * T.idx, by = (n0.y - T.o.y) * T.idy, bz = (n0.z - T.o.z) * T.idz;
  // pivot: entry distance of the node's own box [0, 255]^3, not below 0
  const float tref = fmaxf(fmaxf(fminf(bx, fmaf(255.f, ax, bx)), fminf(by, fmaf(255.f, ay, by))), fmaxf(fminf(bz, fmaf(255.f, az, bz)), 0.f));
  // scale: S = 2^(141 - E) for the steepest slope amax in [2^(E-127), 2^(E-126)), i.e. amax * S in [2^14, 2^15); sg = S * 2^-24
  const unsigned Eb = __float_as_uint(fmaxf(fmaxf(fabsf(ax), fabsf(ay)), fabsf(az))) & 0x7F800000u;
  const float S = __uint_as_float(0x86000000u - Eb), sg = __uint_as_float(0x7A000000u - Eb);
  const float tsg = tref * sg;
  const float sax = ax * S, say = ay * S, saz = az * S;
  const float sbx = fmaf(bx, sg, -tsg), sby = fmaf(by, sg, -tsg), sbz = fmaf(bz, sg, -tsg);
  // error bound per axis: fp32 part kSlabErr (|b| + 512 |a|) sg, half part kH (255 |a| sg + |b'|) + kAbs
  const float c32 = kSlabErr * sg, ca = (kH * kQ);
  const float ex = fmaf(c32, fmaf(512.f, fabsf(ax), fabsf(bx)), fmaf(ca, fabsf(sax), fmaf(kH, fabsf(sbx), kAbs)));
  const float ey = fmaf(c32, fmaf(512.f, fabsf(ay), fabsf(by)), fmaf(ca, fabsf(say), fmaf(kH, fabsf(sby), kAbs)));
  const float ez = fmaf(c32, fmaf(512.f, fabsf(az), fabsf(bz)), fmaf(ca, fabsf(saz), fmaf(kH, fabsf(sbz), kAbs)));
  // the ray's own interval [0, tMax] relative to the pivot, rounded outwards
  const float z0 = fmaf(-tsg, 1.f + kH, -kAbs);
  const float tm0 = fmaf(T.tMax, sg, -tsg), tm = fmaf(kH, fabsf(tm0), tm0) + kAbs;
  const __half2 hA0 = __floats2half2_rn(sax, say), hA1 = __floats2half2_rn(saz, z0);
  const __half2 hN0 = __floats2half2_rn(sbx - ex, sby - ey), hN1 = __floats2half2_rn(sbz - ez, tm);
  const __half2 hF0 = __floats2half2_rn(sbx + ex, sby + ey), hF1 = __floats2half2_rn(sbz + ez, 0.f);
  const __half2 AX = __low2half2(hA0), AY = __high2half2(hA0), AZ = __low2half2(hA1), Z0 = __high2half2(hA1);
  const __half2 BNX = __low2half2(hN0), BNY = __high2half2(hN0), BNZ = __low2half2(hN1), TM = __high2half2(hN1);
  const __half2 BFX = __low2half2(hF0), BFY = __high2half2(hF0), BFZ = __low2half2(hF1);
  const bool px = T.idx >= 0.f, py = T.idy >= 0.f, pz = T.idz >= 0.f;
  unsigned hitmask = 0;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const unsigned meta4 = __float_as_uint(h ? n1.w : n1.z);
    const unsigned isInner4 = (meta4 & (meta4 << 1)) & 0x10101010u;
    const unsigned innerMask4 = (isInner4 >> 4) * 0xffu;
    const unsigned bitIndex4 = meta4 ^ (T.octinv4 & innerMask4);  // bits [4:0] of each byte; the shifts below ignore the rest
    const unsigned lox = __float_as_uint(h ? n2.y : n2.x), loy = __float_as_uint(h ? n2.w : n2.z);
    const unsigned loz = __float_as_uint(h ? n3.y : n3.x), hix = __float_as_uint(h ? n3.w : n3.z);
    const unsigned hiy = __float_as_uint(h ? n4.y : n4.x), hiz = __float_as_uint(h ? n4.w : n4.z);
    const unsigned nx = px ? lox : hix, fx = px ? hix : lox;
    const unsigned ny = py ? loy : hiy, fy = py ? hiy : loy;
    const unsigned nz = pz ? loz : hiz, fz = pz ? hiz : loz;
    unsigned m[2];
#pragma unroll
    for (int p = 0; p < 2; ++p) {  // children (4h + 2p, 4h + 2p + 1)
      const unsigned sel = p ? 0x4342u : 0x4140u;
      const __half2 tnx = __hfma2(u_as_h2(__byte_perm(nx, 0u, sel)), AX, BNX), tny = __hfma2(u_as_h2(__byte_perm(ny, 0u, sel)), AY, BNY);
      const __half2 tnz = __hfma2(u_as_h2(__byte_perm(nz, 0u, sel)), AZ, BNZ);
      const __half2 tfx = __hfma2(u_as_h2(__byte_perm(fx, 0u, sel)), AX, BFX), tfy = __hfma2(u_as_h2(__byte_perm(fy, 0u, sel)), AY, BFY);
      const __half2 tfz = __hfma2(u_as_h2(__byte_perm(fz, 0u, sel)), AZ, BFZ);
      m[p] = h2_le_mask(__hmax2(__hmax2(tnx, tny), __hmax2(tnz, Z0)), __hmin2(__hmin2(tfx, tfy), __hmin2(tfz, TM)));
    }
    // 0xFF per hit child, then (childBits & hit) << bitIndex per child
    const unsigned cb4 = (meta4 >> 5) & 0x07070707u & __byte_perm(m[0], m[1], 0x6420u);
    // x << (s & 31) as a funnel shift: no masking of the shift amount, so byte J of bitIndex4 only has to be moved down
#define PBR_CW_CHILD(J) hitmask |= __funnelshift_l(0u, __byte_perm(cb4, 0u, 0x4440u | J), bitIndex4 >> (8 * J));
    PBR_CW_CHILD(0) PBR_CW_CHILD(1) PBR_CW_CHILD(2) PBR_CW_CHILD(3)
#undef PBR_CW_CHILD
  }
  return hitmask;
}
#endif

// One node step: pick the next inner child of the node group (front-to-back by octant), fetch its
// 80-byte node with five 16-byte loads, slab-test the 8 quantised child boxes, emit a new node group
// and triangle group.
template <int STRIDE = 32>
PBR_D void cw_node_step(const float4* __restrict__ nodes, Trav& T, uint2* sstack, TravSpill& S) {
  const unsigned hits = T.ngroup.y;
  const int bit = 31 - __clz(hits);
  T.ngroup.y &= ~(1u << bit);
  if (T.ngroup.y > 0x00FFFFFFu) trav_push<STRIDE>(T, sstack, S, T.ngroup);
  const unsigned slot = (unsigned)(bit - 24) ^ (T.octinv4 & 0xffu);
  const unsigned rel = __popc(hits & ~(0xFFFFFFFFu << slot));
  const float4* np = nodes + 5 * (size_t)(T.ngroup.x + rel);
  const float4 n0 = __ldg(np), n1 = __ldg(np + 1), n2 = __ldg(np + 2), n3 = __ldg(np + 3), n4 = __ldg(np + 4);
  T.ngroup.x = __float_as_uint(n1.x);
  T.tgroup.x = __float_as_uint(n1.y);
#if PBR_NODE_HALF2
  const unsigned hitmask = (T.octinv4 & kTravWideRange) ? cw_node_test_f32(n0, n1, n2, n3, n4, T) : cw_node_test_h2(n0, n1, n2, n3, n4, T);
#else
  const unsigned hitmask = cw_node_test_f32(n0, n1, n2, n3, n4, T);
#endif
  T.ngroup.y = (hitmask & 0xFF000000u) | (__float_as_uint(n0.w) >> 24);
  T.tgroup.y = hitmask & 0x00FFFFFFu;
}

// One triangle of the triangle group: the reference's watertight test, bit-exact (t, b0, b1, b2).
// `store`: where the hit record goes (k_trace keeps it in shared memory: it is written once or twice per ray and read once,
// and six registers less let a ninth block fit on the SM); NULL = T.hit.
// SPH: the leaf arrays may hold analytic spheres (kPrimSphereBit); their hit record is (prim, t, b0 = t, 0, 0).
template <bool SPH = false>
PBR_D bool cw_tri_step(const SceneDev& sc, Trav& T, HitRec* store = nullptr) {
  const int bit = 31 - __clz(T.tgroup.y);
  T.tgroup.y &= ~(1u << bit);
  const int ti = (int)T.tgroup.x + bit;
  const float4* rec = sc.tv + 3 * (size_t)ti;  // one 48-byte record
  const float4 v1 = __ldg(rec + 1);
  if (SPH && (__float_as_uint(v1.w) & kPrimSphereBit)) {
    const V3 center = ld3(rec, 0);
    float ts;
    if (!sphere_solve(T.rp.o - center, T.dir, v1.x, T.tMax, &ts)) return false;
    T.tMax = ts;
    HitRec& h = store ? *store : T.hit;
    h.tri = ti; h.t = ts; h.b0 = ts; h.b1 = 0.f; h.b2 = 0.f;
    h.w1 = __float_as_uint(v1.w);
    return true;
  }
  V3 p0 = ld3(rec, 0), p1(v1.x, v1.y, v1.z), p2 = ld3(rec, 2);
  float t, b0, b1, b2;
  if (intersect_triangle(T.rp, T.tMax, p0, p1, p2, &t, &b0, &b1, &b2)) {
    T.tMax = t;
    HitRec& h = store ? *store : T.hit;
    h.tri = ti; h.t = t; h.b0 = b0; h.b1 = b1; h.b2 = b2;
    h.w1 = __float_as_uint(v1.w);
    return true;
  }
  return false;
}

// Whole-ray traversal by one thread (parity hooks and small queries).  `sstack` = this lane's column
// of the warp's shared-memory stack.
template <bool ANY_HIT, bool COUNT, bool SPH = false>
PBR_D bool traverse(const SceneDev& sc, V3 o, V3 d, float tMax, HitRec* hit, uint2* sstack, unsigned* nNodes, unsigned* nTris) {
  Trav T;
  TravSpill S;
  trav_init(T, o, d, tMax, sc.f32one);
  bool done = false;
  while (!done) {
    if (T.ngroup.y > 0x00FFFFFFu) {
      cw_node_step(sc.nodes8, T, sstack, S);
      if (COUNT) (*nNodes)++;
    } else {
      T.tgroup = T.ngroup;
      T.ngroup = make_uint2(0u, 0u);
    }
    while (T.tgroup.y != 0u) {
      if (COUNT) (*nTris)++;
      if (cw_tri_step<SPH>(sc, T) && ANY_HIT) { T.tgroup.y = 0u; T.ngroup.y = 0u; T.sp = 0; }
    }
    if (T.ngroup.y <= 0x00FFFFFFu) {
      if (T.sp > 0) T.ngroup = trav_pop(T, sstack, S);
      else done = true;
    }
  }
  *hit = T.hit;
  return T.hit.tri >= 0;
}

}  // namespace pbr
