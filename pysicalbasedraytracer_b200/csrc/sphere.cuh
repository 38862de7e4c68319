// sphere.cuh - analytic sphere primitive.  The reference's PBR::Sphere (Shape/Sphere.cpp:8-37) is a stub; the
// arithmetic here is that of the CPU twin host/AnalyticSphere.h (a subclass of the reference's Shape that implements what
// Shape/README.md section 3 documents), operation for operation.  PARITY UNPINNED by the reference; pins: twin == GPU.
// A sphere lives in the triangle arrays of the BVH leaf order as
//   word 0 = centre.xyz | caller's primitive id     word 1 = radius, -, - | material word with kPrimSphereBit     word 2 = -,-,- | area-light id
#pragma once
#include "device_types.cuh"

namespace pbr {

// AnalyticSphere::SolveHit: A t^2 + B t + C = 0 in double from the float object-space ray
PBR_D bool sphere_solve(V3 o, V3 d, float radius, float tMax, float* tHit) {
  const double ox = o.x, oy = o.y, oz = o.z, dx = d.x, dy = d.y, dz = d.z;
  const double a = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
  const double b = __dmul_rn(2.0, __dadd_rn(__dadd_rn(__dmul_rn(ox, dx), __dmul_rn(oy, dy)), __dmul_rn(oz, dz)));
  const double c = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(ox, ox), __dmul_rn(oy, oy)), __dmul_rn(oz, oz)),
                             -__dmul_rn((double)radius, (double)radius));
  const double discriminant = __dadd_rn(__dmul_rn(b, b), -__dmul_rn(__dmul_rn(4.0, a), c));
  if (!(discriminant > 0)) return false;
  const double root = sqrt(discriminant);
  const double den = __dmul_rn(2.0, a);
  const float f0 = (float)(__dadd_rn(-b, -root) / den), f1 = (float)(__dadd_rn(-b, root) / den);
  if (f0 > 0 && f0 < tMax) { *tHit = f0; return true; }
  if (f1 > 0 && f1 < tMax) { *tHit = f1; return true; }
  return false;
}

// AnalyticSphere::Intersect's SurfaceInteraction for a hit at parameter t of the WORLD ray (o, d)
struct SphereHit {
  V3 p, pError, n, dpdu;
};
PBR_D SphereHit sphere_hit(V3 center, float radius, bool flipNormal, V3 o, V3 d, float t) {
  const V3 oo = o - center;
  V3 pHit = oo + d * t;
  const float dist = sqrtf(pHit.x * pHit.x + pHit.y * pHit.y + pHit.z * pHit.z);
  const float f = radius / dist;
  pHit = V3(pHit.x * f, pHit.y * f, pHit.z * f);
  if (pHit.x == 0 && pHit.y == 0) pHit.x = 1e-5f * radius;
  const float cosTheta = clampf(pHit.z / radius, -1.f, 1.f);
  const float sinTheta = sqrtf(stdmax(0.f, 1 - cosTheta * cosTheta));
  const float zRadius = sqrtf(pHit.x * pHit.x + pHit.y * pHit.y);
  const float invZRadius = 1 / zRadius;
  const float cosPhi = pHit.x * invZRadius, sinPhi = pHit.y * invZRadius;
  const float twoPi = 2 * kPi;
  SphereHit h;
  h.dpdu = V3(-twoPi * pHit.y, twoPi * pHit.x, 0.f);
  const float dth = 0.f - kPi;  // thetaMax - thetaMin
  const V3 dpdv(dth * (pHit.z * cosPhi), dth * (pHit.z * sinPhi), dth * (-radius * sinTheta));
  h.p = pHit + center;
  h.pError = gamma_f(5) * (vabs(pHit) + vabs(h.p));
  V3 n = normalize(cross(h.dpdu, dpdv));
  if (flipNormal) n = -n;
  h.n = n;
  return h;
}

}  // namespace pbr
