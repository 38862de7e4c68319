// media.cuh - participating media for the VolPath wavefront.
//   HomogeneousMedium::{Tr,Sample}          Media/HomogeneousMedium.cpp:10-45
//   PhaseHG, HenyeyGreenstein::{p,Sample_p} Media/Medium.h:21-24, Media/Medium.cpp:7-28
//   MediumInterface / GetMedium             Media/Medium.h:51-59, Core/Interaction.h:48-50,
//                                           GeometricPrimitive::Intersect Core/Primitive.cpp:30-35
// Operation order follows the reference statement by statement; expf / logf / sinf / cosf are CUDA's (libm class
// accuracy), so results agree to the floating-point tolerance of the BSDF code, not bit for bit.
#pragma once
#include "device_types.cuh"

namespace pbr {

constexpr float kMaxFloat = 3.402823466e+38f;  // std::numeric_limits<float>::max()  (Core/PBR.h MaxFloat)

PBR_D Spec medium_sigma_t(const pbrgpu_medium& m) {  // sigma_t(sigma_s + sigma_a), HomogeneousMedium.h:15
  return Spec(m.sigma_s[0] + m.sigma_a[0], m.sigma_s[1] + m.sigma_a[1], m.sigma_s[2] + m.sigma_a[2]);
}
PBR_D Spec spec_exp(Spec s) { return Spec(expf(s.r), expf(s.g), expf(s.b)); }
PBR_D Spec spec_neg(Spec s) { return Spec(-s.r, -s.g, -s.b); }

// HomogeneousMedium::Tr: Exp(-sigma_t * min(ray.tMax * ray.d.Length(), MaxFloat))
PBR_D Spec medium_Tr(const pbrgpu_medium& m, float tMax, float dLen) {
  return spec_exp(spec_neg(medium_sigma_t(m)) * stdmin(tMax * dLen, kMaxFloat));
}

// HomogeneousMedium::Sample, split in two so that the distance can be drawn BEFORE the ray is traced:
//   reference:  found = Intersect(ray)            -> ray.tMax = tHit (or inf)
//               t = min(dist / |d|, ray.tMax); sampledMedium = t < ray.tMax
//   here:       tS = dist / |d| is drawn first and used as the ray's tMax.  A hit found below tS is the reference's hit
//               (the watertight test does not depend on tMax except for the range rejection) and means "no scattering";
//               no hit below tS means the reference's tHit is >= tS, i.e. it scatters at tS.  Same decisions, same
//               values (a hit at exactly t == tS is the only tie), but rays inside a dense medium are traced over
//               their short free path instead of to the far wall.
// The two sampler values are consumed in the reference's order: Intersect draws none.
PBR_D float medium_sample_distance(const pbrgpu_medium& m, float u0, float u1, float dLen) {
  const Spec sigma_t = medium_sigma_t(m);
  int channel = min((int)(u0 * 3), 2);
  const float st = channel == 0 ? sigma_t.r : (channel == 1 ? sigma_t.g : sigma_t.b);
  float dist = -logf(1 - u1) / st;
  return dist / dLen;
}
// t = min(dist / |d|, ray.tMax) and sampledMedium as established by the trace; returns the path weight.
PBR_D Spec medium_sample_weight(const pbrgpu_medium& m, float t, bool sampledMedium, float dLen) {
  const Spec sigma_t = medium_sigma_t(m);
  const Spec sigma_s(m.sigma_s[0], m.sigma_s[1], m.sigma_s[2]);
  Spec Tr = spec_exp(spec_neg(sigma_t) * stdmin(t, kMaxFloat) * dLen);
  Spec density = sampledMedium ? (sigma_t * Tr) : Tr;
  float pdf = 0;
  pdf += density.r; pdf += density.g; pdf += density.b;
  pdf *= 1 / (float)3;
  if (pdf == 0) pdf = 1;
  return sampledMedium ? (Tr * sigma_s / pdf) : (Tr / pdf);
}

PBR_D float phase_hg(float cosTheta, float g) {
  float denom = 1 + g * g + 2 * g * cosTheta;
  return kInv4Pi * (1 - g * g) / (denom * sqrtf(denom));
}
PBR_D float hg_p(V3 wo, V3 wi, float g) { return phase_hg(dot(wo, wi), g); }
PBR_D float hg_sample_p(V3 wo, V3* wi, float u0, float u1, float g) {
  float cosTheta;
  if (fabsf(g) < 1e-3)
    cosTheta = 1 - 2 * u0;
  else {
    float sqrTerm = (1 - g * g) / (1 + g - 2 * g * u0);
    cosTheta = -(1 + g * g - sqrTerm * sqrTerm) / (2 * g);
  }
  float sinTheta = sqrtf(stdmax(0.f, 1 - cosTheta * cosTheta));
  float phi = 2 * kPi * u1;
  V3 v1, v2;
  coordinate_system(wo, &v1, &v2);
  // SphericalDirection(sinTheta, cosTheta, phi, v1, v2, wo)  Core/Geometry.h:1510-1515
  *wi = sinTheta * cosf(phi) * v1 + sinTheta * sinf(phi) * v2 + cosTheta * wo;
  return phase_hg(cosTheta, g);
}

// SurfaceInteraction::mediumInterface as GeometricPrimitive::Intersect leaves it: the primitive's own interface when
// it is a medium transition, otherwise the medium the ray travels in on both sides.
PBR_D void isect_media(const SceneDev& sc, int tri, int rayMedium, int* inside, int* outside) {
  const uint32_t w = sc.triMedium ? __ldg(sc.triMedium + tri) : 0u;
  int in = (int)(w & 0xffffu) - 1, out = (int)(w >> 16) - 1;
  if (in == out) in = out = rayMedium;
  *inside = in; *outside = out;
}

}  // namespace pbr
