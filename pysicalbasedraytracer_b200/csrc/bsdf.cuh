// bsdf.cuh - BSDF / BxDF / Fresnel / Trowbridge-Reitz, register-resident (the reference heap-allocates
// these per hit: Material/MatteMaterial.cpp:17-26, Reflection.h:112-117).  Operation order follows
//   Material/Reflection.h:14-60, Reflection.cpp:7-318   (BxDFs, BSDF::f / Pdf / Sample_f)
//   Material/Fresnel.cpp:7-54                           (FrDielectric, FrConductor)
//   Material/Microfacet.cpp:116-292                     (TrowbridgeReitz D, Lambda, Sample_wh, Pdf)
//   Sampler/Sampling.h:57-63, Sampling.cpp:74-92        (CosineSampleHemisphere, ConcentricSampleDisk)
// libm calls (sinf/cosf/...) differ from glibc by ulps => results after them are tolerance-matched.
#pragma once
#include "vecmath.cuh"

namespace pbr {

enum : int {
  BSDF_REFLECTION = 1, BSDF_TRANSMISSION = 2, BSDF_DIFFUSE = 4, BSDF_GLOSSY = 8, BSDF_SPECULAR = 16, BSDF_ALL = 31
};
enum : int { LOBE_LAMBERT, LOBE_OREN_NAYAR, LOBE_SPEC_REFL, LOBE_SPEC_TRANS, LOBE_FRESNEL_SPEC, LOBE_MICRO_REFL, LOBE_MICRO_TRANS };
enum : int { FR_NOOP, FR_DIELECTRIC, FR_CONDUCTOR };

// Which lobe kinds / Fresnel kinds a material can ever create (Material/*.cpp).  The shading kernels are
// instantiated per material, so every `case` a material cannot reach is removed at compile time; MAT = -1
// keeps everything (generic use).
template <int MAT>
PBR_D constexpr bool mat_has_lobe(int kind) {
  return MAT < 0 ? true
       : MAT == 0 /*Matte*/   ? (kind == 0 /*LAMBERT*/ || kind == 1 /*OREN_NAYAR*/)
       : MAT == 1 /*Mirror*/  ? (kind == 2 /*SPEC_REFL*/)
       : MAT == 2 /*Glass*/   ? (kind >= 2)
       : MAT == 3 /*Metal*/   ? (kind == 5 /*MICRO_REFL*/)
       :            /*Plastic*/ (kind == 0 || kind == 5);
}
template <int MAT>
PBR_D constexpr bool mat_has_fresnel(int fr) {
  return MAT < 0 ? true
       : MAT == 1 ? fr == 0 /*NOOP*/
       : MAT == 2 ? fr == 1 /*DIELECTRIC*/
       : MAT == 3 ? fr == 2 /*CONDUCTOR*/
       : MAT == 4 ? fr == 1
       :            fr == 0;
}

struct Lobe {
  int kind, flags, fresnel;
  Spec R;              // R (reflection lobes) or T (transmission lobes); FresnelSpecular: R
  Spec T;              // FresnelSpecular: T ; conductor: k
  Spec cEta;           // conductor: etaT (etaI = 1)
  float frEtaI, frEtaT;  // dielectric Fresnel arguments
  float etaA, etaB;      // transmission lobes
  float alphax, alphay;  // Trowbridge-Reitz
  float A, B;            // Oren-Nayar
};

struct Bsdf {
  float eta;
  V3 ns, ng, ss, ts;
  int n;
  Lobe lobes[2];
  PBR_D V3 to_local(V3 v) const { return V3(dot(v, ss), dot(v, ts), dot(v, ns)); }
  PBR_D V3 to_world(V3 v) const {
    return V3(ss.x * v.x + ts.x * v.y + ns.x * v.z, ss.y * v.x + ts.y * v.y + ns.y * v.z,
              ss.z * v.x + ts.z * v.y + ns.z * v.z);
  }
  PBR_D int num_components(int flags) const {
    int c = 0;
    for (int i = 0; i < n; ++i) if ((lobes[i].flags & flags) == lobes[i].flags) ++c;
    return c;
  }
};

// ---- local-frame trigonometry (Reflection.h:14-37) ----
PBR_D float cos_theta(V3 w) { return w.z; }
PBR_D float cos2_theta(V3 w) { return w.z * w.z; }
PBR_D float abs_cos_theta(V3 w) { return fabsf(w.z); }
PBR_D float sin2_theta(V3 w) { return stdmax(0.f, 1.f - cos2_theta(w)); }
PBR_D float sin_theta(V3 w) { return sqrtf(sin2_theta(w)); }
PBR_D float tan_theta(V3 w) { return sin_theta(w) / cos_theta(w); }
PBR_D float tan2_theta(V3 w) { return sin2_theta(w) / cos2_theta(w); }
PBR_D float cos_phi(V3 w) { float s = sin_theta(w); return (s == 0) ? 1 : clampf(w.x / s, -1, 1); }
PBR_D float sin_phi(V3 w) { float s = sin_theta(w); return (s == 0) ? 0 : clampf(w.y / s, -1, 1); }
PBR_D float cos2_phi(V3 w) { return cos_phi(w) * cos_phi(w); }
PBR_D float sin2_phi(V3 w) { return sin_phi(w) * sin_phi(w); }
PBR_D bool same_hemisphere(V3 w, V3 wp) { return w.z * wp.z > 0; }
PBR_D V3 reflect_about(V3 wo, V3 n) { return -wo + (2 * dot(wo, n)) * n; }
PBR_D bool refract_dir(V3 wi, V3 n, float eta, V3* wt) {  // Reflection.h:41-51
  float cosThetaI = dot(n, wi);
  float sin2ThetaI = stdmax(0.f, 1 - cosThetaI * cosThetaI);
  float sin2ThetaT = eta * eta * sin2ThetaI;
  if (sin2ThetaT >= 1) return false;
  float cosThetaT = sqrtf(1 - sin2ThetaT);
  *wt = eta * -wi + (eta * cosThetaI - cosThetaT) * n;
  return true;
}

// ---- sampling warps ----
PBR_D void concentric_sample_disk(float u0, float u1, float* dx, float* dy) {
  float ox = 2.f * u0 - 1, oy = 2.f * u1 - 1;
  if (ox == 0 && oy == 0) { *dx = 0; *dy = 0; return; }
  float theta, r;
  if (fabsf(ox) > fabsf(oy)) { r = ox; theta = kPiOver4 * (oy / ox); }
  else { r = oy; theta = kPiOver2 - kPiOver4 * (ox / oy); }
  *dx = r * cosf(theta);
  *dy = r * sinf(theta);
}
PBR_D V3 cosine_sample_hemisphere(float u0, float u1) {
  float dx, dy;
  concentric_sample_disk(u0, u1, &dx, &dy);
  float z = sqrtf(stdmax(0.f, 1 - dx * dx - dy * dy));
  return V3(dx, dy, z);
}
PBR_D float power_heuristic(int nf, float fPdf, int ng, float gPdf) {
  float f = nf * fPdf, g = ng * gPdf;
  return (f * f) / (f * f + g * g);
}

// ---- Fresnel ----
PBR_D float fr_dielectric(float cosThetaI, float etaI, float etaT) {
  cosThetaI = clampf(cosThetaI, -1, 1);
  bool entering = cosThetaI > 0.f;
  if (!entering) { float t = etaI; etaI = etaT; etaT = t; cosThetaI = fabsf(cosThetaI); }
  float sinThetaI = sqrtf(stdmax(0.f, 1 - cosThetaI * cosThetaI));
  float sinThetaT = etaI / etaT * sinThetaI;
  if (sinThetaT >= 1) return 1;
  float cosThetaT = sqrtf(stdmax(0.f, 1 - sinThetaT * sinThetaT));
  float Rparl = ((etaT * cosThetaI) - (etaI * cosThetaT)) / ((etaT * cosThetaI) + (etaI * cosThetaT));
  float Rperp = ((etaI * cosThetaI) - (etaT * cosThetaT)) / ((etaI * cosThetaI) + (etaT * cosThetaT));
  return (Rparl * Rparl + Rperp * Rperp) / 2;
}
PBR_D Spec fr_conductor(float cosThetaI, Spec etai, Spec etat, Spec k) {
  cosThetaI = clampf(cosThetaI, -1, 1);
  Spec eta = etat / etai;
  Spec etak = k / etai;
  float cosThetaI2 = cosThetaI * cosThetaI;
  float sinThetaI2 = 1.f - cosThetaI2;
  Spec eta2 = eta * eta;
  Spec etak2 = etak * etak;
  Spec t0 = eta2 - etak2 - Spec(sinThetaI2);
  Spec a2plusb2 = spec_sqrt(t0 * t0 + 4.f * eta2 * etak2);
  Spec t1 = a2plusb2 + Spec(cosThetaI2);
  Spec a = spec_sqrt(0.5f * (a2plusb2 + t0));
  Spec t2 = (2.f * cosThetaI) * a;
  Spec Rs = (t1 - t2) / (t1 + t2);
  Spec t3 = cosThetaI2 * a2plusb2 + Spec(sinThetaI2 * sinThetaI2);
  Spec t4 = t2 * sinThetaI2;
  Spec Rp = Rs * (t3 - t4) / (t3 + t4);
  return 0.5f * (Rp + Rs);
}
template <int MAT>
PBR_D Spec fresnel_eval(const Lobe& l, float cosI) {
  if (mat_has_fresnel<MAT>(FR_NOOP) && l.fresnel == FR_NOOP) return Spec(1.f);
  if (mat_has_fresnel<MAT>(FR_DIELECTRIC) && l.fresnel == FR_DIELECTRIC) return Spec(fr_dielectric(cosI, l.frEtaI, l.frEtaT));
  if (mat_has_fresnel<MAT>(FR_CONDUCTOR)) return fr_conductor(fabsf(cosI), Spec(1.f), l.cEta, l.T);
  return Spec(1.f);
}

// ---- Trowbridge-Reitz (GGX), sampleVisibleArea = true ----
PBR_D float tr_D(const Lobe& l, V3 wh) {
  float tan2Theta = tan2_theta(wh);
  if (isinf(tan2Theta)) return 0.f;
  const float cos4Theta = cos2_theta(wh) * cos2_theta(wh);
  float e = (cos2_phi(wh) / (l.alphax * l.alphax) + sin2_phi(wh) / (l.alphay * l.alphay)) * tan2Theta;
  return 1 / (kPi * l.alphax * l.alphay * cos4Theta * (1 + e) * (1 + e));
}
PBR_D float tr_lambda(const Lobe& l, V3 w) {
  float absTanTheta = fabsf(tan_theta(w));
  if (isinf(absTanTheta)) return 0.f;
  float alpha = sqrtf(cos2_phi(w) * l.alphax * l.alphax + sin2_phi(w) * l.alphay * l.alphay);
  float alpha2Tan2Theta = (alpha * absTanTheta) * (alpha * absTanTheta);
  return (-1 + sqrtf(1.f + alpha2Tan2Theta)) / 2;
}
PBR_D float tr_G1(const Lobe& l, V3 w) { return 1 / (1 + tr_lambda(l, w)); }
PBR_D float tr_G(const Lobe& l, V3 wo, V3 wi) { return 1 / (1 + tr_lambda(l, wo) + tr_lambda(l, wi)); }
PBR_D float tr_pdf(const Lobe& l, V3 wo, V3 wh) { return tr_D(l, wh) * tr_G1(l, wo) * absdot(wo, wh) / abs_cos_theta(wo); }
PBR_D void tr_sample11(float cosTheta, float U1, float U2, float* slope_x, float* slope_y) {
  if ((double)cosTheta > .9999) {
    float r = sqrtf(U1 / (1 - U1));
    float phi = (float)(6.28318530718 * (double)U2);
    *slope_x = (float)((double)r * cos((double)phi));
    *slope_y = (float)((double)r * sin((double)phi));
    return;
  }
  float sinTheta = sqrtf(stdmax(0.f, 1.f - cosTheta * cosTheta));
  float tanTheta = sinTheta / cosTheta;
  float a = 1 / tanTheta;
  float G1 = 2 / (1 + sqrtf(1.f + 1.f / (a * a)));
  float A = 2 * U1 / G1 - 1;
  float tmp = 1.f / (A * A - 1.f);
  if (tmp > 1e10f) tmp = 1e10f;
  float B = tanTheta;
  float D = sqrtf(stdmax(B * B * tmp * tmp - (A * A - B * B) * tmp, 0.f));
  float slope_x_1 = B * tmp - D;
  float slope_x_2 = B * tmp + D;
  *slope_x = (A < 0 || slope_x_2 > 1.f / tanTheta) ? slope_x_1 : slope_x_2;
  float S;
  if (U2 > 0.5f) { S = 1.f; U2 = 2.f * (U2 - .5f); }
  else { S = -1.f; U2 = 2.f * (.5f - U2); }
  float z = (U2 * (U2 * (U2 * 0.27385f - 0.73369f) + 0.46341f)) /
            (U2 * (U2 * (U2 * 0.093073f + 0.309420f) - 1.000000f) + 0.597999f);
  *slope_y = S * z * sqrtf(1.f + *slope_x * *slope_x);
}
PBR_D V3 tr_sample(V3 wi, float alpha_x, float alpha_y, float U1, float U2) {
  V3 wiStretched = normalize(V3(alpha_x * wi.x, alpha_y * wi.y, wi.z));
  float slope_x, slope_y;
  tr_sample11(cos_theta(wiStretched), U1, U2, &slope_x, &slope_y);
  float tmp = cos_phi(wiStretched) * slope_x - sin_phi(wiStretched) * slope_y;
  slope_y = sin_phi(wiStretched) * slope_x + cos_phi(wiStretched) * slope_y;
  slope_x = tmp;
  slope_x = alpha_x * slope_x;
  slope_y = alpha_y * slope_y;
  return normalize(V3(-slope_x, -slope_y, 1.f));
}
PBR_D V3 tr_sample_wh(const Lobe& l, V3 wo, float u0, float u1) {
  bool flip = wo.z < 0;
  V3 wh = tr_sample(flip ? -wo : wo, l.alphax, l.alphay, u0, u1);
  if (flip) wh = -wh;
  return wh;
}

// ---- BxDF::f ----
template <int MAT>
PBR_D Spec lobe_f(const Lobe& l, V3 wo, V3 wi) {
  switch (l.kind) {
    case LOBE_LAMBERT:
      if (!mat_has_lobe<MAT>(LOBE_LAMBERT)) break;
      return l.R * kInvPi;
    case LOBE_OREN_NAYAR: {
      if (!mat_has_lobe<MAT>(LOBE_OREN_NAYAR)) break;
      float sinThetaI = sin_theta(wi), sinThetaO = sin_theta(wo);
      float maxCos = 0;
      if (sinThetaI > 1e-4 && sinThetaO > 1e-4) {
        float sinPhiI = sin_phi(wi), cosPhiI = cos_phi(wi);
        float sinPhiO = sin_phi(wo), cosPhiO = cos_phi(wo);
        float dCos = cosPhiI * cosPhiO + sinPhiI * sinPhiO;
        maxCos = stdmax(0.f, dCos);
      }
      float sinAlpha, tanBeta;
      if (abs_cos_theta(wi) > abs_cos_theta(wo)) { sinAlpha = sinThetaO; tanBeta = sinThetaI / abs_cos_theta(wi); }
      else { sinAlpha = sinThetaI; tanBeta = sinThetaO / abs_cos_theta(wo); }
      return l.R * kInvPi * (l.A + l.B * maxCos * sinAlpha * tanBeta);
    }
    case LOBE_MICRO_REFL: {
      if (!mat_has_lobe<MAT>(LOBE_MICRO_REFL)) break;
      float cosThetaO = abs_cos_theta(wo), cosThetaI = abs_cos_theta(wi);
      V3 wh = wi + wo;
      if (cosThetaI == 0 || cosThetaO == 0) return Spec(0.f);
      if (wh.x == 0 && wh.y == 0 && wh.z == 0) return Spec(0.f);
      wh = normalize(wh);
      Spec F = fresnel_eval<MAT>(l, dot(wi, faceforward(wh, V3(0, 0, 1))));
      return l.R * tr_D(l, wh) * tr_G(l, wo, wi) * F / (4 * cosThetaI * cosThetaO);
    }
    case LOBE_MICRO_TRANS: {
      if (!mat_has_lobe<MAT>(LOBE_MICRO_TRANS)) break;
      if (same_hemisphere(wo, wi)) return Spec(0.f);
      float cosThetaO = cos_theta(wo), cosThetaI = cos_theta(wi);
      if (cosThetaI == 0 || cosThetaO == 0) return Spec(0.f);
      float eta = cos_theta(wo) > 0 ? (l.etaB / l.etaA) : (l.etaA / l.etaB);
      V3 wh = normalize(wo + wi * eta);
      if (wh.z < 0) wh = -wh;
      if (dot(wo, wh) * dot(wi, wh) > 0) return Spec(0.f);
      Spec F = Spec(fr_dielectric(dot(wo, wh), l.etaA, l.etaB));
      float sqrtDenom = dot(wo, wh) + eta * dot(wi, wh);
      float factor = 1 / eta;  // TransportMode::Radiance
      return (Spec(1.f) - F) * l.R *
             fabsf(tr_D(l, wh) * tr_G(l, wo, wi) * eta * eta * absdot(wi, wh) * absdot(wo, wh) * factor * factor /
                   (cosThetaI * cosThetaO * sqrtDenom * sqrtDenom));
    }
    default: break;  // specular lobes
  }
  return Spec(0.f);
}

// ---- BxDF::Pdf ----
template <int MAT>
PBR_D float lobe_pdf(const Lobe& l, V3 wo, V3 wi) {
  switch (l.kind) {
    case LOBE_LAMBERT:
    case LOBE_OREN_NAYAR: return same_hemisphere(wo, wi) ? abs_cos_theta(wi) * kInvPi : 0;
    case LOBE_MICRO_REFL: {
      if (!mat_has_lobe<MAT>(LOBE_MICRO_REFL)) break;
      if (!same_hemisphere(wo, wi)) return 0;
      V3 wh = normalize(wo + wi);
      return tr_pdf(l, wo, wh) / (4 * dot(wo, wh));
    }
    case LOBE_MICRO_TRANS: {
      if (!mat_has_lobe<MAT>(LOBE_MICRO_TRANS)) break;
      if (same_hemisphere(wo, wi)) return 0;
      float eta = cos_theta(wo) > 0 ? (l.etaB / l.etaA) : (l.etaA / l.etaB);
      V3 wh = normalize(wo + wi * eta);
      if (dot(wo, wh) * dot(wi, wh) > 0) return 0;
      float sqrtDenom = dot(wo, wh) + eta * dot(wi, wh);
      float dwh_dwi = fabsf((eta * eta * dot(wi, wh)) / (sqrtDenom * sqrtDenom));
      return tr_pdf(l, wo, wh) * dwh_dwi;
    }
    default: break;
  }
  return 0;
}

// ---- BxDF::Sample_f.  *pdf is pre-set to 0 by the caller (BSDF::Sample_f, Reflection.cpp:132) and left
// untouched on the early-out paths, as in the reference. ----
template <int MAT>
PBR_D Spec lobe_sample_f(const Lobe& l, V3 wo, V3* wi, float u0, float u1, float* pdf, int* sampledType) {
  switch (l.kind) {
    case LOBE_LAMBERT:
    case LOBE_OREN_NAYAR: {
      if (!mat_has_lobe<MAT>(LOBE_LAMBERT)) break;
      *wi = cosine_sample_hemisphere(u0, u1);
      if (wo.z < 0) wi->z *= -1;
      *pdf = lobe_pdf<MAT>(l, wo, *wi);
      return lobe_f<MAT>(l, wo, *wi);
    }
    case LOBE_SPEC_REFL: {
      if (!mat_has_lobe<MAT>(LOBE_SPEC_REFL)) break;
      *wi = V3(-wo.x, -wo.y, wo.z);
      *pdf = 1;
      return fresnel_eval<MAT>(l, cos_theta(*wi)) * l.R / abs_cos_theta(*wi);
    }
    case LOBE_SPEC_TRANS: {
      if (!mat_has_lobe<MAT>(LOBE_SPEC_TRANS)) break;
      bool entering = cos_theta(wo) > 0;
      float etaI = entering ? l.etaA : l.etaB;
      float etaT = entering ? l.etaB : l.etaA;
      if (!refract_dir(wo, faceforward(V3(0, 0, 1), wo), etaI / etaT, wi)) return Spec(0.f);
      *pdf = 1;
      Spec ft = l.R * (Spec(1.f) - Spec(fr_dielectric(cos_theta(*wi), l.etaA, l.etaB)));
      ft = ft * ((etaI * etaI) / (etaT * etaT));
      return ft / abs_cos_theta(*wi);
    }
    case LOBE_FRESNEL_SPEC: {
      if (!mat_has_lobe<MAT>(LOBE_FRESNEL_SPEC)) break;
      float F = fr_dielectric(cos_theta(wo), l.etaA, l.etaB);
      if (u0 < F) {
        *wi = V3(-wo.x, -wo.y, wo.z);
        *sampledType = BSDF_SPECULAR | BSDF_REFLECTION;
        *pdf = F;
        return F * l.R / abs_cos_theta(*wi);
      } else {
        bool entering = cos_theta(wo) > 0;
        float etaI = entering ? l.etaA : l.etaB;
        float etaT = entering ? l.etaB : l.etaA;
        if (!refract_dir(wo, faceforward(V3(0, 0, 1), wo), etaI / etaT, wi)) return Spec(0.f);
        Spec ft = l.T * (1 - F);
        ft = ft * ((etaI * etaI) / (etaT * etaT));
        *sampledType = BSDF_SPECULAR | BSDF_TRANSMISSION;
        *pdf = 1 - F;
        return ft / abs_cos_theta(*wi);
      }
    }
    case LOBE_MICRO_REFL: {
      if (!mat_has_lobe<MAT>(LOBE_MICRO_REFL)) break;
      if (wo.z == 0) return Spec(0.f);
      V3 wh = tr_sample_wh(l, wo, u0, u1);
      if (dot(wo, wh) < 0) return Spec(0.f);
      *wi = reflect_about(wo, wh);
      if (!same_hemisphere(wo, *wi)) return Spec(0.f);
      *pdf = tr_pdf(l, wo, wh) / (4 * dot(wo, wh));
      return lobe_f<MAT>(l, wo, *wi);
    }
    case LOBE_MICRO_TRANS: {
      if (!mat_has_lobe<MAT>(LOBE_MICRO_TRANS)) break;
      if (wo.z == 0) return Spec(0.f);
      V3 wh = tr_sample_wh(l, wo, u0, u1);
      if (dot(wo, wh) < 0) return Spec(0.f);
      float eta = cos_theta(wo) > 0 ? (l.etaA / l.etaB) : (l.etaB / l.etaA);
      if (!refract_dir(wo, wh, eta, wi)) return Spec(0.f);
      *pdf = lobe_pdf<MAT>(l, wo, *wi);
      return lobe_f<MAT>(l, wo, *wi);
    }
  }
  return Spec(0.f);
}

// ---- BSDF::f / Pdf / Sample_f (Reflection.cpp:56-164) ----
// The lobe loops run to a compile-time bound (the most lobes the material can create) with an `i < b.n`
// guard, so that lobes[] is indexed statically and the whole Bsdf stays in registers.
template <int MAT>
PBR_D constexpr int mat_max_lobes() { return (MAT == 0 || MAT == 1 || MAT == 3) ? 1 : 2; }

template <int MAT>
PBR_D int bsdf_num_components(const Bsdf& b, int flags) {
  int c = 0;
#pragma unroll
  for (int i = 0; i < mat_max_lobes<MAT>(); ++i)
    if (i < b.n && (b.lobes[i].flags & flags) == b.lobes[i].flags) ++c;
  return c;
}
template <int MAT>
PBR_D Spec bsdf_f(const Bsdf& b, V3 woW, V3 wiW, int flags) {
  V3 wi = b.to_local(wiW), wo = b.to_local(woW);
  if (wo.z == 0) return Spec(0.f);
  bool reflect = dot(wiW, b.ng) * dot(woW, b.ng) > 0;
  Spec f(0.f);
#pragma unroll
  for (int i = 0; i < mat_max_lobes<MAT>(); ++i) {
    if (i >= b.n) continue;
    const Lobe& l = b.lobes[i];
    if ((l.flags & flags) == l.flags &&
        ((reflect && (l.flags & BSDF_REFLECTION)) || (!reflect && (l.flags & BSDF_TRANSMISSION))))
      f = f + lobe_f<MAT>(l, wo, wi);
  }
  return f;
}
template <int MAT>
PBR_D float bsdf_pdf(const Bsdf& b, V3 woW, V3 wiW, int flags) {
  if (b.n == 0) return 0.f;
  V3 wo = b.to_local(woW), wi = b.to_local(wiW);
  if (wo.z == 0) return 0.f;
  float pdf = 0.f;
  int matching = 0;
#pragma unroll
  for (int i = 0; i < mat_max_lobes<MAT>(); ++i)
    if (i < b.n && (b.lobes[i].flags & flags) == b.lobes[i].flags) { ++matching; pdf += lobe_pdf<MAT>(b.lobes[i], wo, wi); }
  return matching > 0 ? pdf / matching : 0.f;
}
template <int MAT>
PBR_D Spec bsdf_sample_f(const Bsdf& b, V3 woW, V3* wiW, float u0, float u1, float* pdf, int type, int* sampledType) {
  int matching = bsdf_num_components<MAT>(b, type);
  if (matching == 0) { *pdf = 0; *sampledType = 0; return Spec(0.f); }
  int comp = min((int)floorf(u0 * matching), matching - 1);
  int which = 0, count = comp;
  bool chosen = false;
#pragma unroll
  for (int i = 0; i < mat_max_lobes<MAT>(); ++i)
    if (!chosen && i < b.n && (b.lobes[i].flags & type) == b.lobes[i].flags && count-- == 0) { which = i; chosen = true; }
  const Lobe l = (mat_max_lobes<MAT>() == 1 || which == 0) ? b.lobes[0] : b.lobes[mat_max_lobes<MAT>() - 1];
  float ur0 = stdmin(u0 * matching - comp, kOneMinusEpsilon);
  V3 wi, wo = b.to_local(woW);
  // reference returns 0 here WITHOUT writing *pdf (Reflection.cpp:131); callers then read an
  // uninitialised pdf.  We define it as 0 (documented deviation, measure-zero event).
  if (wo.z == 0) { *pdf = 0; *sampledType = 0; return Spec(0.f); }
  *pdf = 0;
  *sampledType = l.flags;
  Spec f = lobe_sample_f<MAT>(l, wo, &wi, ur0, u1, pdf, sampledType);
  if (*pdf == 0) { *sampledType = 0; return Spec(0.f); }
  *wiW = b.to_world(wi);
  if (!(l.flags & BSDF_SPECULAR) && matching > 1) {
#pragma unroll
    for (int i = 0; i < mat_max_lobes<MAT>(); ++i)
      if (i < b.n && i != which && (b.lobes[i].flags & type) == b.lobes[i].flags) *pdf += lobe_pdf<MAT>(b.lobes[i], wo, wi);
  }
  if (matching > 1) *pdf /= matching;
  if (!(l.flags & BSDF_SPECULAR)) {
    bool reflect = dot(*wiW, b.ng) * dot(woW, b.ng) > 0;
    f = Spec(0.f);
#pragma unroll
    for (int i = 0; i < mat_max_lobes<MAT>(); ++i) {
      if (i >= b.n) continue;
      const Lobe& li = b.lobes[i];
      if ((li.flags & type) == li.flags &&
          ((reflect && (li.flags & BSDF_REFLECTION)) || (!reflect && (li.flags & BSDF_TRANSMISSION))))
        f = f + lobe_f<MAT>(li, wo, wi);
    }
  }
  return f;
}

}  // namespace pbr
