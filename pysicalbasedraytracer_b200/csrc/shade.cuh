// shade.cuh - surface interaction reconstruction, material -> BSDF, light sampling, MIS.
//   Triangle::Intersect attribute part      Shape/Triangle.cpp:147-254
//   SurfaceInteraction ctor                 Core/Interaction.cpp:8-32
//   *Material::ComputeScatteringFunctions   Material/{MatteMaterial,Mirror,GlassMaterial,MetalMaterial,PlasticMaterial}.cpp
//   BSDF ctor (shading frame)               Material/Reflection.h:105-110
//   Light::Sample_Li / Pdf_Li / L / Le      Light/{PointLight,DiffuseLight,InfiniteAreaLight,SkyBoxLight}.cpp
//   Shape::Sample / Pdf (solid angle)       Shape/Shape.cpp:18-42,  Triangle::Sample Shape/Triangle.cpp:360-387
//   Distribution1D / 2D                     Sampler/Sampling.h:77-171
//   Interaction::SpawnRay / SpawnRayTo      Core/Interaction.h:28-44
#pragma once
#include "bsdf.cuh"
#include "device_types.cuh"
#include "trace.cuh"
#include "texture.cuh"

namespace pbr {

struct Isect {
  V3 p, pError, n, ns;  // n = geometric normal (after orientation flip), ns = shading normal
  V3 wo;                // Normalize(-ray.d)  (Interaction ctor normalises)
  V3 ss, ts;            // BSDF shading frame (host-precomputed per triangle)
  int matId, matType, lightId;
};

struct TriData {
  V3 p0, p1, p2;
  uint32_t w0, w1, w2;
};
PBR_D TriData load_tri(const SceneDev& sc, int ti) {
  const float4* rec = sc.tv + 3 * (size_t)ti;
  float4 a = __ldg(rec), b = __ldg(rec + 1), c = __ldg(rec + 2);
  TriData t;
  t.p0 = V3(a.x, a.y, a.z); t.p1 = V3(b.x, b.y, b.z); t.p2 = V3(c.x, c.y, c.z);
  t.w0 = __float_as_uint(a.w); t.w1 = __float_as_uint(b.w); t.w2 = __float_as_uint(c.w);
  return t;
}

// dpdu of Triangle::Intersect (Triangle.cpp:147-160) for the given vertex uvs; zero when the uvs are degenerate
// (the reference leaves dpdu = dpdv = 0 there)
PBR_HD V3 triangle_dpdu(V3 p0, V3 p1, V3 p2, float u0x, float u0y, float u1x, float u1y, float u2x, float u2y) {
  V3 dp02 = p0 - p2, dp12 = p1 - p2;
  const float duv02x = u0x - u2x, duv02y = u0y - u2y, duv12x = u1x - u2x, duv12y = u1y - u2y;
  float determinant = duv02x * duv12y - duv02y * duv12x;
  V3 dpdu;
  if (!(fabsf(determinant) < 1e-8)) {
    float invdet = 1 / determinant;
    dpdu = (duv12y * dp02 - duv02y * dp12) * invdet;
  }
  return dpdu;
}

// dpdu AND dpdv (Triangle.cpp:147-160), needed by SurfaceInteraction::ComputeDifferentials for image textures
PBR_D void triangle_dpdu_dpdv(V3 p0, V3 p1, V3 p2, const float uv[6], V3* dpdu, V3* dpdv) {
  V3 dp02 = p0 - p2, dp12 = p1 - p2;
  const float duv02x = uv[0] - uv[4], duv02y = uv[1] - uv[5], duv12x = uv[2] - uv[4], duv12y = uv[3] - uv[5];
  float determinant = duv02x * duv12y - duv02y * duv12x;
  *dpdu = V3(0, 0, 0); *dpdv = V3(0, 0, 0);
  if (!(fabsf(determinant) < 1e-8)) {
    float invdet = 1 / determinant;
    *dpdu = (duv12y * dp02 - duv02y * dp12) * invdet;
    *dpdv = (-duv12x * dp02 + duv02x * dp12) * invdet;
  }
}

// Host + device: the per-triangle frame of a FLAT-shaded triangle exactly as the reference computes it per hit
//   n  = Normalize(Cross(p0 - p2, p1 - p2)), flipped by reverseOrientation ^ transformSwapsHandedness  (Triangle.cpp:152,182-184)
//   dpdu from the vertex uvs (default (0,0),(1,0),(1,1), Triangle.h:50-54)                             (Triangle.cpp:147-160)
//   ss = Normalize(dpdu), ts = Cross(ns, ss)                                                          (Reflection.h:105-110)
PBR_HD void triangle_frame(V3 p0, V3 p1, V3 p2, const float uv[6], bool flip, V3* n, V3* ss, V3* ts) {
  V3 dp02 = p0 - p2, dp12 = p1 - p2;
  V3 dpdu = triangle_dpdu(p0, p1, p2, uv[0], uv[1], uv[2], uv[3], uv[4], uv[5]);
  V3 nn = normalize(cross(dp02, dp12));
  if (flip) nn = -nn;
  *n = nn;
  *ss = normalize(dpdu);
  *ts = cross(nn, *ss);
}

// SurfaceInteraction of a hit on an analytic sphere (host/AnalyticSphere.h::Intersect); kept out of line: scenes without
// spheres never call it and the shading kernels stay as compact as before.
// (Everything goes in and comes out BY VALUE: a pointer or reference to the caller's Isect / TriData handed to an out-of-line
// function forces those structs into local memory in every shading kernel - 185 local loads / stores per shaded hit.)
static __device__ __noinline__ Isect make_isect_sphere(V3 center, float radius, uint32_t w1, uint32_t w2, V3 rayO, V3 rayD, float tHit) {
  Isect is;
  const uint32_t fl = w1 >> kTriFlagShift;
  const SphereHit h = sphere_hit(center, radius, (fl & kTriFlagFlipNormal) != 0, rayO, rayD, tHit);
  is.p = h.p;
  is.pError = h.pError;
  is.n = h.n;
  is.ns = h.n;
  is.ss = normalize(h.dpdu);   // BSDF ctor (Reflection.h:105-110): ss = Normalize(shading.dpdu), ts = Cross(ns, ss)
  is.ts = cross(h.n, is.ss);
  is.wo = normalize(-rayD);
  is.matId = (int)(w1 & kTriMatMask);
  is.matType = (int)((w1 >> kTriTypeShift) & 0xf);
  is.lightId = (int)w2;
  return is;
}

// rayO: origin of the ray that produced the hit; read for analytic spheres only (their hit record is the ray parameter in b0)
PBR_D void make_isect(const SceneDev& sc, const TriData& t, int tri, float b0, float b1, float b2, V3 rayD, Isect* is,
                      V3 rayO = V3(0.f, 0.f, 0.f)) {
  if (t.w1 & kPrimSphereBit) { const Isect sph = make_isect_sphere(t.p0, t.p1.x, t.w1, t.w2, rayO, rayD, b0); *is = sph; return; }
  const V3 p0 = t.p0, p1 = t.p1, p2 = t.p2;
  float xAbsSum = (fabsf(b0 * p0.x) + fabsf(b1 * p1.x) + fabsf(b2 * p2.x));
  float yAbsSum = (fabsf(b0 * p0.y) + fabsf(b1 * p1.y) + fabsf(b2 * p2.y));
  float zAbsSum = (fabsf(b0 * p0.z) + fabsf(b1 * p1.z) + fabsf(b2 * p2.z));
  is->pError = gamma_f(7) * V3(xAbsSum, yAbsSum, zAbsSum);
  is->p = b0 * p0 + b1 * p1 + b2 * p2;
  const float4* frame = sc.tf + 3 * (size_t)tri;
  const float4 n4 = __ldg(frame);
  const uint32_t fl = t.w1 >> kTriFlagShift;
  if (!(fl & kTriFlagNormals) && !(t.w1 & kTriTangentBit)) {
    // flat-shaded triangle: the whole frame is a per-triangle constant (host-precomputed, uvs included)
    const float4 s4 = __ldg(frame + 1), t4 = __ldg(frame + 2);
    is->n = V3(n4.x, n4.y, n4.z);
    is->ns = is->n;
    is->ss = V3(s4.x, s4.y, s4.z);
    is->ts = V3(t4.x, t4.y, t4.z);
  } else {
    // per-vertex shading normals and / or tangents (`mesh->n || mesh->s`): Triangle.cpp:186-250 +
    // SurfaceInteraction::SetShadingGeometry (Interaction.cpp:98-113)
    float uv[6] = {0.f, 0.f, 1.f, 0.f, 1.f, 1.f};
    if (fl & kTriFlagUVs) {
      const float4 a = __ldg(sc.tuv0 + tri), b = __ldg(sc.tuv1 + tri);
      uv[0] = a.x; uv[1] = a.y; uv[2] = a.z; uv[3] = a.w; uv[4] = b.x; uv[5] = b.y;
    }
    V3 n(n4.x, n4.y, n4.z);
    const V3 dpdu = triangle_dpdu(p0, p1, p2, uv[0], uv[1], uv[2], uv[3], uv[4], uv[5]);
    V3 ns = n;
    if (fl & kTriFlagNormals) {
      const float4 a0 = __ldg(sc.tn0 + tri), a1 = __ldg(sc.tn1 + tri), a2 = __ldg(sc.tn2 + tri);
      ns = b0 * V3(a0.x, a0.y, a0.z) + b1 * V3(a1.x, a1.y, a1.z) + b2 * V3(a2.x, a2.y, a2.z);
      if (length_sq(ns) > 0) ns = normalize(ns);
      else ns = n;
    }
    V3 ss;
    if (t.w1 & kTriTangentBit) {  // mesh->s: interpolated tangent, Normalize(dpdu) when it vanishes (Triangle.cpp:204-213)
      const float4 s0 = __ldg(sc.ts0 + tri), s1 = __ldg(sc.ts1 + tri), s2 = __ldg(sc.ts2 + tri);
      ss = b0 * V3(s0.x, s0.y, s0.z) + b1 * V3(s1.x, s1.y, s1.z) + b2 * V3(s2.x, s2.y, s2.z);
      if (length_sq(ss) > 0) ss = normalize(ss);
      else ss = normalize(dpdu);
    } else {
      ss = normalize(dpdu);
    }
    V3 ts = cross(ss, ns);
    if (length_sq(ts) > 0.f) {
      ts = normalize(ts);
      ss = cross(ts, ns);
    } else {
      coordinate_system(ns, &ss, &ts);
    }
    if (fl & kTriFlagReverse) ts = -ts;
    // SetShadingGeometry(ss, ts, .., orientationIsAuthoritative = true)
    V3 sn = normalize(cross(ss, ts));
    n = faceforward(n, sn);
    is->n = n;
    is->ns = sn;
    // BSDF ctor (Reflection.h:105-110): ss = Normalize(shading.dpdu), ts = Cross(ns, ss)
    is->ss = normalize(ss);
    is->ts = cross(sn, is->ss);
  }
  is->wo = normalize(-rayD);
  is->matId = (int)(t.w1 & kTriMatMask);
  is->matType = (int)((t.w1 >> kTriTypeShift) & 0xf);
  is->lightId = (int)t.w2;
}

// DiffuseAreaLight::L (DiffuseLight.h:17-19)
PBR_D Spec area_light_L(const pbrgpu_light& l, V3 n, V3 w) {
  return (l.two_sided || dot(n, w) > 0) ? Spec(l.L[0], l.L[1], l.L[2]) : Spec(0.f);
}

PBR_D Lobe make_lobe(int kind, int flags, Spec R) {
  Lobe l;
  l.kind = kind; l.flags = flags; l.fresnel = FR_NOOP; l.R = R; l.T = Spec(0.f); l.cEta = Spec(0.f);
  l.frEtaI = l.frEtaT = 1.f; l.etaA = l.etaB = 1.f; l.alphax = l.alphay = 0.001f; l.A = l.B = 0.f;
  return l;
}

// Material::ComputeScatteringFunctions for the five materials.  `mats[...]` roughness values were
// already passed through RoughnessToAlpha on the host when remap_roughness was set (Microfacet.h:78-83).
// kdTex / ksTex / roughTex: Kd->Evaluate(si) (Matte / Plastic), Ks->Evaluate(si) and roughness->Evaluate(si) (Plastic) of an
// ImageTexture, or NULL for the ConstantTexture in the material record.
template <int MAT>
PBR_D void make_bsdf(const SceneDev& sc, const Isect& is, bool allowMultipleLobes, Bsdf* b, const Spec* kdTex = nullptr,
                     const Spec* ksTex = nullptr, const float* roughTex = nullptr) {
  const pbrgpu_material& m = sc.mats[is.matId];
  b->eta = 1.f;
  b->ns = is.ns;
  b->ng = is.n;
  b->ss = is.ss;
  b->ts = is.ts;
  b->n = 0;
  switch (MAT) {  // == m.type: the shading queues are sorted by material type
    case PBRGPU_MAT_MATTE: {
      Spec r = spec_clamp0(kdTex ? *kdTex : Spec(m.kd[0], m.kd[1], m.kd[2]));
      float sig = clampf(m.sigma, 0.f, 90.f);
      if (!r.is_black()) {
        if (sig == 0) b->lobes[b->n++] = make_lobe(LOBE_LAMBERT, BSDF_REFLECTION | BSDF_DIFFUSE, r);
        else {
          Lobe l = make_lobe(LOBE_OREN_NAYAR, BSDF_REFLECTION | BSDF_DIFFUSE, r);
          float sigma = (kPi / 180) * sig;
          float sigma2 = sigma * sigma;
          l.A = 1.f - (sigma2 / (2.f * (sigma2 + 0.33f)));
          l.B = 0.45f * sigma2 / (sigma2 + 0.09f);
          b->lobes[b->n++] = l;
        }
      }
      break;
    }
    case PBRGPU_MAT_MIRROR: {
      Spec R = spec_clamp0(Spec(m.kr[0], m.kr[1], m.kr[2]));
      if (!R.is_black()) b->lobes[b->n++] = make_lobe(LOBE_SPEC_REFL, BSDF_REFLECTION | BSDF_SPECULAR, R);
      break;
    }
    case PBRGPU_MAT_GLASS: {
      float eta = m.eta[0];
      float urough = m.u_rough, vrough = m.v_rough;  // alpha if remapped
      Spec R = spec_clamp0(Spec(m.kr[0], m.kr[1], m.kr[2]));
      Spec T = spec_clamp0(Spec(m.kt[0], m.kt[1], m.kt[2]));
      b->eta = eta;
      if (R.is_black() && T.is_black()) break;
      bool isSpecular = (m.pad3 != 0.f);  // host: original urough == 0 && vrough == 0 (before remapping)
      if (isSpecular && allowMultipleLobes) {
        Lobe l = make_lobe(LOBE_FRESNEL_SPEC, BSDF_REFLECTION | BSDF_TRANSMISSION | BSDF_SPECULAR, R);
        l.T = T; l.etaA = 1.f; l.etaB = eta;
        b->lobes[b->n++] = l;
      } else {
        if (!R.is_black()) {
          Lobe l = make_lobe(isSpecular ? LOBE_SPEC_REFL : LOBE_MICRO_REFL,
                             isSpecular ? (BSDF_REFLECTION | BSDF_SPECULAR) : (BSDF_REFLECTION | BSDF_GLOSSY), R);
          l.fresnel = FR_DIELECTRIC; l.frEtaI = 1.f; l.frEtaT = eta;
          l.alphax = stdmax(0.001f, urough); l.alphay = stdmax(0.001f, vrough);
          b->lobes[b->n++] = l;
        }
        if (!T.is_black()) {
          Lobe l = make_lobe(isSpecular ? LOBE_SPEC_TRANS : LOBE_MICRO_TRANS,
                             isSpecular ? (BSDF_TRANSMISSION | BSDF_SPECULAR) : (BSDF_TRANSMISSION | BSDF_GLOSSY), T);
          l.etaA = 1.f; l.etaB = eta;
          l.alphax = stdmax(0.001f, urough); l.alphay = stdmax(0.001f, vrough);
          b->lobes[b->n++] = l;
        }
      }
      break;
    }
    case PBRGPU_MAT_METAL: {
      Lobe l = make_lobe(LOBE_MICRO_REFL, BSDF_REFLECTION | BSDF_GLOSSY, Spec(1.f));
      l.fresnel = FR_CONDUCTOR;
      l.cEta = Spec(m.eta[0], m.eta[1], m.eta[2]);
      l.T = Spec(m.k[0], m.k[1], m.k[2]);
      l.alphax = stdmax(0.001f, m.u_rough); l.alphay = stdmax(0.001f, m.v_rough);
      b->lobes[b->n++] = l;
      break;
    }
    case PBRGPU_MAT_PLASTIC: {
      Spec kd = spec_clamp0(kdTex ? *kdTex : Spec(m.kd[0], m.kd[1], m.kd[2]));
      if (!kd.is_black()) b->lobes[b->n++] = make_lobe(LOBE_LAMBERT, BSDF_REFLECTION | BSDF_DIFFUSE, kd);
      Spec ks = spec_clamp0(ksTex ? *ksTex : Spec(m.kr[0], m.kr[1], m.kr[2]));
      if (!ks.is_black()) {
        Lobe l = make_lobe(LOBE_MICRO_REFL, BSDF_REFLECTION | BSDF_GLOSSY, ks);
        l.fresnel = FR_DIELECTRIC; l.frEtaI = 1.5f; l.frEtaT = 1.f;
        const float rough = roughTex ? *roughTex : m.u_rough;  // roughness->Evaluate(*si) (PlasticMaterial.cpp:21; remap on the host)
        l.alphax = stdmax(0.001f, rough); l.alphay = stdmax(0.001f, rough);
        b->lobes[b->n++] = l;
      }
      break;
    }
  }
}

// FindInterval (Core/PBR.h:167-181) over a cdf array
PBR_D int find_interval(const float* cdf, int size, float u) {
  int first = 0, len = size;
  while (len > 0) {
    int half = len >> 1, middle = first + half;
    if (__ldg(cdf + middle) <= u) { first = middle + 1; len -= half + 1; }
    else len = half;
  }
  int r = first - 1;
  return r < 0 ? 0 : (r > size - 2 ? size - 2 : r);
}

// Distribution1D::SampleDiscrete (Sampling.h:96-107)
PBR_D int sample_discrete(const float* func, const float* cdf, float funcInt, int count, float u, float* pdf) {
  int offset = find_interval(cdf, count + 1, u);
  *pdf = (funcInt > 0) ? __ldg(func + offset) / (funcInt * count) : 0;
  return offset;
}
// Distribution1D::SampleContinuous (Sampling.h:118-131)
PBR_D float sample_continuous(const float* func, const float* cdf, float funcInt, int count, float u, float* pdf, int* off) {
  int offset = find_interval(cdf, count + 1, u);
  *off = offset;
  float c0 = __ldg(cdf + offset), c1 = __ldg(cdf + offset + 1);
  float du = u - c0;
  if ((c1 - c0) > 0) du /= (c1 - c0);
  *pdf = (funcInt > 0) ? __ldg(func + offset) / funcInt : 0;
  return (offset + du) / count;
}

// MIPMap::Lookup(st, 0.f) == bilinear `triangle(0, st)` on level 0 with Repeat wrap
// (Texture/MIPMap.h:207-213, 256-266, Texel :183-205)
PBR_D Spec env_texel(const EnvDev& e, int s, int t) {
  s = s % e.width; if (s < 0) s += e.width;   // Mod()
  t = t % e.height; if (t < 0) t += e.height;
  const float* p = e.texels + 3 * ((size_t)t * e.width + s);
  return Spec(__ldg(p), __ldg(p + 1), __ldg(p + 2));
}
PBR_D Spec env_lookup_bilinear(const EnvDev& e, float s_, float t_) {
  float s = s_ * e.width - 0.5f;
  float t = t_ * e.height - 0.5f;
  int s0 = (int)floorf(s), t0 = (int)floorf(t);
  float ds = s - s0, dt = t - t0;
  return (1 - ds) * (1 - dt) * env_texel(e, s0, t0) + (1 - ds) * dt * env_texel(e, s0, t0 + 1) +
         ds * (1 - dt) * env_texel(e, s0 + 1, t0) + ds * dt * env_texel(e, s0 + 1, t0 + 1);
}
PBR_D float spherical_theta(V3 v) { return acosf(clampf(v.z, -1, 1)); }
PBR_D float spherical_phi(V3 v) { float p = atan2f(v.y, v.x); return (p < 0) ? (p + 2 * kPi) : p; }
PBR_D V3 xform_vec(const float* m, V3 v) {
  return V3(m[0] * v.x + m[1] * v.y + m[2] * v.z, m[4] * v.x + m[5] * v.y + m[6] * v.z, m[8] * v.x + m[9] * v.y + m[10] * v.z);
}
// SkyBoxLight::getLightValue (SkyBoxLight.cpp:30-43) incl. HDRtoLDR(…, 0.3) (Spectrum.h:219-226)
PBR_D Spec skybox_value(const EnvDev& e, float u, float v) {
  u = clampf(u, 0.f, 1.f);
  v = clampf(v, 0.f, 1.f);
  int w = (int)(u * e.width), h = (int)(v * e.height);
  w = w < 0 ? 0 : (w > e.width - 1 ? e.width - 1 : w);
  h = h < 0 ? 0 : (h > e.height - 1 ? e.height - 1 : h);
  const float* p = e.texels + 3 * ((size_t)w + (size_t)h * e.width);
  float invExposure = (float)(1.0 / (1.0 - (double)0.3f));
  return Spec((float)(1.0 - (double)expf(-__ldg(p) * invExposure)), (float)(1.0 - (double)expf(-__ldg(p + 1) * invExposure)),
              (float)(1.0 - (double)expf(-__ldg(p + 2) * invExposure)));
}
PBR_D void skybox_uv(V3 p, float* u, float* v) {  // get_sphere_uv, SkyBoxLight.cpp:12-17
  float phi = atan2f(p.z, p.x);
  float theta = asinf(p.y);
  *u = 1 - (phi + kPi) * kInv2Pi;
  *v = (theta + kPiOver2) * kInvPi;
}
// Light::Le(ray) for rays that escape (PathIntegrator.cpp:54-55)
PBR_D Spec light_Le(const SceneDev& sc, const pbrgpu_light& l, V3 d) {
  if (l.type == PBRGPU_LIGHT_INFINITE) {
    V3 w = normalize(xform_vec(l.world_to_light, d));
    float s = spherical_phi(w) * kInv2Pi, t = spherical_theta(w) * kInvPi;
    return env_lookup_bilinear(sc.env, s, t);
  }
  if (l.type == PBRGPU_LIGHT_SKYBOX) {
    float u, v;
    skybox_uv(normalize(d), &u, &v);
    return skybox_value(sc.env, u, v);
  }
  return Spec(0.8f);  // Light::Le default (Light.h:58, quirk Q3)
}

// A sampled point on a light as needed by VisibilityTester (p, n, pError of the far endpoint)
struct LightSample {
  Spec Li;
  V3 wi;
  float pdf;
  V3 p1, n1, pErr1;  // Interaction at the light
};

// Triangle::Sample (area) + Shape::Sample (solid angle) + DiffuseAreaLight::Sample_Li
PBR_D void sample_area_light(const SceneDev& sc, const pbrgpu_light& l, V3 refP, float u0, float u1, LightSample* ls) {
  TriData t = load_tri(sc, l.tri);
  float su0 = sqrtf(u0);
  float b0 = 1 - su0, b1 = u1 * su0;
  float b2 = 1 - b0 - b1;
  V3 p = b0 * t.p0 + b1 * t.p1 + b2 * t.p2;
  // n = Normalize(Cross(p1 - p0, p2 - p0)) (orientation flip applied) and Area() of Triangle::Sample
  // (Triangle.cpp:352-387) are per-light constants, precomputed on the host into the light record
  V3 n(l.light_to_world[0], l.light_to_world[1], l.light_to_world[2]);
  if ((t.w1 >> kTriFlagShift) & kTriFlagNormals) {
    // mesh->n: it.n = Faceforward(Normalize(Cross(p1 - p0, p2 - p0)), b0 n0 + b1 n1 + (1 - b0 - b1) n2)   (Triangle.cpp:372-377)
    const float4 a0 = __ldg(sc.tn0 + l.tri), a1 = __ldg(sc.tn1 + l.tri), a2 = __ldg(sc.tn2 + l.tri);
    const V3 ns = b0 * V3(a0.x, a0.y, a0.z) + b1 * V3(a1.x, a1.y, a1.z) + (1 - b0 - b1) * V3(a2.x, a2.y, a2.z);
    n = faceforward(V3(l.light_to_world[8], l.light_to_world[9], l.light_to_world[10]), ns);
  }
  V3 pAbsSum = vabs(b0 * t.p0) + vabs(b1 * t.p1) + vabs(b2 * t.p2);
  ls->pErr1 = gamma_f(6) * pAbsSum;
  float area = l.light_to_world[3];
  float pdf = 1 / area;
  V3 wi = p - refP;
  if (length_sq(wi) == 0) pdf = 0;
  else {
    wi = normalize(wi);
    pdf *= dist_sq(refP, p) / absdot(n, -wi);
    if (isinf(pdf)) pdf = 0.f;
  }
  ls->p1 = p; ls->n1 = n;
  if (pdf == 0 || length_sq(p - refP) == 0) { ls->pdf = 0; ls->Li = Spec(0.f); return; }
  ls->wi = normalize(p - refP);
  ls->pdf = pdf;
  ls->Li = area_light_L(l, n, -ls->wi);
}

// Shape::Pdf(ref, wi) for a triangle light (Shape.cpp:31-42): re-intersects the light's own triangle.
PBR_D float area_light_pdf(const SceneDev& sc, const pbrgpu_light& l, V3 rayO, V3 refP, V3 wi) {
  TriData t = load_tri(sc, l.tri);
  RayPrep rp = prep_ray(rayO, wi);
  float th, b0, b1, b2;
  if (!intersect_triangle(rp, kInfinity, t.p0, t.p1, t.p2, &th, &b0, &b1, &b2)) return 0;
  V3 pl = b0 * t.p0 + b1 * t.p1 + b2 * t.p2;
  V3 nl(l.light_to_world[4], l.light_to_world[5], l.light_to_world[6]);  // Triangle::Intersect's normal of the light
  if (((t.w1 >> kTriFlagShift) & kTriFlagNormals) || (t.w1 & kTriTangentBit)) {  // isectLight.n after SetShadingGeometry
    Isect il;
    make_isect(sc, t, l.tri, b0, b1, b2, wi, &il);
    nl = il.n;
  }
  float area = l.light_to_world[3];
  float pdf = dist_sq(refP, pl) / (absdot(nl, -wi) * area);
  if (isinf(pdf)) pdf = 0.f;
  return pdf;
}

// DiffuseAreaLight over an analytic sphere: AnalyticSphere::Sample(ref, u) (uniform cone outside the sphere, uniform area
// inside) and ::Pdf(ref, wi).  Out of line for the same reason as make_isect_sphere.
static __device__ __noinline__ LightSample sample_sphere_light(const pbrgpu_light& l, V3 refP, V3 refPError, V3 refN, float u0, float u1) {
  LightSample out;
  LightSample* ls = &out;
  out.Li = Spec(0.f); out.wi = V3(0.f, 0.f, 0.f); out.pdf = 0.f;
  const V3 pCenter(l.pos[0], l.pos[1], l.pos[2]);
  const float radius = l.world_radius;
  const bool reverse = l.light_to_world[0] != 0.f;
  const V3 pOrigin = offset_ray_origin(refP, refPError, refN, pCenter - refP);
  V3 p, n, pErr;
  float pdf;
  if (dist_sq(pOrigin, pCenter) <= radius * radius) {
    // Sample(u): uniform by area
    const float z = 1 - 2 * u0;
    const float r = sqrtf(stdmax(0.f, 1.f - z * z));
    const float phi = 2 * kPi * u1;
    V3 pObj = radius * V3(r * cosf(phi), r * sinf(phi), z);
    n = normalize(pObj);
    if (reverse) n = -n;
    const float f = radius / sqrtf(pObj.x * pObj.x + pObj.y * pObj.y + pObj.z * pObj.z);
    pObj = V3(pObj.x * f, pObj.y * f, pObj.z * f);
    p = pObj + pCenter;
    pErr = gamma_f(5) * (vabs(pObj) + vabs(p));
    pdf = 1 / (4 * kPi * radius * radius);
    V3 wi = p - refP;
    if (length_sq(wi) == 0) pdf = 0;
    else {
      wi = normalize(wi);
      pdf *= dist_sq(refP, p) / absdot(n, -wi);
    }
    if (isinf(pdf)) pdf = 0.f;
  } else {
    const float dc = sqrtf(dist_sq(refP, pCenter));
    const float invDc = 1 / dc;
    const V3 wc = (pCenter - refP) * invDc;
    V3 wcX, wcY;
    coordinate_system(wc, &wcX, &wcY);
    const float sinThetaMax = radius * invDc;
    const float sinThetaMax2 = sinThetaMax * sinThetaMax;
    const float invSinThetaMax = 1 / sinThetaMax;
    const float cosThetaMax = sqrtf(stdmax(0.f, 1 - sinThetaMax2));
    float cosTheta = (cosThetaMax - 1) * u0 + 1;
    float sinTheta2 = 1 - cosTheta * cosTheta;
    if (sinThetaMax2 < 0.00068523f) {
      sinTheta2 = sinThetaMax2 * u0;
      cosTheta = sqrtf(1 - sinTheta2);
    }
    const float cosAlpha = sinTheta2 * invSinThetaMax + cosTheta * sqrtf(stdmax(0.f, 1.f - sinTheta2 * invSinThetaMax * invSinThetaMax));
    const float sinAlpha = sqrtf(stdmax(0.f, 1.f - cosAlpha * cosAlpha));
    const float phi = u1 * 2 * kPi;
    // SphericalDirection(sinAlpha, cosAlpha, phi, -wcX, -wcY, -wc)
    const V3 nWorld = sinAlpha * cosf(phi) * (-wcX) + sinAlpha * sinf(phi) * (-wcY) + cosAlpha * (-wc);
    p = pCenter + radius * nWorld;
    pErr = gamma_f(5) * vabs(p);
    n = nWorld;
    if (reverse) n = -n;
    pdf = 1 / (2 * kPi * (1 - cosThetaMax));
  }
  ls->p1 = p; ls->n1 = n; ls->pErr1 = pErr;
  // DiffuseAreaLight::Sample_Li (DiffuseLight.cpp:25-38)
  if (pdf == 0 || length_sq(p - refP) == 0) { ls->pdf = 0; ls->Li = Spec(0.f); return out; }
  ls->wi = normalize(p - refP);
  ls->pdf = pdf;
  ls->Li = area_light_L(l, n, -ls->wi);
  return out;
}
static __device__ __noinline__ float sphere_light_pdf(const pbrgpu_light& l, V3 spawnO, V3 refP, V3 refPError, V3 refN, V3 wi) {
  const V3 pCenter(l.pos[0], l.pos[1], l.pos[2]);
  const float radius = l.world_radius;
  const V3 pOrigin = offset_ray_origin(refP, refPError, refN, pCenter - refP);
  if (dist_sq(pOrigin, pCenter) <= radius * radius) {
    // Shape::Pdf (Shape.cpp:31-42): intersect the light's own shape along ref.SpawnRay(wi)
    float t;
    if (!sphere_solve(spawnO - pCenter, wi, radius, kInfinity, &t)) return 0;
    const SphereHit h = sphere_hit(pCenter, radius, l.light_to_world[0] != 0.f, spawnO, wi, t);
    float pdf = dist_sq(refP, h.p) / (absdot(h.n, -wi) * (4 * kPi * radius * radius));
    if (isinf(pdf)) pdf = 0.f;
    return pdf;
  }
  const float sinThetaMax2 = radius * radius / dist_sq(refP, pCenter);
  const float cosThetaMax = sqrtf(stdmax(0.f, 1 - sinThetaMax2));
  return 1 / (2 * kPi * (1 - cosThetaMax));
}

// refPError / refN: the reference point's error bound and normal (zero for a point inside a medium); read by sphere lights only
PBR_D void sample_light(const SceneDev& sc, const pbrgpu_light& l, V3 refP, float u0, float u1, LightSample* ls,
                        V3 refPError = V3(0.f, 0.f, 0.f), V3 refN = V3(0.f, 0.f, 0.f)) {
  ls->pdf = 0;  // EstimateDirect initialises lightPdf = 0 (Integrator.cpp:77)
  ls->Li = Spec(0.f);
  ls->n1 = V3(0, 0, 0);
  ls->pErr1 = V3(0, 0, 0);
  switch (l.type) {
    case PBRGPU_LIGHT_POINT: {
      V3 pl(l.pos[0], l.pos[1], l.pos[2]);
      ls->wi = normalize(pl - refP);
      ls->pdf = 1.f;
      ls->p1 = pl;
      ls->Li = Spec(l.L[0], l.L[1], l.L[2]) / dist_sq(pl, refP);
      break;
    }
    case PBRGPU_LIGHT_AREA_TRI: sample_area_light(sc, l, refP, u0, u1, ls); break;
    case PBRGPU_LIGHT_AREA_SPHERE: { const LightSample sph = sample_sphere_light(l, refP, refPError, refN, u0, u1); *ls = sph; break; }
    case PBRGPU_LIGHT_INFINITE: {
      const EnvDev& e = sc.env;
      float pdf1, pdf0;
      int v, dummy;
      float d1 = sample_continuous(e.margFunc, e.margCdf, e.margInt, e.height, u1, &pdf1, &v);
      float d0 = sample_continuous(e.condFunc + (size_t)v * e.width, e.condCdf + (size_t)v * (e.width + 1),
                                   __ldg(e.condInt + v), e.width, u0, &pdf0, &dummy);
      float mapPdf = pdf0 * pdf1;
      if (mapPdf == 0) return;
      float theta = d1 * kPi, phi = d0 * 2 * kPi;
      float cosTheta = cosf(theta), sinTheta = sinf(theta);
      float sinPhi = sinf(phi), cosPhi = cosf(phi);
      ls->wi = xform_vec(l.light_to_world, V3(sinTheta * cosPhi, sinTheta * sinPhi, cosTheta));
      ls->pdf = mapPdf / (2 * kPi * kPi * sinTheta);
      if (sinTheta == 0) ls->pdf = 0;
      ls->p1 = refP + ls->wi * (2 * sc.worldRadius);
      ls->Li = env_lookup_bilinear(e, d0, d1);
      break;
    }
    case PBRGPU_LIGHT_SKYBOX: {
      float z = 1 - 2 * u0;
      float r = sqrtf(stdmax(0.f, 1.f - z * z));
      float phi = 2 * kPi * u1;
      ls->wi = V3(r * cosf(phi), r * sinf(phi), z);
      ls->pdf = 1.f / (4 * kPi);
      ls->p1 = refP + ls->wi * (2 * l.world_radius);
      float u, v;
      skybox_uv(normalize(ls->wi), &u, &v);
      ls->Li = skybox_value(sc.env, u, v);
      break;
    }
  }
}

// Light::Pdf_Li
PBR_D float light_pdf_li(const SceneDev& sc, const pbrgpu_light& l, V3 spawnO, V3 refP, V3 wi, V3 refPError = V3(0.f, 0.f, 0.f),
                         V3 refN = V3(0.f, 0.f, 0.f)) {
  switch (l.type) {
    case PBRGPU_LIGHT_AREA_TRI: return area_light_pdf(sc, l, spawnO, refP, wi);
    case PBRGPU_LIGHT_AREA_SPHERE: return sphere_light_pdf(l, spawnO, refP, refPError, refN, wi);
    case PBRGPU_LIGHT_INFINITE: {
      V3 w = xform_vec(l.world_to_light, wi);
      float theta = spherical_theta(w), phi = spherical_phi(w);
      float sinTheta = sinf(theta);
      if (sinTheta == 0) return 0;
      const EnvDev& e = sc.env;
      float pu = phi * kInv2Pi, pv = theta * kInvPi;
      int iu = (int)(pu * e.width); iu = iu < 0 ? 0 : (iu > e.width - 1 ? e.width - 1 : iu);
      int iv = (int)(pv * e.height); iv = iv < 0 ? 0 : (iv > e.height - 1 ? e.height - 1 : iv);
      return (__ldg(e.condFunc + (size_t)iv * e.width + iu) / e.margInt) / (2 * kPi * kPi * sinTheta);
    }
    default: return 0;  // Point, SkyBox
  }
}

PBR_D bool is_delta_light(const pbrgpu_light& l) { return l.type == PBRGPU_LIGHT_POINT; }

}  // namespace pbr
