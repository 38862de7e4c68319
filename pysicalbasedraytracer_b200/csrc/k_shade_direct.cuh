// k_shade_direct.cuh - Whitted / DirectLighting shading (WhittedIntegrator.cpp:11-65, DirectLightingIntegrator.cpp:28-59).
#pragma once
#include "kcommon.cuh"

namespace pbr {

// ---------------------------------------------------------------------------------------------
// Whitted / DirectLighting integrators (BASELINE config 4).  Same wavefront, different
// per-hit state machine: recursion happens only through perfect specular reflection
// (SamplerIntegrator::SpecularReflect, Integrator.cpp:179-222), which is iterated here with `beta`.
//   WhittedIntegrator::Li            Integrator/WhittedIntegrator.cpp:11-65
//   DirectLightingIntegrator::Li     Integrator/DirectLightingIntegrator.cpp:28-59
//   UniformSampleAllLights           Integrator/Integrator.cpp:17-44  (sample arrays: Sampler.cpp:47-64,97-120)
//   UniformSampleOneLight (no distribution)   Integrator/Integrator.cpp:59-62
// A hit emits several shadow / MIS rays (one per light or per light sample), so the contributions are
// added to L[slot] with atomics (RenderParams::atomicL) and the queues are sized per integrator.
// ---------------------------------------------------------------------------------------------
PBR_D void push_shadow_ray(const RenderParams& P, V3 origin, V3 dd, uint32_t slot, Spec c) {
  uint32_t pos = queue_push(&P.ws.ctr->nShadow, true);
  P.ws.shO[pos] = make_float4(origin.x, origin.y, origin.z, 1 - kShadowEpsilon);
  P.ws.shD[pos] = make_float4(dd.x, dd.y, dd.z, __uint_as_float(slot));
  P.ws.shC[pos] = make_float4(c.r, c.g, c.b, 0.f);
}
PBR_D void push_mis_ray(const RenderParams& P, V3 origin, V3 wi, uint32_t slot, Spec c, int lightNum) {
  uint32_t pos = queue_push(&P.ws.ctr->nMis, true);
  P.ws.miO[pos] = make_float4(origin.x, origin.y, origin.z, kInfinity);
  P.ws.miD[pos] = make_float4(wi.x, wi.y, wi.z, __uint_as_float(slot));
  P.ws.miC[pos] = make_float4(c.r, c.g, c.b, __int_as_float(lightNum));
}

// EstimateDirect (Integrator.cpp:71-177) for one (light, uLight, uScattering); contributions are
// scaled by beta / div before they are queued.
template <int MAT>
PBR_D void estimate_direct(const RenderParams& P, const Bsdf& bsdf, const Isect& is, int lightNum, float uL0, float uL1,
                           float uS0, float uS1, Spec beta, float div, uint32_t slot) {
  const SceneDev& sc = P.sc;
  const pbrgpu_light& light = sc.lights[lightNum];
  const int nonSpec = BSDF_ALL & ~BSDF_SPECULAR;
  LightSample ls;
  sample_light(sc, light, is.p, uL0, uL1, &ls, is.pError, is.n);
  if (ls.pdf > 0 && !ls.Li.is_black()) {
    Spec f = bsdf_f<MAT>(bsdf, is.wo, ls.wi, nonSpec) * absdot(ls.wi, is.ns);
    float scatteringPdf = bsdf_pdf<MAT>(bsdf, is.wo, ls.wi, nonSpec);
    if (!f.is_black()) {
      V3 origin = offset_ray_origin(is.p, is.pError, is.n, ls.p1 - is.p);
      V3 target = offset_ray_origin(ls.p1, ls.pErr1, ls.n1, origin - ls.p1);
      Spec Ld;
      if (is_delta_light(light)) Ld = f * ls.Li / ls.pdf;
      else Ld = f * ls.Li * power_heuristic(1, ls.pdf, 1, scatteringPdf) / ls.pdf;
      push_shadow_ray(P, origin, target - origin, slot, beta * (Ld / div));
    }
  }
  if (!is_delta_light(light)) {
    V3 wi;
    float scatteringPdf;
    int sampledType;
    Spec f = bsdf_sample_f<MAT>(bsdf, is.wo, &wi, uS0, uS1, &scatteringPdf, nonSpec, &sampledType);
    f = f * absdot(wi, is.ns);
    if (!f.is_black() && scatteringPdf > 0) {
      V3 origin = offset_ray_origin(is.p, is.pError, is.n, wi);
      float weight = 1;
      bool ok = true;
      if (!(sampledType & BSDF_SPECULAR)) {
        float lightPdf = light_pdf_li(sc, light, origin, is.p, wi, is.pError, is.n);
        if (lightPdf == 0) ok = false;
        else weight = power_heuristic(1, scatteringPdf, 1, lightPdf);
      }
      if (ok) push_mis_ray(P, origin, wi, slot, beta * ((f * weight / scatteringPdf) / div), lightNum);
    }
  }
}

template <int MAT, bool TEX = false>
__global__ void __launch_bounds__(kBlock) k_shade_direct(const __grid_constant__ RenderParams P, int depth) {
  const uint32_t n = P.ws.ctr->nMat[MAT];
  const SceneDev& sc = P.sc;
  const uint32_t* queue = P.ws.qMat + (size_t)MAT * P.ws.capacity;
  for (uint32_t base = blockIdx.x * blockDim.x; base < n; base += gridDim.x * blockDim.x) {
    uint32_t i = base + threadIdx.x;
    bool active = i < n;
    bool pushNext = false;
    uint32_t slot = 0, idxLo = 0, idxHi = 0, nxMeta = 0;
    float4 nxO, nxD, nxB;
    if (active) {
      const uint32_t qpos = queue[i];  // position of this path in ws.cur
      const PathRegs rec = load_path(P.ws.cur + qpos);
      const float4 rd4 = rec.d, h4 = P.ws.hit[qpos], be4 = rec.beta;
      slot = rec.slot; idxLo = rec.idxLo; idxHi = rec.idxHi;
      V3 rayD(rd4.x, rd4.y, rd4.z);
      Spec beta(be4.x, be4.y, be4.z);
      SamplerCursor cur = load_cursor(P, rec);
      TriData tri = load_tri(sc, __float_as_int(h4.x));
      Isect is;
      make_isect(sc, tri, __float_as_int(h4.x), h4.y, h4.z, h4.w, rayD, &is, V3(rec.o.x, rec.o.y, rec.o.z));
      // L += isect.Le(wo) at every depth (WhittedIntegrator.cpp:33, DirectLightingIntegrator.cpp:41)
      if (is.lightId >= 0) {
        Spec Le = beta * area_light_L(sc.lights[is.lightId], is.n, is.wo);
        if (!Le.is_black()) { atomicAdd(&P.ws.L[slot].x, Le.r); atomicAdd(&P.ws.L[slot].y, Le.g); atomicAdd(&P.ws.L[slot].z, Le.b); }
      }
      Bsdf bsdf;
      Spec kdTex, ksTex;
      float roughTex = 0.f;
      bool haveKdTex = false, haveKsTex = false, haveRoughTex = false;
      if (TEX && (MAT == PBRGPU_MAT_MATTE || MAT == PBRGPU_MAT_PLASTIC)) {
        const pbrgpu_material& m = sc.mats[is.matId];
        if (m.kd_texture_plus1 > 0) {  // zero derivatives, like every lookup of the reference (see k_shade)
          kdTex = eval_image_texture(P, m.kd_texture_plus1, tri, __float_as_int(h4.x), h4.y, h4.z, h4.w, is, false, slot, cur.index, V3(0, 0, 0), rayD);
          haveKdTex = true;
        }
        if (MAT == PBRGPU_MAT_PLASTIC && m.ks_texture_plus1 > 0) {
          ksTex = (haveKdTex && m.ks_texture_plus1 == m.kd_texture_plus1) ? kdTex
                  : eval_image_texture(P, m.ks_texture_plus1, tri, __float_as_int(h4.x), h4.y, h4.z, h4.w, is, false, slot, cur.index, V3(0, 0, 0), rayD);
          haveKsTex = true;
        }
        if (MAT == PBRGPU_MAT_PLASTIC && m.rough_texture_plus1 > 0) {
          roughTex = eval_image_texture(P, m.rough_texture_plus1, tri, __float_as_int(h4.x), h4.y, h4.z, h4.w, is, false, slot, cur.index, V3(0, 0, 0), rayD).r;
          haveRoughTex = true;
        }
      }
      make_bsdf<MAT>(sc, is, /*allowMultipleLobes=*/false, &bsdf, haveKdTex ? &kdTex : nullptr, haveKsTex ? &ksTex : nullptr,
                     haveRoughTex ? &roughTex : nullptr);
      if (P.integrator == PBRGPU_INTEGRATOR_WHITTED) {
        for (int j = 0; j < sc.nLights; ++j) {
          float u0, u1;
          cur.get2D(sc.hp, &u0, &u1);
          LightSample ls;
          sample_light(sc, sc.lights[j], is.p, u0, u1, &ls, is.pError, is.n);
          if (ls.Li.is_black() || ls.pdf == 0) continue;
          Spec f = bsdf_f<MAT>(bsdf, is.wo, ls.wi, BSDF_ALL);
          if (f.is_black()) continue;
          V3 origin = offset_ray_origin(is.p, is.pError, is.n, ls.p1 - is.p);
          V3 target = offset_ray_origin(ls.p1, ls.pErr1, ls.n1, origin - ls.p1);
          push_shadow_ray(P, origin, target - origin, slot, beta * (f * ls.Li * absdot(ls.wi, is.ns) / ls.pdf));
        }
      } else if (sc.nLights > 0) {
        if (P.integrator == PBRGPU_INTEGRATOR_DIRECT_ONE) {
          int lightNum = min((int)(cur.get1D(sc.hp) * sc.nLights), sc.nLights - 1);
          float lightPdf = float(1) / sc.nLights;
          float uL0, uL1, uS0, uS1;
          cur.get2D(sc.hp, &uL0, &uL1);
          cur.get2D(sc.hp, &uS0, &uS1);
          estimate_direct<MAT>(P, bsdf, is, lightNum, uL0, uL1, uS0, uS1, beta, lightPdf, slot);
        } else {  // UniformSampleAll: two sample arrays per light per recursion depth
          int arr = 2 * sc.nLights * depth;
          const int nArrays = (P.arrayEndDim - 5) / 2;
          // sample k of pixel-sample s of an array lives at GetIndexForSample(s * n + k) (GlobalSampler::StartPixel, Sampler.cpp:97-120)
          const bool sobol = sc.hp.type == PBRGPU_SAMPLER_SOBOL;
          const int64_t stride = sc.hp.sampleStride;
          int px, py;
          long long sNum;
          uint32_t pix;
          slot_to_pixel(P, slot, &px, &py, &sNum, &pix);
          const int64_t pixelOffset = sobol ? 0 : P.pixelOffsetOf(cur.index);
          const int64_t s = sobol ? (int64_t)sNum : (cur.index - pixelOffset) / stride;  // pixel-sample number
          for (int j = 0; j < sc.nLights; ++j) {
            int nS = sc.lights[j].n_samples;
            if (arr + 1 >= nArrays) {  // arrays exhausted: one plain sample (Integrator.cpp:29-33)
              float uL0, uL1, uS0, uS1;
              cur.get2D(sc.hp, &uL0, &uL1);
              cur.get2D(sc.hp, &uS0, &uS1);
              estimate_direct<MAT>(P, bsdf, is, j, uL0, uL1, uS0, uS1, beta, 1.f, slot);
            } else {
              int dimL = 5 + 2 * arr, dimS = 5 + 2 * (arr + 1);
              for (int k = 0; k < nS; ++k) {
                const int64_t idx = sobol ? sobol_index_for_sample(px, py, s * nS + k, sc.hp) : pixelOffset + (s * nS + k) * stride;
                float uL0 = halton_sample_dimension(idx, dimL, sc.hp), uL1 = halton_sample_dimension(idx, dimL + 1, sc.hp);
                float uS0 = halton_sample_dimension(idx, dimS, sc.hp), uS1 = halton_sample_dimension(idx, dimS + 1, sc.hp);
                estimate_direct<MAT>(P, bsdf, is, j, uL0, uL1, uS0, uS1, beta, (float)nS, slot);
              }
            }
            arr += 2;
          }
        }
      }
      // SpecularReflect (Integrator.cpp:179-222): only while depth + 1 < maxDepth
      if (depth + 1 < P.maxDepth) {
        float u0, u1, pdf;
        int flags;
        V3 wi;
        cur.get2D(sc.hp, &u0, &u1);
        Spec f = bsdf_sample_f<MAT>(bsdf, is.wo, &wi, u0, u1, &pdf, BSDF_REFLECTION | BSDF_SPECULAR, &flags);
        if (pdf > 0.f && !f.is_black() && absdot(wi, is.ns) != 0.f) {
          beta = beta * (f * absdot(wi, is.ns) / pdf);
          V3 o = offset_ray_origin(is.p, is.pError, is.n, wi);
          nxO = make_float4(o.x, o.y, o.z, kInfinity);
          nxD = make_float4(wi.x, wi.y, wi.z, 0.f);
          nxB = make_float4(beta.r, beta.g, beta.b, 1.f);
          nxMeta = (uint32_t)cur.dim;
          pushNext = true;
        }
      }
    }
    uint32_t pos = queue_push(&P.ws.ctr->nNext, pushNext);
    if (pushNext) store_path(P.ws.nxt + pos, nxO, nxD, nxB, idxLo, idxHi, nxMeta, slot);
  }
}


}  // namespace pbr
