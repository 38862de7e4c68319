// k_shade.cuh - material-sorted shading of the Path / VolPath integrators (PathIntegrator.cpp:42-108).
#pragma once
#include "kcommon.cuh"

namespace pbr {

#ifndef PBR_SHADE_STAGE_SMEM
#define PBR_SHADE_STAGE_SMEM 1
#endif
#ifndef PBR_SHADE_MIN_BLOCKS
#define PBR_SHADE_MIN_BLOCKS 6
#endif
// K4/K5/K7: shading for one material queue.
// VOL = true is VolPathIntegrator::Li's surface branch (VolPathIntegrator.cpp:56-93): the bounce count comes from
// the path (`bounce` is ignored), UniformSampleOneLight runs for every BSDF (no NumComponents guard), the shadow
// ray carries what VisibilityTester::Tr needs, and the spawned ray takes the medium on its side of the surface.
// TEX = true (launched only for scenes with image textures): Matte / Plastic Kd may come from an ImageTexture.
template <int MAT, bool VOL = false, bool TEX = false>
__global__ void __launch_bounds__(kShadeBlock, PBR_SHADE_MIN_BLOCKS) k_shade(const __grid_constant__ RenderParams P, int bounce) {
  __shared__ uint32_t sCount, sBase;
  __shared__ uint32_t sCount3[3], sBase3[3];
#if PBR_SHADE_STAGE_SMEM
  // The rays a hit emits (shadow, MIS, continuation: 9-12 float4) wait for the block-wide queue push at the end of the
  // iteration.  Kept in registers they were spilled (80-register budget): local stores are written through to L2.  Staged
  // in shared memory, one 16-byte column per thread and field, they cost neither registers nor L2 traffic.
  __shared__ float4 sStage[VOL ? 12 : 9][kShadeBlock];
#define PBR_OUT(k) sStage[k][threadIdx.x]
#endif
  const uint32_t n = P.ws.ctr->nMat[MAT];
  const SceneDev& sc = P.sc;
  const uint32_t* queue = P.ws.qMat + (size_t)MAT * P.ws.capacity;
  for (uint32_t base = blockIdx.x * blockDim.x; base < n; base += gridDim.x * blockDim.x) {
    uint32_t i = base + threadIdx.x;
    bool active = i < n;
    bool pushShadow = false, pushMis = false, pushNext = false;
    uint32_t slot = 0;
#if PBR_SHADE_STAGE_SMEM
    float4 &shO = PBR_OUT(0), &shD = PBR_OUT(1), &shC = PBR_OUT(2), &miO = PBR_OUT(3), &miD = PBR_OUT(4), &miC = PBR_OUT(5);
    float4 &nxO = PBR_OUT(6), &nxD = PBR_OUT(7), &nxB = PBR_OUT(8);
    float4 &shP1 = PBR_OUT(VOL ? 9 : 0), &shE1 = PBR_OUT(VOL ? 10 : 0), &shN1 = PBR_OUT(VOL ? 11 : 0);
#else
    float4 shO, shD, shC, miO, miD, miC, shP1, shE1, shN1;
    float4 nxO, nxD, nxB;              // the continuation ray's record, written at its compacted position after the queue push
#endif
    uint32_t nxMeta = 0, idxLo = 0, idxHi = 0;
    if (active) {
      const uint32_t pos = queue[i];   // position of this path in ws.cur
      const PathRegs rec = load_path(P.ws.cur + pos);  // 64 bytes = two full sectors
      const float4 rd4 = rec.d, h4 = P.ws.hit[pos], be4 = rec.beta;
      const uint32_t meta = rec.meta;
      slot = rec.slot; idxLo = rec.idxLo; idxHi = rec.idxHi;
      if (VOL) bounce = (int)((meta >> kMetaBounceShift) & kMetaBounceMask);
      int medIn = -1, medOut = -1;
      if (VOL) isect_media(sc, __float_as_int(h4.x), __float_as_int(rd4.w), &medIn, &medOut);
      V3 rayD(rd4.x, rd4.y, rd4.z);
      Spec beta(be4.x, be4.y, be4.z);
      float etaScale = be4.w;
      bool specularBounce = (meta >> 16) & 1u;
      SamplerCursor cur = load_cursor(P, rec);
      TriData tri = load_tri(sc, __float_as_int(h4.x));
      Isect is;
      make_isect(sc, tri, __float_as_int(h4.x), h4.y, h4.z, h4.w, rayD, &is, V3(rec.o.x, rec.o.y, rec.o.z));
      Spec Ladd(0.f);
      // emitted light for camera rays / after specular bounces
      if ((bounce == 0 || specularBounce) && is.lightId >= 0)
        Ladd = Ladd + beta * area_light_L(sc.lights[is.lightId], is.n, -rayD);
      bool alive = bounce < P.maxDepth;
      if (alive) {
        Bsdf bsdf;
        Spec kdTex, ksTex;
        float roughTex = 0.f;
        bool haveKdTex = false, haveKsTex = false, haveRoughTex = false;
        if (TEX && (MAT == PBRGPU_MAT_MATTE || MAT == PBRGPU_MAT_PLASTIC)) {
          const pbrgpu_material& m = sc.mats[is.matId];
          if (m.kd_texture_plus1 > 0) {
            const float4 ro4 = rec.o;
            // Reference quirk: SurfaceInteraction::ComputeScatteringFunctions takes `const Ray&` (Interaction.h:86-89) and
            // hands it to ComputeDifferentials(const RayDifferential&): the implicit RayDifferential(const Ray&) conversion
            // (Geometry.h:1387) sets hasDifferentials = false, so EVERY texture lookup of a render sees zero derivatives -
            // camera rays included.  The differential code stays reachable through PBRGPU_DUMP_HIT_DIFFERENTIALS, where it is
            // checked against the reference's ComputeDifferentials called on the real RayDifferential.
            const bool cameraRay = false;
            kdTex = eval_image_texture(P, m.kd_texture_plus1, tri, __float_as_int(h4.x), h4.y, h4.z, h4.w, is, cameraRay, slot, cur.index,
                                       V3(ro4.x, ro4.y, ro4.z), rayD);
            haveKdTex = true;
          }
          if (MAT == PBRGPU_MAT_PLASTIC && m.ks_texture_plus1 > 0) {  // Ks->Evaluate(*si): the same lookup when Kd and Ks share the texture
            const float4 ro4 = rec.o;
            ksTex = (haveKdTex && m.ks_texture_plus1 == m.kd_texture_plus1) ? kdTex
                    : eval_image_texture(P, m.ks_texture_plus1, tri, __float_as_int(h4.x), h4.y, h4.z, h4.w, is, false, slot, cur.index,
                                         V3(ro4.x, ro4.y, ro4.z), rayD);
            haveKsTex = true;
          }
          if (MAT == PBRGPU_MAT_PLASTIC && m.rough_texture_plus1 > 0) {  // ImageTexture<float, float>: first channel
            const float4 ro4 = rec.o;
            roughTex = eval_image_texture(P, m.rough_texture_plus1, tri, __float_as_int(h4.x), h4.y, h4.z, h4.w, is, false, slot, cur.index,
                                          V3(ro4.x, ro4.y, ro4.z), rayD).r;
            haveRoughTex = true;
          }
        }
        make_bsdf<MAT>(sc, is, /*allowMultipleLobes=*/true, &bsdf, haveKdTex ? &kdTex : nullptr, haveKsTex ? &ksTex : nullptr,
                       haveRoughTex ? &roughTex : nullptr);
        const int nonSpec = BSDF_ALL & ~BSDF_SPECULAR;
        if ((VOL || bsdf_num_components<MAT>(bsdf, nonSpec) > 0) && sc.nLights > 0) {
          // UniformSampleOneLight (Integrator.cpp:46-69)
          float pickPdf;
          int lightNum = sample_discrete(sc.lightFunc, sc.lightCdf, sc.lightFuncInt, sc.nLights, cur.get1D(sc.hp), &pickPdf);
          if (pickPdf != 0) {
            float uL0, uL1, uS0, uS1;
            cur.get2D(sc.hp, &uL0, &uL1);
            cur.get2D(sc.hp, &uS0, &uS1);
            const pbrgpu_light& light = sc.lights[lightNum];
            // EstimateDirect (Integrator.cpp:71-177), strategy 1: sample the light
            LightSample ls;
            sample_light(sc, light, is.p, uL0, uL1, &ls, is.pError, is.n);
            if (ls.pdf > 0 && !ls.Li.is_black()) {
              Spec f = bsdf_f<MAT>(bsdf, is.wo, ls.wi, nonSpec) * absdot(ls.wi, is.ns);
              float scatteringPdf = bsdf_pdf<MAT>(bsdf, is.wo, ls.wi, nonSpec);
              if (!f.is_black()) {
                V3 origin = offset_ray_origin(is.p, is.pError, is.n, ls.p1 - is.p);
                V3 target = offset_ray_origin(ls.p1, ls.pErr1, ls.n1, origin - ls.p1);
                V3 dd = target - origin;
                Spec Ld;
                if (is_delta_light(light)) Ld = f * ls.Li / ls.pdf;
                else Ld = f * ls.Li * power_heuristic(1, ls.pdf, 1, scatteringPdf) / ls.pdf;
                Spec c = beta * (Ld / pickPdf);
                shO = make_float4(origin.x, origin.y, origin.z, 1 - kShadowEpsilon);
                shD = make_float4(dd.x, dd.y, dd.z, __uint_as_float(slot));
                shC = make_float4(c.r, c.g, c.b, 0.f);
                if (VOL) {  // p0.SpawnRayTo(p1): Ray::medium = GetMedium(d)  (Interaction.h:38-50)
                  shP1 = make_float4(ls.p1.x, ls.p1.y, ls.p1.z, __int_as_float(dot(dd, is.n) > 0 ? medOut : medIn));
                  shE1 = make_float4(ls.pErr1.x, ls.pErr1.y, ls.pErr1.z, 0.f);
                  shN1 = make_float4(ls.n1.x, ls.n1.y, ls.n1.z, 0.f);
                }
                pushShadow = true;
              }
            }
            // strategy 2: sample the BSDF
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
                if (ok) {
                  Spec c = beta * ((f * weight / scatteringPdf) / pickPdf);
                  miO = make_float4(origin.x, origin.y, origin.z, kInfinity);
                  miD = make_float4(wi.x, wi.y, wi.z, __uint_as_float(slot));
                  miC = make_float4(c.r, c.g, c.b, __int_as_float(lightNum));
                  pushMis = true;
                }
              }
            }
          }
        }
        // sample the BSDF for the next path segment (PathIntegrator.cpp:84-99)
        V3 wo = -rayD, wi;
        float pdf, u0, u1;
        int flags;
        cur.get2D(sc.hp, &u0, &u1);
        Spec f = bsdf_sample_f<MAT>(bsdf, wo, &wi, u0, u1, &pdf, BSDF_ALL, &flags);
        if (f.is_black() || pdf == 0.f) alive = false;
        if (alive) {
          beta = beta * (f * absdot(wi, is.ns) / pdf);
          specularBounce = (flags & BSDF_SPECULAR) != 0;
          if ((flags & BSDF_SPECULAR) && (flags & BSDF_TRANSMISSION)) {
            float eta = bsdf.eta;
            etaScale *= (dot(wo, is.n) > 0) ? (eta * eta) : 1 / (eta * eta);
          }
          V3 o = offset_ray_origin(is.p, is.pError, is.n, wi);
          Spec rrBeta = beta * etaScale;
          if (rrBeta.max_value() < P.rrThreshold && bounce > 3) {
            float q = stdmax(.05f, 1 - rrBeta.max_value());
            if (cur.get1D(sc.hp) < q) alive = false;
            else beta = beta / (1 - q);
          }
          if (alive) {
            nxO = make_float4(o.x, o.y, o.z, kInfinity);
            const int nextMedium = VOL ? (dot(wi, is.n) > 0 ? medOut : medIn) : -1;
            nxD = make_float4(wi.x, wi.y, wi.z, VOL ? __int_as_float(nextMedium) : 0.f);
            nxB = make_float4(beta.r, beta.g, beta.b, etaScale);
            nxMeta = (uint32_t)cur.dim | (specularBounce ? (1u << kMetaSpecBit) : 0u) |
                     (VOL ? ((uint32_t)(bounce + 1) << kMetaBounceShift) | (nextMedium >= 0 ? (1u << kMetaInMediumBit) : 0u) : 0u);
            pushNext = true;
          }
        }
      }
      if (!Ladd.is_black()) {
        float4 L = P.ws.L[slot];
        P.ws.L[slot] = make_float4(L.x + Ladd.r, L.y + Ladd.g, L.z + Ladd.b, 0.f);
      }
    }
#if PBR_PUSH3
    uint32_t pos, posMis, posNext;
    block_queue_push3(&P.ws.ctr->nShadow, &P.ws.ctr->nMis, &P.ws.ctr->nNext, pushShadow, pushMis, pushNext, sCount3, sBase3, &pos,
                      &posMis, &posNext);
    (void)sCount; (void)sBase;
#else
    uint32_t pos = block_queue_push<kShadeBlock>(&P.ws.ctr->nShadow, pushShadow, &sCount, &sBase);
    uint32_t posMis = block_queue_push<kShadeBlock>(&P.ws.ctr->nMis, pushMis, &sCount, &sBase);
    uint32_t posNext = block_queue_push<kShadeBlock>(&P.ws.ctr->nNext, pushNext, &sCount, &sBase);
    (void)sCount3; (void)sBase3;
#endif
    if (pushShadow) {
      P.ws.shO[pos] = shO; P.ws.shD[pos] = shD; P.ws.shC[pos] = shC;
      if (VOL) { P.ws.shP1[pos] = shP1; P.ws.shE1[pos] = shE1; P.ws.shN1[pos] = shN1; }
    }
    if (pushMis) { P.ws.miO[posMis] = miO; P.ws.miD[posMis] = miD; P.ws.miC[posMis] = miC; }
    if (pushNext) store_path(P.ws.nxt + posNext, nxO, nxD, nxB, idxLo, idxHi, nxMeta, slot);  // compacted: 64 contiguous bytes per lane
  }
}


}  // namespace pbr
