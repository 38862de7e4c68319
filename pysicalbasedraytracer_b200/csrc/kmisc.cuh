// kmisc.cuh - the small non-template kernels of the wavefront (raygen, classification, media, miss, MIS resolve,
// film) and the parity hooks.  Included by pbrgpu.cu only (one definition per kernel).
#pragma once
#include "kcommon.cuh"

namespace pbr {

// K1: sampler + raygen.  Sampler::GetCameraSample (Sampler.cpp:10-21): dims 0,1 film, 2 time, 3,4 lens.
__global__ void __launch_bounds__(kBlock) k_raygen(const __grid_constant__ RenderParams P) {
  uint32_t n = (uint32_t)(P.w * P.h) * (uint32_t)P.nsWave;
  for (uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x; slot < n; slot += gridDim.x * blockDim.x) {
    int px, py; long long s; uint32_t pix;
    slot_to_pixel(P, slot, &px, &py, &s, &pix);
    int64_t idx = sampler_index_for_sample(px, py, s, P.sc.hp);
    float u0 = sampler_film_sample(idx, 0, px, P.sc.hp);
    float u1 = sampler_film_sample(idx, 1, py, P.sc.hp);
    float lu = 0, lv = 0;
    if (P.sc.cam.lensRadius > 0) {
      lu = halton_sample_dimension(idx, 3, P.sc.hp);
      lv = halton_sample_dimension(idx, 4, P.sc.hp);
    }
    V3 o, d;
    generate_ray(P.sc.cam, (float)px + u0, (float)py + u1, lu, lv, &o, &d);
    P.ws.L[slot] = make_float4(0.f, 0.f, 0.f, 0.f);
    // camera paths: queue position == slot.  meta: next dimension 5; specularBounce = false; VolPath: bounces = 0, in-medium flag
    store_path(P.ws.cur + slot, make_float4(o.x, o.y, o.z, kInfinity),
               make_float4(d.x, d.y, d.z, P.integrator == PBRGPU_INTEGRATOR_VOLPATH ? __int_as_float(P.sc.cam.medium) : 0.f),
               make_float4(1.f, 1.f, 1.f, 1.f), (uint32_t)((uint64_t)idx & 0xffffffffu), (uint32_t)((uint64_t)idx >> 32),
               5u | ((P.integrator == PBRGPU_INTEGRATOR_VOLPATH && P.sc.cam.medium >= 0) ? (1u << kMetaInMediumBit) : 0u), slot);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) P.ws.ctr->nCur = n;
}

// Bin the live paths by the material type of what they hit (material-sorted shading queues).  Each block
// bins a tile of kClassifyTile entries: per-queue counts in shared memory, ONE global atomicAdd per
// (block, queue), then scattered writes - three orders of magnitude fewer same-address atomics than one
// per warp.  (The order of entries inside a queue does not influence any result: per-path state only.)
constexpr int kClassifyItems = 8;
constexpr int kClassifyTile = kBlock * kClassifyItems;

__global__ void __launch_bounds__(kBlock) k_classify(const __grid_constant__ RenderParams P) {
  __shared__ uint32_t cnt[kNumQueues], basePos[kNumQueues];
  const uint32_t n = P.ws.ctr->nCur;
  if (blockIdx.x == 0 && threadIdx.x == 0) {  // the shadow / MIS queues of the previous bounce are consumed
    if (P.deferShadow != 1) P.ws.ctr->nShadow = 0;
    P.ws.ctr->nMis = 0;
    P.ws.ctr->cursor[0] = 0;
  }
  for (uint32_t tile = blockIdx.x * kClassifyTile; tile < n; tile += gridDim.x * kClassifyTile) {
    if (threadIdx.x < kNumQueues) cnt[threadIdx.x] = 0;
    __syncthreads();
    uint32_t slot[kClassifyItems], rank[kClassifyItems];
    int q[kClassifyItems];
#pragma unroll
    for (int k = 0; k < kClassifyItems; ++k) {
      const uint32_t i = tile + k * kBlock + threadIdx.x;
      q[k] = -1;
      if (i < n) {
        slot[k] = i;  // position in ws.cur
        q[k] = (int)P.ws.qType[i];
        if (q[k] >= kNumQueues) q[k] = -1;  // kQueueDead: the path ended in k_medium_sample
        else rank[k] = atomicAdd(&cnt[q[k]], 1u);  // shared-memory atomic: rank inside this block's tile
      }
    }
    __syncthreads();
    if (threadIdx.x < kNumQueues) basePos[threadIdx.x] = cnt[threadIdx.x] ? atomicAdd(&P.ws.ctr->nMat[threadIdx.x], cnt[threadIdx.x]) : 0u;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kClassifyItems; ++k)
      if (q[k] >= 0) P.ws.qMat[(size_t)q[k] * P.ws.capacity + basePos[q[k]] + rank[k]] = slot[k];
    __syncthreads();
  }
}

// `if (ray.medium) beta *= ray.medium->Sample(ray, sampler, &mi); if (beta.IsBlack()) break;`  (:35-36) for every
// live path whose segment lies in a medium, in two halves around k_trace (media.cuh: medium_sample_distance):
//   k_medium_presample  draws the two sampler values and makes the sampled free-flight distance the ray's tMax
//   k_medium_sample     after the trace: hit below that distance -> surface, weight Tr/pdf; no hit -> MediumInteraction
//                       at that distance, weight Tr*sigma_s/pdf, path re-routed to the medium queue
__global__ void __launch_bounds__(kBlock) k_medium_presample(const __grid_constant__ RenderParams P) {
  const uint32_t n = P.ws.ctr->nCur;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    PathRec* rec = P.ws.cur + i;
    const uint32_t meta = rec->meta;
    if (!((meta >> kMetaInMediumBit) & 1u)) continue;  // vacuum: nothing to sample (4 bytes read)
    const PathRegs r = load_path(rec);
    const float4 rd4 = r.d;
    const int med = __float_as_int(rd4.w);
    SamplerCursorX4 cur(load_cursor(P, r));
    cur.prefetch(P.sc.hp, 2);
    const float u0 = cur.get1D(P.sc.hp);
    const float u1 = cur.get1D(P.sc.hp);
    rec->o.w = medium_sample_distance(P.sc.media[med], u0, u1, length(V3(rd4.x, rd4.y, rd4.z)));
    rec->meta = (meta & ~0xffffu) | (uint32_t)cur.dim;
  }
}

__global__ void __launch_bounds__(kBlock) k_medium_sample(const __grid_constant__ RenderParams P) {
  const uint32_t n = P.ws.ctr->nCur;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    PathRec* rec = P.ws.cur + i;
    if (!((rec->meta >> kMetaInMediumBit) & 1u)) continue;
    const float4 rd4 = rec->d;
    const int med = __float_as_int(rd4.w);
    const bool hit = __float_as_int(P.ws.hit[i].x) >= 0;
    const float t = rec->o.w;                    // k_trace: hit distance, or the sampled distance when nothing was hit
    const bool sampled = !hit && t < kInfinity;  // t < ray.tMax
    const Spec w = medium_sample_weight(P.sc.media[med], t, sampled, length(V3(rd4.x, rd4.y, rd4.z)));
    const float4 be4 = rec->beta;
    const Spec beta = Spec(be4.x, be4.y, be4.z) * w;
    if (beta.is_black()) { P.ws.qType[i] = kQueueDead; continue; }
    rec->beta = make_float4(beta.r, beta.g, beta.b, be4.w);
    if (sampled) P.ws.qType[i] = (uint32_t)kQueueMedium;  // mi.p = ray(t), t stays in rayO.w
  }
}

// Surface without a material = medium boundary (:65-70): emission rule, maxDepth test, then
// `ray = isect.SpawnRay(ray.d); bounces--; continue;` - same direction, same bounce count, no Russian roulette.
__global__ void __launch_bounds__(kBlock) k_shade_null(const __grid_constant__ RenderParams P) {
  const uint32_t n = P.ws.ctr->nMat[kQueueNull];
  const SceneDev& sc = P.sc;
  const uint32_t* queue = P.ws.qMat + (size_t)kQueueNull * P.ws.capacity;
  for (uint32_t base = blockIdx.x * blockDim.x; base < n; base += gridDim.x * blockDim.x) {
    const uint32_t i = base + threadIdx.x;
    bool pushNext = false;
    PathRegs rec;
    float4 nxO, nxD;
    uint32_t nxMeta = 0;
    if (i < n) {
      const uint32_t qpos = queue[i];
      rec = load_path(P.ws.cur + qpos);
      const uint32_t slot = rec.slot;
      const float4 rd4 = rec.d, h4 = P.ws.hit[qpos];
      const uint32_t meta = rec.meta;
      const int bounce = (int)((meta >> kMetaBounceShift) & kMetaBounceMask);
      const bool specularBounce = (meta >> kMetaSpecBit) & 1u;
      const V3 rayD(rd4.x, rd4.y, rd4.z);
      const int tri = __float_as_int(h4.x);
      const TriData td = load_tri(sc, tri);
      Isect is;
      make_isect(sc, td, tri, h4.y, h4.z, h4.w, rayD, &is, V3(rec.o.x, rec.o.y, rec.o.z));
      if ((bounce == 0 || specularBounce) && is.lightId >= 0) {
        const float4 be4 = rec.beta;
        const Spec Ladd = Spec(be4.x, be4.y, be4.z) * area_light_L(sc.lights[is.lightId], is.n, -rayD);
        if (!Ladd.is_black()) {
          const float4 L = P.ws.L[slot];
          P.ws.L[slot] = make_float4(L.x + Ladd.r, L.y + Ladd.g, L.z + Ladd.b, 0.f);
        }
      }
      if (bounce < P.maxDepth) {
        int mIn, mOut;
        isect_media(sc, tri, __float_as_int(rd4.w), &mIn, &mOut);
        const V3 o = offset_ray_origin(is.p, is.pError, is.n, rayD);
        nxO = make_float4(o.x, o.y, o.z, kInfinity);
        const int nextMedium = dot(rayD, is.n) > 0 ? mOut : mIn;
        nxD = make_float4(rayD.x, rayD.y, rayD.z, __int_as_float(nextMedium));
        nxMeta = (meta & ~(1u << kMetaInMediumBit)) | (nextMedium >= 0 ? (1u << kMetaInMediumBit) : 0u);
        pushNext = true;
      }
    }
    const uint32_t pos = queue_push(&P.ws.ctr->nNext, pushNext);
    if (pushNext) store_path(P.ws.nxt + pos, nxO, nxD, rec.beta, rec.idxLo, rec.idxHi, nxMeta, rec.slot);
  }
}

// Scattering inside a medium (:39-54): direct lighting with the phase function in place of the BSDF
// (EstimateDirect's MediumInteraction branches, Integrator.cpp:95-100,137-142), then Sample_p for the next direction.
__global__ void __launch_bounds__(kShadeBlock) k_shade_medium(const __grid_constant__ RenderParams P) {
  __shared__ uint32_t sCount, sBase;
  const uint32_t n = P.ws.ctr->nMat[kQueueMedium];
  const SceneDev& sc = P.sc;
  const uint32_t* queue = P.ws.qMat + (size_t)kQueueMedium * P.ws.capacity;
  for (uint32_t base = blockIdx.x * blockDim.x; base < n; base += gridDim.x * blockDim.x) {
    const uint32_t i = base + threadIdx.x;
    bool pushShadow = false, pushMis = false, pushNext = false;
    uint32_t slot = 0, idxLo = 0, idxHi = 0, nxMeta = 0;
    float4 shO, shD, shC, miO, miD, miC, shP1, shE1, shN1, nxO, nxD, nxB;
    if (i < n) {
      const PathRegs rec = load_path(P.ws.cur + queue[i]);
      slot = rec.slot; idxLo = rec.idxLo; idxHi = rec.idxHi;
      const float4 ro4 = rec.o, rd4 = rec.d, be4 = rec.beta;
      const uint32_t meta = rec.meta;
      const int bounce = (int)((meta >> kMetaBounceShift) & kMetaBounceMask);
      if (bounce < P.maxDepth) {
        const int med = __float_as_int(rd4.w);
        const float g = sc.media[med].g;
        const V3 rayD(rd4.x, rd4.y, rd4.z);
        const V3 p = V3(ro4.x, ro4.y, ro4.z) + rayD * ro4.w;  // mi.p = ray(t)
        const V3 wo = -rayD;                                  // mi.wo (not normalised: Interaction.h:21-22)
        const V3 zero(0, 0, 0);
        Spec beta(be4.x, be4.y, be4.z);
        const float etaScale = be4.w;
        SamplerCursorX4 cur(load_cursor(P, rec));
        // sampler dimensions in groups with interleaved digit loops: light pick + light sample, then phase sample + Sample_p
        if (sc.nLights > 0) cur.prefetch(sc.hp, 3);
        if (sc.nLights > 0) {
          float pickPdf;
          const int lightNum = sample_discrete(sc.lightFunc, sc.lightCdf, sc.lightFuncInt, sc.nLights, cur.get1D(sc.hp), &pickPdf);
          if (pickPdf != 0) {
            float uL0, uL1, uS0, uS1;
            cur.get2D(sc.hp, &uL0, &uL1);
            cur.prefetch(sc.hp, 4);  // uS0, uS1 and Sample_p's u0, u1
            cur.get2D(sc.hp, &uS0, &uS1);
            const pbrgpu_light& light = sc.lights[lightNum];
            LightSample ls;
            sample_light(sc, light, p, uL0, uL1, &ls);
            if (ls.pdf > 0 && !ls.Li.is_black()) {
              const Spec f(hg_p(wo, ls.wi, g));
              // the reference leaves scatteringPdf = 0 in this branch (Integrator.cpp:95-100), so the light-sampling
              // strategy gets PowerHeuristic(1, lightPdf, 1, 0): reproduced as is
              const float scatteringPdf = 0;
              if (!f.is_black()) {
                const V3 origin = offset_ray_origin(p, zero, zero, ls.p1 - p);
                const V3 target = offset_ray_origin(ls.p1, ls.pErr1, ls.n1, origin - ls.p1);
                const V3 dd = target - origin;
                Spec Ld;
                if (is_delta_light(light)) Ld = f * ls.Li / ls.pdf;
                else Ld = f * ls.Li * power_heuristic(1, ls.pdf, 1, scatteringPdf) / ls.pdf;
                const Spec c = beta * (Ld / pickPdf);
                shO = make_float4(origin.x, origin.y, origin.z, 1 - kShadowEpsilon);
                shD = make_float4(dd.x, dd.y, dd.z, __uint_as_float(slot));
                shC = make_float4(c.r, c.g, c.b, 0.f);
                shP1 = make_float4(ls.p1.x, ls.p1.y, ls.p1.z, __int_as_float(med));
                shE1 = make_float4(ls.pErr1.x, ls.pErr1.y, ls.pErr1.z, 0.f);
                shN1 = make_float4(ls.n1.x, ls.n1.y, ls.n1.z, 0.f);
                pushShadow = true;
              }
            }
            if (!is_delta_light(light)) {
              V3 wi;
              const float ph = hg_sample_p(wo, &wi, uS0, uS1, g);
              const Spec f(ph);
              const float scatteringPdf = ph;
              if (!f.is_black() && scatteringPdf > 0) {
                const float lightPdf = light_pdf_li(sc, light, p, p, wi);
                if (lightPdf != 0) {
                  const float weight = power_heuristic(1, scatteringPdf, 1, lightPdf);
                  const Spec c = beta * ((f * weight / scatteringPdf) / pickPdf);
                  miO = make_float4(p.x, p.y, p.z, kInfinity);
                  miD = make_float4(wi.x, wi.y, wi.z, __uint_as_float(slot));
                  miC = make_float4(c.r, c.g, c.b, __int_as_float(lightNum));
                  pushMis = true;
                }
              }
            }
          }
        }
        // mi.phase->Sample_p(wo, &wi, sampler.Get2D()); ray = mi.SpawnRay(wi); specularBounce = false
        float u0, u1;
        cur.ensure(sc.hp, 2);
        cur.get2D(sc.hp, &u0, &u1);
        V3 wi;
        hg_sample_p(wo, &wi, u0, u1, g);
        bool alive = true;
        const Spec rrBeta = beta * etaScale;
        if (rrBeta.max_value() < P.rrThreshold && bounce > 3) {
          const float q = stdmax(.05f, 1 - rrBeta.max_value());
          if (cur.get1D(sc.hp) < q) alive = false;
          else beta = beta / (1 - q);
        }
        if (alive) {
          nxO = make_float4(p.x, p.y, p.z, kInfinity);
          nxD = make_float4(wi.x, wi.y, wi.z, __int_as_float(med));
          nxB = make_float4(beta.r, beta.g, beta.b, etaScale);
          nxMeta = (uint32_t)cur.dim | ((uint32_t)(bounce + 1) << kMetaBounceShift) | (1u << kMetaInMediumBit);
          pushNext = true;
        }
      }
    }
    uint32_t pos = block_queue_push<kShadeBlock>(&P.ws.ctr->nShadow, pushShadow, &sCount, &sBase);
    if (pushShadow) {
      P.ws.shO[pos] = shO; P.ws.shD[pos] = shD; P.ws.shC[pos] = shC;
      P.ws.shP1[pos] = shP1; P.ws.shE1[pos] = shE1; P.ws.shN1[pos] = shN1;
    }
    pos = block_queue_push<kShadeBlock>(&P.ws.ctr->nMis, pushMis, &sCount, &sBase);
    if (pushMis) { P.ws.miO[pos] = miO; P.ws.miD[pos] = miD; P.ws.miC[pos] = miC; }
    pos = block_queue_push<kShadeBlock>(&P.ws.ctr->nNext, pushNext, &sCount, &sBase);
    if (pushNext) store_path(P.ws.nxt + pos, nxO, nxD, nxB, idxLo, idxHi, nxMeta, slot);
  }
}

// Rays that left the scene: Le of the infinite lights (PathIntegrator.cpp:52-56), path ends (:60).
__global__ void __launch_bounds__(kBlock) k_shade_miss(const __grid_constant__ RenderParams P, int bounce) {
  const uint32_t n = P.ws.ctr->nMat[kQueueMiss];
  const uint32_t* queue = P.ws.qMat + (size_t)kQueueMiss * P.ws.capacity;
  const bool volInteg = P.integrator == PBRGPU_INTEGRATOR_VOLPATH;
  const bool pathInteg = P.integrator == PBRGPU_INTEGRATOR_PATH || volInteg;
  if (pathInteg && P.sc.nInfinite == 0) return;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const PathRec* rec = P.ws.cur + queue[i];
    const uint32_t meta = rec->meta;
    const uint32_t slot = rec->slot;
    bool specularBounce = (meta >> kMetaSpecBit) & 1u;
    if (volInteg) bounce = (int)((meta >> kMetaBounceShift) & kMetaBounceMask);
    if (pathInteg && !(bounce == 0 || specularBounce)) continue;
    float4 rd = rec->d, be = rec->beta;
    Spec beta(be.x, be.y, be.z);
    Spec Ladd(0.f);
    if (pathInteg) {
      for (int k = 0; k < P.sc.nInfinite; ++k)
        Ladd = Ladd + beta * light_Le(P.sc, P.sc.lights[P.sc.infinite[k]], V3(rd.x, rd.y, rd.z));
    } else {
      // Whitted / DirectLighting sum Le over ALL lights; Light::Le defaults to 0.8 (Light.h:58, quirk Q3)
      Spec sum(0.f);
      for (int k = 0; k < P.sc.nLights; ++k) sum = sum + light_Le(P.sc, P.sc.lights[k], V3(rd.x, rd.y, rd.z));
      Ladd = beta * sum;
    }
    float4 L = P.ws.L[slot];
    P.ws.L[slot] = make_float4(L.x + Ladd.r, L.y + Ladd.g, L.z + Ladd.b, 0.f);
  }
}

// K6: resolve the MIS rays traced by k_trace<TRACE_MIS> (Integrator.cpp:157-172): the hit counts only
// if it is the sampled light's own triangle; a miss asks the light for Le.
__global__ void __launch_bounds__(kBlock) k_resolve_mis(const __grid_constant__ RenderParams P) {
  const uint32_t n = P.ws.ctr->nMis;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    float4 o = P.ws.miO[i], d = P.ws.miD[i], c = P.ws.miC[i];
    V3 wi(d.x, d.y, d.z);
    int tri = __float_as_int(o.w);
    int lightId = __float_as_int(c.w);
    const pbrgpu_light& light = P.sc.lights[lightId];
    Spec Li(0.f);
    if (tri >= 0) {
      if (__float_as_int(__ldg(&P.sc.tv[3 * tri + 2].w)) == lightId) {
        if (light.type == PBRGPU_LIGHT_AREA_SPHERE) {  // the hit normal of the analytic sphere along this ray
          const V3 c(light.pos[0], light.pos[1], light.pos[2]), ro(o.x, o.y, o.z);
          float t;
          if (sphere_solve(ro - c, wi, light.world_radius, kInfinity, &t))
            Li = area_light_L(light, sphere_hit(c, light.world_radius, light.light_to_world[0] != 0.f, ro, wi, t).n, -wi);
        } else {
          const float4 n4 = __ldg(P.sc.tf + 3 * tri);
          V3 nh(n4.x, n4.y, n4.z);
          const TriData td = load_tri(P.sc, tri);
          if (((td.w1 >> kTriFlagShift) & kTriFlagNormals) || (td.w1 & kTriTangentBit)) {
            // emissive mesh with shading normals / tangents: the hit normal is face-forwarded by the interpolated shading
            // normal (SetShadingGeometry), which needs the barycentrics of this hit
            const RayPrep rp = prep_ray(V3(o.x, o.y, o.z), wi);
            float th, b0, b1, b2;
            if (intersect_triangle(rp, kInfinity, td.p0, td.p1, td.p2, &th, &b0, &b1, &b2)) {
              Isect il;
              make_isect(P.sc, td, tri, b0, b1, b2, wi, &il);
              nh = il.n;
            }
          }
          Li = area_light_L(light, nh, -wi);
        }
      }
    } else {
      Li = light_Le(P.sc, light, wi);
    }
    if (!Li.is_black()) {
      uint32_t slot = __float_as_uint(d.w);
      if (P.atomicL) {
        atomicAdd(&P.ws.L[slot].x, c.x * Li.r); atomicAdd(&P.ws.L[slot].y, c.y * Li.g); atomicAdd(&P.ws.L[slot].z, c.z * Li.b);
      } else {
        float4 L = P.ws.L[slot];
        P.ws.L[slot] = make_float4(L.x + c.x * Li.r, L.y + c.y * Li.g, L.z + c.z * Li.b, 0.f);
      }
    }
  }
}

// wave start: every queue empty
__global__ void k_wave_begin(Counters* c) {
  c->nCur = c->nNext = c->nShadow = c->nMis = 0;
  for (int k = 0; k < 8; ++k) c->nMat[k] = 0;
  for (int k = 0; k < 4; ++k) c->cursor[k] = 0;
}

// bounce bookkeeping (host swaps qCur/qNext pointers each bounce)
__global__ void k_next_bounce(Counters* c) {
  c->nCur = c->nNext;
  c->nNext = 0;
  for (int k = 0; k < 8; ++k) c->nMat[k] = 0;
  for (int k = 0; k < 4; ++k) c->cursor[k] = 0;
}

// VolPath tail: the shadow rays that the tail iterations emitted were traced together by one k_trace launch
// (RenderParams::deferShadow == 2), which left `contribution x transmittance` in shC.  Add them to their paths, one launch
// per iteration's range of the queue: a path has at most one shadow ray per iteration, so the plain read-modify-write is
// race-free and a path's sum is formed in iteration order (deterministic).
__global__ void __launch_bounds__(kBlock) k_shadow_add(const __grid_constant__ RenderParams P, uint32_t begin, uint32_t end, int resetCount) {
  for (uint32_t e = begin + blockIdx.x * blockDim.x + threadIdx.x; e < end; e += gridDim.x * blockDim.x) {
    const uint32_t slot = __float_as_uint(P.ws.shD[e].w);
    const float4 c = P.ws.shC[e];
    float4 L = P.ws.L[slot];
    P.ws.L[slot] = make_float4(L.x + c.x, L.y + c.y, L.z + c.z, 0.f);
  }
  if (resetCount && blockIdx.x == 0 && threadIdx.x == 0) { P.ws.ctr->nShadow = 0; P.ws.ctr->cursor[0] = 0; }
}

// wave end: per pixel, add the wave's per-path radiance in sample order (deterministic).
__global__ void __launch_bounds__(kBlock) k_film_accumulate(const __grid_constant__ RenderParams P) {
  uint32_t nPix = (uint32_t)(P.w * P.h);
  for (uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < nPix; p += gridDim.x * blockDim.x) {
    float4 a = reinterpret_cast<float4*>(P.accum)[p];
    for (int j = 0; j < P.nsWave; ++j) {
      float4 L = P.ws.L[(size_t)j * nPix + p];
      a.x += L.x; a.y += L.y; a.z += L.z;
    }
    a.w += (float)P.nsWave;
    reinterpret_cast<float4*>(P.accum)[p] = a;
  }
}

// K8: film resolve = mean, RGB->XYZ->RGB, sRGB gamma, 8-bit, Y flip (Integrator.cpp:313-344, Spectrum.h:10-25)
__global__ void k_resolve_film(const float4* sum, int w, int h, float invSpp, int useCount, float* outRgb, unsigned char* outRgba8) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= w * h) return;
  float4 s = sum[p];
  float r, g, b;
  if (useCount) { r = s.x / s.w; g = s.y / s.w; b = s.z / s.w; }  // colObj /= spp : true divide per channel
  else { r = s.x * invSpp; g = s.y * invSpp; b = s.z * invSpp; }
  if (outRgb) { outRgb[3 * p] = r; outRgb[3 * p + 1] = g; outRgb[3 * p + 2] = b; }
  if (outRgba8) {
    float X = 0.412453f * r + 0.357580f * g + 0.180423f * b;
    float Y = 0.212671f * r + 0.715160f * g + 0.072169f * b;
    float Z = 0.019334f * r + 0.119193f * g + 0.950227f * b;
    float rgb[3];
    rgb[0] = 3.240479f * X - 1.537150f * Y - 0.498535f * Z;
    rgb[1] = -0.969256f * X + 1.875991f * Y + 0.041556f * Z;
    rgb[2] = 0.055648f * X - 0.204043f * Y + 1.057311f * Z;
    int x = p % w, y = p / w;
    int q = x + (h - 1 - y) * w;
    for (int k = 0; k < 3; ++k) {
      float v = rgb[k];
      float gc = (v <= 0.0031308f) ? 12.92f * v : 1.055f * powf(v, (1.f / 2.4f)) - 0.055f;
      outRgba8[4 * q + k] = (unsigned char)clampf(255.f * gc + 0.5f, 0.f, 255.f);
    }
    outRgba8[4 * q + 3] = 255;
  }
}

// ---- parity hooks ----
__global__ void k_dump_sampler(HaltonParams hp, int x0, int y0, int w, int h, long long s0, int ns, int nDims, float* out) {
  long long n = (long long)w * h * ns * nDims;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    int d = (int)(i % nDims);
    long long r = i / nDims;
    int sj = (int)(r % ns);
    long long p = r / ns;
    int px = x0 + (int)(p % w), py = y0 + (int)(p / w);
    int64_t idx = sampler_index_for_sample(px, py, s0 + sj, hp);
    out[i] = d < 2 ? sampler_film_sample(idx, d, d == 0 ? px : py, hp) : halton_sample_dimension(idx, d, hp);
  }
}

__global__ void __launch_bounds__(kBlock) k_dump_primary(const __grid_constant__ RenderParams P, int ns, float* outRays, int* outHits,
                                                         float* outDiff = nullptr) {
  __shared__ uint2 sstk[kBlock / 32][kSmemStack][32];
  uint2* sstack = &sstk[threadIdx.x >> 5][0][threadIdx.x & 31];
  long long n = (long long)P.w * P.h * ns;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    int sj = (int)(i % ns);
    long long p = i / ns;
    int px = P.x0 + (int)(p % P.w), py = P.y0 + (int)(p / P.w);
    int64_t idx = sampler_index_for_sample(px, py, P.sBegin + sj, P.sc.hp);
    float u0 = sampler_film_sample(idx, 0, px, P.sc.hp), u1 = sampler_film_sample(idx, 1, py, P.sc.hp);
    float lu = 0, lv = 0;
    if (P.sc.cam.lensRadius > 0) { lu = halton_sample_dimension(idx, 3, P.sc.hp); lv = halton_sample_dimension(idx, 4, P.sc.hp); }
    V3 o, d;
    generate_ray(P.sc.cam, (float)px + u0, (float)py + u1, lu, lv, &o, &d);
    if (outRays) {
      float* r = outRays + 8 * i;
      r[0] = o.x; r[1] = o.y; r[2] = o.z; r[3] = kInfinity; r[4] = d.x; r[5] = d.y; r[6] = d.z; r[7] = 0.f;
    }
    if (outDiff) {  // hit flag, uv, dudx, dvdx, dudy, dvdy of the camera ray (SurfaceInteraction::ComputeDifferentials)
      HitRec h;
      unsigned a = 0, b = 0;
      traverse<false, false, true>(P.sc, o, d, kInfinity, &h, sstack, &a, &b);
      float* q = outDiff + 8 * i;
      for (int k = 0; k < 8; ++k) q[k] = 0.f;
      if (h.tri >= 0) {
        const TriData tri = load_tri(P.sc, h.tri);
        Isect is;
        make_isect(P.sc, tri, h.tri, h.b0, h.b1, h.b2, d, &is, o);
        float u, v;
        UvDiff dd;
        hit_uv_differentials(P, tri, h.tri, h.b0, h.b1, h.b2, is.p, is.n, true, px, py, idx, o, d, &u, &v, &dd);
        q[0] = 1.f; q[1] = u; q[2] = v; q[3] = dd.dudx; q[4] = dd.dvdx; q[5] = dd.dudy; q[6] = dd.dvdy;
      }
    }
    if (outHits) {
      HitRec h;
      unsigned a = 0, b = 0;
      traverse<false, false, true>(P.sc, o, d, kInfinity, &h, sstack, &a, &b);
      int* q = outHits + 4 * i;
      q[0] = h.tri >= 0 ? __float_as_int(__ldg(&P.sc.tv[3 * h.tri].w)) : -1;
      q[1] = __float_as_int(h.t);
      q[2] = h.tri >= 0 ? __float_as_int(h.b0) : 0;
      q[3] = h.tri >= 0 ? __float_as_int(h.b1) : 0;
    }
  }
}

__global__ void __launch_bounds__(kBlock) k_texture_eval(SceneDev sc, int tex, const float* rows, long long n, float* out) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float* r = rows + 8 * i;
    UvDiff dd{r[2], r[3], r[4], r[5]};
    const Spec v = tex_evaluate(sc.textures[tex], sc.ewaLut, r[0], r[1], dd);
    out[4 * i] = v.r; out[4 * i + 1] = v.g; out[4 * i + 2] = v.b; out[4 * i + 3] = 0.f;
  }
}

__global__ void __launch_bounds__(kBlock) k_trace_user(const __grid_constant__ RenderParams P, const float* rays, long long n, int anyHit, int* out) {
  __shared__ uint2 sstk[kBlock / 32][kSmemStack][32];
  uint2* sstack = &sstk[threadIdx.x >> 5][0][threadIdx.x & 31];
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float* r = rays + 8 * i;
    HitRec h;
    unsigned a = 0, b = 0;
    V3 o(r[0], r[1], r[2]), d(r[4], r[5], r[6]);
    if (anyHit) {
      out[i] = traverse<true, false, true>(P.sc, o, d, r[3], &h, sstack, &a, &b) ? 1 : 0;
    } else {
      traverse<false, false, true>(P.sc, o, d, r[3], &h, sstack, &a, &b);
      int* q = out + 4 * i;
      q[0] = h.tri >= 0 ? __float_as_int(__ldg(&P.sc.tv[3 * h.tri].w)) : -1;
      q[1] = __float_as_int(h.t);
      q[2] = h.tri >= 0 ? __float_as_int(h.b0) : 0;
      q[3] = h.tri >= 0 ? __float_as_int(h.b1) : 0;
    }
  }
}


// Read bandwidth of a buffer that is re-read `iters` times (pbrgpu_measure_read_bandwidth): with a buffer smaller than the
// L2 this is the L2 -> SM bandwidth the traversal kernel's node / triangle fetches are bounded by (SURVEY 8(d) regime (i));
// with a multi-GB buffer it is the HBM read bandwidth.  16-byte loads that bypass the L1 (ld.global.cg), 8 in flight per thread.
__global__ void __launch_bounds__(256) k_read_bandwidth(const float4* __restrict__ p, size_t n, int iters, float* sink) {
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (int it = 0; it < iters; ++it) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 7 * stride < n; i += 8 * stride) {
      float4 v[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = __ldcg(p + i + k * stride);
#pragma unroll
      for (int k = 0; k < 8; ++k) { acc.x += v[k].x; acc.y += v[k].y; acc.z += v[k].z; acc.w += v[k].w; }
    }
    for (; i < n; i += stride) { const float4 v = __ldcg(p + i); acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w; }
  }
  if (acc.x + acc.y + acc.z + acc.w == 123456.789f) *sink = acc.x;  // never true: keeps the loads alive
}

// Parity hook: Light::Sample_Li / Pdf_Li of one light for n reference points.  rows float[n][16] = ref.p, ref.pError, ref.n,
// u0, u1, query wi, -, -;  out float[n][16] = Li.rgb, pdf, wi, p1, n1, Pdf_Li(ref, query wi), -, -.
__global__ void __launch_bounds__(kBlock) k_light_eval(SceneDev sc, int light, const float* rows, long long n, float* out) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float* r = rows + 16 * i;
    const V3 p(r[0], r[1], r[2]), pe(r[3], r[4], r[5]), nn(r[6], r[7], r[8]), wq(r[11], r[12], r[13]);
    const pbrgpu_light& l = sc.lights[light];
    LightSample ls;
    ls.wi = V3(0, 0, 0); ls.p1 = V3(0, 0, 0);
    sample_light(sc, l, p, r[9], r[10], &ls, pe, nn);
    const V3 spawnO = offset_ray_origin(p, pe, nn, wq);  // ref.SpawnRay(wi).o
    const float pq = light_pdf_li(sc, l, spawnO, p, wq, pe, nn);
    float* o = out + 16 * i;
    o[0] = ls.Li.r; o[1] = ls.Li.g; o[2] = ls.Li.b; o[3] = ls.pdf;
    o[4] = ls.wi.x; o[5] = ls.wi.y; o[6] = ls.wi.z;
    o[7] = ls.p1.x; o[8] = ls.p1.y; o[9] = ls.p1.z;
    o[10] = ls.n1.x; o[11] = ls.n1.y; o[12] = ls.n1.z;
    o[13] = pq; o[14] = o[15] = 0.f;
  }
}

// ---- parity hooks on the PRODUCTION traversal kernel (k_trace) ----
// PBRGPU_DUMP_WAVE_HITS: after k_wave_begin -> k_raygen -> k_trace, the hit records of the wave in the layout of
// PBRGPU_DUMP_PRIMARY_HITS ([pixel][sample][prim, t, b0, b1]).
__global__ void __launch_bounds__(kBlock) k_collect_wave_hits(const __grid_constant__ RenderParams P, int ns, int* out) {
  const uint32_t nPix = (uint32_t)(P.w * P.h);
  const uint32_t n = nPix * (uint32_t)ns;
  for (uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x; slot < n; slot += gridDim.x * blockDim.x) {
    const uint32_t j = slot / nPix, p = slot - j * nPix;
    const float4 h = P.ws.hit[slot];
    const int tri = __float_as_int(h.x);
    int* q = out + 4 * ((size_t)p * ns + j);
    q[0] = tri >= 0 ? __float_as_int(__ldg(&P.sc.tv[3 * tri].w)) : -1;
    q[1] = __float_as_int(P.ws.cur[slot].o.w);  // k_trace stored the hit distance (RenderParams::storeHitT), tMax on a miss
    q[2] = tri >= 0 ? __float_as_int(h.y) : 0;
    q[3] = tri >= 0 ? __float_as_int(h.z) : 0;
  }
}

// pbrgpu_trace_rays through k_trace's own queues: closest-hit rays enter BOTH the path queue and the MIS queue, any-hit
// rays the shadow queue (contribution 1 -> L[i].x == 1 means "unoccluded").
__global__ void __launch_bounds__(kBlock) k_feed_rays(const __grid_constant__ RenderParams P, const float* rays, uint32_t n, int anyHit) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float* r = rays + 8 * (size_t)i;
    const float4 o = make_float4(r[0], r[1], r[2], r[3]), d = make_float4(r[4], r[5], r[6], 0.f);
    P.ws.L[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (anyHit) {
      P.ws.shO[i] = o;
      P.ws.shD[i] = make_float4(d.x, d.y, d.z, __uint_as_float(i));
      P.ws.shC[i] = make_float4(1.f, 0.f, 0.f, 0.f);
    } else {
      store_path(P.ws.cur + i, o, d, make_float4(1.f, 1.f, 1.f, 1.f), 0u, 0u, 5u, i);
      P.ws.miO[i] = o;
      P.ws.miD[i] = make_float4(d.x, d.y, d.z, __uint_as_float(i));
      P.ws.miC[i] = make_float4(0.f, 0.f, 0.f, __int_as_float(0));
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    P.ws.ctr->nCur = anyHit ? 0u : n;
    P.ws.ctr->nMis = anyHit ? 0u : n;
    P.ws.ctr->nShadow = anyHit ? n : 0u;
  }
}
__global__ void __launch_bounds__(kBlock) k_collect_rays(const __grid_constant__ RenderParams P, uint32_t n, int anyHit, int* out, int* nDisagree) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    if (anyHit) { out[i] = P.ws.L[i].x == 1.f ? 0 : 1; continue; }
    const float4 h = P.ws.hit[i];
    const int tri = __float_as_int(h.x);
    if (__float_as_int(P.ws.miO[i].w) != tri) atomicAdd(nDisagree, 1);  // the MIS queue must find the same triangle
    int* q = out + 4 * (size_t)i;
    q[0] = tri >= 0 ? __float_as_int(__ldg(&P.sc.tv[3 * tri].w)) : -1;
    q[1] = __float_as_int(P.ws.cur[i].o.w);
    q[2] = tri >= 0 ? __float_as_int(h.y) : 0;
    q[3] = tri >= 0 ? __float_as_int(h.z) : 0;
  }
}

}  // namespace pbr
