// kernels.cuh - the wavefront kernels (sm_100a).  One bounce of the wavefront maps onto
// PathIntegrator::Li's loop body (Integrator/PathIntegrator.cpp:42-108):
//   k_trace (path)    scene.Intersect(ray)                                   :44
//   k_classify        bins the hits into material-sorted shading queues
//   k_shade<MAT>      emission :47-51, maxDepth :60, BSDF :63, UniformSampleOneLight :74-81,
//                     Sample_f :88-99, Russian roulette :100-107
//   k_shade_miss      infinite-light Le for escaped rays                     :52-56
//   k_trace (shadow)  VisibilityTester::Unoccluded (Integrator.cpp:107, Light.cpp:19-22)
//   k_trace (MIS) + k_resolve_mis   EstimateDirect's BSDF-sampled closest-hit ray (Integrator.cpp:157-172)
// k_trace is ONE persistent kernel: warps pull rays from a device queue and replace finished rays
// while the rest of the warp keeps traversing (ballot / shfl aggregated fetch).
// All paths of a wave are at the same bounce, so `bounce` is a launch argument, not path state.
//
// VolPathIntegrator::Li (Integrator/VolPathIntegrator.cpp:21-107) runs on the same wavefront with VOL = true
// instantiations: rays carry a medium id, k_trace stores the hit distance, k_medium_sample draws
// Medium::Sample per segment, two more queues (surface without material, scattering inside a medium) are shaded
// by k_shade_null / k_shade_medium, shadow rays march through medium boundaries inside k_trace
// (VisibilityTester::Tr), and the bounce count is path state (a boundary crossing does not count as a bounce).
#pragma once
#include "shade.cuh"
#include "media.cuh"

namespace pbr {

constexpr int kBlock = 256;
constexpr int kShadeBlock = 128;  // shading kernels: ~100 registers/thread, 128-thread blocks pack the register file better

PBR_D V3 xform_point(const float* m, V3 p) {  // Transform::operator()(Point3) Core/Transform.h:138-150
  float x = p.x, y = p.y, z = p.z;
  float xp = m[0] * x + m[1] * y + m[2] * z + m[3];
  float yp = m[4] * x + m[5] * y + m[6] * z + m[7];
  float zp = m[8] * x + m[9] * y + m[10] * z + m[11];
  float wp = m[12] * x + m[13] * y + m[14] * z + m[15];
  if (wp == 1) return V3(xp, yp, zp);
  float inv = 1.f / wp;
  return V3(inv * xp, inv * yp, inv * zp);
}

// Camera::GenerateRay* (o, d only: differentials are dead code on this path)
PBR_D void generate_ray(const CameraDev& cam, float fx, float fy, float lensU, float lensV, V3* o, V3* d) {
  V3 pCamera = xform_point(cam.r2c, V3(fx, fy, 0));
  V3 ro, rd;
  if (cam.type == PBRGPU_CAMERA_PERSPECTIVE) {
    ro = V3(0, 0, 0);
    rd = normalize(pCamera);
  } else {
    ro = pCamera;
    rd = V3(0, 0, 1);
  }
  if (cam.lensRadius > 0) {
    float lx, ly;
    concentric_sample_disk(lensU, lensV, &lx, &ly);
    lx = cam.lensRadius * lx; ly = cam.lensRadius * ly;
    float ft = cam.focalDistance / rd.z;
    V3 pFocus = ro + rd * ft;
    ro = V3(lx, ly, 0);
    rd = normalize(pFocus - ro);
  }
  *o = xform_point(cam.c2w, ro);
  *d = xform_vec(cam.c2w, rd);
}

PBR_D void slot_to_pixel(const RenderParams& P, uint32_t slot, int* px, int* py, long long* s, uint32_t* pix) {
  uint32_t nPix = (uint32_t)(P.w * P.h);
  uint32_t j = slot / nPix, p = slot - j * nPix;
  *pix = p;
  *py = P.y0 + (int)(p / (uint32_t)P.w);
  *px = P.x0 + (int)(p % (uint32_t)P.w);
  *s = P.sBegin + j;
}

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
    P.ws.rayO[slot] = make_float4(o.x, o.y, o.z, kInfinity);
    P.ws.rayD[slot] = make_float4(d.x, d.y, d.z, P.integrator == PBRGPU_INTEGRATOR_VOLPATH ? __int_as_float(P.sc.cam.medium) : 0.f);
    P.ws.beta[slot] = make_float4(1.f, 1.f, 1.f, 1.f);
    P.ws.L[slot] = make_float4(0.f, 0.f, 0.f, 0.f);
    P.ws.hidx[slot] = make_uint2((uint32_t)((uint64_t)idx & 0xffffffffu), (uint32_t)((uint64_t)idx >> 32));
    // next dimension; specularBounce = false; VolPath: bounces = 0, in-medium flag
    P.ws.meta[slot] = 5u | ((P.integrator == PBRGPU_INTEGRATOR_VOLPATH && P.sc.cam.medium >= 0) ? (1u << kMetaInMediumBit) : 0u);
    P.ws.qCur[slot] = slot;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) P.ws.ctr->nCur = n;
}

// The offset rays of a camera sample (rxOrigin/rxDirection, ryOrigin/ryDirection of GenerateRayDifferential for a pinhole
// camera, Perspective.cpp:72-78 / Orthographic.cpp:54-58), transformed to world space and scaled by
// RayDifferential::ScaleDifferentials(1 / sqrt(spp)) as the render loop does (Integrator.cpp:305).
PBR_D void camera_differentials(const RenderParams& P, float fx, float fy, V3 o, V3 d, V3* rxO, V3* rxD, V3* ryO, V3* ryD) {
  const CameraDev& cam = P.sc.cam;
  const V3 pCamera = xform_point(cam.r2c, V3(fx, fy, 0));
  V3 xo, xd, yo, yd;
  if (cam.type == PBRGPU_CAMERA_PERSPECTIVE) {
    const V3 p00 = xform_point(cam.r2c, V3(0, 0, 0));
    const V3 dxCamera = xform_point(cam.r2c, V3(1, 0, 0)) - p00, dyCamera = xform_point(cam.r2c, V3(0, 1, 0)) - p00;
    xo = yo = V3(0, 0, 0);
    xd = normalize(pCamera + dxCamera);
    yd = normalize(pCamera + dyCamera);
  } else {
    const V3 dxCamera = xform_vec(cam.r2c, V3(1, 0, 0)), dyCamera = xform_vec(cam.r2c, V3(0, 1, 0));
    xo = pCamera + dxCamera;
    yo = pCamera + dyCamera;
    xd = yd = V3(0, 0, 1);
  }
  xo = xform_point(cam.c2w, xo); yo = xform_point(cam.c2w, yo);
  xd = xform_vec(cam.c2w, xd); yd = xform_vec(cam.c2w, yd);
  const float s = P.diffScale;
  *rxO = o + (xo - o) * s; *ryO = o + (yo - o) * s;
  *rxD = d + (xd - d) * s; *ryD = d + (yd - d) * s;
}

// Kd->Evaluate(si) for a Matte / Plastic material whose Kd is an ImageTexture: uv and dpdu/dpdv of the hit
// (Triangle.cpp:147-170), ComputeDifferentials for camera rays (every other ray has no differentials: Interaction.cpp:91-97).
PBR_D void hit_uv_differentials(const RenderParams& P, const TriData& tri, int triId, float b0, float b1, float b2, V3 p, V3 n,
                                bool cameraRay, int px, int py, int64_t index, V3 rayO, V3 rayD, float* u, float* v, UvDiff* dd) {
  const SceneDev& sc = P.sc;
  float uv[6] = {0.f, 0.f, 1.f, 0.f, 1.f, 1.f};  // Triangle::GetUVs default (Triangle.h:50-54)
  if ((tri.w1 >> kTriFlagShift) & kTriFlagUVs) {
    const float4 a = __ldg(sc.tuv0 + triId), b = __ldg(sc.tuv1 + triId);
    uv[0] = a.x; uv[1] = a.y; uv[2] = a.z; uv[3] = a.w; uv[4] = b.x; uv[5] = b.y;
  }
  *u = (b0 * uv[0] + b1 * uv[2]) + b2 * uv[4];
  *v = (b0 * uv[1] + b1 * uv[3]) + b2 * uv[5];
  *dd = UvDiff{0.f, 0.f, 0.f, 0.f};
  if (cameraRay) {
    V3 dpdu, dpdv;
    triangle_dpdu_dpdv(tri.p0, tri.p1, tri.p2, uv, &dpdu, &dpdv);
    const float fx = (float)px + sampler_film_sample(index, 0, px, sc.hp);
    const float fy = (float)py + sampler_film_sample(index, 1, py, sc.hp);
    V3 rxO, rxD, ryO, ryD;
    camera_differentials(P, fx, fy, rayO, rayD, &rxO, &rxD, &ryO, &ryD);
    *dd = compute_differentials(p, n, dpdu, dpdv, rxO, rxD, ryO, ryD);
  }
}

PBR_D Spec eval_kd_texture(const RenderParams& P, const pbrgpu_material& m, const TriData& tri, int triId, float b0, float b1, float b2,
                           const Isect& is, bool cameraRay, uint32_t slot, int64_t index, V3 rayO, V3 rayD) {
  int px, py; long long sIdx; uint32_t pix;
  slot_to_pixel(P, slot, &px, &py, &sIdx, &pix);
  float u, v;
  UvDiff dd;
  hit_uv_differentials(P, tri, triId, b0, b1, b2, is.p, is.n, cameraRay, px, py, index, rayO, rayD, &u, &v, &dd);
  return tex_evaluate(P.sc.textures[m.kd_texture_plus1 - 1], P.sc.ewaLut, u, v, dd);
}

// warp-aggregated append: one atomicAdd per (warp, queue)
PBR_D uint32_t queue_push(uint32_t* counter, bool active) {
  unsigned mask = __ballot_sync(__activemask(), active);
  if (!active) return 0;
  int lane = threadIdx.x & 31;
  int leader = __ffs(mask) - 1;
  uint32_t base = 0;
  if (lane == leader) base = atomicAdd(counter, (uint32_t)__popc(mask));
  base = __shfl_sync(mask, base, leader);
  return base + __popc(mask & ((1u << lane) - 1u));
}

// K2/K3/K6: the traversal kernel.  Persistent warps; each lane owns one ray at a time.  When at least
// P.fetchMinIdle lanes of a warp have finished their rays, the idle lanes fetch new ones with a single
// warp-aggregated atomicAdd on the queue cursor (__ballot_sync / __shfl_sync), so a warp is not held
// hostage by its longest ray.  Per iteration a lane performs one node step (80-byte node, 8 child
// boxes) and then drains the triangle group that step produced.
// One launch per bounce serves three queues at once: the path rays of bounce b and the shadow / MIS
// rays emitted by the shading of bounce b-1 (independent work, so the machine stays full).
enum { TRACE_PATH = 0, TRACE_SHADOW = 1, TRACE_MIS = 2 };
constexpr int kTraceBlock = 128;

// Triangle batching: a node step leaves a lane with 0..24 candidate triangles.  Draining them right away
// (a while loop) would run the long watertight test with only the few lanes that happen to have several
// candidates, so a lane parks unfinished triangle groups in a small shared-memory list and keeps
// descending; every loop iteration tests at most ONE triangle per lane.
#ifndef PBR_TRI_QUEUE
#define PBR_TRI_QUEUE 6
#endif
constexpr int kTriQueue = PBR_TRI_QUEUE;
#ifndef PBR_NODE_STEPS
#define PBR_NODE_STEPS 1
#endif
constexpr int kNodeStepsPerIter = PBR_NODE_STEPS;

#ifndef PBR_TRACE_MIN_BLOCKS
#define PBR_TRACE_MIN_BLOCKS 8
#endif
#ifndef PBR_SHADE_MIN_BLOCKS
#define PBR_SHADE_MIN_BLOCKS 6
#endif
// VisibilityTester::Tr (Light/Light.cpp:31-47) for one finished segment of a VolPath shadow ray: returns true when the
// ray continues behind a medium boundary (state re-initialised in T), false when it is done (blocked, or arrived and
// its contribution x transmittance added to L).
constexpr int kMaxShadowCrossings = 256;
PBR_D bool vol_shadow_next_segment(const RenderParams& P, uint32_t ray, Trav& T, const HitRec& hit, int* hops) {
  const SceneDev& sc = P.sc;
  if (hit.tri >= 0 && ((hit.w1 >> kTriTypeShift) & 0xfu) != (uint32_t)kQueueNull) return false;  // opaque surface: Tr = 0
  float4 c = P.ws.shC[ray];
  const float4 p1m = P.ws.shP1[ray], dd = P.ws.shD[ray];
  const int med = __float_as_int(p1m.w);
  const V3 d(dd.x, dd.y, dd.z);
  if (med >= 0) {  // ray.medium->Tr(ray): ray.tMax is the hit distance after Intersect, else 1 - ShadowEpsilon
    const Spec tr = medium_Tr(sc.media[med], hit.t, length(d));
    c.x *= tr.r; c.y *= tr.g; c.z *= tr.b;
  }
  if (hit.tri < 0) {
    const uint32_t slot = __float_as_uint(dd.w);
    atomicAdd(&P.ws.L[slot].x, c.x); atomicAdd(&P.ws.L[slot].y, c.y); atomicAdd(&P.ws.L[slot].z, c.z);
    return false;
  }
  if (++*hops > kMaxShadowCrossings) { atomicAdd(&P.ws.ctr->volCapHits, 1u); return false; }
  if (P.countStats) atomicMax(&P.ws.ctr->volMaxHops, (uint32_t)*hops);
  // ray = isect.SpawnRayTo(p1)   (Interaction.h:38-43; isect.mediumInterface from GeometricPrimitive::Intersect)
  const TriData tri = load_tri(sc, hit.tri);
  Isect is;
  make_isect(sc, tri, hit.tri, hit.b0, hit.b1, hit.b2, d, &is);
  const float4 e1 = P.ws.shE1[ray], n1 = P.ws.shN1[ray];
  const V3 p1(p1m.x, p1m.y, p1m.z);
  const V3 origin = offset_ray_origin(is.p, is.pError, is.n, p1 - is.p);
  const V3 target = offset_ray_origin(p1, V3(e1.x, e1.y, e1.z), V3(n1.x, n1.y, n1.z), origin - p1);
  const V3 nd = target - origin;
  int mIn, mOut;
  isect_media(sc, hit.tri, med, &mIn, &mOut);
  const int nmed = dot(nd, is.n) > 0 ? mOut : mIn;
  P.ws.shC[ray] = c;
  P.ws.shP1[ray].w = __int_as_float(nmed);
  P.ws.shD[ray] = make_float4(nd.x, nd.y, nd.z, dd.w);
  trav_init(T, origin, nd, 1 - kShadowEpsilon, sc.f32one);
  return true;
}

template <bool COUNT, bool VOL = false>
__global__ void __launch_bounds__(kTraceBlock, PBR_TRACE_MIN_BLOCKS) k_trace(const __grid_constant__ RenderParams P) {
  __shared__ uint2 sstk[kTraceBlock / 32][kSmemStack][32];
  __shared__ uint2 stri[kTraceBlock / 32][kTriQueue][32];
  __shared__ HitRec shit[kTraceBlock];  // closest hit so far of each lane's ray (see cw_tri_step)
  HitRec& hit = shit[threadIdx.x];
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint2* sstack = &sstk[warp][0][lane];
  uint2* tq = &stri[warp][0][lane];
  Counters* C = P.ws.ctr;
  const uint32_t nPath = C->nCur, nShadow = C->nShadow, nMis = C->nMis;
  const uint32_t n = nPath + nShadow + nMis;
  Trav T;
  TravSpill S;
  T.ngroup = make_uint2(0u, 0u); T.tgroup = make_uint2(0u, 0u); T.sp = 0;
  int nt = 0;  // parked triangle groups
  bool busy = false, exhausted = false;
  int mode = TRACE_PATH;
  uint32_t ray = 0;  // PATH: wavefront slot; SHADOW / MIS: index in that queue
  uint32_t qi = 0;   // PATH: position in the path queue
  int hops = 0;      // VOL: medium boundaries this shadow ray has crossed
  unsigned segs = 0; // VOL: shadow-ray segments traced by this lane (each is one Scene::Intersect of VisibilityTester::Tr)
  unsigned nodes = 0, tris = 0;
  unsigned rayWork0 = 0;  // COUNT: nodes + tris at the start of the current ray (-> Counters::maxRayWork)
  while (true) {
    // ---- fetch: idle lanes take new rays with one warp-aggregated atomicAdd ----
    // (a per-warp reserve of 128 rays per atomicAdd was measured: no faster - the cursor is not a bottleneck)
    const unsigned idle = __ballot_sync(full, !busy);
    if (idle != 0u && !exhausted && (idle == full || __popc(idle) >= P.fetchMinIdle)) {
      const int nIdle = __popc(idle), leader = __ffs(idle) - 1;
      uint32_t base = 0;
      if (lane == leader) base = atomicAdd(&C->cursor[0], (uint32_t)nIdle);
      base = __shfl_sync(full, base, leader);
      exhausted = base + (uint32_t)nIdle >= n;
      if (!busy) {
        uint32_t i = base + __popc(idle & ((1u << lane) - 1u));
        if (i < n) {
          float4 ro, rd;
          if (i < nPath) { mode = TRACE_PATH; qi = i; ray = P.ws.qCur[i]; ro = P.ws.rayO[ray]; rd = P.ws.rayD[ray]; }
          else if (i < nPath + nShadow) { mode = TRACE_SHADOW; ray = i - nPath; ro = P.ws.shO[ray]; rd = P.ws.shD[ray]; }
          else { mode = TRACE_MIS; ray = i - nPath - nShadow; ro = P.ws.miO[ray]; rd = P.ws.miD[ray]; }
          trav_init(T, V3(ro.x, ro.y, ro.z), V3(rd.x, rd.y, rd.z), ro.w, P.sc.f32one);
          nt = 0;
          hops = 0;
          busy = true;
          hit.tri = -1;
          if (COUNT) rayWork0 = nodes + tris;
        }
      }
    }
    if (exhausted && __ballot_sync(full, busy) == 0u) break;
    // ---- node phase: kNodeStepsPerIter x (pop if the node group is exhausted, then one node step).  More than one
    // node step per triangle round was measured (2: +4 % time, 3: +19 %): the extra speculative node visits cost more
    // than the fuller triangle rounds save, so kNodeStepsPerIter = 1 ----
    if (busy) {
#pragma unroll 1
      for (int rep = 0; rep < kNodeStepsPerIter; ++rep) {
        if (T.ngroup.y <= 0x00FFFFFFu && T.sp > 0) T.ngroup = trav_pop(T, sstack, S);
        if (T.ngroup.y > 0x00FFFFFFu && (T.tgroup.y == 0u || nt < kTriQueue)) {
          if (T.tgroup.y != 0u) { tq[nt * 32] = T.tgroup; nt++; }  // park the unfinished triangle group
          cw_node_step(P.sc.nodes8, T, sstack, S);
          if (COUNT) nodes++;
        }
      }
      // ---- triangle round: ONE candidate triangle per lane (a threshold on the number of lanes with candidates
      // was measured: lanes always have a backlog, so the round runs every iteration anyway) ----
      if (T.tgroup.y != 0u || nt > 0) {
        if (T.tgroup.y == 0u) { nt--; T.tgroup = tq[nt * 32]; }
        if (COUNT) tris++;
        // any-hit exit for shadow rays; VolPath shadow rays need the CLOSEST surface (it may be a medium boundary)
        // (unless the scene has no such boundary at all: then the first surface found ends the ray either way)
        if (cw_tri_step(P.sc, T, &hit) && mode == TRACE_SHADOW && (!VOL || P.volShadowAnyHit)) { T.tgroup.y = 0u; nt = 0; T.ngroup.y = 0u; T.sp = 0; }
      }
      // ---- ray finished?  (checked in the same iteration, so the lane is idle for the next fetch) ----
      if (T.ngroup.y <= 0x00FFFFFFu && T.sp == 0 && T.tgroup.y == 0u && nt == 0) {
        bool again = false;
        if (COUNT) { atomicMax(&C->maxRayWork, nodes + tris - rayWork0); rayWork0 = nodes + tris; }
        if (mode == TRACE_PATH) {
          P.ws.hit[ray] = make_float4(__int_as_float(hit.tri), hit.b0, hit.b1, hit.b2);
          // shading-queue id of this path, written in QUEUE order so that k_classify reads it coalesced
          P.ws.qType[qi] = hit.tri >= 0 ? ((hit.w1 >> kTriTypeShift) & 0xfu) : (uint32_t)kQueueMiss;
          if (VOL) P.ws.rayO[ray].w = T.tMax;  // ray.tMax after Scene::Intersect (hit distance, else unchanged), read by Medium::Sample
        } else if (mode == TRACE_MIS) {
          P.ws.miO[ray].w = __int_as_float(hit.tri);
        } else if (VOL) {
          segs++;
          hit.t = T.tMax;
          again = vol_shadow_next_segment(P, ray, T, hit, &hops);  // true: next segment, same lane
          if (again) hit.tri = -1;
        } else if (hit.tri < 0) {  // unoccluded: add the light-sampling contribution
          const uint32_t slot = __float_as_uint(P.ws.shD[ray].w);
          const float4 c = P.ws.shC[ray];
          if (P.atomicL) {
            atomicAdd(&P.ws.L[slot].x, c.x); atomicAdd(&P.ws.L[slot].y, c.y); atomicAdd(&P.ws.L[slot].z, c.z);
          } else {
            float4 L = P.ws.L[slot];
            P.ws.L[slot] = make_float4(L.x + c.x, L.y + c.y, L.z + c.z, 0.f);
          }
        }
        busy = again;
      }
    }
  }
  if (COUNT) {
    atomicAdd(&C->nodesVisited, (unsigned long long)nodes);
    atomicAdd(&C->trisTested, (unsigned long long)tris);
  }
  if (VOL) {  // VolPath has no IntersectP queries: every shadow segment is a closest-hit query
    segs = __reduce_add_sync(full, segs);
    if (lane == 0 && segs) atomicAdd(&C->raysClosest, (unsigned long long)segs);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    if (!VOL) atomicAdd(&C->raysShadow, (unsigned long long)nShadow);
    atomicAdd(&C->raysClosest, (unsigned long long)(nPath + nMis));
    atomicAdd(&C->raysPath, (unsigned long long)nPath);
  }
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
    P.ws.ctr->nShadow = 0;
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
        slot[k] = P.ws.qCur[i];
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

// Block-aggregated queue append for the shading kernels: every thread of the block calls it (uniform
// control flow); one global atomicAdd per block per call.
template <int BLOCK>
PBR_D uint32_t block_queue_push(uint32_t* counter, bool active, uint32_t* sCount, uint32_t* sBase) {
  if (threadIdx.x == 0) *sCount = 0;
  __syncthreads();
  const unsigned mask = __ballot_sync(0xffffffffu, active);
  const int lane = threadIdx.x & 31;
  uint32_t warpBase = 0;
  if (lane == 0 && mask) warpBase = atomicAdd(sCount, (uint32_t)__popc(mask));
  warpBase = __shfl_sync(0xffffffffu, warpBase, 0);
  __syncthreads();
  if (threadIdx.x == 0) *sBase = *sCount ? atomicAdd(counter, *sCount) : 0u;
  __syncthreads();
  return *sBase + warpBase + __popc(mask & ((1u << lane) - 1u));
}

// Three appends (shadow, MIS, next-bounce queues) with ONE round of barriers: per-warp ballots, one shared-memory atomicAdd
// per (warp, queue), three threads do the three global atomicAdds.  (Three separate block_queue_push calls cost nine
// __syncthreads per loop iteration: ncu attributed 9 % of k_shade<Matte>'s stall cycles to the barrier.)
#ifndef PBR_PUSH3
#define PBR_PUSH3 1
#endif
PBR_D void block_queue_push3(uint32_t* c0, uint32_t* c1, uint32_t* c2, bool a0, bool a1, bool a2, uint32_t* sCount3, uint32_t* sBase3,
                             uint32_t* p0, uint32_t* p1, uint32_t* p2) {
  if (threadIdx.x < 3) sCount3[threadIdx.x] = 0;
  __syncthreads();
  const unsigned m0 = __ballot_sync(0xffffffffu, a0), m1 = __ballot_sync(0xffffffffu, a1), m2 = __ballot_sync(0xffffffffu, a2);
  const int lane = threadIdx.x & 31;
  uint32_t w0 = 0, w1 = 0, w2 = 0;
  if (lane == 0) {
    if (m0) w0 = atomicAdd(&sCount3[0], (uint32_t)__popc(m0));
    if (m1) w1 = atomicAdd(&sCount3[1], (uint32_t)__popc(m1));
    if (m2) w2 = atomicAdd(&sCount3[2], (uint32_t)__popc(m2));
  }
  w0 = __shfl_sync(0xffffffffu, w0, 0); w1 = __shfl_sync(0xffffffffu, w1, 0); w2 = __shfl_sync(0xffffffffu, w2, 0);
  __syncthreads();
  if (threadIdx.x < 3) {
    uint32_t* c = threadIdx.x == 0 ? c0 : (threadIdx.x == 1 ? c1 : c2);
    const uint32_t n = sCount3[threadIdx.x];
    sBase3[threadIdx.x] = n ? atomicAdd(c, n) : 0u;
  }
  __syncthreads();
  const unsigned below = (1u << lane) - 1u;
  *p0 = sBase3[0] + w0 + __popc(m0 & below);
  *p1 = sBase3[1] + w1 + __popc(m1 & below);
  *p2 = sBase3[2] + w2 + __popc(m2 & below);
}

// K4/K5/K7: shading for one material queue.
// VOL = true is VolPathIntegrator::Li's surface branch (VolPathIntegrator.cpp:56-93): the bounce count comes from
// the path (`bounce` is ignored), UniformSampleOneLight runs for every BSDF (no NumComponents guard), the shadow
// ray carries what VisibilityTester::Tr needs, and the spawned ray takes the medium on its side of the surface.
// TEX = true (launched only for scenes with image textures): Matte / Plastic Kd may come from an ImageTexture.
template <int MAT, bool VOL = false, bool TEX = false>
__global__ void __launch_bounds__(kShadeBlock, PBR_SHADE_MIN_BLOCKS) k_shade(const __grid_constant__ RenderParams P, int bounce) {
  __shared__ uint32_t sCount, sBase;
  __shared__ uint32_t sCount3[3], sBase3[3];
  const uint32_t n = P.ws.ctr->nMat[MAT];
  const SceneDev& sc = P.sc;
  const uint32_t* queue = P.ws.qMat + (size_t)MAT * P.ws.capacity;
  for (uint32_t base = blockIdx.x * blockDim.x; base < n; base += gridDim.x * blockDim.x) {
    uint32_t i = base + threadIdx.x;
    bool active = i < n;
    bool pushShadow = false, pushMis = false, pushNext = false;
    uint32_t slot = 0;
    float4 shO, shD, shC, miO, miD, miC, shP1, shE1, shN1;
    if (active) {
      slot = queue[i];
      float4 rd4 = P.ws.rayD[slot], h4 = P.ws.hit[slot], be4 = P.ws.beta[slot];
      uint2 hi = P.ws.hidx[slot];
      uint32_t meta = P.ws.meta[slot];
      if (VOL) bounce = (int)((meta >> kMetaBounceShift) & kMetaBounceMask);
      int medIn = -1, medOut = -1;
      if (VOL) isect_media(sc, __float_as_int(h4.x), __float_as_int(rd4.w), &medIn, &medOut);
      V3 rayD(rd4.x, rd4.y, rd4.z);
      Spec beta(be4.x, be4.y, be4.z);
      float etaScale = be4.w;
      bool specularBounce = (meta >> 16) & 1u;
      SamplerCursor cur;
      cur.index = (int64_t)(((uint64_t)hi.y << 32) | hi.x);
      cur.dim = (int)(meta & 0xffffu);
      cur.arrayEndDim = P.arrayEndDim;
      TriData tri = load_tri(sc, __float_as_int(h4.x));
      Isect is;
      make_isect(sc, tri, __float_as_int(h4.x), h4.y, h4.z, h4.w, rayD, &is);
      Spec Ladd(0.f);
      // emitted light for camera rays / after specular bounces
      if ((bounce == 0 || specularBounce) && is.lightId >= 0)
        Ladd = Ladd + beta * area_light_L(sc.lights[is.lightId], is.n, -rayD);
      bool alive = bounce < P.maxDepth;
      if (alive) {
        Bsdf bsdf;
        Spec kdTex;
        bool haveKdTex = false;
        if (TEX && (MAT == PBRGPU_MAT_MATTE || MAT == PBRGPU_MAT_PLASTIC)) {
          const pbrgpu_material& m = sc.mats[is.matId];
          if (m.kd_texture_plus1 > 0) {
            const float4 ro4 = P.ws.rayO[slot];
            // Reference quirk: SurfaceInteraction::ComputeScatteringFunctions takes `const Ray&` (Interaction.h:86-89) and
            // hands it to ComputeDifferentials(const RayDifferential&): the implicit RayDifferential(const Ray&) conversion
            // (Geometry.h:1387) sets hasDifferentials = false, so EVERY texture lookup of a render sees zero derivatives -
            // camera rays included.  The differential code stays reachable through PBRGPU_DUMP_HIT_DIFFERENTIALS, where it is
            // checked against the reference's ComputeDifferentials called on the real RayDifferential.
            const bool cameraRay = false;
            kdTex = eval_kd_texture(P, m, tri, __float_as_int(h4.x), h4.y, h4.z, h4.w, is, cameraRay, slot, cur.index,
                                    V3(ro4.x, ro4.y, ro4.z), rayD);
            haveKdTex = true;
          }
        }
        make_bsdf<MAT>(sc, is, /*allowMultipleLobes=*/true, &bsdf, haveKdTex ? &kdTex : nullptr);
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
            sample_light(sc, light, is.p, uL0, uL1, &ls);
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
                  float lightPdf = light_pdf_li(sc, light, origin, is.p, wi);
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
            P.ws.rayO[slot] = make_float4(o.x, o.y, o.z, kInfinity);
            const int nextMedium = VOL ? (dot(wi, is.n) > 0 ? medOut : medIn) : -1;
            P.ws.rayD[slot] = make_float4(wi.x, wi.y, wi.z, VOL ? __int_as_float(nextMedium) : 0.f);
            P.ws.beta[slot] = make_float4(beta.r, beta.g, beta.b, etaScale);
            P.ws.meta[slot] = (uint32_t)cur.dim | (specularBounce ? (1u << kMetaSpecBit) : 0u) |
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
    if (pushNext) P.ws.qNext[posNext] = slot;
  }
}

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
  sample_light(sc, light, is.p, uL0, uL1, &ls);
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
        float lightPdf = light_pdf_li(sc, light, origin, is.p, wi);
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
    uint32_t slot = 0;
    if (active) {
      slot = queue[i];
      float4 rd4 = P.ws.rayD[slot], h4 = P.ws.hit[slot], be4 = P.ws.beta[slot];
      uint2 hi = P.ws.hidx[slot];
      uint32_t meta = P.ws.meta[slot];
      V3 rayD(rd4.x, rd4.y, rd4.z);
      Spec beta(be4.x, be4.y, be4.z);
      SamplerCursor cur;
      cur.index = (int64_t)(((uint64_t)hi.y << 32) | hi.x);
      cur.dim = (int)(meta & 0xffffu);
      cur.arrayEndDim = P.arrayEndDim;
      TriData tri = load_tri(sc, __float_as_int(h4.x));
      Isect is;
      make_isect(sc, tri, __float_as_int(h4.x), h4.y, h4.z, h4.w, rayD, &is);
      // L += isect.Le(wo) at every depth (WhittedIntegrator.cpp:33, DirectLightingIntegrator.cpp:41)
      if (is.lightId >= 0) {
        Spec Le = beta * area_light_L(sc.lights[is.lightId], is.n, is.wo);
        if (!Le.is_black()) { atomicAdd(&P.ws.L[slot].x, Le.r); atomicAdd(&P.ws.L[slot].y, Le.g); atomicAdd(&P.ws.L[slot].z, Le.b); }
      }
      Bsdf bsdf;
      Spec kdTex;
      bool haveKdTex = false;
      if (TEX && (MAT == PBRGPU_MAT_MATTE || MAT == PBRGPU_MAT_PLASTIC)) {
        const pbrgpu_material& m = sc.mats[is.matId];
        if (m.kd_texture_plus1 > 0) {  // zero derivatives, like every lookup of the reference (see k_shade)
          kdTex = eval_kd_texture(P, m, tri, __float_as_int(h4.x), h4.y, h4.z, h4.w, is, false, slot, cur.index, V3(0, 0, 0), rayD);
          haveKdTex = true;
        }
      }
      make_bsdf<MAT>(sc, is, /*allowMultipleLobes=*/false, &bsdf, haveKdTex ? &kdTex : nullptr);
      if (P.integrator == PBRGPU_INTEGRATOR_WHITTED) {
        for (int j = 0; j < sc.nLights; ++j) {
          float u0, u1;
          cur.get2D(sc.hp, &u0, &u1);
          LightSample ls;
          sample_light(sc, sc.lights[j], is.p, u0, u1, &ls);
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
          const int64_t stride = sc.hp.sampleStride;
          const int64_t s = (cur.index - P.pixelOffsetOf(cur.index)) / stride;  // pixel-sample number
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
                int64_t idx = P.pixelOffsetOf(cur.index) + (s * nS + k) * stride;  // GetIndexForSample(s*n+k)
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
          P.ws.rayO[slot] = make_float4(o.x, o.y, o.z, kInfinity);
          P.ws.rayD[slot] = make_float4(wi.x, wi.y, wi.z, 0.f);
          P.ws.beta[slot] = make_float4(beta.r, beta.g, beta.b, 1.f);
          P.ws.meta[slot] = (uint32_t)cur.dim;
          pushNext = true;
        }
      }
    }
    uint32_t pos = queue_push(&P.ws.ctr->nNext, pushNext);
    if (pushNext) P.ws.qNext[pos] = slot;
  }
}

// ---------------------------------------------------------------------------------------------
// VolPathIntegrator (Integrator/VolPathIntegrator.cpp:21-107), the parts that have no PathIntegrator twin.
// ---------------------------------------------------------------------------------------------
PBR_D SamplerCursor load_cursor(const RenderParams& P, uint32_t slot, uint32_t meta) {
  const uint2 hi = P.ws.hidx[slot];
  SamplerCursor cur;
  cur.index = (int64_t)(((uint64_t)hi.y << 32) | hi.x);
  cur.dim = (int)(meta & 0xffffu);
  cur.arrayEndDim = P.arrayEndDim;
  return cur;
}

// `if (ray.medium) beta *= ray.medium->Sample(ray, sampler, &mi); if (beta.IsBlack()) break;`  (:35-36) for every
// live path whose segment lies in a medium, in two halves around k_trace (media.cuh: medium_sample_distance):
//   k_medium_presample  draws the two sampler values and makes the sampled free-flight distance the ray's tMax
//   k_medium_sample     after the trace: hit below that distance -> surface, weight Tr/pdf; no hit -> MediumInteraction
//                       at that distance, weight Tr*sigma_s/pdf, path re-routed to the medium queue
__global__ void __launch_bounds__(kBlock) k_medium_presample(const __grid_constant__ RenderParams P) {
  const uint32_t n = P.ws.ctr->nCur;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const uint32_t slot = P.ws.qCur[i];
    const uint32_t meta = P.ws.meta[slot];
    if (!((meta >> kMetaInMediumBit) & 1u)) continue;  // vacuum: nothing to sample (4 bytes read)
    const float4 rd4 = P.ws.rayD[slot];
    const int med = __float_as_int(rd4.w);
    SamplerCursor cur = load_cursor(P, slot, meta);
    const float u0 = cur.get1D(P.sc.hp);
    const float u1 = cur.get1D(P.sc.hp);
    P.ws.rayO[slot].w = medium_sample_distance(P.sc.media[med], u0, u1, length(V3(rd4.x, rd4.y, rd4.z)));
    P.ws.meta[slot] = (meta & ~0xffffu) | (uint32_t)cur.dim;
  }
}

__global__ void __launch_bounds__(kBlock) k_medium_sample(const __grid_constant__ RenderParams P) {
  const uint32_t n = P.ws.ctr->nCur;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const uint32_t slot = P.ws.qCur[i];
    if (!((P.ws.meta[slot] >> kMetaInMediumBit) & 1u)) continue;
    const float4 rd4 = P.ws.rayD[slot];
    const int med = __float_as_int(rd4.w);
    const bool hit = __float_as_int(P.ws.hit[slot].x) >= 0;
    const float t = P.ws.rayO[slot].w;           // k_trace: hit distance, or the sampled distance when nothing was hit
    const bool sampled = !hit && t < kInfinity;  // t < ray.tMax
    const Spec w = medium_sample_weight(P.sc.media[med], t, sampled, length(V3(rd4.x, rd4.y, rd4.z)));
    const float4 be4 = P.ws.beta[slot];
    const Spec beta = Spec(be4.x, be4.y, be4.z) * w;
    if (beta.is_black()) { P.ws.qType[i] = kQueueDead; continue; }
    P.ws.beta[slot] = make_float4(beta.r, beta.g, beta.b, be4.w);
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
    uint32_t slot = 0;
    if (i < n) {
      slot = queue[i];
      const float4 rd4 = P.ws.rayD[slot], h4 = P.ws.hit[slot];
      const uint32_t meta = P.ws.meta[slot];
      const int bounce = (int)((meta >> kMetaBounceShift) & kMetaBounceMask);
      const bool specularBounce = (meta >> kMetaSpecBit) & 1u;
      const V3 rayD(rd4.x, rd4.y, rd4.z);
      const int tri = __float_as_int(h4.x);
      const TriData td = load_tri(sc, tri);
      Isect is;
      make_isect(sc, td, tri, h4.y, h4.z, h4.w, rayD, &is);
      if ((bounce == 0 || specularBounce) && is.lightId >= 0) {
        const float4 be4 = P.ws.beta[slot];
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
        P.ws.rayO[slot] = make_float4(o.x, o.y, o.z, kInfinity);
        const int nextMedium = dot(rayD, is.n) > 0 ? mOut : mIn;
        P.ws.rayD[slot] = make_float4(rayD.x, rayD.y, rayD.z, __int_as_float(nextMedium));
        P.ws.meta[slot] = (meta & ~(1u << kMetaInMediumBit)) | (nextMedium >= 0 ? (1u << kMetaInMediumBit) : 0u);
        pushNext = true;
      }
    }
    const uint32_t pos = queue_push(&P.ws.ctr->nNext, pushNext);
    if (pushNext) P.ws.qNext[pos] = slot;
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
    uint32_t slot = 0;
    float4 shO, shD, shC, miO, miD, miC, shP1, shE1, shN1;
    if (i < n) {
      slot = queue[i];
      const float4 ro4 = P.ws.rayO[slot], rd4 = P.ws.rayD[slot], be4 = P.ws.beta[slot];
      const uint32_t meta = P.ws.meta[slot];
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
        SamplerCursor cur = load_cursor(P, slot, meta);
        if (sc.nLights > 0) {
          float pickPdf;
          const int lightNum = sample_discrete(sc.lightFunc, sc.lightCdf, sc.lightFuncInt, sc.nLights, cur.get1D(sc.hp), &pickPdf);
          if (pickPdf != 0) {
            float uL0, uL1, uS0, uS1;
            cur.get2D(sc.hp, &uL0, &uL1);
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
          P.ws.rayO[slot] = make_float4(p.x, p.y, p.z, kInfinity);
          P.ws.rayD[slot] = make_float4(wi.x, wi.y, wi.z, __int_as_float(med));
          P.ws.beta[slot] = make_float4(beta.r, beta.g, beta.b, etaScale);
          P.ws.meta[slot] = (uint32_t)cur.dim | ((uint32_t)(bounce + 1) << kMetaBounceShift) | (1u << kMetaInMediumBit);
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
    if (pushNext) P.ws.qNext[pos] = slot;
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
    uint32_t slot = queue[i];
    const uint32_t meta = P.ws.meta[slot];
    bool specularBounce = (meta >> kMetaSpecBit) & 1u;
    if (volInteg) bounce = (int)((meta >> kMetaBounceShift) & kMetaBounceMask);
    if (pathInteg && !(bounce == 0 || specularBounce)) continue;
    float4 rd = P.ws.rayD[slot], be = P.ws.beta[slot];
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
      if (__float_as_int(__ldg(&P.sc.tv2[tri].w)) == lightId) {
        const float4 n4 = __ldg(P.sc.tvN + tri);
        Li = area_light_L(light, V3(n4.x, n4.y, n4.z), -wi);
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
      traverse<false, false>(P.sc, o, d, kInfinity, &h, sstack, &a, &b);
      float* q = outDiff + 8 * i;
      for (int k = 0; k < 8; ++k) q[k] = 0.f;
      if (h.tri >= 0) {
        const TriData tri = load_tri(P.sc, h.tri);
        Isect is;
        make_isect(P.sc, tri, h.tri, h.b0, h.b1, h.b2, d, &is);
        float u, v;
        UvDiff dd;
        hit_uv_differentials(P, tri, h.tri, h.b0, h.b1, h.b2, is.p, is.n, true, px, py, idx, o, d, &u, &v, &dd);
        q[0] = 1.f; q[1] = u; q[2] = v; q[3] = dd.dudx; q[4] = dd.dvdx; q[5] = dd.dudy; q[6] = dd.dvdy;
      }
    }
    if (outHits) {
      HitRec h;
      unsigned a = 0, b = 0;
      traverse<false, false>(P.sc, o, d, kInfinity, &h, sstack, &a, &b);
      int* q = outHits + 4 * i;
      q[0] = h.tri >= 0 ? __float_as_int(__ldg(&P.sc.tv0[h.tri].w)) : -1;
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
      out[i] = traverse<true, false>(P.sc, o, d, r[3], &h, sstack, &a, &b) ? 1 : 0;
    } else {
      traverse<false, false>(P.sc, o, d, r[3], &h, sstack, &a, &b);
      int* q = out + 4 * i;
      q[0] = h.tri >= 0 ? __float_as_int(__ldg(&P.sc.tv0[h.tri].w)) : -1;
      q[1] = __float_as_int(h.t);
      q[2] = h.tri >= 0 ? __float_as_int(h.b0) : 0;
      q[3] = h.tri >= 0 ? __float_as_int(h.b1) : 0;
    }
  }
}

}  // namespace pbr
