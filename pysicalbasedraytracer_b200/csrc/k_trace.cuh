// k_trace.cuh - the persistent traversal kernel (BVHAccel::Intersect / IntersectP, Accelerator/BVHAccel.cpp:285-366).
#pragma once
#include "kcommon.cuh"

namespace pbr {

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
// VisibilityTester::Tr (Light/Light.cpp:31-47) for one finished segment of a VolPath shadow ray: returns true when the
// ray continues behind a medium boundary (state re-initialised in T), false when it is done (blocked, or arrived and
// its contribution x transmittance added to L).
constexpr int kMaxShadowCrossings = 256;
PBR_D bool vol_shadow_next_segment(const RenderParams& P, uint32_t ray, Trav& T, const HitRec& hit, int* hops) {
  const SceneDev& sc = P.sc;
  const bool toRecord = P.deferShadow == 2;
  if (hit.tri >= 0 && ((hit.w1 >> kTriTypeShift) & 0xfu) != (uint32_t)kQueueNull) {  // opaque surface: Tr = 0
    if (toRecord) P.ws.shC[ray] = make_float4(0.f, 0.f, 0.f, 0.f);
    return false;
  }
  float4 c = P.ws.shC[ray];
  const float4 p1m = P.ws.shP1[ray], dd = P.ws.shD[ray];
  const int med = __float_as_int(p1m.w);
  const V3 d(dd.x, dd.y, dd.z);
  if (med >= 0) {  // ray.medium->Tr(ray): ray.tMax is the hit distance after Intersect, else 1 - ShadowEpsilon
    const Spec tr = medium_Tr(sc.media[med], hit.t, length(d));
    c.x *= tr.r; c.y *= tr.g; c.z *= tr.b;
  }
  if (hit.tri < 0) {
    if (toRecord) { P.ws.shC[ray] = c; return false; }
    const uint32_t slot = __float_as_uint(dd.w);
    atomicAdd(&P.ws.L[slot].x, c.x); atomicAdd(&P.ws.L[slot].y, c.y); atomicAdd(&P.ws.L[slot].z, c.z);
    return false;
  }
  if (++*hops > kMaxShadowCrossings) {
    atomicAdd(&P.ws.ctr->volCapHits, 1u);
    if (toRecord) P.ws.shC[ray] = make_float4(0.f, 0.f, 0.f, 0.f);
    return false;
  }
  if (P.countStats) atomicMax(&P.ws.ctr->volMaxHops, (uint32_t)*hops);
  // ray = isect.SpawnRayTo(p1)   (Interaction.h:38-43; isect.mediumInterface from GeometricPrimitive::Intersect)
  const TriData tri = load_tri(sc, hit.tri);
  Isect is;
  const float4 so = P.ws.shO[ray];  // origin of the segment just traced (an analytic-sphere boundary needs it)
  make_isect(sc, tri, hit.tri, hit.b0, hit.b1, hit.b2, d, &is, V3(so.x, so.y, so.z));
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
  P.ws.shO[ray] = make_float4(origin.x, origin.y, origin.z, 1 - kShadowEpsilon);
  trav_init(T, origin, nd, 1 - kShadowEpsilon, sc.f32one);
  return true;
}

// VolPathIntegrator::Li's loop for a path ray whose segment has just been traced, in the lane that traced it:
//   * `beta *= ray.medium->Sample(ray, sampler, &mi)` (VolPathIntegrator.cpp:35-36): the weight of the segment (the distance
//     was drawn before the trace, media.cuh) - what k_medium_sample does as a separate pass when this is switched off;
//   * closest hit on a surface WITHOUT a material, i.e. a medium boundary (:65-70): `ray = isect.SpawnRay(ray.d); bounces--;
//     continue;` including the free-flight sample of the next segment (k_shade_null + k_medium_presample).
// A path that crosses the ripples of a boundary mesh dozens of times stays in its lane instead of costing one wavefront
// iteration (trace + classify + shade launches and a host read-back) per crossing.  Same arithmetic, same sampler
// dimensions, same order as the separate kernels.  *q = shading queue of the path when it leaves the lane; returns true when
// the lane continues with the re-spawned ray (*no, *nTMax; the direction is unchanged).
PBR_D bool vol_path_next_segment(const RenderParams& P, uint32_t ray, float tEnd, const HitRec& hit, uint32_t* q, V3* no, float* nTMax) {
  const SceneDev& sc = P.sc;
  PathRec* rec = P.ws.cur + ray;
  uint32_t meta = rec->meta;
  const float4 rd4 = rec->d;
  const V3 rayD(rd4.x, rd4.y, rd4.z);
  const int med = __float_as_int(rd4.w);
  if ((meta >> kMetaInMediumBit) & 1u) {
    const bool sampled = hit.tri < 0 && tEnd < kInfinity;  // t < ray.tMax: scattering event before any surface
    const Spec w = medium_sample_weight(sc.media[med], tEnd, sampled, length(rayD));
    const float4 be4 = rec->beta;
    const Spec beta = Spec(be4.x, be4.y, be4.z) * w;
    if (beta.is_black()) { *q = kQueueDead; return false; }
    rec->beta = make_float4(beta.r, beta.g, beta.b, be4.w);
    if (sampled) *q = (uint32_t)kQueueMedium;  // mi.p = ray(t): t goes to rayO.w with the hit record
  }
  if (*q != (uint32_t)kQueueNull) return false;
  const TriData td = load_tri(sc, hit.tri);
  if ((int)td.w2 >= 0) return false;  // an emissive boundary: k_shade_null applies the emission rule
  const int bounce = (int)((meta >> kMetaBounceShift) & kMetaBounceMask);
  if (bounce >= P.maxDepth) { *q = kQueueDead; return false; }  // :60 `if (bounces >= maxDepth) break;`
  const float4 ro4 = rec->o;
  Isect is;
  make_isect(sc, td, hit.tri, hit.b0, hit.b1, hit.b2, rayD, &is, V3(ro4.x, ro4.y, ro4.z));
  int mIn, mOut;
  isect_media(sc, hit.tri, med, &mIn, &mOut);
  const V3 o = offset_ray_origin(is.p, is.pError, is.n, rayD);
  const int nextMedium = dot(rayD, is.n) > 0 ? mOut : mIn;
  float tMax = kInfinity;
  meta &= ~(1u << kMetaInMediumBit);
  if (nextMedium >= 0) {  // Medium::Sample's two sampler values for the next segment, drawn before it is traced
    const int64_t index = (int64_t)(((uint64_t)rec->idxHi << 32) | rec->idxLo);
    int dim = (int)(meta & 0xffffu);
    if (dim >= 5 && dim < P.arrayEndDim) dim = P.arrayEndDim;
    const float u0 = sample_dimension_outofline(index, dim, &sc.hp);
    const float u1 = sample_dimension_outofline(index, dim + 1, &sc.hp);
    tMax = medium_sample_distance(sc.media[nextMedium], u0, u1, length(rayD));
    meta = (meta & ~0xffffu) | (uint32_t)(dim + 2) | (1u << kMetaInMediumBit);
  }
  rec->o = make_float4(o.x, o.y, o.z, tMax);
  rec->d.w = __int_as_float(nextMedium);
  rec->meta = meta;
  *no = o; *nTMax = tMax;
  return true;
}

template <bool COUNT, bool VOL = false, bool SPH = false>
__global__ void __launch_bounds__(kTraceBlock, PBR_TRACE_MIN_BLOCKS) k_trace(const __grid_constant__ RenderParams P) {
  // traversal stack (first kSmemStack levels) and parked triangle groups of every lane: ONE [slot][thread] array, so that a
  // single per-thread base register addresses all of it with immediate offsets (two separate [warp][slot][lane] arrays had
  // the compiler re-derive two base addresses from threadIdx in every loop iteration: ~5 % of the kernel's instructions)
  __shared__ uint2 slane[kSmemStack + kTriQueue][kTraceBlock];
  __shared__ HitRec shit[kTraceBlock];  // closest hit so far of each lane's ray (see cw_tri_step)
  HitRec& hit = shit[threadIdx.x];
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  uint2* sstack = &slane[0][threadIdx.x];
  uint2* tq = sstack + kSmemStack * kTraceBlock;
  Counters* C = P.ws.ctr;
  // VolPath tail (RenderParams::deferShadow): 1 = leave the shadow queue alone, 2 = the shadow queue only
  const uint32_t nPath = (VOL && P.deferShadow == 2) ? 0u : C->nCur, nShadow = (VOL && P.deferShadow == 1) ? 0u : C->nShadow;
  const uint32_t nMis = (VOL && P.deferShadow == 2) ? 0u : C->nMis;
  const uint32_t n = nPath + nShadow + nMis;
  Trav T;
  TravSpill S;
  T.ngroup = make_uint2(0u, 0u); T.tgroup = make_uint2(0u, 0u); T.sp = 0;
  int nt = 0;  // parked triangle groups
  bool busy = false, pend = false, exhausted = false;  // pend: traversal finished, result not yet written
  int mode = TRACE_PATH;
  uint32_t ray = 0;  // index in the queue of `mode`: PATH = position in ws.cur, SHADOW / MIS = position in that ray queue
  int hops = 0;      // VOL: medium boundaries this shadow ray has crossed
  unsigned segs = 0; // VOL: shadow-ray segments traced by this lane (each is one Scene::Intersect of VisibilityTester::Tr)
  unsigned psegs = 0; // VOL: path segments continued in this lane behind a medium boundary (each is one Scene::Intersect of Li's loop)
  unsigned nodes = 0, tris = 0;
  unsigned rayWork0 = 0;  // COUNT: nodes + tris at the start of the current ray (-> Counters::maxRayWork)
  unsigned long long lp[8] = {0, 0, 0, 0, 0, 0, 0, 0};  // COUNT: lane-activity profile (Counters::laneProf), kept by lane 0
  while (true) {
    // ---- completion + fetch, batched: a lane whose ray has finished parks (pend) until at least P.fetchMinIdle lanes of the
    // warp are without traversal work; then the finished rays write their results together and the free lanes take new
    // rays with one warp-aggregated atomicAdd.  (Writing each result the moment its ray finished ran that code - for VolPath
    // the whole of VisibilityTester::Tr's next-segment set-up - with 1-4 active lanes.)
    // (a per-warp reserve of 128 rays per atomicAdd was measured: no faster - the cursor is not a bottleneck)
    const unsigned idle = __ballot_sync(full, !busy);
    if (idle != 0u && (exhausted || idle == full || __popc(idle) >= P.fetchMinIdle)) {
      if (pend) {
        pend = false;
        if (COUNT) { atomicMax(&C->maxRayWork, nodes + tris - rayWork0); rayWork0 = nodes + tris; }
        if (mode == TRACE_PATH) {
          uint32_t q = hit.tri >= 0 ? ((hit.w1 >> kTriTypeShift) & 0xfu) : (uint32_t)kQueueMiss;
          if (VOL && P.volInLane) {
            V3 no;
            float nTMax;
            if (vol_path_next_segment(P, ray, T.tMax, hit, &q, &no, &nTMax)) {  // crossed a medium boundary: next segment, same lane
              const float4 rd = P.ws.cur[ray].d;
              trav_init(T, no, V3(rd.x, rd.y, rd.z), nTMax, P.sc.f32one);
              busy = true; hit.tri = -1; nt = 0;
              psegs++;
            }
          }
          if (!busy) {
            P.ws.hit[ray] = make_float4(__int_as_float(hit.tri), hit.b0, hit.b1, hit.b2);
            // hit record and shading-queue id at the path's queue position
            P.ws.qType[ray] = q;
            if (VOL || P.storeHitT) P.ws.cur[ray].o.w = T.tMax;  // ray.tMax after Scene::Intersect (hit distance, else unchanged), read by Medium::Sample
          }
        } else if (mode == TRACE_MIS) {
          P.ws.miO[ray].w = __int_as_float(hit.tri);
        } else if (VOL) {
          segs++;
          hit.t = T.tMax;
          busy = vol_shadow_next_segment(P, ray, T, hit, &hops);  // true: next segment, same lane
          if (busy) { hit.tri = -1; nt = 0; }
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
      }
      const unsigned want = VOL ? __ballot_sync(full, !busy) : idle;  // VOL: a shadow ray continuing behind a boundary keeps its lane
      if (!exhausted && want != 0u) {
        const int nWant = __popc(want), leader = __ffs(want) - 1;
        uint32_t base = 0;
        if (lane == leader) base = atomicAdd(&C->cursor[0], (uint32_t)nWant);
        base = __shfl_sync(full, base, leader);
        exhausted = base + (uint32_t)nWant >= n;
        if (COUNT) { lp[5]++; lp[6] += (unsigned)nWant; }
        if (!busy) {
          uint32_t i = base + __popc(want & ((1u << lane) - 1u));
          if (i < n) {
            float4 ro, rd;
            if (i < nPath) { mode = TRACE_PATH; ray = i; ro = P.ws.cur[i].o; rd = P.ws.cur[i].d; }  // coalesced: records are in queue order
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
    }
    if (exhausted && __ballot_sync(full, busy) == 0u) break;
    bool didNode = false, didTri = false, triHit = false;
    if (COUNT) { lp[0]++; lp[1] += __popc(__ballot_sync(full, busy)); }
    // ---- node phase: kNodeStepsPerIter x (pop if the node group is exhausted, then one node step).  More than one
    // node step per triangle round was measured (2: +4 % time, 3: +19 %): the extra speculative node visits cost more
    // than the fuller triangle rounds save, so kNodeStepsPerIter = 1 ----
    if (busy) {
#pragma unroll 1
      for (int rep = 0; rep < kNodeStepsPerIter; ++rep) {
        if (T.ngroup.y <= 0x00FFFFFFu && T.sp > 0) T.ngroup = trav_pop<kTraceBlock>(T, sstack, S);
        if (T.ngroup.y > 0x00FFFFFFu && (T.tgroup.y == 0u || nt < kTriQueue)) {
          if (T.tgroup.y != 0u) { tq[nt * kTraceBlock] = T.tgroup; nt++; }  // park the unfinished triangle group
          cw_node_step<kTraceBlock>(P.sc.nodes8, T, sstack, S);
          if (COUNT) { nodes++; didNode = true; }
        }
      }
    }
    // ---- triangle round: ONE candidate triangle per lane, every iteration.  (Skipping the round while fewer than N lanes have
    // a candidate - with and without a "stalled lanes" override - was measured for N = 6 .. 14: never faster, the lanes always
    // have a backlog.) ----
    if (busy) {
      if (T.tgroup.y != 0u || nt > 0) {
        if (T.tgroup.y == 0u) { nt--; T.tgroup = tq[nt * kTraceBlock]; }
        if (COUNT) { tris++; didTri = true; }
        // any-hit exit for shadow rays; VolPath shadow rays need the CLOSEST surface (it may be a medium boundary)
        // (unless the scene has no such boundary at all: then the first surface found ends the ray either way)
        const bool th = cw_tri_step<SPH>(P.sc, T, &hit);
        if (COUNT) triHit = th;
        if (th && mode == TRACE_SHADOW && (!VOL || P.volShadowAnyHit)) { T.tgroup.y = 0u; nt = 0; T.ngroup.y = 0u; T.sp = 0; }
      }
      // ---- ray finished?  (checked in the same iteration, so the lane counts as idle at the next loop head) ----
      if (T.ngroup.y <= 0x00FFFFFFu && T.sp == 0 && T.tgroup.y == 0u && nt == 0) { busy = false; pend = true; }
    }
    if (COUNT) {
      const unsigned mN = __ballot_sync(full, didNode), mT = __ballot_sync(full, didTri), mH = __ballot_sync(full, triHit);
      lp[2] += __popc(mN); lp[3] += __popc(mT); lp[4] += __popc(mH);
      lp[7] += (mT != 0u);  // iterations in which the triangle test ran at all
    }
  }
  if (COUNT) {
    atomicAdd(&C->nodesVisited, (unsigned long long)nodes);
    atomicAdd(&C->trisTested, (unsigned long long)tris);
    if (lane == 0)
      for (int k = 0; k < 8; ++k) atomicAdd(&C->laneProf[k], lp[k]);
  }
  if (VOL) {  // VolPath has no IntersectP queries: every shadow segment is a closest-hit query
    segs = __reduce_add_sync(full, segs);
    psegs = __reduce_add_sync(full, psegs);
    if (lane == 0 && segs + psegs) atomicAdd(&C->raysClosest, (unsigned long long)(segs + psegs));
    if (lane == 0 && psegs) atomicAdd(&C->raysPath, (unsigned long long)psegs);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    if (!VOL) atomicAdd(&C->raysShadow, (unsigned long long)nShadow);
    atomicAdd(&C->raysClosest, (unsigned long long)(nPath + nMis));
    atomicAdd(&C->raysPath, (unsigned long long)nPath);
  }
}


}  // namespace pbr
