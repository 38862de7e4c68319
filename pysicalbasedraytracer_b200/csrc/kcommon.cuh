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

// texPlus1: pbrgpu_material::kd_texture_plus1 or ::ks_texture_plus1 (> 0)
PBR_D Spec eval_image_texture(const RenderParams& P, int texPlus1, const TriData& tri, int triId, float b0, float b1, float b2,
                              const Isect& is, bool cameraRay, uint32_t slot, int64_t index, V3 rayO, V3 rayD) {
  int px, py; long long sIdx; uint32_t pix;
  slot_to_pixel(P, slot, &px, &py, &sIdx, &pix);
  float u, v;
  UvDiff dd;
  hit_uv_differentials(P, tri, triId, b0, b1, b2, is.p, is.n, cameraRay, px, py, index, rayO, rayD, &u, &v, &dd);
  return tex_evaluate(P.sc.textures[texPlus1 - 1], P.sc.ewaLut, u, v, dd);
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

// A path record as four 16-byte loads (the fourth word holds the sampler index, meta and the film slot)
struct PathRegs {
  float4 o, d, beta;
  uint32_t idxLo, idxHi, meta, slot;
};
PBR_D PathRegs load_path(const PathRec* rec) {
  const float4* p = reinterpret_cast<const float4*>(rec);
  PathRegs r;
  r.o = p[0]; r.d = p[1]; r.beta = p[2];
  const float4 w = p[3];
  r.idxLo = __float_as_uint(w.x); r.idxHi = __float_as_uint(w.y); r.meta = __float_as_uint(w.z); r.slot = __float_as_uint(w.w);
  return r;
}
PBR_D void store_path(PathRec* rec, float4 o, float4 d, float4 beta, uint32_t idxLo, uint32_t idxHi, uint32_t meta, uint32_t slot) {
  float4* p = reinterpret_cast<float4*>(rec);
  p[0] = o; p[1] = d; p[2] = beta;
  p[3] = make_float4(__uint_as_float(idxLo), __uint_as_float(idxHi), __uint_as_float(meta), __uint_as_float(slot));
}

PBR_D SamplerCursor load_cursor(const RenderParams& P, const PathRegs& r) {
  SamplerCursor cur;
  cur.index = (int64_t)(((uint64_t)r.idxHi << 32) | r.idxLo);
  cur.dim = (int)(r.meta & 0xffffu);
  cur.arrayEndDim = P.arrayEndDim;
  return cur;
}


}  // namespace pbr
