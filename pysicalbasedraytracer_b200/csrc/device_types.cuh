// device_types.cuh - HBM-resident layout of the flattened scene and of the wavefront state.
#pragma once
#include "../../include/pbrgpu.h"
#include "bvh.h"
#include "sampler.cuh"
#include "vecmath.cuh"

namespace pbr {

// ---- shading queues: 5 material types + "ray escaped"; VolPath adds "surface without a material" (medium
// boundary) and "scattered inside a medium" ----
constexpr int kQueueMiss = PBRGPU_MAT_COUNT;        // 5
constexpr int kQueueNull = PBRGPU_MAT_COUNT + 1;    // 6: GeometricPrimitive with a NULL material (also the triangle's type code)
constexpr int kQueueMedium = PBRGPU_MAT_COUNT + 2;  // 7: MediumInteraction sampled on this segment
constexpr int kNumQueues = PBRGPU_MAT_COUNT + 3;
constexpr uint32_t kQueueDead = 15u;                // path ended before shading (beta went black in Medium::Sample)
// per-path word `meta`: sampler dimension (bits 0-15) | specularBounce (bit 16) | VolPath: bounces (bits 17-23), in-medium (24)
constexpr int kMetaSpecBit = 16;
constexpr int kMetaBounceShift = 17;
constexpr uint32_t kMetaBounceMask = 0x7fu;
constexpr int kMetaInMediumBit = 24;  // VolPath: Ray::medium != nullptr (lets the medium kernels skip a path after one 4-byte read)

// ---- triangle record (leaf order): three float4 = three 16-byte loads from ONE 48-byte record (tv[3 * i + 0..2]: two
// sectors, usually one cache line; three separate arrays cost three sectors in three lines per triangle) ----
//   word 0 = p0.xyz | caller's triangle index     word 1 = p1.xyz | material id (bits 0-21), tangents (22), sphere (23), material type (24-27), flags (28-31)
//   word 2 = p2.xyz | area-light id or -1
constexpr uint32_t kTriMatMask = 0x003fffffu;
constexpr uint32_t kTriTangentBit = 0x00400000u;  // the triangle's mesh has per-vertex tangents (TriangleMesh::s)
constexpr uint32_t kPrimSphereBit = 0x00800000u;  // the record is an analytic sphere (sphere.cuh): word 0 = centre | id, word 1 .x = radius
constexpr int kTriTypeShift = 24;
constexpr int kTriFlagShift = 28;
constexpr uint32_t kTriFlagFlipNormal = 1u;  // reverseOrientation ^ transformSwapsHandedness
constexpr uint32_t kTriFlagReverse = 2u;     // reverseOrientation alone (Triangle.cpp:248: `if (reverseOrientation) ts = -ts`)
constexpr uint32_t kTriFlagNormals = 4u;     // the triangle's mesh has per-vertex shading normals (TriangleMesh::n)
constexpr uint32_t kTriFlagUVs = 8u;         // the triangle's mesh has per-vertex uv (TriangleMesh::uv)

struct CameraDev {
  int type, width, height;
  float lensRadius, focalDistance;
  float r2c[16];
  float c2w[16];
  int medium;  // Camera::medium (index into SceneDev::media) or -1
};

struct EnvDev {
  const float* texels;    // RGB, INFINITE: premultiplied by L ; SKYBOX: raw
  int width, height;
  const float* condFunc;  // [h][w]
  const float* condCdf;   // [h][w+1]
  const float* condInt;   // [h] funcInt per row
  const float* margFunc;  // [h]
  const float* margCdf;   // [h+1]
  float margInt;
};

// ImageTexture on the device: MIP pyramid (level 0 first) as float4 RGB- texels + the UVMapping2D (texture.cuh)
struct TexDev {
  const float4* texels;
  int width, height, nLevels;
  int doTrilinear, wrap;
  float maxAnisotropy;
  float su, sv, du, dv;
  uint32_t levelOffset[16];
};

struct SceneDev {
  const float4* tv;    // [3 * nTris] triangle records (above)
  // per-triangle shading frame, host-precomputed with the reference's own operation order (bit-identical to
  // what Triangle::Intersect / the BSDF ctor would compute per hit): removes two fp64 cross products, two
  // normalisations and ~30 conversions per shaded hit (the shading kernels were XU-pipe bound on them)
  const float4* tf;    // [3 * nTris] frame records: word 0 = geometric normal of Triangle::Intersect (Triangle.cpp:182-184),
                       // orientation flip applied; word 1 = ss = Normalize(dpdu) (Reflection.h:105-110); word 2 = ts = Cross(ns, ss)
  // optional per-vertex shading data (NULL when no mesh has it): smooth normals are interpolated per hit
  const float4* tn0;   // normal of vertex 0 | -          (triangles of meshes with TriangleMesh::n)
  const float4* tn1;
  const float4* tn2;
  const float4* ts0;   // tangent of vertex 0 | -         (triangles of meshes with TriangleMesh::s; NULL when no mesh has them)
  const float4* ts1;
  const float4* ts2;
  const float4* tuv0;  // uv0.xy, uv1.xy
  const float4* tuv1;  // uv2.xy, -, -
  const float4* nodes8;         // Bvh8Node[] viewed as 5 x float4 per node
  const pbrgpu_material* mats;  // roughness fields already remapped to alpha on the host (glibc logf)
  const pbrgpu_light* lights;   // AREA_TRI: .tri = LEAF-ORDER triangle index
  const TexDev* textures;       // ImageTexture table (NULL when no material uses one)
  const float* ewaLut;          // MIPMap::weightLut[128]
  const pbrgpu_medium* media;   // HomogeneousMedium table (VolPath)
  const uint32_t* triMedium;    // per leaf-order triangle: (inside + 1) | (outside + 1) << 16, 0 = none; NULL = no interfaces
  int nMedia;
  int nTris, nLights, nInfinite;  // nTris: primitives in the leaf arrays (triangles + analytic spheres)
  int nSpheres;                   // analytic spheres among them (0: the sphere code is never instantiated)
  unsigned f32one;              // 0x3F800000 (opaque to ptxas: keeps the PRMT selector an immediate, trace.cuh q2f)
  int infinite[4];              // indices of infinite lights (Scene::infiniteLights)
  const float* lightFunc;       // Distribution1D of the active light-sampling strategy
  const float* lightCdf;
  float lightFuncInt;
  float worldRadius;            // Bounds3::BoundingSphere of the scene bound (InfiniteAreaLight::Preprocess)
  V3 worldCenter;
  EnvDev env;
  CameraDev cam;
  HaltonParams hp;
};

struct Counters {
  uint32_t nCur, nNext, nShadow, nMis;
  uint32_t nMat[8];
  uint32_t cursor[4];  // persistent-kernel fetch cursors (TRACE_PATH, TRACE_MIS, TRACE_SHADOW)
  unsigned long long raysClosest, raysShadow, nodesVisited, trisTested, raysPath;
  uint32_t maxRayWork;              // COUNT: most node steps + triangle tests spent on one ray (segment)
  uint32_t volMaxHops, volCapHits;  // VolPath shadow rays: most medium boundaries crossed by one ray / rays stopped at the cap
  // COUNT build of k_trace, lane-activity profile (PBRGPU_LANE_LOG): warp loop iterations, and summed over them the lanes that
  // were busy / did a node step / did a triangle test / had a triangle pass; fetch events and lanes fetched
  unsigned long long laneProf[8];
};

// One live path of the current bounce: a 64-byte record = two full 32-byte sectors.  Records live in QUEUE ORDER: position i
// of `cur` is the i-th live path of this bounce, the shading kernels write the survivors compacted into `nxt` (ping-pong), so
//   * k_trace reads its rays with coalesced 32-byte loads (record offset 0..31) and writes hit[i] / qType[i] at the same index,
//   * the shading kernels gather whole records through the material queues (positions in a material queue are ascending up
//     to the tile-local order of k_classify) - every fetched sector is used completely.
// (Round 1 kept five separate 16-byte arrays indexed by a fixed per-path slot: after a few bounces the live slots were
// scattered, every 16-byte access pulled a half-used sector and k_shade moved ~430 bytes of DRAM per path.)
struct alignas(16) PathRec {
  float4 o;      // ray origin, tMax (VolPath: k_trace stores the hit distance here for Medium::Sample)
  float4 d;      // ray direction, time (VolPath: as_float(Ray::medium id or -1))
  float4 beta;   // beta.rgb, etaScale
  uint32_t idxLo, idxHi;  // Halton / Sobol index of the camera sample
  uint32_t meta;          // dim (bits 0-15) | specularBounce (bit 16) | VolPath: bounces (bits 17-23), in-medium (24)
  uint32_t slot;          // j * nPix + pixel (j = sample within the wave): where this path's radiance accumulates (L[slot])
};
static_assert(sizeof(PathRec) == 64, "PathRec is two 32-byte sectors");

// Wavefront state.  `capacity` paths per wave.
struct WaveDev {
  PathRec* cur;     // live paths of this bounce, queue order
  PathRec* nxt;     // survivors, written compacted by the shading kernels (the host swaps cur / nxt after every bounce)
  float4* hit;      // [position in cur] as_float(leaf tri index or -1), b0, b1, b2
  float4* L;        // [slot] L.rgb (per-path radiance, summed in launch order => deterministic)
  // shadow-ray queue (EstimateDirect light-sampling strategy)
  float4* shO;      // o.xyz, tMax
  float4* shD;      // d.xyz, as_float(slot)
  float4* shC;      // contribution rgb added to L[slot] when unoccluded
  // VolPath only: far endpoint of the shadow segment (VisibilityTester::p1) so that the ray can be re-spawned at
  // every medium boundary it crosses (VisibilityTester::Tr, Light.cpp:31-47), and the medium of the current segment
  float4* shP1;     // p1.xyz, as_float(medium id of the segment or -1)
  float4* shE1;     // p1.pError
  float4* shN1;     // p1.n
  // MIS-ray queue (EstimateDirect BSDF-sampling strategy; closest hit)
  float4* miO;      // o.xyz, tMax -> after k_trace<MIS>: as_float(hit triangle or -1)
  float4* miD;      // d.xyz, as_float(slot)
  float4* miC;      // throughput rgb, as_float(light id)
  uint32_t* qMat;   // [kNumQueues][capacity] positions in `cur`, one queue per shading kernel
  uint32_t* qType;  // [position in cur] shading-queue id (written by k_trace, read by k_classify)
  Counters* ctr;
  uint32_t capacity;
};

struct RenderParams {
  SceneDev sc;
  WaveDev ws;
  int integrator, maxDepth, lightStrategyUniformPick;
  float rrThreshold;
  int x0, y0, w, h;       // pixel region
  long long sBegin;       // first sample index of this wave
  int nsWave;             // samples per pixel in this wave
  int arrayEndDim;        // GlobalSampler::arrayEndDim (5 unless UniformSampleAll)
  int countStats;
  int fetchMinIdle;       // idle lanes of a warp that trigger a ray fetch
  float diffScale;        // 1 / sqrt(samplesPerPixel): RayDifferential::ScaleDifferentials of the render loop (image textures)
  int volShadowAnyHit;    // VolPath, scene without material-less surfaces: a shadow ray ends at ANY surface it finds
  int storeHitT;          // parity hooks: k_trace also stores the hit distance of path rays (rayO.w), like the VOL build
  int atomicL;            // several shadow / MIS rays per path and bounce (Whitted / DirectLighting): atomics on L
  int deferShadow;        // VolPath tail (pbrgpu.cu): 1 = the shadow queue is left alone and keeps growing (k_trace skips it, k_classify
                          // does not reset it); 2 = the flush launch: k_trace serves ONLY the shadow queue and leaves each ray's
                          // contribution x transmittance in shC (0 when blocked) for k_shadow_add
  int volInLane;          // VolPath: k_trace applies Medium::Sample's weight and continues paths across material-less medium
                          // boundaries in the lane (k_trace.cuh vol_path_next_segment); 0 = the separate kernels do it
  // Halton offset of the pixel a global index belongs to: index = offset(pixel) + sample * stride, offset < stride
  PBR_D int64_t pixelOffsetOf(int64_t index) const { return index % sc.hp.sampleStride; }
  float* accum;           // [h*w] RGBA sums (A = sample count)
};

}  // namespace pbr
