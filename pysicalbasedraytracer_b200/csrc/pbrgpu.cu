// pbrgpu.cu - implementation of the C-ABI in include/pbrgpu.h: host-side flattening of the scene
// description into HBM-resident SoA buffers and the wavefront launch schedule.
//
// Launch schedule for one wave of paths (all at the same bounce; PathIntegrator.cpp:42-108):
//   k_raygen -> for bounce b in 0..maxDepth: k_trace (path rays of b + shadow/MIS rays of b-1) -> k_resolve_mis
//               -> k_classify -> k_shade<MAT> x5 + k_shade_miss -> k_next_bounce ;  k_film_accumulate
// Queue sizes live in device counters; the host never reads them inside a render (no sync), grids
// are sized for the wave capacity and the kernels loop over the live count.
// There is NO CPU fallback: every compute entry point needs a CUDA device.
#include <cuda_runtime.h>

#include <cmath>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <memory>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "host_tables.h"
#include "kmisc.cuh"
#include "launch.h"
#include "nccl_dyn.h"

using namespace pbr;

namespace {

thread_local std::string g_err;
int fail(const std::string& m) { g_err = m; return 1; }

#define CK(call)                                                                                  \
  do {                                                                                            \
    cudaError_t e_ = (call);                                                                      \
    if (e_ != cudaSuccess)                                                                        \
      return fail(std::string(#call) + ": " + cudaGetErrorString(e_) + " (" __FILE__ ":" + std::to_string(__LINE__) + ")"); \
  } while (0)

template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  cudaError_t alloc(size_t count) {
    release();
    n = count;
    if (count == 0) return cudaSuccess;
    return cudaMalloc((void**)&p, count * sizeof(T));
  }
  cudaError_t upload(const std::vector<T>& v, cudaStream_t s) {
    cudaError_t e = alloc(v.size());
    if (e != cudaSuccess || v.empty()) return e;
    return cudaMemcpyAsync(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, s);
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
  }
  ~DevBuf() { release(); }
};

}  // namespace

struct pbrgpu_ctx {
  int device = 0;
  int smCount = 148;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[8] = {};
  std::vector<cudaEvent_t> pool;  // PBRGPU_RENDER_TIMERS
  // two waves in flight (render_device_single): the second wave's stream, fork / join events, one "film updated" event per wave set
  cudaStream_t stream2 = nullptr;
  cudaEvent_t evFork = nullptr, evJoin = nullptr, evFilm[2] = {};
  size_t l2PersistBytes = 0;      // persisting-L2 carve-out granted by the device (0 = none)
  size_t l2WindowMax = 0;         // cudaDevAttrMaxAccessPolicyWindowSize
  // ---- multi-GPU (scene replicated, sample range sharded, ONE ncclReduce(sum) of the RGBA film to rank 0) ----
  // (a) pbrgpu_init_multi: this context is a GROUP - it owns one ordinary context per device and one communicator each
  std::vector<pbrgpu_ctx*> subs;
  std::vector<NcclApi::Comm> comms;
  // (b) pbrgpu_comm_attach: one process per GPU (torchrun / MPI); this ordinary context is rank `rank` of `nranks`
  NcclApi::Comm comm = nullptr;
  int nranks = 1, rank = 0;
};

// CUDA-event timeline: mark(cat) closes the previous interval and opens one of category `cat`.
enum { T_RAYGEN, T_CLOSEST, T_SHADE, T_SHADOW, T_MIS, T_FILM, T_END, T_COUNT };
struct Timeline {
  pbrgpu_ctx* c;
  cudaStream_t st;
  bool on;
  size_t used = 0;
  std::vector<int> cats;
  cudaError_t mark(int cat) {
    if (!on) return cudaSuccess;
    if (used == c->pool.size()) {
      cudaEvent_t e;
      cudaError_t r = cudaEventCreate(&e);
      if (r != cudaSuccess) return r;
      c->pool.push_back(e);
    }
    cats.push_back(cat);
    return cudaEventRecord(c->pool[used++], st);
  }
  cudaError_t collect(double out[T_COUNT], uint64_t counts[T_COUNT]) {
    for (int k = 0; k < T_COUNT; ++k) { out[k] = 0; counts[k] = 0; }
    for (size_t i = 0; i + 1 < used; ++i) {
      float ms = 0;
      cudaError_t r = cudaEventElapsedTime(&ms, c->pool[i], c->pool[i + 1]);
      if (r != cudaSuccess) return r;
      out[cats[i]] += ms;
      counts[cats[i]]++;
    }
    return cudaSuccess;
  }
};

struct LastRender {
  uint64_t samples = 0, launches = 0;
  Timeline tl{nullptr, nullptr, false};
};

struct pbrgpu_scene {
  LastRender last;
  pbrgpu_ctx* ctx = nullptr;
  SceneDev sc{};
  BuiltBvh8 bvh;
  std::vector<int32_t> leafOf;  // caller's triangle index -> leaf-order position
  std::vector<pbrgpu_light> lights;  // device copy (tri remapped to leaf order)
  std::vector<float> lightPower;     // Light::Power().y() per light
  bool haveCamera = false, haveSampler = false;
  int64_t spp = 0;
  // device buffers
  DevBuf<float4> tfr, tn0, tn1, tn2, ts0, ts1, ts2, tuv0, tuv1;  // tfr: per-triangle frame records (SceneDev::tf)
  DevBuf<float4> geom;  // ONE slab: Bvh8Node[] | triangle records - everything k_trace reads, covered by one L2 access-policy window
  DevBuf<pbrgpu_material> mats;
  DevBuf<pbrgpu_light> dlights;
  DevBuf<float> lightFunc, lightCdf;
  DevBuf<float> envTexels, condFunc, condCdf, condInt, margFunc, margCdf;
  DevBuf<uint16_t> perms;
  DevBuf<int32_t> primes, primeSums;
  DevBuf<uint32_t> primeMagic, sobolMat;
  DevBuf<float> dimConst;
  DevBuf<uint64_t> sobolVdc, sobolVdcInv;
  int lightStrategyBuilt = -1;
  int traceBlocksPerSm = 0;
  int fetchMinIdle = 8;
  int waveLog2 = 25;
  bool hasMat[PBRGPU_MAT_COUNT] = {};  // material types present in the scene (absent ones: no shade launch)
  bool hasNullMat = false;             // some triangle has no material (medium boundary): VolPath only
  bool hasTextures = false;            // some material takes Kd from an ImageTexture: TEX shading kernels, Path / VolPath only
  std::vector<DevBuf<float4>> texPyramids;
  DevBuf<TexDev> texTable;
  DevBuf<float> ewaLut;
  DevBuf<pbrgpu_medium> media;
  DevBuf<uint32_t> triMedium;
  DevBuf<float4> shP1, shE1, shN1;     // VolPath shadow-ray endpoints (allocated on first VolPath render)
  size_t volRayCapacity = 0;
  uint32_t* hNCur = nullptr;           // pinned: live-path count read back by the VolPath loop
  // wavefront state
  uint32_t capacity = 0;
  size_t rayCapacity = 0;
  DevBuf<PathRec> recA, recB;  // live paths in queue order, ping-pong (WaveDev::cur / nxt)
  DevBuf<float4> hit, L, shO, shD, shC, miO, miD, miC;
  DevBuf<uint32_t> qMat, qType;
  DevBuf<Counters> ctr;
  // second wave set (two waves in flight): same buffers once more, allocated by the first render that runs two waves
  struct WaveSet {
    DevBuf<PathRec> recA, recB;
    DevBuf<float4> hit, L, shO, shD, shC, miO, miD, miC;
    DevBuf<uint32_t> qMat, qType;
    DevBuf<Counters> ctr;
    uint32_t capacity = 0;
    size_t rayCapacity = 0;
  };
  std::unique_ptr<WaveSet> alt;
  bool altUsed = false;  // the last render ran waves on the second set: its ray counters belong to the totals
  int wavesInFlight = 2;
  DevBuf<float> accum;
  size_t accumPix = 0;
  // host-buffer path (pbrgpu_render): persistent device RGB buffer + pinned staging so that the image leaves the
  // GPU at PCIe speed whatever memory the caller's buffer lives in
  DevBuf<float> drgb;
  DevBuf<unsigned char> d8;
  void* hstage = nullptr;
  size_t hstageBytes = 0;
  std::vector<pbrgpu_scene*> replicas;  // scene of a GROUP context: one ordinary scene per device (this object holds nothing else)
  ~pbrgpu_scene() { if (hstage) cudaFreeHost(hstage); if (hNCur) cudaFreeHost(hNCur); }
};

namespace {

int ensure_wave(pbrgpu_scene* s, uint32_t cap, uint32_t raysPerHit, bool vol) {
  // shadow / MIS queues hold up to raysPerHit rays per path and bounce (1 for the Path integrator)
  size_t rcap = (size_t)cap * raysPerHit;
  if (vol && s->volRayCapacity < rcap) {
    CK(s->shP1.alloc(rcap)); CK(s->shE1.alloc(rcap)); CK(s->shN1.alloc(rcap));
    s->volRayCapacity = rcap;
  }
  if (vol && !s->hNCur) CK(cudaHostAlloc((void**)&s->hNCur, 4 * sizeof(uint32_t), cudaHostAllocDefault));
  if (s->rayCapacity < rcap) {
    CK(s->shO.alloc(rcap)); CK(s->shD.alloc(rcap)); CK(s->shC.alloc(rcap));
    CK(s->miO.alloc(rcap)); CK(s->miD.alloc(rcap)); CK(s->miC.alloc(rcap));
    s->rayCapacity = rcap;
  }
  if (s->capacity >= cap) return 0;
  CK(s->recA.alloc(cap)); CK(s->recB.alloc(cap)); CK(s->hit.alloc(cap)); CK(s->L.alloc(cap));
  CK(s->qMat.alloc((size_t)cap * kNumQueues));
  CK(s->qType.alloc(cap));
  if (!s->ctr.p) CK(s->ctr.alloc(1));
  s->capacity = cap;
  return 0;
}

void fill_wave(pbrgpu_scene* s, WaveDev* w) {
  w->cur = s->recA.p; w->nxt = s->recB.p; w->hit = s->hit.p; w->L = s->L.p;
  w->shO = s->shO.p; w->shD = s->shD.p; w->shC = s->shC.p;
  w->shP1 = s->shP1.p; w->shE1 = s->shE1.p; w->shN1 = s->shN1.p;
  w->miO = s->miO.p; w->miD = s->miD.p; w->miC = s->miC.p;
  w->qMat = s->qMat.p; w->qType = s->qType.p;
  w->ctr = s->ctr.p;
  w->capacity = s->capacity;
}

// Second wave set for a render that keeps two waves in flight.  Returns 1 on a CUDA error, sets *ok = false (and frees
// nothing) when the device does not have the memory to spare: the render then runs its waves one after the other.
int ensure_alt_wave(pbrgpu_scene* s, uint32_t cap, uint32_t raysPerHit, bool* ok) {
  *ok = false;
  if (!s->alt) s->alt.reset(new pbrgpu_scene::WaveSet());
  pbrgpu_scene::WaveSet& a = *s->alt;
  const size_t rcap = (size_t)cap * raysPerHit;
  size_t need = 0;
  if (a.rayCapacity < rcap) need += rcap * 6 * sizeof(float4);
  if (a.capacity < cap) need += (size_t)cap * (2 * sizeof(PathRec) + 2 * sizeof(float4) + (kNumQueues + 1) * sizeof(uint32_t));
  if (need) {
    size_t freeB = 0, totalB = 0;
    CK(cudaMemGetInfo(&freeB, &totalB));
    if (freeB < need + ((size_t)2 << 30)) return 0;
  }
  if (a.rayCapacity < rcap) {
    CK(a.shO.alloc(rcap)); CK(a.shD.alloc(rcap)); CK(a.shC.alloc(rcap));
    CK(a.miO.alloc(rcap)); CK(a.miD.alloc(rcap)); CK(a.miC.alloc(rcap));
    a.rayCapacity = rcap;
  }
  if (a.capacity < cap) {
    CK(a.recA.alloc(cap)); CK(a.recB.alloc(cap)); CK(a.hit.alloc(cap)); CK(a.L.alloc(cap));
    CK(a.qMat.alloc((size_t)cap * kNumQueues));
    CK(a.qType.alloc(cap));
    a.capacity = cap;
  }
  if (!a.ctr.p) CK(a.ctr.alloc(1));
  *ok = true;
  return 0;
}

void fill_alt_wave(pbrgpu_scene* s, WaveDev* w) {
  pbrgpu_scene::WaveSet& a = *s->alt;
  w->cur = a.recA.p; w->nxt = a.recB.p; w->hit = a.hit.p; w->L = a.L.p;
  w->shO = a.shO.p; w->shD = a.shD.p; w->shC = a.shC.p;
  w->shP1 = nullptr; w->shE1 = nullptr; w->shN1 = nullptr;  // VolPath never runs on the second set
  w->miO = a.miO.p; w->miD = a.miD.p; w->miC = a.miC.p;
  w->qMat = a.qMat.p; w->qType = a.qType.p;
  w->ctr = a.ctr.p;
  w->capacity = a.capacity;
}

// Light sampling distribution (LightDistrib.cpp:8-45): "uniform" (also what "spatial" falls back to, and
// what a single-light scene always gets) or "power".
int build_light_distribution(pbrgpu_scene* s, int strategy) {
  int n = (int)s->lights.size();
  if (n == 0) { s->sc.lightFunc = s->sc.lightCdf = nullptr; s->sc.lightFuncInt = 0; return 0; }
  if (n == 1) strategy = PBRGPU_LIGHTSTRATEGY_UNIFORM;
  if (s->lightStrategyBuilt == strategy) return 0;
  std::vector<float> f((size_t)n, 1.f);
  if (strategy == PBRGPU_LIGHTSTRATEGY_POWER) f = s->lightPower;
  host::Dist1D d;
  d.build(f.data(), n);
  CK(s->lightFunc.upload(d.func, s->ctx->stream));
  CK(s->lightCdf.upload(d.cdf, s->ctx->stream));
  CK(cudaStreamSynchronize(s->ctx->stream));
  s->sc.lightFunc = s->lightFunc.p;
  s->sc.lightCdf = s->lightCdf.p;
  s->sc.lightFuncInt = d.funcInt;
  s->lightStrategyBuilt = strategy;
  return 0;
}

inline int mod_i(int a, int b) { int r = a - (a / b) * b; return r < 0 ? r + b : r; }

// host twin of MIPMap::triangle(0, st) (Texture/MIPMap.h:256-266) used to build the sampling image
void env_bilinear_host(const std::vector<float>& tex, int W, int H, float s_, float t_, float out[3]) {
  float s = s_ * W - 0.5f, t = t_ * H - 0.5f;
  int s0 = (int)std::floor(s), t0 = (int)std::floor(t);
  float ds = s - s0, dt = t - t0;
  auto texel = [&](int a, int b) { return &tex[3 * ((size_t)mod_i(b, H) * W + mod_i(a, W))]; };
  const float *a = texel(s0, t0), *b = texel(s0, t0 + 1), *c = texel(s0 + 1, t0), *d = texel(s0 + 1, t0 + 1);
  float w0 = (1 - ds) * (1 - dt), w1 = (1 - ds) * dt, w2 = ds * (1 - dt), w3 = ds * dt;
  for (int k = 0; k < 3; ++k) out[k] = ((w0 * a[k] + w1 * b[k]) + w2 * c[k]) + w3 * d[k];
}

uint32_t grid_for(const pbrgpu_ctx* c, uint64_t n, int perSm) {
  uint64_t blocks = (n + kBlock - 1) / kBlock;
  uint64_t cap = (uint64_t)c->smCount * perSm;
  if (blocks < 1) blocks = 1;
  return (uint32_t)(blocks < cap ? blocks : cap);
}

// Grid of a grid-stride kernel = exactly one resident wave (SMs x the kernel's occupancy), never more blocks than work:
// a fixed "8 blocks per SM" made 2.67 waves of the 80-register shading kernels, whose partial last wave left two thirds
// of the SMs idle.
template <typename K>
uint32_t grid_occ(const pbrgpu_ctx* c, K kernel, int block, uint64_t n) {
  static std::unordered_map<const void*, int> cache;
  const void* key = reinterpret_cast<const void*>(kernel);
  auto it = cache.find(key);
  if (it == cache.end()) {
    int nb = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kernel, block, 0) != cudaSuccess || nb < 1) { cudaGetLastError(); nb = 4; }
    if (const char* e = getenv("PBRGPU_SHADE_WAVES")) nb *= std::max(1, atoi(e));
    it = cache.emplace(key, nb).first;
  }
  const uint64_t blocks = std::max<uint64_t>(1, (n + block - 1) / block);
  return (uint32_t)std::min<uint64_t>(blocks, (uint64_t)c->smCount * it->second);
}

// one resident wave of k_shade<mat, vol, tex>
uint32_t shade_grid(const pbrgpu_ctx* c, int mat, bool vol, bool tex, uint64_t n) {
  static int cache[PBRGPU_MAT_COUNT][2][2] = {};
  int& nb = cache[mat][vol ? 1 : 0][tex ? 1 : 0];
  if (nb == 0) {
    nb = std::max(1, shade_blocks_per_sm(mat, vol, tex));
    if (const char* e = getenv("PBRGPU_SHADE_WAVES")) nb *= std::max(1, atoi(e));
  }
  const uint64_t blocks = std::max<uint64_t>(1, (n + kShadeBlock - 1) / kShadeBlock);
  return (uint32_t)std::min<uint64_t>(blocks, (uint64_t)c->smCount * nb);
}

}  // namespace

extern "C" {

const char* pbrgpu_last_error(void) { return g_err.c_str(); }

int pbrgpu_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

int pbrgpu_init(int device_id, pbrgpu_ctx** out) {
  if (!out) return fail("pbrgpu_init: out == NULL");
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    cudaGetLastError();
    return fail("pbrgpu_init: no CUDA device available (this library has no CPU fallback)");
  }
  if (device_id < 0 || device_id >= n) return fail("pbrgpu_init: bad device id");
  CK(cudaSetDevice(device_id));
  auto* c = new pbrgpu_ctx;
  c->device = device_id;
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, device_id));
  c->smCount = prop.multiProcessorCount;
  {
    // Optional L2 set-aside for the traversal data (BVH nodes + triangle vertices, ~50 MB on the dragon-class scene).
    // OFF by default - measured on B200: k_trace does not get faster (25.4 ms either way: it is issue-bound, its L2 misses
    // are not the limiter) while the shading kernels, which stream the wavefront state through what is left of the L2,
    // slow down (12.7 -> 13.2 ms with 32 MB set aside, 18.7 ms with 64 MB, 22.5 ms with 96 MB).
    size_t want = 0;
    if (const char* e = getenv("PBRGPU_L2_PERSIST_MB")) want = (size_t)std::max(0, atoi(e)) << 20;
    want = std::min<size_t>(want, (size_t)prop.persistingL2CacheMaxSize);
    if (want > 0 && cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want) == cudaSuccess) {
      c->l2PersistBytes = want;
      c->l2WindowMax = (size_t)prop.accessPolicyMaxWindowSize;
    } else {
      cudaGetLastError();
    }
    if (getenv("PBRGPU_VERBOSE"))
      fprintf(stderr, "pbrgpu: L2 %d MB, persisting max %d MB, granted %zu MB, window max %d MB\n", prop.l2CacheSize >> 20,
              prop.persistingL2CacheMaxSize >> 20, c->l2PersistBytes >> 20, prop.accessPolicyMaxWindowSize >> 20);
  }
  CK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  for (auto& ev : c->ev) CK(cudaEventCreate(&ev));
  *out = c;
  return 0;
}

void pbrgpu_shutdown(pbrgpu_ctx* c) {
  if (!c) return;
  if (const NcclApi* api = (c->comm || !c->comms.empty()) ? nccl_api(nullptr) : nullptr) {
    if (c->comm) api->CommDestroy(c->comm);
    for (NcclApi::Comm m : c->comms) if (m) api->CommDestroy(m);
  }
  for (pbrgpu_ctx* sub : c->subs) pbrgpu_shutdown(sub);
  cudaSetDevice(c->device);
  for (auto& ev : c->ev) if (ev) cudaEventDestroy(ev);
  for (auto& ev : c->pool) cudaEventDestroy(ev);
  for (cudaEvent_t ev : {c->evFork, c->evJoin, c->evFilm[0], c->evFilm[1]}) if (ev) cudaEventDestroy(ev);
  if (c->stream2) cudaStreamDestroy(c->stream2);
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
}

int pbrgpu_halton_params(const int32_t sample_bounds[4], int32_t out7[7]) {
  host::halton_params(sample_bounds, out7);
  return 0;
}

int64_t pbrgpu_compute_perms(uint16_t* out, int64_t capacity) {
  std::vector<int32_t> primes, sums;
  host::prime_tables(&primes, &sums);
  std::vector<uint16_t> p = host::radical_inverse_permutations(primes);
  if (out) memcpy(out, p.data(), sizeof(uint16_t) * (size_t)std::min<int64_t>(capacity, (int64_t)p.size()));
  return (int64_t)p.size();
}

static int scene_create_single(pbrgpu_ctx* ctx, const pbrgpu_scene_desc* d, const BuiltBvh8* prebuilt, pbrgpu_scene** out) {
  if (!ctx || !d || !out) return fail("pbrgpu_scene_create: NULL argument");
  *out = nullptr;
  if (d->n_spheres < 0 || (d->n_spheres > 0 && !d->spheres)) return fail("pbrgpu_scene_create: bad sphere table");
  if (d->n_tris < 0 || (d->n_tris == 0 && d->n_spheres == 0) || (d->n_tris > 0 && (d->n_verts <= 0 || !d->verts || !d->indices || !d->tri_material)))
    return fail("pbrgpu_scene_create: empty scene");
  for (int i = 0; i < d->n_spheres; ++i) {
    const pbrgpu_sphere& sp = d->spheres[i];
    if (!(sp.radius > 0)) return fail("pbrgpu_scene_create: sphere radius must be positive");
    if (sp.material < 0 || sp.material >= d->n_materials) return fail("pbrgpu_scene_create: sphere material out of range");
    if (sp.light >= d->n_lights) return fail("pbrgpu_scene_create: sphere light out of range");
    if (d->materials[sp.material].kd_texture_plus1 > 0 || d->materials[sp.material].ks_texture_plus1 > 0 ||
        d->materials[sp.material].rough_texture_plus1 > 0) return fail("pbrgpu_scene_create: image textures on analytic spheres are not supported");
  }
  if (d->n_tris >= (1 << 29)) return fail("pbrgpu_scene_create: too many triangles");
  if (d->n_materials < 0 || d->n_materials > (int)kTriMatMask || (d->n_materials > 0 && !d->materials))
    return fail("pbrgpu_scene_create: bad material count");
  if (d->n_media < 0 || d->n_media > 0xfffe || (d->n_media > 0 && !d->media)) return fail("pbrgpu_scene_create: bad media table");
  for (int i = 0; i < d->n_media; ++i)
    if (d->media[i].type != PBRGPU_MEDIUM_HOMOGENEOUS) return fail("pbrgpu_scene_create: unknown medium type (HomogeneousMedium only)");
  for (int64_t i = 0; i < (int64_t)d->n_tris * 3; ++i)
    if (d->indices[i] < 0 || d->indices[i] >= d->n_verts) return fail("pbrgpu_scene_create: vertex index out of range");
  for (int i = 0; i < d->n_tris; ++i) {
    if (d->tri_light && d->tri_light[i] >= 0 && d->lights[d->tri_light[i]].type != PBRGPU_LIGHT_AREA_TRI)
      return fail("pbrgpu_scene_create: tri_light must name an AREA_TRI light");
    if (d->tri_material[i] < -1 || d->tri_material[i] >= d->n_materials)
      return fail("pbrgpu_scene_create: tri_material out of range");
    for (const int32_t* tm : {d->tri_medium_inside, d->tri_medium_outside})
      if (tm && (tm[i] < -1 || tm[i] >= d->n_media)) return fail("pbrgpu_scene_create: triangle medium index out of range");
    if (d->tri_light && d->tri_light[i] >= d->n_lights) return fail("pbrgpu_scene_create: tri_light out of range");
  }
  CK(cudaSetDevice(ctx->device));
  std::unique_ptr<pbrgpu_scene> s(new pbrgpu_scene);
  s->ctx = ctx;
  cudaStream_t st = ctx->stream;

  int maxLeaf = 3;
  if (const char* e = getenv("PBRGPU_MAX_LEAF")) maxLeaf = std::max(1, std::min(3, atoi(e)));
  const int nTriangles = d->n_tris;
  const int nT = d->n_tris + d->n_spheres;  // primitives in the leaf arrays: triangles, then analytic spheres
  if (prebuilt) s->bvh = *prebuilt;  // group context: built once, uploaded to every device
  else {
    std::vector<Aabb> sb((size_t)d->n_spheres);  // WorldBound() = ObjectToWorld(ObjectBound()): centre -/+ radius per axis
    for (int i = 0; i < d->n_spheres; ++i)
      for (int a = 0; a < 3; ++a) {
        sb[i].lo[a] = -d->spheres[i].radius + d->spheres[i].center[a];
        sb[i].hi[a] = d->spheres[i].radius + d->spheres[i].center[a];
      }
    build_bvh8(d->verts, d->indices, nTriangles, maxLeaf, &s->bvh, sb.data(), d->n_spheres);
  }
  s->leafOf.assign((size_t)nT, -1);
  std::vector<float4> vrec((size_t)nT * 3), vfr((size_t)nT * 3);  // 48-byte triangle / frame records, leaf order
  std::vector<float4> vn0, vn1, vn2, vuv0, vuv1, vs0, vs1, vs2;  // only when some mesh has shading normals / uvs / tangents
  bool anyNormals = false;
  for (int k = 0; k < nT; ++k) {
    int ti = s->bvh.triOrder[k];
    s->leafOf[ti] = k;
    if (ti >= nTriangles) {  // analytic sphere (sphere.cuh): word 0 = centre | id, word 1 = radius | material word, word 2 = - | light
      const pbrgpu_sphere& sp = d->spheres[ti - nTriangles];
      const uint32_t flags = sp.reverse_orientation ? (kTriFlagFlipNormal | kTriFlagReverse) : 0u;
      const uint32_t w1 = (uint32_t)sp.material | kPrimSphereBit | ((uint32_t)d->materials[sp.material].type << kTriTypeShift) | (flags << kTriFlagShift);
      vrec[3 * (size_t)k] = make_float4(sp.center[0], sp.center[1], sp.center[2], u2f((uint32_t)ti));
      vrec[3 * (size_t)k + 1] = make_float4(sp.radius, 0.f, 0.f, u2f(w1));
      vrec[3 * (size_t)k + 2] = make_float4(0.f, 0.f, 0.f, u2f((uint32_t)sp.light));
      vfr[3 * (size_t)k] = vfr[3 * (size_t)k + 1] = vfr[3 * (size_t)k + 2] = make_float4(0.f, 0.f, 0.f, 0.f);
      continue;
    }
    const float* a = d->verts + 3 * (size_t)d->indices[3 * (size_t)ti];
    const float* b = d->verts + 3 * (size_t)d->indices[3 * (size_t)ti + 1];
    const float* c = d->verts + 3 * (size_t)d->indices[3 * (size_t)ti + 2];
    int m = d->tri_material[ti];
    // tri_flags: bit0 reverseOrientation ^ transformSwapsHandedness, bit1 reverseOrientation, bit2 / bit3 "this triangle's
    // mesh has per-vertex normals / uvs" (default when tri_flags is NULL: every triangle uses the arrays that are given)
    const uint8_t tf = d->tri_flags ? d->tri_flags[ti] : (uint8_t)((d->normals ? 4 : 0) | (d->uvs ? 8 : 0) | (d->tangents ? 16 : 0));
    uint32_t flags = (tf & 1 ? kTriFlagFlipNormal : 0u) | (tf & 2 ? kTriFlagReverse : 0u) |
                     ((tf & 4) && d->normals ? kTriFlagNormals : 0u) | ((tf & 8) && d->uvs ? kTriFlagUVs : 0u);
    // a triangle without a material carries the queue id of "medium boundary" as its type code
    if (m < 0) s->hasNullMat = true;
    const bool hasTangents = (tf & 16) && d->tangents;  // TriangleMesh::s
    uint32_t w1 = m < 0 ? ((uint32_t)kQueueNull << kTriTypeShift) | (flags << kTriFlagShift)
                        : (uint32_t)m | ((uint32_t)d->materials[m].type << kTriTypeShift) | (flags << kTriFlagShift);
    if (hasTangents) w1 |= kTriTangentBit;
    int32_t lid = d->tri_light ? d->tri_light[ti] : -1;
    vrec[3 * (size_t)k] = make_float4(a[0], a[1], a[2], u2f((uint32_t)ti));
    vrec[3 * (size_t)k + 1] = make_float4(b[0], b[1], b[2], u2f(w1));
    vrec[3 * (size_t)k + 2] = make_float4(c[0], c[1], c[2], u2f((uint32_t)lid));
    V3 n, ss, ts;
    float uv[6] = {0.f, 0.f, 1.f, 0.f, 1.f, 1.f};  // Triangle::GetUVs default (Triangle.h:50-54)
    if (flags & kTriFlagUVs)
      for (int q = 0; q < 3; ++q) {
        uv[2 * q] = d->uvs[2 * (size_t)d->indices[3 * (size_t)ti + q]];
        uv[2 * q + 1] = d->uvs[2 * (size_t)d->indices[3 * (size_t)ti + q] + 1];
      }
    triangle_frame(V3(a[0], a[1], a[2]), V3(b[0], b[1], b[2]), V3(c[0], c[1], c[2]), uv, (flags & kTriFlagFlipNormal) != 0, &n, &ss, &ts);
    if (flags & kTriFlagNormals) {
      anyNormals = true;
      const float* n0 = d->normals + 3 * (size_t)d->indices[3 * (size_t)ti];
      const float* n1 = d->normals + 3 * (size_t)d->indices[3 * (size_t)ti + 1];
      const float* n2 = d->normals + 3 * (size_t)d->indices[3 * (size_t)ti + 2];
      if (vn0.empty()) { vn0.assign((size_t)nT, make_float4(0, 0, 0, 0)); vn1 = vn0; vn2 = vn0; }
      vn0[k] = make_float4(n0[0], n0[1], n0[2], 0.f);
      vn1[k] = make_float4(n1[0], n1[1], n1[2], 0.f);
      vn2[k] = make_float4(n2[0], n2[1], n2[2], 0.f);
    }
    if (hasTangents) {
      if (vs0.empty()) { vs0.assign((size_t)nT, make_float4(0, 0, 0, 0)); vs1 = vs0; vs2 = vs0; }
      for (int q = 0; q < 3; ++q) {
        const float* sv = d->tangents + 3 * (size_t)d->indices[3 * (size_t)ti + q];
        (q == 0 ? vs0 : (q == 1 ? vs1 : vs2))[k] = make_float4(sv[0], sv[1], sv[2], 0.f);
      }
    }
    if (flags & kTriFlagUVs) {
      if (vuv0.empty()) { vuv0.assign((size_t)nT, make_float4(0, 0, 0, 0)); vuv1 = vuv0; }
      vuv0[k] = make_float4(uv[0], uv[1], uv[2], uv[3]);
      vuv1[k] = make_float4(uv[4], uv[5], 0.f, 0.f);
    }
    vfr[3 * (size_t)k] = make_float4(n.x, n.y, n.z, 0.f);
    vfr[3 * (size_t)k + 1] = make_float4(ss.x, ss.y, ss.z, 0.f);
    vfr[3 * (size_t)k + 2] = make_float4(ts.x, ts.y, ts.z, 0.f);
  }
  CK(s->tfr.upload(vfr, st));
  if (anyNormals) { CK(s->tn0.upload(vn0, st)); CK(s->tn1.upload(vn1, st)); CK(s->tn2.upload(vn2, st)); }
  if (!vuv0.empty()) { CK(s->tuv0.upload(vuv0, st)); CK(s->tuv1.upload(vuv1, st)); }
  if (!vs0.empty()) { CK(s->ts0.upload(vs0, st)); CK(s->ts1.upload(vs1, st)); CK(s->ts2.upload(vs2, st)); }
  {
    static_assert(sizeof(Bvh8Node) == 5 * sizeof(float4), "Bvh8Node is five 16-byte words");
    const size_t nN = s->bvh.nodes.size() * 5, nTv = (size_t)nT;
    CK(s->geom.alloc(nN + 3 * nTv));
    CK(cudaMemcpyAsync(s->geom.p, s->bvh.nodes.data(), nN * sizeof(float4), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(s->geom.p + nN, vrec.data(), 3 * nTv * sizeof(float4), cudaMemcpyHostToDevice, st));
    CK(cudaStreamSynchronize(st));  // vrec is a local
    s->sc.nodes8 = s->geom.p;
    s->sc.tv = s->geom.p + nN;
  }
  // participating media + per-triangle MediumInterface (leaf order)
  if (d->n_media > 0) {
    std::vector<pbrgpu_medium> media(d->media, d->media + d->n_media);
    CK(s->media.upload(media, st));
    s->sc.media = s->media.p;
    s->sc.nMedia = d->n_media;
    if (d->tri_medium_inside || d->tri_medium_outside) {
      std::vector<uint32_t> tm((size_t)nT);
      for (int k = 0; k < nT; ++k) {
        const int ti = s->bvh.triOrder[k];
        const bool isTri = ti < nTriangles;  // analytic spheres carry no MediumInterface
        const int in = isTri && d->tri_medium_inside ? d->tri_medium_inside[ti] : -1;
        const int out = isTri && d->tri_medium_outside ? d->tri_medium_outside[ti] : -1;
        tm[k] = (uint32_t)(in + 1) | ((uint32_t)(out + 1) << 16);
      }
      CK(s->triMedium.upload(tm, st));
      s->sc.triMedium = s->triMedium.p;
    }
  }

  // materials: RoughnessToAlpha applied on the host (glibc logf, as the reference would)
  std::vector<pbrgpu_material> mats;
  if (d->n_materials > 0) mats.assign(d->materials, d->materials + d->n_materials);
  else mats.push_back(pbrgpu_material{});  // a scene of medium boundaries only: keep the table non-empty
  for (auto& m : mats) {
    if (m.type < 0 || m.type >= PBRGPU_MAT_COUNT) return fail("pbrgpu_scene_create: unknown material type");
    s->hasMat[m.type] = true;
    m.pad3 = (m.type == PBRGPU_MAT_GLASS && m.u_rough == 0 && m.v_rough == 0) ? 1.f : 0.f;  // isSpecular (device copy only)
    if (m.remap_roughness && m.type != PBRGPU_MAT_MATTE && m.type != PBRGPU_MAT_MIRROR) {
      m.u_rough = host::roughness_to_alpha(m.u_rough);
      m.v_rough = host::roughness_to_alpha(m.v_rough);
    }
  }
  // image textures: MIP pyramid per texture (MIPMap.h:146-168: each level is the 2x2 box filter of the previous one, texels
  // fetched through the wrap mode), EWA weight table (MIPMap.h:170-176)
  if (d->n_textures < 0 || (d->n_textures > 0 && !d->textures)) return fail("pbrgpu_scene_create: bad texture table");
  for (auto& m : mats) {
    if (m.kd_texture_plus1 < 0 || m.kd_texture_plus1 > d->n_textures) return fail("pbrgpu_scene_create: kd_texture_plus1 out of range");
    if (m.kd_texture_plus1 > 0) {
      if (m.type != PBRGPU_MAT_MATTE && m.type != PBRGPU_MAT_PLASTIC)
        return fail("pbrgpu_scene_create: image textures are supported for the Kd of Matte and Plastic materials");
      s->hasTextures = true;
    }
    if (m.ks_texture_plus1 < 0 || m.ks_texture_plus1 > d->n_textures) return fail("pbrgpu_scene_create: ks_texture_plus1 out of range");
    if (m.ks_texture_plus1 > 0) {
      if (m.type != PBRGPU_MAT_PLASTIC) return fail("pbrgpu_scene_create: a Ks image texture is supported for Plastic materials only");
      s->hasTextures = true;
    }
    if (m.rough_texture_plus1 < 0 || m.rough_texture_plus1 > d->n_textures) return fail("pbrgpu_scene_create: rough_texture_plus1 out of range");
    if (m.rough_texture_plus1 > 0) {
      if (m.type != PBRGPU_MAT_PLASTIC) return fail("pbrgpu_scene_create: a roughness image texture is supported for Plastic materials only");
      // RoughnessToAlpha of a constant is evaluated on the host with glibc's logf (above); per hit it would be the device's logf
      if (m.remap_roughness) return fail("pbrgpu_scene_create: a roughness image texture needs remap_roughness == 0");
      s->hasTextures = true;
    }
  }
  if (d->n_textures > 0) {
    std::vector<TexDev> table((size_t)d->n_textures);
    s->texPyramids.resize((size_t)d->n_textures);
    for (int ti = 0; ti < d->n_textures; ++ti) {
      const pbrgpu_texture& t = d->textures[ti];
      int W = t.width, H = t.height;
      if (W <= 0 || H <= 0 || !t.texels || W > 32768 || H > 32768) return fail("pbrgpu_scene_create: bad texture size");
      if (t.wrap < PBRGPU_WRAP_REPEAT || t.wrap > PBRGPU_WRAP_CLAMP) return fail("pbrgpu_scene_create: bad texture wrap mode");
      TexDev& td = table[(size_t)ti];
      std::vector<float> prev(t.texels, t.texels + (size_t)W * H * 3), cur;
      if (!host::is_pow2(W) || !host::is_pow2(H))  // MIPMap ctor: Lanczos resampling up to powers of two (MIPMap.h:93-144)
        prev = host::resample_to_pow2(t.texels, t.width, t.height, t.wrap, &W, &H);
      td.width = W; td.height = H; td.doTrilinear = t.do_trilinear ? 1 : 0; td.wrap = t.wrap; td.maxAnisotropy = t.max_anisotropy;
      td.su = t.su; td.sv = t.sv; td.du = t.du; td.dv = t.dv;
      int nLevels = 1;
      for (int m = std::max(W, H); m > 1; m >>= 1) ++nLevels;  // 1 + Log2Int(max(w, h))
      td.nLevels = nLevels;
      std::vector<float4> all;
      int pw = W, ph = H;
      for (int lv = 0; lv < nLevels; ++lv) {
        if (lv > 0) {
          const int sRes = std::max(1, pw / 2), tRes = std::max(1, ph / 2);
          cur.assign((size_t)sRes * tRes * 3, 0.f);
          auto texel = [&](int a, int b, int c) -> float {  // MIPMap::Texel on the previous level
            if (t.wrap == PBRGPU_WRAP_REPEAT) { a = mod_i(a, pw); b = mod_i(b, ph); }
            else if (t.wrap == PBRGPU_WRAP_CLAMP) { a = std::min(std::max(a, 0), pw - 1); b = std::min(std::max(b, 0), ph - 1); }
            else if (a < 0 || a >= pw || b < 0 || b >= ph) return 0.f;
            return prev[3 * ((size_t)b * pw + a) + c];
          };
          for (int tt = 0; tt < tRes; ++tt)
            for (int u = 0; u < sRes; ++u)
              for (int c = 0; c < 3; ++c)
                cur[3 * ((size_t)tt * sRes + u) + c] =
                    .25f * (((texel(2 * u, 2 * tt, c) + texel(2 * u + 1, 2 * tt, c)) + texel(2 * u, 2 * tt + 1, c)) + texel(2 * u + 1, 2 * tt + 1, c));
          prev.swap(cur);
          pw = sRes; ph = tRes;
        }
        td.levelOffset[lv] = (uint32_t)all.size();
        for (size_t k = 0; k < (size_t)pw * ph; ++k) all.push_back(make_float4(prev[3 * k], prev[3 * k + 1], prev[3 * k + 2], 0.f));
      }
      CK(s->texPyramids[(size_t)ti].upload(all, st));
      CK(cudaStreamSynchronize(st));
      td.texels = s->texPyramids[(size_t)ti].p;
    }
    CK(s->texTable.upload(table, st));
    std::vector<float> lut(kWeightLutSize);
    for (int i = 0; i < kWeightLutSize; ++i) {
      float alpha = 2;
      float r2 = float(i) / float(kWeightLutSize - 1);
      lut[i] = std::exp(-alpha * r2) - std::exp(-alpha);
    }
    CK(s->ewaLut.upload(lut, st));
    CK(cudaStreamSynchronize(st));
    s->sc.textures = s->texTable.p;
    s->sc.ewaLut = s->ewaLut.p;
  }
  CK(s->mats.upload(mats, st));

  // scene bound -> InfiniteAreaLight::Preprocess (BoundingSphere, Geometry.h)
  const Aabb& wb = s->bvh.bounds;
  V3 center((wb.lo[0] + wb.hi[0]) * 0.5f, (wb.lo[1] + wb.hi[1]) * 0.5f, (wb.lo[2] + wb.hi[2]) * 0.5f);
  bool inside = center.x >= wb.lo[0] && center.x <= wb.hi[0] && center.y >= wb.lo[1] && center.y <= wb.hi[1] &&
                center.z >= wb.lo[2] && center.z <= wb.hi[2];
  V3 dv(center.x - wb.hi[0], center.y - wb.hi[1], center.z - wb.hi[2]);
  s->sc.worldCenter = center;
  s->sc.worldRadius = inside ? std::sqrt(dv.x * dv.x + dv.y * dv.y + dv.z * dv.z) : 0.f;

  // lights
  s->lights.assign(d->lights, d->lights + d->n_lights);
  s->lightPower.assign((size_t)d->n_lights, 0.f);
  s->sc.nInfinite = 0;
  int envUser = -1;
  for (int i = 0; i < d->n_lights; ++i) {
    pbrgpu_light& l = s->lights[i];
    switch (l.type) {
      case PBRGPU_LIGHT_POINT: {
        Spec p = (4 * kPi) * Spec(l.L[0], l.L[1], l.L[2]);
        s->lightPower[i] = p.y();
        break;
      }
      case PBRGPU_LIGHT_AREA_TRI: {
        if (l.tri < 0 || l.tri >= nTriangles) return fail("pbrgpu_scene_create: area light triangle out of range");
        if (!d->tri_light || d->tri_light[l.tri] != i) return fail("pbrgpu_scene_create: tri_light / light.tri mismatch");
        const float* a = d->verts + 3 * (size_t)d->indices[3 * (size_t)l.tri];
        const float* b = d->verts + 3 * (size_t)d->indices[3 * (size_t)l.tri + 1];
        const float* c = d->verts + 3 * (size_t)d->indices[3 * (size_t)l.tri + 2];
        V3 p0(a[0], a[1], a[2]), p1(b[0], b[1], b[2]), p2(c[0], c[1], c[2]);
        float area = 0.5f * length(cross(p1 - p0, p2 - p0));
        Spec p = ((l.two_sided ? 2.f : 1.f) * Spec(l.L[0], l.L[1], l.L[2])) * area * kPi;
        s->lightPower[i] = p.y();
        // device-side constants of this light (the light_to_world slots are unused by area lights):
        //   [0..2] Triangle::Sample's normal (Triangle.cpp:371), [3] Area(), [4..6] Triangle::Intersect's normal
        const bool flip = d->tri_flags && (d->tri_flags[l.tri] & 1);
        V3 nS = normalize(cross(p1 - p0, p2 - p0));
        if (flip) nS = nS * -1.f;
        V3 nI, ss, ts;
        const float uvDefault[6] = {0.f, 0.f, 1.f, 0.f, 1.f, 1.f};
        triangle_frame(p0, p1, p2, uvDefault, flip, &nI, &ss, &ts);
        l.light_to_world[0] = nS.x; l.light_to_world[1] = nS.y; l.light_to_world[2] = nS.z; l.light_to_world[3] = area;
        l.light_to_world[4] = nI.x; l.light_to_world[5] = nI.y; l.light_to_world[6] = nI.z;
        //   [8..10] Normalize(Cross(p1 - p0, p2 - p0)) WITHOUT the orientation flip: Triangle::Sample of a mesh with shading
        //   normals face-forwards this one to the interpolated normal instead (Triangle.cpp:371-377)
        {
          const V3 nU = normalize(cross(p1 - p0, p2 - p0));
          l.light_to_world[8] = nU.x; l.light_to_world[9] = nU.y; l.light_to_world[10] = nU.z;
        }
        l.tri = s->leafOf[l.tri];
        break;
      }
      case PBRGPU_LIGHT_AREA_SPHERE: {
        if (l.tri < 0 || l.tri >= d->n_spheres) return fail("pbrgpu_scene_create: sphere light index out of range");
        const pbrgpu_sphere& sp = d->spheres[l.tri];
        if (sp.light != i) return fail("pbrgpu_scene_create: sphere.light / light.tri mismatch");
        const float area = 4 * kPi * sp.radius * sp.radius;  // Sphere::Area (Sphere.cpp:31)
        Spec p = ((l.two_sided ? 2.f : 1.f) * Spec(l.L[0], l.L[1], l.L[2])) * area * kPi;  // DiffuseAreaLight::Power
        s->lightPower[i] = p.y();
        l.pos[0] = sp.center[0]; l.pos[1] = sp.center[1]; l.pos[2] = sp.center[2];
        l.world_radius = sp.radius;
        l.light_to_world[0] = sp.reverse_orientation ? 1.f : 0.f;
        l.tri = s->leafOf[nTriangles + l.tri];
        break;
      }
      case PBRGPU_LIGHT_INFINITE:
      case PBRGPU_LIGHT_SKYBOX:
        if (envUser >= 0) return fail("pbrgpu_scene_create: at most one environment-map light is supported");
        envUser = i;
        if (s->sc.nInfinite >= 4) return fail("pbrgpu_scene_create: too many infinite lights");
        s->sc.infinite[s->sc.nInfinite++] = i;
        break;
      default: return fail("pbrgpu_scene_create: unknown light type");
    }
  }
  if (envUser >= 0) {
    int W = d->env_width, H = d->env_height;
    if (!d->env_texels || W <= 0 || H <= 0) return fail("pbrgpu_scene_create: environment light without env_texels");
    const pbrgpu_light& l = s->lights[envUser];
    std::vector<float> tex(d->env_texels, d->env_texels + (size_t)W * H * 3);
    if (l.type == PBRGPU_LIGHT_INFINITE) {
      for (size_t k = 0; k < (size_t)W * H; ++k)
        for (int c = 0; c < 3; ++c) tex[3 * k + c] = l.L[c] * tex[3 * k + c];  // InfiniteAreaLight.cpp:20-31
      // Lmap = MIPMap(resolution, texels): a map whose sides are not powers of two is Lanczos-resampled (Repeat wrap) and the
      // sampling distribution is built over the RESAMPLED size (InfiniteAreaLight.cpp:40-59).  SkyBoxLight reads its raw image.
      if (!host::is_pow2(W) || !host::is_pow2(H)) {
        std::vector<float> rs = host::resample_to_pow2(tex.data(), W, H, PBRGPU_WRAP_REPEAT, &W, &H);
        tex.swap(rs);
      }
    }
    s->sc.env.width = W;
    s->sc.env.height = H;
    if (l.type == PBRGPU_LIGHT_INFINITE) {
      // sampling image (InfiniteAreaLight.cpp:44-59) and Distribution2D (Sampling.cpp:121-133)
      std::vector<float> img((size_t)W * H);
      for (int v = 0; v < H; ++v) {
        float vp = (v + .5f) / (float)H;
        float sinTheta = std::sin(kPi * (v + .5f) / H);
        for (int u = 0; u < W; ++u) {
          float up = (u + .5f) / (float)W;
          float rgb[3];
          env_bilinear_host(tex, W, H, up, vp, rgb);
          float y = 0.212671f * rgb[0] + 0.715160f * rgb[1] + 0.072169f * rgb[2];
          img[(size_t)u + (size_t)v * W] = y * sinTheta;
        }
      }
      std::vector<float> cFunc((size_t)W * H), cCdf((size_t)(W + 1) * H), cInt((size_t)H);
      for (int v = 0; v < H; ++v) {
        host::Dist1D row;
        row.build(&img[(size_t)v * W], W);
        memcpy(&cFunc[(size_t)v * W], row.func.data(), sizeof(float) * W);
        memcpy(&cCdf[(size_t)v * (W + 1)], row.cdf.data(), sizeof(float) * (W + 1));
        cInt[v] = row.funcInt;
      }
      host::Dist1D marg;
      marg.build(cInt.data(), H);
      CK(s->condFunc.upload(cFunc, st)); CK(s->condCdf.upload(cCdf, st)); CK(s->condInt.upload(cInt, st));
      CK(s->margFunc.upload(marg.func, st)); CK(s->margCdf.upload(marg.cdf, st));
      s->sc.env.condFunc = s->condFunc.p; s->sc.env.condCdf = s->condCdf.p; s->sc.env.condInt = s->condInt.p;
      s->sc.env.margFunc = s->margFunc.p; s->sc.env.margCdf = s->margCdf.p;
      s->sc.env.margInt = marg.funcInt;
      // Power() (InfiniteAreaLight.cpp:63-67) = 4pi * pi * r^2 * Lmap->Lookup((.5,.5), .5f): a trilinear lookup, so
      // the MIP pyramid (MIPMap.h:146-168: 2x2 box filter per level) is rebuilt here, host-side, once.
      {
        std::vector<std::vector<float>> pyr;
        std::vector<int> pw, ph;
        pyr.push_back(tex); pw.push_back(W); ph.push_back(H);
        int nLevels = 1;
        for (int m = std::max(W, H); m > 1; m >>= 1) ++nLevels;
        for (int i = 1; i < nLevels; ++i) {
          int sRes = std::max(1, pw[i - 1] / 2), tRes = std::max(1, ph[i - 1] / 2);
          std::vector<float> lv((size_t)sRes * tRes * 3);
          const std::vector<float>& pr = pyr[i - 1];
          const int w0 = pw[i - 1], h0 = ph[i - 1];
          auto texel = [&](int a, int b, int c) { return pr[3 * ((size_t)mod_i(b, h0) * w0 + mod_i(a, w0)) + c]; };
          for (int t = 0; t < tRes; ++t)
            for (int u = 0; u < sRes; ++u)
              for (int c = 0; c < 3; ++c)
                lv[3 * ((size_t)t * sRes + u) + c] =
                    .25f * (((texel(2 * u, 2 * t, c) + texel(2 * u + 1, 2 * t, c)) + texel(2 * u, 2 * t + 1, c)) + texel(2 * u + 1, 2 * t + 1, c));
          pyr.push_back(std::move(lv)); pw.push_back(sRes); ph.push_back(tRes);
        }
        auto tri = [&](int level, float out[3]) {  // MIPMap::triangle(level, (.5,.5))
          level = std::max(0, std::min(nLevels - 1, level));
          env_bilinear_host(pyr[level], pw[level], ph[level], .5f, .5f, out);
        };
        const float invLog2 = 1.442695040888963387004650940071f;
        float level = nLevels - 1 + std::log(std::max(.5f, (float)1e-8)) * invLog2;  // MIPMap::Lookup(st, width = .5f)
        float rgb[3];
        if (level < 0) tri(0, rgb);
        else if (level >= nLevels - 1) { for (int c = 0; c < 3; ++c) rgb[c] = pyr[nLevels - 1][c]; }
        else {
          int iLevel = (int)std::floor(level);
          float delta = level - iLevel, a[3], b[3];
          tri(iLevel, a); tri(iLevel + 1, b);
          for (int c = 0; c < 3; ++c) rgb[c] = (1 - delta) * a[c] + delta * b[c];
        }
        const float r = s->sc.worldRadius;
        const float k = (4 * kPi) * kPi * r * r;
        Spec pwr = k * Spec(rgb[0], rgb[1], rgb[2]);
        s->lightPower[envUser] = pwr.y();
      }
    }
    CK(s->envTexels.upload(tex, st));
    s->sc.env.texels = s->envTexels.p;
  }
  CK(s->dlights.upload(s->lights, st));

  s->sc.tf = s->tfr.p;
  s->sc.ts0 = s->ts0.p; s->sc.ts1 = s->ts1.p; s->sc.ts2 = s->ts2.p;
  s->sc.tn0 = s->tn0.p; s->sc.tn1 = s->tn1.p; s->sc.tn2 = s->tn2.p; s->sc.tuv0 = s->tuv0.p; s->sc.tuv1 = s->tuv1.p;
  s->sc.mats = s->mats.p;
  s->sc.lights = s->dlights.p;
  s->sc.nTris = nT;
  s->sc.nSpheres = d->n_spheres;
  s->sc.f32one = 0x3F800000u;
  s->sc.nLights = d->n_lights;

  // sampler tables
  std::vector<int32_t> primes, sums;
  host::prime_tables(&primes, &sums);
  CK(s->primes.upload(primes, st));
  CK(s->primeSums.upload(sums, st));
  s->sc.hp.primes = s->primes.p;
  s->sc.hp.primeSums = s->primeSums.p;
  std::vector<uint32_t> magic;
  host::prime_magic(primes, &magic);
  CK(s->primeMagic.upload(magic, st));
  s->sc.hp.primeMagic = reinterpret_cast<const uint2*>(s->primeMagic.p);
  CK(cudaStreamSynchronize(st));
  *out = s.release();
  return 0;
}

void pbrgpu_scene_destroy(pbrgpu_scene* s) {
  if (!s) return;
  for (pbrgpu_scene* r : s->replicas) pbrgpu_scene_destroy(r);
  s->replicas.clear();
  cudaSetDevice(s->ctx->device);
  delete s;
}

static int camera_set_single(pbrgpu_scene* s, const pbrgpu_camera_desc* c) {
  if (!s || !c) return fail("pbrgpu_camera_set: NULL argument");
  if (c->type != PBRGPU_CAMERA_PERSPECTIVE && c->type != PBRGPU_CAMERA_ORTHOGRAPHIC) return fail("pbrgpu_camera_set: bad type");
  if (c->width <= 0 || c->height <= 0) return fail("pbrgpu_camera_set: bad resolution");
  CameraDev& cd = s->sc.cam;
  cd.type = c->type; cd.width = c->width; cd.height = c->height;
  cd.lensRadius = c->lens_radius; cd.focalDistance = c->focal_distance;
  memcpy(cd.r2c, c->raster_to_camera, sizeof(cd.r2c));
  memcpy(cd.c2w, c->camera_to_world, sizeof(cd.c2w));
  if (c->medium_plus1 < 0 || c->medium_plus1 > s->sc.nMedia) return fail("pbrgpu_camera_set: camera medium out of range");
  cd.medium = c->medium_plus1 - 1;
  s->haveCamera = true;
  return 0;
}

static int sampler_set_single(pbrgpu_scene* s, const pbrgpu_sampler_desc* d) {
  if (!s || !d) return fail("pbrgpu_sampler_set: NULL argument");
  if (d->samples_per_pixel <= 0) return fail("pbrgpu_sampler_set: samples_per_pixel <= 0");
  CK(cudaSetDevice(s->ctx->device));
  HaltonParams& hp = s->sc.hp;
  if (d->type == PBRGPU_SAMPLER_SOBOL) {
    // Sobol' (parity unpinned: the reference has the tables, no sampler).  The tables are the reference's data and
    // come through the descriptor; CPU twin = host/SobolSampler.h.
    if (!d->sobol_matrices32 || !d->sobol_vdc || !d->sobol_vdc_inv || d->sobol_n_dims <= 0 || d->sobol_n_dims > 1024)
      return fail("pbrgpu_sampler_set: Sobol needs the reference's SobolMatrices32 / VdCSobolMatrices(Inv) tables");
    int dx = d->sample_bounds[2] - d->sample_bounds[0], dy = d->sample_bounds[3] - d->sample_bounds[1];
    int res = 1, lg = 0;
    while (res < std::max(dx, dy)) { res <<= 1; ++lg; }
    if (lg > 26) return fail("pbrgpu_sampler_set: Sobol resolution too large");
    std::vector<uint32_t> m32(d->sobol_matrices32, d->sobol_matrices32 + (size_t)d->sobol_n_dims * kSobolMatrixSize);
    std::vector<uint64_t> v(d->sobol_vdc, d->sobol_vdc + kSobolMatrixSize), vi(d->sobol_vdc_inv, d->sobol_vdc_inv + kSobolMatrixSize);
    CK(s->sobolMat.upload(m32, s->ctx->stream));
    CK(s->sobolVdc.upload(v, s->ctx->stream));
    CK(s->sobolVdcInv.upload(vi, s->ctx->stream));
    CK(cudaStreamSynchronize(s->ctx->stream));
    hp.type = PBRGPU_SAMPLER_SOBOL;
    hp.sobolDims = d->sobol_n_dims;
    hp.resolution = res;
    hp.log2Resolution = lg;
    hp.boundsMinX = d->sample_bounds[0];
    hp.boundsMinY = d->sample_bounds[1];
    hp.sobolMatrices32 = s->sobolMat.p;
    hp.sobolVdc = s->sobolVdc.p;
    hp.sobolVdcInv = s->sobolVdcInv.p;
    hp.sampleAtPixelCenter = 0;
    hp.sampleStride = 1;
    int64_t spp = 1;
    while (spp < d->samples_per_pixel) spp <<= 1;
    s->spp = spp;
    s->haveSampler = true;
    return 0;
  }
  if (d->type != PBRGPU_SAMPLER_HALTON) return fail("pbrgpu_sampler_set: unknown sampler type");
  hp.type = PBRGPU_SAMPLER_HALTON;
  int32_t p7[7];
  host::halton_params(d->sample_bounds, p7);
  hp.baseScales[0] = p7[0]; hp.baseScales[1] = p7[1];
  hp.baseExponents[0] = p7[2]; hp.baseExponents[1] = p7[3];
  hp.sampleStride = p7[4];
  hp.multInverse[0] = p7[5]; hp.multInverse[1] = p7[6];
  hp.sampleAtPixelCenter = d->sample_at_pixel_center;
  std::vector<uint16_t> perms;
  if (d->perms && d->n_perms > 0) perms.assign(d->perms, d->perms + d->n_perms);
  else {
    std::vector<int32_t> primes, sums;
    host::prime_tables(&primes, &sums);
    perms = host::radical_inverse_permutations(primes);
  }
  CK(s->perms.upload(perms, s->ctx->stream));
  {  // per-dimension float constants of (Scrambled)RadicalInverseSpecialized (LowDiscrepancy.cpp:214,2303,2314)
    std::vector<int32_t> primes, sums;
    host::prime_tables(&primes, &sums);
    std::vector<float> dc(2 * primes.size());
    for (size_t i = 0; i < primes.size(); ++i) {
      const float invBase = (float)1 / (float)primes[i];
      float tail = 0.f;
      if ((size_t)sums[i] < perms.size()) tail = invBase * (float)(int)perms[(size_t)sums[i]] / (1 - invBase);
      dc[2 * i] = invBase;
      dc[2 * i + 1] = tail;
    }
    CK(s->dimConst.upload(dc, s->ctx->stream));
    hp.dimConst = reinterpret_cast<const float2*>(s->dimConst.p);
  }
  CK(cudaStreamSynchronize(s->ctx->stream));
  hp.perms = s->perms.p;
  hp.nPerms = (int64_t)perms.size();
  s->spp = d->samples_per_pixel;
  s->haveSampler = true;
  return 0;
}

int pbrgpu_texture_eval(pbrgpu_scene* s, int texture, const float* rows, int64_t n, float* out) {
  if (s && !s->replicas.empty()) s = s->replicas[0];  // group scene: the hooks run on the first device
  if (!s || !rows || !out || n < 0) return fail("pbrgpu_texture_eval: bad argument");
  if (texture < 0 || (size_t)texture >= s->texPyramids.size()) return fail("pbrgpu_texture_eval: texture index out of range");
  if (n == 0) return 0;
  pbrgpu_ctx* c = s->ctx;
  CK(cudaSetDevice(c->device));
  DevBuf<float> din, dout;
  CK(din.alloc((size_t)n * 8)); CK(dout.alloc((size_t)n * 4));
  CK(cudaMemcpyAsync(din.p, rows, sizeof(float) * 8 * (size_t)n, cudaMemcpyHostToDevice, c->stream));
  k_texture_eval<<<grid_for(c, (uint64_t)n, 8), kBlock, 0, c->stream>>>(s->sc, texture, din.p, n, dout.p);
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(out, dout.p, sizeof(float) * 4 * (size_t)n, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return 0;
}

int pbrgpu_light_eval(pbrgpu_scene* s, int light, const float* rows, int64_t n, float* out) {
  if (s && !s->replicas.empty()) s = s->replicas[0];
  if (!s || !rows || !out || n < 0) return fail("pbrgpu_light_eval: bad argument");
  if (light < 0 || light >= s->sc.nLights) return fail("pbrgpu_light_eval: light index out of range");
  if (n == 0) return 0;
  pbrgpu_ctx* c = s->ctx;
  CK(cudaSetDevice(c->device));
  if (build_light_distribution(s, PBRGPU_LIGHTSTRATEGY_UNIFORM)) return 1;
  DevBuf<float> din, dout;
  CK(din.alloc((size_t)n * 16)); CK(dout.alloc((size_t)n * 16));
  CK(cudaMemcpyAsync(din.p, rows, sizeof(float) * 16 * (size_t)n, cudaMemcpyHostToDevice, c->stream));
  k_light_eval<<<grid_for(c, (uint64_t)n, 8), kBlock, 0, c->stream>>>(s->sc, light, din.p, n, dout.p);
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(out, dout.p, sizeof(float) * 16 * (size_t)n, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return 0;
}

int pbrgpu_scene_info(pbrgpu_scene* s, int64_t out[8]) {
  if (s && !s->replicas.empty()) s = s->replicas[0];  // group scene: the hooks run on the first device
  if (!s || !out) return fail("pbrgpu_scene_info: NULL argument");
  memset(out, 0, sizeof(int64_t) * 8);
  out[0] = (int64_t)s->bvh.nodes.size();
  out[1] = s->bvh.nLeaves;
  out[2] = s->bvh.maxLeafSize;
  out[3] = (int64_t)(s->bvh.nodes.size() * sizeof(Bvh8Node));
  out[4] = (int64_t)s->sc.nTris * 48;
  out[5] = s->bvh.maxDepth;
  // FNV-1a over the built nodes and the leaf order: the build is deterministic and independent of its thread count
  uint64_t h = 1469598103934665603ull;
  auto mix = [&h](const void* p, size_t n) {
    const unsigned char* b = static_cast<const unsigned char*>(p);
    for (size_t i = 0; i < n; ++i) { h ^= b[i]; h *= 1099511628211ull; }
  };
  mix(s->bvh.nodes.data(), s->bvh.nodes.size() * sizeof(Bvh8Node));
  mix(s->bvh.triOrder.data(), s->bvh.triOrder.size() * sizeof(int32_t));
  out[6] = (int64_t)(h >> 1);
  return 0;
}

static int check_render_args(pbrgpu_scene* s, const pbrgpu_render_desc* rd) {
  if (!s || !rd) return fail("pbrgpu_render: NULL argument");
  if (!s->haveCamera || !s->haveSampler) return fail("pbrgpu_render: camera / sampler not set");
  if (rd->x1 <= rd->x0 || rd->y1 <= rd->y0) return fail("pbrgpu_render: empty pixel region");
  if (rd->s1 < rd->s0 || rd->s0 < 0) return fail("pbrgpu_render: bad sample range");
  if (rd->max_depth < 0 || rd->max_depth > 1000) return fail("pbrgpu_render: bad max_depth");
  if (rd->integrator < PBRGPU_INTEGRATOR_PATH || rd->integrator > PBRGPU_INTEGRATOR_VOLPATH) return fail("pbrgpu_render: unknown integrator");
  if (rd->integrator != PBRGPU_INTEGRATOR_VOLPATH && s->hasNullMat)
    return fail("pbrgpu_render: the scene has triangles without a material (medium boundaries); only the VolPath integrator renders them");
  if (rd->integrator == PBRGPU_INTEGRATOR_VOLPATH && rd->max_depth > (int)kMetaBounceMask - 1)
    return fail("pbrgpu_render: VolPath max_depth > 126");
  if (s->bvh.maxDepth >= kSmemStack + kLocalStack) return fail("pbrgpu_render: BVH deeper than the traversal stack");
  return 0;
}

static int fetch_stats(pbrgpu_scene* s, cudaStream_t st, pbrgpu_stats* stats) {
  pbrgpu_ctx* c = s->ctx;
  Counters hc{};
  CK(cudaMemcpyAsync(&hc, s->ctr.p, sizeof(Counters), cudaMemcpyDeviceToHost, st));
  Counters hc2{};
  if (s->altUsed) CK(cudaMemcpyAsync(&hc2, s->alt->ctr.p, sizeof(Counters), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  hc.raysClosest += hc2.raysClosest; hc.raysShadow += hc2.raysShadow; hc.raysPath += hc2.raysPath;
  hc.nodesVisited += hc2.nodesVisited; hc.trisTested += hc2.trisTested;
  for (int k = 0; k < 8; ++k) hc.laneProf[k] += hc2.laneProf[k];
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, c->ev[0], c->ev[1]));
  memset(stats, 0, sizeof(*stats));
  stats->camera_samples = s->last.samples;
  stats->rays_closest = hc.raysClosest;
  stats->rays_shadow = hc.raysShadow;
  stats->nodes_visited = hc.nodesVisited;
  stats->tris_tested = hc.trisTested;
  stats->kernel_launches = s->last.launches;
  stats->ms_total = ms;
  stats->rays_path = hc.raysPath;
  if (getenv("PBRGPU_LANE_LOG") && hc.laneProf[0])  // development aid (COUNT build of k_trace): where the lanes of a warp go
    fprintf(stderr, "k_trace lanes: %llu warp iterations; per iteration busy %.2f, node step %.2f, triangle test %.2f (ran in %.1f %% of iterations), "
            "triangle hit %.3f; %llu fetches of %.2f rays\n", hc.laneProf[0], (double)hc.laneProf[1] / hc.laneProf[0], (double)hc.laneProf[2] / hc.laneProf[0],
            (double)hc.laneProf[3] / hc.laneProf[0], 100.0 * hc.laneProf[7] / hc.laneProf[0], (double)hc.laneProf[4] / hc.laneProf[0], hc.laneProf[5],
            (double)hc.laneProf[6] / std::max<unsigned long long>(hc.laneProf[5], 1));
  if (getenv("PBRGPU_VOL_LOG")) fprintf(stderr, "vol shadow rays: max hops %u, stopped at the cap %u\n", hc.volMaxHops, hc.volCapHits);
  Timeline& tl = s->last.tl;
  if (tl.on) {
    double t[T_COUNT];
    uint64_t n[T_COUNT];
    CK(tl.collect(t, n));
    stats->ms_raygen = t[T_RAYGEN]; stats->ms_trace_closest = t[T_CLOSEST]; stats->ms_shade = t[T_SHADE];
    stats->ms_trace_shadow = t[T_SHADOW]; stats->ms_trace_mis = t[T_MIS]; stats->ms_film = t[T_FILM];
    stats->n_trace_closest_launches = n[T_CLOSEST];
  }
  return 0;
}

static int render_device_single(pbrgpu_scene* s, const pbrgpu_render_desc* rd, float* d_rgba_sum, void* stream_, pbrgpu_stats* stats) {
  if (check_render_args(s, rd)) return 1;
  if (!d_rgba_sum) return fail("pbrgpu_render_device: NULL accumulator");
  pbrgpu_ctx* c = s->ctx;
  CK(cudaSetDevice(c->device));
  cudaStream_t st = stream_ ? (cudaStream_t)stream_ : c->stream;
  if (build_light_distribution(s, rd->light_strategy)) return 1;
  if (s->traceBlocksPerSm <= 0) {
    int nb = trace_blocks_per_sm();
    if (const char* e = getenv("PBRGPU_TRACE_BLOCKS_PER_SM")) nb = atoi(e);
    s->traceBlocksPerSm = std::max(1, nb);
    if (const char* e = getenv("PBRGPU_FETCH_MIN_IDLE")) s->fetchMinIdle = std::max(1, std::min(32, atoi(e)));
    if (const char* e = getenv("PBRGPU_WAVE_LOG2")) s->waveLog2 = std::max(14, std::min(25, atoi(e)));
    if (const char* e = getenv("PBRGPU_WAVES_IN_FLIGHT")) s->wavesInFlight = std::max(1, std::min(2, atoi(e)));
  }

  const int W = rd->x1 - rd->x0, H = rd->y1 - rd->y0;
  const int64_t nS = rd->s1 - rd->s0;
  const bool count = (rd->flags & PBRGPU_RENDER_COUNTERS) != 0;
  const bool pathInteg = rd->integrator == PBRGPU_INTEGRATOR_PATH;
  const bool volInteg = rd->integrator == PBRGPU_INTEGRATOR_VOLPATH;
  uint32_t raysPerHit = 1;
  if (rd->integrator == PBRGPU_INTEGRATOR_WHITTED) raysPerHit = (uint32_t)std::max<size_t>(1, s->lights.size());
  if (rd->integrator == PBRGPU_INTEGRATOR_DIRECT_ALL) {
    raysPerHit = 0;
    for (const auto& l : s->lights) {
      if (l.n_samples < 1) return fail("pbrgpu_render: light with nSamples < 1");
      raysPerHit += (uint32_t)l.n_samples;
    }
    raysPerHit = std::max(1u, raysPerHit);
  }
  if (raysPerHit > 4096) return fail("pbrgpu_render: too many light samples per hit");
  const uint32_t capMax = std::max(1u << 14, std::min(1u << s->waveLog2, (2u << s->waveLog2) / raysPerHit));
  // row bands of <= capMax pixels; inside a band, waves of nsWave samples per pixel
  int bandRows = (int)std::max<int64_t>(1, std::min<int64_t>(H, capMax / (uint32_t)W));
  if ((uint64_t)W > capMax) return fail("pbrgpu_render: region wider than the wave capacity");
  int nsWaveMax = (int)std::max<int64_t>(1, capMax / ((uint64_t)bandRows * W));
  if (rd->samples_per_wave > 0) nsWaveMax = std::min(nsWaveMax, rd->samples_per_wave);
  // Two waves in flight: consecutive waves alternate between two wave sets and two streams, so that the tail of one
  // wave's kernels (a persistent k_trace launch ends with its longest rays, every launch with a partial last wave of
  // blocks) overlaps the other wave's work; the film is updated in wave order (an event chain), sums stay deterministic.
  // Measured on the bench workload: 2 x 16 spp in flight 972.7 Msamples/s against 935.2 one after the other.  A call that
  // fits ONE wave is split into two half waves.  Not for VolPath (its loop reads counters back per iteration) and not
  // under the per-kernel timers / counters of the diagnostics flags.
  bool dual = s->wavesInFlight >= 2 && !volInteg && !(rd->flags & (PBRGPU_RENDER_TIMERS | PBRGPU_RENDER_COUNTERS));
  if (dual && bandRows >= H && nS <= nsWaveMax) {
    if (nS >= 2 && (uint64_t)W * H * (uint64_t)nS >= (1u << 20)) nsWaveMax = (int)((nS + 1) / 2);
    else dual = false;
  }
  uint32_t cap = (uint32_t)((uint64_t)bandRows * W * std::min<int64_t>(nsWaveMax, std::max<int64_t>(nS, 1)));
  if (ensure_wave(s, cap, raysPerHit, volInteg)) return 1;
  if (dual && ensure_alt_wave(s, cap, raysPerHit, &dual)) return 1;
  if (dual && !c->stream2) {
    CK(cudaStreamCreateWithFlags(&c->stream2, cudaStreamNonBlocking));
    for (cudaEvent_t* e : {&c->evFork, &c->evJoin, &c->evFilm[0], &c->evFilm[1]}) CK(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
  }
  s->altUsed = dual;

  RenderParams P{};
  P.sc = s->sc;
  fill_wave(s, &P.ws);
  P.integrator = rd->integrator;
  P.maxDepth = rd->max_depth;
  P.rrThreshold = rd->rr_threshold;
  // GlobalSampler::arrayEndDim (Sampler.cpp:100-101): DirectLighting/UniformSampleAll requests two 2D arrays per
  // light per recursion depth (DirectLightingIntegrator.cpp:8-26); nobody else requests arrays
  P.arrayEndDim = 5;
  if (rd->integrator == PBRGPU_INTEGRATOR_DIRECT_ALL) P.arrayEndDim = 5 + 2 * (2 * (int)s->lights.size() * rd->max_depth);
  P.atomicL = (pathInteg || volInteg) ? 0 : 1;
  P.volShadowAnyHit = s->hasNullMat ? 0 : 1;
  // VolPath tail: once fewer than volInLaneBelow paths are alive, k_trace continues a path across
  // material-less medium boundaries inside its lane (k_trace.cuh vol_path_next_segment).  With millions of live paths the
  // separate full-width kernels are faster (measured on the dragon-class smoke volume: in-lane from the first iteration
  // 150 ms, never 143 ms); in the tail every crossing otherwise costs a whole wavefront iteration for a handful of paths.
  uint32_t volInLaneBelow = 2u << 20;
  if (const char* e = getenv("PBRGPU_VOL_INLANE_BELOW")) volInLaneBelow = (uint32_t)std::max(0ll, std::min(atoll(e), 0xffffffffll));
  P.volInLane = (volInteg && volInLaneBelow == 0xffffffffu) ? 1 : 0;
  P.diffScale = 1 / std::sqrt((float)s->spp);
  P.countStats = count ? 1 : 0;
  P.fetchMinIdle = s->fetchMinIdle;

  // traversal data persists in L2 for the duration of the render (window = the geometry slab; when the slab is larger than
  // the carve-out, hitRatio keeps a random carve-out-sized fraction of its lines persisting instead of thrashing)
  bool l2Window = false;
  if (c->l2PersistBytes > 0 && s->geom.n > 0) {
    cudaStreamAttrValue av{};
    const size_t bytes = std::min(s->geom.n * sizeof(float4), c->l2WindowMax);
    av.accessPolicyWindow.base_ptr = s->geom.p;
    av.accessPolicyWindow.num_bytes = bytes;
    av.accessPolicyWindow.hitRatio = (float)std::min(1.0, (double)c->l2PersistBytes / (double)bytes);
    av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    if (cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &av) == cudaSuccess) l2Window = true;
    else cudaGetLastError();
  }
  if (!(rd->flags & PBRGPU_RENDER_ACCUMULATE)) CK(cudaMemsetAsync(d_rgba_sum, 0, sizeof(float) * 4 * (size_t)W * H, st));
  CK(cudaMemsetAsync(s->ctr.p, 0, sizeof(Counters), st));  // (a memset, not a pageable H2D copy: never blocks the host)
  CK(cudaEventRecord(c->ev[0], st));
  uint64_t launches = 0;
  Timeline tl{c, st, (rd->flags & PBRGPU_RENDER_TIMERS) != 0};
  cudaStream_t const stMain = st;
  WaveDev wsSet[2];
  wsSet[0] = P.ws;
  if (dual) {
    fill_alt_wave(s, &wsSet[1]);
    CK(cudaMemsetAsync(s->alt->ctr.p, 0, sizeof(Counters), stMain));
    CK(cudaEventRecord(c->evFork, stMain));  // the second stream starts behind the film / counter resets (and the caller's earlier work)
    CK(cudaStreamWaitEvent(c->stream2, c->evFork, 0));
  }
  int64_t waveIdx = 0;
  for (int yb = 0; yb < H; yb += bandRows) {
    int rows = std::min(bandRows, H - yb);
    P.x0 = rd->x0; P.y0 = rd->y0 + yb; P.w = W; P.h = rows;
    P.accum = d_rgba_sum + 4 * (size_t)yb * W;
    for (int64_t sb = rd->s0; sb < rd->s1; sb += nsWaveMax) {
      int ns = (int)std::min<int64_t>(nsWaveMax, rd->s1 - sb);
      P.sBegin = sb;
      P.nsWave = ns;
      uint64_t nPaths = (uint64_t)rows * W * ns;
      const int set = dual ? (int)(waveIdx & 1) : 0;
      cudaStream_t st = set ? c->stream2 : stMain;  // (shadows the function's `st` for the body of the wave)
      P.ws = wsSet[set];
      uint32_t g = grid_for(c, nPaths, 8);
      uint32_t gc = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>((nPaths + kClassifyTile - 1) / kClassifyTile, (uint64_t)c->smCount * 8));
      // persistent traversal grid: every SM filled to the kernel's occupancy, never more blocks than rays
      uint32_t gt = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>((nPaths * (pathInteg ? 1 : 1 + raysPerHit) + kTraceBlockThreads - 1) / kTraceBlockThreads,
                                                                      (uint64_t)c->smCount * s->traceBlocksPerSm));
      CK(tl.mark(T_RAYGEN));
      k_wave_begin<<<1, 1, 0, st>>>(P.ws.ctr);
      k_raygen<<<grid_occ(c, k_raygen, kBlock, nPaths), kBlock, 0, st>>>(P);
      launches += 2;
      // bounce b: ONE traversal launch serves the path rays of bounce b and the shadow + MIS rays that the
      // shading of bounce b-1 emitted; then MIS resolve (b-1), classification, material-sorted shading (b).
      if (volInteg) {
        // VolPathIntegrator::Li: the bounce count is path state (crossing a medium boundary repeats the bounce), so the
        // number of iterations is not known up front: max_depth + 1 iterations are launched blind, then the live-path
        // and pending-ray counters are read back after every iteration until the wave has drained.
        //
        // Tail: from the first read-back on, the shadow rays that shading emits are NOT traced by the
        // next iteration's k_trace.  They stay in the shadow queue, which keeps growing (RenderParams::deferShadow = 1),
        // and ONE launch traces all of them when the paths have drained (deferShadow = 2) - a shadow ray marches through
        // up to dozens of medium boundaries inside its lane, and with a few thousand live paths every iteration waited
        // for its longest such chain (1-3 ms, 100+ times).  k_shadow_add then adds the contributions per iteration range.
        const int kMaxExtraIters = 4096;
        const bool volLog = getenv("PBRGPU_VOL_LOG") != nullptr;  // development aid: per-iteration queue sizes + wall time
        const bool deferOk = getenv("PBRGPU_VOL_NO_DEFER") == nullptr;
        const size_t shadowCap = (size_t)s->capacity;  // entries of the shadow-ray arrays this wave may use (raysPerHit == 1)
        // Scenes WITH material-less medium boundaries defer from the first iteration on (one 16-byte read-back per iteration;
        // the queue is flushed whenever the next iteration might overflow it): their shadow rays are the expensive, serial
        // ones, and traced on their own - every lane in the same mode, one longest chain per flush instead of one per
        // iteration - they cost about half.  Other scenes keep the blind launches: their shadow rays are single any-hit segments.
        const int firstRead = (deferOk && s->hasNullMat) ? 0 : rd->max_depth;
        std::vector<uint32_t> ranges;  // deferred shadow queue: start of each iteration's range, then the end
        auto flush_deferred = [&]() -> int {
          if (ranges.empty()) return 0;
          if (ranges.back() > ranges.front()) {
            P.deferShadow = 2;
            CK(tl.mark(T_CLOSEST));
            launch_trace(count, true, gt, st, P);
            launches++;
            CK(tl.mark(T_SHADE));
            for (size_t k = 0; k + 1 < ranges.size(); ++k) {
              const uint32_t b = ranges[k], e = ranges[k + 1];
              const bool last = k + 2 == ranges.size();
              if (e > b || last) {
                k_shadow_add<<<(unsigned)std::max<uint64_t>(1, std::min<uint64_t>(((uint64_t)(e - b) + kBlock - 1) / kBlock, (uint64_t)c->smCount * 8)), kBlock, 0, st>>>(P, b, e, last ? 1 : 0);
                launches++;
              }
            }
          }
          ranges.clear();
          P.deferShadow = 0;
          return 0;
        };
        for (int iter = 0;; ++iter) {
          std::chrono::steady_clock::time_point t0;
          if (volLog) {
            CK(cudaMemcpyAsync(s->hNCur, s->ctr.p, 4 * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
            t0 = std::chrono::steady_clock::now();
          }
          CK(tl.mark(T_SHADE));
          if (s->sc.nMedia > 0) { k_medium_presample<<<g, kBlock, 0, st>>>(P); launches++; }
          CK(tl.mark(T_CLOSEST));
          launch_trace(count, true, gt, st, P);
          if (volLog) {
            CK(cudaStreamSynchronize(st));
            const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
            static Counters prev{};
            Counters hc{};
            CK(cudaMemcpy(&hc, s->ctr.p, sizeof(Counters), cudaMemcpyDeviceToHost));
            if (iter == 0) prev = Counters{};
            const double rays = (double)(hc.raysClosest - prev.raysClosest);
            fprintf(stderr, "vol iter %d: paths %u shadow %u mis %u  k_trace %.3f ms  segments %.0f nodes/seg %.2f tris/seg %.2f  max work of one ray %u%s\n", iter,
                    s->hNCur[0], s->hNCur[2], s->hNCur[3], ms, rays, (hc.nodesVisited - prev.nodesVisited) / std::max(rays, 1.0),
                    (hc.trisTested - prev.trisTested) / std::max(rays, 1.0), hc.maxRayWork, P.deferShadow ? "  (shadow queue deferred)" : "");
            prev = hc;
            CK(cudaMemset(&s->ctr.p->maxRayWork, 0, sizeof(uint32_t)));
          }
          CK(tl.mark(T_MIS));
          if (iter > 0) { k_resolve_mis<<<g, kBlock, 0, st>>>(P); launches++; }
          CK(tl.mark(T_SHADE));
          if (s->sc.nMedia > 0 && !P.volInLane) { k_medium_sample<<<g, kBlock, 0, st>>>(P); launches++; }
          k_classify<<<gc, kBlock, 0, st>>>(P);
          for (int M = 0; M < PBRGPU_MAT_COUNT; ++M)
            if (s->hasMat[M]) {
              const bool tex = s->hasTextures && (M == PBRGPU_MAT_MATTE || M == PBRGPU_MAT_PLASTIC);
              launch_shade(M, true, tex, shade_grid(c, M, true, tex, nPaths), st, P, 0);
              launches++;
            }
          if (s->hasNullMat) { k_shade_null<<<g, kBlock, 0, st>>>(P); launches++; }
          if (s->sc.nMedia > 0) { k_shade_medium<<<grid_occ(c, k_shade_medium, kShadeBlock, nPaths), kShadeBlock, 0, st>>>(P); launches++; }
          if (s->sc.nInfinite > 0) { k_shade_miss<<<g, kBlock, 0, st>>>(P, 0); launches++; }
          k_next_bounce<<<1, 1, 0, st>>>(s->ctr.p);
          launches += 3;
          std::swap(P.ws.cur, P.ws.nxt);
          if (iter >= firstRead) {
            CK(cudaMemcpyAsync(s->hNCur, s->ctr.p, 4 * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));  // nCur, nNext, nShadow, nMis
            CK(cudaStreamSynchronize(st));
            const uint32_t nLive = s->hNCur[0], nSh = s->hNCur[2], nMi = s->hNCur[3];
            P.volInLane = nLive < volInLaneBelow ? 1 : 0;
            if (deferOk) {
              if (ranges.empty()) ranges.push_back(0u);  // deferral starts: the rays just emitted are its first range
              ranges.push_back(nSh);
              const bool drained = nLive + nMi == 0;
              // the next iteration appends at most one shadow ray per live path
              if (drained || (size_t)nSh + nLive > shadowCap || iter >= rd->max_depth + kMaxExtraIters) {
                if (flush_deferred()) return 1;
                if (drained || iter >= rd->max_depth + kMaxExtraIters) break;
                ranges.push_back(0u);
              }
              P.deferShadow = 1;
            } else if (nLive + nSh + nMi == 0 || iter >= rd->max_depth + kMaxExtraIters) break;
          }
        }
        P.deferShadow = 0;
      }
      const int lastBounce = volInteg ? -1 : (pathInteg ? rd->max_depth : std::max(rd->max_depth - 1, 0));
      for (int bounce = 0; bounce <= lastBounce; ++bounce) {
        CK(tl.mark(T_CLOSEST));
        launch_trace(count, false, gt, st, P);
        CK(tl.mark(T_MIS));
        if (bounce > 0) { k_resolve_mis<<<g, kBlock, 0, st>>>(P); launches++; }
        CK(tl.mark(T_SHADE));
        k_classify<<<gc, kBlock, 0, st>>>(P);
        // one shading launch per material type PRESENT in the scene
        for (int M = 0; M < PBRGPU_MAT_COUNT; ++M)
          if (s->hasMat[M]) {
            const bool tex = s->hasTextures && (M == PBRGPU_MAT_MATTE || M == PBRGPU_MAT_PLASTIC);
            if (pathInteg) launch_shade(M, false, tex, shade_grid(c, M, false, tex, nPaths), st, P, bounce);
            else launch_shade_direct(M, tex, g, st, P, bounce);
            launches++;
          }
        if (!pathInteg || s->sc.nInfinite > 0) { k_shade_miss<<<g, kBlock, 0, st>>>(P, bounce); launches++; }
        k_next_bounce<<<1, 1, 0, st>>>(P.ws.ctr);
        launches += 3;
        std::swap(P.ws.cur, P.ws.nxt);
      }
      if (!pathInteg && !volInteg) {  // shadow / MIS rays of the last shaded depth
        CK(tl.mark(T_CLOSEST));
        launch_trace(count, false, gt, st, P);
        CK(tl.mark(T_MIS));
        k_resolve_mis<<<g, kBlock, 0, st>>>(P);
        launches += 2;
      }
      CK(tl.mark(T_FILM));
      if (dual && waveIdx > 0) CK(cudaStreamWaitEvent(st, c->evFilm[set ^ 1], 0));  // the film takes the waves in order
      k_film_accumulate<<<grid_for(c, (uint64_t)rows * W, 8), kBlock, 0, st>>>(P);
      if (dual) CK(cudaEventRecord(c->evFilm[set], st));
      launches += 1;
      waveIdx++;
    }
  }
  if (dual) {  // join: the caller's stream continues when both wave streams are done
    CK(cudaEventRecord(c->evJoin, c->stream2));
    CK(cudaStreamWaitEvent(stMain, c->evJoin, 0));
  }
  CK(tl.mark(T_END));
  CK(cudaGetLastError());
  CK(cudaEventRecord(c->ev[1], st));
  if (l2Window) {  // the stream may be the caller's: leave it as it was
    cudaStreamAttrValue av{};
    av.accessPolicyWindow.num_bytes = 0;
    cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &av);
  }
  s->last = LastRender{(uint64_t)W * H * (uint64_t)nS, launches, tl};
  if (stats && fetch_stats(s, st, stats)) return 1;
  return 0;
}

int pbrgpu_resolve_film(pbrgpu_ctx* c, const float* d_rgba_sum, int width, int height, float inv_spp, float* out_rgb_mean,
                        unsigned char* out_rgba8) {
  if (!c || !d_rgba_sum || width <= 0 || height <= 0) return fail("pbrgpu_resolve_film: bad argument");
  if (!c->subs.empty()) c = c->subs[0];  // group context: the film lives on its first device
  CK(cudaSetDevice(c->device));
  size_t n = (size_t)width * height;
  DevBuf<float> drgb;
  DevBuf<unsigned char> d8;
  if (out_rgb_mean) CK(drgb.alloc(n * 3));
  if (out_rgba8) CK(d8.alloc(n * 4));
  k_resolve_film<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>((const float4*)d_rgba_sum, width, height, inv_spp,
                                                                      inv_spp <= 0 ? 1 : 0, drgb.p, d8.p);
  CK(cudaGetLastError());
  if (out_rgb_mean) CK(cudaMemcpyAsync(out_rgb_mean, drgb.p, n * 3 * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
  if (out_rgba8) CK(cudaMemcpyAsync(out_rgba8, d8.p, n * 4, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return 0;
}

__global__ void k_rgba_to_rgb(const float4* in, size_t n, float* out) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  float4 v = in[i];
  out[3 * i] = v.x; out[3 * i + 1] = v.y; out[3 * i + 2] = v.z;
}

// D2H through the scene's pinned staging buffer in chunks: the copy of chunk k+1 overlaps the host memcpy of chunk k.
static int staged_d2h(pbrgpu_scene* s, void* dst, const void* dsrc, size_t bytes, bool accumulateFloat) {
  pbrgpu_ctx* c = s->ctx;
  if (!accumulateFloat) {
    // the caller's buffer is itself page-locked (cudaHostAlloc / cudaHostRegister): one DMA, no staging, no host copy
    cudaPointerAttributes at{};
    if (cudaPointerGetAttributes(&at, dst) == cudaSuccess && at.type == cudaMemoryTypeHost) {
      CK(cudaMemcpyAsync(dst, dsrc, bytes, cudaMemcpyDeviceToHost, c->stream));
      CK(cudaStreamSynchronize(c->stream));
      return 0;
    }
    cudaGetLastError();
  }
  const size_t chunk = 4u << 20;
  if (s->hstageBytes < 2 * chunk) {
    if (s->hstage) cudaFreeHost(s->hstage);
    s->hstage = nullptr;
    CK(cudaHostAlloc(&s->hstage, 2 * chunk, cudaHostAllocDefault));
    s->hstageBytes = 2 * chunk;
  }
  char* stage[2] = {(char*)s->hstage, (char*)s->hstage + chunk};
  size_t nChunks = (bytes + chunk - 1) / chunk;
  auto len = [&](size_t k) { return std::min(chunk, bytes - k * chunk); };
  if (nChunks == 0) return 0;
  CK(cudaMemcpyAsync(stage[0], dsrc, len(0), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaEventRecord(c->ev[2], c->stream));
  for (size_t k = 0; k < nChunks; ++k) {
    if (k + 1 < nChunks) {
      CK(cudaMemcpyAsync(stage[(k + 1) & 1], (const char*)dsrc + (k + 1) * chunk, len(k + 1), cudaMemcpyDeviceToHost, c->stream));
      CK(cudaEventRecord(c->ev[3 + ((k + 1) & 1)], c->stream));
    }
    CK(cudaEventSynchronize(k == 0 ? c->ev[2] : c->ev[3 + (k & 1)]));
    if (accumulateFloat) {
      float* d = (float*)((char*)dst + k * chunk);
      const float* h = (const float*)stage[k & 1];
      for (size_t i = 0, n = len(k) / sizeof(float); i < n; ++i) d[i] += h[i];
    } else {
      memcpy((char*)dst + k * chunk, stage[k & 1], len(k));
    }
  }
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// Multi-GPU.  Every (pixel, sample index) is independent and addressable in O(1)
// (HaltonSampler::GetIndexForSample, Sampler/Halton.cpp:61-81), so a render shards by SAMPLE RANGE: the scene is
// replicated, device g of G renders [s0 + g*n/G, s0 + (g+1)*n/G) of every pixel into a private fp32 RGBA sum film
// (A = sample count), and ONE ncclReduce(sum, root 0) over NVLink adds the films; rank 0 resolves.  There is no other
// exchange.  Two ways to get there behind the same pbrgpu_render call:
//   pbrgpu_init_multi      one process, N devices (what a C++ application does: Main/main.cpp:413 stays one call)
//   pbrgpu_comm_attach     one process per device (torchrun): every rank calls pbrgpu_render with the same arguments
// ---------------------------------------------------------------------------------------------------------------
namespace {

// contiguous, balanced split of [s0, s1): the first (n % world) ranks take one extra sample (== shard.py)
void shard_samples(int64_t s0, int64_t s1, int rank, int world, int64_t* lo, int64_t* hi) {
  const int64_t n = s1 - s0, base = n / world, extra = n % world;
  *lo = s0 + rank * base + std::min<int64_t>(rank, extra);
  *hi = *lo + base + (rank < extra ? 1 : 0);
}

int nccl_fail(const NcclApi* api, const char* what, int rc) {
  return fail(std::string(what) + ": NCCL error " + std::to_string(rc) + " (" + (api->GetErrorString ? api->GetErrorString(rc) : "?") + ")");
}

void add_stats(pbrgpu_stats* a, const pbrgpu_stats& b) {
  a->camera_samples += b.camera_samples; a->rays_closest += b.rays_closest; a->rays_shadow += b.rays_shadow;
  a->nodes_visited += b.nodes_visited; a->tris_tested += b.tris_tested; a->kernel_launches += b.kernel_launches;
  a->rays_path += b.rays_path; a->n_trace_closest_launches += b.n_trace_closest_launches;
  a->ms_total = std::max(a->ms_total, b.ms_total);  // the devices run concurrently
  a->ms_raygen = std::max(a->ms_raygen, b.ms_raygen); a->ms_trace_closest = std::max(a->ms_trace_closest, b.ms_trace_closest);
  a->ms_shade = std::max(a->ms_shade, b.ms_shade); a->ms_trace_shadow = std::max(a->ms_trace_shadow, b.ms_trace_shadow);
  a->ms_trace_mis = std::max(a->ms_trace_mis, b.ms_trace_mis); a->ms_film = std::max(a->ms_film, b.ms_film);
}

// RGBA sum film on the scene's device -> the caller's host buffers (RGB sums and / or the tone-mapped 8-bit image)
int film_to_host(pbrgpu_scene* s, const float* d_rgba, int W, int H, float* out_rgb_sum, unsigned char* out_rgba8, bool accumulate) {
  pbrgpu_ctx* c = s->ctx;
  const size_t n = (size_t)W * H;
  if (out_rgb_sum) {
    if (s->drgb.n < n * 3) CK(s->drgb.alloc(n * 3));
    k_rgba_to_rgb<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>((const float4*)d_rgba, n, s->drgb.p);
    CK(cudaGetLastError());
    if (staged_d2h(s, out_rgb_sum, s->drgb.p, n * 3 * sizeof(float), accumulate)) return 1;
  }
  if (out_rgba8) {
    if (s->d8.n < n * 4) CK(s->d8.alloc(n * 4));
    k_resolve_film<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>((const float4*)d_rgba, W, H, 0.f, 1, nullptr, s->d8.p);
    CK(cudaGetLastError());
    if (staged_d2h(s, out_rgba8, s->d8.p, n * 4, false)) return 1;
  }
  return 0;
}

// GROUP scene: every replica renders its shard on its own device (one host thread each: the VolPath loop reads counters
// back, and nothing should serialise the devices), then one grouped ncclReduce into `d_root` on device 0.
// d_root: the root's film (device 0).  rootStream: the stream the root renders / reduces on.
int render_group_device(pbrgpu_scene* g, const pbrgpu_render_desc* rd, float* d_root, cudaStream_t rootStream, pbrgpu_stats* stats) {
  const int N = (int)g->replicas.size();
  const int W = rd->x1 - rd->x0, H = rd->y1 - rd->y0;
  const size_t n = (size_t)W * H;
  std::vector<int> rc((size_t)N, 0);
  std::vector<std::string> err((size_t)N);
  std::vector<pbrgpu_stats> st((size_t)N);
  std::vector<float*> film((size_t)N, nullptr);
  std::vector<std::thread> th;
  for (int i = 0; i < N; ++i) {
    th.emplace_back([&, i] {
      pbrgpu_scene* r = g->replicas[i];
      pbrgpu_render_desc r2 = *rd;
      shard_samples(rd->s0, rd->s1, i, N, &r2.s0, &r2.s1);
      float* dst = d_root;
      cudaStream_t stream = rootStream;
      if (i > 0) {  // private film, never accumulating
        r2.flags &= ~PBRGPU_RENDER_ACCUMULATE;
        if (cudaSetDevice(r->ctx->device) != cudaSuccess) { rc[i] = 1; err[i] = "cudaSetDevice failed"; return; }
        if (r->accumPix < n) {
          if (r->accum.alloc(n * 4) != cudaSuccess) { rc[i] = 1; err[i] = "out of device memory for the film"; return; }
          r->accumPix = n;
        }
        dst = r->accum.p;
        stream = r->ctx->stream;
      }
      film[i] = dst;
      rc[i] = render_device_single(r, &r2, dst, stream, stats ? &st[i] : nullptr);
      if (rc[i]) err[i] = g_err;
    });
  }
  for (auto& t : th) t.join();
  for (int i = 0; i < N; ++i) if (rc[i]) return fail("device " + std::to_string(g->replicas[i]->ctx->device) + ": " + err[i]);
  if (N > 1) {
    std::string e;
    const NcclApi* api = nccl_api(&e);
    if (!api) return fail(e);
    int r = api->GroupStart();
    if (r) return nccl_fail(api, "ncclGroupStart", r);
    for (int i = 0; i < N; ++i) {
      r = api->Reduce(film[i], film[0], n * 4, NcclApi::kFloat32, NcclApi::kSum, 0, g->ctx->comms[i], i == 0 ? rootStream : g->replicas[i]->ctx->stream);
      if (r) { api->GroupEnd(); return nccl_fail(api, "ncclReduce", r); }
    }
    r = api->GroupEnd();
    if (r) return nccl_fail(api, "ncclGroupEnd", r);
  }
  CK(cudaSetDevice(g->replicas[0]->ctx->device));
  if (stats) {
    *stats = st[0];
    for (int i = 1; i < N; ++i) add_stats(stats, st[i]);
  }
  return 0;
}

// ATTACHED context (one process per GPU): render this rank's shard, reduce to rank 0.
int render_rank_device(pbrgpu_scene* s, const pbrgpu_render_desc* rd, float* d_film, cudaStream_t st, pbrgpu_stats* stats) {
  pbrgpu_ctx* c = s->ctx;
  pbrgpu_render_desc r2 = *rd;
  shard_samples(rd->s0, rd->s1, c->rank, c->nranks, &r2.s0, &r2.s1);
  if (c->rank != 0) r2.flags &= ~PBRGPU_RENDER_ACCUMULATE;  // only the root's film carries earlier sums
  if (render_device_single(s, &r2, d_film, st, stats)) return 1;
  std::string e;
  const NcclApi* api = nccl_api(&e);
  if (!api) return fail(e);
  const size_t n = (size_t)(rd->x1 - rd->x0) * (rd->y1 - rd->y0);
  const int r = api->Reduce(d_film, d_film, n * 4, NcclApi::kFloat32, NcclApi::kSum, 0, c->comm, st);
  if (r) return nccl_fail(api, "ncclReduce", r);
  return 0;
}

inline bool is_group(const pbrgpu_scene* s) { return !s->replicas.empty(); }
inline bool sharded(const pbrgpu_scene* s, const pbrgpu_render_desc* rd) {
  if (rd->flags & PBRGPU_RENDER_LOCAL_ONLY) return false;
  return is_group(s) ? s->replicas.size() > 1 : (s->ctx->comm != nullptr && s->ctx->nranks > 1);
}

}  // namespace

int pbrgpu_render_device(pbrgpu_scene* s, const pbrgpu_render_desc* rd, float* d_rgba_sum, void* stream_, pbrgpu_stats* stats) {
  if (!s) return fail("pbrgpu_render_device: NULL scene");
  pbrgpu_scene* s0 = is_group(s) ? s->replicas[0] : s;
  if (check_render_args(s0, rd)) return 1;
  if (!d_rgba_sum) return fail("pbrgpu_render_device: NULL accumulator");
  if (!sharded(s, rd)) return render_device_single(s0, rd, d_rgba_sum, stream_, stats);
  cudaStream_t st = stream_ ? (cudaStream_t)stream_ : s0->ctx->stream;
  if (is_group(s)) return render_group_device(s, rd, d_rgba_sum, st, stats);
  return render_rank_device(s, rd, d_rgba_sum, st, stats);
}

int pbrgpu_render(pbrgpu_scene* s, const pbrgpu_render_desc* rd, float* out_rgb_sum, unsigned char* out_rgba8, pbrgpu_stats* stats) {
  if (!s) return fail("pbrgpu_render: NULL scene");
  pbrgpu_scene* s0 = is_group(s) ? s->replicas[0] : s;  // the scene on the root device
  if (check_render_args(s0, rd)) return 1;
  pbrgpu_ctx* c = s0->ctx;
  CK(cudaSetDevice(c->device));
  const int W = rd->x1 - rd->x0, H = rd->y1 - rd->y0;
  size_t n = (size_t)W * H;
  if (s0->accumPix < n) { CK(s0->accum.alloc(n * 4)); s0->accumPix = n; }
  pbrgpu_render_desc r2 = *rd;
  r2.flags &= ~PBRGPU_RENDER_ACCUMULATE;  // PBRGPU_RENDER_ACCUMULATE on host buffers: the caller's sums are added after the render
  bool root = true;
  if (!sharded(s, rd)) {
    if (render_device_single(s0, &r2, s0->accum.p, c->stream, nullptr)) return 1;
  } else if (is_group(s)) {
    if (render_group_device(s, &r2, s0->accum.p, c->stream, stats)) return 1;
  } else {
    if (render_rank_device(s, &r2, s0->accum.p, c->stream, nullptr)) return 1;
    root = c->rank == 0;
  }
  // the reduced film leaves the GPU on the root only; the other ranks of a process-per-GPU job return once their
  // shard has been handed to the reduce (their output buffers are not touched and may be NULL)
  if (root && film_to_host(s0, s0->accum.p, W, H, out_rgb_sum, out_rgba8, (rd->flags & PBRGPU_RENDER_ACCUMULATE) != 0)) return 1;
  if (!root) CK(cudaStreamSynchronize(c->stream));
  if (stats && !(is_group(s) && sharded(s, rd)) && fetch_stats(s0, c->stream, stats)) return 1;
  return 0;
}

int pbrgpu_scene_create(pbrgpu_ctx* ctx, const pbrgpu_scene_desc* d, pbrgpu_scene** out) {
  if (!ctx || !d || !out) return fail("pbrgpu_scene_create: NULL argument");
  *out = nullptr;
  if (ctx->subs.empty()) return scene_create_single(ctx, d, nullptr, out);
  // GROUP context: validate + build on the first device, then replicate with the same BVH on the others (in parallel)
  std::unique_ptr<pbrgpu_scene> g(new pbrgpu_scene);
  g->ctx = ctx;
  const int N = (int)ctx->subs.size();
  g->replicas.assign((size_t)N, nullptr);
  auto destroy_all = [&] { for (auto*& r : g->replicas) { if (r) pbrgpu_scene_destroy(r); r = nullptr; } };
  if (scene_create_single(ctx->subs[0], d, nullptr, &g->replicas[0])) { destroy_all(); g->replicas.clear(); return 1; }
  std::vector<int> rc((size_t)N, 0);
  std::vector<std::string> err((size_t)N);
  std::vector<std::thread> th;
  for (int i = 1; i < N; ++i)
    th.emplace_back([&, i] {
      rc[i] = scene_create_single(ctx->subs[i], d, &g->replicas[0]->bvh, &g->replicas[i]);
      if (rc[i]) err[i] = g_err;
    });
  for (auto& t : th) t.join();
  for (int i = 1; i < N; ++i)
    if (rc[i]) { destroy_all(); g->replicas.clear(); return fail("device " + std::to_string(ctx->subs[i]->device) + ": " + err[i]); }
  CK(cudaSetDevice(ctx->subs[0]->device));
  *out = g.release();
  return 0;
}

int pbrgpu_camera_set(pbrgpu_scene* s, const pbrgpu_camera_desc* c) {
  if (!s || !c) return fail("pbrgpu_camera_set: NULL argument");
  if (!is_group(s)) return camera_set_single(s, c);
  for (pbrgpu_scene* r : s->replicas) if (camera_set_single(r, c)) return 1;
  return 0;
}

int pbrgpu_sampler_set(pbrgpu_scene* s, const pbrgpu_sampler_desc* d) {
  if (!s || !d) return fail("pbrgpu_sampler_set: NULL argument");
  if (!is_group(s)) return sampler_set_single(s, d);
  for (pbrgpu_scene* r : s->replicas) if (sampler_set_single(r, d)) return 1;
  return cudaSetDevice(s->replicas[0]->ctx->device) == cudaSuccess ? 0 : fail("cudaSetDevice failed");
}

int pbrgpu_init_multi(int n_devices, const int* device_ids, pbrgpu_ctx** out) {
  if (!out) return fail("pbrgpu_init_multi: out == NULL");
  *out = nullptr;
  const int avail = pbrgpu_device_count();
  if (avail == 0) return fail("pbrgpu_init_multi: no CUDA device available (this library has no CPU fallback)");
  if (n_devices <= 0) n_devices = avail;  // "all devices"
  if (n_devices > avail && !device_ids) return fail("pbrgpu_init_multi: more devices requested than present");
  std::vector<int> ids((size_t)n_devices);
  for (int i = 0; i < n_devices; ++i) {
    ids[i] = device_ids ? device_ids[i] : i;
    if (ids[i] < 0 || ids[i] >= avail) return fail("pbrgpu_init_multi: bad device id");
    for (int j = 0; j < i; ++j) if (ids[j] == ids[i]) return fail("pbrgpu_init_multi: duplicate device id");
  }
  std::unique_ptr<pbrgpu_ctx> g(new pbrgpu_ctx);
  auto cleanup = [&] { for (pbrgpu_ctx* c : g->subs) pbrgpu_shutdown(c); g->subs.clear(); };
  for (int i = 0; i < n_devices; ++i) {
    pbrgpu_ctx* c = nullptr;
    if (pbrgpu_init(ids[i], &c)) { cleanup(); return 1; }
    g->subs.push_back(c);
  }
  g->device = ids[0];
  g->smCount = g->subs[0]->smCount;
  if (n_devices > 1) {
    std::string e;
    const NcclApi* api = nccl_api(&e);
    if (!api) { cleanup(); return fail("pbrgpu_init_multi: " + e); }
    g->comms.assign((size_t)n_devices, nullptr);
    const int r = api->CommInitAll(g->comms.data(), n_devices, ids.data());
    if (r) { g->comms.clear(); cleanup(); return nccl_fail(api, "ncclCommInitAll", r); }
  }
  cudaSetDevice(ids[0]);
  *out = g.release();
  return 0;
}

int pbrgpu_comm_unique_id(void* out128) {
  if (!out128) return fail("pbrgpu_comm_unique_id: NULL argument");
  std::string e;
  const NcclApi* api = nccl_api(&e);
  if (!api) return fail("pbrgpu_comm_unique_id: " + e);
  NcclApi::UniqueId id;
  const int r = api->GetUniqueId(&id);
  if (r) return nccl_fail(api, "ncclGetUniqueId", r);
  memcpy(out128, &id, sizeof(id));
  return 0;
}

int pbrgpu_comm_attach(pbrgpu_ctx* c, int nranks, int rank, const void* unique_id128) {
  if (!c || !unique_id128) return fail("pbrgpu_comm_attach: NULL argument");
  if (!c->subs.empty()) return fail("pbrgpu_comm_attach: a pbrgpu_init_multi context already owns its devices");
  if (nranks < 1 || rank < 0 || rank >= nranks) return fail("pbrgpu_comm_attach: bad rank / nranks");
  if (c->comm) return fail("pbrgpu_comm_attach: already attached");
  std::string e;
  const NcclApi* api = nccl_api(&e);
  if (!api) return fail("pbrgpu_comm_attach: " + e);
  CK(cudaSetDevice(c->device));
  NcclApi::UniqueId id;
  memcpy(&id, unique_id128, sizeof(id));
  const int r = api->CommInitRank(&c->comm, nranks, id, rank);
  if (r) { c->comm = nullptr; return nccl_fail(api, "ncclCommInitRank", r); }
  c->nranks = nranks;
  c->rank = rank;
  return 0;
}

int pbrgpu_resample_pow2(const float* rgb, int32_t width, int32_t height, int32_t wrap, float* out, int64_t out_floats, int32_t out_size[2]) {
  if (!rgb || !out_size || width <= 0 || height <= 0 || wrap < PBRGPU_WRAP_REPEAT || wrap > PBRGPU_WRAP_CLAMP)
    return fail("pbrgpu_resample_pow2: bad argument");
  int W = width, H = height;
  std::vector<float> r;
  if (host::is_pow2(W) && host::is_pow2(H)) r.assign(rgb, rgb + (size_t)W * H * 3);
  else r = host::resample_to_pow2(rgb, width, height, wrap, &W, &H);
  out_size[0] = W; out_size[1] = H;
  if (out) {
    if (out_floats < (int64_t)r.size()) return fail("pbrgpu_resample_pow2: output too small");
    memcpy(out, r.data(), r.size() * sizeof(float));
  }
  return 0;
}

int pbrgpu_shard_samples(int64_t s0, int64_t s1, int rank, int world, int64_t out2[2]) {
  if (!out2 || world < 1 || rank < 0 || rank >= world || s1 < s0) return fail("pbrgpu_shard_samples: bad argument");
  shard_samples(s0, s1, rank, world, &out2[0], &out2[1]);
  return 0;
}

int pbrgpu_measure_read_bandwidth(pbrgpu_ctx* c, size_t bytes, int iters, double* out_gb_per_s) {
  if (!c || !out_gb_per_s || bytes < (1u << 20) || iters < 1) return fail("pbrgpu_measure_read_bandwidth: bad argument");
  if (!c->subs.empty()) c = c->subs[0];
  CK(cudaSetDevice(c->device));
  DevBuf<float4> buf;
  DevBuf<float> sink;
  const size_t n = bytes / sizeof(float4);
  CK(buf.alloc(n));
  CK(sink.alloc(1));
  CK(cudaMemsetAsync(buf.p, 0, n * sizeof(float4), c->stream));
  const unsigned grid = (unsigned)c->smCount * 8;
  double best = 0;
  for (int rep = 0; rep < 4; ++rep) {  // rep 0 warms the L2 (and the clocks)
    CK(cudaEventRecord(c->ev[5], c->stream));
    k_read_bandwidth<<<grid, 256, 0, c->stream>>>(buf.p, n, iters, sink.p);
    CK(cudaGetLastError());
    CK(cudaEventRecord(c->ev[6], c->stream));
    CK(cudaEventSynchronize(c->ev[6]));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, c->ev[5], c->ev[6]));
    if (rep > 0 && ms > 0) best = std::max(best, (double)n * sizeof(float4) * iters / (ms * 1e-3) / 1e9);
  }
  *out_gb_per_s = best;
  return 0;
}

int pbrgpu_ctx_info(pbrgpu_ctx* c, int32_t out[4]) {
  if (!c || !out) return fail("pbrgpu_ctx_info: NULL argument");
  out[0] = c->subs.empty() ? 1 : (int32_t)c->subs.size();
  out[1] = c->rank;
  out[2] = c->nranks;
  out[3] = 0;
  if (!c->subs.empty() || c->comm) {
    const NcclApi* api = nccl_api(nullptr);
    int v = 0;
    if (api && api->GetVersion && api->GetVersion(&v) == 0) out[3] = v;
  }
  return 0;
}

int pbrgpu_debug_dump(pbrgpu_scene* s, const pbrgpu_dump_desc* dd, void* out, size_t out_bytes) {
  if (s && !s->replicas.empty()) s = s->replicas[0];  // group scene: the hooks run on the first device
  if (!s || !dd || !out) return fail("pbrgpu_debug_dump: NULL argument");
  pbrgpu_ctx* c = s->ctx;
  CK(cudaSetDevice(c->device));
  if (dd->what == PBRGPU_DUMP_PERMS) {
    if (!s->haveSampler) return fail("pbrgpu_debug_dump: sampler not set");
    size_t bytes = std::min(out_bytes, (size_t)s->sc.hp.nPerms * sizeof(uint16_t));
    CK(cudaMemcpy(out, s->perms.p, bytes, cudaMemcpyDeviceToHost));
    return 0;
  }
  if (!s->haveSampler) return fail("pbrgpu_debug_dump: sampler not set");
  const int W = dd->x1 - dd->x0, H = dd->y1 - dd->y0;
  const int64_t ns = dd->s1 - dd->s0;
  if (W <= 0 || H <= 0 || ns <= 0 || ns > (1 << 20)) return fail("pbrgpu_debug_dump: bad region");
  if (dd->what == PBRGPU_DUMP_SAMPLER) {
    if (dd->n_dims <= 0) return fail("pbrgpu_debug_dump: n_dims <= 0");
    size_t n = (size_t)W * H * ns * dd->n_dims;
    if (out_bytes < n * sizeof(float)) return fail("pbrgpu_debug_dump: output too small");
    DevBuf<float> d;
    CK(d.alloc(n));
    k_dump_sampler<<<grid_for(c, n, 8), kBlock, 0, c->stream>>>(s->sc.hp, dd->x0, dd->y0, W, H, dd->s0, (int)ns, dd->n_dims, d.p);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out, d.p, n * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return 0;
  }
  if (dd->what == PBRGPU_DUMP_PRIMARY_RAYS || dd->what == PBRGPU_DUMP_PRIMARY_HITS || dd->what == PBRGPU_DUMP_HIT_DIFFERENTIALS) {
    if (!s->haveCamera) return fail("pbrgpu_debug_dump: camera not set");
    bool rays = dd->what == PBRGPU_DUMP_PRIMARY_RAYS;
    const bool diff = dd->what == PBRGPU_DUMP_HIT_DIFFERENTIALS;
    size_t n = (size_t)W * H * ns;
    size_t bytes = n * ((rays || diff) ? 8 * sizeof(float) : 4 * sizeof(int));
    if (out_bytes < bytes) return fail("pbrgpu_debug_dump: output too small");
    DevBuf<unsigned char> d;
    CK(d.alloc(bytes));
    RenderParams P{};
    P.sc = s->sc;
    P.x0 = dd->x0; P.y0 = dd->y0; P.w = W; P.h = H; P.sBegin = dd->s0;
    P.diffScale = 1 / std::sqrt((float)s->spp);
    k_dump_primary<<<grid_for(c, n, 8), kBlock, 0, c->stream>>>(P, (int)ns, rays ? (float*)d.p : nullptr,
                                                                 (rays || diff) ? nullptr : (int*)d.p, diff ? (float*)d.p : nullptr);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out, d.p, bytes, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return 0;
  }
  if (dd->what == PBRGPU_DUMP_WAVE_HITS) {
    // the PRODUCTION kernels: k_wave_begin -> k_raygen -> k_trace<false, false>, then ws.hit in caller order
    if (!s->haveCamera) return fail("pbrgpu_debug_dump: camera not set");
    const uint64_t n = (uint64_t)W * H * ns;
    if (n > (1u << 25)) return fail("pbrgpu_debug_dump: region x samples exceeds one wave (2^25 paths)");
    if (out_bytes < n * 4 * sizeof(int)) return fail("pbrgpu_debug_dump: output too small");
    if (s->bvh.maxDepth >= kSmemStack + kLocalStack) return fail("pbrgpu_debug_dump: BVH deeper than the traversal stack");
    if (ensure_wave(s, (uint32_t)n, 1, false)) return 1;
    if (s->traceBlocksPerSm <= 0) s->traceBlocksPerSm = std::max(1, trace_blocks_per_sm());
    DevBuf<int> d;
    CK(d.alloc(n * 4));
    RenderParams P{};
    P.sc = s->sc;
    fill_wave(s, &P.ws);
    P.integrator = PBRGPU_INTEGRATOR_PATH;
    P.x0 = dd->x0; P.y0 = dd->y0; P.w = W; P.h = H; P.sBegin = dd->s0; P.nsWave = (int)ns;
    P.arrayEndDim = 5;
    P.fetchMinIdle = s->fetchMinIdle;
    P.storeHitT = 1;
    cudaStream_t st = c->stream;
    CK(cudaMemsetAsync(s->ctr.p, 0, sizeof(Counters), st));
    k_wave_begin<<<1, 1, 0, st>>>(s->ctr.p);
    k_raygen<<<grid_for(c, n, 8), kBlock, 0, st>>>(P);
    const uint32_t gt = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>((n + kTraceBlockThreads - 1) / kTraceBlockThreads,
                                                                            (uint64_t)c->smCount * s->traceBlocksPerSm));
    launch_trace(false, false, gt, st, P);
    k_collect_wave_hits<<<grid_for(c, n, 8), kBlock, 0, st>>>(P, (int)ns, d.p);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out, d.p, n * 4 * sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return 0;
  }
  return fail("pbrgpu_debug_dump: unknown dump kind");
}

int pbrgpu_trace_rays(pbrgpu_scene* s, const float* rays, int64_t n, int any_hit, void* out_hits) {
  if (s && !s->replicas.empty()) s = s->replicas[0];  // group scene: the hooks run on the first device
  if (!s || !rays || !out_hits || n <= 0) return fail("pbrgpu_trace_rays: bad argument");
  pbrgpu_ctx* c = s->ctx;
  CK(cudaSetDevice(c->device));
  DevBuf<float> dr;
  DevBuf<int> dh;
  CK(dr.alloc((size_t)n * 8));
  CK(dh.alloc((size_t)n * ((any_hit & 1) ? 1 : 4)));
  CK(cudaMemcpyAsync(dr.p, rays, (size_t)n * 8 * sizeof(float), cudaMemcpyHostToDevice, c->stream));
  RenderParams P{};
  P.sc = s->sc;
  if (any_hit & PBRGPU_TRACE_VIA_WAVE_KERNEL) {
    // through the production kernel k_trace and its three queues (path + MIS for closest hits, shadow for any-hit)
    const bool anyHit = (any_hit & 1) != 0;
    if (n > (1 << 25)) return fail("pbrgpu_trace_rays: more rays than one wave holds (2^25)");
    if (s->bvh.maxDepth >= kSmemStack + kLocalStack) return fail("pbrgpu_trace_rays: BVH deeper than the traversal stack");
    if (ensure_wave(s, (uint32_t)n, 1, false)) return 1;
    if (s->traceBlocksPerSm <= 0) s->traceBlocksPerSm = std::max(1, trace_blocks_per_sm());
    fill_wave(s, &P.ws);
    P.fetchMinIdle = s->fetchMinIdle;
    P.storeHitT = 1;
    DevBuf<int> dis;
    CK(dis.alloc(1));
    cudaStream_t st = c->stream;
    CK(cudaMemsetAsync(dis.p, 0, sizeof(int), st));
    CK(cudaMemsetAsync(s->ctr.p, 0, sizeof(Counters), st));
    k_wave_begin<<<1, 1, 0, st>>>(s->ctr.p);
    k_feed_rays<<<grid_for(c, (uint64_t)n, 8), kBlock, 0, st>>>(P, dr.p, (uint32_t)n, anyHit ? 1 : 0);
    const uint32_t gt = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>((2 * (uint64_t)n + kTraceBlockThreads - 1) / kTraceBlockThreads,
                                                                            (uint64_t)c->smCount * s->traceBlocksPerSm));
    launch_trace(false, false, gt, st, P);
    k_collect_rays<<<grid_for(c, (uint64_t)n, 8), kBlock, 0, st>>>(P, (uint32_t)n, anyHit ? 1 : 0, dh.p, dis.p);
    CK(cudaGetLastError());
    int nDis = 0;
    CK(cudaMemcpyAsync(out_hits, dh.p, dh.n * sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(&nDis, dis.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (nDis) return fail("pbrgpu_trace_rays: k_trace's path queue and MIS queue found different triangles for " + std::to_string(nDis) + " rays");
    return 0;
  }
  k_trace_user<<<grid_for(c, (uint64_t)n, 8), kBlock, 0, c->stream>>>(P, dr.p, n, any_hit & 1, dh.p);
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(out_hits, dh.p, dh.n * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return 0;
}

}  // extern "C"
