/*
 * pbrgpu.h - C-ABI of the B200-native wavefront path tracer (libpbrgpu.so).
 *
 * This is the drop-in boundary for ONE hot path of G0T-cha/PysicalBasedRaytracer: everything that
 * happens inside  Integrator::Render(const Scene&, double&)  (Integrator/Integrator.h:14) for the
 * Path / DirectLighting / Whitted integrators:
 *     Sampler::GetCameraSample            Sampler/Sampler.cpp:10-21
 *     HaltonSampler::SampleDimension      Sampler/Halton.cpp:61-92, Sampler/LowDiscrepancy.cpp:211-225,2300-2316
 *     PerspectiveCamera::GenerateRay...   Camera/Perspective.cpp:21-80, Camera/Orthographic.cpp:7-63
 *     BVHAccel::Intersect / IntersectP    Accelerator/BVHAccel.cpp:285-366, Shape/Triangle.cpp:62-350
 *     PathIntegrator::Li                  Integrator/PathIntegrator.cpp:32-110
 *     VolPathIntegrator::Li               Integrator/VolPathIntegrator.cpp:21-107  (+ Media/HomogeneousMedium.cpp:10-45,
 *                                         Media/Medium.cpp:7-28, VisibilityTester::Tr Light/Light.cpp:31-47)
 *     UniformSampleOneLight/EstimateDirect Integrator/Integrator.cpp:46-177
 *     Material::ComputeScatteringFunctions + BSDF::{f,Pdf,Sample_f}   Material/ (all .cpp)
 *     Light::{Sample_Li,Pdf_Li,Le,L}      Light/ (all .cpp)
 *     per-pixel mean + XYZ/sRGB/gamma/8-bit  Integrator/Integrator.cpp:313-344
 *
 * The reference has no FFI layer: its boundary is the C++ virtual above, whose inputs are object
 * graphs (Scene -> BVHAccel -> GeometricPrimitive -> Triangle/Material/AreaLight, Camera, Sampler).
 * A host shim (pysicalbasedraytracer_b200/host/GpuIntegrator.h for C++, pysicalbasedraytracer_b200/api.py for
 * Python) flattens those graphs into the POD descriptors below and calls these entry points.
 *
 * Conventions: plain C, POD structs, opaque handles, int status (0 = OK, non-zero = error; text via
 * pbrgpu_last_error()), no exceptions cross the boundary.  All host arrays are copied during the
 * call; the library owns device memory behind the handles.  Handles are not thread-safe.
 * There is NO CPU fallback: without a CUDA device every compute entry point fails with an error.
 */
#ifndef PBRGPU_H
#define PBRGPU_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pbrgpu_ctx pbrgpu_ctx;     /* one CUDA device + stream, or a group of them (pbrgpu_init_multi) */
typedef struct pbrgpu_scene pbrgpu_scene; /* flattened Scene + Camera + Sampler resident in HBM */

/* ---- materials: Material/{MatteMaterial,Mirror,GlassMaterial,MetalMaterial,PlasticMaterial}.cpp ---- */
enum {
  PBRGPU_MAT_MATTE = 0,   /* Kd=kd, sigma               MatteMaterial.cpp:13-28   */
  PBRGPU_MAT_MIRROR = 1,  /* Kr=kr                      Mirror.cpp:5-15           */
  PBRGPU_MAT_GLASS = 2,   /* Kr=kr Kt=kt index=eta[0]   GlassMaterial.cpp:9-57    */
  PBRGPU_MAT_METAL = 3,   /* eta, k, u/v roughness      MetalMaterial.cpp:25-43   */
  PBRGPU_MAT_PLASTIC = 4, /* Kd=kd Ks=kr roughness=u_rough  PlasticMaterial.cpp:8-30 */
  PBRGPU_MAT_COUNT = 5
};
typedef struct pbrgpu_material {
  int32_t type;
  int32_t remap_roughness;
  float sigma;   /* Matte (degrees) */
  float u_rough; /* Glass/Metal u roughness, Plastic roughness */
  float kd[3];
  float v_rough;
  float kr[3];
  int32_t ks_texture_plus1; /* Plastic: Ks comes from textures[ks_texture_plus1 - 1] instead of kr (Main/main.cpp:63-77 builds its
                               Plastic with ONE ImageTexture as Kd and Ks); 0 = ConstantTexture(kr).  Was padding: zero-filled records keep their meaning */
  float kt[3];
  int32_t kd_texture_plus1; /* Matte / Plastic: Kd comes from textures[kd_texture_plus1 - 1] (an ImageTexture) instead of kd;
                               0 = ConstantTexture(kd) */
  float eta[3]; /* Metal eta (RGB); Glass index of refraction in eta[0] */
  int32_t rough_texture_plus1; /* Plastic with remap_roughness == 0: roughness comes from the first channel of
                                  textures[rough_texture_plus1 - 1] - an ImageTexture<float, float> (Shape/ModelLoad.cpp:147-164 builds its
                                  Plastic that way); the caller passes the float texels of its MIPMap replicated in R, G, B.
                                  0 = ConstantTexture(u_rough).  Was padding */
  float k[3]; /* Metal absorption */
  float pad3;
} pbrgpu_material; /* 96 bytes */

/* ---- image textures: ImageTexture<RGBSpectrum, Spectrum> (Texture/ImageTexture.h) = UVMapping2D + MIPMap (Texture/MIPMap.h) ---- */
enum { PBRGPU_WRAP_REPEAT = 0, PBRGPU_WRAP_BLACK = 1, PBRGPU_WRAP_CLAMP = 2 }; /* ImageWrap */
typedef struct pbrgpu_texture {
  int32_t width, height;  /* any size: sides that are not powers of two are Lanczos-resampled up like MIPMap does (MIPMap.h:93-144) */
  const float* texels;    /* width*height*3 RGB, row-major, as the MIPMap receives them, i.e. after ImageTexture::convertIn
                             (scale, optional inverse gamma) and after the loader's vertical flip (ImageTexture.cpp:16) */
  int32_t do_trilinear;   /* MIPMap::doTrilinear: trilinear filter instead of EWA */
  float max_anisotropy;   /* MIPMap::maxAnisotropy (EWA) */
  int32_t wrap;           /* PBRGPU_WRAP_* */
  float su, sv, du, dv;   /* UVMapping2D (Texture/Texture.h) */
  int32_t pad0;
} pbrgpu_texture; /* 48 bytes */

/* ---- participating media: Media/HomogeneousMedium.{h,cpp}, phase function Media/Medium.{h,cpp} ---- */
enum { PBRGPU_MEDIUM_HOMOGENEOUS = 0 };
typedef struct pbrgpu_medium {
  int32_t type;
  float g;          /* Henyey-Greenstein asymmetry */
  float sigma_a[3]; /* absorption */
  float sigma_s[3]; /* scattering; sigma_t = sigma_s + sigma_a (HomogeneousMedium.h:13-16) */
} pbrgpu_medium; /* 32 bytes */

/* ---- lights: Light/{PointLight,DiffuseLight,InfiniteAreaLight,SkyBoxLight}.cpp ---- */
enum {
  PBRGPU_LIGHT_POINT = 0,    /* pos, L = intensity I */
  PBRGPU_LIGHT_AREA_TRI = 1, /* DiffuseAreaLight over ONE triangle (the reference creates one light per Triangle) */
  PBRGPU_LIGHT_INFINITE = 2, /* InfiniteAreaLight: env map * L, importance sampled, light_to_world */
  PBRGPU_LIGHT_SKYBOX = 3,   /* SkyBoxLight: env map, uniform sphere sampling, tone-mapped texels */
  PBRGPU_LIGHT_AREA_SPHERE = 4 /* DiffuseAreaLight over ONE analytic sphere (pbrgpu_sphere): .tri = index into spheres; sampled by
                                  solid angle (uniform cone outside, uniform area inside: host/AnalyticSphere.h) */
};
typedef struct pbrgpu_light {
  int32_t type;
  int32_t tri;       /* AREA_TRI: triangle index in the caller's order */
  int32_t two_sided; /* AREA_TRI */
  int32_t n_samples; /* Light::nSamples (used by UniformSampleAllLights) */
  float pos[3];      /* POINT: world position (LightToWorld(0,0,0)) */
  float world_radius; /* SKYBOX: ctor argument.  INFINITE: ignored (computed from the scene bound) */
  float L[3];        /* POINT: I.  AREA_TRI: Lemit.  INFINITE: scale applied to texels */
  float pad0;
  float light_to_world[16]; /* INFINITE, row-major Matrix4x4::m */
  float world_to_light[16]; /* INFINITE, row-major inverse */
} pbrgpu_light; /* 176 bytes */

/* ---- analytic spheres.  The reference's PBR::Sphere is a stub (Shape/Sphere.cpp:8-37: no hit record, IntersectP
 *      false); Shape/README.md section 3 documents the intended quadratic.  CPU twin: host/AnalyticSphere.h (a subclass of the
 *      reference's Shape) - PARITY UNPINNED by the reference, pins are twin == GPU.  Optional: a scene without spheres never
 *      runs this code (separate kernel instantiation).  Sphere k is primitive n_tris + k in the parity hooks. ---- */
typedef struct pbrgpu_sphere {
  float center[3];   /* ObjectToWorld must be a pure translation */
  float radius;
  int32_t material;  /* index into materials */
  int32_t light;     /* index into lights (PBRGPU_LIGHT_AREA_SPHERE) or -1 */
  int32_t reverse_orientation;
  int32_t pad0;
} pbrgpu_sphere; /* 32 bytes */

/* ---- scene: Core/Scene.h:13-30 + Accelerator/BVHAccel.h + Shape/Triangle.h ---- */
typedef struct pbrgpu_scene_desc {
  int32_t n_verts;
  int32_t n_tris;
  const float* verts;          /* n_verts*3, WORLD space (TriangleMesh ctor transforms, Triangle.cpp:27-28) */
  const int32_t* indices;      /* n_tris*3 */
  const float* normals;        /* optional n_verts*3 world-space shading normals (mesh->n) or NULL */
  const float* uvs;            /* optional n_verts*2 (mesh->uv) or NULL -> default (0,0),(1,0),(1,1) Triangle.h:50-54 */
  const int32_t* tri_material; /* n_tris: index into materials; -1 = no material (GeometricPrimitive with a NULL
                                  Material: a pure medium boundary, Core/Primitive.cpp:39-46) - such scenes render
                                  with PBRGPU_INTEGRATOR_VOLPATH only */
  const int32_t* tri_light;    /* n_tris: index into lights for emissive triangles, else -1 */
  const uint8_t* tri_flags;    /* optional n_tris: bit0 = reverseOrientation ^ transformSwapsHandedness, bit1 = reverseOrientation,
                                  bit2 / bit3 / bit4 = the triangle's mesh has per-vertex normals / uvs / tangents (when NULL: no
                                  flips, and every triangle uses whichever of `normals` / `uvs` / `tangents` is given) */
  int32_t n_materials;
  int32_t n_lights;
  const pbrgpu_material* materials;
  const pbrgpu_light* lights;
  int32_t env_width;  /* env map for INFINITE / SKYBOX lights; INFINITE: sides that are not powers of two are Lanczos-resampled
                         up like InfiniteAreaLight's MIPMap does; SKYBOX reads the raw image at any size */
  int32_t env_height;
  const float* env_texels; /* env_width*env_height*3 RGB, row-major, raw (unscaled) */
  /* participating media (GeometricPrimitive::mediumInterface, Core/Primitive.h:43): all optional */
  int32_t n_media;
  int32_t pad0;
  const pbrgpu_medium* media;
  const int32_t* tri_medium_inside;  /* n_tris: index into media or -1 (MediumInterface::inside);  NULL = none */
  const int32_t* tri_medium_outside; /* n_tris: index into media or -1 (MediumInterface::outside); NULL = none */
  /* image textures referenced by pbrgpu_material::kd_texture_plus1 / ks_texture_plus1 */
  int32_t n_textures;
  int32_t pad1;
  const pbrgpu_texture* textures;
  /* analytic spheres (optional) */
  int32_t n_spheres;
  int32_t pad2;
  const pbrgpu_sphere* spheres;
  const float* tangents;       /* optional n_verts*3 world-space per-vertex tangents (mesh->s, Triangle.cpp:204-213) or NULL;
                                  tri_flags bit4 = this triangle's mesh has them (NULL tri_flags: every triangle uses them) */
} pbrgpu_scene_desc;

/* ---- camera: Camera/Camera.h:34-54, Perspective.cpp, Orthographic.cpp ---- */
enum { PBRGPU_CAMERA_PERSPECTIVE = 0, PBRGPU_CAMERA_ORTHOGRAPHIC = 1 };
typedef struct pbrgpu_camera_desc {
  int32_t type;
  int32_t width;
  int32_t height;
  float lens_radius;
  float focal_distance;
  float raster_to_camera[16]; /* ProjectiveCamera::RasterToCamera.m, row-major */
  float camera_to_world[16];  /* Camera::CameraToWorld.m, row-major */
  int32_t medium_plus1;       /* Camera::medium as (index into media) + 1; 0 = vacuum.  Read by VOLPATH only.  NOTE: the
                                 reference itself never delivers Camera::medium to the integrator (Perspective.cpp:44-80 does
                                 not set ray->medium; Orthographic.cpp:59-60 sets it before CameraToWorld(*ray), which rebuilds
                                 the ray without it, Transform.h:169), so the host mirrors pass 0 to stay identical */
} pbrgpu_camera_desc;

/* ---- sampler: Sampler/Halton.cpp:30-58; Sobol: host/SobolSampler.h over the reference's tables
 *      (Sampler/SobolMatrices.h:42-47) - the reference itself has no Sobol sampler, parity unpinned ---- */
enum { PBRGPU_SAMPLER_HALTON = 0, PBRGPU_SAMPLER_SOBOL = 1 };
typedef struct pbrgpu_sampler_desc {
  int32_t type;
  int32_t sample_at_pixel_center;
  int64_t samples_per_pixel;
  int32_t sample_bounds[4]; /* x0,y0,x1,y1 passed to HaltonSampler (drives baseScales/Exponents) */
  const uint16_t* perms;    /* optional: HaltonSampler::radicalInversePermutations; NULL -> library
                               computes them with its own PCG32 (RNG.h:93-108, LowDiscrepancy.cpp:2284-2298) */
  int64_t n_perms;
  /* SOBOL only: the reference's own generator tables, supplied by the host (they are data of the reference,
   * not of this library).  samples_per_pixel is rounded up to a power of two like a pbrt GlobalSampler. */
  const uint32_t* sobol_matrices32; /* SobolMatrices32[sobol_n_dims * 52] */
  int32_t sobol_n_dims;             /* <= 1024 */
  int32_t pad0;
  const uint64_t* sobol_vdc;        /* VdCSobolMatrices[log2Resolution - 1][52]    */
  const uint64_t* sobol_vdc_inv;    /* VdCSobolMatricesInv[log2Resolution - 1][52] */
} pbrgpu_sampler_desc;

/* ---- render: Integrator ctor arguments (PathIntegrator.h:14-18, WhittedIntegrator.h:8-11,
 *      DirectLightingIntegrator.h:10-17) + the pixel / sample sub-range to compute ---- */
enum {
  PBRGPU_INTEGRATOR_PATH = 0,
  PBRGPU_INTEGRATOR_DIRECT_ONE = 1, /* LightStrategy::UniformSampleOne */
  PBRGPU_INTEGRATOR_DIRECT_ALL = 2, /* LightStrategy::UniformSampleAll */
  PBRGPU_INTEGRATOR_WHITTED = 3,
  PBRGPU_INTEGRATOR_VOLPATH = 4     /* VolPathIntegrator (Integrator/VolPathIntegrator.cpp): media + null-material boundaries */
};
enum { PBRGPU_LIGHTSTRATEGY_UNIFORM = 0, PBRGPU_LIGHTSTRATEGY_POWER = 1 };
typedef struct pbrgpu_render_desc {
  int32_t integrator;
  int32_t max_depth;
  float rr_threshold;
  int32_t light_strategy;
  int32_t x0, y0, x1, y1; /* pixel region [x0,x1) x [y0,y1) */
  int64_t s0, s1;         /* sample indices [s0,s1) of each pixel (multi-GPU shards this range) */
  int32_t samples_per_wave; /* 0 = auto */
  int32_t flags;            /* PBRGPU_RENDER_* */
} pbrgpu_render_desc;
enum {
  PBRGPU_RENDER_ACCUMULATE = 1, /* add into the output sums instead of overwriting */
  PBRGPU_RENDER_COUNTERS = 2,   /* count nodes visited / triangles tested (slower) */
  PBRGPU_RENDER_TIMERS = 4,     /* CUDA events around every kernel group -> pbrgpu_stats.ms_* */
  PBRGPU_RENDER_LOCAL_ONLY = 8  /* multi-GPU context: do NOT shard / reduce - this device renders [s0,s1) alone (parity checks) */
};

typedef struct pbrgpu_stats {
  uint64_t camera_samples;
  uint64_t rays_closest; /* Scene::Intersect queries (path + MIS rays) */
  uint64_t rays_shadow;  /* Scene::IntersectP queries */
  uint64_t nodes_visited; /* only with PBRGPU_RENDER_COUNTERS */
  uint64_t tris_tested;   /* only with PBRGPU_RENDER_COUNTERS */
  uint64_t kernel_launches;
  double ms_total; /* CUDA-event time of the whole render on the launching stream */
  /* per-kernel CUDA-event times, only with PBRGPU_RENDER_TIMERS (sum over launches) */
  double ms_raygen;
  double ms_trace_closest;
  double ms_shade;
  double ms_trace_shadow;
  double ms_trace_mis;
  double ms_film;
  uint64_t n_trace_closest_launches;
  uint64_t rays_path; /* closest-hit rays that were path segments (rays_closest - MIS rays) */
} pbrgpu_stats;

/* ---- lifecycle ---- */
int pbrgpu_init(int device_id, pbrgpu_ctx** out);
void pbrgpu_shutdown(pbrgpu_ctx* ctx);
const char* pbrgpu_last_error(void);
int pbrgpu_device_count(void);

/* ---- multi-GPU: scene replicated per GPU, the sample range [s0,s1) of a render sharded across the GPUs,
 *      ONE ncclReduce(sum, fp32, root 0) of the RGBA sum films over NVLink, film resolve on the root.  The application still
 *      makes the single call Integrator::Render -> pbrgpu_render (Integrator/Integrator.h:14, Main/main.cpp:413). ----
 * (a) one process, N devices: a GROUP context.  device_ids == NULL: devices 0..n_devices-1; n_devices <= 0: all devices.
 *     Every entry point that takes this context / its scenes behaves as on one GPU; pbrgpu_render(_device) shards.
 *     With pbrgpu_render_device the accumulator lives on the FIRST device of the group. */
int pbrgpu_init_multi(int n_devices, const int* device_ids, pbrgpu_ctx** out);
/* (b) one process per GPU (torchrun, MPI): rank 0 obtains a 128-byte id, the launcher broadcasts it, every rank attaches its
 *     ordinary single-device context.  Afterwards every rank calls pbrgpu_render(_device) with the SAME arguments: each renders
 *     its share of [s0,s1), the films are reduced to rank 0, and only rank 0's output buffers are written (the other ranks may
 *     pass NULL).  PBRGPU_RENDER_ACCUMULATE is honoured on rank 0's accumulator. */
int pbrgpu_comm_unique_id(void* out128);
int pbrgpu_comm_attach(pbrgpu_ctx* ctx, int nranks, int rank, const void* unique_id128);
/* out[0] = devices driven by this context, out[1] = rank, out[2] = nranks, out[3] = NCCL version code (0 = NCCL not in use) */
int pbrgpu_ctx_info(pbrgpu_ctx* ctx, int32_t out[4]);

int pbrgpu_scene_create(pbrgpu_ctx* ctx, const pbrgpu_scene_desc* desc, pbrgpu_scene** out);
void pbrgpu_scene_destroy(pbrgpu_scene* scene);
int pbrgpu_camera_set(pbrgpu_scene* scene, const pbrgpu_camera_desc* cam);
int pbrgpu_sampler_set(pbrgpu_scene* scene, const pbrgpu_sampler_desc* smp);

/* Render sample range [s0,s1) of the pixel region.  out_rgb_sum (host, (x1-x0)*(y1-y0)*3 floats,
 * row-major, y down as raster) receives the SUM over samples of Li (the reference's `colObj` before
 * `/= spp`, Integrator.cpp:309).  out_rgba8 (optional, W*H*4) receives the tone-mapped 8-bit image of
 * sum/(s1-s0) with the reference's Y flip (Integrator.cpp:327-344). */
int pbrgpu_render(pbrgpu_scene* scene, const pbrgpu_render_desc* rd, float* out_rgb_sum,
                  unsigned char* out_rgba8, pbrgpu_stats* stats);

/* Same, but the sums stay in HBM: d_rgb_sum is a DEVICE pointer ((x1-x0)*(y1-y0)*4 floats, RGBA with
 * A = number of samples) on the context's device, e.g. a torch tensor's data_ptr() so that the
 * multi-GPU driver can NCCL-reduce it.  `stream` is a cudaStream_t; 0 means the context's own stream, which is created
 * with cudaStreamNonBlocking and is therefore NOT ordered against the legacy default stream: a caller that fills
 * d_rgb_sum on another stream (or on the default stream) must pass that stream, or synchronise before the call. */
int pbrgpu_render_device(pbrgpu_scene* scene, const pbrgpu_render_desc* rd, float* d_rgb_sum,
                         void* stream, pbrgpu_stats* stats);

/* Film resolve on device data (mean, XYZ->sRGB, gamma, 8-bit, Y flip; Integrator.cpp:313-344). */
int pbrgpu_resolve_film(pbrgpu_ctx* ctx, const float* d_rgba_sum, int width, int height,
                        float inv_spp, float* out_rgb_mean, unsigned char* out_rgba8);

/* ---- parity hooks (used by tests; cheap) ---- */
enum {
  PBRGPU_DUMP_SAMPLER = 0,      /* float[n_pix][n_samples][n_dims] : SampleDimension(index(pixel,s), d) */
  PBRGPU_DUMP_PRIMARY_RAYS = 1, /* float[n_pix][n_samples][8] : o.xyz, tMax, d.xyz, time */
  PBRGPU_DUMP_PRIMARY_HITS = 2, /* int32 prim (caller's triangle index, -1 miss) + float t,b0,b1 : [..][4] */
  PBRGPU_DUMP_PERMS = 3,        /* uint16[n_perms] */
  PBRGPU_DUMP_SHADOW = 4,       /* int32[n] occlusion of caller-supplied rays (see pbrgpu_trace_rays) */
  PBRGPU_DUMP_WAVE_HITS = 6,    /* same layout as PRIMARY_HITS, but produced by the PRODUCTION kernels of a render: k_wave_begin ->
                                   k_raygen -> k_trace (persistent warps, parked triangle groups, shared-memory hit record);
                                   PRIMARY_HITS uses the one-thread-per-ray helper loop.  region x samples <= 2^25 */
  PBRGPU_DUMP_HIT_DIFFERENTIALS = 5 /* float[n_pix][n_samples][8]: hit flag, uv, dudx, dvdx, dudy, dvdy of the camera ray's hit
                                       (Triangle.cpp:147-170, SurfaceInteraction::ComputeDifferentials) */
};
typedef struct pbrgpu_dump_desc {
  int32_t what;
  int32_t x0, y0, x1, y1;
  int64_t s0, s1;
  int32_t n_dims; /* SAMPLER only */
  int32_t pad;
} pbrgpu_dump_desc;
int pbrgpu_debug_dump(pbrgpu_scene* scene, const pbrgpu_dump_desc* dd, void* out, size_t out_bytes);

/* Trace caller-supplied rays (float[n][8]: o.xyz,tMax,d.xyz,pad).  any_hit=0: closest hit ->
 * out_hits int32/float [n][4] (prim, t, b0, b1 as bit patterns); any_hit=1: out_hits int32[n].
 * any_hit | PBRGPU_TRACE_VIA_WAVE_KERNEL: the rays go through the production traversal kernel k_trace (closest-hit rays
 * through its path AND MIS queues - the call fails if the two disagree -, any-hit rays through its shadow queue) instead
 * of the one-thread-per-ray helper loop; n <= 2^25. */
enum { PBRGPU_TRACE_VIA_WAVE_KERNEL = 2 };
int pbrgpu_trace_rays(pbrgpu_scene* scene, const float* rays, int64_t n, int any_hit, void* out_hits);

/* Parity hook: ImageTexture::Evaluate of textures[texture] for n surface points.  rows float[n][8] = u, v, dudx, dvdx, dudy,
 * dvdy, -, - (SurfaceInteraction::uv and its screen-space derivatives); out float[n][4] = rgb, -. */
int pbrgpu_texture_eval(pbrgpu_scene* scene, int texture, const float* rows, int64_t n, float* out);

/* Parity hook: Light::Sample_Li / Pdf_Li of lights[light] for n reference points.  rows float[n][16] = ref.p, ref.pError, ref.n,
 * u0, u1, query wi, -, -;  out float[n][16] = Li.rgb, pdf, wi, p1, n1, Pdf_Li(ref, query wi), -, -. */
int pbrgpu_light_eval(pbrgpu_scene* scene, int light, const float* rows, int64_t n, float* out);

/* Halton helper mirrors (host-side integer code; Halton.cpp:40-81).  Outputs 7 ints:
 * baseScales[2], baseExponents[2], sampleStride, multInverse[2]. */
int pbrgpu_halton_params(const int32_t sample_bounds[4], int32_t out7[7]);
/* Radical-inverse permutations for the first n_bases primes (n_bases<=1000); returns count written. */
int64_t pbrgpu_compute_perms(uint16_t* out, int64_t capacity);

/* Scene ingestion helper (host only): the reference's private ".3d" mesh text format, Shape/plyRead.h:17-47
 * ("vertex N face M", N vertices scaled by `scale` - the reference uses 20 -, M faces "<count> i0 i1 i2").
 * Returns malloc'd arrays (release with pbrgpu_free); non-zero on error. */
int pbrgpu_load_3d(const char* path, float scale, float** out_verts, int32_t* out_n_verts, int32_t** out_indices,
                   int32_t* out_n_tris);
/* Radiance ".hdr" (RGBE, flat or run-length coded scanlines) environment map, what InfiniteAreaLight / SkyBoxLight read
 * through stbi_loadf (Light/InfiniteAreaLight.cpp:17-18, Light/SkyBoxLight.cpp:19-22).  Returns a malloc'd
 * width*height*3 float array, row-major, top row first unless flip_vertically (= stbi_set_flip_vertically_on_load);
 * texels are bit-identical to stb's.  Non-zero on error (2 = cannot open, 3 = not a Radiance RGBE file, 5 = truncated). */
int pbrgpu_load_hdr(const char* path, int flip_vertically, float** out_rgb, int32_t* out_width, int32_t* out_height);
void pbrgpu_free(void* p);

/* Host-only helper: the level-0 image MIPMap builds from an image whose sides are not powers of two (separable 4-tap Lanczos
 * up to RoundUpPow2, Texture/MIPMap.h:45-64,93-144) - what pbrgpu_scene_create applies to such textures / env maps.
 * out_size receives the resampled width, height; out (may be NULL) width*height*3 floats. */
int pbrgpu_resample_pow2(const float* rgb, int32_t width, int32_t height, int32_t wrap, float* out, int64_t out_floats,
                         int32_t out_size[2]);
/* Host-only helper: the sample sub-range rank `rank` of `world` renders for [s0,s1) (contiguous, balanced; what
 * pbrgpu_render applies on a multi-GPU context). */
int pbrgpu_shard_samples(int64_t s0, int64_t s1, int rank, int world, int64_t out2[2]);
/* Measurement helper (bench.py): read bandwidth in GB/s of a `bytes`-sized device buffer re-read `iters` times
 * with 16-byte loads that bypass the L1.  bytes << L2 size (e.g. 48 MB on B200's 126 MB): L2 -> SM bandwidth, the bound of
 * the traversal kernel's node / triangle fetches; bytes >> L2: HBM read bandwidth. */
int pbrgpu_measure_read_bandwidth(pbrgpu_ctx* ctx, size_t bytes, int iters, double* out_gb_per_s);

/* BVH statistics: out[0]=n_nodes, out[1]=n_leaves, out[2]=max_leaf_size, out[3]=bytes_nodes, out[4]=bytes_tris, out[5]=depth of
 * the wide tree, out[6]=hash of the built nodes + leaf order (the host build is multi-threaded - PBRGPU_BUILD_THREADS, default
 * min(16, cores) - and its result does not depend on the thread count) */
int pbrgpu_scene_info(pbrgpu_scene* scene, int64_t out[8]);

#ifdef __cplusplus
}
#endif
#endif /* PBRGPU_H */
