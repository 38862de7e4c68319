// TEST INFRASTRUCTURE ONLY - never linked into the product.
//
// ref_harness.cpp: a C-ABI (ctypes-friendly) harness around the UNMODIFIED reference renderer
// (G0T-cha/PysicalBasedRaytracer), compiled by oracle/build_ref.sh against the staged reference
// headers into oracle/_ref/libpbr_ref.so.  It turns the same POD scene description the GPU library
// consumes (include/pbrgpu.h) into the reference's own object graph (TriangleMesh / Triangle /
// GeometricPrimitive / Material / Light / BVHAccel / Scene / PerspectiveCamera / HaltonSampler) and
// runs the reference's own public entry points on it:
//   sampler streams   HaltonSampler::{StartPixel,SetSampleNumber,Get1D}   Sampler/Sampler.cpp:97-143
//   primary rays      Camera::GenerateRayDifferential                    Camera/Perspective.cpp:44-80
//   primary hits      Scene::Intersect                                   Core/Scene.cpp:20-23
//   radiance          {Path,VolPath,DirectLighting,Whitted}Integrator::Li  Integrator/*.cpp
// It does NOT use SamplerIntegrator::Render for parity dumps (quirk Q1: swapped loop bounds); its own
// pixel loop iterates x in [x0,x1), y in [y0,y1) and returns linear-float sums.  ref_render_shipped()
// calls the reference's Render() verbatim (4 threads, square images) for the as-shipped CPU baseline.

#include <atomic>
#include <cstdio>
#include <map>
#include <unordered_map>
#include <omp.h>

#define private public
#define protected public
#include "Core/PBR.h"
#include "Core/Geometry.h"
#include "Core/Transform.h"
#include "Core/Spectrum.h"
#include "Core/Interaction.h"
#include "Core/Primitive.h"
#include "Core/Scene.h"
#include "Core/FrameBuffer.h"
#include "Shape/Shape.h"
#include "Shape/Triangle.h"
#include "Shape/plyRead.h"
#include "Accelerator/BVHAccel.h"
#include "Camera/Camera.h"
#include "Camera/Perspective.h"
#include "Camera/Orthographic.h"
#include "Sampler/Sampler.h"
#include "Sampler/Halton.h"
#include "Sampler/LowDiscrepancy.h"
#include "Sampler/Sampling.h"
#include "Texture/Texture.h"
#include "Texture/ConstantTexture.h"
#include "Texture/MIPMap.h"
#include "Texture/ImageTexture.h"
#include "Material/Material.h"
#include "Material/MatteMaterial.h"
#include "Material/Mirror.h"
#include "Material/GlassMaterial.h"
#include "Material/MetalMaterial.h"
#include "Material/PlasticMaterial.h"
#include "Material/Reflection.h"
#include "Light/Light.h"
#include "Light/PointLight.h"
#include "Light/DiffuseLight.h"
#include "Light/InfiniteAreaLight.h"
#include "Light/SkyBoxLight.h"
#include "Light/LightDistrib.h"
#include "Integrator/Integrator.h"
#include "Integrator/PathIntegrator.h"
#include "Integrator/WhittedIntegrator.h"
#include "Integrator/DirectLightingIntegrator.h"
#include "Integrator/VolPathIntegrator.h"
#include "Media/Medium.h"
#include "Media/HomogeneousMedium.h"
#undef private
#undef protected

#include "SobolSampler.h"  // our CPU twin of the GPU Sobol sampler (the reference has only the tables)
#include "AnalyticSphere.h"  // our CPU twin of the GPU analytic sphere (the reference's Sphere is a stub)
#include "pbrgpu.h"

namespace PBR {
OrthographicCamera* CreateOrthographicCamera(const int RasterWidth, const int RasterHeight,
                                             const Transform& cam2world, Medium* media);
}
using namespace PBR;

namespace {

// Forwarding aggregate that counts Scene::Intersect / IntersectP queries per thread.
struct CountingAggregate : public Aggregate {
  std::shared_ptr<Primitive> inner;
  mutable std::vector<uint64_t> nI, nP;  // padded per-thread counters
  explicit CountingAggregate(std::shared_ptr<Primitive> p) : inner(std::move(p)), nI(256 * 8), nP(256 * 8) {}
  Bounds3f WorldBound() const override { return inner->WorldBound(); }
  bool Intersect(const Ray& r, SurfaceInteraction* si) const override {
    nI[omp_get_thread_num() * 8]++;
    return inner->Intersect(r, si);
  }
  bool IntersectP(const Ray& r) const override {
    nP[omp_get_thread_num() * 8]++;
    return inner->IntersectP(r);
  }
  void reset() const { std::fill(nI.begin(), nI.end(), 0); std::fill(nP.begin(), nP.end(), 0); }
  uint64_t sumI() const { uint64_t s = 0; for (auto v : nI) s += v; return s; }
  uint64_t sumP() const { uint64_t s = 0; for (auto v : nP) s += v; return s; }
};

struct RefScene {
  Transform identity;
  std::vector<std::shared_ptr<TriangleMesh>> meshes;
  std::vector<std::shared_ptr<Shape>> tris;
  std::vector<std::shared_ptr<Shape>> spheres;                // AnalyticSphere twins (pbrgpu_sphere)
  std::vector<std::unique_ptr<Transform>> sphereXf;            // their ObjectToWorld / WorldToObject (translations)
  std::vector<std::shared_ptr<Material>> materials;
  std::vector<std::unique_ptr<HomogeneousMedium>> media;
  std::vector<std::shared_ptr<Light>> lights;
  std::vector<std::shared_ptr<Primitive>> prims;  // caller's order
  std::unordered_map<const Primitive*, int> primIndex;
  std::shared_ptr<BVHAccel> bvh;
  std::shared_ptr<CountingAggregate> agg;
  std::unique_ptr<Scene> scene;
  double bvhBuildSeconds = 0;
  std::string tmpfile;
};

Spectrum rgb(const float c[3]) {
  Spectrum s;
  s[0] = c[0]; s[1] = c[1]; s[2] = c[2];
  return s;
}
std::shared_ptr<Texture<Spectrum>> ctex(const float c[3]) {
  return std::make_shared<ConstantTexture<Spectrum>>(rgb(c));
}
std::shared_ptr<Texture<float>> ftex(float v) { return std::make_shared<ConstantTexture<float>>(v); }

// ImageTexture<RGBSpectrum, Spectrum> over the caller's texels.  The class only loads files (ImageTexture.cpp:10-37), so it
// is constructed on a non-existent file (-> its 1x1 placeholder) and the placeholder MIPMap is replaced by the reference's own
// MIPMap built from the texels (pyramid, EWA table: MIPMap.h:86-178) - the same trick as for the InfiniteAreaLight below.
std::shared_ptr<Texture<Spectrum>> makeImageTexture(const pbrgpu_texture& t) {
  static std::atomic<int> counter{0};
  std::unique_ptr<TextureMapping2D> map(new UVMapping2D(t.su, t.sv, t.du, t.dv));
  const std::string name = "/nonexistent/pbr_ref_texture_" + std::to_string(counter++);
  const ImageWrap wrap = t.wrap == PBRGPU_WRAP_CLAMP ? ImageWrap::Clamp : (t.wrap == PBRGPU_WRAP_BLACK ? ImageWrap::Black : ImageWrap::Repeat);
  std::streambuf* old = std::cerr.rdbuf(nullptr);
  auto tex = std::make_shared<ImageTexture<RGBSpectrum, Spectrum>>(std::move(map), name, t.do_trilinear != 0, t.max_anisotropy, wrap, 1.f, false);
  std::cerr.rdbuf(old);
  std::vector<RGBSpectrum> texels((size_t)t.width * t.height);
  for (size_t k = 0; k < texels.size(); ++k)
    for (int c = 0; c < 3; ++c) texels[k][c] = t.texels[3 * k + c];
  tex->mipmap = new MIPMap<RGBSpectrum>(Point2i(t.width, t.height), texels.data(), t.do_trilinear != 0, t.max_anisotropy, wrap);
  return tex;
}

// ImageTexture<float, float> over the first channel of the caller's texels (same construction)
std::shared_ptr<Texture<float>> makeFloatImageTexture(const pbrgpu_texture& t) {
  static std::atomic<int> counter{0};
  std::unique_ptr<TextureMapping2D> map(new UVMapping2D(t.su, t.sv, t.du, t.dv));
  const std::string name = "/nonexistent/pbr_ref_float_texture_" + std::to_string(counter++);
  const ImageWrap wrap = t.wrap == PBRGPU_WRAP_CLAMP ? ImageWrap::Clamp : (t.wrap == PBRGPU_WRAP_BLACK ? ImageWrap::Black : ImageWrap::Repeat);
  std::streambuf* old = std::cerr.rdbuf(nullptr);
  auto tex = std::make_shared<ImageTexture<float, float>>(std::move(map), name, t.do_trilinear != 0, t.max_anisotropy, wrap, 1.f, false);
  std::cerr.rdbuf(old);
  std::vector<float> texels((size_t)t.width * t.height);
  for (size_t k = 0; k < texels.size(); ++k) texels[k] = t.texels[3 * k];
  tex->mipmap = new MIPMap<float>(Point2i(t.width, t.height), texels.data(), t.do_trilinear != 0, t.max_anisotropy, wrap);
  return tex;
}

std::shared_ptr<Material> makeMaterial(const pbrgpu_material& m, const pbrgpu_scene_desc* d) {
  auto bump = ftex(0.f);
  std::shared_ptr<Texture<Spectrum>> kd = m.kd_texture_plus1 > 0 ? makeImageTexture(d->textures[m.kd_texture_plus1 - 1]) : nullptr;
  // Ks of a Plastic: the SAME texture object when it names the same texture (Main/main.cpp:63-77 passes one ImageTexture twice)
  std::shared_ptr<Texture<Spectrum>> ks = m.ks_texture_plus1 > 0
      ? (m.ks_texture_plus1 == m.kd_texture_plus1 ? kd : makeImageTexture(d->textures[m.ks_texture_plus1 - 1])) : nullptr;
  switch (m.type) {
    case PBRGPU_MAT_MATTE: return std::make_shared<MatteMaterial>(kd ? kd : ctex(m.kd), ftex(m.sigma), bump);
    case PBRGPU_MAT_MIRROR: return std::make_shared<MirrorMaterial>(ctex(m.kr), bump);
    case PBRGPU_MAT_GLASS:
      return std::make_shared<GlassMaterial>(ctex(m.kr), ctex(m.kt), ftex(m.u_rough), ftex(m.v_rough),
                                             ftex(m.eta[0]), bump, m.remap_roughness != 0);
    case PBRGPU_MAT_METAL:
      return std::make_shared<MetalMaterial>(ctex(m.eta), ctex(m.k), ftex(m.u_rough), ftex(m.u_rough),
                                             ftex(m.v_rough), bump, m.remap_roughness != 0);
    case PBRGPU_MAT_PLASTIC:
      return std::make_shared<PlasticMaterial>(kd ? kd : ctex(m.kd), ks ? ks : ctex(m.kr),
                                               m.rough_texture_plus1 > 0 ? makeFloatImageTexture(d->textures[m.rough_texture_plus1 - 1]) : ftex(m.u_rough), bump,
                                               m.remap_roughness != 0);
  }
  return nullptr;
}

Transform fromRowMajor(const float m[16], const float minv[16]) {
  Matrix4x4 a, b;
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) { a.m[i][j] = m[i * 4 + j]; b.m[i][j] = minv[i * 4 + j]; }
  return Transform(a, b);
}

struct RefCamera { std::shared_ptr<Camera> cam; int W, H, type; bool honourMedium = false; };
struct RefSampler { std::shared_ptr<Sampler> s; };
std::shared_ptr<Sampler> makeSampler(int type, int spp, const Bounds2i& b) {
  if (type == PBRGPU_SAMPLER_SOBOL) return std::make_shared<SobolSampler>(spp, b);
  return std::make_shared<HaltonSampler>(spp, b);
}

}  // namespace

extern "C" {

void* ref_scene_create(const pbrgpu_scene_desc* d) {
  auto* rs = new RefScene;
  // one TriangleMesh for the whole soup: vertices are already in world space -> identity transform
  // (Transform::operator() with an identity matrix returns the point unchanged: wp == 1 branch).
  std::vector<Point3f> P(d->n_verts);
  for (int i = 0; i < d->n_verts; ++i) P[i] = Point3f(d->verts[3 * i], d->verts[3 * i + 1], d->verts[3 * i + 2]);
  std::vector<Normal3f> N;
  if (d->normals) {
    N.resize(d->n_verts);
    for (int i = 0; i < d->n_verts; ++i) N[i] = Normal3f(d->normals[3 * i], d->normals[3 * i + 1], d->normals[3 * i + 2]);
  }
  std::vector<Point2f> UV;
  if (d->uvs) {
    UV.resize(d->n_verts);
    for (int i = 0; i < d->n_verts; ++i) UV[i] = Point2f(d->uvs[2 * i], d->uvs[2 * i + 1]);
  }
  std::vector<Vector3f> TG;
  if (d->tangents) {
    TG.resize(d->n_verts);
    for (int i = 0; i < d->n_verts; ++i) TG[i] = Vector3f(d->tangents[3 * i], d->tangents[3 * i + 1], d->tangents[3 * i + 2]);
  }
  // tri_flags bit2 / bit3 / bit4 say whether a triangle's mesh has per-vertex normals / uvs / tangents: one TriangleMesh per
  // combination in use (they share the vertex and index arrays), so that triangles of meshes WITHOUT uvs get Triangle::GetUVs' defaults
  std::shared_ptr<TriangleMesh> variant[8];
  auto meshFor = [&](int i) {
    int key = 0;
    if (d->tri_flags) key = ((d->tri_flags[i] & 4) && d->normals ? 1 : 0) | ((d->tri_flags[i] & 8) && d->uvs ? 2 : 0) |
                            ((d->tri_flags[i] & 16) && d->tangents ? 4 : 0);
    else key = (d->normals ? 1 : 0) | (d->uvs ? 2 : 0) | (d->tangents ? 4 : 0);
    if (!variant[key]) {
      variant[key] = std::make_shared<TriangleMesh>(rs->identity, d->n_tris, d->indices, d->n_verts, P.data(), (key & 4) ? TG.data() : nullptr,
                                                    (key & 1) ? N.data() : nullptr, (key & 2) ? UV.data() : nullptr, nullptr);
      rs->meshes.push_back(variant[key]);
    }
    return variant[key];
  };
  for (int i = 0; i < d->n_materials; ++i) rs->materials.push_back(makeMaterial(d->materials[i], d));
  for (int i = 0; i < d->n_media; ++i)
    rs->media.emplace_back(new HomogeneousMedium(rgb(d->media[i].sigma_a), rgb(d->media[i].sigma_s), d->media[i].g));
  rs->tris.reserve(d->n_tris);
  for (int i = 0; i < d->n_tris; ++i) {
    bool rev = d->tri_flags ? (d->tri_flags[i] & 1) != 0 : false;
    rs->tris.push_back(std::make_shared<Triangle>(&rs->identity, &rs->identity, rev, meshFor(i), i));
  }
  for (int i = 0; i < d->n_spheres; ++i) {
    const pbrgpu_sphere& sp = d->spheres[i];
    rs->sphereXf.emplace_back(new Transform(Translate(Vector3f(sp.center[0], sp.center[1], sp.center[2]))));
    Transform* o2w = rs->sphereXf.back().get();
    rs->sphereXf.emplace_back(new Transform(Inverse(*o2w)));
    rs->spheres.push_back(std::make_shared<AnalyticSphere>(o2w, rs->sphereXf.back().get(), sp.reverse_orientation != 0, sp.radius));
  }
  // lights (one entry per emissive triangle, as the reference does: Main/main.cpp:353-375)
  rs->lights.resize(d->n_lights);
  std::vector<std::shared_ptr<AreaLight>> areaOfLight(d->n_lights);
  for (int i = 0; i < d->n_lights; ++i) {
    const pbrgpu_light& L = d->lights[i];
    switch (L.type) {
      case PBRGPU_LIGHT_POINT:
        rs->lights[i] = std::make_shared<PointLight>(Translate(Vector3f(L.pos[0], L.pos[1], L.pos[2])),
                                                     MediumInterface(), rgb(L.L));
        break;
      case PBRGPU_LIGHT_AREA_TRI: {
        auto a = std::make_shared<DiffuseAreaLight>(rs->identity, MediumInterface(), rgb(L.L), L.n_samples,
                                                    rs->tris[L.tri], L.two_sided != 0);
        areaOfLight[i] = a;
        rs->lights[i] = a;
        break;
      }
      case PBRGPU_LIGHT_AREA_SPHERE: {
        auto a = std::make_shared<DiffuseAreaLight>(rs->identity, MediumInterface(), rgb(L.L), L.n_samples,
                                                    rs->spheres[L.tri], L.two_sided != 0);
        areaOfLight[i] = a;
        rs->lights[i] = a;
        break;
      }
      case PBRGPU_LIGHT_INFINITE: {
        // The ctor insists on loading a file through stb; give it a 1x1 Radiance file, then swap in
        // the caller's float texels through the reference's own MIPMap / Distribution2D classes,
        // repeating InfiniteAreaLight.cpp:20-31 (premultiply by L) and :40-59 (sampling image).
        char name[] = "/tmp/pbr_ref_env_XXXXXX.hdr";
        int fd = mkstemps(name, 4);
        const char hdr[] = "#?RADIANCE\nFORMAT=32-bit_rle_rgbe\n\n-Y 1 +X 1\n";
        unsigned char px[4] = {128, 128, 128, 129};
        if (fd >= 0) { (void)!write(fd, hdr, sizeof(hdr) - 1); (void)!write(fd, px, 4); close(fd); }
        Transform l2w = fromRowMajor(L.light_to_world, L.world_to_light);
        auto il = std::make_shared<InfiniteAreaLight>(l2w, rgb(L.L), L.n_samples, std::string(name));
        unlink(name);
        int W = d->env_width, H = d->env_height;
        std::unique_ptr<RGBSpectrum[]> tex(new RGBSpectrum[(size_t)W * H]);
        for (size_t k = 0; k < (size_t)W * H; ++k) {
          RGBSpectrum r;
          for (int c = 0; c < 3; ++c) { r[c] = d->env_texels[3 * k + c]; r[c] = L.L[c] * r[c]; }
          tex[k] = r;
        }
        il->Lmap.reset(new MIPMap<RGBSpectrum>(Point2i(W, H), tex.get()));
        int width = il->Lmap->Width(), height = il->Lmap->Height();
        std::unique_ptr<float[]> img(new float[(size_t)width * height]);
        for (int v = 0; v < height; v++) {
          float vp = (v + .5f) / (float)height;
          float sinTheta = std::sin(Pi * (v + .5f) / height);
          for (int u = 0; u < width; ++u) {
            float up = (u + .5f) / (float)width;
            img[u + v * width] = il->Lmap->Lookup(Point2f(up, vp), 0.f).y();
            img[u + v * width] *= sinTheta;
          }
        }
        il->distribution.reset(new Distribution2D(img.get(), width, height));
        rs->lights[i] = il;
        break;
      }
      case PBRGPU_LIGHT_SKYBOX: {
        auto sb = std::make_shared<SkyBoxLight>(rs->identity, Point3f(0, 0, 0), L.world_radius, "/nonexistent", L.n_samples);
        size_t n = (size_t)d->env_width * d->env_height * 3;
        sb->data = (float*)malloc(n * sizeof(float));
        memcpy(sb->data, d->env_texels, n * sizeof(float));
        sb->imageWidth = d->env_width; sb->imageHeight = d->env_height; sb->nrComponents = 3;
        rs->lights[i] = sb;
        break;
      }
    }
  }
  rs->prims.reserve(d->n_tris);
  for (int i = 0; i < d->n_tris; ++i) {
    std::shared_ptr<AreaLight> al = d->tri_light[i] >= 0 ? areaOfLight[d->tri_light[i]] : nullptr;
    const Medium* mIn = d->tri_medium_inside && d->tri_medium_inside[i] >= 0 ? rs->media[d->tri_medium_inside[i]].get() : nullptr;
    const Medium* mOut = d->tri_medium_outside && d->tri_medium_outside[i] >= 0 ? rs->media[d->tri_medium_outside[i]].get() : nullptr;
    std::shared_ptr<Material> mat = d->tri_material[i] >= 0 ? rs->materials[d->tri_material[i]] : nullptr;
    auto gp = std::make_shared<GeometricPrimitive>(rs->tris[i], mat, al, MediumInterface(mIn, mOut));
    rs->primIndex[gp.get()] = i;
    rs->prims.push_back(gp);
  }
  for (int i = 0; i < d->n_spheres; ++i) {  // primitive ids n_tris .. n_tris + n_spheres - 1
    const pbrgpu_sphere& sp = d->spheres[i];
    std::shared_ptr<AreaLight> al = sp.light >= 0 ? areaOfLight[sp.light] : nullptr;
    auto gp = std::make_shared<GeometricPrimitive>(rs->spheres[i], rs->materials[sp.material], al, MediumInterface());
    rs->primIndex[gp.get()] = d->n_tris + i;
    rs->prims.push_back(gp);
  }
  double t0 = omp_get_wtime();
  rs->bvh = std::make_shared<BVHAccel>(rs->prims, 1, BVHAccel::SplitMethod::SAH);  // Main/main.cpp:385
  rs->bvhBuildSeconds = omp_get_wtime() - t0;
  rs->agg = std::make_shared<CountingAggregate>(rs->bvh);
  rs->scene.reset(new Scene(rs->agg, rs->lights));
  return rs;
}
void ref_scene_destroy(void* h) { delete (RefScene*)h; }
double ref_scene_bvh_seconds(void* h) { return ((RefScene*)h)->bvhBuildSeconds; }
void ref_scene_bound(void* h, float out[6]) {
  const Bounds3f& b = ((RefScene*)h)->scene->WorldBound();
  out[0] = b.pMin.x; out[1] = b.pMin.y; out[2] = b.pMin.z; out[3] = b.pMax.x; out[4] = b.pMax.y; out[5] = b.pMax.z;
}

// camera = Create{Perspective,Orthographic}Camera(W,H, Inverse(LookAt(eye,look,up)))   Main/main.cpp:196-200
void* ref_camera_create(int type, int W, int H, const float eye[3], const float look[3], const float up[3]) {
  Transform lookat = LookAt(Point3f(eye[0], eye[1], eye[2]), Point3f(look[0], look[1], look[2]), Vector3f(up[0], up[1], up[2]));
  Transform c2w = Inverse(lookat);
  auto* rc = new RefCamera;
  rc->W = W; rc->H = H; rc->type = type;
  if (type == PBRGPU_CAMERA_PERSPECTIVE) rc->cam.reset(CreatePerspectiveCamera(W, H, c2w, nullptr));
  else rc->cam.reset(CreateOrthographicCamera(W, H, c2w, nullptr));
  return rc;
}
// thin-lens variant: the factories fix lensRadius = 0 (Perspective.cpp:84-104), the class itself supports a lens
void* ref_camera_create_lens(int type, int W, int H, const float eye[3], const float look[3], const float up[3], float lensRadius,
                             float focalDistance) {
  Transform c2w = Inverse(LookAt(Point3f(eye[0], eye[1], eye[2]), Point3f(look[0], look[1], look[2]), Vector3f(up[0], up[1], up[2])));
  float frame = (float)W / (float)H;
  Bounds2f screen;
  if (frame > 1.f) { screen.pMin.x = -frame; screen.pMax.x = frame; screen.pMin.y = -1.f; screen.pMax.y = 1.f; }
  else { screen.pMin.x = -1.f; screen.pMax.x = 1.f; screen.pMin.y = -1.f / frame; screen.pMax.y = 1.f / frame; }
  auto* rc = new RefCamera;
  rc->W = W; rc->H = H; rc->type = type;
  if (type == PBRGPU_CAMERA_PERSPECTIVE) rc->cam.reset(new PerspectiveCamera(W, H, c2w, screen, lensRadius, focalDistance, 90.0f, nullptr));
  else {
    screen.pMin.x *= 2.f; screen.pMax.x *= 2.f; screen.pMin.y *= 2.f; screen.pMax.y *= 2.f;
    rc->cam.reset(new OrthographicCamera(W, H, c2w, screen, lensRadius, focalDistance, nullptr));
  }
  return rc;
}
void ref_camera_destroy(void* h) { delete (RefCamera*)h; }
// Camera::medium (Camera.h:29): the medium the camera sits in (index into the scene's media, -1 = vacuum).
// The reference never delivers it: PerspectiveCamera does not set ray->medium at all, OrthographicCamera sets it
// BEFORE `*ray = CameraToWorld(*ray)` and Transform::operator()(const Ray&) rebuilds the ray without it
// (Core/Transform.h:169).  honour != 0 makes the harness loop re-attach it after GenerateRayDifferential, so that the
// GPU library's camera-medium support can still be checked against the reference's own VolPathIntegrator::Li.
void ref_camera_set_medium(void* h, void* hscene, int medium, int honour) {
  auto* rc = (RefCamera*)h;
  auto* rs = (RefScene*)hscene;
  rc->cam->medium = medium >= 0 ? rs->media[medium].get() : nullptr;
  rc->honourMedium = honour != 0;
}
void ref_camera_desc(void* h, pbrgpu_camera_desc* out) {
  auto* rc = (RefCamera*)h;
  auto* pc = dynamic_cast<const ProjectiveCamera*>(rc->cam.get());
  out->type = rc->type; out->width = rc->W; out->height = rc->H;
  out->lens_radius = pc->lensRadius; out->focal_distance = pc->focalDistance;
  out->medium_plus1 = 0;  // the caller knows the scene's medium indices, the camera holds only a pointer
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) {
      out->raster_to_camera[i * 4 + j] = pc->RasterToCamera.m.m[i][j];
      out->camera_to_world[i * 4 + j] = pc->CameraToWorld.m.m[i][j];
    }
}

void* ref_sampler_create(int spp, int x0, int y0, int x1, int y1) {
  auto* s = new RefSampler;
  s->s = std::make_shared<HaltonSampler>(spp, Bounds2i(Point2i(x0, y0), Point2i(x1, y1)));
  return s;
}
void* ref_sampler_create_typed(int type, int spp, int x0, int y0, int x1, int y1) {
  auto* s = new RefSampler;
  s->s = makeSampler(type, spp, Bounds2i(Point2i(x0, y0), Point2i(x1, y1)));
  return s;
}
// The reference's Sobol generator tables (Sampler/SobolMatrices.h:42-47), for handing to pbrgpu_sampler_set.
void ref_sobol_tables(uint32_t* m32, int log2Resolution, uint64_t* vdc, uint64_t* vdcInv) {
  if (m32) memcpy(m32, SobolMatrices32, sizeof(uint32_t) * NumSobolDimensions * SobolMatrixSize);
  if (log2Resolution >= 1) {
    if (vdc) memcpy(vdc, VdCSobolMatrices[log2Resolution - 1], sizeof(uint64_t) * SobolMatrixSize);
    if (vdcInv) memcpy(vdcInv, VdCSobolMatricesInv[log2Resolution - 1], sizeof(uint64_t) * SobolMatrixSize);
  }
}
void ref_sampler_destroy(void* h) { delete (RefSampler*)h; }
void ref_halton_params(void* h, int32_t out7[7]) {
  auto& s = *dynamic_cast<HaltonSampler*>(((RefSampler*)h)->s.get());
  out7[0] = s.baseScales[0]; out7[1] = s.baseScales[1]; out7[2] = s.baseExponents[0]; out7[3] = s.baseExponents[1];
  out7[4] = s.sampleStride; out7[5] = s.multInverse[0]; out7[6] = s.multInverse[1];
}
int64_t ref_perms(uint16_t* out, int64_t cap) {
  RNG rng;
  std::vector<uint16_t> p = ComputeRadicalInversePermutations(rng);
  int64_t n = std::min<int64_t>(cap, (int64_t)p.size());
  if (out) memcpy(out, p.data(), n * sizeof(uint16_t));
  return (int64_t)p.size();
}
float ref_radical_inverse(int baseIndex, uint64_t a) { return RadicalInverse(baseIndex, a); }
float ref_scrambled_radical_inverse(int baseIndex, uint64_t a) {
  static std::vector<uint16_t> p;
  if (p.empty()) { RNG rng; p = ComputeRadicalInversePermutations(rng); }
  return ScrambledRadicalInverse(baseIndex, a, &p[PrimeSums[baseIndex]]);
}
int ref_prime(int i) { return Primes[i]; }
int ref_prime_sum(int i) { return PrimeSums[i]; }

// out[pixel][sample][dim] for pixels of region (row-major), samples s0..s1, first n_dims Get1D() draws.
void ref_sampler_dump(void* h, int x0, int y0, int x1, int y1, int64_t s0, int64_t s1, int n_dims, float* out) {
  auto& proto = *((RefSampler*)h)->s;
  size_t k = 0;
  for (int y = y0; y < y1; ++y)
    for (int x = x0; x < x1; ++x) {
      std::unique_ptr<Sampler> ps = proto.Clone(0);
      ps->StartPixel(Point2i(x, y));
      for (int64_t s = s0; s < s1; ++s) {
        ps->SetSampleNumber(s);
        for (int d = 0; d < n_dims; ++d) out[k++] = ps->Get1D();
      }
    }
}
int64_t ref_halton_index(void* h, int x, int y, int64_t s) {
  auto& proto = *((RefSampler*)h)->s;
  std::unique_ptr<Sampler> ps = proto.Clone(0);
  ps->StartPixel(Point2i(x, y));
  return dynamic_cast<GlobalSampler*>(ps.get())->GetIndexForSample(s);
}

// out[pixel][sample][8] = o.xyz, tMax, d.xyz, time   (the `Ray ray(r)` slice PathIntegrator::Li starts from)
void ref_primary_rays(void* hc, void* hs, int x0, int y0, int x1, int y1, int64_t s0, int64_t s1, float* out) {
  auto* rc = (RefCamera*)hc;
  auto& proto = *((RefSampler*)hs)->s;
  size_t k = 0;
  for (int y = y0; y < y1; ++y)
    for (int x = x0; x < x1; ++x) {
      std::unique_ptr<Sampler> ps = proto.Clone(0);
      Point2i pixel(x, y);
      ps->StartPixel(pixel);
      for (int64_t s = s0; s < s1; ++s) {
        ps->SetSampleNumber(s);
        CameraSample cs = ps->GetCameraSample(pixel);
        RayDifferential r;
        rc->cam->GenerateRayDifferential(cs, &r);
        out[k++] = r.o.x; out[k++] = r.o.y; out[k++] = r.o.z; out[k++] = r.tMax;
        out[k++] = r.d.x; out[k++] = r.d.y; out[k++] = r.d.z; out[k++] = r.time;
      }
    }
}

// out[pixel][sample][4] = prim index (int bits; -1 miss), tHit, uv.x?, uv.y? -> we return t and the hit point's
// barycentric-free identifiers: (prim, t, n.x sign marker unused, 0).  Only prim and t are compared.
void ref_primary_hits(void* hscene, void* hc, void* hs, int x0, int y0, int x1, int y1, int64_t s0, int64_t s1, int32_t* out) {
  auto* rs = (RefScene*)hscene;
  auto* rc = (RefCamera*)hc;
  auto& proto = *((RefSampler*)hs)->s;
  int W = x1 - x0;
  int64_t ns = s1 - s0;
#pragma omp parallel for schedule(dynamic, 1)
  for (int y = y0; y < y1; ++y)
    for (int x = x0; x < x1; ++x) {
      std::unique_ptr<Sampler> ps = proto.Clone(0);
      Point2i pixel(x, y);
      ps->StartPixel(pixel);
      for (int64_t s = s0; s < s1; ++s) {
        ps->SetSampleNumber(s);
        CameraSample cs = ps->GetCameraSample(pixel);
        RayDifferential r;
        rc->cam->GenerateRayDifferential(cs, &r);
        Ray ray(r);
        SurfaceInteraction isect;
        size_t k = (((size_t)(y - y0) * W + (x - x0)) * ns + (s - s0)) * 4;
        if (rs->scene->Intersect(ray, &isect)) {
          out[k] = rs->primIndex[isect.primitive];
          float t = ray.tMax;
          memcpy(&out[k + 1], &t, 4);
        } else {
          out[k] = -1;
          float t = ray.tMax;
          memcpy(&out[k + 1], &t, 4);
        }
        out[k + 2] = 0; out[k + 3] = 0;
      }
    }
}

// Trace caller rays through Scene::Intersect / IntersectP.  rays float[n][8] (o,tMax,d,pad).
void ref_trace_rays(void* hscene, const float* rays, int64_t n, int any_hit, int32_t* out) {
  auto* rs = (RefScene*)hscene;
#pragma omp parallel for schedule(static, 1024)
  for (int64_t i = 0; i < n; ++i) {
    const float* r = rays + 8 * i;
    Ray ray(Point3f(r[0], r[1], r[2]), Vector3f(r[4], r[5], r[6]), r[3]);
    if (any_hit) {
      out[i] = rs->scene->IntersectP(ray) ? 1 : 0;
    } else {
      SurfaceInteraction isect;
      bool hit = rs->scene->Intersect(ray, &isect);
      out[4 * i] = hit ? rs->primIndex[isect.primitive] : -1;
      float t = ray.tMax;
      memcpy(&out[4 * i + 1], &t, 4);
      out[4 * i + 2] = out[4 * i + 3] = 0;
    }
  }
}

struct RefRenderStats { uint64_t samples, rays_closest, rays_shadow; double seconds; int threads; };

// Linear-float SUM over samples [s0,s1) of Li for each pixel of the region; own pixel loop over the
// reference's public API (Clone / StartPixel / SetSampleNumber / GetCameraSample /
// GenerateRayDifferential / ScaleDifferentials / Li), OpenMP over rows with `threads` threads.
int ref_render_typed(void* hscene, void* hc, int sampler_type, int spp, const int32_t sample_bounds[4], const pbrgpu_render_desc* rd,
                     int threads, float* out_rgb_sum, RefRenderStats* st);
int ref_render(void* hscene, void* hc, int spp, const int32_t sample_bounds[4], const pbrgpu_render_desc* rd,
               int threads, float* out_rgb_sum, RefRenderStats* st) {
  return ref_render_typed(hscene, hc, PBRGPU_SAMPLER_HALTON, spp, sample_bounds, rd, threads, out_rgb_sum, st);
}
int ref_render_typed(void* hscene, void* hc, int sampler_type, int spp, const int32_t sample_bounds[4], const pbrgpu_render_desc* rd,
                     int threads, float* out_rgb_sum, RefRenderStats* st) {
  auto* rs = (RefScene*)hscene;
  auto* rc = (RefCamera*)hc;
  Bounds2i sb(Point2i(sample_bounds[0], sample_bounds[1]), Point2i(sample_bounds[2], sample_bounds[3]));
  std::shared_ptr<Sampler> sampler = makeSampler(sampler_type, spp, sb);
  Bounds2i pb(Point2i(rd->x0, rd->y0), Point2i(rd->x1, rd->y1));
  std::unique_ptr<SamplerIntegrator> integ;
  const char* strat = rd->light_strategy == PBRGPU_LIGHTSTRATEGY_POWER ? "power" : "uniform";
  switch (rd->integrator) {
    case PBRGPU_INTEGRATOR_PATH:
      integ.reset(new PathIntegrator(rd->max_depth, rc->cam, sampler, pb, rd->rr_threshold, strat, nullptr));
      break;
    case PBRGPU_INTEGRATOR_DIRECT_ONE:
      integ.reset(new DirectLightingIntegrator(LightStrategy::UniformSampleOne, rd->max_depth, rc->cam, sampler, pb, nullptr));
      break;
    case PBRGPU_INTEGRATOR_DIRECT_ALL:
      integ.reset(new DirectLightingIntegrator(LightStrategy::UniformSampleAll, rd->max_depth, rc->cam, sampler, pb, nullptr));
      break;
    case PBRGPU_INTEGRATOR_WHITTED:
      integ.reset(new WhittedIntegrator(rd->max_depth, rc->cam, sampler, pb, nullptr));
      break;
    case PBRGPU_INTEGRATOR_VOLPATH:
      integ.reset(new VolPathIntegrator(rd->max_depth, rc->cam, sampler, pb, rd->rr_threshold, strat, nullptr));
      break;
    default: return 1;
  }
  if (threads <= 0) threads = omp_get_num_procs();
  omp_set_num_threads(threads);
  rs->agg->reset();
  double t0 = omp_get_wtime();
  integ->Preprocess(*rs->scene, *sampler);
  int W = rd->x1 - rd->x0;
#pragma omp parallel for schedule(dynamic, 1)
  for (int y = rd->y0; y < rd->y1; ++y)
    for (int x = rd->x0; x < rd->x1; ++x) {
      std::unique_ptr<Sampler> ps = sampler->Clone(0);
      Point2i pixel(x, y);
      ps->StartPixel(pixel);
      Spectrum col(0.f);
      for (int64_t s = rd->s0; s < rd->s1; ++s) {
        ps->SetSampleNumber(s);
        CameraSample cs = ps->GetCameraSample(pixel);
        RayDifferential r;
        rc->cam->GenerateRayDifferential(cs, &r);
        if (rc->honourMedium) r.medium = rc->cam->medium;
        r.ScaleDifferentials(1 / std::sqrt((float)ps->samplesPerPixel));
        col += integ->Li(r, *rs->scene, *ps, 0);
      }
      size_t k = ((size_t)(y - rd->y0) * W + (x - rd->x0)) * 3;
      if (rd->flags & PBRGPU_RENDER_ACCUMULATE) { out_rgb_sum[k] += col[0]; out_rgb_sum[k + 1] += col[1]; out_rgb_sum[k + 2] += col[2]; }
      else { out_rgb_sum[k] = col[0]; out_rgb_sum[k + 1] = col[1]; out_rgb_sum[k + 2] = col[2]; }
    }
  double t1 = omp_get_wtime();
  if (st) {
    st->samples = (uint64_t)(rd->x1 - rd->x0) * (rd->y1 - rd->y0) * (rd->s1 - rd->s0);
    st->rays_closest = rs->agg->sumI();
    st->rays_shadow = rs->agg->sumP();
    st->seconds = t1 - t0;
    st->threads = threads;
  }
  return 0;
}

// The reference's OWN SamplerIntegrator::Render, verbatim (4 threads hard-coded, square images only,
// Integrator.cpp:280-356).  Returns wall seconds reported by the reference; rgba8 = FrameBuffer bytes.
double ref_render_shipped_integrator(void* hscene, void* hc, int integrator, int spp, int W, int H, int max_depth, float rr,
                                     unsigned char* rgba8, RefRenderStats* st);
double ref_render_shipped(void* hscene, void* hc, int spp, int W, int H, int max_depth, float rr, unsigned char* rgba8,
                          RefRenderStats* st) {
  return ref_render_shipped_integrator(hscene, hc, PBRGPU_INTEGRATOR_PATH, spp, W, H, max_depth, rr, rgba8, st);
}
double ref_render_shipped_integrator(void* hscene, void* hc, int integrator, int spp, int W, int H, int max_depth, float rr,
                                     unsigned char* rgba8, RefRenderStats* st) {
  auto* rs = (RefScene*)hscene;
  auto* rc = (RefCamera*)hc;
  Bounds2i b(Point2i(0, 0), Point2i(W, H));
  std::shared_ptr<Sampler> sampler = std::make_shared<HaltonSampler>(spp, b);
  FrameBuffer fb;
  fb.InitBuffer(W, H, 4);
  std::unique_ptr<SamplerIntegrator> integ;
  switch (integrator) {
    case PBRGPU_INTEGRATOR_WHITTED: integ.reset(new WhittedIntegrator(max_depth, rc->cam, sampler, b, &fb)); break;
    case PBRGPU_INTEGRATOR_DIRECT_ONE: integ.reset(new DirectLightingIntegrator(LightStrategy::UniformSampleOne, max_depth, rc->cam, sampler, b, &fb)); break;
    case PBRGPU_INTEGRATOR_DIRECT_ALL: integ.reset(new DirectLightingIntegrator(LightStrategy::UniformSampleAll, max_depth, rc->cam, sampler, b, &fb)); break;
    case PBRGPU_INTEGRATOR_VOLPATH: integ.reset(new VolPathIntegrator(max_depth, rc->cam, sampler, b, rr, "uniform", &fb)); break;
    default: integ.reset(new PathIntegrator(max_depth, rc->cam, sampler, b, rr, "uniform", &fb)); break;
  }
  rs->agg->reset();
  double t = 0;
  // silence the progress bar
  std::streambuf* old = std::cout.rdbuf(nullptr);
  integ->Render(*rs->scene, t);
  std::cout.rdbuf(old);
  if (rgba8) memcpy(rgba8, fb.getUCbuffer(), (size_t)W * H * 4);
  if (st) {
    st->samples = (uint64_t)W * H * spp;
    st->rays_closest = rs->agg->sumI();
    st->rays_shadow = rs->agg->sumP();
    st->seconds = t;
    st->threads = 4;
  }
  return t;
}

// Tone-map exactly as Integrator.cpp:327-344 (for the film-resolve parity test).
void ref_tonemap(const float* rgb_mean, int n, unsigned char* rgba8) {
  for (int i = 0; i < n; ++i) {
    Spectrum c;
    c[0] = rgb_mean[3 * i]; c[1] = rgb_mean[3 * i + 1]; c[2] = rgb_mean[3 * i + 2];
    float xyz[3], rgbv[3];
    c.ToXYZ(xyz);
    XYZToRGB(xyz, rgbv);
    for (int k = 0; k < 3; ++k) rgba8[4 * i + k] = (unsigned char)PBR::Clamp(255.f * GammaCorrect(rgbv[k]) + 0.5f, 0.f, 255.f);
    rgba8[4 * i + 3] = 255;
  }
}

int ref_num_procs() { return omp_get_num_procs(); }

// The reference's own .3d loader (Shape/plyRead.h): copies what plyInfo parsed into caller buffers.
int ref_ply_read(const char* path, float* verts, int cap_verts, int32_t* indices, int cap_tris, int32_t* n_verts, int32_t* n_tris) {
  plyInfo info(path);
  *n_verts = info.nVertices; *n_tris = info.nTriangles;
  if (info.nVertices > cap_verts || info.nTriangles > cap_tris) return 1;
  for (int i = 0; i < info.nVertices; ++i)
    for (int k = 0; k < 3; ++k) verts[3 * i + k] = info.vertexArray[i][k];
  memcpy(indices, info.vertexIndices, sizeof(int32_t) * 3 * (size_t)info.nTriangles);
  return 0;
}

// Surface differentials of the camera ray of (pixel, sample s): out[16] = hit flag, uv.xy, dudx, dvdx, dudy, dvdy, dpdu.xyz, dpdv.xyz, -, -
// (Triangle::Intersect + SurfaceInteraction::ComputeDifferentials, Core/Interaction.cpp:48-99)
void ref_camera_hit_differentials(void* hscene, void* hc, int spp, int W, int H, int px, int py, int64_t s, float* out) {
  auto* rs = (RefScene*)hscene;
  auto* rc = (RefCamera*)hc;
  HaltonSampler smp(spp, Bounds2i(Point2i(0, 0), Point2i(W, H)));
  Point2i pixel(px, py);
  smp.StartPixel(pixel);
  smp.SetSampleNumber(s);
  CameraSample cs = smp.GetCameraSample(pixel);
  RayDifferential r;
  rc->cam->GenerateRayDifferential(cs, &r);
  r.ScaleDifferentials(1 / std::sqrt((float)smp.samplesPerPixel));
  SurfaceInteraction isect;
  memset(out, 0, 16 * sizeof(float));
  if (!rs->scene->Intersect(r, &isect)) return;
  isect.ComputeDifferentials(r);
  out[0] = 1.f; out[1] = isect.uv[0]; out[2] = isect.uv[1];
  out[3] = isect.dudx; out[4] = isect.dvdx; out[5] = isect.dudy; out[6] = isect.dvdy;
  for (int k = 0; k < 3; ++k) { out[7 + k] = isect.dpdu[k]; out[10 + k] = isect.dpdv[k]; }
}

// ImageTexture::Evaluate known answers: in[n][8] = u, v, dudx, dvdx, dudy, dvdy, -, -  ->  out[n][4] = rgb, -
void ref_texture_kat(const pbrgpu_texture* t, const float* in, int n, float* out) {
  auto tex = makeImageTexture(*t);
  for (int i = 0; i < n; ++i) {
    const float* r = in + 8 * i;
    SurfaceInteraction si;
    si.uv = Point2f(r[0], r[1]);
    si.dudx = r[2]; si.dvdx = r[3]; si.dudy = r[4]; si.dvdy = r[5];
    Spectrum v = tex->Evaluate(si);
    out[4 * i] = v[0]; out[4 * i + 1] = v[1]; out[4 * i + 2] = v[2]; out[4 * i + 3] = 0.f;
  }
}

// The reference's own image loader (stb, compiled into the reference by Light/*.cpp) on a Radiance .hdr file:
// what InfiniteAreaLight's ctor (InfiniteAreaLight.cpp:17-18) and SkyBoxLight::loadImage (SkyBoxLight.cpp:19-22) get.
extern "C" float* stbi_loadf(char const* filename, int* x, int* y, int* channels_in_file, int desired_channels);
extern "C" void stbi_set_flip_vertically_on_load(int flag_true_if_should_flip);
extern "C" void stbi_image_free(void* retval_from_stbi_load);
int ref_load_hdr(const char* path, int flip, float* out, int cap_floats, int32_t* w, int32_t* h, int32_t* comp) {
  stbi_set_flip_vertically_on_load(flip);
  int x = 0, y = 0, c = 0;
  float* d = stbi_loadf(path, &x, &y, &c, 0);
  stbi_set_flip_vertically_on_load(0);
  if (!d) return 1;
  *w = x; *h = y; *comp = c;
  int rc = 0;
  if ((long long)x * y * c > cap_floats) rc = 2;
  else memcpy(out, d, sizeof(float) * (size_t)x * y * c);
  stbi_image_free(d);
  return rc;
}

// ---- known-answer hooks for the participating-media functions (pin oracle/restate.c and the GPU twins) ----
namespace {
// a Sampler that replays a fixed list of values (HomogeneousMedium::Sample draws two Get1D)
struct ReplaySampler : public Sampler {
  std::vector<float> v;
  size_t k = 0;
  ReplaySampler(std::initializer_list<float> l) : Sampler(1), v(l) {}
  float Get1D() override { return v[k++ % v.size()]; }
  Point2f Get2D() override { float a = Get1D(); float b = Get1D(); return Point2f(a, b); }
  std::unique_ptr<Sampler> Clone(int) override { return nullptr; }
};
}  // namespace

// in[n][8] = d.xyz, tMax, u0, u1, -, -  ->  out[n][8] = weight rgb, mi.p along d's dominant axis (or tMax), sampled flag, Tr rgb
// HomogeneousMedium::Sample / ::Tr (Media/HomogeneousMedium.cpp:10-45) on the ray (0, d, tMax)
void ref_medium_kat(const pbrgpu_medium* m, const float* in, int n, float* out) {
  HomogeneousMedium med(rgb(m->sigma_a), rgb(m->sigma_s), m->g);
  for (int i = 0; i < n; ++i) {
    const float* r = in + 8 * i;
    Ray ray(Point3f(0, 0, 0), Vector3f(r[0], r[1], r[2]), r[3], 0.f, &med);
    ReplaySampler smp{r[4], r[5]};
    MediumInteraction mi;
    Spectrum w = med.Sample(ray, smp, &mi);
    float* o = out + 8 * i;
    o[0] = w[0]; o[1] = w[1]; o[2] = w[2];
    // t itself is not stored by the reference: report mi.p = ray(t) = d * t along the dominant axis of d
    int ax = std::abs(r[0]) > std::abs(r[1]) ? (std::abs(r[0]) > std::abs(r[2]) ? 0 : 2) : (std::abs(r[1]) > std::abs(r[2]) ? 1 : 2);
    o[3] = mi.IsValid() ? mi.p[ax] : r[3];
    o[4] = mi.IsValid() ? 1.f : 0.f;
    Spectrum tr = med.Tr(ray, smp);
    o[5] = tr[0]; o[6] = tr[1]; o[7] = tr[2];
  }
}
// in[n][8] = wo.xyz, u0, u1, wi2.xyz  ->  out[n][8] = Sample_p: wi.xyz, p ; p(wo, wi2), -, -, -    (Media/Medium.cpp:7-28)
void ref_hg_kat(float g, const float* in, int n, float* out) {
  HenyeyGreenstein hg(g);
  for (int i = 0; i < n; ++i) {
    const float* r = in + 8 * i;
    Vector3f wo(r[0], r[1], r[2]), wi;
    float p = hg.Sample_p(wo, &wi, Point2f(r[3], r[4]));
    float* o = out + 8 * i;
    o[0] = wi.x; o[1] = wi.y; o[2] = wi.z; o[3] = p;
    o[4] = hg.p(wo, Vector3f(r[5], r[6], r[7]));
    o[5] = o[6] = o[7] = 0.f;
  }
}

// Light::Sample_Li / Pdf_Li of lights[light] for n reference points (same row layout as pbrgpu_light_eval)
void ref_light_eval(void* hscene, int light, const float* rows, int n, float* out) {
  auto* rs = (RefScene*)hscene;
  const Light& L = *rs->lights[light];
  for (int i = 0; i < n; ++i) {
    const float* r = rows + 16 * i;
    Interaction ref;
    ref.p = Point3f(r[0], r[1], r[2]);
    ref.pError = Vector3f(r[3], r[4], r[5]);
    ref.n = Normal3f(r[6], r[7], r[8]);
    Vector3f wi(0, 0, 0);
    float pdf = 0;
    VisibilityTester vis;
    Spectrum Li = L.Sample_Li(ref, Point2f(r[9], r[10]), &wi, &pdf, &vis);
    const float pq = L.Pdf_Li(ref, Vector3f(r[11], r[12], r[13]));
    float* o = out + 16 * i;
    o[0] = Li[0]; o[1] = Li[1]; o[2] = Li[2]; o[3] = pdf;
    o[4] = wi.x; o[5] = wi.y; o[6] = wi.z;
    o[7] = vis.P1().p.x; o[8] = vis.P1().p.y; o[9] = vis.P1().p.z;
    o[10] = vis.P1().n.x; o[11] = vis.P1().n.y; o[12] = vis.P1().n.z;
    o[13] = pq; o[14] = o[15] = 0.f;
  }
}

// Accessors for oracle/dropin_test.cpp (the C++ drop-in test lives in its own library, libdropin_test.so, which links the
// product; this library never does): the application-side objects of a harness scene / camera.
const void* ref_scene_bvh_shared(void* hscene) { return &((RefScene*)hscene)->bvh; }        // std::shared_ptr<BVHAccel>*
const void* ref_scene_lights_vector(void* hscene) { return &((RefScene*)hscene)->lights; }  // std::vector<std::shared_ptr<Light>>*
const void* ref_camera_shared(void* hc) { return &((RefCamera*)hc)->cam; }                  // std::shared_ptr<Camera>*

}  // extern "C"
