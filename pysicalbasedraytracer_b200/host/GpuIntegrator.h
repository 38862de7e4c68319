// GpuIntegrator.h - C++ drop-in for the reference renderer (G0T-cha/PysicalBasedRaytracer).
//
//   PBR::GpuPathIntegrator : public PBR::Integrator        (+ GpuVolPathIntegrator, GpuWhittedIntegrator,
//                                                            GpuDirectLightingIntegrator)
//
// has the SAME constructor signature as PBR::PathIntegrator (Integrator/PathIntegrator.h:14-18) and the
// same virtual  void Render(const Scene&, double& timeConsume)  (Integrator/Integrator.h:14), so the
// application (Main/main.cpp:399-413) switches to the GPU by changing one type name.  Render():
//   1. flattens the reference object graph  Scene -> BVHAccel -> GeometricPrimitive -> Triangle /
//      Material / AreaLight, scene.lights, the ProjectiveCamera and the HaltonSampler  into the POD
//      descriptors of include/pbrgpu.h (world-space triangle soup + material / light tables),
//   2. calls the C-ABI of libpbrgpu.so (pbrgpu_scene_create / camera_set / sampler_set / render),
//   3. writes the 8-bit image into the caller's FrameBuffer through set_uc, exactly where
//      SamplerIntegrator::Render would (Integrator/Integrator.cpp:339-343), and sets timeConsume.
// The reference's own BVH is NOT used on the GPU: only the primitive list it owns is read.
//
// The reference exposes none of the members the flattener needs (Scene::aggregate,
// BVHAccel::primitives, GeometricPrimitive::shape, Triangle::mesh, material textures, ...), so this
// header includes the reference headers with `private` / `protected` redefined to `public`
// (option 1; object layout under GCC/MSVC does not depend on access specifiers).
// => Include this header FIRST in the translation unit that uses it, before any reference header.
//
// Error convention = the reference's: no exceptions, problems go to std::cerr and Render returns
// with timeConsume = -1 (the reference has no error channel at all).
#pragma once
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <iostream>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

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
#include "Accelerator/BVHAccel.h"
#include "Camera/Camera.h"
#include "Sampler/Sampler.h"
#include "Sampler/Halton.h"
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
#include "Light/Light.h"
#include "Light/PointLight.h"
#include "Light/DiffuseLight.h"
#include "Light/InfiniteAreaLight.h"
#include "Light/SkyBoxLight.h"
#include "Media/Medium.h"
#include "Media/HomogeneousMedium.h"
#include "Integrator/Integrator.h"
#undef private
#undef protected

#include "pbrgpu.h"
#include "SobolSampler.h"
#include "AnalyticSphere.h"

namespace PBR {

namespace gpu_detail {

struct Flat {
  std::vector<float> verts, normals, uvs, tangents;  // normals / uvs / tangents: per emitted vertex, filled only if some mesh has them
  bool anyNormals = false, anyUVs = false, anyTangents = false;
  std::vector<int32_t> indices, triMaterial, triLight;
  std::vector<uint8_t> triFlags;
  std::vector<pbrgpu_material> materials;
  std::vector<pbrgpu_light> lights;
  std::vector<pbrgpu_texture> textures;          // ImageTexture objects used as Kd (texels point into texelStore)
  std::vector<std::vector<float>> texelStore;
  std::unordered_map<const void*, int> textureId;
  std::vector<pbrgpu_medium> media;              // distinct HomogeneousMedium objects met on the primitives
  std::vector<int32_t> triMediumIn, triMediumOut;  // MediumInterface::inside / ::outside per triangle (-1 = none)
  std::vector<pbrgpu_sphere> spheres;            // AnalyticSphere shapes (host/AnalyticSphere.h)
  std::vector<float> env;
  int envW = 0, envH = 0;
  std::string error;
};

inline bool constSpectrum(const std::shared_ptr<Texture<Spectrum>>& t, float out[3], std::string* err) {
  if (!t) { out[0] = out[1] = out[2] = 0.f; return true; }
  if (!dynamic_cast<const ConstantTexture<Spectrum>*>(t.get())) { *err = "only ConstantTexture is supported on the GPU path"; return false; }
  SurfaceInteraction si;
  Spectrum s = t->Evaluate(si);  // ConstantTexture::Evaluate ignores the interaction (ConstantTexture.h:14)
  out[0] = s[0]; out[1] = s[1]; out[2] = s[2];
  return true;
}
inline bool constFloat(const std::shared_ptr<Texture<float>>& t, float* out, std::string* err) {
  if (!t) { *out = 0.f; return true; }
  if (!dynamic_cast<const ConstantTexture<float>*>(t.get())) { *err = "only ConstantTexture is supported on the GPU path"; return false; }
  SurfaceInteraction si;
  *out = t->Evaluate(si);
  return true;
}

// Kd of Matte / Plastic: a ConstantTexture, or an ImageTexture<RGBSpectrum, Spectrum> with a UVMapping2D whose MIPMap level 0
// (already converted by ImageTexture::convertIn) becomes a pbrgpu_texture
// A spectrum texture of a material: a ConstantTexture goes to `rgb`, an ImageTexture to the texture table and `*texPlus1`
inline bool flattenSpectrumTexture(const std::shared_ptr<Texture<Spectrum>>& t, float rgb[3], int32_t* texPlus1, Flat* f) {
  auto* img = dynamic_cast<const ImageTexture<RGBSpectrum, Spectrum>*>(t.get());
  if (!img) return constSpectrum(t, rgb, &f->error);
  auto it = f->textureId.find(img);
  if (it == f->textureId.end()) {
    auto* map = dynamic_cast<const UVMapping2D*>(img->mapping.get());
    if (!map) { f->error = "ImageTexture with a mapping other than UVMapping2D"; return false; }
    const MIPMap<RGBSpectrum>* mm = img->mipmap;
    const auto& lvl0 = *mm->pyramid[0];
    pbrgpu_texture pt;
    std::memset(&pt, 0, sizeof(pt));
    pt.width = lvl0.uSize(); pt.height = lvl0.vSize();
    pt.do_trilinear = mm->doTrilinear ? 1 : 0;
    pt.max_anisotropy = mm->maxAnisotropy;
    pt.wrap = mm->wrapMode == ImageWrap::Clamp ? PBRGPU_WRAP_CLAMP : (mm->wrapMode == ImageWrap::Black ? PBRGPU_WRAP_BLACK : PBRGPU_WRAP_REPEAT);
    pt.su = map->su; pt.sv = map->sv; pt.du = map->du; pt.dv = map->dv;
    f->texelStore.emplace_back((size_t)pt.width * pt.height * 3);
    std::vector<float>& tx = f->texelStore.back();
    for (int v = 0; v < pt.height; ++v)
      for (int u = 0; u < pt.width; ++u) {
        const RGBSpectrum& s = lvl0(u, v);
        float* d = &tx[3 * ((size_t)v * pt.width + u)];
        d[0] = s[0]; d[1] = s[1]; d[2] = s[2];
      }
    it = f->textureId.emplace(img, (int)f->textures.size()).first;
    f->textures.push_back(pt);  // .texels is filled in once texelStore stops growing (flattenScene's caller)
  }
  *texPlus1 = it->second + 1;
  return true;
}
// A float texture of a material: a ConstantTexture goes to `*value`, an ImageTexture<float, float> to the texture table with its
// float texels replicated in R, G, B (the device reads the first channel)
inline bool flattenFloatTexture(const std::shared_ptr<Texture<float>>& t, float* value, int32_t* texPlus1, Flat* f) {
  auto* img = dynamic_cast<const ImageTexture<float, float>*>(t.get());
  if (!img) return constFloat(t, value, &f->error);
  auto it = f->textureId.find(img);
  if (it == f->textureId.end()) {
    auto* map = dynamic_cast<const UVMapping2D*>(img->mapping.get());
    if (!map) { f->error = "ImageTexture with a mapping other than UVMapping2D"; return false; }
    const MIPMap<float>* mm = img->mipmap;
    const auto& lvl0 = *mm->pyramid[0];
    pbrgpu_texture pt;
    std::memset(&pt, 0, sizeof(pt));
    pt.width = lvl0.uSize(); pt.height = lvl0.vSize();
    pt.do_trilinear = mm->doTrilinear ? 1 : 0;
    pt.max_anisotropy = mm->maxAnisotropy;
    pt.wrap = mm->wrapMode == ImageWrap::Clamp ? PBRGPU_WRAP_CLAMP : (mm->wrapMode == ImageWrap::Black ? PBRGPU_WRAP_BLACK : PBRGPU_WRAP_REPEAT);
    pt.su = map->su; pt.sv = map->sv; pt.du = map->du; pt.dv = map->dv;
    f->texelStore.emplace_back((size_t)pt.width * pt.height * 3);
    std::vector<float>& tx = f->texelStore.back();
    for (int v = 0; v < pt.height; ++v)
      for (int u = 0; u < pt.width; ++u) {
        float* d = &tx[3 * ((size_t)v * pt.width + u)];
        d[0] = d[1] = d[2] = lvl0(u, v);
      }
    it = f->textureId.emplace(img, (int)f->textures.size()).first;
    f->textures.push_back(pt);
  }
  *texPlus1 = it->second + 1;
  return true;
}
inline bool flattenKd(const std::shared_ptr<Texture<Spectrum>>& t, pbrgpu_material* o, Flat* f) {
  return flattenSpectrumTexture(t, o->kd, &o->kd_texture_plus1, f);
}

inline bool flattenMaterial(const Material* m, pbrgpu_material* o, Flat* f) {
  std::string* err = &f->error;
  std::memset(o, 0, sizeof(*o));
  if (auto* p = dynamic_cast<const MatteMaterial*>(m)) {
    o->type = PBRGPU_MAT_MATTE;
    return flattenKd(p->Kd, o, f) && constFloat(p->sigma, &o->sigma, err);
  }
  if (auto* p = dynamic_cast<const MirrorMaterial*>(m)) {
    o->type = PBRGPU_MAT_MIRROR;
    return constSpectrum(p->Kr, o->kr, err);
  }
  if (auto* p = dynamic_cast<const GlassMaterial*>(m)) {
    o->type = PBRGPU_MAT_GLASS;
    o->remap_roughness = p->remapRoughness ? 1 : 0;
    float idx = 1.f;
    bool ok = constSpectrum(p->Kr, o->kr, err) && constSpectrum(p->Kt, o->kt, err) && constFloat(p->uRoughness, &o->u_rough, err) &&
              constFloat(p->vRoughness, &o->v_rough, err) && constFloat(p->index, &idx, err);
    o->eta[0] = o->eta[1] = o->eta[2] = idx;
    return ok;
  }
  if (auto* p = dynamic_cast<const MetalMaterial*>(m)) {
    o->type = PBRGPU_MAT_METAL;
    o->remap_roughness = p->remapRoughness ? 1 : 0;
    float r = 0.f;
    bool ok = constSpectrum(p->eta, o->eta, err) && constSpectrum(p->k, o->k, err) && constFloat(p->roughness, &r, err);
    // MetalMaterial.cpp:29-32: u/vRoughness when given, else `roughness`
    if (ok && p->uRoughness) ok = constFloat(p->uRoughness, &o->u_rough, err); else o->u_rough = r;
    if (ok && p->vRoughness) ok = constFloat(p->vRoughness, &o->v_rough, err); else o->v_rough = r;
    return ok;
  }
  if (auto* p = dynamic_cast<const PlasticMaterial*>(m)) {
    o->type = PBRGPU_MAT_PLASTIC;
    o->remap_roughness = p->remapRoughness ? 1 : 0;
    bool ok = flattenKd(p->Kd, o, f) && flattenSpectrumTexture(p->Ks, o->kr, &o->ks_texture_plus1, f) &&
              flattenFloatTexture(p->roughness, &o->u_rough, &o->rough_texture_plus1, f);
    if (ok && o->rough_texture_plus1 > 0 && o->remap_roughness) { *err = "Plastic: a roughness ImageTexture with remapRoughness is not supported on the GPU path"; ok = false; }
    o->v_rough = o->u_rough;
    return ok;
  }
  *err = "unsupported Material subclass (GPU path: Matte, Mirror, Glass, Metal, Plastic)";
  return false;
}

// Medium* -> index into Flat::media (-1 for nullptr); only HomogeneousMedium has a GPU twin
inline bool flattenMedium(const Medium* m, std::unordered_map<const Medium*, int>* ids, Flat* f, int* out) {
  *out = -1;
  if (!m) return true;
  auto it = ids->find(m);
  if (it == ids->end()) {
    auto* h = dynamic_cast<const HomogeneousMedium*>(m);
    if (!h) { f->error = "unsupported Medium subclass (GPU path: HomogeneousMedium)"; return false; }
    pbrgpu_medium o;
    std::memset(&o, 0, sizeof(o));
    o.type = PBRGPU_MEDIUM_HOMOGENEOUS;
    o.g = h->g;
    for (int c = 0; c < 3; ++c) { o.sigma_a[c] = h->sigma_a[c]; o.sigma_s[c] = h->sigma_s[c]; }
    it = ids->emplace(m, (int)f->media.size()).first;
    f->media.push_back(o);
  }
  *out = it->second;
  return true;
}

inline void copyMatrix(const Matrix4x4& m, float out[16]) {
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) out[i * 4 + j] = m.m[i][j];
}

inline bool flattenScene(const Scene& scene, Flat* f) {
  auto* bvh = dynamic_cast<const BVHAccel*>(scene.aggregate.get());
  if (!bvh) { f->error = "Scene::aggregate is not a BVHAccel"; return false; }
  const auto& prims = bvh->primitives;
  const size_t n = prims.size();
  f->verts.reserve(n * 9);
  f->indices.reserve(n * 3);
  f->triMaterial.reserve(n);
  std::unordered_map<const Medium*, int> mediumId;
  std::unordered_map<const Material*, int> matId;
  std::unordered_map<const Light*, int> lightId;
  for (size_t i = 0; i < scene.lights.size(); ++i) lightId[scene.lights[i].get()] = (int)i;
  f->lights.resize(scene.lights.size());
  std::vector<int> lightTri(scene.lights.size(), -1), lightSphere(scene.lights.size(), -1);
  auto materialOf = [&](const GeometricPrimitive* gp, int* out) {
    if (!gp->material) { *out = -1; return true; }  // medium boundary: rendered by GpuVolPathIntegrator only
    auto it = matId.find(gp->material.get());
    if (it == matId.end()) {
      pbrgpu_material m;
      if (!flattenMaterial(gp->material.get(), &m, f)) return false;
      it = matId.emplace(gp->material.get(), (int)f->materials.size()).first;
      f->materials.push_back(m);
    }
    *out = it->second;
    return true;
  };
  for (size_t pi = 0; pi < n; ++pi) {
    auto* gp = dynamic_cast<const GeometricPrimitive*>(prims[pi].get());
    if (!gp) { f->error = "primitive is not a GeometricPrimitive"; return false; }
    if (auto* sph = dynamic_cast<const AnalyticSphere*>(gp->shape.get())) {
      // analytic sphere (the reference's own Sphere is a stub: Shape/Sphere.cpp:8-37); ObjectToWorld must be a translation
      const Matrix4x4& m = sph->ObjectToWorld->m;
      for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c)
          if (m.m[r][c] != (r == c ? 1.f : 0.f)) { f->error = "AnalyticSphere: ObjectToWorld must be a pure translation"; return false; }
      if (gp->mediumInterface.inside || gp->mediumInterface.outside) { f->error = "AnalyticSphere with a MediumInterface is not supported"; return false; }
      pbrgpu_sphere o;
      std::memset(&o, 0, sizeof(o));
      o.center[0] = m.m[0][3]; o.center[1] = m.m[1][3]; o.center[2] = m.m[2][3];
      o.radius = sph->radius;
      o.reverse_orientation = sph->reverseOrientation ? 1 : 0;
      o.light = -1;
      if (!gp->material) { f->error = "AnalyticSphere without a material is not supported"; return false; }
      if (!materialOf(gp, &o.material)) return false;
      if (gp->areaLight) {
        auto li = lightId.find(gp->areaLight.get());
        if (li == lightId.end()) { f->error = "area light of a primitive is not in scene.lights"; return false; }
        o.light = li->second;
        lightSphere[li->second] = (int)f->spheres.size();
      }
      f->spheres.push_back(o);
      continue;
    }
    auto* tri = dynamic_cast<const Triangle*>(gp->shape.get());
    if (!tri) { f->error = "shape is neither a Triangle nor an AnalyticSphere (the reference's Sphere is a stub, SURVEY Q4)"; return false; }
    const size_t i = f->triMaterial.size();  // triangle index (analytic spheres are numbered separately)
    f->triMaterial.push_back(-1); f->triLight.push_back(-1); f->triFlags.push_back(0);
    f->triMediumIn.push_back(-1); f->triMediumOut.push_back(-1);
    if (!flattenMedium(gp->mediumInterface.inside, &mediumId, f, &f->triMediumIn[i]) ||
        !flattenMedium(gp->mediumInterface.outside, &mediumId, f, &f->triMediumOut[i]))
      return false;
    if ((tri->mesh->n && !f->anyNormals) || (tri->mesh->uv && !f->anyUVs)) {
      // first mesh with shading normals / uvs: materialise the optional arrays (zero for everything before it)
      if (tri->mesh->n && !f->anyNormals) { f->anyNormals = true; f->normals.assign(n * 9, 0.f); }
      if (tri->mesh->uv && !f->anyUVs) { f->anyUVs = true; f->uvs.assign(n * 6, 0.f); }
    }
    if (tri->mesh->s && !f->anyTangents) { f->anyTangents = true; f->tangents.assign(n * 9, 0.f); }
    for (int k = 0; k < 3; ++k) {
      const Point3f& p = tri->mesh->p[tri->v[k]];  // already world space (TriangleMesh ctor, Triangle.cpp:27-28)
      f->verts.push_back(p.x); f->verts.push_back(p.y); f->verts.push_back(p.z);
      f->indices.push_back((int32_t)(3 * i + k));
      if (tri->mesh->n) {
        const Normal3f& nn = tri->mesh->n[tri->v[k]];
        float* d = &f->normals[9 * i + 3 * k];
        d[0] = nn.x; d[1] = nn.y; d[2] = nn.z;
      }
      if (tri->mesh->s) {
        const Vector3f& sv = tri->mesh->s[tri->v[k]];
        float* d = &f->tangents[9 * i + 3 * k];
        d[0] = sv.x; d[1] = sv.y; d[2] = sv.z;
      }
      if (tri->mesh->uv) {
        const Point2f& t = tri->mesh->uv[tri->v[k]];
        f->uvs[6 * i + 2 * k] = t.x; f->uvs[6 * i + 2 * k + 1] = t.y;
      }
    }
    f->triFlags[i] = (uint8_t)(((tri->reverseOrientation ^ tri->transformSwapsHandedness) ? 1 : 0) | (tri->reverseOrientation ? 2 : 0) |
                               (tri->mesh->n ? 4 : 0) | (tri->mesh->uv ? 8 : 0) | (tri->mesh->s ? 16 : 0));
    if (!materialOf(gp, &f->triMaterial[i])) return false;
    if (gp->areaLight) {
      auto li = lightId.find(gp->areaLight.get());
      if (li == lightId.end()) { f->error = "area light of a primitive is not in scene.lights"; return false; }
      f->triLight[i] = li->second;
      lightTri[li->second] = (int)i;
    }
  }
  for (size_t i = 0; i < scene.lights.size(); ++i) {
    pbrgpu_light& o = f->lights[i];
    std::memset(&o, 0, sizeof(o));
    const Light* l = scene.lights[i].get();
    o.n_samples = l->nSamples;
    o.tri = -1;
    if (auto* p = dynamic_cast<const PointLight*>(l)) {
      o.type = PBRGPU_LIGHT_POINT;
      o.pos[0] = p->pLight.x; o.pos[1] = p->pLight.y; o.pos[2] = p->pLight.z;
      o.L[0] = p->I[0]; o.L[1] = p->I[1]; o.L[2] = p->I[2];
    } else if (auto* p = dynamic_cast<const DiffuseAreaLight*>(l)) {
      o.type = lightSphere[i] >= 0 ? PBRGPU_LIGHT_AREA_SPHERE : PBRGPU_LIGHT_AREA_TRI;
      o.tri = lightSphere[i] >= 0 ? lightSphere[i] : lightTri[i];
      if (o.tri < 0) { f->error = "DiffuseAreaLight whose shape is not a scene primitive"; return false; }
      o.two_sided = p->twoSided ? 1 : 0;
      o.L[0] = p->Lemit[0]; o.L[1] = p->Lemit[1]; o.L[2] = p->Lemit[2];
    } else if (auto* p = dynamic_cast<const InfiniteAreaLight*>(l)) {
      o.type = PBRGPU_LIGHT_INFINITE;
      o.L[0] = o.L[1] = o.L[2] = 1.f;  // the MIPMap texels are already multiplied by L (InfiniteAreaLight.cpp:20-31)
      copyMatrix(p->LightToWorld.m, o.light_to_world);
      copyMatrix(p->WorldToLight.m, o.world_to_light);
      const auto& lvl0 = *p->Lmap->pyramid[0];
      f->envW = lvl0.uSize(); f->envH = lvl0.vSize();
      f->env.resize((size_t)f->envW * f->envH * 3);
      for (int v = 0; v < f->envH; ++v)
        for (int u = 0; u < f->envW; ++u) {
          const RGBSpectrum& t = lvl0(u, v);
          float* d = &f->env[3 * ((size_t)v * f->envW + u)];
          d[0] = t[0]; d[1] = t[1]; d[2] = t[2];
        }
    } else if (auto* p = dynamic_cast<const SkyBoxLight*>(l)) {
      o.type = PBRGPU_LIGHT_SKYBOX;
      o.world_radius = p->worldRadius;
      if (!p->data || p->nrComponents < 3) { f->error = "SkyBoxLight without image data"; return false; }
      f->envW = p->imageWidth; f->envH = p->imageHeight;
      f->env.resize((size_t)f->envW * f->envH * 3);
      for (size_t k = 0; k < (size_t)f->envW * f->envH; ++k)
        for (int c = 0; c < 3; ++c) f->env[3 * k + c] = p->data[k * p->nrComponents + c];
    } else {
      f->error = "unsupported Light subclass (GPU path: Point, DiffuseArea, InfiniteArea, SkyBox)";
      return false;
    }
  }
  return true;
}

}  // namespace gpu_detail

// Shared Render() of the GPU integrators: flatten -> pbrgpu_* -> FrameBuffer.
class GpuSamplerIntegrator : public Integrator {
 public:
  GpuSamplerIntegrator(std::shared_ptr<const Camera> camera, std::shared_ptr<Sampler> sampler, const Bounds2i& pixelBounds,
                       FrameBuffer* framebuffer, int device = 0)
      : camera(camera), sampler(sampler), pixelBounds(pixelBounds), m_FrameBuffer(framebuffer), device(device) {}
  ~GpuSamplerIntegrator() override {
    if (gscene) pbrgpu_scene_destroy(gscene);
    if (ctx) pbrgpu_shutdown(ctx);
  }
  void Render(const Scene& scene, double& timeConsume) override {
    auto t0 = std::chrono::steady_clock::now();
    timeConsume = -1;
    if (!ctx) {  // one GPU (pbrgpu_init) or a group of GPUs driven by this process (pbrgpu_init_multi)
      if (devices.size() > 1 ? pbrgpu_init_multi((int)devices.size(), devices.data(), &ctx) : pbrgpu_init(devices.empty() ? device : devices[0], &ctx))
        return fail("pbrgpu_init");
    }
    if (!gscene || flattenedScene != &scene) {  // flatten once per Scene object
      if (gscene) { pbrgpu_scene_destroy(gscene); gscene = nullptr; }
      gpu_detail::Flat f;
      if (!gpu_detail::flattenScene(scene, &f)) { std::cerr << "GpuIntegrator: " << f.error << std::endl; return; }
      pbrgpu_scene_desc d;
      std::memset(&d, 0, sizeof(d));
      d.n_verts = (int32_t)(f.verts.size() / 3);
      d.n_tris = (int32_t)(f.indices.size() / 3);
      d.verts = f.verts.data(); d.indices = f.indices.data();
      d.tri_material = f.triMaterial.data(); d.tri_light = f.triLight.data(); d.tri_flags = f.triFlags.data();
      d.normals = f.anyNormals ? f.normals.data() : nullptr;
      d.uvs = f.anyUVs ? f.uvs.data() : nullptr;
      d.tangents = f.anyTangents ? f.tangents.data() : nullptr;
      d.n_materials = (int32_t)f.materials.size(); d.materials = f.materials.data();
      d.n_lights = (int32_t)f.lights.size(); d.lights = f.lights.data();
      d.env_width = f.envW; d.env_height = f.envH; d.env_texels = f.env.empty() ? nullptr : f.env.data();
      for (size_t k = 0; k < f.textures.size(); ++k) f.textures[k].texels = f.texelStore[k].data();
      d.n_textures = (int32_t)f.textures.size();
      d.textures = f.textures.empty() ? nullptr : f.textures.data();
      if (!f.media.empty()) {
        d.n_media = (int32_t)f.media.size(); d.media = f.media.data();
        d.tri_medium_inside = f.triMediumIn.data(); d.tri_medium_outside = f.triMediumOut.data();
      }
      d.n_spheres = (int32_t)f.spheres.size();
      d.spheres = f.spheres.empty() ? nullptr : f.spheres.data();
      if (pbrgpu_scene_create(ctx, &d, &gscene)) return fail("pbrgpu_scene_create");
      flattenedScene = &scene;
    }
    // camera (Camera/Camera.h:34-54)
    auto* pc = dynamic_cast<const ProjectiveCamera*>(camera.get());
    if (!pc) { std::cerr << "GpuIntegrator: camera is not a ProjectiveCamera" << std::endl; return; }
    const int W = pixelBounds.pMax.x, H = pixelBounds.pMax.y;
    pbrgpu_camera_desc cd;
    std::memset(&cd, 0, sizeof(cd));
    // an OrthographicCamera keeps its ray direction fixed at (0,0,1); tell the two apart by the projection's w row
    cd.type = (pc->CameraToScreen.m.m[3][2] != 0.f) ? PBRGPU_CAMERA_PERSPECTIVE : PBRGPU_CAMERA_ORTHOGRAPHIC;
    cd.width = W; cd.height = H;
    cd.lens_radius = pc->lensRadius; cd.focal_distance = pc->focalDistance;
    gpu_detail::copyMatrix(pc->RasterToCamera.m, cd.raster_to_camera);
    gpu_detail::copyMatrix(pc->CameraToWorld.m, cd.camera_to_world);
    // Camera::medium is deliberately NOT forwarded: the reference loses it before the integrator sees the ray
    // (Perspective.cpp:44-80 never sets it; Orthographic.cpp:59-60 + Transform.h:169), so cd.medium_plus1 stays 0.
    if (pbrgpu_camera_set(gscene, &cd)) return fail("pbrgpu_camera_set");
    // sampler (Sampler/Halton.cpp:30-58): the reference passes the same bounds to sampler and integrator
    auto* hs = dynamic_cast<const HaltonSampler*>(sampler.get());
    auto* ss = dynamic_cast<const SobolSampler*>(sampler.get());
    if (!hs && !ss) { std::cerr << "GpuIntegrator: sampler is neither a HaltonSampler nor a SobolSampler" << std::endl; return; }
    pbrgpu_sampler_desc sd;
    std::memset(&sd, 0, sizeof(sd));
    sd.samples_per_pixel = sampler->samplesPerPixel;
    sd.sample_bounds[0] = pixelBounds.pMin.x; sd.sample_bounds[1] = pixelBounds.pMin.y;
    sd.sample_bounds[2] = pixelBounds.pMax.x; sd.sample_bounds[3] = pixelBounds.pMax.y;
    if (hs) {
      sd.type = PBRGPU_SAMPLER_HALTON;
      sd.sample_at_pixel_center = hs->sampleAtPixelCenter ? 1 : 0;
      int32_t p7[7];
      pbrgpu_halton_params(sd.sample_bounds, p7);
      if (p7[0] != hs->baseScales[0] || p7[1] != hs->baseScales[1] || p7[4] != hs->sampleStride) {
        std::cerr << "GpuIntegrator: HaltonSampler was built for other bounds than the integrator's pixelBounds" << std::endl;
        return;
      }
      sd.perms = HaltonSampler::radicalInversePermutations.data();
      sd.n_perms = (int64_t)HaltonSampler::radicalInversePermutations.size();
    } else {
      sd.type = PBRGPU_SAMPLER_SOBOL;
      sd.sample_bounds[0] = ss->sampleBounds.pMin.x; sd.sample_bounds[1] = ss->sampleBounds.pMin.y;
      sd.sample_bounds[2] = ss->sampleBounds.pMax.x; sd.sample_bounds[3] = ss->sampleBounds.pMax.y;
      if (ss->log2Resolution < 1) { std::cerr << "GpuIntegrator: SobolSampler over a 1x1 image" << std::endl; return; }
      sd.sobol_matrices32 = SobolMatrices32;
      sd.sobol_n_dims = NumSobolDimensions;
      sd.sobol_vdc = VdCSobolMatrices[ss->log2Resolution - 1];
      sd.sobol_vdc_inv = VdCSobolMatricesInv[ss->log2Resolution - 1];
    }
    if (pbrgpu_sampler_set(gscene, &sd)) return fail("pbrgpu_sampler_set");
    pbrgpu_render_desc rd;
    std::memset(&rd, 0, sizeof(rd));
    fillRenderDesc(&rd);
    rd.x0 = pixelBounds.pMin.x; rd.y0 = pixelBounds.pMin.y; rd.x1 = pixelBounds.pMax.x; rd.y1 = pixelBounds.pMax.y;
    rd.s0 = 0; rd.s1 = sampler->samplesPerPixel;
    const int rw = rd.x1 - rd.x0, rh = rd.y1 - rd.y0;
    rgbSum.assign((size_t)rw * rh * 3, 0.f);
    std::vector<unsigned char> rgba((size_t)rw * rh * 4);
    if (pbrgpu_render(gscene, &rd, rgbSum.data(), rgba.data(), &stats)) return fail("pbrgpu_render");
    if (m_FrameBuffer) {
      // pbrgpu_render returns rows already flipped (row 0 = bottom), as set_uc(i, pMax.y - j - 1, ..) stores them
      for (int r = 0; r < rh; ++r)
        for (int x = 0; x < rw; ++x) {
          const unsigned char* px = &rgba[4 * ((size_t)r * rw + x)];
          const int fbRow = pixelBounds.pMax.y - 1 - (rd.y0 + (rh - 1 - r));
          for (int c = 0; c < 4; ++c) m_FrameBuffer->set_uc(rd.x0 + x, fbRow, c, px[c]);
        }
    }
    timeConsume = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    IntegratorRenderTime = (float)timeConsume;
  }
  pbrgpu_stats stats{};
  std::vector<float> rgbSum;  // linear per-pixel SUM of Li over the samples (the reference's colObj before /= spp)

  // Multi-GPU: call before the first Render().  The scene is replicated on these devices, every Render()
  // shards the samples of each pixel across them and sums the per-GPU films with one ncclReduce; the application still makes
  // the single Integrator::Render call of Main/main.cpp:413.
  void setDevices(const std::vector<int>& ids) { devices = ids; }

 protected:
  virtual void fillRenderDesc(pbrgpu_render_desc* rd) const = 0;
  void fail(const char* what) { std::cerr << "GpuIntegrator: " << what << " failed: " << pbrgpu_last_error() << std::endl; }
  std::shared_ptr<const Camera> camera;
  std::shared_ptr<Sampler> sampler;
  const Bounds2i pixelBounds;
  FrameBuffer* m_FrameBuffer;
  int device;
  std::vector<int> devices;  // setDevices(): more than one -> pbrgpu_init_multi
  pbrgpu_ctx* ctx = nullptr;
  pbrgpu_scene* gscene = nullptr;
  const Scene* flattenedScene = nullptr;
};

// Same constructor as PathIntegrator (Integrator/PathIntegrator.h:14-18).
class GpuPathIntegrator : public GpuSamplerIntegrator {
 public:
  GpuPathIntegrator(int maxDepth, std::shared_ptr<const Camera> camera, std::shared_ptr<Sampler> sampler, const Bounds2i& pixelBounds,
                    float rrThreshold = 1, const std::string& lightSampleStrategy = "spatial", FrameBuffer* framebuffer = nullptr)
      : GpuSamplerIntegrator(camera, sampler, pixelBounds, framebuffer), maxDepth(maxDepth), rrThreshold(rrThreshold),
        lightSampleStrategy(lightSampleStrategy) {}

 protected:
  void fillRenderDesc(pbrgpu_render_desc* rd) const override {
    rd->integrator = PBRGPU_INTEGRATOR_PATH;
    rd->max_depth = maxDepth;
    rd->rr_threshold = rrThreshold;
    // LightDistrib.cpp:8-23: "power" -> PowerLightDistribution, everything else ("uniform", "spatial") -> uniform
    rd->light_strategy = lightSampleStrategy == "power" ? PBRGPU_LIGHTSTRATEGY_POWER : PBRGPU_LIGHTSTRATEGY_UNIFORM;
  }
  const int maxDepth;
  const float rrThreshold;
  const std::string lightSampleStrategy;
};

// Same constructor as VolPathIntegrator (Integrator/VolPathIntegrator.h:11-19) - the integrator Main/main.cpp:399 builds.
class GpuVolPathIntegrator : public GpuPathIntegrator {
 public:
  using GpuPathIntegrator::GpuPathIntegrator;

 protected:
  void fillRenderDesc(pbrgpu_render_desc* rd) const override {
    GpuPathIntegrator::fillRenderDesc(rd);
    rd->integrator = PBRGPU_INTEGRATOR_VOLPATH;
  }
};

// Same constructor as WhittedIntegrator (Integrator/WhittedIntegrator.h:8-11).
class GpuWhittedIntegrator : public GpuSamplerIntegrator {
 public:
  GpuWhittedIntegrator(int maxDepth, std::shared_ptr<const Camera> camera, std::shared_ptr<Sampler> sampler, const Bounds2i& pixelBounds,
                       FrameBuffer* framebuffer = nullptr)
      : GpuSamplerIntegrator(camera, sampler, pixelBounds, framebuffer), maxDepth(maxDepth) {}

 protected:
  void fillRenderDesc(pbrgpu_render_desc* rd) const override {
    rd->integrator = PBRGPU_INTEGRATOR_WHITTED;
    rd->max_depth = maxDepth;
  }
  const int maxDepth;
};

// Same constructor as DirectLightingIntegrator (Integrator/DirectLightingIntegrator.h:10-17); `strategy` is the
// reference's own LightStrategy enum when that header is included, else an int (0 = UniformSampleAll, 1 = UniformSampleOne).
class GpuDirectLightingIntegrator : public GpuSamplerIntegrator {
 public:
  template <typename Strategy>
  GpuDirectLightingIntegrator(Strategy strategy, int maxDepth, std::shared_ptr<const Camera> camera, std::shared_ptr<Sampler> sampler,
                              const Bounds2i& pixelBounds, FrameBuffer* framebuffer = nullptr)
      : GpuSamplerIntegrator(camera, sampler, pixelBounds, framebuffer), sampleAll((int)strategy == 0), maxDepth(maxDepth) {}

 protected:
  void fillRenderDesc(pbrgpu_render_desc* rd) const override {
    rd->integrator = sampleAll ? PBRGPU_INTEGRATOR_DIRECT_ALL : PBRGPU_INTEGRATOR_DIRECT_ONE;
    rd->max_depth = maxDepth;
  }
  const bool sampleAll;
  const int maxDepth;
};

}  // namespace PBR
