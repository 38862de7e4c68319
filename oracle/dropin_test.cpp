// TEST INFRASTRUCTURE ONLY - never linked into the product, and the only oracle-side object that links the product.
//
// dropin_test.cpp -> oracle/_ref/libdropin_test.so (built by oracle/build_ref.sh): the C++ drop-in test.  The reference's
// OWN object graph (Scene / BVHAccel / GeometricPrimitive / Triangle / Material / Light / PerspectiveCamera / HaltonSampler),
// created by the harness in libpbr_ref.so, is rendered through PBR::GpuPathIntegrator::Render, i.e. through the flattener of
// pysicalbasedraytracer_b200/host/GpuIntegrator.h and the C-ABI of libpbrgpu.so.  Needs a GPU.
// Kept out of libpbr_ref.so so that the reference arm of bench.py (`--impl reference`) loads no product code at all.
// The C++ drop-in under test (flattener + GpuPathIntegrator); must precede every reference header.
#include "GpuIntegrator.h"

#include "Accelerator/BVHAccel.h"
#include "Camera/Camera.h"
#include "Core/FrameBuffer.h"
#include "Core/Scene.h"
#include "Integrator/DirectLightingIntegrator.h"
#include "Light/Light.h"
#include "Sampler/Halton.h"

using namespace PBR;

extern "C" {
// libpbr_ref.so (oracle/ref_harness.cpp): the application-side objects behind the harness handles
const void* ref_scene_bvh_shared(void* hscene);
const void* ref_scene_lights_vector(void* hscene);
const void* ref_camera_shared(void* hc);

int ref_gpu_dropin_render(void* hscene, void* hc, int integrator, int spp, int W, int H, int max_depth, float rr, const char* strategy,
                          unsigned char* rgba8, float* rgb_sum, double* seconds, int n_devices) {
  const auto& bvh = *(const std::shared_ptr<BVHAccel>*)ref_scene_bvh_shared(hscene);
  const auto& lights = *(const std::vector<std::shared_ptr<Light>>*)ref_scene_lights_vector(hscene);
  const auto& cam = *(const std::shared_ptr<Camera>*)ref_camera_shared(hc);
  Scene raw(bvh, lights);  // the app's Scene: aggregate = BVHAccel (no counting wrapper)
  Bounds2i b(Point2i(0, 0), Point2i(W, H));
  std::shared_ptr<Sampler> sampler = std::make_shared<HaltonSampler>(spp, b);
  FrameBuffer fb;
  fb.InitBuffer(W, H, 4);
  std::unique_ptr<GpuSamplerIntegrator> integ;
  switch (integrator) {
    case PBRGPU_INTEGRATOR_PATH: integ.reset(new GpuPathIntegrator(max_depth, cam, sampler, b, rr, strategy, &fb)); break;
    case PBRGPU_INTEGRATOR_VOLPATH: integ.reset(new GpuVolPathIntegrator(max_depth, cam, sampler, b, rr, strategy, &fb)); break;
    case PBRGPU_INTEGRATOR_WHITTED: integ.reset(new GpuWhittedIntegrator(max_depth, cam, sampler, b, &fb)); break;
    case PBRGPU_INTEGRATOR_DIRECT_ONE:
      integ.reset(new GpuDirectLightingIntegrator(LightStrategy::UniformSampleOne, max_depth, cam, sampler, b, &fb)); break;
    case PBRGPU_INTEGRATOR_DIRECT_ALL:
      integ.reset(new GpuDirectLightingIntegrator(LightStrategy::UniformSampleAll, max_depth, cam, sampler, b, &fb)); break;
    default: return 2;
  }
  // n_devices > 1: the same Render() call on a pbrgpu_init_multi context (devices 0..n-1): sample range sharded, films
  // reduced with NCCL, resolve on device 0
  if (n_devices > 1) { std::vector<int> ids; for (int i = 0; i < n_devices; ++i) ids.push_back(i); integ->setDevices(ids); }
  double t = -1;
  integ->Render(raw, t);
  if (seconds) *seconds = t;
  if (t < 0) return 1;
  if (rgba8) memcpy(rgba8, fb.getUCbuffer(), (size_t)W * H * 4);
  if (rgb_sum) memcpy(rgb_sum, integ->rgbSum.data(), sizeof(float) * (size_t)W * H * 3);
  return 0;
}

}  // extern "C"
