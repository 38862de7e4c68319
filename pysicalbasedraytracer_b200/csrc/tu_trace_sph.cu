// tu_trace_sph.cu - k_trace<COUNT, VOL, SPH = true>: scenes that contain analytic spheres (pbrgpu_sphere, sphere.cuh).
#include "k_trace.cuh"
#include "launch.h"

namespace pbr {
void launch_trace_sph(bool count, bool vol, unsigned grid, cudaStream_t st, const RenderParams& P) {
  if (vol) {
    if (count) k_trace<true, true, true><<<grid, kTraceBlock, 0, st>>>(P);
    else k_trace<false, true, true><<<grid, kTraceBlock, 0, st>>>(P);
  } else {
    if (count) k_trace<true, false, true><<<grid, kTraceBlock, 0, st>>>(P);
    else k_trace<false, false, true><<<grid, kTraceBlock, 0, st>>>(P);
  }
}
}  // namespace pbr
