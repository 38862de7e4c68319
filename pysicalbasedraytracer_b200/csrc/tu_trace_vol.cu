// tu_trace_vol.cu - k_trace<COUNT, VOL = true>: VolPath traversal (shadow rays march through medium boundaries).
#include "k_trace.cuh"
#include "launch.h"

namespace pbr {
void launch_trace_vol(bool count, unsigned grid, cudaStream_t st, const RenderParams& P) {
  if (count) k_trace<true, true, false><<<grid, kTraceBlock, 0, st>>>(P);
  else k_trace<false, true, false><<<grid, kTraceBlock, 0, st>>>(P);
}
}  // namespace pbr
