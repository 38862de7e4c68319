// tu_trace.cu - k_trace<COUNT, VOL = false, SPH = false>: the traversal kernel of the Path / Whitted / DirectLighting integrators.
#include "k_trace.cuh"
#include "launch.h"

namespace pbr {
static_assert(kTraceBlock == kTraceBlockThreads, "launch.h block size");
void launch_trace_vol(bool count, unsigned grid, cudaStream_t st, const RenderParams& P);          // tu_trace_vol.cu
void launch_trace_sph(bool count, bool vol, unsigned grid, cudaStream_t st, const RenderParams& P);  // tu_trace_sph.cu

int trace_blocks_per_sm() {
  int nb = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_trace<false, false, false>, kTraceBlock, 0) != cudaSuccess || nb < 1) {
    cudaGetLastError();
    nb = PBR_TRACE_MIN_BLOCKS;
  }
  return nb;
}

void launch_trace(bool count, bool vol, unsigned grid, cudaStream_t st, const RenderParams& P) {
  if (P.sc.nSpheres > 0) { launch_trace_sph(count, vol, grid, st, P); return; }
  if (vol) { launch_trace_vol(count, grid, st, P); return; }
  if (count) k_trace<true, false, false><<<grid, kTraceBlock, 0, st>>>(P);
  else k_trace<false, false, false><<<grid, kTraceBlock, 0, st>>>(P);
}
}  // namespace pbr
