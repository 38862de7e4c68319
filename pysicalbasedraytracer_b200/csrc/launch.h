// launch.h - host-side launchers of the template kernels.  Every group of instantiations lives in its own
// translation unit (tu_*.cu) so that the library builds in parallel; pbrgpu.cu only sees these declarations.
#pragma once
#include <cuda_runtime.h>

#include "device_types.cuh"

namespace pbr {

constexpr int kTraceBlockThreads = 128;   // == kTraceBlock (k_trace.cuh)
constexpr int kShadeBlockThreads = 128;   // == kShadeBlock (kcommon.cuh)
constexpr int kDirectBlockThreads = 256;  // == kBlock

// k_trace<COUNT, VOL>: resident blocks per SM (occupancy of the plain build), launch
int trace_blocks_per_sm();
// (P.sc.nSpheres > 0 selects the instantiations that know the analytic-sphere primitive)
void launch_trace(bool count, bool vol, unsigned grid, cudaStream_t st, const RenderParams& P);

// k_shade<MAT, VOL, TEX>: resident blocks per SM of that instantiation (0 = no such instantiation), launch
int shade_blocks_per_sm(int mat, bool vol, bool tex);
void launch_shade(int mat, bool vol, bool tex, unsigned grid, cudaStream_t st, const RenderParams& P, int bounce);

// k_shade_direct<MAT, TEX>
void launch_shade_direct(int mat, bool tex, unsigned grid, cudaStream_t st, const RenderParams& P, int depth);

}  // namespace pbr
