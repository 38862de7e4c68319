// tu_shade_dispatch.cu - material id -> the per-material launchers of tu_shade_<material>.cu
#include "launch.h"

namespace pbr {
#define PBR_DECL(n)                                                                                              \
  int shade_blocks_per_sm_##n(bool vol, bool tex);                                                               \
  void launch_shade_##n(bool vol, bool tex, unsigned grid, cudaStream_t st, const RenderParams& P, int bounce); \
  void launch_shade_direct_##n(bool tex, unsigned grid, cudaStream_t st, const RenderParams& P, int depth);
PBR_DECL(matte) PBR_DECL(mirror) PBR_DECL(glass) PBR_DECL(metal) PBR_DECL(plastic)
#undef PBR_DECL

int shade_blocks_per_sm(int mat, bool vol, bool tex) {
  switch (mat) {
    case PBRGPU_MAT_MATTE: return shade_blocks_per_sm_matte(vol, tex);
    case PBRGPU_MAT_MIRROR: return shade_blocks_per_sm_mirror(vol, tex);
    case PBRGPU_MAT_GLASS: return shade_blocks_per_sm_glass(vol, tex);
    case PBRGPU_MAT_METAL: return shade_blocks_per_sm_metal(vol, tex);
    case PBRGPU_MAT_PLASTIC: return shade_blocks_per_sm_plastic(vol, tex);
  }
  return 0;
}
void launch_shade(int mat, bool vol, bool tex, unsigned grid, cudaStream_t st, const RenderParams& P, int bounce) {
  switch (mat) {
    case PBRGPU_MAT_MATTE: launch_shade_matte(vol, tex, grid, st, P, bounce); break;
    case PBRGPU_MAT_MIRROR: launch_shade_mirror(vol, tex, grid, st, P, bounce); break;
    case PBRGPU_MAT_GLASS: launch_shade_glass(vol, tex, grid, st, P, bounce); break;
    case PBRGPU_MAT_METAL: launch_shade_metal(vol, tex, grid, st, P, bounce); break;
    case PBRGPU_MAT_PLASTIC: launch_shade_plastic(vol, tex, grid, st, P, bounce); break;
  }
}
void launch_shade_direct(int mat, bool tex, unsigned grid, cudaStream_t st, const RenderParams& P, int depth) {
  switch (mat) {
    case PBRGPU_MAT_MATTE: launch_shade_direct_matte(tex, grid, st, P, depth); break;
    case PBRGPU_MAT_MIRROR: launch_shade_direct_mirror(tex, grid, st, P, depth); break;
    case PBRGPU_MAT_GLASS: launch_shade_direct_glass(tex, grid, st, P, depth); break;
    case PBRGPU_MAT_METAL: launch_shade_direct_metal(tex, grid, st, P, depth); break;
    case PBRGPU_MAT_PLASTIC: launch_shade_direct_plastic(tex, grid, st, P, depth); break;
  }
}
}  // namespace pbr
