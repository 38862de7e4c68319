// tu_shade_glass.cu - k_shade<GLASS, *, *> and k_shade_direct<GLASS, *> (see tu_shade.inc)
#define PBR_TU_MAT PBRGPU_MAT_GLASS
#define PBR_TU_FN(name) name##_glass
#include "tu_shade.inc"
