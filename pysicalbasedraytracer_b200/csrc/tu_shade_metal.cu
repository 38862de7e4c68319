// tu_shade_metal.cu - k_shade<METAL, *, *> and k_shade_direct<METAL, *> (see tu_shade.inc)
#define PBR_TU_MAT PBRGPU_MAT_METAL
#define PBR_TU_FN(name) name##_metal
#include "tu_shade.inc"
