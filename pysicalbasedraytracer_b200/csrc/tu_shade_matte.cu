// tu_shade_matte.cu - k_shade<MATTE, *, *> and k_shade_direct<MATTE, *> (see tu_shade.inc)
#define PBR_TU_MAT PBRGPU_MAT_MATTE
#define PBR_TU_FN(name) name##_matte
#include "tu_shade.inc"
