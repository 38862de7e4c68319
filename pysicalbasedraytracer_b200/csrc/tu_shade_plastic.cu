// tu_shade_plastic.cu - k_shade<PLASTIC, *, *> and k_shade_direct<PLASTIC, *> (see tu_shade.inc)
#define PBR_TU_MAT PBRGPU_MAT_PLASTIC
#define PBR_TU_FN(name) name##_plastic
#include "tu_shade.inc"
