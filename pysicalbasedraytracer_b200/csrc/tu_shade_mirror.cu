// tu_shade_mirror.cu - k_shade<MIRROR, *, *> and k_shade_direct<MIRROR, *> (see tu_shade.inc)
#define PBR_TU_MAT PBRGPU_MAT_MIRROR
#define PBR_TU_FN(name) name##_mirror
#include "tu_shade.inc"
