#include "combo.cuh"
C8_DEFINE_COMBO(2d_mixed_small_j2, 2, MECH_MIXED, SmallJ2, 4)
