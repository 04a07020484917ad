#include "combo.cuh"
C8_DEFINE_COMBO(3d_mixed_small_j2, 3, MECH_MIXED, SmallJ2, C8_G3D)
