#include "combo.cuh"
C8_DEFINE_COMBO(3d_mixed_small_hill, 3, MECH_MIXED, SmallHill, C8_G3D)
