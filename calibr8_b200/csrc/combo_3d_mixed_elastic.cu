#include "combo.cuh"
C8_DEFINE_COMBO(3d_mixed_elastic, 3, MECH_MIXED, Elastic, C8_G3D)
