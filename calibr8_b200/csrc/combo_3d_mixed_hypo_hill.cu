#include "combo.cuh"
C8_DEFINE_COMBO(3d_mixed_hypo_hill, 3, MECH_MIXED, HypoHill, C8_G3D)
