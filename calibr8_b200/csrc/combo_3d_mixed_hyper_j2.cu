#include "combo.cuh"
C8_DEFINE_COMBO(3d_mixed_hyper_j2, 3, MECH_MIXED, HyperJ2, C8_G3D)
