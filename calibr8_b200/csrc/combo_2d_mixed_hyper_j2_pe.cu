#include "combo.cuh"
C8_DEFINE_COMBO(2d_mixed_hyper_j2_pe, 2, MECH_MIXED, HyperJ2PlaneStrain, 4)
