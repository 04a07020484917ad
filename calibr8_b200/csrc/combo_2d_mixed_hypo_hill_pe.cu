#include "combo.cuh"
C8_DEFINE_COMBO(2d_mixed_hypo_hill_pe, 2, MECH_MIXED, HypoHillPlaneStrain, 4)
