#include "combo.cuh"
C8_DEFINE_COMBO(2d_ps_small_hill, 2, MECH_PLANE_STRESS, SmallHillPlaneStress, 4)
