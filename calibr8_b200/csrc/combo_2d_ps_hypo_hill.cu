#include "combo.cuh"
C8_DEFINE_COMBO(2d_ps_hypo_hill, 2, MECH_PLANE_STRESS, HypoHillPlaneStress, 4)
