#include "combo.cuh"
C8_DEFINE_COMBO(2d_ps_hyper_j2, 2, MECH_PLANE_STRESS, HyperJ2PlaneStress, 4)
