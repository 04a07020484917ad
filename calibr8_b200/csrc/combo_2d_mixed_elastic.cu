#include "combo.cuh"
C8_DEFINE_COMBO(2d_mixed_elastic, 2, MECH_MIXED, Elastic, 4)
