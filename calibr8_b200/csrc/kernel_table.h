// One table of kernel launchers per (dimension, mechanics type, constitutive model)
// combination; each combination is compiled in its own translation unit
// (combo_*.cu) and looked up by the C-ABI context at c8_set_model time.
#pragma once
#include "args.h"
#include "adj_args.h"

namespace c8 {

struct KernelTable {
  int dim, mech, local_type;
  int nn, nb, nx, nxi, npar, group;
  bool finite;
  void (*forward_jacobian)(const FwdArgs&, cudaStream_t);
  void (*global_residual)(const FwdArgs&, cudaStream_t);
  void (*init_xi)(double* xi, long long xi_ld, int n_elems, cudaStream_t);
  void (*adjoint_jacobian)(const AdjArgs&, cudaStream_t);
  void (*adjoint_local)(const AdjArgs&, cudaStream_t);
  void (*qoi_gradient)(const AdjArgs&, cudaStream_t);
  void (*qoi_value)(const AdjArgs&, int mode, cudaStream_t);
  // virtual fields method: single-residual (plane stress) mechanics only, else nullptr
  void (*vfm_forward)(const VfmArgs&, cudaStream_t);
  void (*vfm_adjoint)(const VfmArgs&, cudaStream_t);
};

const KernelTable* find_kernel_table(int dim, int mech, int local_type);

}  // namespace c8
