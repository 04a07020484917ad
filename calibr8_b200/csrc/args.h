// Plain-old-data argument blocks shared by the kernels, their launchers and the
// C-ABI context.  All pointers are device pointers.
#pragma once
#include <cuda_runtime.h>

namespace c8 {

// HBM layout (see DESIGN.md "Data layout"):
//   conn    [n_elems][NN] int32          coords [n_nodes][DIM] f64
//   x       [n_nodes][NB] f64 (node-interleaved u.., p)
//   xi      [NXI][xi_ld] f64  structure-of-arrays, xi_ld >= n_elems
//   eoff    [n_elems][NN*NN] int32  BSR block index of (node_a, node_b)
//   vals    [nnzb][NB][NB] f64 BSR values, b [n_nodes][NB]
struct MeshArgs {
  int n_elems;
  int n_nodes;
  int n_row_nodes;        // rows of nodes >= n_row_nodes are not scattered (ghost nodes of a partition)
  const int* conn;
  const double* coords;
  const int* elem_es;     // [n_elems] element-set id, or nullptr (all 0)
  const int* eoff;
  const int* gptr;        // [nnzb+1] gather plan of the two-phase assembly (inverse of eoff)
  const int* gsrc;        // [n_elems*NN*NN] positions in eoff, sorted by block then element
  int nnzb;
  int n_row_blocks;       // blocks in the rows of nodes < n_row_nodes (= nnzb on one GPU)
};

struct ModelArgs {
  const double* params;   // [n_es][npar]
  int npar;
  int max_iters;
  double abs_tol, rel_tol;
  double stab_mult;       // Mechanics "stabilization multiplier"
  double thickness;       // MechanicsPlaneStress "thickness"
};

struct FwdArgs {
  MeshArgs mesh;
  ModelArgs model;
  const double* x;
  const double* x_prev;
  const double* xi_prev;
  double* xi;             // in: current-field values (initial guess for some models); out: solved
  long long xi_ld;
  double* vals;           // BSR values (overwritten with the assembled matrix), may be nullptr
  double* emat;           // element-matrix scratch [n_elems+1][NX][NX] (required when vals is given)
  double* b;              // residual (+=), may be nullptr
  signed char* path;      // per-element branch (0 elastic / 1 plastic), may be nullptr
  int* n_failed;          // device counter of failed local solves
  double* elem_J;         // optional [n_elems][NX][NX] element Jacobians (reference dof order)
  double* elem_R;         // optional [n_elems][NX]
  cudaEvent_t elements_done;  // host side only: recorded between the element kernel and the BSR gather (may be null)
};

}  // namespace c8
