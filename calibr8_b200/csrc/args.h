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
  // chunked two-phase assembly (forward path): the elements are processed in chunks of chunk_elems
  // so that a chunk's element matrices are still in L2 when its gather reads them.  Plan entries of
  // chunk c are [cg_ptr[c], cg_ptr[c+1]): block cg_blk[i] (bit 31 set: first chunk touching the
  // block -> overwrite, else accumulate), contributions gsrc[cg_k[2i] .. cg_k[2i+1])
  int chunk_elems;        // 0: one pass over the whole mesh (full scratch)
  int n_chunks;
  const int* cg_ptr;      // host copy of the entry ranges lives in the context; device array for the kernel
  const unsigned* cg_blk;
  const int* cg_k;
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
  double* emat;           // element-matrix scratch [n_elems+1][NX][NX] (required when vals is given);
                          // chunked: [chunk_elems+1][NX][NX], slot = element - elem_begin
  int elem_begin, elem_end;   // element range of this launch (0, n_elems unless chunked)
  double* b;              // residual (+=), may be nullptr
  signed char* path;      // per-element branch (0 elastic / 1 plastic), may be nullptr
  int* n_failed;          // device counter of failed local solves; n_failed[1]: tile counter of the persistent
                          // element kernel (both zeroed by the caller before the launch)
  double* elem_J;         // optional [n_elems][NX][NX] element Jacobians (reference dof order)
  double* elem_R;         // optional [n_elems][NX]
  const int* cg_ptr_host;     // host side only: chunk -> first plan entry [n_chunks + 1], nullptr: one pass
  const int* cg_end_host;     // host side only: chunk -> end of its entries in the OWNED rows (a prefix)
  cudaEvent_t elements_done;  // host side only: recorded between the element kernel and the BSR gather (may be null)
};

}  // namespace c8
