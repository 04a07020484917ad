// K9 / K10: the global linear algebra of the Newton and adjoint solves on the device.
// Replaces Tpetra (src/linear_alg.cpp), Belos/Teko/MueLu (src/linear_solve.cpp:22-124) and the
// Dirichlet row replacement of src/dbcs.cpp:56-119 for this path:
//   * BSR (NB x NB node blocks) SpMV            -- HBM-bound, 8 B/value + 4 B/block index
//   * fused multi-dot / axpy / norm reductions  -- deterministic two-stage sums
//   * block-Jacobi (one NB x NB inverse per node) right preconditioner
//   * GMRES(m) with device-resident Arnoldi (krylov.cu), block-Jacobi / aggregation-AMG (amg.cu)
//   * Dirichlet rows: zero the row, keep the diagonal, R = diag*(u - g)
#include "linalg.cuh"


using namespace c8;

extern "C" {

int c8_spmv(c8_ctx* ctx, const double* A_vals_dev, const double* x_dev, double* y_dev) {
  C8_REQUIRE(ctx, ctx->kt != nullptr, "c8_set_model must be called first");
  LinAlg la(ctx);
  la.halo(const_cast<double*>(x_dev));  // ghost entries of the input (no-op on one GPU)
  la.spmv(A_vals_dev, x_dev, y_dev);
  C8_CUDA(ctx, cudaGetLastError());
  return C8_OK;
}

int c8_apply_dbc(c8_ctx* ctx, double* A_vals_dev, double* R_dev, const double* x_dev,
                 const int32_t* dbc_node_dev, const int32_t* dbc_eq_dev, const double* dbc_val_dev,
                 int n_dbc, int is_adjoint) {
  C8_REQUIRE(ctx, ctx->kt != nullptr, "c8_set_model must be called first");
  if (n_dbc == 0) return C8_OK;
  const int block = 128, grid = (n_dbc + block - 1) / block;
  switch (ctx->kt->nb) {
    case 2: k_apply_dbc<2><<<grid, block, 0, ctx->stream>>>(ctx->d_rowptr, ctx->d_colind, A_vals_dev, R_dev, x_dev, dbc_node_dev, dbc_eq_dev, dbc_val_dev, n_dbc, is_adjoint); break;
    case 3: k_apply_dbc<3><<<grid, block, 0, ctx->stream>>>(ctx->d_rowptr, ctx->d_colind, A_vals_dev, R_dev, x_dev, dbc_node_dev, dbc_eq_dev, dbc_val_dev, n_dbc, is_adjoint); break;
    default: k_apply_dbc<4><<<grid, block, 0, ctx->stream>>>(ctx->d_rowptr, ctx->d_colind, A_vals_dev, R_dev, x_dev, dbc_node_dev, dbc_eq_dev, dbc_val_dev, n_dbc, is_adjoint); break;
  }
  C8_CUDA(ctx, cudaGetLastError());
  return C8_OK;
}

int c8_apply_tbc(c8_ctx* ctx, double* R_dev, const int32_t* side_nodes_dev, const double* traction_dev,
                 int n_sides) {
  C8_REQUIRE(ctx, ctx->kt != nullptr, "c8_set_model must be called first");
  if (n_sides == 0) return C8_OK;
  const int block = 128, grid = (n_sides + block - 1) / block;
  if (ctx->dim == 3)
    k_apply_tbc<3><<<grid, block, 0, ctx->stream>>>(ctx->d_coords, side_nodes_dev, traction_dev, R_dev, n_sides,
                                                    ctx->kt->nb, ctx->n_owned_nodes);
  else
    k_apply_tbc<2><<<grid, block, 0, ctx->stream>>>(ctx->d_coords, side_nodes_dev, traction_dev, R_dev, n_sides,
                                                    ctx->kt->nb, ctx->n_owned_nodes);
  C8_CUDA(ctx, cudaGetLastError());
  return C8_OK;
}

}  // extern "C"
