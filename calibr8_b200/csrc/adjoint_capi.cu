// C-ABI implementation, part 3: adjoint / QoI kernels (include/c8b200.h).
#include "c8b200.h"
#include "context.cuh"

using namespace c8;

namespace c8 {
__global__ void k_axpby_pub(double a, const double* __restrict__ x, double b, double* __restrict__ y,
                            long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    y[i] = a * x[i] + (b == 0.0 ? 0.0 : b * y[i]);
}
}  // namespace c8

static QoiArgs to_args(const c8_qoi* q) {
  QoiArgs a{};
  if (!q) { a.type = QOI_AVG_DISP; return a; }
  a.type = q->type;
  for (int k = 0; k < 3; ++k) a.weights[k] = q->weights[k];
  a.balance_factor = q->balance_factor;
  a.dt_over_T = q->dt_over_T;
  a.inv_area = q->inv_area;
  a.load_mismatch = q->load_mismatch;
  a.coord_idx = q->coord_idx;
  a.coord_value = q->coord_value;
  a.coord_tol = q->coord_tol;
  a.reaction_force_comp = q->reaction_force_comp;
  a.measured = q->measured_dev;
  a.facet = (const signed char*)q->facet_dev;
  a.compute_torque = q->compute_torque;
  a.normal_2d[0] = q->normal_2d[0]; a.normal_2d[1] = q->normal_2d[1];
  return a;
}

static AdjArgs base_args(c8_ctx* ctx, const c8_qoi* q, const double* x, const double* xp,
                         const double* xi, const double* xip) {
  AdjArgs a{};
  a.mesh = ctx->mesh_args();
  a.model = ctx->model;
  a.qoi = to_args(q);
  a.x = x; a.x_prev = xp; a.xi = xi; a.xi_prev = xip; a.xi_ld = ctx->xi_ld;
  return a;
}

extern "C" {

int c8_adjoint_jacobian(c8_ctx* ctx, const c8_qoi* qoi, const double* x, const double* xp,
                        const double* xi, const double* xip, double* g, const double* f,
                        double* AT_vals, double* rhs) {
  C8_REQUIRE(ctx, ctx->kt != nullptr, "c8_set_model must be called first");
  AdjArgs a = base_args(ctx, qoi, x, xp, xi, xip);
  a.g = g; a.f = const_cast<double*>(f); a.vals = AT_vals; a.b = rhs;
  if (AT_vals) {
    a.emat = element_scratch(ctx);
    if (!a.emat) return C8_ERR_CUDA;
  }
  a.tile_counter = ctx->d_nfailed + 1;
  C8_CUDA(ctx, cudaMemsetAsync(a.tile_counter, 0, sizeof(int), ctx->stream));
  ctx->kt->adjoint_jacobian(a, ctx->stream);
  C8_CUDA(ctx, cudaGetLastError());
  return C8_OK;
}

int c8_adjoint_local(c8_ctx* ctx, const double* x, const double* xp, const double* xi,
                     const double* xip, const double* z, double* phi, double* g, double* f) {
  C8_REQUIRE(ctx, ctx->kt != nullptr, "c8_set_model must be called first");
  AdjArgs a = base_args(ctx, nullptr, x, xp, xi, xip);
  a.z = z; a.phi = phi; a.g = g; a.f = f;
  ctx->kt->adjoint_local(a, ctx->stream);
  C8_CUDA(ctx, cudaGetLastError());
  return C8_OK;
}

int c8_qoi_value(c8_ctx* ctx, const c8_qoi* qoi, const double* x, const double* xp,
                 const double* xi, const double* xip, int mode, double* scalars_dev) {
  C8_REQUIRE(ctx, ctx->kt != nullptr, "c8_set_model must be called first");
  AdjArgs a = base_args(ctx, qoi, x, xp, xi, xip);
  a.mesh.n_elems = ctx->n_owned_elems;  // every element is counted once across ranks
  a.scalars = scalars_dev;
  ctx->kt->qoi_value(a, mode, ctx->stream);
  C8_CUDA(ctx, cudaGetLastError());
  return C8_OK;
}

int c8_qoi_gradient(c8_ctx* ctx, const c8_qoi* qoi, const double* x, const double* xp,
                    const double* xi, const double* xip, const double* z, const double* phi,
                    double* grad_dev) {
  C8_REQUIRE(ctx, ctx->kt != nullptr, "c8_set_model must be called first");
  AdjArgs a = base_args(ctx, qoi, x, xp, xi, xip);
  a.mesh.n_elems = ctx->n_owned_elems;
  a.z = z; a.phi = const_cast<double*>(phi); a.grad = grad_dev;
  ctx->kt->qoi_gradient(a, ctx->stream);
  C8_CUDA(ctx, cudaGetLastError());
  return C8_OK;
}

int c8_axpby(c8_ctx* ctx, double a, const double* x_dev, double b, double* y_dev, int64_t n) {
  if (n <= 0) return C8_OK;
  long long g = (n + 255) / 256;
  if (g > 148 * 8) g = 148 * 8;
  k_axpby_pub<<<(unsigned)g, 256, 0, ctx->stream>>>(a, x_dev, b, y_dev, n);
  C8_CUDA(ctx, cudaGetLastError());
  return C8_OK;
}

int c8_get_coords(c8_ctx* ctx, double* coords_host) {
  for (int n = 0; n < ctx->n_nodes; ++n) {
    for (int k = 0; k < 3; ++k)
      coords_host[size_t(n) * 3 + k] = k < ctx->dim ? ctx->h_coords[size_t(n) * ctx->dim + k] : 0.0;
  }
  return C8_OK;
}

int c8_get_conn(c8_ctx* ctx, int32_t* conn_host) {
  for (size_t k = 0; k < ctx->h_conn.size(); ++k) conn_host[k] = ctx->h_conn[k];
  return C8_OK;
}

void* c8_get_stream(c8_ctx* ctx) { return (void*)ctx->stream; }

}  // extern "C"

// ---- virtual fields method ---------------------------------------------------------------
static VfmArgs vfm_args(c8_ctx* ctx, const double* x, const double* xp, const double* xip,
                        double* xi) {
  VfmArgs a{};
  a.mesh = ctx->mesh_args();
  a.model = ctx->model;
  a.x = x; a.x_prev = xp; a.xi_prev = xip; a.xi = xi; a.xi_ld = ctx->xi_ld;
  a.n_failed = ctx->d_nfailed;
  return a;
}

extern "C" {

int c8_vfm_forward(c8_ctx* ctx, const double* x_meas, const double* x_meas_prev,
                   const double* xi_prev, double* xi, double* b, double* dR, double* local_sens,
                   int* n_failed) {
  C8_REQUIRE(ctx, ctx->kt != nullptr, "c8_set_model must be called first");
  C8_REQUIRE(ctx, ctx->kt->vfm_forward != nullptr,
             "virtual fields need a single-residual global residual (mechanics_plane_stress)");
  VfmArgs a = vfm_args(ctx, x_meas, x_meas_prev, xi_prev, xi);
  a.b = b; a.dR = dR; a.local_sens = local_sens;
  C8_CUDA(ctx, cudaMemsetAsync(ctx->d_nfailed, 0, sizeof(int), ctx->stream));
  ctx->kt->vfm_forward(a, ctx->stream);
  C8_CUDA(ctx, cudaGetLastError());
  if (n_failed) {
    int rc = c8::fetch_n_failed(ctx, n_failed);
    if (rc != C8_OK) return rc;
  }
  return C8_OK;
}

int c8_vfm_adjoint(c8_ctx* ctx, const double* x_meas, const double* x_meas_prev,
                   const double* xi, const double* xi_prev, const double* w, double s,
                   double* hist, double* grad) {
  C8_REQUIRE(ctx, ctx->kt != nullptr, "c8_set_model must be called first");
  C8_REQUIRE(ctx, ctx->kt->vfm_adjoint != nullptr,
             "virtual fields need a single-residual global residual (mechanics_plane_stress)");
  VfmArgs a = vfm_args(ctx, x_meas, x_meas_prev, xi_prev, const_cast<double*>(xi));
  a.mesh.n_elems = ctx->n_owned_elems;
  a.w = w; a.s = s; a.hist = hist; a.grad = grad;
  ctx->kt->vfm_adjoint(a, ctx->stream);
  C8_CUDA(ctx, cudaGetLastError());
  return C8_OK;
}

}  // extern "C"
