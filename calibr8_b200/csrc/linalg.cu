// K9 / K10: the global linear algebra of the Newton and adjoint solves on the device.
// Replaces Tpetra (src/linear_alg.cpp), Belos/Teko/MueLu (src/linear_solve.cpp:22-124) and the
// Dirichlet row replacement of src/dbcs.cpp:56-119 for this path:
//   * BSR (NB x NB node blocks) SpMV            -- HBM-bound, 8 B/value + 4 B/block index
//   * fused multi-dot / axpy / norm reductions  -- deterministic two-stage sums
//   * block-Jacobi (one NB x NB inverse per node) right preconditioner
//   * restarted GMRES(m), classical Gram-Schmidt with one fused multi-dot per iteration
//   * Dirichlet rows: zero the row, keep the diagonal, R = diag*(u - g)
#include <cmath>
#include <cstdio>
#include <vector>

#include "c8b200.h"
#include "context.cuh"

namespace c8 {

// y = A x ; one thread per scalar row (node, r); loops over the node's blocks
template <int NB>
__global__ void k_bsr_spmv(const int* __restrict__ rowptr, const int* __restrict__ colind,
                           const double* __restrict__ vals, const double* __restrict__ x,
                           double* __restrict__ y, int n_nodes) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_nodes * NB) return;
  const int node = i / NB, r = i % NB;
  double s = 0.0;
  const int b0 = rowptr[node], b1 = rowptr[node + 1];
  for (int k = b0; k < b1; ++k) {
    const double* a = vals + (size_t(k) * NB + r) * NB;
    const double* xv = x + size_t(__ldg(&colind[k])) * NB;
#pragma unroll
    for (int c = 0; c < NB; ++c) s = fma(__ldg(&a[c]), __ldg(&xv[c]), s);
  }
  y[i] = s;
}

// block-Jacobi: Dinv[node] = inverse of the diagonal NB x NB block (Gauss-Jordan, partial pivoting)
template <int NB>
__global__ void k_block_jacobi_setup(const int* __restrict__ rowptr, const int* __restrict__ colind,
                                     const double* __restrict__ vals, double* __restrict__ dinv,
                                     int n_nodes) {
  const int node = blockIdx.x * blockDim.x + threadIdx.x;
  if (node >= n_nodes) return;
  int kd = -1;
  for (int k = rowptr[node]; k < rowptr[node + 1]; ++k)
    if (colind[k] == node) { kd = k; break; }
  double a[NB][2 * NB];
  for (int r = 0; r < NB; ++r)
    for (int c = 0; c < NB; ++c) {
      a[r][c] = kd >= 0 ? vals[(size_t(kd) * NB + r) * NB + c] : (r == c ? 1.0 : 0.0);
      a[r][NB + c] = (r == c) ? 1.0 : 0.0;
    }
  for (int k = 0; k < NB; ++k) {
    int p = k;
    double best = fabs(a[k][k]);
    for (int r = k + 1; r < NB; ++r)
      if (fabs(a[r][k]) > best) { best = fabs(a[r][k]); p = r; }
    if (best == 0.0) { a[k][k] = 1.0; p = k; }  // empty row (isolated dof): identity
    for (int c = 0; c < 2 * NB; ++c) { const double t = a[k][c]; a[k][c] = a[p][c]; a[p][c] = t; }
    const double inv = 1.0 / a[k][k];
    for (int c = 0; c < 2 * NB; ++c) a[k][c] *= inv;
    for (int r = 0; r < NB; ++r) {
      if (r == k) continue;
      const double m = a[r][k];
      for (int c = 0; c < 2 * NB; ++c) a[r][c] -= m * a[k][c];
    }
  }
  for (int r = 0; r < NB; ++r)
    for (int c = 0; c < NB; ++c) dinv[(size_t(node) * NB + r) * NB + c] = a[r][NB + c];
}
template <int NB>
__global__ void k_block_jacobi_apply(const double* __restrict__ dinv, const double* __restrict__ x,
                                     double* __restrict__ y, int n_nodes) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_nodes * NB) return;
  const int node = i / NB, r = i % NB;
  double s = 0.0;
#pragma unroll
  for (int c = 0; c < NB; ++c) s = fma(dinv[(size_t(node) * NB + r) * NB + c], x[size_t(node) * NB + c], s);
  y[i] = s;
}

// out[j] = sum_i V[j*ld + i] * w[i], j < nv  (stage 1: per-block partials, stage 2: final)
constexpr int DOT_BLOCK = 256;
__global__ void k_multi_dot_partial(const double* __restrict__ V, long long ld,
                                    const double* __restrict__ w, int nv, long long n,
                                    double* __restrict__ partial) {
  __shared__ double sh[DOT_BLOCK / 32];
  for (int j = blockIdx.y; j < nv; j += gridDim.y) {
    double s = 0.0;
    const double* v = V + j * ld;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x)
      s = fma(v[i], w[i], s);
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
      double t = threadIdx.x < DOT_BLOCK / 32 ? sh[threadIdx.x] : 0.0;
      for (int o = 16; o > 0; o >>= 1) t += __shfl_down_sync(0xffffffffu, t, o);
      if (threadIdx.x == 0) partial[size_t(j) * gridDim.x + blockIdx.x] = t;
    }
    __syncthreads();
  }
}
__global__ void k_multi_dot_final(const double* __restrict__ partial, int nparts, int nv,
                                  double* __restrict__ out) {
  const int j = blockIdx.x;
  if (j >= nv) return;
  __shared__ double sh[32];
  double s = 0.0;
  for (int i = threadIdx.x; i < nparts; i += blockDim.x) s += partial[size_t(j) * nparts + i];
  for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    double t = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0.0;
    for (int o = 16; o > 0; o >>= 1) t += __shfl_down_sync(0xffffffffu, t, o);
    if (threadIdx.x == 0) out[j] = t;
  }
}
// w -= sum_j h[j] V[j]   (h on the device)
__global__ void k_multi_axpy_neg(const double* __restrict__ V, long long ld,
                                 const double* __restrict__ h, int nv, long long n,
                                 double* __restrict__ w) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    double s = w[i];
    for (int j = 0; j < nv; ++j) s = fma(-h[j], V[j * ld + i], s);
    w[i] = s;
  }
}
// y = a*x + b*y
__global__ void k_axpby(double a, const double* __restrict__ x, double b, double* __restrict__ y,
                        long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    y[i] = a * x[i] + (b == 0.0 ? 0.0 : b * y[i]);
}
// x += sum_j c[j] V[j]  (c on the host -> passed through a device array)
__global__ void k_multi_axpy(const double* __restrict__ V, long long ld, const double* __restrict__ c,
                             int nv, long long n, double* __restrict__ x) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    double s = x[i];
    for (int j = 0; j < nv; ++j) s = fma(c[j], V[j * ld + i], s);
    x[i] = s;
  }
}

// Dirichlet rows (src/dbcs.cpp:56-119): row (node, eq): zero everything but the diagonal entry,
// R = diag * (u - g)   (adjoint: R = 0).  One thread per constrained dof.
template <int NB>
__global__ void k_apply_dbc(const int* __restrict__ rowptr, const int* __restrict__ colind,
                            double* __restrict__ vals, double* __restrict__ R,
                            const double* __restrict__ x, const int* __restrict__ dbc_node,
                            const int* __restrict__ dbc_eq, const double* __restrict__ dbc_val,
                            int n_dbc, int is_adjoint) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_dbc) return;
  const int node = dbc_node[i], eq = dbc_eq[i];
  double diag = 0.0;
  for (int k = rowptr[node]; k < rowptr[node + 1]; ++k) {
    double* a = vals + (size_t(k) * NB + eq) * NB;
    const bool dblk = (colind[k] == node);
#pragma unroll
    for (int c = 0; c < NB; ++c) {
      if (dblk && c == eq) diag = a[c];
      else a[c] = 0.0;
    }
  }
  const size_t row = size_t(node) * NB + eq;
  R[row] = is_adjoint ? 0.0 : diag * (x[row] - dbc_val[i]);
}

static int grid_for(long long n, int block, int sms) {
  long long g = (n + block - 1) / block;
  const long long cap = (long long)sms * 8;
  return int(g < cap ? (g > 0 ? g : 1) : cap);
}

struct LinAlg {
  c8_ctx* ctx;
  int nb, n_nodes, sms;
  long long n;
  cudaStream_t s;
  explicit LinAlg(c8_ctx* c) : ctx(c) {
    // rows = owned nodes (all nodes on one GPU); vectors are allocated for all local nodes so
    // that SpMV can read ghost entries filled by the halo callback
    nb = c->kt->nb; n_nodes = c->n_owned_nodes; n = (long long)n_nodes * nb; s = c->stream;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device);
  }
  void halo(double* x) const { if (ctx->halo_cb) ctx->halo_cb(ctx->comm_user, x, nb); }
  void allreduce(double* buf_dev, int cnt) const {
    if (ctx->allreduce_cb) ctx->allreduce_cb(ctx->comm_user, buf_dev, cnt);
  }
  void spmv(const double* A, const double* x, double* y) const {
    const int block = 128, grid = int((n + block - 1) / block);
    switch (nb) {
      case 2: k_bsr_spmv<2><<<grid, block, 0, s>>>(ctx->d_rowptr, ctx->d_colind, A, x, y, n_nodes); break;
      case 3: k_bsr_spmv<3><<<grid, block, 0, s>>>(ctx->d_rowptr, ctx->d_colind, A, x, y, n_nodes); break;
      default: k_bsr_spmv<4><<<grid, block, 0, s>>>(ctx->d_rowptr, ctx->d_colind, A, x, y, n_nodes); break;
    }
  }
  void jacobi_setup(const double* A, double* dinv) const {
    const int block = 128, grid = (n_nodes + block - 1) / block;
    switch (nb) {
      case 2: k_block_jacobi_setup<2><<<grid, block, 0, s>>>(ctx->d_rowptr, ctx->d_colind, A, dinv, n_nodes); break;
      case 3: k_block_jacobi_setup<3><<<grid, block, 0, s>>>(ctx->d_rowptr, ctx->d_colind, A, dinv, n_nodes); break;
      default: k_block_jacobi_setup<4><<<grid, block, 0, s>>>(ctx->d_rowptr, ctx->d_colind, A, dinv, n_nodes); break;
    }
  }
  void jacobi_apply(const double* dinv, const double* x, double* y) const {
    const int block = 128, grid = int((n + block - 1) / block);
    switch (nb) {
      case 2: k_block_jacobi_apply<2><<<grid, block, 0, s>>>(dinv, x, y, n_nodes); break;
      case 3: k_block_jacobi_apply<3><<<grid, block, 0, s>>>(dinv, x, y, n_nodes); break;
      default: k_block_jacobi_apply<4><<<grid, block, 0, s>>>(dinv, x, y, n_nodes); break;
    }
  }
  // out_dev[j] = V[j] . w
  void multi_dot(const double* V, long long ld, const double* w, int nv, double* partial,
                 double* out_dev) const {
    const int gx = grid_for(n, DOT_BLOCK, sms) > 256 ? 256 : grid_for(n, DOT_BLOCK, sms);
    dim3 grid(gx, nv < 64 ? nv : 64);
    k_multi_dot_partial<<<grid, DOT_BLOCK, 0, s>>>(V, ld, w, nv, n, partial);
    k_multi_dot_final<<<nv, 256, 0, s>>>(partial, gx, nv, out_dev);
    allreduce(out_dev, nv);
  }
};

}  // namespace c8

using namespace c8;

// ---- per-context solver workspace ------------------------------------------------
struct c8_solver_ws {
  int m = 0;
  long long n = 0;
  double *V = nullptr, *w = nullptr, *z = nullptr, *dinv = nullptr, *partial = nullptr,
         *dots = nullptr, *coef = nullptr;
  double* h_dots = nullptr;  // pinned
};
static c8_solver_ws g_ws_dummy;

static int ensure_ws(c8_ctx* ctx, c8_solver_ws& ws, int m) {
  const long long n = (long long)ctx->n_nodes * ctx->kt->nb;  // all local nodes incl. ghosts
  if (ws.m >= m && ws.n == n) return C8_OK;
  cudaFree(ws.V); cudaFree(ws.w); cudaFree(ws.z); cudaFree(ws.dinv); cudaFree(ws.partial);
  cudaFree(ws.dots); cudaFree(ws.coef);
  if (ws.h_dots) cudaFreeHost(ws.h_dots);
  ws = c8_solver_ws();
  C8_CUDA(ctx, cudaMalloc(&ws.V, size_t(m + 1) * n * sizeof(double)));
  C8_CUDA(ctx, cudaMalloc(&ws.w, n * sizeof(double)));
  C8_CUDA(ctx, cudaMalloc(&ws.z, n * sizeof(double)));
  C8_CUDA(ctx, cudaMalloc(&ws.dinv, size_t(ctx->n_nodes) * ctx->kt->nb * ctx->kt->nb * sizeof(double)));
  C8_CUDA(ctx, cudaMalloc(&ws.partial, size_t(m + 2) * 256 * sizeof(double)));
  C8_CUDA(ctx, cudaMalloc(&ws.dots, size_t(m + 2) * sizeof(double)));
  C8_CUDA(ctx, cudaMalloc(&ws.coef, size_t(m + 2) * sizeof(double)));
  C8_CUDA(ctx, cudaMallocHost(&ws.h_dots, size_t(m + 2) * sizeof(double)));
  ws.m = m; ws.n = n;
  return C8_OK;
}

// one workspace per context, keyed by pointer (contexts are few)
#include <map>
static std::map<c8_ctx*, c8_solver_ws> g_ws;

extern "C" {

int c8_spmv(c8_ctx* ctx, const double* A_vals_dev, const double* x_dev, double* y_dev) {
  C8_REQUIRE(ctx, ctx->kt != nullptr, "c8_set_model must be called first");
  LinAlg la(ctx);
  la.halo(const_cast<double*>(x_dev));  // ghost entries of the input (no-op on one GPU)
  la.spmv(A_vals_dev, x_dev, y_dev);
  C8_CUDA(ctx, cudaGetLastError());
  return C8_OK;
}

int c8_dot(c8_ctx* ctx, const double* x_dev, const double* y_dev, double* out_host) {
  C8_REQUIRE(ctx, ctx->kt != nullptr, "c8_set_model must be called first");
  c8_solver_ws& ws = g_ws[ctx];
  int rc = ensure_ws(ctx, ws, ws.m > 0 ? ws.m : 8);
  if (rc != C8_OK) return rc;
  LinAlg la(ctx);
  la.multi_dot(x_dev, 0, y_dev, 1, ws.partial, ws.dots);
  C8_CUDA(ctx, cudaMemcpyAsync(ws.h_dots, ws.dots, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  C8_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  *out_host = ws.h_dots[0];
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

// Restarted GMRES(m) with block-Jacobi right preconditioning: solves A x = b, x_dev in: initial
// guess, out: solution.  Converges on ||b - A x|| <= max(rel_tol*||b - A x0||, abs_tol).
// info_host[0] = iterations, info_host[1] = final residual norm, info_host[2] = initial norm.
int c8_gmres(c8_ctx* ctx, const double* A, const double* b, double* x, int restart, int max_iters,
             double rel_tol, double abs_tol, double* info_host) {
  C8_REQUIRE(ctx, ctx->kt != nullptr, "c8_set_model must be called first");
  const int m = restart;
  c8_solver_ws& ws = g_ws[ctx];
  int rc = ensure_ws(ctx, ws, m);
  if (rc != C8_OK) return rc;
  LinAlg la(ctx);
  const long long n = la.n;      // owned dofs: the range every vector operation runs over
  const long long ld = ws.n;     // allocation stride of a Krylov vector (owned + ghost)
  cudaStream_t s = ctx->stream;
  const int ag = grid_for(n, 256, la.sms);
  la.jacobi_setup(A, ws.dinv);
  std::vector<double> H(size_t(m + 1) * m, 0.0), cs(m), sn(m), g(m + 1), y(m);
  auto nrm2 = [&](const double* v, double* out) -> int {
    la.multi_dot(v, 0, v, 1, ws.partial, ws.dots);
    C8_CUDA(ctx, cudaMemcpyAsync(ws.h_dots, ws.dots, sizeof(double), cudaMemcpyDeviceToHost, s));
    C8_CUDA(ctx, cudaStreamSynchronize(s));
    *out = std::sqrt(ws.h_dots[0]);
    return C8_OK;
  };
  int total = 0;
  double beta0 = -1.0, beta = 0.0, target = 0.0;
  while (true) {
    // r = b - A x  -> V[0]
    la.halo(x);
    la.spmv(A, x, ws.w);
    C8_CUDA(ctx, cudaMemcpyAsync(ws.V, b, n * sizeof(double), cudaMemcpyDeviceToDevice, s));
    k_axpby<<<ag, 256, 0, s>>>(-1.0, ws.w, 1.0, ws.V, n);
    if ((rc = nrm2(ws.V, &beta)) != C8_OK) return rc;
    if (beta0 < 0) { beta0 = beta; target = std::max(rel_tol * beta0, abs_tol); }
    if (beta <= target || total >= max_iters || !(beta == beta)) break;
    k_axpby<<<ag, 256, 0, s>>>(0.0, ws.V, 1.0 / beta, ws.V, n);  // V0 *= 1/beta
    std::fill(g.begin(), g.end(), 0.0);
    g[0] = beta;
    int j = 0;
    for (; j < m && total < max_iters; ++j, ++total) {
      double* vj1 = ws.V + size_t(j + 1) * ld;
      la.jacobi_apply(ws.dinv, ws.V + size_t(j) * ld, ws.z);
      la.halo(ws.z);
      la.spmv(A, ws.z, vj1);
      // classical Gram-Schmidt, two passes (CGS2) for orthogonality: h = V^T w ; w -= V h
      for (int pass = 0; pass < 2; ++pass) {
        la.multi_dot(ws.V, ld, vj1, j + 1, ws.partial, ws.dots);
        k_multi_axpy_neg<<<ag, 256, 0, s>>>(ws.V, ld, ws.dots, j + 1, n, vj1);
        C8_CUDA(ctx, cudaMemcpyAsync(ws.h_dots, ws.dots, (j + 1) * sizeof(double),
                                     cudaMemcpyDeviceToHost, s));
        C8_CUDA(ctx, cudaStreamSynchronize(s));
        for (int i = 0; i <= j; ++i) H[size_t(i) * m + j] = (pass == 0 ? 0.0 : H[size_t(i) * m + j]) + ws.h_dots[i];
      }
      double hn;
      if ((rc = nrm2(vj1, &hn)) != C8_OK) return rc;
      H[size_t(j + 1) * m + j] = hn;
      if (hn > 0) k_axpby<<<ag, 256, 0, s>>>(0.0, vj1, 1.0 / hn, vj1, n);
      // Givens
      for (int i = 0; i < j; ++i) {
        const double t = cs[i] * H[size_t(i) * m + j] + sn[i] * H[size_t(i + 1) * m + j];
        H[size_t(i + 1) * m + j] = -sn[i] * H[size_t(i) * m + j] + cs[i] * H[size_t(i + 1) * m + j];
        H[size_t(i) * m + j] = t;
      }
      const double a = H[size_t(j) * m + j], bb = H[size_t(j + 1) * m + j];
      const double d = std::hypot(a, bb);
      cs[j] = d > 0 ? a / d : 1.0;
      sn[j] = d > 0 ? bb / d : 0.0;
      H[size_t(j) * m + j] = d;
      H[size_t(j + 1) * m + j] = 0.0;
      g[j + 1] = -sn[j] * g[j];
      g[j] = cs[j] * g[j];
      if (std::fabs(g[j + 1]) <= target) { ++j; ++total; break; }
    }
    // back substitution, x += M^-1 (V y)
    for (int i = j - 1; i >= 0; --i) {
      double t = g[i];
      for (int k = i + 1; k < j; ++k) t -= H[size_t(i) * m + k] * y[k];
      y[i] = t / H[size_t(i) * m + i];
    }
    for (int i = 0; i < j; ++i) ws.h_dots[i] = y[i];
    C8_CUDA(ctx, cudaMemcpyAsync(ws.coef, ws.h_dots, j * sizeof(double), cudaMemcpyHostToDevice, s));
    C8_CUDA(ctx, cudaMemsetAsync(ws.w, 0, n * sizeof(double), s));
    k_multi_axpy<<<ag, 256, 0, s>>>(ws.V, ld, ws.coef, j, n, ws.w);
    la.jacobi_apply(ws.dinv, ws.w, ws.z);
    k_axpby<<<ag, 256, 0, s>>>(1.0, ws.z, 1.0, x, n);
    C8_CUDA(ctx, cudaStreamSynchronize(s));
  }
  if (info_host) { info_host[0] = total; info_host[1] = beta; info_host[2] = beta0; }
  C8_CUDA(ctx, cudaGetLastError());
  return (beta <= target) ? C8_OK : C8_ERR_USAGE - 1;  // -4: not converged
}

void c8_linalg_release(c8_ctx* ctx) {
  auto it = g_ws.find(ctx);
  if (it == g_ws.end()) return;
  c8_solver_ws& ws = it->second;
  cudaFree(ws.V); cudaFree(ws.w); cudaFree(ws.z); cudaFree(ws.dinv); cudaFree(ws.partial);
  cudaFree(ws.dots); cudaFree(ws.coef);
  if (ws.h_dots) cudaFreeHost(ws.h_dots);
  g_ws.erase(it);
}

}  // extern "C"
