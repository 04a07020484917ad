// Device linear-algebra building blocks shared by linalg.cu (C ABI), krylov.cu (GMRES) and amg.cu
// (multilevel preconditioner): BSR SpMV, block-Jacobi, fused multi-dot / multi-axpy, Dirichlet rows.
#pragma once
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "c8b200.h"
#include "context.cuh"

namespace c8 {

// Programmatic dependent launch (sm_90+): the solver's iteration is a chain of ~50 small dependent
// kernels; launched with the programmatic-stream-serialization attribute, a kernel's grid is
// scheduled while its predecessor still runs and parks at griddepcontrol.wait until the
// predecessor's grid has completed and flushed -- the launch latency leaves the critical path, the
// data dependency stays.  Every kernel launched through pdl_launch() calls pdl_wait() first.
// C8_PDL=0 launches the same kernels without the attribute.
__device__ __forceinline__ void pdl_wait() {
#if defined(__CUDA_ARCH__) && __CUDA_ARCH__ >= 900
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
}
inline bool pdl_enabled() {
  static const bool on = [] { const char* e = getenv("C8_PDL"); return !(e && e[0] == '0'); }();
  return on;
}
struct PdlLaunch {
  dim3 grid, block;
  size_t smem;
  cudaStream_t stream;
  template <class... P, class... A>
  void operator()(void (*kernel)(P...), A&&... args) const {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    cudaLaunchKernelEx(&cfg, kernel, P(args)...);
  }
};
inline PdlLaunch pdl_launch(dim3 grid, dim3 block, size_t smem, cudaStream_t stream) {
  return PdlLaunch{grid, block, smem, stream};
}


// one row of an NB x NB block (or the NB entries of a node) with the widest loads the alignment
// allows: 16-byte LDG for NB = 4 (float4 / 2 x double2) and NB = 2 doubles; values widened to fp64
template <int NB, class F>
__device__ __forceinline__ void ld_row(const F* __restrict__ a, double (&v)[NB]) {
  if constexpr (NB == 4 && sizeof(F) == 4) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(a));
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  } else if constexpr (NB == 4 && sizeof(F) == 8) {
    const double2 t0 = __ldg(reinterpret_cast<const double2*>(a));
    const double2 t1 = __ldg(reinterpret_cast<const double2*>(a) + 1);
    v[0] = t0.x; v[1] = t0.y; v[2] = t1.x; v[3] = t1.y;
  } else if constexpr (NB == 2 && sizeof(F) == 8) {
    const double2 t = __ldg(reinterpret_cast<const double2*>(a));
    v[0] = t.x; v[1] = t.y;
  } else if constexpr (NB == 2 && sizeof(F) == 4) {
    const float2 t = __ldg(reinterpret_cast<const float2*>(a));
    v[0] = t.x; v[1] = t.y;
  } else {
#pragma unroll
    for (int c = 0; c < NB; ++c) v[c] = double(__ldg(&a[c]));
  }
}

// y = A x ; one thread per scalar row (node, r); loops over the node's blocks
template <int NB>
__global__ void k_bsr_spmv(const int* __restrict__ rowptr, const int* __restrict__ colind,
                           const double* __restrict__ vals, const double* __restrict__ x,
                           double* __restrict__ y, int n_nodes) {
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_nodes * NB) return;
  const int node = i / NB, r = i % NB;
  double s = 0.0;
  const int b0 = rowptr[node], b1 = rowptr[node + 1];
#pragma unroll 4
  for (int k = b0; k < b1; ++k) {
    double a[NB], xv[NB];
    ld_row<NB, double>(vals + (size_t(k) * NB + r) * NB, a);
    ld_row<NB, double>(x + size_t(__ldg(&colind[k])) * NB, xv);
#pragma unroll
    for (int c = 0; c < NB; ++c) s = fma(a[c], xv[c], s);
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
  pdl_wait();
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
static __global__ void k_multi_dot_partial(const double* __restrict__ V, long long ld,
                                    const double* __restrict__ w, int nv, long long n,
                                    double* __restrict__ partial) {
  pdl_wait();
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
static __global__ void k_multi_dot_final(const double* __restrict__ partial, int nparts, int nv,
                                  double* __restrict__ out) {
  pdl_wait();
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
static __global__ void k_multi_axpy_neg(const double* __restrict__ V, long long ld,
                                 const double* __restrict__ h, int nv, long long n,
                                 double* __restrict__ w) {
  pdl_wait();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    double s = w[i];
    for (int j = 0; j < nv; ++j) s = fma(-h[j], V[j * ld + i], s);
    w[i] = s;
  }
}
// y = a*x + b*y
static __global__ void k_axpby(double a, const double* __restrict__ x, double b, double* __restrict__ y,
                        long long n) {
  pdl_wait();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    y[i] = a * x[i] + (b == 0.0 ? 0.0 : b * y[i]);
}
// x += sum_j c[j] V[j]  (c on the host -> passed through a device array)
static __global__ void k_multi_axpy(const double* __restrict__ V, long long ld, const double* __restrict__ c,
                             int nv, long long n, double* __restrict__ x) {
  pdl_wait();
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

// Traction boundary conditions, apply_primal_tbc of the reference (src/tbcs.cpp:17-86):
//   R[n, d] -= T_d(x_q, t) N_n(x_q) w dv   over the side quadrature of the local variables' order
// (getIPFitShape(dim, 1): ONE point at the side centroid, where every linear side basis function is
// 1/DIM and w dv is the side's area (3-D) or length (2-D)).  One thread per side; trac holds the
// traction vector at the side's quadrature point (the host evaluates the deck's expression strings
// there).  Rows of ghost nodes belong to another part and are skipped.
template <int DIM>
__global__ void k_apply_tbc(const double* __restrict__ coords, const int* __restrict__ side_nodes,
                            const double* __restrict__ trac, double* __restrict__ R, int n_sides,
                            int nb, int n_row_nodes) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_sides) return;
  int nd[DIM];
  double X[DIM][DIM];
#pragma unroll
  for (int a = 0; a < DIM; ++a) {
    nd[a] = side_nodes[s * DIM + a];
#pragma unroll
    for (int k = 0; k < DIM; ++k) X[a][k] = coords[size_t(nd[a]) * DIM + k];
  }
  double wdv;
  if constexpr (DIM == 3) {
    const double ax = X[1][0] - X[0][0], ay = X[1][1] - X[0][1], az = X[1][2] - X[0][2];
    const double bx = X[2][0] - X[0][0], by = X[2][1] - X[0][1], bz = X[2][2] - X[0][2];
    const double cx = ay * bz - az * by, cy = az * bx - ax * bz, cz = ax * by - ay * bx;
    wdv = 0.5 * sqrt(cx * cx + cy * cy + cz * cz);
  } else {
    const double ax = X[1][0] - X[0][0], ay = X[1][1] - X[0][1];
    wdv = sqrt(ax * ax + ay * ay);
  }
#pragma unroll
  for (int a = 0; a < DIM; ++a) {
    if (nd[a] >= n_row_nodes) continue;
#pragma unroll
    for (int d = 0; d < DIM; ++d)
      atomicAdd(&R[size_t(nd[a]) * nb + d], -trac[s * DIM + d] * (1.0 / DIM) * wdv);
  }
}

static inline int grid_for(long long n, int block, int sms) {
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
      case 2: pdl_launch(grid, block, 0, s)(k_bsr_spmv<2>, ctx->d_rowptr, ctx->d_colind, A, x, y, n_nodes); break;
      case 3: pdl_launch(grid, block, 0, s)(k_bsr_spmv<3>, ctx->d_rowptr, ctx->d_colind, A, x, y, n_nodes); break;
      default: pdl_launch(grid, block, 0, s)(k_bsr_spmv<4>, ctx->d_rowptr, ctx->d_colind, A, x, y, n_nodes); break;
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
      case 2: pdl_launch(grid, block, 0, s)(k_block_jacobi_apply<2>, dinv, x, y, n_nodes); break;
      case 3: pdl_launch(grid, block, 0, s)(k_block_jacobi_apply<3>, dinv, x, y, n_nodes); break;
      default: pdl_launch(grid, block, 0, s)(k_block_jacobi_apply<4>, dinv, x, y, n_nodes); break;
    }
  }
  // out_dev[j] = V[j] . w
  void multi_dot(const double* V, long long ld, const double* w, int nv, double* partial,
                 double* out_dev) const {
    const int gx = grid_for(n, DOT_BLOCK, sms) > 256 ? 256 : grid_for(n, DOT_BLOCK, sms);
    dim3 grid(gx, nv < 64 ? nv : 64);
    pdl_launch(grid, DOT_BLOCK, 0, s)(k_multi_dot_partial, V, ld, w, nv, n, partial);
    pdl_launch(nv, 256, 0, s)(k_multi_dot_final, partial, gx, nv, out_dev);
    allreduce(out_dev, nv);
  }
};

}  // namespace c8
