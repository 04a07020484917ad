// Aggregation multigrid preconditioner (see amg.cuh).
#include "amg.cuh"
#include "comm.cuh"

#include <algorithm>
#include <cstdlib>
#include <cstdint>
#include <numeric>

namespace c8 {

// r = b - A x ; one thread per scalar row.  F = storage type of the matrix values: the fine level of
// the preconditioner reads an fp32 copy (half the HBM traffic of the pass; the smoother then works
// with a fixed, slightly perturbed linear operator, the Krylov residuals stay fp64)
template <int NB, class F>
__global__ void k_bsr_residual(const int* __restrict__ rowptr, const int* __restrict__ colind,
                               const F* __restrict__ vals, const double* __restrict__ x,
                               const double* __restrict__ b, double* __restrict__ r, int n_nodes) {
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_nodes * NB) return;
  const int node = i / NB, row = i % NB;
  double s = b[i];
  const int b0 = rowptr[node], b1 = rowptr[node + 1];
#pragma unroll 4
  for (int k = b0; k < b1; ++k) {
    double a[NB], xv[NB];
    ld_row<NB, F>(vals + (size_t(k) * NB + row) * NB, a);
    ld_row<NB, double>(x + size_t(__ldg(&colind[k])) * NB, xv);
#pragma unroll
    for (int c = 0; c < NB; ++c) s = fma(-a[c], xv[c], s);
  }
  r[i] = s;
}

// One damped block-Jacobi sweep, out of place: xout = x' + omega Dinv (b - A x'), with
// x' = xin (+ pscale * P xc when PROLONG: the coarse-grid correction is folded into the sweep).
// Four lanes per node (lane r < NB owns row r); the node's residual rows meet through shuffles.
template <int NB, class F, bool PROLONG>
__global__ void k_smooth(const int* __restrict__ rowptr, const int* __restrict__ colind,
                         const F* __restrict__ vals, const double* __restrict__ dinv,
                         const double* __restrict__ b, const double* __restrict__ xin,
                         double* __restrict__ xout, const int* __restrict__ agg,
                         const double* __restrict__ xc, double pscale, double omega, int n, int npc,
                         const double* xprev = nullptr, double c1 = 0.0) {
  pdl_wait();
  const int gid = blockIdx.x * blockDim.x + threadIdx.x;
  int node = gid >> 2;
  const int lane4 = gid & 3;
  const bool valid = node < n;
  if (!valid) node = n - 1;            // keep every lane for the shuffles
  const bool active = lane4 < NB;
  const int row = active ? lane4 : 0;
  double s = b[size_t(node) * NB + row];
  const int b0 = rowptr[node], b1 = rowptr[node + 1];
#pragma unroll 4
  for (int k = b0; k < b1; ++k) {
    const int col = __ldg(&colind[k]);
    double a[NB], xv[NB];
    ld_row<NB, F>(vals + (size_t(k) * NB + row) * NB, a);
    ld_row<NB, double>(xin + size_t(col) * NB, xv);
    // columns >= npc carry no coarse correction (the ghost columns of a part whose hierarchy acts
    // on its owned block only; npc = all columns when the hierarchy spans the parts); the index is
    // clamped rather than branched around because the compiler may speculate a read-only load
    const double ps = (PROLONG && col < npc) ? pscale : 0.0;
    const double* pc = PROLONG ? xc + size_t(agg[col < npc ? col : 0]) * NB : nullptr;
    if (PROLONG) {
      double pv[NB];
      ld_row<NB, double>(pc, pv);
#pragma unroll
      for (int c = 0; c < NB; ++c) xv[c] = fma(ps, pv[c], xv[c]);
    }
#pragma unroll
    for (int c = 0; c < NB; ++c) s = fma(-a[c], xv[c], s);
  }
  const int base = (threadIdx.x & 31) & ~3;
  double upd = 0.0;
#pragma unroll
  for (int c = 0; c < NB; ++c) {
    const double sc = __shfl_sync(0xffffffffu, s, base + c);
    upd = fma(dinv[(size_t(node) * NB + row) * NB + c], sc, upd);
  }
  double xo = xin[size_t(node) * NB + row];
  if (PROLONG) xo = fma(pscale, xc[size_t(agg[node]) * NB + row], xo);
  // second step of a Chebyshev pair: x2 = x1 + c1 (x1 - x0) + c2 Dinv r1 (xprev = x0, may alias xout: a thread
  // reads only its own entry of it, before it writes)
  if (c1 != 0.0) {
    const double xp = xprev ? xprev[size_t(node) * NB + row] : 0.0;
    xo = fma(c1, xo - xp, xo);
  }
  if (valid && active) xout[size_t(node) * NB + row] = xo + omega * upd;
}

// Coarse levels: rows are long (the coarse graphs fill in) and few, so one thread per scalar row is
// a serial chain of dependent gathers (measured ~17 us per launch whatever the level size).  Here a
// WARP owns a node: the lanes split the row's blocks, partial sums meet in a shuffle reduction.
// RESID_ONLY: r = b - A x' instead of the Jacobi update.
template <int NB, bool PROLONG, bool RESID_ONLY>
__global__ void k_smooth_warp(const int* __restrict__ rowptr, const int* __restrict__ colind,
                              const double* __restrict__ vals, const double* __restrict__ dinv,
                              const double* __restrict__ b, const double* __restrict__ xin,
                              double* __restrict__ xout, const int* __restrict__ agg,
                              const double* __restrict__ xc, double pscale, double omega, int n) {
  pdl_wait();
  const int node = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (node >= n) return;  // the whole warp leaves together
  double s[NB];
#pragma unroll
  for (int r = 0; r < NB; ++r) s[r] = 0.0;
  for (int k = rowptr[node] + lane; k < rowptr[node + 1]; k += 32) {
    const int col = colind[k];
    double xv[NB];
#pragma unroll
    for (int c = 0; c < NB; ++c) xv[c] = xin[size_t(col) * NB + c];
    if (PROLONG) {
      const double* pc = xc + size_t(agg[col]) * NB;
#pragma unroll
      for (int c = 0; c < NB; ++c) xv[c] = fma(pscale, pc[c], xv[c]);
    }
    const double* a = vals + size_t(k) * NB * NB;
#pragma unroll
    for (int r = 0; r < NB; ++r)
#pragma unroll
      for (int c = 0; c < NB; ++c) s[r] = fma(-a[r * NB + c], xv[c], s[r]);
  }
#pragma unroll
  for (int r = 0; r < NB; ++r) {
    for (int o = 16; o > 0; o >>= 1) s[r] += __shfl_down_sync(0xffffffffu, s[r], o);
    s[r] = __shfl_sync(0xffffffffu, s[r], 0) + b[size_t(node) * NB + r];
  }
  if (lane < NB) {
    if (RESID_ONLY) {
      double v = s[0];
#pragma unroll
      for (int r = 1; r < NB; ++r) v = (lane == r) ? s[r] : v;
      xout[size_t(node) * NB + lane] = v;
    } else {
      double upd = 0.0;
#pragma unroll
      for (int c = 0; c < NB; ++c) upd = fma(dinv[(size_t(node) * NB + lane) * NB + c], s[c], upd);
      double xo = xin[size_t(node) * NB + lane];
      if (PROLONG) xo = fma(pscale, xc[size_t(agg[node]) * NB + lane], xo);
      xout[size_t(node) * NB + lane] = xo + omega * upd;
    }
  }
}

// NB = 4, a warp per node with COALESCED block reads: the values of a node's row are contiguous in
// memory, so the 32 lanes read 512 consecutive bytes per step -- a lane owns one 16-byte piece of one
// block (fp64: 8 lanes per block, 4 blocks per step, the piece is half a row; fp32: 4 lanes per
// block, 8 blocks per step, the piece is a whole row) and multiplies it with the matching entries
// of x[col].  Lanes holding pieces of the same block row meet in 3 xor-shuffles.  The
// thread-per-row kernels above touch 8 separate segments per load instruction and the earlier
// warp-per-node kernel 128-byte-strided scalars (16 LDG.64 per block and lane); this one issues one
// LDG.128 per 16 bytes of matrix.  RESID_ONLY: xout = b - A x'; else one damped block-Jacobi sweep.
template <class F, bool PROLONG, bool RESID_ONLY>
__global__ void k_smooth_cw4(const int* __restrict__ rowptr, const int* __restrict__ colind,
                             const F* __restrict__ vals, const double* __restrict__ dinv,
                             const double* __restrict__ b, const double* __restrict__ xin,
                             double* __restrict__ xout, const int* __restrict__ agg,
                             const double* __restrict__ xc, double pscale, double omega, int n, int npc) {
  pdl_wait();
  constexpr int LPB = sizeof(F) == 8 ? 8 : 4;   // lanes per block
  constexpr int BPS = 32 / LPB;                 // blocks per step
  const int node = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (node >= n) return;  // the whole warp leaves together
  const int j = lane / LPB, p = lane % LPB;
  double s = 0.0;
  const int b1 = rowptr[node + 1];
  for (int k = rowptr[node] + j; k < b1; k += BPS) {
    const int col = __ldg(&colind[k]);
    double ps = 0.0;
    const double* pc = nullptr;
    if (PROLONG) { ps = col < npc ? pscale : 0.0; pc = xc + size_t(agg[col < npc ? col : 0]) * 4; }
    if constexpr (sizeof(F) == 8) {
      const int c0 = (p & 1) * 2;
      const double2 a = __ldg(reinterpret_cast<const double2*>(vals + size_t(k) * 16) + p);
      double2 xv = __ldg(reinterpret_cast<const double2*>(xin + size_t(col) * 4 + c0));
      if (PROLONG) {
        const double2 q = __ldg(reinterpret_cast<const double2*>(pc + c0));
        xv.x = fma(ps, q.x, xv.x); xv.y = fma(ps, q.y, xv.y);
      }
      s = fma(-double(a.x), xv.x, s); s = fma(-double(a.y), xv.y, s);
    } else {
      const float4 a = __ldg(reinterpret_cast<const float4*>(vals + size_t(k) * 16) + p);
      double2 x0 = __ldg(reinterpret_cast<const double2*>(xin + size_t(col) * 4));
      double2 x1 = __ldg(reinterpret_cast<const double2*>(xin + size_t(col) * 4) + 1);
      if (PROLONG) {
        const double2 q0 = __ldg(reinterpret_cast<const double2*>(pc));
        const double2 q1 = __ldg(reinterpret_cast<const double2*>(pc) + 1);
        x0.x = fma(ps, q0.x, x0.x); x0.y = fma(ps, q0.y, x0.y);
        x1.x = fma(ps, q1.x, x1.x); x1.y = fma(ps, q1.y, x1.y);
      }
      s = fma(-double(a.x), x0.x, s); s = fma(-double(a.y), x0.y, s);
      s = fma(-double(a.z), x1.x, s); s = fma(-double(a.w), x1.y, s);
    }
  }
  // block row of a lane: fp64 p >> 1 (pieces 2r, 2r+1), fp32 p
  if constexpr (sizeof(F) == 8) {
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 8);
    s += __shfl_xor_sync(0xffffffffu, s, 16);
  } else {
    s += __shfl_xor_sync(0xffffffffu, s, 4);
    s += __shfl_xor_sync(0xffffffffu, s, 8);
    s += __shfl_xor_sync(0xffffffffu, s, 16);
  }
  const int row = lane & 3;
  const int src = sizeof(F) == 8 ? row * 2 : row;   // a lane that holds the sum of block row `row`
  const double res = __shfl_sync(0xffffffffu, s, src) + b[size_t(node) * 4 + row];
  if (RESID_ONLY) {
    if (lane < 4) xout[size_t(node) * 4 + row] = res;
    return;
  }
  double upd = 0.0;
#pragma unroll
  for (int c = 0; c < 4; ++c)
    upd = fma(dinv[(size_t(node) * 4 + row) * 4 + c], __shfl_sync(0xffffffffu, res, c), upd);
  if (lane < 4) {
    double xo = xin[size_t(node) * 4 + row];
    if (PROLONG) xo = fma(pscale, xc[size_t(agg[node]) * 4 + row], xo);
    xout[size_t(node) * 4 + row] = xo + omega * upd;
  }
}

__global__ void k_to_float(const double* __restrict__ in, float* __restrict__ out, size_t n) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) out[i] = float(in[i]);
}

// x = (zero_guess ? 0 : x) + omega * Dinv r
template <int NB>
__global__ void k_jacobi_update(const double* __restrict__ dinv, const double* __restrict__ r,
                                double* __restrict__ x, double omega, int zero_guess, int n_nodes) {
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_nodes * NB) return;
  const int node = i / NB, row = i % NB;
  double s = 0.0;
#pragma unroll
  for (int c = 0; c < NB; ++c) s = fma(dinv[(size_t(node) * NB + row) * NB + c], r[size_t(node) * NB + c], s);
  x[i] = (zero_guess ? 0.0 : x[i]) + omega * s;
}

// bc[I] = sum over the members of aggregate I
template <int NB>
__global__ void k_restrict(const int* __restrict__ aggptr, const int* __restrict__ aggmem,
                           const double* __restrict__ r, double* __restrict__ bc, int nc) {
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nc * NB) return;
  const int I = i / NB, c = i % NB;
  double s = 0.0;
  for (int k = aggptr[I]; k < aggptr[I + 1]; ++k) s += r[size_t(aggmem[k]) * NB + c];
  bc[i] = s;
}

template <int NB>
__global__ void k_prolong_add(const int* __restrict__ agg, const double* __restrict__ xc,
                              double* __restrict__ x, double scale, int n) {
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * NB) return;
  const int node = i / NB, c = i % NB;
  x[i] += scale * xc[size_t(agg[node]) * NB + c];
}

// coarse block values = sum of the fine blocks mapped to it; one thread per (coarse block, entry)
template <int NB>
__global__ void k_galerkin(const int* __restrict__ cptr, const int* __restrict__ cmem,
                           const double* __restrict__ fine, double* __restrict__ coarse, int nnzb_c) {
  constexpr int BB = NB * NB;
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= (long long)nnzb_c * BB) return;
  const int blk = int(i / BB), e = int(i % BB);
  double s = 0.0;
  for (int k = cptr[blk]; k < cptr[blk + 1]; ++k) s += fine[size_t(cmem[k]) * BB + e];
  coarse[i] = s;
}

// dense [N][2N] = [A | I] from the coarsest BSR operator
template <int NB>
__global__ void k_dense_fill(const int* __restrict__ rowptr, const int* __restrict__ colind,
                             const double* __restrict__ vals, double* __restrict__ M, int n_nodes) {
  const int N = n_nodes * NB;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const int node = i / NB, row = i % NB;
  double* Mi = M + size_t(i) * 2 * N;
  for (int c = 0; c < 2 * N; ++c) Mi[c] = (c == N + i) ? 1.0 : 0.0;
  for (int k = rowptr[node]; k < rowptr[node + 1]; ++k)
#pragma unroll
    for (int c = 0; c < NB; ++c) Mi[colind[k] * NB + c] = vals[(size_t(k) * NB + row) * NB + c];
}

// Gauss-Jordan inverse with partial pivoting, one CTA, matrix in global memory (L2 resident)
__global__ void __launch_bounds__(1024) k_dense_inverse(double* __restrict__ M, int N) {
  extern __shared__ double colk[];  // N doubles
  __shared__ double s_best[32];
  __shared__ int s_idx[32];
  __shared__ int s_p;
  const int tid = threadIdx.x, nt = blockDim.x;
  const int W = 2 * N;
  for (int k = 0; k < N; ++k) {
    // pivot search
    double best = -1.0;
    int bi = k;
    for (int i = k + tid; i < N; i += nt) {
      const double a = fabs(M[size_t(i) * W + k]);
      if (a > best) { best = a; bi = i; }
    }
    for (int o = 16; o > 0; o >>= 1) {
      const double ob = __shfl_down_sync(0xffffffffu, best, o);
      const int oi = __shfl_down_sync(0xffffffffu, bi, o);
      if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
    }
    if ((tid & 31) == 0) { s_best[tid >> 5] = best; s_idx[tid >> 5] = bi; }
    __syncthreads();
    if (tid == 0) {
      double b = s_best[0];
      int p = s_idx[0];
      for (int w = 1; w < (nt >> 5); ++w)
        if (s_best[w] > b || (s_best[w] == b && s_idx[w] < p)) { b = s_best[w]; p = s_idx[w]; }
      s_p = p;
      if (!(b > 0.0)) { M[size_t(k) * W + k] = 1.0; s_p = k; }  // empty row: identity
    }
    __syncthreads();
    const int p = s_p;
    if (p != k)
      for (int c = tid; c < W; c += nt) {
        const double t = M[size_t(k) * W + c];
        M[size_t(k) * W + c] = M[size_t(p) * W + c];
        M[size_t(p) * W + c] = t;
      }
    __syncthreads();
    const double inv = 1.0 / M[size_t(k) * W + k];
    for (int i = tid; i < N; i += nt) colk[i] = M[size_t(i) * W + k];
    __syncthreads();
    for (int c = tid; c < W; c += nt) M[size_t(k) * W + c] *= inv;
    __syncthreads();
    // eliminate column k from every other row; columns < k of the left half are already zero
    const int c0 = k, ncol = W - c0;
    const long long work = (long long)N * ncol;
    for (long long q = tid; q < work; q += nt) {
      const int i = int(q / ncol), c = c0 + int(q % ncol);
      if (i == k) continue;
      const double f = colk[i];
      if (f != 0.0) M[size_t(i) * W + c] = fma(-f, M[size_t(k) * W + c], M[size_t(i) * W + c]);
    }
    __syncthreads();
  }
}

// x = Ainv b with Ainv = right half of M
__global__ void k_dense_apply(const double* __restrict__ M, const double* __restrict__ b,
                              double* __restrict__ x, int N) {
  pdl_wait();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= N) return;
  const double* row = M + size_t(warp) * 2 * N + N;
  double s = 0.0;
  for (int c = lane; c < N; c += 32) s = fma(row[c], b[c], s);
  for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
  if (lane == 0) x[warp] = s;
}

// C8_AMG_DEBUG=1: synchronise and report after every stage of the cycle (bad-access hunting)
static void amg_dbg(cudaStream_t s, const char* what, int level) {
  static const bool on = getenv("C8_AMG_DEBUG") != nullptr;
  if (!on) return;
  const cudaError_t e = cudaStreamSynchronize(s);
  if (e != cudaSuccess) {
    fprintf(stderr, "[amg] %s at level %d: %s\n", what, level, cudaGetErrorString(e));
    fflush(stderr);
    abort();
  }
}

// experiment switches (defaults = the measured best)
static bool env_flag(const char* name, bool dflt) {
  const char* e = getenv(name);
  return e ? e[0] != '0' : dflt;
}
static bool cw_coarse() { static const bool v = env_flag("C8_CW_COARSE", true); return v; }   // coalesced warp kernel on levels >= 1
static bool cw_fine() { static const bool v = env_flag("C8_CW_FINE", false); return v; }      // ... and on the fine level
static bool fine_prolong_add() { static const bool v = env_flag("C8_FINE_PROLONG_ADD", true); return v; }
// fine-level Chebyshev smoothing: upper bound of the spectrum of Dinv A (0: damped Jacobi) and hi / lo ratio
static double env_num(const char* name, double dflt) { const char* e = getenv(name); return e ? atof(e) : dflt; }
static double cheb_lmax() { static const double v = env_num("C8_CHEB_LMAX", 0.0); return v; }
static double cheb_ratio() { static const double v = env_num("C8_CHEB_RATIO", 4.0); return v; }

#define C8_NB_SWITCH(nb, CALL)  \
  switch (nb) {                 \
    case 2: { constexpr int NB = 2; CALL; } break; \
    case 3: { constexpr int NB = 3; CALL; } break; \
    default: { constexpr int NB = 4; CALL; } break; \
  }

template <class T>
static int upload(c8_ctx* ctx, const std::vector<T>& h, T** d) {
  *d = nullptr;
  if (h.empty()) return C8_OK;
  C8_CUDA(ctx, cudaMalloc(d, h.size() * sizeof(T)));
  C8_CUDA(ctx, cudaMemcpy(*d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
  return C8_OK;
}

Amg::~Amg() {
  for (AmgLevel& L : lv_) {
    void* ptrs[] = {L.own_rowptr, L.own_colind, L.own_vals, L.dinv, L.x, L.b, L.r, L.agg, L.aggptr,
                    L.aggmem, L.cptr, L.cmem, L.xt, L.vals32};
    for (void* p : ptrs)
      if (p) cudaFree(p);
  }
  if (dense_) cudaFree(dense_);
  if (r0_) cudaFree(r0_);
}

// the build's collectives on top of the context's transport: host vectors staged through a device buffer
struct DeviceCollectives : AmgCollectives {
  c8_ctx* ctx = nullptr;
  double* d = nullptr;
  size_t cap = 0;
  bool ok = true;
  ~DeviceCollectives() override { if (d) cudaFree(d); }
  double* buf(size_t n) {
    if (n > cap) {
      if (d) cudaFree(d);
      d = nullptr; cap = 0;
      if (cudaMalloc(&d, (n + 1024) * sizeof(double)) != cudaSuccess) { ok = false; return nullptr; }
      cap = n + 1024;
    }
    return d;
  }
  template <class F>
  void round_trip(std::vector<double>& v, F&& op) {
    if (v.empty()) return;   // the same on every part
    double* p = buf(v.size());
    if (!p) return;
    cudaStream_t s = ctx->stream;
    cudaMemcpyAsync(p, v.data(), v.size() * sizeof(double), cudaMemcpyHostToDevice, s);
    op(p);
    cudaMemcpyAsync(v.data(), p, v.size() * sizeof(double), cudaMemcpyDeviceToHost, s);
    if (cudaStreamSynchronize(s) != cudaSuccess) ok = false;
  }
  void allreduce(std::vector<double>& v) override {
    round_trip(v, [&](double* p) { ctx->allreduce_cb(ctx->comm_user, p, int(v.size())); });
  }
  void halo(int level, std::vector<double>& v) override {
    round_trip(v, [&](double* p) { comm_halo_level(ctx, level, p, 1); });
  }
  int add_level(const HaloPlanHost& plan) override {
    const int id = comm_add_level(ctx, plan);
    if (id < 0) ok = false;
    return id;
  }
  HaloPlanHost plan(int level) const override { return comm_plan(ctx, level); }
};

int Amg::build() {
  nb_ = ctx_->kt->nb;
  const int n0 = ctx_->n_owned_nodes;
  // the hierarchy spans the parts when one of the library's transports is bound (amg_host.hpp);
  // with the caller's own hooks it acts on the owned x owned block of the part
  dist_ = opt.distributed && comm_library_transport(ctx_);
  comm_drop_levels(ctx_);
  DeviceCollectives dc;
  dc.ctx = ctx_; dc.rank = comm_rank(ctx_); dc.nranks = comm_nranks(ctx_);
  AmgBuildOptions bo;
  bo.coarsest_max_nodes = opt.coarsest_max_nodes; bo.max_levels = opt.max_levels;
  bo.max_aggregate_size = opt.max_aggregate_size; bo.coarse_aggregate_size = opt.coarse_aggregate_size;
  bo.replicate_max_nodes = opt.replicate_max_nodes;
  std::vector<AmgLevelHost> hl;
  amg_build_host(n0, ctx_->n_nodes, ctx_->h_rowptr.data(), ctx_->h_colind.data(), dist_ ? &dc : nullptr, 0, bo, hl);
  if (!dc.ok) return fail(ctx_, C8_ERR_CUDA, "multigrid build: a collective of the hierarchy build failed");
  lv_.clear();
  lv_.resize(hl.size());
  int rc;
  for (size_t l = 0; l < hl.size(); ++l) {
    AmgLevel& L = lv_[l];
    const AmgLevelHost& H = hl[l];
    L.n = H.n; L.halo_level = H.halo_level;
    if (l == 0) {
      L.ld = ctx_->n_nodes; L.nnzb = ctx_->nnzb;
      L.rowptr = ctx_->d_rowptr; L.colind = ctx_->d_colind;
    } else {
      L.ld = H.ld; L.nnzb = H.nnzb;
      if ((rc = upload(ctx_, H.rowptr, &L.own_rowptr)) != C8_OK) return rc;
      if ((rc = upload(ctx_, H.colind, &L.own_colind)) != C8_OK) return rc;
      L.rowptr = L.own_rowptr; L.colind = L.own_colind;
      const size_t nv = size_t(L.ld > 0 ? L.ld : 1) * nb_;
      C8_CUDA(ctx_, cudaMalloc(&L.own_vals, size_t(L.nnzb > 0 ? L.nnzb : 1) * nb_ * nb_ * sizeof(double)));
      L.vals = L.own_vals;
      C8_CUDA(ctx_, cudaMalloc(&L.x, nv * sizeof(double)));
      C8_CUDA(ctx_, cudaMalloc(&L.b, nv * sizeof(double)));
      C8_CUDA(ctx_, cudaMalloc(&L.r, nv * sizeof(double)));
      C8_CUDA(ctx_, cudaMemsetAsync(L.x, 0, nv * sizeof(double), ctx_->stream));
      C8_CUDA(ctx_, cudaMemsetAsync(L.b, 0, nv * sizeof(double), ctx_->stream));
      C8_CUDA(ctx_, cudaMemsetAsync(L.r, 0, nv * sizeof(double), ctx_->stream));
    }
    if (l + 1 < hl.size()) {
      L.nc = H.nc_rows;
      L.npc = int(H.agg.size());
      L.coarse_replicated = H.coarse_replicated;
      L.n_cmem = int(H.cmem.size());
      if ((rc = upload(ctx_, H.agg, &L.agg)) != C8_OK) return rc;
      if ((rc = upload(ctx_, H.aggptr, &L.aggptr)) != C8_OK) return rc;
      if ((rc = upload(ctx_, H.aggmem, &L.aggmem)) != C8_OK) return rc;
      if ((rc = upload(ctx_, H.cptr, &L.cptr)) != C8_OK) return rc;
      if ((rc = upload(ctx_, H.cmem, &L.cmem)) != C8_OK) return rc;
    }
    C8_CUDA(ctx_, cudaMalloc(&L.dinv, size_t(L.n > 0 ? L.n : 1) * nb_ * nb_ * sizeof(double)));
    const size_t nx = size_t(L.ld > 0 ? L.ld : 1) * nb_;
    C8_CUDA(ctx_, cudaMalloc(&L.xt, nx * sizeof(double)));
    C8_CUDA(ctx_, cudaMemsetAsync(L.xt, 0, nx * sizeof(double), ctx_->stream));
  }
  if (opt.fp32_fine_level && lv_.size() > 1)
    C8_CUDA(ctx_, cudaMalloc(&lv_[0].vals32, size_t(lv_[0].nnzb) * nb_ * nb_ * sizeof(float)));
  C8_CUDA(ctx_, cudaMalloc(&r0_, size_t(ctx_->n_nodes) * nb_ * sizeof(double)));
  C8_CUDA(ctx_, cudaMemsetAsync(r0_, 0, size_t(ctx_->n_nodes) * nb_ * sizeof(double), ctx_->stream));
  nd_ = 0;
  if (lv_.size() > 1 && lv_.back().halo_level < 0 && lv_.back().n > 0 && lv_.back().n * nb_ <= 1024) {
    nd_ = lv_.back().n * nb_;
    C8_CUDA(ctx_, cudaMalloc(&dense_, size_t(nd_) * 2 * nd_ * sizeof(double)));
  }
  return C8_OK;
}

double Amg::operator_complexity() const {
  double s = 0.0;
  for (const AmgLevel& L : lv_) s += L.nnzb;
  return lv_.empty() ? 0.0 : s / lv_[0].nnzb;
}

int Amg::setup(const double* A) {
  cudaStream_t s = ctx_->stream;
  lv_[0].vals = A;
  if (lv_[0].vals32)
    k_to_float<<<148 * 8, 256, 0, s>>>(A, lv_[0].vals32, size_t(lv_[0].nnzb) * nb_ * nb_);
  for (size_t l = 0; l < lv_.size(); ++l) {
    AmgLevel& L = lv_[l];
    if (L.n > 0) {
      const int gj = (L.n + 127) / 128;
      C8_NB_SWITCH(nb_, (k_block_jacobi_setup<NB><<<gj, 128, 0, s>>>(L.rowptr, L.colind, L.vals, L.dinv, L.n)));
    }
    if (l + 1 < lv_.size()) {
      AmgLevel& C = lv_[l + 1];
      const long long work = (long long)C.nnzb * nb_ * nb_;
      if (work > 0) {
        const unsigned g = unsigned((work + 255) / 256);
        C8_NB_SWITCH(nb_, (k_galerkin<NB><<<g, 256, 0, s>>>(L.cptr, L.cmem, L.vals, C.own_vals, C.nnzb)));
        // a replicated level: every part computed the blocks of its own rows (zeros elsewhere)
        if (L.coarse_replicated) ctx_->allreduce_cb(ctx_->comm_user, C.own_vals, int(work));
      }
    }
  }
  if (nd_ > 0) {
    const AmgLevel& C = lv_.back();
    C8_NB_SWITCH(nb_, (k_dense_fill<NB><<<(nd_ + 127) / 128, 128, 0, s>>>(C.rowptr, C.colind, C.vals, dense_, C.n)));
    k_dense_inverse<<<1, 1024, nd_ * sizeof(double), s>>>(dense_, nd_);
  }
  C8_CUDA(ctx_, cudaGetLastError());
  return C8_OK;
}

// fills the ghost entries of a level's vector (no-op on serial / replicated levels)
void Amg::halo(int l, const double* v) {
  if (lv_[l].halo_level >= 0) comm_halo_level(ctx_, lv_[l].halo_level, const_cast<double*>(v), nb_);
}

void Amg::smooth(int l, const double* b, double* x, int sweeps, bool zero_guess) {
  AmgLevel& L = lv_[l];
  cudaStream_t s = ctx_->stream;
  double* r = (l == 0) ? r0_ : L.r;
  const int g = (L.n * nb_ + 127) / 128;
  for (int k = 0; k < sweeps; ++k) {
    if (k == 0 && zero_guess) {
      if (L.n > 0) { C8_NB_SWITCH(nb_, (pdl_launch(g, 128, 0, s)(k_jacobi_update<NB>, L.dinv, b, x, opt.omega, 1, L.n))); }
    } else {
      halo(l, x);
      if (L.n == 0) continue;
      C8_NB_SWITCH(nb_, (pdl_launch(g, 128, 0, s)(k_bsr_residual<NB, double>, L.rowptr, L.colind, L.vals, x, b, r, L.n)));
      C8_NB_SWITCH(nb_, (pdl_launch(g, 128, 0, s)(k_jacobi_update<NB>, L.dinv, r, x, opt.omega, 0, L.n)));
    }
  }
}

// one out-of-place sweep on level l (fp32 matrix copy on the fine level when present); the ghost
// entries of xin (and of xc) must be current
void Amg::sweep(int l, const double* b, const double* xin, double* xout, const double* xc, const double* xprev,
                double c1, double c2) {
  AmgLevel& L = lv_[l];
  if (L.n == 0) return;
  cudaStream_t s = ctx_->stream;
  const int g = (L.n * 4 + 127) / 128;
  const double oc = opt.over_correction, om = (c2 != 0.0) ? c2 : opt.omega;
  const int gw = (L.n * 32 + 127) / 128;
  if (nb_ == 4 && (l > 0 ? cw_coarse() : cw_fine())) {  // a warp per node, coalesced block reads
    if (l == 0 && L.vals32) {
      if (xc) pdl_launch(gw, 128, 0, s)(k_smooth_cw4<float, true, false>, L.rowptr, L.colind, L.vals32, L.dinv, b, xin, xout, L.agg, xc, oc, om, L.n, L.npc);
      else pdl_launch(gw, 128, 0, s)(k_smooth_cw4<float, false, false>, L.rowptr, L.colind, L.vals32, L.dinv, b, xin, xout, nullptr, nullptr, 0.0, om, L.n, 0);
    } else {
      if (xc) pdl_launch(gw, 128, 0, s)(k_smooth_cw4<double, true, false>, L.rowptr, L.colind, L.vals, L.dinv, b, xin, xout, L.agg, xc, oc, om, L.n, L.npc);
      else pdl_launch(gw, 128, 0, s)(k_smooth_cw4<double, false, false>, L.rowptr, L.colind, L.vals, L.dinv, b, xin, xout, nullptr, nullptr, 0.0, om, L.n, 0);
    }
    return;
  }
  if (l > 0) {  // coarse levels: a warp per node
    if (xc) { C8_NB_SWITCH(nb_, (pdl_launch(gw, 128, 0, s)(k_smooth_warp<NB, true, false>, L.rowptr, L.colind, L.vals, L.dinv, b, xin, xout, L.agg, xc, oc, om, L.n))); }
    else { C8_NB_SWITCH(nb_, (pdl_launch(gw, 128, 0, s)(k_smooth_warp<NB, false, false>, L.rowptr, L.colind, L.vals, L.dinv, b, xin, xout, nullptr, nullptr, 0.0, om, L.n))); }
    return;
  }
  if (L.vals32) {
    if (xc) { C8_NB_SWITCH(nb_, (pdl_launch(g, 128, 0, s)(k_smooth<NB, float, true>, L.rowptr, L.colind, L.vals32, L.dinv, b, xin, xout, L.agg, xc, oc, om, L.n, L.npc, xprev, c1))); }
    else { C8_NB_SWITCH(nb_, (pdl_launch(g, 128, 0, s)(k_smooth<NB, float, false>, L.rowptr, L.colind, L.vals32, L.dinv, b, xin, xout, (const int*)nullptr, (const double*)nullptr, 0.0, om, L.n, 0, xprev, c1))); }
  } else {
    if (xc) { C8_NB_SWITCH(nb_, (pdl_launch(g, 128, 0, s)(k_smooth<NB, double, true>, L.rowptr, L.colind, L.vals, L.dinv, b, xin, xout, L.agg, xc, oc, om, L.n, L.npc, xprev, c1))); }
    else { C8_NB_SWITCH(nb_, (pdl_launch(g, 128, 0, s)(k_smooth<NB, double, false>, L.rowptr, L.colind, L.vals, L.dinv, b, xin, xout, (const int*)nullptr, (const double*)nullptr, 0.0, om, L.n, 0, xprev, c1))); }
  }
}

// V(nu_pre, nu_post) cycle; the result lands in xout.  Launches per level: nu_pre + nu_post sweeps
// (the first from a zero guess is a block-diagonal product, the first after the coarse solve carries
// the prolongation) + residual + restriction.  On a partitioned level every sweep / residual is
// preceded by the halo copy of its input vector; the restricted right-hand side of a replicated
// coarse level is summed over the parts (each part fills the rows of its own aggregates).
void Amg::cycle(int l, const double* b, double* xout) {
  AmgLevel& L = lv_[l];
  cudaStream_t s = ctx_->stream;
  const bool part = L.halo_level >= 0;
  if (L.n == 0 && !part) return;
  const bool coarsest = (l + 1 == int(lv_.size()));
  if (coarsest) {
    if (nd_ > 0 && l > 0) {
      pdl_launch((nd_ * 32 + 255) / 256, 256, 0, s)(k_dense_apply, dense_, b, xout, nd_);
    } else {
      smooth(l, b, xout, 4 * (opt.nu_pre + opt.nu_post), true);
    }
    return;
  }
  int nu1 = opt.nu_pre < 1 ? 1 : opt.nu_pre, nu2 = opt.nu_post < 1 ? 1 : opt.nu_post;
  if (l > 0 && opt.coarse_nu > 0) nu1 = nu2 = opt.coarse_nu;  // cheaper cycle below the fine level
  const int writes = nu1 + nu2;
  double* bufs[2] = {xout, L.xt};
  auto buf = [&](int w) { return bufs[(writes - 1 - w) & 1]; };  // the last write goes to xout
  const int g = (L.n * nb_ + 127) / 128;
  amg_dbg(s, "enter", l);
  // Fine level, V(2,2) only: the two sweeps of a side as ONE Chebyshev polynomial of degree 2 in Dinv A on
  // [lmax / ratio, lmax] (same kernels and traffic as two damped Jacobi sweeps):
  //   x1 = x0 + 1/theta Dinv r0,  x2 = x1 + rho1 rho0 (x1 - x0) + 2 rho1 / delta Dinv r1
  const bool cheb = (l == 0) && cheb_lmax() > 0.0 && nu1 == 2 && nu2 == 2 && !(nb_ == 4 && cw_fine()) &&
                    fine_prolong_add();
  double ch_a0 = 0.0, ch_c1 = 0.0, ch_c2 = 0.0;
  if (cheb) {
    const double hi = cheb_lmax(), lo = hi / cheb_ratio();
    const double theta = 0.5 * (hi + lo), delta = 0.5 * (hi - lo), sigma = theta / delta;
    const double rho0 = 1.0 / sigma, rho1 = 1.0 / (2.0 * sigma - rho0);
    ch_a0 = 1.0 / theta; ch_c1 = rho1 * rho0; ch_c2 = 2.0 * rho1 / delta;
  }
  if (L.n > 0) { C8_NB_SWITCH(nb_, (pdl_launch(g, 128, 0, s)(k_jacobi_update<NB>, L.dinv, b, buf(0), cheb ? ch_a0 : opt.omega, 1, L.n))); }
  amg_dbg(s, "jacobi zero-guess", l);
  for (int w = 1; w < nu1; ++w) {
    halo(l, buf(w - 1));
    if (cheb) sweep(l, b, buf(w - 1), buf(w), nullptr, nullptr, ch_c1, ch_c2);   // x0 = 0
    else sweep(l, b, buf(w - 1), buf(w), nullptr);
    amg_dbg(s, "pre sweep", l);
  }
  const double* cur = buf(nu1 - 1);
  halo(l, cur);
  double* r = (l == 0) ? r0_ : L.r;
  if (L.n > 0) {
    if (nb_ == 4 && (l > 0 ? cw_coarse() : cw_fine())) {
      const int gw = (L.n * 32 + 127) / 128;
      if (l == 0 && L.vals32) pdl_launch(gw, 128, 0, s)(k_smooth_cw4<float, false, true>, L.rowptr, L.colind, L.vals32, nullptr, b, cur, r, nullptr, nullptr, 0.0, 0.0, L.n, 0);
      else pdl_launch(gw, 128, 0, s)(k_smooth_cw4<double, false, true>, L.rowptr, L.colind, L.vals, nullptr, b, cur, r, nullptr, nullptr, 0.0, 0.0, L.n, 0);
    }
    else if (l > 0) { C8_NB_SWITCH(nb_, (pdl_launch((L.n * 32 + 127) / 128, 128, 0, s)(k_smooth_warp<NB, false, true>, L.rowptr, L.colind, L.vals, nullptr, b, cur, r, nullptr, nullptr, 0.0, 0.0, L.n))); }
    else if (L.vals32) { C8_NB_SWITCH(nb_, (pdl_launch(g, 128, 0, s)(k_bsr_residual<NB, float>, L.rowptr, L.colind, L.vals32, cur, b, r, L.n))); }
    else { C8_NB_SWITCH(nb_, (pdl_launch(g, 128, 0, s)(k_bsr_residual<NB, double>, L.rowptr, L.colind, L.vals, cur, b, r, L.n))); }
  }
  amg_dbg(s, "residual", l);
  AmgLevel& C = lv_[l + 1];
  if (C.n > 0) {
    const int gc = (C.n * nb_ + 127) / 128;
    C8_NB_SWITCH(nb_, (pdl_launch(gc, 128, 0, s)(k_restrict<NB>, L.aggptr, L.aggmem, r, C.b, C.n)));
    if (L.coarse_replicated) ctx_->allreduce_cb(ctx_->comm_user, C.b, C.n * nb_);
  }
  amg_dbg(s, "restrict", l);
  cycle(l + 1, C.b, C.x);
  halo(l + 1, C.x);   // the prolongation reads the aggregates of ghost columns
  amg_dbg(s, "coarse cycle", l);
  if (l == 0 && fine_prolong_add() && L.npc > 0) {
    // fine level: the correction as an elementwise pass (3 vectors) + a plain sweep, instead of two
    // more dependent gathers per block inside the sweep
    C8_NB_SWITCH(nb_, (pdl_launch((L.npc * nb_ + 255) / 256, 256, 0, s)(k_prolong_add<NB>, L.agg, C.x, const_cast<double*>(cur), opt.over_correction, L.npc)));
    if (cheb) sweep(l, b, cur, buf(nu1), nullptr, nullptr, 0.0, ch_a0);
    else sweep(l, b, cur, buf(nu1), nullptr);
  } else
  sweep(l, b, cur, buf(nu1), C.x);   // the ghost entries of cur are still current
  amg_dbg(s, "prolong sweep", l);
  for (int w = nu1 + 1; w < writes; ++w) {
    halo(l, buf(w - 1));
    if (cheb) sweep(l, b, buf(w - 1), buf(w), nullptr, cur, ch_c1, ch_c2);   // x0 = cur (aliases the output buffer)
    else sweep(l, b, buf(w - 1), buf(w), nullptr);
    amg_dbg(s, "post sweep", l);
  }
}

void Amg::apply(const double* r, double* z) {
  // ghost entries of z: zero when the hierarchy acts on the owned block only; GMRES refreshes them
  // by its own halo copy before the SpMV either way
  if (lv_[0].ld > lv_[0].n)
    cudaMemsetAsync(z + size_t(lv_[0].n) * nb_, 0, size_t(lv_[0].ld - lv_[0].n) * nb_ * sizeof(double),
                    ctx_->stream);
  cycle(0, r, z);
}

}  // namespace c8
