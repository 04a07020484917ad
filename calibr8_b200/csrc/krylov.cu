// K10: restarted GMRES(m) with a device-resident Arnoldi process.
//
// Replaces Belos' BlockGmres solver manager of the reference (src/linear_solve.cpp:22-124).  The
// Hessenberg matrix, the Givens rotations and the least-squares right-hand side live on the
// device; one iteration is a fixed sequence of kernel launches (preconditioner, BSR SpMV, two
// passes of classical Gram-Schmidt with fused multi-dots, normalisation, a one-thread Givens
// update) with NO host synchronisation.  The host only looks at the residual estimates every
// `check_every` iterations (one small D2H copy + stream sync), so small and medium systems are
// not launch/sync-latency bound and a partitioned run overlaps NCCL reductions on the same stream.
// Right preconditioning: block-Jacobi or the aggregation multigrid of amg.cu.
#include <algorithm>
#include <map>
#include <memory>

#include "amg.cuh"

namespace c8 {

// column j of H from the two Gram-Schmidt passes, previous rotations applied, new rotation
// generated, least-squares rhs updated; res[j] = |g[j+1]| (the residual-norm estimate).
// The norm of the new basis vector needs no reduction of its own: the second pass also returned
// ww = w.w of the vector it corrected (d2[j+1]), and the basis is orthonormal, so
// |w - V d2|^2 = ww - sum d2^2 (d2 is at rounding level after the first pass: no cancellation).
// nrm2[0] receives that squared norm for k_normalize.
__global__ void k_givens(double* __restrict__ H, int ldh, double* __restrict__ cs,
                         double* __restrict__ sn, double* __restrict__ g,
                         const double* __restrict__ d1, const double* __restrict__ d2,
                         double* __restrict__ nrm2, int j, double* __restrict__ res) {
  pdl_wait();
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double* h = H + size_t(j) * ldh;
  double ww = d2[j + 1];
  for (int i = 0; i <= j; ++i) { h[i] = d1[i] + d2[i]; ww = fma(-d2[i], d2[i], ww); }
  nrm2[0] = ww;
  h[j + 1] = sqrt(fmax(ww, 0.0));
  for (int i = 0; i < j; ++i) {
    const double t = cs[i] * h[i] + sn[i] * h[i + 1];
    h[i + 1] = -sn[i] * h[i] + cs[i] * h[i + 1];
    h[i] = t;
  }
  const double a = h[j], b = h[j + 1];
  const double d = hypot(a, b);
  const double c = d > 0.0 ? a / d : 1.0, s = d > 0.0 ? b / d : 0.0;
  cs[j] = c; sn[j] = s;
  h[j] = d; h[j + 1] = 0.0;
  g[j + 1] = -s * g[j];
  g[j] = c * g[j];
  res[j] = fabs(g[j + 1]);
}

// y = R^-1 g for the leading k columns
__global__ void k_backsolve(const double* __restrict__ H, int ldh, const double* __restrict__ g,
                            double* __restrict__ y, int k) {
  pdl_wait();
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  for (int i = k - 1; i >= 0; --i) {
    double t = g[i];
    for (int c = i + 1; c < k; ++c) t -= H[size_t(c) * ldh + i] * y[c];
    y[i] = t / H[size_t(i) * ldh + i];
  }
}

// v *= 1/sqrt(nrm2[0]) (device scalar); g0 (optional) receives the norm
__global__ void k_normalize(double* __restrict__ v, const double* __restrict__ nrm2, long long n,
                            double* __restrict__ g0) {
  pdl_wait();
  const double nr = sqrt(fmax(nrm2[0], 0.0));
  const double inv = nr > 0.0 ? 1.0 / nr : 0.0;
  if (g0 && blockIdx.x == 0 && threadIdx.x == 0) g0[0] = nr;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    v[i] *= inv;
}

struct SolverState {
  int m = 0;
  long long n = 0;  // vector allocation length (all local nodes incl. ghosts)
  double *V = nullptr, *w = nullptr, *z = nullptr, *dinv = nullptr, *partial = nullptr;
  double *dots = nullptr, *dots2 = nullptr, *nrm = nullptr, *coef = nullptr;
  double *H = nullptr, *cs = nullptr, *sn = nullptr, *g = nullptr, *res = nullptr;
  double* h_buf = nullptr;  // pinned
  int pc_type = C8_PC_AMG;
  AmgOptions amg_opt;
  std::unique_ptr<Amg> amg;
  bool amg_built = false;
  long long total_iters = 0, total_solves = 0;
  // CUDA graphs of one Arnoldi iteration, one per column index j (the launch sequence of an
  // iteration depends on j only through counts and pointers into the fixed workspace); valid for
  // one matrix pointer / preconditioner configuration
  std::vector<cudaGraphExec_t> iter_graph;
  const double* graph_A = nullptr;
  int graph_pc = -1;
  bool use_graphs = true;
  void drop_graphs() {
    for (cudaGraphExec_t g : iter_graph) if (g) cudaGraphExecDestroy(g);
    iter_graph.clear();
    graph_A = nullptr; graph_pc = -1;
  }
  void free_ws() {
    drop_graphs();
    void* p[] = {V, w, z, dinv, partial, dots, dots2, nrm, coef, H, cs, sn, g, res};
    for (void* q : p) if (q) cudaFree(q);
    if (h_buf) cudaFreeHost(h_buf);
    V = w = z = dinv = partial = dots = dots2 = nrm = coef = H = cs = sn = g = res = nullptr;
    h_buf = nullptr;
    m = 0; n = 0;
  }
};
static std::map<c8_ctx*, SolverState> g_state;

static int ensure_ws(c8_ctx* ctx, SolverState& ws, int m) {
  const long long n = (long long)ctx->n_nodes * ctx->kt->nb;
  if (ws.m >= m && ws.n == n) return C8_OK;
  ws.free_ws();
  C8_CUDA(ctx, cudaMalloc(&ws.V, size_t(m + 1) * n * sizeof(double)));
  C8_CUDA(ctx, cudaMemsetAsync(ws.V, 0, size_t(m + 1) * n * sizeof(double), ctx->stream));
  C8_CUDA(ctx, cudaMalloc(&ws.w, n * sizeof(double)));
  C8_CUDA(ctx, cudaMalloc(&ws.z, n * sizeof(double)));
  C8_CUDA(ctx, cudaMemsetAsync(ws.w, 0, n * sizeof(double), ctx->stream));
  C8_CUDA(ctx, cudaMemsetAsync(ws.z, 0, n * sizeof(double), ctx->stream));
  C8_CUDA(ctx, cudaMalloc(&ws.dinv, size_t(ctx->n_nodes) * ctx->kt->nb * ctx->kt->nb * sizeof(double)));
  C8_CUDA(ctx, cudaMalloc(&ws.partial, size_t(m + 2) * 256 * sizeof(double)));
  C8_CUDA(ctx, cudaMalloc(&ws.dots, size_t(m + 2) * sizeof(double)));
  C8_CUDA(ctx, cudaMalloc(&ws.dots2, size_t(m + 2) * sizeof(double)));
  C8_CUDA(ctx, cudaMalloc(&ws.nrm, 2 * sizeof(double)));
  C8_CUDA(ctx, cudaMalloc(&ws.coef, size_t(m + 2) * sizeof(double)));
  C8_CUDA(ctx, cudaMalloc(&ws.H, size_t(m + 2) * (m + 1) * sizeof(double)));
  C8_CUDA(ctx, cudaMalloc(&ws.cs, size_t(m + 2) * sizeof(double)));
  C8_CUDA(ctx, cudaMalloc(&ws.sn, size_t(m + 2) * sizeof(double)));
  C8_CUDA(ctx, cudaMalloc(&ws.g, size_t(m + 2) * sizeof(double)));
  C8_CUDA(ctx, cudaMalloc(&ws.res, size_t(m + 2) * sizeof(double)));
  C8_CUDA(ctx, cudaMallocHost(&ws.h_buf, size_t(m + 2) * sizeof(double)));
  ws.m = m; ws.n = n;
  return C8_OK;
}

}  // namespace c8

using namespace c8;

extern "C" {

int c8_set_preconditioner(c8_ctx* ctx, int type, const double* opts, int n_opts) {
  C8_REQUIRE(ctx, type == C8_PC_BLOCK_JACOBI || type == C8_PC_AMG, "unknown preconditioner type");
  SolverState& st = g_state[ctx];
  st.pc_type = type;
  AmgOptions o;
  if (opts) {
    if (n_opts > 0) o.nu_pre = int(opts[0]);
    if (n_opts > 1) o.nu_post = int(opts[1]);
    if (n_opts > 2) o.omega = opts[2];
    if (n_opts > 3) o.over_correction = opts[3];
    if (n_opts > 4) o.coarsest_max_nodes = int(opts[4]);
    if (n_opts > 5) o.max_aggregate_size = int(opts[5]);
    if (n_opts > 6) o.coarse_aggregate_size = int(opts[6]);
    if (n_opts > 7) o.coarse_nu = int(opts[7]);
    if (n_opts > 8) o.distributed = opts[8] != 0.0;
    if (n_opts > 9) o.replicate_max_nodes = int(opts[9]);
  }
  const bool rebuild = o.coarsest_max_nodes != st.amg_opt.coarsest_max_nodes ||
                       o.max_aggregate_size != st.amg_opt.max_aggregate_size ||
                       o.coarse_aggregate_size != st.amg_opt.coarse_aggregate_size ||
                       o.distributed != st.amg_opt.distributed ||
                       o.replicate_max_nodes != st.amg_opt.replicate_max_nodes;
  st.amg_opt = o;
  st.drop_graphs();
  if (st.amg) st.amg->opt = o;
  if (rebuild) { st.amg.reset(); st.amg_built = false; }
  return C8_OK;
}

int c8_dot(c8_ctx* ctx, const double* x_dev, const double* y_dev, double* out_host) {
  C8_REQUIRE(ctx, ctx->kt != nullptr, "c8_set_model must be called first");
  SolverState& ws = g_state[ctx];
  int rc = ensure_ws(ctx, ws, ws.m > 0 ? ws.m : 8);
  if (rc != C8_OK) return rc;
  LinAlg la(ctx);
  la.multi_dot(x_dev, 0, y_dev, 1, ws.partial, ws.dots);
  C8_CUDA(ctx, cudaMemcpyAsync(ws.h_buf, ws.dots, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  C8_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  *out_host = ws.h_buf[0];
  return C8_OK;
}

// Restarted GMRES(m), right preconditioned: solves A x = b, x_dev in: initial guess, out: solution.
// Converges on ||b - A x|| <= max(rel_tol*||b - A x0||, abs_tol).
// info_host[0] = iterations, info_host[1] = final residual norm, info_host[2] = initial norm.
int c8_gmres(c8_ctx* ctx, const double* A, const double* b, double* x, int restart, int max_iters,
             double rel_tol, double abs_tol, double* info_host) {
  C8_REQUIRE(ctx, ctx->kt != nullptr, "c8_set_model must be called first");
  C8_REQUIRE(ctx, restart >= 1, "restart must be positive");
  const int m = restart;
  SolverState& ws = g_state[ctx];
  int rc = ensure_ws(ctx, ws, m);
  if (rc != C8_OK) return rc;
  LinAlg la(ctx);
  const long long n = la.n;      // owned dofs: the range every vector operation runs over
  const long long ld = ws.n;     // allocation stride of a Krylov vector (owned + ghost)
  const int ldh = ws.m + 2;
  cudaStream_t s = ctx->stream;
  const int ag = grid_for(n, 256, la.sms);
  const int check_every = 4;

  // ---- preconditioner set-up for this matrix
  bool use_amg = ws.pc_type == C8_PC_AMG;
  if (use_amg) {
    if (!ws.amg_built) {
      ws.amg.reset(new Amg(ctx));
      ws.amg->opt = ws.amg_opt;
      if ((rc = ws.amg->build()) != C8_OK) return rc;
      ws.amg_built = true;
    }
    if ((rc = ws.amg->setup(A)) != C8_OK) return rc;
  } else {
    la.jacobi_setup(A, ws.dinv);
  }
  auto precond = [&](const double* r, double* z) {
    if (use_amg) ws.amg->apply(r, z);
    else la.jacobi_apply(ws.dinv, r, z);
  };
  auto fetch = [&](const double* dev, int cnt) -> int {
    C8_CUDA(ctx, cudaMemcpyAsync(ws.h_buf, dev, cnt * sizeof(double), cudaMemcpyDeviceToHost, s));
    C8_CUDA(ctx, cudaStreamSynchronize(s));
    return C8_OK;
  };

  // one Arnoldi iteration for column j: a fixed launch sequence (captured into a CUDA graph on one
  // GPU; a partitioned run launches directly because the transports enqueue host-side work)
  auto iteration = [&](int j) {
    double* vj1 = ws.V + size_t(j + 1) * ld;
    precond(ws.V + size_t(j) * ld, ws.z);
    la.halo(ws.z);
    la.spmv(A, ws.z, vj1);
    // classical Gram-Schmidt, two passes: h = V^T w ; w -= V h
    la.multi_dot(ws.V, ld, vj1, j + 1, ws.partial, ws.dots);
    pdl_launch(ag, 256, 0, s)(k_multi_axpy_neg, ws.V, ld, ws.dots, j + 1, n, vj1);
    la.multi_dot(ws.V, ld, vj1, j + 2, ws.partial, ws.dots2);   // V[j+1] is vj1 itself: dots2[j+1] = w.w
    pdl_launch(ag, 256, 0, s)(k_multi_axpy_neg, ws.V, ld, ws.dots2, j + 1, n, vj1);
    pdl_launch(1, 32, 0, s)(k_givens, ws.H, ldh, ws.cs, ws.sn, ws.g, ws.dots, ws.dots2, ws.nrm, j, ws.res);
    pdl_launch(ag, 256, 0, s)(k_normalize, vj1, ws.nrm, n, nullptr);
  };
  // capture needs a real stream; a partitioned run captures too when its transport only enqueues
  // stream work (NCCL), not when it stages through the host
  static const bool graphs_env = [] { const char* e = getenv("C8_GRAPHS"); return !(e && e[0] == '0'); }();
  bool graphs = graphs_env && ws.use_graphs && ((!ctx->halo_cb && !ctx->allreduce_cb) || ctx->comm_capturable) &&
                s != nullptr && s != cudaStreamLegacy && s != cudaStreamPerThread;
  if (graphs) {
    const int pc_key = use_amg ? 1 : 0;
    if (ws.graph_A != A || ws.graph_pc != pc_key || int(ws.iter_graph.size()) != ws.m) {
      ws.drop_graphs();
      ws.iter_graph.assign(ws.m, nullptr);
      ws.graph_A = A; ws.graph_pc = pc_key;
    }
  }

  int total = 0;
  double beta0 = -1.0, beta = 0.0, target = 0.0;
  while (true) {
    // r = b - A x  -> V[0], beta = |r|
    la.halo(x);
    la.spmv(A, x, ws.w);
    C8_CUDA(ctx, cudaMemcpyAsync(ws.V, b, n * sizeof(double), cudaMemcpyDeviceToDevice, s));
    pdl_launch(ag, 256, 0, s)(k_axpby, -1.0, ws.w, 1.0, ws.V, n);
    la.multi_dot(ws.V, 0, ws.V, 1, ws.partial, ws.nrm);
    if ((rc = fetch(ws.nrm, 1)) != C8_OK) return rc;
    beta = std::sqrt(ws.h_buf[0]);
    if (beta0 < 0) { beta0 = beta; target = std::max(rel_tol * beta0, abs_tol); }
    if (beta <= target || total >= max_iters || !(beta == beta)) break;
    C8_CUDA(ctx, cudaMemsetAsync(ws.g, 0, size_t(m + 2) * sizeof(double), s));
    pdl_launch(ag, 256, 0, s)(k_normalize, ws.V, ws.nrm, n, ws.g);   // V0 = r/beta, g[0] = beta
    int j = 0, used = 0;
    bool done = false;
    while (j < m && total < max_iters && !done) {
      const int j_end = std::min(std::min(j + check_every, m), j + (max_iters - total));
      for (; j < j_end; ++j, ++total) {
        if (graphs) {
          if (!ws.iter_graph[j]) {
            cudaGraph_t graph = nullptr;
            bool ok = cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
            if (ok) {
              iteration(j);
              ok = cudaStreamEndCapture(s, &graph) == cudaSuccess && graph != nullptr;
            }
            if (ok) ok = cudaGraphInstantiate(&ws.iter_graph[j], graph, 0) == cudaSuccess;
            if (graph) cudaGraphDestroy(graph);
            if (!ok) {  // capture not possible here (e.g. a transport that cannot be captured)
              cudaGetLastError();
              ws.iter_graph[j] = nullptr;
              ws.use_graphs = false;
              graphs = false;
              iteration(j);
              continue;
            }
          }
          C8_CUDA(ctx, cudaGraphLaunch(ws.iter_graph[j], s));
        } else {
          iteration(j);
        }
      }
      // look at the residual estimates of the iterations just queued
      if ((rc = fetch(ws.res, j)) != C8_OK) return rc;
      used = j;
      for (int i = 0; i < j; ++i)
        if (!(ws.h_buf[i] > target)) { used = i + 1; done = true; break; }  // also stops on NaN
    }
    total -= (j - used);  // iterations queued past convergence are not counted
    // x += M^-1 (V y), y from the leading `used` columns
    pdl_launch(1, 32, 0, s)(k_backsolve, ws.H, ldh, ws.g, ws.coef, used);
    C8_CUDA(ctx, cudaMemsetAsync(ws.w, 0, n * sizeof(double), s));
    pdl_launch(ag, 256, 0, s)(k_multi_axpy, ws.V, ld, ws.coef, used, n, ws.w);
    precond(ws.w, ws.z);
    pdl_launch(ag, 256, 0, s)(k_axpby, 1.0, ws.z, 1.0, x, n);
  }
  ws.total_iters += total; ws.total_solves += 1;
  if (info_host) { info_host[0] = total; info_host[1] = beta; info_host[2] = beta0; }
  C8_CUDA(ctx, cudaGetLastError());
  return (beta <= target) ? C8_OK : C8_ERR_NOT_CONVERGED;
}

// out[0] = levels, out[1] = operator complexity, out[2..] = nodes per level
int c8_preconditioner_info(c8_ctx* ctx, double* out, int n_out) {
  SolverState& st = g_state[ctx];
  for (int i = 0; i < n_out; ++i) out[i] = 0.0;
  if (!st.amg) return C8_OK;
  if (n_out > 0) out[0] = st.amg->num_levels();
  if (n_out > 1) out[1] = st.amg->operator_complexity();
  for (int l = 0; l < st.amg->num_levels() && 2 + l < n_out; ++l) out[2 + l] = st.amg->levels()[l].n;
  return C8_OK;
}

// the mesh / partition / model changed: drop the workspace and the hierarchy, keep the options
void c8_linalg_invalidate(c8_ctx* ctx) {
  auto it = g_state.find(ctx);
  if (it == g_state.end()) return;
  cudaSetDevice(ctx->device);
  it->second.amg.reset();
  it->second.amg_built = false;
  it->second.free_ws();
}

void c8_linalg_release(c8_ctx* ctx) {
  auto it = g_state.find(ctx);
  if (it == g_state.end()) return;
  cudaSetDevice(ctx->device);
  it->second.amg.reset();
  it->second.free_ws();
  g_state.erase(it);
}

}  // extern "C"
