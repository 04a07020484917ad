// C-ABI implementation, part 1: context, discretisation, model set-up, layout
// helpers and the forward hot-path entry points (include/c8b200.h).
#include "c8b200.h"
#include "context.cuh"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace c8 {

// tables defined in the combo_*.cu translation units
const KernelTable* table_3d_mixed_elastic();
const KernelTable* table_3d_mixed_small_j2();
const KernelTable* table_3d_mixed_small_hill();
const KernelTable* table_3d_mixed_hyper_j2();
const KernelTable* table_2d_mixed_elastic();
const KernelTable* table_2d_mixed_small_j2();
const KernelTable* table_2d_mixed_small_hill_pe();
const KernelTable* table_2d_mixed_hyper_j2_pe();
const KernelTable* table_2d_ps_small_hill();
const KernelTable* table_2d_ps_hyper_j2();
const KernelTable* table_3d_mixed_hypo_hill();
const KernelTable* table_2d_mixed_hypo_hill_pe();
const KernelTable* table_2d_ps_hypo_hill();

const KernelTable* find_kernel_table(int dim, int mech, int local_type) {
  const KernelTable* all[] = {
      table_3d_mixed_elastic(),  table_3d_mixed_small_j2(),     table_3d_mixed_small_hill(),
      table_3d_mixed_hyper_j2(), table_2d_mixed_elastic(),      table_2d_mixed_small_j2(),
      table_2d_mixed_small_hill_pe(), table_2d_mixed_hyper_j2_pe(), table_2d_ps_small_hill(),
      table_2d_ps_hyper_j2(),    table_3d_mixed_hypo_hill(),    table_2d_mixed_hypo_hill_pe(),
      table_2d_ps_hypo_hill()};
  for (const KernelTable* t : all)
    if (t->dim == dim && t->mech == mech && t->local_type == local_type) return t;
  return nullptr;
}

int fail(c8_ctx* ctx, int code, const std::string& msg) {
  if (ctx) ctx->err = msg;
  return code;
}
bool cuda_ok(c8_ctx* ctx, cudaError_t e, const char* what) {
  if (e == cudaSuccess) return true;
  if (ctx) ctx->err = std::string(what) + ": " + cudaGetErrorString(e);
  return false;
}
double* stage(c8_ctx* ctx, size_t bytes) {
  if (bytes > ctx->stage_bytes) {
    if (ctx->d_stage) cudaFree(ctx->d_stage);
    ctx->d_stage = nullptr;
    if (cudaMalloc(&ctx->d_stage, bytes) != cudaSuccess) { ctx->stage_bytes = 0; return nullptr; }
    ctx->stage_bytes = bytes;
  }
  return ctx->d_stage;
}
double* pinned(c8_ctx* ctx, size_t bytes) {
  if (bytes > ctx->pinned_bytes) {
    if (ctx->d_stage_out) cudaFree(ctx->d_stage_out);
  if (ctx->h_pinned) cudaFreeHost(ctx->h_pinned);
    ctx->h_pinned = nullptr;
    if (cudaMallocHost(&ctx->h_pinned, bytes) != cudaSuccess) { ctx->pinned_bytes = 0; return nullptr; }
    ctx->pinned_bytes = bytes;
  }
  return ctx->h_pinned;
}

template <class T>
static bool upload(c8_ctx* ctx, T** d, const std::vector<T>& h) {
  if (*d) cudaFree(*d);
  *d = nullptr;
  if (h.empty()) return true;
  if (!cuda_ok(ctx, cudaMalloc(d, h.size() * sizeof(T)), "cudaMalloc")) return false;
  return cuda_ok(ctx, cudaMemcpy(*d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice),
                 "cudaMemcpy H2D");
}

// ---- layout kernels -----------------------------------------------------------
// per-residual host layout (u [n][dim], p [n]) <-> node-interleaved [n][nb]
__global__ void k_interleave(const double* __restrict__ u, const double* __restrict__ p,
                             double* __restrict__ x, int n_nodes, int dim, int nb) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_nodes * nb) return;
  const int n = i / nb, q = i % nb;
  x[i] = q < dim ? u[size_t(n) * dim + q] : p[n];
}
__global__ void k_deinterleave(const double* __restrict__ x, double* __restrict__ u,
                               double* __restrict__ p, int n_nodes, int dim, int nb) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_nodes * nb) return;
  const int n = i / nb, q = i % nb;
  if (q < dim) u[size_t(n) * dim + q] = x[i];
  else p[n] = x[i];
}
// AoS [n_elems][nxi] <-> SoA [nxi][ld]
__global__ void k_aos_to_soa(const double* __restrict__ a, double* __restrict__ s, int n, int nxi,
                             long long ld) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= (long long)n * nxi) return;
  const int q = int(i / n);
  const long long e = i % n;
  s[q * ld + e] = a[e * nxi + q];
}
__global__ void k_soa_to_aos(const double* __restrict__ s, double* __restrict__ a, int n, int nxi,
                             long long ld) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= (long long)n * nxi) return;
  const long long e = i / nxi;
  const int q = int(i % nxi);
  a[i] = s[q * ld + e];
}
// BSR values -> the reference's block (i,j) CSR values.  Block (i,j) row r = node*neq_i+eq_i has
// (row nnz of node) * neq_j entries ordered by (adjacent node, eq_j): the same order as the BSR row.
__global__ void k_bsr_to_block_csr(const double* __restrict__ bsr, const int* __restrict__ rowptr,
                                   double* __restrict__ out, int n_nodes, int nb, int eq_i0,
                                   int neq_i, int eq_j0, int neq_j) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_nodes * neq_i) return;
  const int node = r / neq_i, ei = r % neq_i;
  const int b0 = rowptr[node], b1 = rowptr[node + 1];
  // CSR row start: sum over previous rows = (rowptr[node]*neq_i + ei*(b1-b0)) * neq_j
  size_t o = (size_t(b0) * neq_i + size_t(ei) * (b1 - b0)) * neq_j;
  for (int k = b0; k < b1; ++k)
    for (int ej = 0; ej < neq_j; ++ej)
      out[o++] = bsr[size_t(k) * nb * nb + (eq_i0 + ei) * nb + (eq_j0 + ej)];
}

}  // namespace c8

using namespace c8;

static void residual_split(const c8_ctx* ctx, int i, int* eq0, int* neq) {
  // residual 0 = displacement (dim eqs), residual 1 = pressure (mixed only)
  if (i == 0) { *eq0 = 0; *neq = ctx->dim; }
  else { *eq0 = ctx->dim; *neq = 1; }
}
static int num_resid(const c8_ctx* ctx) { return (ctx->kt && ctx->kt->nb > ctx->dim) ? 2 : 1; }

namespace c8 {
// element-matrix scratch of the two-phase assembly: [n_elems + 1][NX][NX] (the last slot is a
// sink for thread groups that must not contribute)
double* element_scratch(c8_ctx* ctx) {
  const size_t need = size_t(ctx->n_elems) + 1;
  if (ctx->d_emat && ctx->emat_elems >= need && ctx->emat_nx == ctx->kt->nx) return ctx->d_emat;
  if (ctx->d_emat) cudaFree(ctx->d_emat);
  ctx->d_emat = nullptr;
  const size_t bytes = need * ctx->kt->nx * ctx->kt->nx * sizeof(double);
  if (!cuda_ok(ctx, cudaMalloc(&ctx->d_emat, bytes), "cudaMalloc(element scratch)")) return nullptr;
  ctx->emat_elems = need; ctx->emat_nx = ctx->kt->nx;
  return ctx->d_emat;
}

bool chunked_assembly(c8_ctx* ctx, FwdArgs& a) {
  a.cg_ptr_host = nullptr;
  if (ctx->n_chunks < 2 || !a.vals || a.elem_J || a.elem_R) return false;
  // a partition keeps the rows of its owned nodes only: those are the leading blocks, and inside a chunk
  // the entries are sorted by block, so the owned rows are a prefix of every chunk's entries
  const int n_row_blocks = ctx->h_rowptr[ctx->n_owned_nodes];
  if (ctx->cg_end_rows != n_row_blocks) {
    ctx->h_cg_end.assign(ctx->n_chunks, 0);
    for (int c = 0; c < ctx->n_chunks; ++c) {
      const unsigned* b0 = ctx->h_cg_blk.data() + ctx->h_cg_ptr[c];
      const unsigned* b1 = ctx->h_cg_blk.data() + ctx->h_cg_ptr[c + 1];
      const unsigned* it = std::partition_point(b0, b1, [&](unsigned w) { return int(w & 0x7fffffffu) < n_row_blocks; });
      ctx->h_cg_end[c] = int(it - ctx->h_cg_blk.data());
    }
    ctx->cg_end_rows = n_row_blocks;
  }
  const int nx = ctx->kt->nx;
  if (!ctx->d_emat_chunk || ctx->emat_chunk_nx != nx) {
    if (ctx->d_emat_chunk) cudaFree(ctx->d_emat_chunk);
    ctx->d_emat_chunk = nullptr;
    const size_t bytes = (size_t(ctx->chunk_elems) + 1) * nx * nx * sizeof(double);
    if (cudaMalloc(&ctx->d_emat_chunk, bytes) != cudaSuccess) { cudaGetLastError(); return false; }
    ctx->emat_chunk_nx = nx;
  }
  a.emat = ctx->d_emat_chunk;
  a.cg_ptr_host = ctx->h_cg_ptr.data();
  a.cg_end_host = ctx->h_cg_end.data();
  return true;
}

__global__ void k_int_to_double(const int* i, double* d) { *d = double(*i); }

// number of failed local solves over ALL parts (PCU_Add_Int of the status, src/primal.cpp:96)
int fetch_n_failed(c8_ctx* ctx, int* out) {
  if (ctx->allreduce_cb) {
    if (!ctx->d_scalar) C8_CUDA(ctx, cudaMalloc(&ctx->d_scalar, 8 * sizeof(double)));
    k_int_to_double<<<1, 1, 0, ctx->stream>>>(ctx->d_nfailed, ctx->d_scalar);
    ctx->allreduce_cb(ctx->comm_user, ctx->d_scalar, 1);
    double h = 0.0;
    C8_CUDA(ctx, cudaMemcpyAsync(&h, ctx->d_scalar, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    C8_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *out = int(h + 0.5);
  } else {
    C8_CUDA(ctx, cudaMemcpyAsync(out, ctx->d_nfailed, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    C8_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  return C8_OK;
}
}  // namespace c8

extern "C" {

const char* c8_version(void) { return "calibr8_b200 0.1 (sm_100a)"; }

c8_ctx* c8_create(int device) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0 || device >= n) {
    std::fprintf(stderr, "c8_create: no CUDA device %d available (this library has no CPU path)\n",
                 device);
    return nullptr;
  }
  if (cudaSetDevice(device) != cudaSuccess) return nullptr;
  c8_ctx* ctx = new c8_ctx;
  ctx->device = device;
  if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
    delete ctx;
    return nullptr;
  }
  ctx->own_stream = true;
  if (cudaMalloc(&ctx->d_nfailed, 2 * sizeof(int)) != cudaSuccess) {   // [0] failed local solves, [1] tile counter of the persistent K1
    cudaStreamDestroy(ctx->stream);
    delete ctx;
    return nullptr;
  }
  return ctx;
}

void c8_destroy(c8_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  c8_linalg_release(ctx);
  c8_comm_release(ctx);
  void* ptrs[] = {ctx->d_conn, ctx->d_coords, ctx->d_elem_es, ctx->d_rowptr, ctx->d_colind,
                  ctx->d_eoff, ctx->d_params, ctx->d_nfailed, ctx->d_A, ctx->d_b, ctx->d_x,
                  ctx->d_xp, ctx->d_xi, ctx->d_xip, ctx->d_stage, ctx->d_scalar, ctx->d_gptr,
                  ctx->d_gsrc, ctx->d_emat, ctx->d_cg_ptr, ctx->d_cg_blk, ctx->d_cg_k, ctx->d_emat_chunk};
  for (void* p : ptrs)
    if (p) cudaFree(p);
  if (ctx->h_pinned) cudaFreeHost(ctx->h_pinned);
  if (ctx->ev_elements) cudaEventDestroy(ctx->ev_elements);
  if (ctx->side_stream) cudaStreamDestroy(ctx->side_stream);
  if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

const char* c8_last_error(c8_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int c8_set_stream(c8_ctx* ctx, void* s) {
  if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
  ctx->stream = (cudaStream_t)s;
  ctx->own_stream = false;
  return C8_OK;
}
int c8_synchronize(c8_ctx* ctx) {
  C8_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return C8_OK;
}

int c8_set_mesh(c8_ctx* ctx, int dim, int n_elems, int n_nodes, const int32_t* conn,
                const double* coords, const int32_t* elem_set, int n_es) {
  C8_REQUIRE(ctx, dim == 2 || dim == 3, "dim must be 2 or 3");
  C8_CUDA(ctx, cudaSetDevice(ctx->device));
  const int nn = dim + 1;
  ctx->dim = dim; ctx->nn = nn; ctx->n_elems = n_elems; ctx->n_nodes = n_nodes;
  ctx->n_es = n_es > 0 ? n_es : 1;
  ctx->h_conn.assign(conn, conn + size_t(n_elems) * nn);
  ctx->h_coords.resize(size_t(n_nodes) * dim);
  for (int n = 0; n < n_nodes; ++n)
    for (int k = 0; k < dim; ++k) ctx->h_coords[size_t(n) * dim + k] = coords[size_t(n) * 3 + k];
  for (size_t k = 0; k < ctx->h_conn.size(); ++k)
    C8_REQUIRE(ctx, ctx->h_conn[k] >= 0 && ctx->h_conn[k] < n_nodes, "connectivity out of range");

  // node graph (the reference's CrsGraph per block is this graph expanded by the equation
  // counts, src/disc.cpp:356-387): node -> sorted unique nodes sharing an element
  std::vector<int> cnt(n_nodes + 1, 0);
  for (int v : ctx->h_conn) cnt[v + 1]++;
  for (int n = 0; n < n_nodes; ++n) cnt[n + 1] += cnt[n];
  std::vector<int> n2e(ctx->h_conn.size());
  {
    std::vector<int> pos(cnt.begin(), cnt.end() - 1);
    for (int e = 0; e < n_elems; ++e)
      for (int a = 0; a < nn; ++a) n2e[pos[ctx->h_conn[size_t(e) * nn + a]]++] = e;
  }
  ctx->h_rowptr.assign(n_nodes + 1, 0);
  ctx->h_colind.clear();
  ctx->h_colind.reserve(size_t(n_nodes) * (dim == 3 ? 16 : 8));
  std::vector<int> tmp;
  for (int n = 0; n < n_nodes; ++n) {
    tmp.clear();
    for (int k = cnt[n]; k < cnt[n + 1]; ++k) {
      const int e = n2e[k];
      for (int a = 0; a < nn; ++a) tmp.push_back(ctx->h_conn[size_t(e) * nn + a]);
    }
    if (tmp.empty()) tmp.push_back(n);  // isolated node keeps its diagonal
    std::sort(tmp.begin(), tmp.end());
    tmp.erase(std::unique(tmp.begin(), tmp.end()), tmp.end());
    ctx->h_colind.insert(ctx->h_colind.end(), tmp.begin(), tmp.end());
    ctx->h_rowptr[n + 1] = int(ctx->h_colind.size());
  }
  ctx->nnzb = int(ctx->h_colind.size());
  // element -> BSR block offsets (the role of Disc::compute_scatter_offsets, disc.cpp:414-459)
  std::vector<int> eoff(size_t(n_elems) * nn * nn);
  for (int e = 0; e < n_elems; ++e)
    for (int a = 0; a < nn; ++a) {
      const int ra = ctx->h_conn[size_t(e) * nn + a];
      const int* rb = &ctx->h_colind[ctx->h_rowptr[ra]];
      const int* re = &ctx->h_colind[ctx->h_rowptr[ra + 1]];
      for (int b = 0; b < nn; ++b) {
        const int* it = std::lower_bound(rb, re, ctx->h_conn[size_t(e) * nn + b]);
        eoff[(size_t(e) * nn + a) * nn + b] = int(it - ctx->h_colind.data());
      }
    }
  // gather plan = the inverse of eoff: for every BSR block the (element, node pair) slots that
  // contribute to it, sorted by element id (deterministic summation order)
  {
    std::vector<int> gptr(size_t(ctx->nnzb) + 1, 0), gsrc(eoff.size());
    for (int v : eoff) gptr[v + 1]++;
    for (int k = 0; k < ctx->nnzb; ++k) gptr[k + 1] += gptr[k];
    std::vector<int> pos(gptr.begin(), gptr.end() - 1);
    for (size_t q = 0; q < eoff.size(); ++q) gsrc[pos[eoff[q]]++] = int(q);
    if (!upload(ctx, &ctx->d_gptr, gptr)) return C8_ERR_CUDA;
    if (!upload(ctx, &ctx->d_gsrc, gsrc)) return C8_ERR_CUDA;
    // chunked forward assembly (args.h): split every block's contribution list at the chunk
    // boundaries of the element index; entries sorted by chunk, then block.  OFF by default (one pass):
    // measured on B200 at 1 M tets it is slower at every chunk size (2.15 ms one pass; 2.62 / 2.81 / 2.90 /
    // 3.26 ms at 75 776 / 37 888 / 18 944 / 9 472 elements per chunk) -- every element-kernel launch pays
    // a fixed ~10-15 us (cold instruction cache of the 140 KB program on every SM + the tail of its last
    // wave), which outweighs reading the scratch from L2.  c8_set_assembly_chunk / C8_ASM_CHUNK enable it.
    static const int chunk_default = [] { const char* e = getenv("C8_ASM_CHUNK"); return e ? atoi(e) : 0; }();
    const int chunk_env = ctx->chunk_request >= 0 ? ctx->chunk_request : chunk_default;
    ctx->chunk_elems = 0; ctx->n_chunks = 0; ctx->h_cg_ptr.clear();
    if (ctx->d_emat_chunk) { cudaFree(ctx->d_emat_chunk); ctx->d_emat_chunk = nullptr; }
    if (chunk_env > 0 && n_elems > 2 * chunk_env) {
      const int CE = chunk_env, nch = (n_elems + CE - 1) / CE;
      const int npair = nn * nn;
      std::vector<std::vector<unsigned>> blk(nch);
      std::vector<std::vector<int>> kk(nch);
      for (int b = 0; b < ctx->nnzb; ++b) {
        int k = gptr[b];
        const int k1 = gptr[b + 1];
        bool first = true;
        if (k == k1) {   // a block without contributions (isolated node): cleared by chunk 0
          blk[0].push_back(unsigned(b) | 0x80000000u);
          kk[0].push_back(k); kk[0].push_back(k);
        }
        while (k < k1) {
          const int c = (gsrc[k] / npair) / CE;
          int k2 = k + 1;
          while (k2 < k1 && (gsrc[k2] / npair) / CE == c) ++k2;
          blk[c].push_back(unsigned(b) | (first ? 0x80000000u : 0u));
          kk[c].push_back(k); kk[c].push_back(k2);
          first = false;
          k = k2;
        }
      }
      std::vector<unsigned> cblk;
      std::vector<int> ck;
      ctx->h_cg_ptr.assign(nch + 1, 0);
      for (int c = 0; c < nch; ++c) {
        cblk.insert(cblk.end(), blk[c].begin(), blk[c].end());
        ck.insert(ck.end(), kk[c].begin(), kk[c].end());
        ctx->h_cg_ptr[c + 1] = int(cblk.size());
      }
      ctx->h_cg_blk = cblk; ctx->cg_end_rows = -1;
      if (!upload(ctx, &ctx->d_cg_blk, cblk)) return C8_ERR_CUDA;
      if (!upload(ctx, &ctx->d_cg_k, ck)) return C8_ERR_CUDA;
      if (!upload(ctx, &ctx->d_cg_ptr, ctx->h_cg_ptr)) return C8_ERR_CUDA;
      ctx->chunk_elems = CE; ctx->n_chunks = nch;
    }
  }
  if (ctx->d_emat) { cudaFree(ctx->d_emat); ctx->d_emat = nullptr; ctx->emat_elems = 0; }
  if (!upload(ctx, &ctx->d_conn, ctx->h_conn)) return C8_ERR_CUDA;
  if (!upload(ctx, &ctx->d_coords, ctx->h_coords)) return C8_ERR_CUDA;
  if (!upload(ctx, &ctx->d_rowptr, ctx->h_rowptr)) return C8_ERR_CUDA;
  if (!upload(ctx, &ctx->d_colind, ctx->h_colind)) return C8_ERR_CUDA;
  if (!upload(ctx, &ctx->d_eoff, eoff)) return C8_ERR_CUDA;
  if (elem_set && ctx->n_es > 1) {
    std::vector<int> es(elem_set, elem_set + n_elems);
    for (int v : es) C8_REQUIRE(ctx, v >= 0 && v < ctx->n_es, "element set id out of range");
    if (!upload(ctx, &ctx->d_elem_es, es)) return C8_ERR_CUDA;
  } else {
    if (ctx->d_elem_es) cudaFree(ctx->d_elem_es);
    ctx->d_elem_es = nullptr;
  }
  ctx->n_owned_nodes = n_nodes;
  ctx->n_owned_elems = n_elems;
  // a new mesh is a one-part mesh until c8_set_partition / c8_set_halo_plan say otherwise: the old
  // halo plan, transport and hooks describe another node numbering
  c8_comm_release(ctx);
  ctx->halo_cb = nullptr; ctx->allreduce_cb = nullptr; ctx->comm_user = nullptr;
  ctx->comm_capturable = false;
  c8_linalg_invalidate(ctx);
  ctx->xi_ld = (long long)((n_elems + 31) / 32) * 32;  // 256-byte aligned component rows
  return C8_OK;
}

int c8_set_assembly_chunk(c8_ctx* ctx, int chunk_elems) {
  C8_REQUIRE(ctx, chunk_elems >= -1, "chunk size must be >= 0 (0: one pass), or -1 for the default");
  ctx->chunk_request = chunk_elems;
  return C8_OK;
}

int c8_set_model(c8_ctx* ctx, int global_type, int local_type, const double* params,
                 int max_iters, double abs_tol, double rel_tol, double stab_mult,
                 double thickness) {
  C8_REQUIRE(ctx, ctx->dim != 0, "c8_set_mesh must be called first");
  const int mech = (global_type == C8_MECHANICS) ? 0 : 1;
  const KernelTable* kt = find_kernel_table(ctx->dim, mech, local_type);
  C8_REQUIRE(ctx, kt != nullptr,
             "no kernels for this (dim, global residual, local residual) combination");
  ctx->kt = kt;
  c8_linalg_invalidate(ctx);
  ctx->global_type = global_type;
  ctx->local_type = local_type;
  ctx->model.npar = kt->npar;
  ctx->model.max_iters = max_iters;
  ctx->model.abs_tol = abs_tol;
  ctx->model.rel_tol = rel_tol;
  ctx->model.stab_mult = stab_mult;
  ctx->model.thickness = thickness;
  // (re)allocate the resident system for this block size
  const size_t nA = size_t(ctx->nnzb) * kt->nb * kt->nb, nb_ = size_t(ctx->n_nodes) * kt->nb;
  const size_t nxi = size_t(ctx->xi_ld) * kt->nxi;
  double** bufs[] = {&ctx->d_A, &ctx->d_b, &ctx->d_x, &ctx->d_xp, &ctx->d_xi, &ctx->d_xip};
  const size_t sizes[] = {nA, nb_, nb_, nb_, nxi, nxi};
  for (int k = 0; k < 6; ++k) {
    if (*bufs[k]) cudaFree(*bufs[k]);
    *bufs[k] = nullptr;
    C8_CUDA(ctx, cudaMalloc(bufs[k], std::max<size_t>(sizes[k], 1) * sizeof(double)));
    C8_CUDA(ctx, cudaMemsetAsync(*bufs[k], 0, std::max<size_t>(sizes[k], 1) * sizeof(double), ctx->stream));
  }
  return c8_set_params(ctx, params);
}

int c8_set_params(c8_ctx* ctx, const double* params) {
  C8_REQUIRE(ctx, ctx->kt != nullptr, "c8_set_model must be called first");
  const size_t n = size_t(ctx->n_es) * ctx->kt->npar;
  if (n > ctx->params_cap) {   // another model / more element sets than the first call sized it for
    if (ctx->d_params) cudaFree(ctx->d_params);
    ctx->d_params = nullptr; ctx->params_cap = 0; ctx->model.params = nullptr;
    C8_CUDA(ctx, cudaMalloc(&ctx->d_params, n * sizeof(double)));
    ctx->params_cap = n;
  }
  C8_CUDA(ctx, cudaMemcpyAsync(ctx->d_params, params, n * sizeof(double), cudaMemcpyHostToDevice,
                               ctx->stream));
  C8_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  ctx->model.params = ctx->d_params;
  ctx->h_params.assign(params, params + n);
  return C8_OK;
}

int c8_get_params(c8_ctx* ctx, double* params_host) {
  C8_REQUIRE(ctx, ctx->kt != nullptr, "c8_set_model must be called first");
  std::memcpy(params_host, ctx->h_params.data(), ctx->h_params.size() * sizeof(double));
  return C8_OK;
}

int c8_info(c8_ctx* ctx, int64_t* out) {
  C8_REQUIRE(ctx, ctx->kt != nullptr, "c8_set_model must be called first");
  const KernelTable* k = ctx->kt;
  const int64_t v[12] = {ctx->dim, k->nn, k->nb, k->nx, k->nxi, k->npar, ctx->n_elems,
                         ctx->n_nodes, ctx->nnzb, int64_t(ctx->n_nodes) * k->nb, k->group,
                         ctx->xi_ld};
  std::memcpy(out, v, sizeof(v));
  return C8_OK;
}

int c8_bsr_pattern(c8_ctx* ctx, int32_t* rowptr, int32_t* colind) {
  std::memcpy(rowptr, ctx->h_rowptr.data(), ctx->h_rowptr.size() * sizeof(int));
  std::memcpy(colind, ctx->h_colind.data(), ctx->h_colind.size() * sizeof(int));
  return C8_OK;
}
int c8_bsr_pattern_dev(c8_ctx* ctx, const int32_t** rowptr, const int32_t** colind,
                       const int32_t** eoff) {
  if (rowptr) *rowptr = ctx->d_rowptr;
  if (colind) *colind = ctx->d_colind;
  if (eoff) *eoff = ctx->d_eoff;
  return C8_OK;
}

int c8_csr_block_size(c8_ctx* ctx, int i, int j, int64_t* n_rows, int64_t* nnz) {
  C8_REQUIRE(ctx, ctx->kt != nullptr, "c8_set_model must be called first");
  C8_REQUIRE(ctx, i < num_resid(ctx) && j < num_resid(ctx), "block index out of range");
  int e0, ni, f0, nj;
  residual_split(ctx, i, &e0, &ni);
  residual_split(ctx, j, &f0, &nj);
  *n_rows = int64_t(ctx->n_nodes) * ni;
  *nnz = int64_t(ctx->nnzb) * ni * nj;
  return C8_OK;
}
int c8_csr_block_pattern(c8_ctx* ctx, int i, int j, int32_t* rowptr, int32_t* colind) {
  C8_REQUIRE(ctx, ctx->kt != nullptr, "c8_set_model must be called first");
  int e0, ni, f0, nj;
  residual_split(ctx, i, &e0, &ni);
  residual_split(ctx, j, &f0, &nj);
  size_t o = 0;
  rowptr[0] = 0;
  for (int n = 0; n < ctx->n_nodes; ++n)
    for (int ei = 0; ei < ni; ++ei) {
      for (int k = ctx->h_rowptr[n]; k < ctx->h_rowptr[n + 1]; ++k)
        for (int ej = 0; ej < nj; ++ej) colind[o++] = ctx->h_colind[k] * nj + ej;
      rowptr[n * ni + ei + 1] = int(o);
    }
  return C8_OK;
}
int c8_csr_block_values(c8_ctx* ctx, int i, int j, const double* bsr, double* vals_host) {
  C8_REQUIRE(ctx, ctx->kt != nullptr, "c8_set_model must be called first");
  int e0, ni, f0, nj;
  residual_split(ctx, i, &e0, &ni);
  residual_split(ctx, j, &f0, &nj);
  const size_t nnz = size_t(ctx->nnzb) * ni * nj;
  double* d = stage(ctx, nnz * sizeof(double));
  C8_REQUIRE(ctx, d != nullptr, "staging allocation failed");
  const int rows = ctx->n_nodes * ni;
  k_bsr_to_block_csr<<<(rows + 127) / 128, 128, 0, ctx->stream>>>(bsr, ctx->d_rowptr, d,
                                                                   ctx->n_nodes, ctx->kt->nb, e0,
                                                                   ni, f0, nj);
  C8_CUDA(ctx, cudaGetLastError());
  C8_CUDA(ctx, cudaMemcpyAsync(vals_host, d, nnz * sizeof(double), cudaMemcpyDeviceToHost,
                               ctx->stream));
  C8_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return C8_OK;
}

// ---- layout helpers ---------------------------------------------------------------
int c8_pack_x(c8_ctx* ctx, const double* u, const double* p, double* x_dev) {
  C8_REQUIRE(ctx, ctx->kt != nullptr, "c8_set_model must be called first");
  const int nb = ctx->kt->nb, dim = ctx->dim, n = ctx->n_nodes;
  const size_t nu = size_t(n) * dim, np = (nb > dim) ? size_t(n) : 0;
  double* d = stage(ctx, (nu + np) * sizeof(double));
  C8_REQUIRE(ctx, d != nullptr, "staging allocation failed");
  C8_CUDA(ctx, cudaMemcpyAsync(d, u, nu * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  if (np) {
    C8_REQUIRE(ctx, p != nullptr, "pressure field required for the mixed formulation");
    C8_CUDA(ctx, cudaMemcpyAsync(d + nu, p, np * sizeof(double), cudaMemcpyHostToDevice,
                                 ctx->stream));
  }
  k_interleave<<<(n * nb + 255) / 256, 256, 0, ctx->stream>>>(d, d + nu, x_dev, n, dim, nb);
  C8_CUDA(ctx, cudaGetLastError());
  C8_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // staging buffer is reused
  return C8_OK;
}
int c8_unpack_x(c8_ctx* ctx, const double* x_dev, double* u, double* p) {
  C8_REQUIRE(ctx, ctx->kt != nullptr, "c8_set_model must be called first");
  const int nb = ctx->kt->nb, dim = ctx->dim, n = ctx->n_nodes;
  const size_t nu = size_t(n) * dim, np = (nb > dim) ? size_t(n) : 0;
  double* d = stage(ctx, (nu + np) * sizeof(double));
  C8_REQUIRE(ctx, d != nullptr, "staging allocation failed");
  k_deinterleave<<<(n * nb + 255) / 256, 256, 0, ctx->stream>>>(x_dev, d, d + nu, n, dim, nb);
  C8_CUDA(ctx, cudaGetLastError());
  C8_CUDA(ctx, cudaMemcpyAsync(u, d, nu * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  if (np && p)
    C8_CUDA(ctx, cudaMemcpyAsync(p, d + nu, np * sizeof(double), cudaMemcpyDeviceToHost,
                                 ctx->stream));
  C8_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return C8_OK;
}
int c8_pack_xi(c8_ctx* ctx, const double* xi_aos, double* xi_dev) {
  C8_REQUIRE(ctx, ctx->kt != nullptr, "c8_set_model must be called first");
  const int nxi = ctx->kt->nxi, n = ctx->n_elems;
  const size_t cnt = size_t(n) * nxi;
  double* d = stage(ctx, cnt * sizeof(double));
  C8_REQUIRE(ctx, d != nullptr, "staging allocation failed");
  C8_CUDA(ctx, cudaMemcpyAsync(d, xi_aos, cnt * sizeof(double), cudaMemcpyHostToDevice,
                               ctx->stream));
  k_aos_to_soa<<<(unsigned)((cnt + 255) / 256), 256, 0, ctx->stream>>>(d, xi_dev, n, nxi,
                                                                        ctx->xi_ld);
  C8_CUDA(ctx, cudaGetLastError());
  C8_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return C8_OK;
}
int c8_unpack_xi(c8_ctx* ctx, const double* xi_dev, double* xi_aos) {
  C8_REQUIRE(ctx, ctx->kt != nullptr, "c8_set_model must be called first");
  const int nxi = ctx->kt->nxi, n = ctx->n_elems;
  const size_t cnt = size_t(n) * nxi;
  double* d = stage(ctx, cnt * sizeof(double));
  C8_REQUIRE(ctx, d != nullptr, "staging allocation failed");
  k_soa_to_aos<<<(unsigned)((cnt + 255) / 256), 256, 0, ctx->stream>>>(xi_dev, d, n, nxi,
                                                                        ctx->xi_ld);
  C8_CUDA(ctx, cudaGetLastError());
  C8_CUDA(ctx, cudaMemcpyAsync(xi_aos, d, cnt * sizeof(double), cudaMemcpyDeviceToHost,
                               ctx->stream));
  C8_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return C8_OK;
}
int c8_init_xi(c8_ctx* ctx, double* xi_dev) {
  C8_REQUIRE(ctx, ctx->kt != nullptr, "c8_set_model must be called first");
  ctx->kt->init_xi(xi_dev, ctx->xi_ld, ctx->n_elems, ctx->stream);
  C8_CUDA(ctx, cudaGetLastError());
  return C8_OK;
}

// ---- forward hot path -----------------------------------------------------------

static int forward_impl(c8_ctx* ctx, const double* x, const double* xp, const double* xip,
                        double* xi, double* A, double* b, int8_t* path, double* eJ, double* eR,
                        int transpose, int* n_failed) {
  C8_REQUIRE(ctx, ctx->kt != nullptr, "c8_set_model must be called first");
  FwdArgs a{};
  a.mesh = ctx->mesh_args();
  a.model = ctx->model;
  a.x = x; a.x_prev = xp; a.xi_prev = xip; a.xi = xi; a.xi_ld = ctx->xi_ld;
  a.vals = A; a.b = b; a.path = (signed char*)path; a.n_failed = ctx->d_nfailed;
  a.elem_J = eJ; a.elem_R = eR;
  if (A && !c8::chunked_assembly(ctx, a)) {
    a.emat = element_scratch(ctx);
    if (!a.emat) return C8_ERR_CUDA;
  }
  (void)transpose;
  C8_CUDA(ctx, cudaMemsetAsync(ctx->d_nfailed, 0, 2 * sizeof(int), ctx->stream));
  ctx->kt->forward_jacobian(a, ctx->stream);
  C8_CUDA(ctx, cudaGetLastError());
  if (n_failed) {
    // the status is the reference's return value: the caller needs it before continuing
    int rc = c8::fetch_n_failed(ctx, n_failed);
    if (rc != C8_OK) return rc;
    if (*n_failed > 0) return C8_ERR_LOCAL_SOLVE;
  }
  return C8_OK;
}

}  // extern "C"

// The reference-facing call with host buffers (Newton iterate in, residual + local-solve status out;
// the matrix stays resident for the linear solve).  The residual, the local state and the status are
// final when the element kernel ends, so their way back to the host (de-interleave + D2H on a side
// stream) overlaps the BSR gather of the matrix.
int c8::forward_state_host(c8_ctx* ctx, const double* u, const double* p, double* b_u, double* b_p,
                           int* n_failed) {
  C8_REQUIRE(ctx, ctx->kt != nullptr, "c8_set_model must be called first");
  const int nb = ctx->kt->nb, dim = ctx->dim, n = ctx->n_nodes;
  const size_t nu = size_t(n) * dim, np = (nb > dim) ? size_t(n) : 0;
  int rc;
  // inputs: H2D into the input staging buffer + interleave, NO host synchronisation before the element kernel is
  // launched (the results leave through a second staging buffer on the side stream, and the call ends with both
  // streams drained, so the next call's copies cannot overtake anything)
  {
    C8_REQUIRE(ctx, np == 0 || p != nullptr, "pressure field required for the mixed formulation");
    double* din = stage(ctx, (nu + np) * sizeof(double));
    C8_REQUIRE(ctx, din != nullptr, "staging allocation failed");
    C8_CUDA(ctx, cudaMemcpyAsync(din, u, nu * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    if (np) C8_CUDA(ctx, cudaMemcpyAsync(din + nu, p, np * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    k_interleave<<<(n * nb + 255) / 256, 256, 0, ctx->stream>>>(din, din + nu, ctx->d_x, n, dim, nb);
    C8_CUDA(ctx, cudaGetLastError());
  }
  (void)rc;
  if (!ctx->side_stream) C8_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->side_stream, cudaStreamNonBlocking));
  if (!ctx->ev_elements) C8_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_elements, cudaEventDisableTiming));
  C8_CUDA(ctx, cudaMemsetAsync(ctx->d_b, 0, size_t(n) * nb * sizeof(double), ctx->stream));
  C8_CUDA(ctx, cudaMemsetAsync(ctx->d_nfailed, 0, 2 * sizeof(int), ctx->stream));
  FwdArgs a{};
  a.mesh = ctx->mesh_args();
  a.model = ctx->model;
  a.x = ctx->d_x; a.x_prev = ctx->d_xp; a.xi_prev = ctx->d_xip; a.xi = ctx->d_xi; a.xi_ld = ctx->xi_ld;
  a.vals = ctx->d_A; a.b = ctx->d_b; a.n_failed = ctx->d_nfailed;   // A is overwritten block by block
  if (!chunked_assembly(ctx, a)) {
    a.emat = element_scratch(ctx);
    if (!a.emat) return C8_ERR_CUDA;
  }
  a.elements_done = ctx->ev_elements;
  ctx->kt->forward_jacobian(a, ctx->stream);
  C8_CUDA(ctx, cudaGetLastError());
  cudaStream_t side = ctx->side_stream;
  C8_CUDA(ctx, cudaStreamWaitEvent(side, ctx->ev_elements, 0));
  if ((nu + np) * sizeof(double) > ctx->stage_out_bytes) {
    if (ctx->d_stage_out) cudaFree(ctx->d_stage_out);
    ctx->d_stage_out = nullptr; ctx->stage_out_bytes = 0;
    C8_CUDA(ctx, cudaMalloc(&ctx->d_stage_out, (nu + np) * sizeof(double)));
    ctx->stage_out_bytes = (nu + np) * sizeof(double);
  }
  double* d = ctx->d_stage_out;
  k_deinterleave<<<(n * nb + 255) / 256, 256, 0, side>>>(ctx->d_b, d, d + nu, n, dim, nb);
  C8_CUDA(ctx, cudaMemcpyAsync(b_u, d, nu * sizeof(double), cudaMemcpyDeviceToHost, side));
  if (np && b_p) C8_CUDA(ctx, cudaMemcpyAsync(b_p, d + nu, np * sizeof(double), cudaMemcpyDeviceToHost, side));
  int nf = 0;
  if (ctx->allreduce_cb) {
    C8_CUDA(ctx, cudaStreamSynchronize(side));
    if ((rc = fetch_n_failed(ctx, &nf)) != C8_OK) return rc;   // summed over the parts, main stream
  } else {
    C8_CUDA(ctx, cudaMemcpyAsync(&nf, ctx->d_nfailed, sizeof(int), cudaMemcpyDeviceToHost, side));
    C8_CUDA(ctx, cudaStreamSynchronize(side));
    C8_CUDA(ctx, cudaStreamSynchronize(ctx->stream));   // the matrix is complete when the call returns
  }
  if (n_failed) *n_failed = nf;
  return nf > 0 ? C8_ERR_LOCAL_SOLVE : C8_OK;
}

extern "C" {

int c8_forward_jacobian(c8_ctx* ctx, const double* x, const double* xp, const double* xip,
                        double* xi, double* A, double* b, int8_t* path, int* n_failed) {
  return forward_impl(ctx, x, xp, xip, xi, A, b, path, nullptr, nullptr, 0, n_failed);
}
int c8_forward_jacobian_elem(c8_ctx* ctx, const double* x, const double* xp, const double* xip,
                             double* xi, double* A, double* b, int8_t* path, double* eJ,
                             double* eR, int* n_failed) {
  return forward_impl(ctx, x, xp, xip, xi, A, b, path, eJ, eR, 0, n_failed);
}

int c8_global_residual(c8_ctx* ctx, const double* x, const double* xp, const double* xi,
                       const double* xip, double* b) {
  C8_REQUIRE(ctx, ctx->kt != nullptr, "c8_set_model must be called first");
  FwdArgs a{};
  a.mesh = ctx->mesh_args();
  a.model = ctx->model;
  a.x = x; a.x_prev = xp; a.xi_prev = xip; a.xi = const_cast<double*>(xi); a.xi_ld = ctx->xi_ld;
  a.b = b;
  ctx->kt->global_residual(a, ctx->stream);
  C8_CUDA(ctx, cudaGetLastError());
  return C8_OK;
}

int c8_resident_matrix(c8_ctx* ctx, double** A) {
  *A = ctx->d_A;
  return C8_OK;
}

int c8_forward_jacobian_host(c8_ctx* ctx, const double* u, const double* p, const double* up,
                             const double* pp, const double* xip_h, double* xi_h, double* bu,
                             double* bp, int* n_failed) {
  C8_REQUIRE(ctx, ctx->kt != nullptr, "c8_set_model must be called first");
  const KernelTable* k = ctx->kt;
  int rc;
  if ((rc = c8_pack_x(ctx, u, p, ctx->d_x)) != C8_OK) return rc;
  if ((rc = c8_pack_x(ctx, up, pp, ctx->d_xp)) != C8_OK) return rc;
  if ((rc = c8_pack_xi(ctx, xip_h, ctx->d_xip)) != C8_OK) return rc;
  if ((rc = c8_pack_xi(ctx, xi_h, ctx->d_xi)) != C8_OK) return rc;
  C8_CUDA(ctx, cudaMemsetAsync(ctx->d_A, 0, size_t(ctx->nnzb) * k->nb * k->nb * sizeof(double),
                               ctx->stream));
  C8_CUDA(ctx, cudaMemsetAsync(ctx->d_b, 0, size_t(ctx->n_nodes) * k->nb * sizeof(double),
                               ctx->stream));
  int nf = 0;
  rc = forward_impl(ctx, ctx->d_x, ctx->d_xp, ctx->d_xip, ctx->d_xi, ctx->d_A, ctx->d_b, nullptr,
                    nullptr, nullptr, 0, &nf);
  if (n_failed) *n_failed = nf;
  if (rc != C8_OK) return rc;
  if ((rc = c8_unpack_xi(ctx, ctx->d_xi, xi_h)) != C8_OK) return rc;
  return c8_unpack_x(ctx, ctx->d_b, bu, bp);
}

}  // extern "C"
