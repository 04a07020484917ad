// The C-ABI context: resident mesh, BSR pattern, element->block offsets, model
// parameters, scratch.  Replaces what Disc / LinearAlg / State hold for this
// path in the reference (src/disc.hpp:72-483, src/linear_alg.hpp:6-84).
#pragma once
#include <cuda_runtime.h>
#include <string>
#include <vector>
#include "kernel_table.h"

struct c8_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  cudaStream_t side_stream = nullptr;   // host-buffer calls: result copies that overlap the BSR gather
  cudaEvent_t ev_elements = nullptr;
  std::string err;

  // mesh
  int dim = 0, nn = 0, n_elems = 0, n_nodes = 0, n_es = 1;
  std::vector<int> h_conn;
  std::vector<double> h_coords;  // [n_nodes][dim]
  std::vector<int> h_rowptr, h_colind;  // node graph (BSR pattern)
  int nnzb = 0;
  int* d_conn = nullptr;
  double* d_coords = nullptr;
  int* d_elem_es = nullptr;
  int* d_rowptr = nullptr;
  int* d_colind = nullptr;
  int* d_eoff = nullptr;
  int* d_gptr = nullptr;       // gather plan: block -> contributing (element, node pair) slots
  int* d_gsrc = nullptr;
  double* d_emat = nullptr;    // element-matrix scratch [n_elems + 1][nx][nx]
  size_t emat_elems = 0;
  int emat_nx = 0;
  // chunked forward assembly (args.h): plan + a scratch of one chunk
  int chunk_elems = 0, n_chunks = 0;
  int chunk_request = -1;     // c8_set_assembly_chunk: elements per chunk, 0 one pass, -1 default / C8_ASM_CHUNK
  std::vector<int> h_cg_ptr, h_cg_end;   // entry range per chunk; end of the owned-row prefix
  std::vector<unsigned> h_cg_blk;
  int cg_end_rows = -1;                  // n_row_blocks the ends were computed for
  int* d_cg_ptr = nullptr;
  unsigned* d_cg_blk = nullptr;
  int* d_cg_k = nullptr;
  double* d_emat_chunk = nullptr;
  int emat_chunk_nx = 0;
  long long xi_ld = 0;
  // partition (multi-GPU): local nodes [0, n_owned_nodes) are owned, the rest are ghosts; local
  // elements [0, n_owned_elems) are owned, the rest are halo elements computed redundantly
  int n_owned_nodes = 0, n_owned_elems = 0;
  void (*halo_cb)(void*, double*, int) = nullptr;
  void (*allreduce_cb)(void*, double*, int) = nullptr;
  void* comm_user = nullptr;
  bool comm_capturable = false;   // the hooks only enqueue stream work (CUDA-graph capturable)

  // model
  const c8::KernelTable* kt = nullptr;
  int global_type = -1, local_type = -1;
  c8::ModelArgs model{};
  double* d_params = nullptr;
  size_t params_cap = 0;          // doubles allocated at d_params
  std::vector<double> h_params;   // [n_es][npar] host copy

  // scratch / resident system
  int* d_nfailed = nullptr;
  double* d_scalar = nullptr;  // 8 doubles of device scratch for reduced scalars
  double* d_A = nullptr;       // resident BSR values [nnzb*nb*nb]
  double* d_b = nullptr;       // [n_nodes*nb]
  double* d_x = nullptr, *d_xp = nullptr;
  double* d_xi = nullptr, *d_xip = nullptr;
  double* d_stage = nullptr;   // staging for host<->device layout conversion
  size_t stage_bytes = 0;
  double* d_stage_out = nullptr;   // second staging buffer: results on the side stream while d_stage holds the inputs
  size_t stage_out_bytes = 0;
  double* h_pinned = nullptr;
  size_t pinned_bytes = 0;

  c8::MeshArgs mesh_args() const {
    c8::MeshArgs m;
    m.n_elems = n_elems; m.n_nodes = n_nodes; m.conn = d_conn; m.coords = d_coords;
    m.n_row_nodes = n_owned_nodes;
    m.elem_es = d_elem_es; m.eoff = d_eoff; m.gptr = d_gptr; m.gsrc = d_gsrc; m.nnzb = nnzb;
    m.n_row_blocks = h_rowptr.empty() ? 0 : h_rowptr[n_owned_nodes];
    m.chunk_elems = chunk_elems; m.n_chunks = n_chunks;
    m.cg_ptr = d_cg_ptr; m.cg_blk = d_cg_blk; m.cg_k = d_cg_k;
    return m;
  }
};

namespace c8 {
int fail(c8_ctx* ctx, int code, const std::string& msg);
bool cuda_ok(c8_ctx* ctx, cudaError_t e, const char* what);
double* stage(c8_ctx* ctx, size_t bytes);
double* pinned(c8_ctx* ctx, size_t bytes);
int fetch_n_failed(c8_ctx* ctx, int* out);
double* element_scratch(c8_ctx* ctx);
// scratch + plan of the chunked forward assembly into a (fills emat, cg_ptr_host); false: not active
bool chunked_assembly(c8_ctx* ctx, c8::FwdArgs& a);
// eval_forward_jacobian on the resident state with HOST nodal buffers (c8_state_forward_jacobian)
int forward_state_host(c8_ctx* ctx, const double* u, const double* p, double* b_u, double* b_p, int* n_failed);
}  // namespace c8

#define C8_CUDA(ctx, call)                                        \
  do {                                                            \
    if (!c8::cuda_ok((ctx), (call), #call)) return C8_ERR_CUDA;   \
  } while (0)
#define C8_REQUIRE(ctx, cond, msg)                                \
  do {                                                            \
    if (!(cond)) return c8::fail((ctx), C8_ERR_USAGE, (msg));     \
  } while (0)
