/* c8b200.h -- C ABI of the B200-native calibr8 hot path.
 *
 * calibr8 has no plugin/FFI interface of its own; its seam for this path is the set of
 * free functions of source/calibr8/src/evaluations.hpp:23-191 operating on State/Disc.
 * Each entry point below names the reference function it replaces (file:line under
 * /root/reference/source/calibr8/src).  A calibr8 maintainer binds them from C++ with a
 * plain `extern "C"` include; see INTEGRATION.md for the stub.
 *
 * Conventions
 *   - return 0 on success, C8_ERR_LOCAL_SOLVE (-1) when at least one local constitutive
 *     Newton did not converge (the reference's `return -1`, evaluations.cpp:95-97),
 *     <= -2 on CUDA / usage errors (message via c8_last_error).
 *   - all floating point is fp64, indices int32, as in the reference (defines.hpp:17-20).
 *   - "dev" pointers are device pointers on the context's GPU; "host" pointers are host memory.
 *   - nodal fields are node-interleaved [n_nodes][NB] (u_0..u_{dim-1}[, p]); local state is
 *     structure-of-arrays [NXI][xi_ld]; the matrix is BSR with NB x NB blocks on the node graph
 *     (c8_bsr_pattern).  c8_csr_block_* export the reference's per-block CSR view
 *     (disc.cpp:356-459) for parity checks.
 */
#ifndef C8B200_H
#define C8B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct c8_ctx c8_ctx;

#define C8_OK 0
#define C8_ERR_LOCAL_SOLVE (-1)
#define C8_ERR_CUDA (-2)
#define C8_ERR_USAGE (-3)

/* global residual types: global_residual.cpp:619-630 */
#define C8_MECHANICS 0
#define C8_MECHANICS_PLANE_STRESS 1
/* local residual types: local_residual.cpp:892-933 (in-scope subset) */
#define C8_ELASTIC 0
#define C8_SMALL_J2 1
#define C8_SMALL_HILL 2
#define C8_SMALL_HILL_PLANE_STRESS 3
#define C8_HYPER_J2 4
#define C8_HYPER_J2_PLANE_STRESS 5
#define C8_SMALL_HILL_PLANE_STRAIN 6
#define C8_HYPER_J2_PLANE_STRAIN 7
#define C8_HYPO_HILL 8              /* hypo_hill.cpp, 3-D */
#define C8_HYPO_HILL_PLANE_STRAIN 9 /* hypo_hill_plane_strain.cpp */
#define C8_HYPO_HILL_PLANE_STRESS 10 /* hypo_hill_plane_stress.cpp */

/* ---- context / discretisation (replaces Disc::build_data, disc.cpp:563-583) ---- */
c8_ctx* c8_create(int device);
void c8_destroy(c8_ctx* ctx);
const char* c8_last_error(c8_ctx* ctx);
const char* c8_version(void);

/* host arrays: conn [n_elems][dim+1], coords [n_nodes][3], elem_set [n_elems] or NULL */
int c8_set_mesh(c8_ctx* ctx, int dim, int n_elems, int n_nodes, const int32_t* conn_host,
                const double* coords_host, const int32_t* elem_set_host, int n_elem_sets);

/* Two-phase assembly of eval_forward_jacobian (scatter_lhs, global_residual.cpp:556-586): the element
 * matrices of a CHUNK of elements go to a scratch that stays in L2 and are summed into the BSR blocks by
 * the chunk's gather pass, chunk after chunk (bit-identical to one pass over the whole mesh).  Call before
 * c8_set_mesh: elements per chunk (0 = one pass with a full-size scratch, -1 = default: $C8_ASM_CHUNK, else
 * one pass -- measured faster on B200, profiles/README.md).  Meshes below two chunks use one pass. */
int c8_set_assembly_chunk(c8_ctx* ctx, int chunk_elems);
/* replaces create_global_residual / create_local_residual + LocalResidual::init_params;
 * params_host [n_elem_sets][npar] in the model's parameter order */
int c8_set_model(c8_ctx* ctx, int global_type, int local_type, const double* params_host,
                 int local_max_iters, double local_abs_tol, double local_rel_tol,
                 double stabilization_multiplier, double thickness);
int c8_set_params(c8_ctx* ctx, const double* params_host); /* LocalResidual::set_params */
int c8_get_params(c8_ctx* ctx, double* params_host);       /* LocalResidual::params, [n_elem_sets][npar] */

/* out[0..11] = dim, nn, nb, nx, nxi, npar, n_elems, n_nodes, nnzb, n_dofs(=n_nodes*nb), group, xi_ld */
int c8_info(c8_ctx* ctx, int64_t* out12);

/* node-graph BSR pattern (host out): rowptr [n_nodes+1], colind [nnzb] */
int c8_bsr_pattern(c8_ctx* ctx, int32_t* rowptr_host, int32_t* colind_host);
/* device pointers of the resident pattern / element->block offsets (read-only) */
int c8_bsr_pattern_dev(c8_ctx* ctx, const int32_t** rowptr_dev, const int32_t** colind_dev,
                       const int32_t** eoff_dev);
/* the reference's block (i,j) CSR (disc.cpp:356-387): sizes, pattern, and values gathered
 * from a BSR value array on the device */
int c8_csr_block_size(c8_ctx* ctx, int i, int j, int64_t* n_rows, int64_t* nnz);
int c8_csr_block_pattern(c8_ctx* ctx, int i, int j, int32_t* rowptr_host, int32_t* colind_host);
int c8_csr_block_values(c8_ctx* ctx, int i, int j, const double* bsr_vals_dev, double* vals_host);

/* layout helpers between the reference's per-residual host arrays and the device layout */
int c8_pack_x(c8_ctx* ctx, const double* u_host, const double* p_host, double* x_dev);
int c8_unpack_x(c8_ctx* ctx, const double* x_dev, double* u_host, double* p_host);
int c8_pack_xi(c8_ctx* ctx, const double* xi_host_aos, double* xi_dev_soa);
int c8_unpack_xi(c8_ctx* ctx, const double* xi_dev_soa, double* xi_host_aos);
/* LocalResidual::init_variables (local_residual.cpp:34-74): initial local state */
int c8_init_xi(c8_ctx* ctx, double* xi_dev);

/* ---- the hot path, device pointers ---- */
/* eval_forward_jacobian, evaluations.cpp:12-154.  A_vals_dev is OVERWRITTEN with the assembled
 * Jacobian (two-phase assembly: element matrices -> scratch, then a deterministic gather per BSR
 * block; rows of ghost nodes are left untouched); b_dev is accumulated into (zero it first, like
 * LinearAlg::zero_all); xi_dev in: current-field values, out: solved.
 * path_dev (int8 per element) and the element-level outputs may be NULL.
 * n_failed (host out, may be NULL): number of local solves that failed. */
int c8_forward_jacobian(c8_ctx* ctx, const double* x_dev, const double* x_prev_dev,
                        const double* xi_prev_dev, double* xi_dev, double* A_vals_dev,
                        double* b_dev, int8_t* path_dev, int* n_failed);
/* same, also writing per-element Jacobians [n_elems][nx][nx] and residuals [n_elems][nx]
 * in the reference's element dof order (global_residual.cpp:21-23) -- parity-test hook */
int c8_forward_jacobian_elem(c8_ctx* ctx, const double* x_dev, const double* x_prev_dev,
                             const double* xi_prev_dev, double* xi_dev, double* A_vals_dev,
                             double* b_dev, int8_t* path_dev, double* elem_J_dev,
                             double* elem_R_dev, int* n_failed);
/* eval_global_residual, evaluations.cpp:156-259 (xi given, no local solve) */
int c8_global_residual(c8_ctx* ctx, const double* x_dev, const double* x_prev_dev,
                       const double* xi_dev, const double* xi_prev_dev, double* b_dev);

/* ---- the hot path, HOST buffers (copies inside): the drop-in call a calibr8 caller makes ----
 * u/p: the reference's per-residual nodal arrays; xi: [n_elems][nxi] packed like the apf IP
 * fields; b_u/b_p out; the assembled matrix stays resident on the device (c8_resident_matrix)
 * for the device-side linear solver, exactly where LinearAlg::A[GHOST] would be consumed. */
int c8_forward_jacobian_host(c8_ctx* ctx, const double* u_host, const double* p_host,
                             const double* u_prev_host, const double* p_prev_host,
                             const double* xi_prev_host, double* xi_host, double* b_u_host,
                             double* b_p_host, int* n_failed);
int c8_resident_matrix(c8_ctx* ctx, double** A_vals_dev);

/* ---- resident step state: what Primal::solve_at_step (primal.cpp:31-208) iterates on ----
 * c8_state_set_prev uploads the converged previous step (global + local fields) and starts the
 * current step as its copy (Disc::create_primal, disc.cpp:643-683).  c8_state_forward_jacobian
 * is one Newton-iteration's assembly: HOST nodal iterate in, HOST residual + status out; the
 * local state, its history and the Jacobian stay resident on the device. */
int c8_state_set_prev(c8_ctx* ctx, const double* u_prev_host, const double* p_prev_host,
                      const double* xi_prev_host);
int c8_state_forward_jacobian(c8_ctx* ctx, const double* u_host, const double* p_host,
                              double* b_u_host, double* b_p_host, int* n_failed);
int c8_state_get_xi(c8_ctx* ctx, double* xi_host);
int c8_state_ptrs(c8_ctx* ctx, double** x_dev, double** x_prev_dev, double** xi_dev,
                  double** xi_prev_dev, double** A_vals_dev, double** b_dev);

/* ---- adjoint pass and objective integrands ----
 * QoI description (QoI<T> of the reference).  NULL = average displacement.
 *   type 0 "average displacement"  avg_disp.cpp:15-33
 *   type 1 "calibration"           calibration.cpp:414-478 (displacement mismatch + plane load)
 *   type 2 "reaction mismatch"     reaction_mismatch.cpp:58-212 (plane load or torque; the caller
 *                                  passes balance_factor = dt_over_T = 1)
 *   type 3 "load mismatch"         load_mismatch.cpp:79-259 (normal load over the side-set facets)
 *   type 4 "surface mismatch"      surface_mismatch.cpp:32-117 (the caller passes weights = 2,
 *                                  inv_area = dt_over_T = 1: sum |u - u_meas|^2 w dv over the facets) */
typedef struct c8_qoi {
  int type;
  double weights[3];        /* "displacement weights" */
  double balance_factor;    /* "balance factor" */
  double dt_over_T;         /* step size / total time */
  double inv_area;          /* 1 / objective area */
  double load_mismatch;     /* total load - measured load of the step (after the preprocess pass) */
  int coord_idx;            /* "coordinate index" of the load plane */
  double coord_value, coord_tol;
  int reaction_force_comp;  /* "reaction force component" */
  const double* measured_dev;  /* [n_nodes][dim] measured displacement of the step */
  const int8_t* facet_dev;     /* [n_elems][3] local vertex ids of the side-set facet or -1 (2-D: 2 ids) */
  int compute_torque;          /* type 2 "compute torque": moment about axis reaction_force_comp */
  double normal_2d[2];         /* type 3 in 2-D: "2D surface normal" */
} c8_qoi;

/* History arrays: g [nxi][xi_ld], f [nx][xi_ld] (element dofs node-interleaved), phi [nxi][xi_ld]. */
/* eval_adjoint_jacobian, evaluations.cpp:349-526: AT_vals = dR/dx_total^T (overwritten), rhs +=,
 * g -= dJ/dxi */
int c8_adjoint_jacobian(c8_ctx* ctx, const c8_qoi* qoi, const double* x_dev,
                        const double* x_prev_dev, const double* xi_dev, const double* xi_prev_dev,
                        double* g_dev, const double* f_dev, double* AT_vals_dev, double* rhs_dev);
/* solve_adjoint_local, evaluations.cpp:528-659: phi, then f and g of the previous step */
int c8_adjoint_local(c8_ctx* ctx, const double* x_dev, const double* x_prev_dev,
                     const double* xi_dev, const double* xi_prev_dev, const double* z_dev,
                     double* phi_dev, double* g_dev, double* f_dev);
/* eval_qoi / preprocess_qoi, evaluations.cpp:662-756, 261-347.  mode 0: scalars_dev[0] += sum of
 * the per-point QoI; mode 1: scalars_dev[1] += load on the coordinate plane (calibration) */
int c8_qoi_value(c8_ctx* ctx, const c8_qoi* qoi, const double* x_dev, const double* x_prev_dev,
                 const double* xi_dev, const double* xi_prev_dev, int mode, double* scalars_dev);
/* eval_qoi_gradient, evaluations.cpp:758-925: grad_dev [n_elem_sets][npar] += derivative w.r.t.
 * every model parameter (the caller selects the active ones) */
int c8_qoi_gradient(c8_ctx* ctx, const c8_qoi* qoi, const double* x_dev, const double* x_prev_dev,
                    const double* xi_dev, const double* xi_prev_dev, const double* z_dev,
                    const double* phi_dev, double* grad_dev);

/* ---- virtual fields method (single-residual mechanics, i.e. mechanics_plane_stress) ----
 * eval_measured_residual[_and_grad], evaluations.cpp:1750-1973: local solve at u = measured,
 * b_dev += R; when dR_dev != NULL also the forward sensitivities: dR_dev [npar][n_dofs] +=
 * dR/dp_total and local_sens_dev [nxi*npar][xi_ld] (dxi/dp, in/out), for EVERY model parameter. */
int c8_vfm_forward(c8_ctx* ctx, const double* x_meas_dev, const double* x_meas_prev_dev,
                   const double* xi_prev_dev, double* xi_dev, double* b_dev, double* dR_dev,
                   double* local_sens_dev, int* n_failed);
/* eval_vfm_adjoint_gradient, evaluations.cpp:1975-2143: hist_dev [nxi][xi_ld] in/out,
 * grad_dev [npar] += s w^T dR/dp + phi^T dC/dp */
int c8_vfm_adjoint(c8_ctx* ctx, const double* x_meas_dev, const double* x_meas_prev_dev,
                   const double* xi_dev, const double* xi_prev_dev, const double* w_dev, double s,
                   double* hist_dev, double* grad_dev);

/* ---- global linear algebra on the device (replaces linear_alg.cpp / linear_solve.cpp) ---- */
#define C8_ERR_NOT_CONVERGED (-4)
int c8_spmv(c8_ctx* ctx, const double* A_vals_dev, const double* x_dev, double* y_dev);
int c8_dot(c8_ctx* ctx, const double* x_dev, const double* y_dev, double* out_host);
int c8_axpby(c8_ctx* ctx, double a, const double* x_dev, double b, double* y_dev, int64_t n);
/* apply_expression_primal_dbcs, dbcs.cpp:28-121: per constrained dof (node, eq) zero the row, keep
 * the diagonal, R = diag*(x - value) (adjoint: R = 0) */
int c8_apply_dbc(c8_ctx* ctx, double* A_vals_dev, double* R_dev, const double* x_dev,
                 const int32_t* dbc_node_dev, const int32_t* dbc_eq_dev, const double* dbc_val_dev,
                 int n_dbc, int is_adjoint);
/* apply_primal_tbcs, tbcs.cpp:17-98 (called from Primal::solve_at_step, primal.cpp:107, between the
 * assembly and the Dirichlet rows): R[n, d] -= T_d N_n w dv over the one-point side quadrature.
 * side_nodes_dev [n_sides][dim] local node ids of every side of the set, traction_dev [n_sides][dim]
 * the traction vector at each side's quadrature point (= centroid), residual 0 (displacement) rows */
int c8_apply_tbc(c8_ctx* ctx, double* R_dev, const int32_t* side_nodes_dev, const double* traction_dev,
                 int n_sides);
/* restarted GMRES(m), device-resident Arnoldi, right preconditioned (c8_set_preconditioner);
 * info_host[3] = iterations, final |r|, initial |r| */
int c8_gmres(c8_ctx* ctx, const double* A_vals_dev, const double* b_dev, double* x_dev,
             int restart, int max_iters, double rel_tol, double abs_tol, double* info_host);
void c8_linalg_release(c8_ctx* ctx);
void c8_linalg_invalidate(c8_ctx* ctx); /* after a mesh / partition / model change (called internally) */
/* right preconditioner of c8_gmres (replaces the Teko/MueLu/Ifpack2 stack of linear_solve.cpp:74-105):
 * block-Jacobi (one NB x NB inverse per node) or aggregation multigrid on the node graph with
 * block-Jacobi smoothing (default).  opts (may be NULL): nu_pre, nu_post, omega, over_correction,
 * coarsest_max_nodes. */
#define C8_PC_BLOCK_JACOBI 0
#define C8_PC_AMG 1
int c8_set_preconditioner(c8_ctx* ctx, int type, const double* opts, int n_opts);
/* out[0] = levels, out[1] = operator complexity, out[2..] = nodes per level */
int c8_preconditioner_info(c8_ctx* ctx, double* out, int n_out);

/* ---- partition (one context per GPU; replaces the OWNED/GHOST maps of disc.cpp:271-314 and the
 * import/export of linear_alg.cpp:53-86) ----
 * The local mesh of a part lists its owned nodes first, then its ghost nodes; its owned elements
 * first, then the halo elements (elements of other parts that touch an owned node).  Every part
 * assembles the COMPLETE rows of its owned nodes (halo elements are evaluated redundantly), so no
 * matrix or residual entries travel; only ghost entries of vectors (halo copy) and scalars do.
 * Objective integrands run over owned elements only. */
int c8_set_partition(c8_ctx* ctx, int n_owned_nodes, int n_owned_elems);
/* communication hooks: halo(user, vec_dev, nb) fills the ghost entries of a nodal vector
 * [n_nodes][nb]; allreduce(user, buf_dev, n) sums n doubles over the parts (both enqueue on the
 * context's stream) */
typedef void (*c8_halo_fn)(void* user, double* vec_dev, int nb);
typedef void (*c8_allreduce_fn)(void* user, double* buf_dev, int n);
int c8_set_comm(c8_ctx* ctx, c8_halo_fn halo, c8_allreduce_fn allreduce, void* user);
int c8_get_partition(c8_ctx* ctx, int* n_owned_nodes, int* n_owned_elems);
/* neighbour plan (host arrays): for neighbour k = 0..n_nbr-1 of rank nbr_rank[k], the owned local
 * nodes send_nodes[send_ptr[k]..send_ptr[k+1]) are sent, and the ghost nodes
 * n_owned_nodes + [recv_ptr[k], recv_ptr[k+1]) are received (ghosts are sorted by owner rank, then
 * by the owner's ordering, so each neighbour fills one contiguous range) */
int c8_set_halo_plan(c8_ctx* ctx, int n_nbr, const int32_t* nbr_rank, const int32_t* send_ptr,
                     const int32_t* send_nodes, const int32_t* recv_ptr);
/* transport 1: NCCL over NVLink (libnccl.so.2 bound at run time).  Rank 0 creates the id, the host
 * program distributes the 128 bytes (MPI_Bcast / torch.distributed), every rank calls init. */
int c8_nccl_unique_id(char* id_out128);
int c8_nccl_init(c8_ctx* ctx, const char* id128, int rank, int nranks);
/* transport 2: host-staged, for an MPI (PCU) or gloo host program without NCCL.
 * exchange(user, send_host, recv_host, nb): send_host holds [n_send][nb] packed in send_ptr order,
 * recv_host must be filled with [n_ghost][nb] in recv_ptr order; allreduce sums n host doubles. */
typedef void (*c8_host_exchange_fn)(void* user, const double* send_host, double* recv_host, int nb);
typedef void (*c8_host_allreduce_fn)(void* user, double* buf_host, int n);
int c8_set_comm_host(c8_ctx* ctx, c8_host_exchange_fn exchange, c8_host_allreduce_fn allreduce,
                     void* user);
/* rank of this part and number of parts, for the host-staged transport (c8_nccl_init takes them
 * itself): lets the multigrid preconditioner span the parts (per-level halo copies, replicated
 * coarse levels) instead of acting on each part's owned block only */
int c8_set_comm_rank(c8_ctx* ctx, int rank, int nranks);
int c8_halo(c8_ctx* ctx, double* vec_dev);              /* nodal vector [n_nodes][NB] */
int c8_halo_nb(c8_ctx* ctx, double* vec_dev, int nb);   /* nodal vector [n_nodes][nb], nb <= 4 */
int c8_allreduce(c8_ctx* ctx, double* buf_dev, int n);
int c8_comm_stats(c8_ctx* ctx, int64_t* out3 /* halo calls, allreduce calls, halo bytes */);
/* bit 0: the NVLink push halo (C8_P2P=1) is active on this part, bit 1: the one-CTA small allreduce */
int c8_comm_p2p_active(c8_ctx* ctx);
void c8_comm_release(c8_ctx* ctx);

int c8_get_coords(c8_ctx* ctx, double* coords_host /* [n_nodes][3] */);
int c8_get_conn(c8_ctx* ctx, int32_t* conn_host);
void* c8_get_stream(c8_ctx* ctx);

/* ---- roofline denominators measured on the box with the same timer as the kernels ---- */
int c8_bench_dfma(c8_ctx* ctx, int iters, double* tflops_out);
int c8_bench_copy(c8_ctx* ctx, double* gbs_out);

/* stream control: all launches go to this stream (default: a context-owned stream) */
int c8_set_stream(c8_ctx* ctx, void* cuda_stream);
int c8_synchronize(c8_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* C8B200_H */
