// Host-side step solvers above the C ABI (see calibr8_host.hpp) and their C driver API
// (c8h_*) used by Python tests / bench through ctypes.
#include "calibr8_host.hpp"

#include <algorithm>
#include <cstring>
#include <limits>

namespace c8host {

#define C8H_CUDA(call)                                                          \
  do {                                                                          \
    cudaError_t e_ = (call);                                                    \
    if (e_ != cudaSuccess)                                                      \
      throw std::runtime_error(std::string(#call) + ": " + cudaGetErrorString(e_)); \
  } while (0)

Problem::Problem(c8_ctx* c) : ctx(c) {
  int64_t info[12];
  check(c8_info(ctx, info), "c8_info");
  dim = int(info[0]); nn = int(info[1]); nb = int(info[2]); nx = int(info[3]); nxi = int(info[4]);
  npar = int(info[5]); n_elems = int(info[6]); n_nodes = int(info[7]); nnzb = int(info[8]);
  n_dofs = info[9]; xi_ld = info[11];
  check(c8_get_partition(ctx, &n_owned_nodes, &n_owned_elems), "c8_get_partition");
  coords.resize(size_t(n_nodes) * 3);
  c8_get_coords(ctx, coords.data());
  A.resize(size_t(nnzb) * nb * nb);
  b.resize(n_dofs); dx.resize(n_dofs); Adx.resize(n_dofs); work.resize(n_dofs);
  saved_xi.resize(size_t(xi_ld) * nxi);
}

Problem::~Problem() {
  if (d_dbc_node) cudaFree(d_dbc_node);
  if (d_dbc_eq) cudaFree(d_dbc_eq);
  if (d_dbc_val) cudaFree(d_dbc_val);
  for (double* p : cal.d_measured) if (p) cudaFree(p);
  if (cal.d_facet) cudaFree(cal.d_facet);
  for (Tbc& t : tbcs) {
    if (t.d_side_nodes) cudaFree(t.d_side_nodes);
    if (t.d_traction) cudaFree(t.d_traction);
  }
}

void Problem::check(int rc, const char* what) const {
  if (rc != C8_OK) throw std::runtime_error(std::string(what) + " failed (" + std::to_string(rc) +
                                            "): " + c8_last_error(ctx));
}

static double wall_now() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
PhaseTimer::PhaseTimer(Problem& p, int i) : P(p), id(i) {
  if (P.profile) { c8_synchronize(P.ctx); t0 = wall_now(); }
}
PhaseTimer::~PhaseTimer() {
  if (P.profile) { c8_synchronize(P.ctx); P.t_phase[id] += wall_now() - t0; }
}

void Problem::set_time(int n_steps, double dt) { num_steps = n_steps; step_size = dt; }

void Problem::add_dbc(int resid, int eq, const int* nodes, int n, const std::string& expr) {
  Dbc d{resid, eq, std::vector<int>(nodes, nodes + n), Expr(expr)};
  dbcs.push_back(std::move(d));
}

void Problem::finalize_dbcs() {
  std::vector<int> node, eq;
  for (const Dbc& d : dbcs)
    for (int nd : d.nodes) {
      if (nd >= n_owned_nodes) continue;  // rows of ghost nodes belong to another part
      node.push_back(nd);
      eq.push_back(d.resid == 0 ? d.eq : dim);  // residual 1 = pressure -> interleaved eq index dim
    }
  n_dbc = int(node.size());
  h_dbc_val.assign(n_dbc, 0.0);
  if (d_dbc_node) { cudaFree(d_dbc_node); cudaFree(d_dbc_eq); cudaFree(d_dbc_val); }
  d_dbc_node = d_dbc_eq = nullptr; d_dbc_val = nullptr;
  if (!n_dbc) return;
  C8H_CUDA(cudaMalloc(&d_dbc_node, n_dbc * sizeof(int)));
  C8H_CUDA(cudaMalloc(&d_dbc_eq, n_dbc * sizeof(int)));
  C8H_CUDA(cudaMalloc(&d_dbc_val, n_dbc * sizeof(double)));
  C8H_CUDA(cudaMemcpy(d_dbc_node, node.data(), n_dbc * sizeof(int), cudaMemcpyHostToDevice));
  C8H_CUDA(cudaMemcpy(d_dbc_eq, eq.data(), n_dbc * sizeof(int), cudaMemcpyHostToDevice));
}

void Problem::eval_dbc_values(double t) {
  size_t k = 0;
  for (const Dbc& d : dbcs)
    for (int nd : d.nodes) {
      if (nd >= n_owned_nodes) continue;
      const double* X = &coords[size_t(nd) * 3];
      h_dbc_val[k++] = d.expr(X[0], X[1], X[2], t);
    }
  if (n_dbc)
    C8H_CUDA(cudaMemcpy(d_dbc_val, h_dbc_val.data(), n_dbc * sizeof(double), cudaMemcpyHostToDevice));
}

void Problem::add_tbc(int resid, const int* side_nodes, int n_sides, const std::vector<std::string>& exprs) {
  if (resid != 0) throw std::runtime_error("traction bcs act on the displacement residual (index 0)");
  if (int(exprs.size()) < dim) throw std::runtime_error("traction bc: one expression per dimension");
  Tbc t;
  t.resid = resid;
  for (int d = 0; d < dim; ++d) t.exprs.emplace_back(exprs[d]);
  // keep the sides that touch an owned node (their other nodes are local: the part holds every
  // element touching an owned node); rows of ghost nodes are skipped by the kernel
  for (int s = 0; s < n_sides; ++s) {
    bool any_owned = false, all_local = true;
    for (int a = 0; a < dim; ++a) {
      const int nd = side_nodes[s * dim + a];
      all_local = all_local && nd >= 0 && nd < n_nodes;
      any_owned = any_owned || (nd >= 0 && nd < n_owned_nodes);
    }
    if (!any_owned) continue;
    if (!all_local) throw std::runtime_error("traction bc: a side of an owned node has a non-local node");
    for (int a = 0; a < dim; ++a) t.side_nodes.push_back(side_nodes[s * dim + a]);
  }
  t.n_sides = int(t.side_nodes.size()) / dim;
  if (t.n_sides) {
    C8H_CUDA(cudaMalloc(&t.d_side_nodes, t.side_nodes.size() * sizeof(int)));
    C8H_CUDA(cudaMemcpy(t.d_side_nodes, t.side_nodes.data(), t.side_nodes.size() * sizeof(int), cudaMemcpyHostToDevice));
    C8H_CUDA(cudaMalloc(&t.d_traction, t.side_nodes.size() * sizeof(double)));
  }
  tbcs.push_back(std::move(t));
}

void Problem::eval_tbc_values(double t) {
  for (Tbc& bc : tbcs) {
    if (!bc.n_sides) continue;
    std::vector<double> T(size_t(bc.n_sides) * dim);
    for (int s = 0; s < bc.n_sides; ++s) {
      double xq[3] = {0., 0., 0.};   // mapLocalToGlobal of the one-point rule = the side centroid
      for (int a = 0; a < dim; ++a)
        for (int k = 0; k < 3; ++k) xq[k] += coords[size_t(bc.side_nodes[s * dim + a]) * 3 + k] / dim;
      for (int d = 0; d < dim; ++d) T[size_t(s) * dim + d] = bc.exprs[d](xq[0], xq[1], xq[2], t);
    }
    C8H_CUDA(cudaMemcpy(bc.d_traction, T.data(), T.size() * sizeof(double), cudaMemcpyHostToDevice));
  }
}

void Problem::apply_tbcs(double* R) const {
  for (const Tbc& bc : tbcs)
    check(c8_apply_tbc(ctx, R, bc.d_side_nodes, bc.d_traction, bc.n_sides), "c8_apply_tbc");
}

// (re)size a per-step history to [0..num_steps] arrays of n doubles, zero-filled; arrays of the
// right size are reused (cudaMalloc / cudaFree of ~100 MB blocks cost milliseconds each and an
// objective is evaluated many times on one mesh)
static void reset_history(Problem& P, std::vector<DevVec>& h, size_t n) {
  cudaStream_t s = (cudaStream_t)c8_get_stream(P.ctx);
  const size_t want = size_t(P.num_steps) + 1;
  bool ok = h.size() == want;
  for (size_t k = 0; ok && k < h.size(); ++k) ok = h[k].size() == n;
  if (!ok) {
    h.clear();
    for (size_t k = 0; k < want; ++k) h.emplace_back(n);
  } else {
    for (DevVec& v : h) cudaMemsetAsync(v.get(), 0, n * sizeof(double), s);
  }
}

void Problem::allocate_history() {
  reset_history(*this, x, size_t(n_dofs));
  reset_history(*this, xi, size_t(xi_ld) * nxi);
  check(c8_init_xi(ctx, xi[0].get()), "c8_init_xi");
  check(c8_synchronize(ctx), "sync");
}

double Problem::dot(const double* a, const double* c) const {
  double v = 0.0;
  check(c8_dot(ctx, a, c, &v), "c8_dot");
  return v;
}
double Problem::norm(const double* v) const { return std::sqrt(dot(v, v)); }
void Problem::axpy(double a, const double* xs, double* y) const {
  check(c8_axpby(ctx, a, xs, 1.0, y, n_dofs), "c8_axpby");
}

// ---------------------------------------------------------------------------------------------
static c8_qoi make_qoi(const Problem& P, int step, double load_mismatch) {
  c8_qoi q{};
  q.type = P.qoi_type;
  if (P.qoi_type >= 1) {
    for (int k = 0; k < 3; ++k) q.weights[k] = P.cal.weights[k];
    q.balance_factor = P.cal.balance_factor;
    q.dt_over_T = P.step_size / (P.num_steps * P.step_size);
    q.inv_area = 1.0 / P.cal.area;
    q.load_mismatch = load_mismatch;
    q.coord_idx = P.cal.coord_idx;
    q.coord_value = P.cal.coord_value;
    q.coord_tol = P.cal.coord_tol;
    q.reaction_force_comp = P.cal.reaction_force_comp;
    q.measured_dev = (step >= 1 && step <= int(P.cal.d_measured.size())) ? P.cal.d_measured[step - 1] : nullptr;
    q.facet_dev = (const int8_t*)P.cal.d_facet;
    q.compute_torque = P.cal.compute_torque ? 1 : 0;
    q.normal_2d[0] = P.cal.normal_2d[0]; q.normal_2d[1] = P.cal.normal_2d[1];
  }
  if (P.qoi_type == 2 || P.qoi_type == 3) {
    // reaction / load mismatch: the integrand is mismatch * load and J += 1/2 mismatch^2
    // (src/reaction_mismatch.cpp:150-153,208-211): the calibration load term with unit factors
    q.balance_factor = 1.0; q.dt_over_T = 1.0;
  }
  if (P.qoi_type == 4) {
    // surface mismatch: sum |u - u_meas|^2 w dv over the facets (src/surface_mismatch.cpp:98-105) = the
    // calibration surface integrand 1/2 sum_d w_d (.)^2 / area * dt/T with w_d = 2, area = dt/T = 1
    for (int k = 0; k < 3; ++k) q.weights[k] = 2.0;
    q.inv_area = 1.0; q.dt_over_T = 1.0; q.balance_factor = 0.0;
  }
  return q;
}

static bool qoi_has_load(int type) { return type == 1 || type == 2 || type == 3; }

// preprocess_qoi + preprocess_finalize: total load on the plane -> load mismatch of the step
static double preprocess_load_mismatch(Problem& P, int step, double* total_load_out = nullptr) {
  if (!qoi_has_load(P.qoi_type)) return 0.0;
  c8_qoi q = make_qoi(P, step, 0.0);
  cudaStream_t s = (cudaStream_t)c8_get_stream(P.ctx);
  C8H_CUDA(cudaMemsetAsync(P.work.get(), 0, 2 * sizeof(double), s));
  P.check(c8_qoi_value(P.ctx, &q, P.x[step].get(), P.x[step - 1].get(), P.xi[step].get(),
                       P.xi[step - 1].get(), 1, P.work.get()), "c8_qoi_value(load)");
  P.check(c8_allreduce(P.ctx, P.work.get(), 2), "c8_allreduce");  // PCU_Add_Double, calibration.cpp:349-353
  double h[2];
  C8H_CUDA(cudaMemcpyAsync(h, P.work.get(), 2 * sizeof(double), cudaMemcpyDeviceToHost, s));
  C8H_CUDA(cudaStreamSynchronize(s));
  const double meas = (step - 1 < int(P.cal.load_data.size())) ? P.cal.load_data[step - 1] : 0.0;
  if (total_load_out) *total_load_out = h[1];
  return h[1] - meas;
}

bool Primal::assemble(int step, double* R_norm) {
  PhaseTimer pt(P, 0);
  cudaStream_t s = (cudaStream_t)c8_get_stream(P.ctx);
  C8H_CUDA(cudaMemsetAsync(P.A.get(), 0, P.A.size() * sizeof(double), s));
  C8H_CUDA(cudaMemsetAsync(P.b.get(), 0, P.b.size() * sizeof(double), s));
  int nf = 0;
  P.check(c8_halo(P.ctx, P.x[step].get()), "c8_halo");  // ghost entries of the Newton iterate
  const int rc = c8_forward_jacobian(P.ctx, P.x[step].get(), P.x[step - 1].get(),
                                     P.xi[step - 1].get(), P.xi[step].get(), P.A.get(), P.b.get(),
                                     nullptr, &nf);
  ++P.n_assemblies;
  if (rc == C8_ERR_LOCAL_SOLVE) return false;
  P.check(rc, "c8_forward_jacobian");
  P.apply_tbcs(P.b.get());   // apply_primal_tbcs(tbcs, disc, R_ghost, t), src/primal.cpp:107
  P.check(c8_apply_dbc(P.ctx, P.A.get(), P.b.get(), P.x[step].get(), P.d_dbc_node, P.d_dbc_eq,
                       P.d_dbc_val, P.n_dbc, 0), "c8_apply_dbc");
  *R_norm = P.norm(P.b.get());
  return true;
}

static double cubic_min(double phi_0, double dphi_0, double a, double phi, double slope_a) {
  const double d1 = dphi_0 + slope_a - 3. * (phi_0 - phi) / (0. - a);
  const double radicand = d1 * d1 - dphi_0 * slope_a;
  if (radicand < 0.) return 0.5 * a;
  const double d2 = std::sqrt(radicand);
  const double denom = slope_a - dphi_0 + 2. * d2;
  if (denom == 0.) return 0.5 * a;
  return a - a * (slope_a + d2 - d1) / denom;
}

void Primal::solve_at_step(int step) {
  cudaStream_t s = (cudaStream_t)c8_get_stream(P.ctx);
  const size_t xb = P.n_dofs * sizeof(double), xib = size_t(P.xi_ld) * P.nxi * sizeof(double);
  // create_primal(step): copy of step-1 (src/disc.cpp:643-683)
  C8H_CUDA(cudaMemcpyAsync(P.x[step].get(), P.x[step - 1].get(), xb, cudaMemcpyDeviceToDevice, s));
  C8H_CUDA(cudaMemcpyAsync(P.xi[step].get(), P.xi[step - 1].get(), xib, cudaMemcpyDeviceToDevice, s));
  P.eval_dbc_values(P.time(step));
  P.eval_tbc_values(P.time(step));
  int iter = 1;
  bool converged = false;
  double resid_norm_0 = 1.;
  const SolverParams& sp = P.sp;
  // When the line search accepts the step it evaluated last, A and R already hold the assembly at the
  // accepted point; the reference re-assembles the identical state at the top of the next iteration
  // (src/primal.cpp:95) -- skipped here (SolverParams::reuse_accepted_assembly).
  bool have_assembly = false;
  double rn_kept = 0.;
  while (iter <= sp.newton_max_iters && !converged) {
    double rn = 0.;
    if (have_assembly) {
      rn = rn_kept;
      have_assembly = false;
    } else if (!assemble(step, &rn)) {
      throw std::runtime_error("primal: local solve failed at the base point (step " +
                               std::to_string(step) + ")");
    }
    if (iter == 1) resid_norm_0 = rn;
    const double rel = rn / resid_norm_0;
    if (sp.print) std::printf("  step %d it %d |R| = %.6e rel %.3e\n", step, iter, rn, rel);
    if (rn < sp.newton_abs_tol || rel < sp.newton_rel_tol) { converged = true; break; }
    // solve A dx = -R
    P.check(c8_axpby(P.ctx, -1.0, P.b.get(), 0.0, P.work.get(), P.n_dofs), "scale_b");
    C8H_CUDA(cudaMemsetAsync(P.dx.get(), 0, xb, s));
    double info[3];
    PhaseTimer* pt1 = new PhaseTimer(P, 1);
    int rc = c8_gmres(P.ctx, P.A.get(), P.work.get(), P.dx.get(), sp.gmres_restart, sp.gmres_max_iters,
                      sp.linear_tol, 0.0, info);
    delete pt1;
    P.n_linear_iters += int(info[0]);
    if (rc != C8_OK && rc != C8_ERR_NOT_CONVERGED) P.check(rc, "c8_gmres");
    if (sp.print) std::printf("    gmres its %d |r| %.3e -> %.3e\n", int(info[0]), info[2], info[1]);
    P.axpy(1.0, P.dx.get(), P.x[step].get());  // add_to_soln
    // line search (src/primal.cpp:141-201, src/line_search.hpp:85-135)
    const double psi_0 = 0.5 * rn * rn, dpsi_0 = -2. * psi_0;
    C8H_CUDA(cudaMemcpyAsync(P.saved_xi.get(), P.xi[step].get(), xib, cudaMemcpyDeviceToDevice, s));
    double alpha_applied = 1.;
    bool last_eval_ok = false;
    double last_eval_rn = 0.;
    auto eval = [&](double alpha, double& phi, double& slope) -> bool {
      C8H_CUDA(cudaMemcpyAsync(P.xi[step].get(), P.saved_xi.get(), xib, cudaMemcpyDeviceToDevice, s));
      P.axpy(alpha - alpha_applied, P.dx.get(), P.x[step].get());
      alpha_applied = alpha;
      double ra = 0.;
      last_eval_ok = false;
      if (!assemble(step, &ra)) return false;
      last_eval_ok = true; last_eval_rn = ra;
      phi = 0.5 * ra * ra;
      PhaseTimer pt5(P, 5);
      P.check(c8_spmv(P.ctx, P.A.get(), P.dx.get(), P.Adx.get()), "c8_spmv");
      slope = P.dot(P.b.get(), P.Adx.get());
      return true;
    };
    const double armijo = sp.ls_c1 * dpsi_0;
    double alpha = 1., best_alpha = 1., best_phi = std::numeric_limits<double>::max();
    bool assembled_any = false, accepted = false;
    for (int n = 1; n <= sp.ls_max_evals; ++n) {
      double phi, slope;
      if (!eval(alpha, phi, slope)) { alpha *= 0.5; continue; }
      assembled_any = true;
      if (phi < best_phi) { best_phi = phi; best_alpha = alpha; }
      if (phi <= psi_0 + alpha * armijo) { accepted = true; break; }
      const double am = cubic_min(psi_0, dpsi_0, alpha, phi, slope);
      alpha = std::min(std::max(am, sp.ls_bmin * alpha), sp.ls_bmax * alpha);
    }
    if (!assembled_any) throw std::runtime_error("primal: line search could not assemble");
    const double a_final = accepted ? alpha : best_alpha;
    if (sp.print && (!accepted || a_final != 1.0))
      std::printf("    line search: alpha = %.3e%s\n", a_final,
                  accepted ? "" : (best_phi >= psi_0 ? " (max evals reached, ||R|| not reduced: increase 'max evals' "
                                                       "or reduce the load increment)" : " (max evals reached)"));
    if (sp.reuse_accepted_assembly && last_eval_ok && a_final == alpha_applied) {
      have_assembly = true;       // x, xi, A, R are exactly the state of the last trial
      rn_kept = last_eval_rn;
    }
    P.axpy(a_final - alpha_applied, P.dx.get(), P.x[step].get());
    ++iter;
  }
  if (!converged) throw std::runtime_error("Newton's method failed in " +
                                           std::to_string(sp.newton_max_iters) + " iterations");
}

double Primal::eval_qoi(int step) {
  cudaStream_t s = (cudaStream_t)c8_get_stream(P.ctx);
  double total_load = 0.0;
  const double mismatch = preprocess_load_mismatch(P, step, &total_load);
  if (int(P.cal.total_load.size()) < P.num_steps) P.cal.total_load.assign(P.num_steps, 0.0);
  P.cal.total_load[step - 1] = total_load;   // the "load out file" line of the step
  c8_qoi q = make_qoi(P, step, mismatch);
  C8H_CUDA(cudaMemsetAsync(P.work.get(), 0, 2 * sizeof(double), s));
  P.check(c8_qoi_value(P.ctx, &q, P.x[step].get(), P.x[step - 1].get(), P.xi[step].get(),
                       P.xi[step - 1].get(), 0, P.work.get()), "c8_qoi_value");
  P.check(c8_allreduce(P.ctx, P.work.get(), 2), "c8_allreduce");
  double h[2];
  C8H_CUDA(cudaMemcpyAsync(h, P.work.get(), 2 * sizeof(double), cudaMemcpyDeviceToHost, s));
  C8H_CUDA(cudaStreamSynchronize(s));
  double J = h[0];
  // Calibration::postprocess, src/calibration.cpp:373-393; Reaction/LoadMismatch::postprocess (unit factors)
  if (qoi_has_load(P.qoi_type)) J += 0.5 * q.balance_factor * q.dt_over_T * mismatch * mismatch;
  return J;
}

double Primal::solve_all() {
  P.allocate_history();
  double J = 0.;
  for (int step = 1; step <= P.num_steps; ++step) {
    solve_at_step(step);
    J += eval_qoi(step);
  }
  return J;
}

// ---------------------------------------------------------------------------------------------
void Adjoint::gradient(std::vector<double>& grad) {
  cudaStream_t s = (cudaStream_t)c8_get_stream(P.ctx);
  const int N = P.num_steps;
  const SolverParams& sp = P.sp;
  const size_t xb = P.n_dofs * sizeof(double);
  // reverse-sweep work arrays live in the Problem and are reused across gradient evaluations
  auto fit = [&](DevVec& v, size_t n) {
    if (v.size() != n) v.resize(n);
    else cudaMemsetAsync(v.get(), 0, n * sizeof(double), s);
  };
  fit(P.adj_g, size_t(P.xi_ld) * P.nxi); fit(P.adj_f, size_t(P.xi_ld) * P.nx);
  fit(P.adj_rhs, size_t(P.n_dofs)); fit(P.adj_grad, 64);
  DevVec &g = P.adj_g, &f = P.adj_f, &rhs = P.adj_rhs, &d_grad = P.adj_grad;
  reset_history(P, P.z, size_t(P.n_dofs));
  reset_history(P, P.phi, size_t(P.xi_ld) * P.nxi);
  int n_es = 1;
  grad.assign(size_t(n_es) * P.npar, 0.0);
  for (int step = N; step >= 1; --step) {
    const double mismatch = preprocess_load_mismatch(P, step);
    c8_qoi q = make_qoi(P, step, mismatch);
    const double *x = P.x[step].get(), *xp = P.x[step - 1].get(), *xi = P.xi[step].get(),
                 *xip = P.xi[step - 1].get();
    C8H_CUDA(cudaMemsetAsync(P.A.get(), 0, P.A.size() * sizeof(double), s));
    C8H_CUDA(cudaMemsetAsync(rhs.get(), 0, xb, s));
    {
      PhaseTimer pt2(P, 2);
      P.check(c8_adjoint_jacobian(P.ctx, &q, x, xp, xi, xip, g.get(), f.get(), P.A.get(), rhs.get()),
              "c8_adjoint_jacobian");
    }
    double* z = P.z[step].get();
    P.check(c8_apply_dbc(P.ctx, P.A.get(), rhs.get(), z, P.d_dbc_node, P.d_dbc_eq, P.d_dbc_val,
                         P.n_dbc, 1), "c8_apply_dbc(adjoint)");
    // iterative refinement of the linear adjoint solve, src/adjoint.cpp:113-180
    int iter = 1;
    double r0 = 1.;
    while (true) {
      C8H_CUDA(cudaMemsetAsync(P.dx.get(), 0, xb, s));
      double info[3];
      PhaseTimer* pt1 = new PhaseTimer(P, 1);
      int rc = c8_gmres(P.ctx, P.A.get(), rhs.get(), P.dx.get(), sp.gmres_restart,
                        sp.gmres_max_iters, 0.1 * sp.newton_rel_tol, 0.0, info);
      delete pt1;
      P.n_linear_iters += int(info[0]);
      if (rc != C8_OK && rc != C8_ERR_NOT_CONVERGED) P.check(rc, "c8_gmres(adjoint)");
      P.axpy(1.0, P.dx.get(), z);
      P.check(c8_spmv(P.ctx, P.A.get(), P.dx.get(), P.Adx.get()), "c8_spmv");
      P.axpy(-1.0, P.Adx.get(), rhs.get());
      const double rn = P.norm(rhs.get());
      if (iter == 1) r0 = rn;
      if (rn < sp.newton_abs_tol || (r0 > 0 && rn / r0 < sp.newton_rel_tol)) break;
      if (++iter > sp.newton_max_iters) throw std::runtime_error("adjoint solve failed to converge");
    }
    P.check(c8_halo(P.ctx, z), "c8_halo");
    {
      PhaseTimer pt3(P, 3);
      P.check(c8_adjoint_local(P.ctx, x, xp, xi, xip, z, P.phi[step].get(), g.get(), f.get()),
              "c8_adjoint_local");
    }
    PhaseTimer pt4(P, 4);
    C8H_CUDA(cudaMemsetAsync(d_grad.get(), 0, 64 * sizeof(double), s));
    P.check(c8_qoi_gradient(P.ctx, &q, x, xp, xi, xip, z, P.phi[step].get(), d_grad.get()),
            "c8_qoi_gradient");
    P.check(c8_allreduce(P.ctx, d_grad.get(), n_es * P.npar), "c8_allreduce");
    double h[64];
    C8H_CUDA(cudaMemcpyAsync(h, d_grad.get(), 64 * sizeof(double), cudaMemcpyDeviceToHost, s));
    C8H_CUDA(cudaStreamSynchronize(s));
    for (int k = 0; k < n_es * P.npar; ++k) grad[k] += h[k];
  }
}

}  // namespace c8host

// =============================================================================================
// C driver API over the host classes (for ctypes)
using namespace c8host;

struct c8h_problem {
  Problem P;
  std::string err;
  explicit c8h_problem(c8_ctx* ctx) : P(ctx) {}
};

c8host::Problem& c8h_problem_ref(c8h_problem* h) { return h->P; }

#define C8H_TRY(h, body)                                  \
  try { body; return 0; }                                 \
  catch (const std::exception& ex) { (h)->err = ex.what(); return -1; }

extern "C" {

c8h_problem* c8h_create(c8_ctx* ctx) {
  try { return new c8h_problem(ctx); } catch (...) { return nullptr; }
}
void c8h_destroy(c8h_problem* h) { delete h; }
const char* c8h_last_error(c8h_problem* h) { return h->err.c_str(); }

int c8h_set_time(c8h_problem* h, int num_steps, double step_size) {
  C8H_TRY(h, h->P.set_time(num_steps, step_size));
}
int c8h_add_dbc(c8h_problem* h, int resid, int eq, const int32_t* nodes, int n, const char* expr) {
  C8H_TRY(h, h->P.add_dbc(resid, eq, nodes, n, expr));
}
int c8h_finalize_dbcs(c8h_problem* h) { C8H_TRY(h, h->P.finalize_dbcs()); }
// side_nodes [n_sides][dim] local node ids; exprs: dim expressions in x, y, z, t separated by ';'
int c8h_add_tbc(c8h_problem* h, int resid, const int32_t* side_nodes, int n_sides, const char* exprs) {
  C8H_TRY(h, {
    std::vector<std::string> ex;
    std::string cur;
    for (const char* c = exprs; ; ++c) {
      if (*c == ';' || *c == 0) { ex.push_back(cur); cur.clear(); if (*c == 0) break; }
      else cur += *c;
    }
    h->P.add_tbc(resid, side_nodes, n_sides, ex);
  });
}
int c8h_set_solver(c8h_problem* h, int newton_max_iters, double abs_tol, double rel_tol,
                   int gmres_restart, int gmres_max_iters, double linear_tol, int print) {
  SolverParams& s = h->P.sp;
  s.newton_max_iters = newton_max_iters; s.newton_abs_tol = abs_tol; s.newton_rel_tol = rel_tol;
  s.gmres_restart = gmres_restart; s.gmres_max_iters = gmres_max_iters; s.linear_tol = linear_tol;
  s.print = print != 0;
  return 0;
}
// the deck's "line search" sublist of the global residual (src/line_search.hpp:33-49): "sufficient
// decrease", "min backtrack factor", "max backtrack factor", "max evals"
int c8h_set_line_search(c8h_problem* h, double c1, double backtrack_min, double backtrack_max, int max_evals) {
  C8H_TRY(h, {
    if (!(c1 > 0.) || !(backtrack_min > 0. && backtrack_min <= backtrack_max && backtrack_max < 1.) || max_evals < 1)
      throw std::runtime_error("line search: need c1 > 0, 0 < min <= max backtrack factor < 1, max evals >= 1");
    SolverParams& s = h->P.sp;
    s.ls_c1 = c1; s.ls_bmin = backtrack_min; s.ls_bmax = backtrack_max; s.ls_max_evals = max_evals;
  });
}
int c8h_set_qoi_avg_disp(c8h_problem* h) { h->P.qoi_type = 0; return 0; }
// measured_host: [num_steps][n_nodes][dim]; load_data [num_steps]; facet_host [n_elems][3] or NULL
int c8h_set_qoi_calibration(c8h_problem* h, double balance_factor, int coord_idx,
                            double coord_value, double coord_tol, int reaction_force_comp,
                            const double* weights, const double* measured_host,
                            const double* load_data, const int8_t* facet_host, double area) {
  C8H_TRY(h, {
    Problem& P = h->P;
    P.qoi_type = 1;
    CalibrationQoi& c = P.cal;
    c.enabled = true;
    c.balance_factor = balance_factor; c.coord_idx = coord_idx; c.coord_value = coord_value;
    c.coord_tol = coord_tol; c.reaction_force_comp = reaction_force_comp; c.area = area;
    for (int k = 0; k < 3; ++k) c.weights[k] = weights ? weights[k] : 1.0;
    c.load_data.assign(load_data, load_data + P.num_steps);
    for (double* p : c.d_measured) if (p) cudaFree(p);
    c.d_measured.assign(P.num_steps, nullptr);
    const size_t nm = size_t(P.n_nodes) * P.dim;
    for (int s = 0; s < P.num_steps; ++s) {
      C8H_CUDA(cudaMalloc(&c.d_measured[s], nm * sizeof(double)));
      C8H_CUDA(cudaMemcpy(c.d_measured[s], measured_host + size_t(s) * nm, nm * sizeof(double),
                          cudaMemcpyHostToDevice));
    }
    if (c.d_facet) { cudaFree(c.d_facet); c.d_facet = nullptr; }
    if (facet_host) {
      C8H_CUDA(cudaMalloc(&c.d_facet, size_t(P.n_elems) * 3));
      C8H_CUDA(cudaMemcpy(c.d_facet, facet_host, size_t(P.n_elems) * 3, cudaMemcpyHostToDevice));
    }
  });
}

// "reaction mismatch" (kind 2), "load mismatch" (3), "surface mismatch" (4) of src/qoi.cpp:272-287.
// facet_host [n_elems][3] local vertex ids of the side-set facet (-1: none; 2-D: two ids), kinds 3, 4;
// measured_host [num_steps][n_nodes][dim], kind 4; load_data [num_steps] = "load input file" or NULL
// (= "load out file" mode: mismatch against zero, the loads are read back with c8h_get_loads)
int c8h_set_qoi_mismatch(c8h_problem* h, int kind, int coord_idx, double coord_value, double coord_tol,
                         int reaction_force_comp, int compute_torque, const int8_t* facet_host,
                         const double* normal_2d, const double* measured_host, const double* load_data) {
  C8H_TRY(h, {
    Problem& P = h->P;
    if (kind < 2 || kind > 4) throw std::runtime_error("mismatch QoI kind must be 2, 3 or 4");
    if (kind == 4 && P.dim != 3) throw std::runtime_error("surface mismatch is a 3-D QoI (src/surface_mismatch.cpp:98-101)");
    if (kind >= 3 && !facet_host) throw std::runtime_error("load / surface mismatch need the side-set facets");
    P.qoi_type = kind;
    CalibrationQoi& c = P.cal;
    c.enabled = true;
    c.balance_factor = 1.; c.area = 1.;
    c.coord_idx = coord_idx; c.coord_value = coord_value; c.coord_tol = coord_tol;
    c.reaction_force_comp = reaction_force_comp; c.compute_torque = compute_torque != 0;
    c.normal_2d[0] = normal_2d ? normal_2d[0] : 0.; c.normal_2d[1] = normal_2d ? normal_2d[1] : 0.;
    c.load_data.clear();
    if (load_data) c.load_data.assign(load_data, load_data + P.num_steps);
    for (double* p : c.d_measured) if (p) cudaFree(p);
    c.d_measured.clear();
    if (measured_host) {
      c.d_measured.assign(P.num_steps, nullptr);
      const size_t nm = size_t(P.n_nodes) * P.dim;
      for (int s = 0; s < P.num_steps; ++s) {
        C8H_CUDA(cudaMalloc(&c.d_measured[s], nm * sizeof(double)));
        C8H_CUDA(cudaMemcpy(c.d_measured[s], measured_host + size_t(s) * nm, nm * sizeof(double), cudaMemcpyHostToDevice));
      }
    }
    if (c.d_facet) { cudaFree(c.d_facet); c.d_facet = nullptr; }
    if (facet_host) {
      C8H_CUDA(cudaMalloc(&c.d_facet, size_t(P.n_elems) * 3));
      C8H_CUDA(cudaMemcpy(c.d_facet, facet_host, size_t(P.n_elems) * 3, cudaMemcpyHostToDevice));
    }
  });
}
// total load of every step of the last primal solve (ReactionMismatch "load out file", load.dat)
int c8h_get_loads(c8h_problem* h, double* loads_out /* [num_steps] */) {
  C8H_TRY(h, {
    for (int s = 0; s < h->P.num_steps; ++s)
      loads_out[s] = s < int(h->P.cal.total_load.size()) ? h->P.cal.total_load[s] : 0.0;
  });
}

int c8h_primal_solve(c8h_problem* h, double* J_out) {
  C8H_TRY(h, { Primal pr(h->P); *J_out = pr.solve_all(); });
}
int c8h_adjoint_gradient(c8h_problem* h, double* grad_out /* [npar] */) {
  C8H_TRY(h, {
    Adjoint ad(h->P);
    std::vector<double> g;
    ad.gradient(g);
    std::memcpy(grad_out, g.data(), g.size() * sizeof(double));
  });
}
// host copies of the stored history: u [n_nodes*dim], p [n_nodes] (or NULL), xi [n_elems][nxi]
int c8h_get_step(c8h_problem* h, int step, double* u, double* p, double* xi) {
  C8H_TRY(h, {
    Problem& P = h->P;
    if (u) P.check(c8_unpack_x(P.ctx, P.x[step].get(), u, p), "c8_unpack_x");
    if (xi) P.check(c8_unpack_xi(P.ctx, P.xi[step].get(), xi), "c8_unpack_xi");
  });
}
int c8h_get_adjoint_step(c8h_problem* h, int step, double* zu, double* zp, double* phi) {
  C8H_TRY(h, {
    Problem& P = h->P;
    if (zu) P.check(c8_unpack_x(P.ctx, P.z[step].get(), zu, zp), "c8_unpack_x");
    if (phi) P.check(c8_unpack_xi(P.ctx, P.phi[step].get(), phi), "c8_unpack_xi");
  });
}
// enable != 0 switches the wall-clock phase profile on (adds stream syncs); out8 (may be NULL)
// receives the accumulated seconds per phase, see Problem::t_phase
int c8h_profile(c8h_problem* h, int enable, double* out8) {
  h->P.profile = enable != 0;
  if (out8) for (int k = 0; k < 8; ++k) out8[k] = h->P.t_phase[k];
  return 0;
}
// JSON description of create_local_residual(local_type, ndims) + create_global_residual(global_type, ndims)
// (host only; returns 0, or -1 with the message in out)
int c8h_describe_residuals(const char* local_type, const char* global_type, int ndims, char* out, int len) {
  auto put = [&](const std::string& s) { std::strncpy(out, s.c_str(), len - 1); out[len - 1] = 0; };
  try {
    const LocalResidual l = create_local_residual(local_type, ndims);
    const GlobalResidual g = create_global_residual(global_type, ndims);
    auto strs = [](const std::vector<std::string>& v) {
      std::string s = "[";
      for (size_t i = 0; i < v.size(); ++i) s += (i ? ", \"" : "\"") + v[i] + "\"";
      return s + "]";
    };
    auto ints = [](const std::vector<int>& v) {
      std::string s = "[";
      for (size_t i = 0; i < v.size(); ++i) s += (i ? ", " : "") + std::to_string(v[i]);
      return s + "]";
    };
    put("{\"local\": {\"c8_type\": " + std::to_string(l.c8_type) + ", \"resid_names\": " + strs(l.resid_names) +
        ", \"var_types\": " + ints(l.var_types) + ", \"num_eqs\": " + ints(l.num_eqs) + ", \"num_dofs\": " +
        std::to_string(l.num_dofs()) + ", \"param_names\": " + strs(l.param_names) + ", \"finite_deformation\": " +
        (l.finite_deformation ? "true" : "false") + ", \"z_stretch_idx\": " + std::to_string(l.z_stretch_idx) +
        "}, \"global\": {\"c8_type\": " + std::to_string(g.c8_type) + ", \"resid_names\": " + strs(g.resid_names) +
        ", \"num_eqs\": " + ints(g.num_eqs) + ", \"num_ip_sets\": " + std::to_string(g.num_ip_sets) + "}}");
    return 0;
  } catch (const std::exception& ex) {
    put(ex.what());
    return -1;
  }
}

// evaluate a boundary-condition / virtual-field expression at n points (xyz [n][3]) and time t;
// returns 0, or -1 with the message in err_out (may be NULL)
int c8h_eval_expr(const char* expr, const double* xyz, int n, double t, double* out, char* err_out,
                  int err_len) {
  try {
    Expr e(expr);
    for (int i = 0; i < n; ++i) out[i] = e(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], t);
    return 0;
  } catch (const std::exception& ex) {
    if (err_out && err_len > 0) { std::strncpy(err_out, ex.what(), err_len - 1); err_out[err_len - 1] = 0; }
    return -1;
  }
}
int c8h_stats(c8h_problem* h, int* n_assemblies, int* n_linear_iters) {
  *n_assemblies = h->P.n_assemblies;
  *n_linear_iters = h->P.n_linear_iters;
  return 0;
}

}  // extern "C"
