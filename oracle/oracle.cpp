// ORACLE (test infrastructure only).
//
// CPU restatement of calibr8's hot-path evaluation loops
// (/root/reference/source/calibr8/src/evaluations.cpp) behind a small C API
// so that tests / bench.py's cpu_baseline leg / smoke() can drive it with
// ctypes.  Nothing in the product path (calibr8_b200/) links, imports or
// executes this file.
//
//   orc_forward_jacobian        <- eval_forward_jacobian        :12-154
//   orc_global_residual         <- eval_global_residual         :156-259
//   orc_adjoint_jacobian        <- eval_adjoint_jacobian        :349-526 (+ preprocess_qoi :261-347)
//   orc_adjoint_local           <- solve_adjoint_local          :528-659
//   orc_qoi                     <- eval_qoi                     :662-756
//   orc_qoi_gradient            <- eval_qoi_gradient            :758-925
//   orc_measured_residual       <- eval_measured_residual       :1750-1845
//   orc_measured_residual_grad  <- eval_measured_residual_and_grad :1847-1973
//   orc_vfm_adjoint_gradient    <- eval_vfm_adjoint_gradient    :1975-2143
//
// PARITY PINNING: the reference has no element-level golden vectors
// (SURVEY.md 8(c)).  This restatement is pinned end-to-end against the
// reference's own regression constants (test/primal/*.yaml.in `regression:
// QoI`) and finite-difference gradient checks by tests/test_oracle_golden.py.
//
// Field layout (all host, fp64): global residual i is [n_nodes * neq_i];
// local state is packed [n_elems * n_xi] in the reference's packing order
// (sym tensors 3-D (00,01,02,11,12,22), 2-D (00,01,11); src/local_residual.cpp:196-218).
// Matrices: one CSR per block (i,j), row = node*neq_i+eq, sorted columns
// (src/disc.cpp:356-387); values scattered through lower_bound offsets like
// src/disc.cpp:414-459.
#include <cstring>
#include <cstdio>
#include "models.hpp"
#include "qoi.hpp"

namespace orc {

#ifdef C8_ORACLE_COUNT
thread_local long long g_flops = 0;
#endif

struct Problem {
  Disc disc;
  int global_type = G_MECHANICS;
  bool mixed = true;
  double stab_mult = 1., thickness = 1.;
  int local_type = L_ELASTIC;
  std::unique_ptr<GlobalResidual<double>> g_d;
  std::unique_ptr<GlobalResidual<Fad>> g_f;
  std::unique_ptr<LocalResidual<double>> l_d;
  std::unique_ptr<LocalResidual<Fad>> l_f;
  CsrGraph ngraph;
  CsrGraph bgraph[2][2];
  std::vector<int> offsets[2][2];  // scatter offsets, src/disc.cpp:414-459
  int qoi_type = 0;                // 0 avg disp, 1 calibration, 2 reaction / 3 load / 4 surface mismatch
  CalibrationData cal;
  std::unique_ptr<QoI<double>> q_d;
  std::unique_ptr<QoI<Fad>> q_f;
  double time = 0., dt = 1.;
  int n_failed_elem = -1;

  int num_resid() const { return g_d->num_residuals(); }
  int neq(int i) const { return g_d->num_eqs(i); }

  void make_global() {
    if (global_type == G_MECHANICS) {
      g_d = std::make_unique<Mechanics<double>>(disc.dim, mixed, stab_mult);
      g_f = std::make_unique<Mechanics<Fad>>(disc.dim, mixed, stab_mult);
    } else {
      g_d = std::make_unique<MechanicsPlaneStress<double>>(disc.dim, thickness);
      g_f = std::make_unique<MechanicsPlaneStress<Fad>>(disc.dim, thickness);
    }
    int const nr = num_resid();
    for (int i = 0; i < nr; ++i)
      for (int j = 0; j < nr; ++j) {
        bgraph[i][j] = block_graph(ngraph, neq(i), neq(j));
        compute_scatter_offsets(i, j);
      }
  }
  void compute_scatter_offsets(int i, int j) {
    int const nn = disc.nn, ni = neq(i), nj = neq(j);
    int const stride = ni * nn * nj * nn;
    CsrGraph const& g = bgraph[i][j];
    offsets[i][j].assign(size_t(disc.n_elems) * stride, -1);
    for (int e = 0; e < disc.n_elems; ++e) {
      int* off = &offsets[i][j][size_t(e) * stride];
      for (int in = 0; in < nn; ++in)
        for (int ie = 0; ie < ni; ++ie) {
          int const row = disc.conn[size_t(e) * nn + in] * ni + ie;
          int const* rb = &g.colind[g.rowptr[row]];
          int const* re = &g.colind[g.rowptr[row + 1]];
          int const row_off = (in * ni + ie) * (nj * nn);
          for (int jn = 0; jn < nn; ++jn)
            for (int je = 0; je < nj; ++je) {
              int const col = disc.conn[size_t(e) * nn + jn] * nj + je;
              int const* it = std::lower_bound(rb, re, col);
              off[row_off + jn * nj + je] = int(it - &g.colind[0]);
            }
        }
    }
  }
};

// src/global_residual.cpp:556-586
static void scatter_lhs(Problem& P, GlobalResidual<Fad>& global, int elem,
                        EMatrix const& dtotal, double* const* vals) {
  int const nr = global.num_residuals(), nn = global.num_nodes();
  for (int i = 0; i < nr; ++i) {
    int const ni = global.num_eqs(i);
    for (int j = 0; j < nr; ++j) {
      int const nj = global.num_eqs(j);
      int const dofs_j = nn * nj;
      double* v = vals[i * nr + j];
      int const* off = &P.offsets[i][j][size_t(elem) * (ni * nn * dofs_j)];
      for (int in = 0; in < nn; ++in)
        for (int ie = 0; ie < ni; ++ie) {
          int const i_idx = global.dx_idx(i, in, ie);
          int const row_offset = (in * ni + ie) * dofs_j;
          for (int jn = 0; jn < nn; ++jn)
            for (int je = 0; je < nj; ++je) {
              int const j_idx = global.dx_idx(j, jn, je);
              v[off[row_offset + jn * nj + je]] += dtotal(i_idx, j_idx);
            }
        }
    }
  }
}

// src/global_residual.cpp:462-479
template <class T>
static void scatter_rhs(Problem& P, GlobalResidual<T>& global, int elem, EVector const& rhs,
                        double* const* RHS) {
  int const nn = global.num_nodes();
  for (int i = 0; i < global.num_residuals(); ++i)
    for (int n = 0; n < nn; ++n)
      for (int eq = 0; eq < global.num_eqs(i); ++eq) {
        int const row = P.disc.conn[size_t(elem) * nn + n] * global.num_eqs(i) + eq;
        RHS[i][row] += rhs[global.dx_idx(i, n, eq)];
      }
}

// src/global_residual.cpp:481-501 ; MV[i] is [num_params][n_rows_i]
static void scatter_sens(Problem& P, GlobalResidual<Fad>& global, int elem, EMatrix const& sens,
                         double* const* MV) {
  int const nn = global.num_nodes();
  int const np = sens.cols();
  for (int i = 0; i < global.num_residuals(); ++i) {
    size_t const nrows = size_t(P.disc.n_nodes) * global.num_eqs(i);
    for (int p = 0; p < np; ++p)
      for (int n = 0; n < nn; ++n)
        for (int eq = 0; eq < global.num_eqs(i); ++eq) {
          int const row = P.disc.conn[size_t(elem) * nn + n] * global.num_eqs(i) + eq;
          MV[i][p * nrows + row] += sens(global.dx_idx(i, n, eq), p);
        }
  }
}

struct QuadSets {
  std::vector<std::vector<QPoint>> sets;
  QuadSets(int dim, std::vector<int> const& orders) {
    for (int o : orders) sets.push_back(quadrature(dim, o));
  }
};

// ---------------------------------------------------------------------------
// eval_forward_jacobian, src/evaluations.cpp:12-154
// elem_dtotal (optional): [n_elems][n_x*n_x] sum over all ip sets of the
// element Jacobian, elem_R likewise -- element-level outputs for parity tests.
static int eval_forward_jacobian(Problem& P, double const* const* x, double const* const* x_prev,
                                 double* xi, double const* xi_prev, double* const* LHS,
                                 double* const* RHS, double* elem_dtotal, double* elem_R,
                                 int* path_out, int* iters_out, int elem_begin, int elem_end) {
  Disc const& disc = P.disc;
  auto& local = *P.l_f;
  auto& global = *P.g_f;
  global.set_time_info(P.time, P.dt);
  global.before_elems(disc);
  QuadSets Q(disc.dim, global.ip_sets());
  int nderivs = -1;
  for (int es = 0; es < disc.n_es; ++es) {
    local.before_elems(es, disc);
    for (int elem : disc.es_elems[es]) {
      if (elem < elem_begin || elem >= elem_end) continue;
      global.set_elem(elem);
      global.gather(x, x_prev);
      int const nx = global.num_dofs();
      for (size_t ip_set = 0; ip_set < Q.sets.size(); ++ip_set) {
        for (auto const& qp : Q.sets[ip_set]) {
          double const* iota = qp.xi;
          double const w = qp.w;
          double const dv = global.geom().dv;
          if (ip_set == 0) {
            global.interpolate(iota);
            local.gather(elem, xi, xi_prev);
            nderivs = local.seed_wrt_xi();
            int path = local.solve_nonlinear(global);
            if (iters_out) iters_out[elem] = local.last_iters;
            if (path == -1) { P.n_failed_elem = elem; return path; }
            if (path_out) path_out[elem] = path;
            local.scatter(elem, xi);
            EMatrix const dC_dxi = local.eigen_jacobian(nderivs);
            local.unseed_wrt_xi();
            nderivs = global.seed_wrt_x();
            global.interpolate(iota);
            local.evaluate(global);
            EMatrix const dC_dx = local.eigen_jacobian(nderivs);
            EMatrix const dxi_dx = full_piv_lu_solve(dC_dxi, neg(dC_dx));
            local.seed_wrt_x(dxi_dx);
          } else {
            nderivs = global.seed_wrt_x();
            global.interpolate(iota);
          }
          global.zero_residual();
          global.evaluate(local, iota, w, dv, int(ip_set));
          EMatrix const dtotal = global.eigen_jacobian(nderivs);
          EVector const elem_resid = global.eigen_residual();
          if (LHS) scatter_lhs(P, global, elem, dtotal, LHS);
          if (RHS) scatter_rhs(P, global, elem, elem_resid, RHS);
          if (elem_dtotal)
            for (int a = 0; a < nx; ++a)
              for (int b = 0; b < nx; ++b)
                elem_dtotal[(size_t(elem) * nx + a) * nx + b] += dtotal(a, b);
          if (elem_R)
            for (int a = 0; a < nx; ++a) elem_R[size_t(elem) * nx + a] += elem_resid[a];
          global.unseed_wrt_x();
        }
      }
    }
  }
  return 0;
}

// ---------------------------------------------------------------------------
// eval_global_residual, src/evaluations.cpp:156-259
static void eval_global_residual(Problem& P, double const* const* x, double const* const* x_prev,
                                 double const* xi, double const* xi_prev, double* const* RHS) {
  Disc const& disc = P.disc;
  auto& local = *P.l_d;
  auto& global = *P.g_d;
  global.set_time_info(P.time, P.dt);
  global.before_elems(disc);
  QuadSets Q(disc.dim, global.ip_sets());
  for (int es = 0; es < disc.n_es; ++es) {
    local.before_elems(es, disc);
    for (int elem : disc.es_elems[es]) {
      global.set_elem(elem);
      global.gather(x, x_prev);
      for (size_t ip_set = 0; ip_set < Q.sets.size(); ++ip_set)
        for (auto const& qp : Q.sets[ip_set]) {
          if (ip_set == 0) local.gather(elem, xi, xi_prev);
          global.interpolate(qp.xi);
          global.zero_residual();
          global.evaluate(local, qp.xi, qp.w, global.geom().dv, int(ip_set));
          scatter_rhs(P, global, elem, global.eigen_residual(), RHS);
        }
    }
  }
}

// ---------------------------------------------------------------------------
// preprocess_qoi, src/evaluations.cpp:261-347
template <class T>
static void preprocess_qoi(Problem& P, QoI<T>& qoi, LocalResidual<T>& local,
                           GlobalResidual<T>& global, double const* const* x,
                           double const* const* x_prev, double const* xi, double const* xi_prev,
                           int step) {
  Disc const& disc = P.disc;
  global.before_elems(disc);
  qoi.before_elems(disc, step);
  auto const qps = quadrature(disc.dim, 1);
  for (int es = 0; es < disc.n_es; ++es) {
    local.before_elems(es, disc);
    for (int elem : disc.es_elems[es]) {
      global.set_elem(elem);
      qoi.set_elem(elem);
      global.gather(x, x_prev);
      for (auto const& qp : qps) {
        global.interpolate(qp.xi);
        local.gather(elem, xi, xi_prev);
        qoi.preprocess(es, elem, global, local, qp.xi, qp.w, global.geom().dv);
      }
    }
  }
  qoi.preprocess_finalize(step);
}

// ---------------------------------------------------------------------------
// eval_adjoint_jacobian, src/evaluations.cpp:349-526
// g [n_elems*n_xi] (in/out), f [n_elems*n_x] (in)
static void eval_adjoint_jacobian(Problem& P, double const* const* x, double const* const* x_prev,
                                  double const* xi, double const* xi_prev, double* g,
                                  double const* f, double* const* LHS, double* const* RHS,
                                  int step) {
  Disc const& disc = P.disc;
  auto& local = *P.l_f;
  auto& global = *P.g_f;
  auto& qoi = *P.q_f;
  global.set_time_info(P.time, P.dt);
  preprocess_qoi(P, qoi, local, global, x, x_prev, xi, xi_prev, step);
  global.before_elems(disc);
  qoi.before_elems(disc, step);
  QuadSets Q(disc.dim, global.ip_sets());
  int nderivs = -1;
  for (int es = 0; es < disc.n_es; ++es) {
    local.before_elems(es, disc);
    int const nxi = local.num_dofs();
    for (int elem : disc.es_elems[es]) {
      global.set_elem(elem);
      qoi.set_elem(elem);
      global.gather(x, x_prev);
      int const nx = global.num_dofs();
      for (size_t ip_set = 0; ip_set < Q.sets.size(); ++ip_set)
        for (auto const& qp : Q.sets[ip_set]) {
          double const* iota = qp.xi;
          double const w = qp.w, dv = global.geom().dv;
          if (ip_set == 0) {
            global.interpolate(iota);
            local.gather(elem, xi, xi_prev);
            nderivs = local.seed_wrt_xi();
            local.evaluate(global);
            EMatrix const dC_dxi = local.eigen_jacobian(nderivs);
            local.unseed_wrt_xi();
            nderivs = global.seed_wrt_x();
            global.interpolate(iota);
            local.evaluate(global);
            EMatrix const dC_dx = local.eigen_jacobian(nderivs);
            EMatrix const dxi_dx = full_piv_lu_solve(dC_dxi, neg(dC_dx));
            local.seed_wrt_x(dxi_dx);
            global.zero_residual();
            global.evaluate(local, iota, w, dv, int(ip_set));
            EMatrix const dtotal = global.eigen_jacobian(nderivs);
            EMatrix const dtotalT = dtotal.transpose();
            scatter_lhs(P, global, elem, dtotalT, LHS);
            local.unseed_wrt_xi();
            qoi.evaluate(es, elem, global, local, iota, w, dv);
            EVector const dJ_dx = qoi.eigen_dvector(nderivs);
            global.unseed_wrt_x();
            nderivs = local.seed_wrt_xi();
            global.interpolate(iota);
            qoi.evaluate(es, elem, global, local, iota, w, dv);
            EVector const dJ_dxi = qoi.eigen_dvector(nderivs);
            local.unseed_wrt_xi();
            double* g_pt = &g[size_t(elem) * nxi];
            double const* f_pt = &f[size_t(elem) * nx];
            for (int a = 0; a < nxi; ++a) g_pt[a] -= dJ_dxi[a];
            EVector rhs(nx);
            for (int a = 0; a < nx; ++a) {
              double s = 0.;
              for (int b = 0; b < nxi; ++b) s += dxi_dx(b, a) * g_pt[b];
              rhs[a] = -dJ_dx[a] + f_pt[a] + s;
            }
            scatter_rhs(P, global, elem, rhs, RHS);
          } else {
            nderivs = global.seed_wrt_x();
            global.interpolate(iota);
            global.zero_residual();
            global.evaluate(local, iota, w, dv, int(ip_set));
            EMatrix const dtotalT = global.eigen_jacobian(nderivs).transpose();
            scatter_lhs(P, global, elem, dtotalT, LHS);
            global.unseed_wrt_x();  // (state hygiene; the reference re-seeds on the next pass)
          }
        }
    }
  }
}

// ---------------------------------------------------------------------------
// solve_adjoint_local, src/evaluations.cpp:528-659
static void solve_adjoint_local(Problem& P, double const* const* x, double const* const* x_prev,
                                double const* xi, double const* xi_prev, double const* const* z,
                                double* phi, double* g, double* f) {
  Disc const& disc = P.disc;
  auto& local = *P.l_f;
  auto& global = *P.g_f;
  global.set_time_info(P.time, P.dt);
  global.before_elems(disc);
  auto const qps = quadrature(disc.dim, 1);
  int nderivs = -1;
  for (int es = 0; es < disc.n_es; ++es) {
    local.before_elems(es, disc);
    int const nxi = local.num_dofs();
    for (int elem : disc.es_elems[es]) {
      global.set_elem(elem);
      global.gather(x, x_prev);
      int const nx = global.num_dofs();
      EVector const z_nodes = global.gather_adjoint(z);
      for (auto const& qp : qps) {
        double const* iota = qp.xi;
        double const w = qp.w, dv = global.geom().dv;
        global.interpolate(iota);
        local.gather(elem, xi, xi_prev);
        nderivs = local.seed_wrt_xi();
        global.zero_residual();
        global.evaluate(local, iota, w, dv, 0);
        local.evaluate(global);
        EMatrix const dC_dxiT = local.eigen_jacobian(nderivs).transpose();
        EMatrix const dR_dxiT = global.eigen_jacobian(nderivs).transpose();
        EVector rhs(nxi);
        EVector const t = matvec(dR_dxiT, z_nodes);
        for (int a = 0; a < nxi; ++a) rhs[a] = g[size_t(elem) * nxi + a] - t[a];
        EVector const phi_pt = full_piv_lu_solve(dC_dxiT, rhs);
        for (int a = 0; a < nxi; ++a) phi[size_t(elem) * nxi + a] = phi_pt[a];
        local.unseed_wrt_xi();
        nderivs = global.seed_wrt_x_prev();
        global.interpolate(iota);
        local.evaluate(global);
        EMatrix const dC_dx_prevT = local.eigen_jacobian(nderivs).transpose();
        EVector const fv = matvec(dC_dx_prevT, phi_pt);
        for (int a = 0; a < nx; ++a) f[size_t(elem) * nx + a] = -fv[a];
        global.unseed_wrt_x_prev();
        global.interpolate(iota);
        nderivs = local.seed_wrt_xi_prev();
        local.evaluate(global);
        EMatrix const dC_dxi_prevT = local.eigen_jacobian(nderivs).transpose();
        EVector const gv = matvec(dC_dxi_prevT, phi_pt);
        for (int a = 0; a < nxi; ++a) g[size_t(elem) * nxi + a] = -gv[a];
        local.unseed_wrt_xi_prev();
      }
    }
  }
}

// ---------------------------------------------------------------------------
// eval_qoi, src/evaluations.cpp:662-756
static double eval_qoi(Problem& P, double const* const* x, double const* const* x_prev,
                       double const* xi, double const* xi_prev, int step) {
  Disc const& disc = P.disc;
  auto& local = *P.l_d;
  auto& global = *P.g_d;
  auto& qoi = *P.q_d;
  global.set_time_info(P.time, P.dt);
  preprocess_qoi(P, qoi, local, global, x, x_prev, xi, xi_prev, step);
  global.before_elems(disc);
  qoi.before_elems(disc, step);
  auto const qps = quadrature(disc.dim, 1);
  double J = 0.;
  for (int es = 0; es < disc.n_es; ++es) {
    local.before_elems(es, disc);
    for (int elem : disc.es_elems[es]) {
      global.set_elem(elem);
      qoi.set_elem(elem);
      global.gather(x, x_prev);
      for (auto const& qp : qps) {
        global.interpolate(qp.xi);
        local.gather(elem, xi, xi_prev);
        qoi.evaluate(es, elem, global, local, qp.xi, qp.w, global.geom().dv);
        qoi.scatter(J);
      }
    }
  }
  qoi.postprocess(J);
  return J;
}

// ---------------------------------------------------------------------------
// eval_qoi_gradient, src/evaluations.cpp:758-925 (num_dfad_params == 0)
// grad is [num_active_params_total]; grad_indices[es] maps es-local -> global
static void eval_qoi_gradient(Problem& P, double const* const* x, double const* const* x_prev,
                              double const* xi, double const* xi_prev, double const* const* z,
                              double const* phi, std::vector<std::vector<int>> const& grad_indices,
                              double* grad, int n_grad, int step) {
  Disc const& disc = P.disc;
  auto& local = *P.l_f;
  auto& global = *P.g_f;
  auto& qoi = *P.q_f;
  global.set_time_info(P.time, P.dt);
  preprocess_qoi(P, qoi, local, global, x, x_prev, xi, xi_prev, step);
  global.before_elems(disc);
  qoi.before_elems(disc, step);
  QuadSets Q(disc.dim, global.ip_sets());
  for (int k = 0; k < n_grad; ++k) grad[k] = 0.;
  int nderivs = -1;
  for (int es = 0; es < disc.n_es; ++es) {
    int const np = int(local.active_indices()[es].size());
    EVector es_grad(np, 0.);
    local.before_elems(es, disc);
    int const nxi = local.num_dofs();
    for (int elem : disc.es_elems[es]) {
      global.set_elem(elem);
      qoi.set_elem(elem);
      global.gather(x, x_prev);
      EVector const z_nodes = global.gather_adjoint(z);
      for (size_t ip_set = 0; ip_set < Q.sets.size(); ++ip_set)
        for (auto const& qp : Q.sets[ip_set]) {
          double const* iota = qp.xi;
          double const w = qp.w, dv = global.geom().dv;
          global.interpolate(iota);
          nderivs = local.seed_wrt_params(es);
          if (ip_set == 0) {
            local.gather(elem, xi, xi_prev);
            local.evaluate(global);
            EMatrix const dC_dpT = local.eigen_jacobian(nderivs).transpose();
            EVector phi_pt(nxi);
            for (int a = 0; a < nxi; ++a) phi_pt[a] = phi[size_t(elem) * nxi + a];
            EVector const t = matvec(dC_dpT, phi_pt);
            for (int p = 0; p < np; ++p) es_grad[p] += t[p];
            qoi.evaluate(es, elem, global, local, iota, w, dv);
            EVector const dJ_dp = qoi.eigen_dvector(nderivs);
            for (int p = 0; p < np; ++p) es_grad[p] += dJ_dp[p];
          }
          global.zero_residual();
          global.evaluate(local, iota, w, dv, int(ip_set));
          EMatrix const dR_dpT = global.eigen_jacobian(nderivs).transpose();
          EVector const t2 = matvec(dR_dpT, z_nodes);
          for (int p = 0; p < np; ++p) es_grad[p] += t2[p];
          local.unseed_wrt_params(es);
        }
    }
    // src/local_residual.cpp:858-867 (assignment, not accumulation)
    for (int p = 0; p < np; ++p) grad[grad_indices[es][p]] = es_grad[p];
  }
}

// ---------------------------------------------------------------------------
// eval_measured_residual, src/evaluations.cpp:1750-1845
static int eval_measured_residual(Problem& P, double const* const* x,
                                  double const* const* x_prev, double* xi, double const* xi_prev,
                                  double* const* RHS) {
  Disc const& disc = P.disc;
  auto& local = *P.l_f;
  auto& global = *P.g_f;
  global.set_time_info(P.time, P.dt);
  global.before_elems(disc);
  auto const qps = quadrature(disc.dim, global.ip_sets()[0]);
  int status = 0;
  for (int es = 0; es < disc.n_es; ++es) {
    local.before_elems(es, disc);
    for (int elem : disc.es_elems[es]) {
      global.set_elem(elem);
      global.gather(x, x_prev);
      for (auto const& qp : qps) {
        global.interpolate(qp.xi);
        local.gather(elem, xi, xi_prev);
        local.seed_wrt_xi();
        int path = local.solve_nonlinear(global);
        if (path == -1) status = -1;
        local.scatter(elem, xi);
        local.unseed_wrt_xi();
        global.zero_residual();
        global.evaluate(local, qp.xi, qp.w, global.geom().dv, 0);
        scatter_rhs(P, global, elem, global.eigen_residual(), RHS);
      }
    }
  }
  return status;
}

// ---------------------------------------------------------------------------
// eval_measured_residual_and_grad, src/evaluations.cpp:1847-1973
// local_sens [n_elems][n_xi*n_p] row-major (in/out); dR[i] [n_p][n_rows_i]
static int eval_measured_residual_and_grad(Problem& P, double const* const* x,
                                           double const* const* x_prev, double* xi,
                                           double const* xi_prev, double* const* RHS,
                                           double* const* dR, double* local_sens) {
  Disc const& disc = P.disc;
  auto& local = *P.l_f;
  auto& global = *P.g_f;
  global.set_time_info(P.time, P.dt);
  global.before_elems(disc);
  auto const qps = quadrature(disc.dim, global.ip_sets()[0]);
  int nderivs = -1;
  int status = 0;
  for (int es = 0; es < disc.n_es; ++es) {
    local.before_elems(es, disc);
    int const nxi = local.num_dofs();
    for (int elem : disc.es_elems[es]) {
      global.set_elem(elem);
      global.gather(x, x_prev);
      for (auto const& qp : qps) {
        double const* iota = qp.xi;
        double const w = qp.w, dv = global.geom().dv;
        global.interpolate(iota);
        local.gather(elem, xi, xi_prev);
        nderivs = local.seed_wrt_xi();
        int path = local.solve_nonlinear(global);
        if (path == -1) status = -1;
        local.scatter(elem, xi);
        EMatrix const dC_dxi = local.eigen_jacobian(nderivs);
        global.zero_residual();
        global.evaluate(local, iota, w, dv, 0);
        EVector const elem_resid = global.eigen_residual();
        EMatrix const dR_dxi = global.eigen_jacobian(nderivs);
        local.unseed_wrt_xi();
        nderivs = local.seed_wrt_xi_prev();
        local.evaluate(global);
        EMatrix const dC_dxi_prev = local.eigen_jacobian(nderivs);
        local.unseed_wrt_xi_prev();
        nderivs = local.seed_wrt_params(es);
        int const np = nderivs;
        local.evaluate(global);
        EMatrix const dC_dp = local.eigen_jacobian(nderivs);
        global.zero_residual();
        global.evaluate(local, iota, w, dv, 0);
        EMatrix const dR_dp = global.eigen_jacobian(nderivs);
        local.unseed_wrt_params(es);
        EMatrix prev(nxi, np);
        for (int a = 0; a < nxi; ++a)
          for (int p = 0; p < np; ++p) prev(a, p) = local_sens[(size_t(elem) * nxi + a) * np + p];
        EMatrix rhs = matmul(dC_dxi_prev, prev);
        for (int a = 0; a < nxi; ++a)
          for (int p = 0; p < np; ++p) rhs(a, p) = -dC_dp(a, p) - rhs(a, p);
        EMatrix const dxi_dp = full_piv_lu_solve(dC_dxi, rhs);
        for (int a = 0; a < nxi; ++a)
          for (int p = 0; p < np; ++p) local_sens[(size_t(elem) * nxi + a) * np + p] = dxi_dp(a, p);
        EMatrix tot = matmul(dR_dxi, dxi_dp);
        for (int a = 0; a < tot.r; ++a)
          for (int p = 0; p < np; ++p) tot(a, p) += dR_dp(a, p);
        scatter_rhs(P, global, elem, elem_resid, RHS);
        scatter_sens(P, global, elem, tot, dR);
      }
    }
  }
  return status;
}

// ---------------------------------------------------------------------------
// eval_vfm_adjoint_gradient, src/evaluations.cpp:1975-2143
// hist [n_elems*n_xi] in/out, vf[i] nodal virtual field, grad += over elements
static void eval_vfm_adjoint_gradient(Problem& P, double const* const* x,
                                      double const* const* x_prev, double const* xi,
                                      double const* xi_prev, double const* const* vf, double* hist,
                                      double s, double* grad, int n_grad) {
  Disc const& disc = P.disc;
  auto& local = *P.l_f;
  auto& global = *P.g_f;
  global.set_time_info(P.time, P.dt);
  global.before_elems(disc);
  auto const qps = quadrature(disc.dim, global.ip_sets()[0]);
  for (int k = 0; k < n_grad; ++k) grad[k] = 0.;
  int nderivs = -1;
  for (int es = 0; es < disc.n_es; ++es) {
    local.before_elems(es, disc);
    int const nxi = local.num_dofs();
    for (int elem : disc.es_elems[es]) {
      global.set_elem(elem);
      global.gather(x, x_prev);
      int const nx = global.num_dofs();
      EVector const w_nodes = global.gather_adjoint(vf);
      for (auto const& qp : qps) {
        double const* iota = qp.xi;
        double const w = qp.w, dv = global.geom().dv;
        global.interpolate(iota);
        local.gather(elem, xi, xi_prev);
        nderivs = local.seed_wrt_xi();
        local.evaluate(global);
        EMatrix const dC_dxiT = local.eigen_jacobian(nderivs).transpose();
        global.zero_residual();
        global.evaluate(local, iota, w, dv, 0);
        EMatrix const dR_dxiT = global.eigen_jacobian(nderivs).transpose();
        local.unseed_wrt_xi();
        nderivs = local.seed_wrt_xi_prev();
        local.evaluate(global);
        EMatrix const dC_dxi_prevT = local.eigen_jacobian(nderivs).transpose();
        local.unseed_wrt_xi_prev();
        nderivs = local.seed_wrt_params(es);
        int const np = nderivs;
        local.evaluate(global);
        EMatrix const dC_dp = local.eigen_jacobian(nderivs);
        global.zero_residual();
        global.evaluate(local, iota, w, dv, 0);
        EMatrix const dR_dp = global.eigen_jacobian(nderivs);
        local.unseed_wrt_params(es);
        EVector const t = matvec(dR_dxiT, w_nodes);
        EVector rhs(nxi);
        for (int a = 0; a < nxi; ++a) rhs[a] = s * -t[a] - hist[size_t(elem) * nxi + a];
        EVector const phi = full_piv_lu_solve(dC_dxiT, rhs);
        EVector const hn = matvec(dC_dxi_prevT, phi);
        for (int a = 0; a < nxi; ++a) hist[size_t(elem) * nxi + a] = hn[a];
        for (int p = 0; p < np; ++p) {
          double a1 = 0., a2 = 0.;
          for (int a = 0; a < nx; ++a) a1 += w_nodes[a] * dR_dp(a, p);
          for (int a = 0; a < nxi; ++a) a2 += phi[a] * dC_dp(a, p);
          grad[p] += s * a1 + a2;
        }
      }
    }
  }
}

}  // namespace orc

// ===========================================================================
// C API
// ===========================================================================
using namespace orc;

extern "C" {

void* orc_create(int dim, int n_elems, int n_nodes, int const* conn, double const* coords,
                 int const* elem_es, int n_es) {
  Problem* P = new Problem;
  P->disc.dim = dim;
  P->disc.nn = dim + 1;
  P->disc.n_elems = n_elems;
  P->disc.n_nodes = n_nodes;
  P->disc.n_es = n_es;
  P->disc.conn.assign(conn, conn + size_t(n_elems) * (dim + 1));
  P->disc.coords.assign(coords, coords + size_t(n_nodes) * 3);
  if (elem_es) P->disc.elem_es.assign(elem_es, elem_es + n_elems);
  else P->disc.elem_es.assign(n_elems, 0);
  P->disc.finalize();
  P->ngraph = node_graph(P->disc);
  return P;
}
void orc_destroy(void* p) { delete static_cast<Problem*>(p); }

void orc_set_global(void* p, int type, int mixed, double stab_mult, double thickness) {
  Problem& P = *static_cast<Problem*>(p);
  P.global_type = type; P.mixed = mixed != 0; P.stab_mult = stab_mult; P.thickness = thickness;
  P.make_global();
}

// params: [n_es][n_par] ; active: flattened per-es lists with active_ptr [n_es+1] (may be NULL)
void orc_set_local(void* p, int type, int max_iters, double abs_tol, double rel_tol,
                   double const* params, int const* active_ptr, int const* active_idx) {
  Problem& P = *static_cast<Problem*>(p);
  P.local_type = type;
  P.l_d = create_local_residual<double>(type, P.disc.dim);
  P.l_f = create_local_residual<Fad>(type, P.disc.dim);
  int const np = local_num_params(type);
  std::vector<std::vector<double>> pv(P.disc.n_es, std::vector<double>(np));
  for (int es = 0; es < P.disc.n_es; ++es)
    for (int k = 0; k < np; ++k) pv[es][k] = params[es * np + k];
  std::vector<std::vector<int>> act(P.disc.n_es);
  if (active_ptr)
    for (int es = 0; es < P.disc.n_es; ++es)
      act[es].assign(active_idx + active_ptr[es], active_idx + active_ptr[es + 1]);
  P.l_d->set_param_values(pv); P.l_f->set_param_values(pv);
  P.l_d->set_active_indices(act); P.l_f->set_active_indices(act);
  P.l_d->set_tolerances(max_iters, abs_tol, rel_tol);
  P.l_f->set_tolerances(max_iters, abs_tol, rel_tol);
}

void orc_set_time(void* p, double time, double dt) {
  Problem& P = *static_cast<Problem*>(p);
  P.time = time; P.dt = dt;
}

// out: [num_resid, neq0, neq1, n_xi, n_params, n_x, nn]
void orc_info(void* p, int* out) {
  Problem& P = *static_cast<Problem*>(p);
  out[0] = P.num_resid();
  out[1] = P.neq(0);
  out[2] = P.num_resid() > 1 ? P.neq(1) : 0;
  out[3] = P.l_d ? P.l_d->total_eqs() : 0;
  out[4] = P.l_d ? P.l_d->num_params() : 0;
  int nx = 0;
  for (int i = 0; i < P.num_resid(); ++i) nx += P.neq(i) * P.disc.nn;
  out[5] = nx;
  out[6] = P.disc.nn;
}

int orc_graph_nnz(void* p, int i, int j) {
  Problem& P = *static_cast<Problem*>(p);
  return int(P.bgraph[i][j].colind.size());
}
int orc_graph_rows(void* p, int i, int j) {
  return static_cast<Problem*>(p)->bgraph[i][j].n_rows;
}
void orc_graph(void* p, int i, int j, int* rowptr, int* colind) {
  CsrGraph const& g = static_cast<Problem*>(p)->bgraph[i][j];
  std::memcpy(rowptr, g.rowptr.data(), g.rowptr.size() * sizeof(int));
  std::memcpy(colind, g.colind.data(), g.colind.size() * sizeof(int));
}
int orc_node_graph_nnz(void* p) { return int(static_cast<Problem*>(p)->ngraph.colind.size()); }
void orc_node_graph(void* p, int* rowptr, int* colind) {
  CsrGraph const& g = static_cast<Problem*>(p)->ngraph;
  std::memcpy(rowptr, g.rowptr.data(), g.rowptr.size() * sizeof(int));
  std::memcpy(colind, g.colind.data(), g.colind.size() * sizeof(int));
}
// scatter offsets of block (i,j), [n_elems * (nn*neq_i) * (nn*neq_j)], src/disc.cpp:414-459
void orc_scatter_offsets(void* p, int i, int j, int* out) {
  auto const& v = static_cast<Problem*>(p)->offsets[i][j];
  std::memcpy(out, v.data(), v.size() * sizeof(int));
}

void orc_init_xi(void* p, double* xi) {
  Problem& P = *static_cast<Problem*>(p);
  P.l_f->init_variables(P.disc, xi);
}

int orc_forward_jacobian(void* p, double const* const* x, double const* const* x_prev, double* xi,
                         double const* xi_prev, double* const* LHS, double* const* RHS,
                         double* elem_dtotal, double* elem_R, int* path_out, int* iters_out) {
  Problem& P = *static_cast<Problem*>(p);
  return eval_forward_jacobian(P, x, x_prev, xi, xi_prev, LHS, RHS, elem_dtotal, elem_R,
                               path_out, iters_out, 0, P.disc.n_elems);
}
// element range version (cpu baseline threads over disjoint ranges with private LHS==NULL)
int orc_forward_jacobian_range(void* p, double const* const* x, double const* const* x_prev,
                               double* xi, double const* xi_prev, double* elem_dtotal,
                               double* elem_R, int elem_begin, int elem_end) {
  Problem& P = *static_cast<Problem*>(p);
  return eval_forward_jacobian(P, x, x_prev, xi, xi_prev, nullptr, nullptr, elem_dtotal, elem_R,
                               nullptr, nullptr, elem_begin, elem_end);
}
void orc_debug_newton(void* p, int on) { static_cast<Problem*>(p)->l_f->debug_newton = on != 0; }
int orc_failed_elem(void* p) { return static_cast<Problem*>(p)->n_failed_elem; }

void orc_global_residual(void* p, double const* const* x, double const* const* x_prev,
                         double const* xi, double const* xi_prev, double* const* RHS) {
  eval_global_residual(*static_cast<Problem*>(p), x, x_prev, xi, xi_prev, RHS);
}

// ---- QoI ----
void orc_set_qoi_avg_disp(void* p) {
  Problem& P = *static_cast<Problem*>(p);
  P.qoi_type = 0;
  P.q_d = std::make_unique<AvgDisp<double>>();
  P.q_f = std::make_unique<AvgDisp<Fad>>();
}
void orc_set_qoi_calibration(void* p, double balance_factor, int coord_idx, double coord_value,
                             double coord_tol, int reaction_force_comp, double const* weights,
                             int const* facet /* [n_elems*3] or NULL */) {
  Problem& P = *static_cast<Problem*>(p);
  P.qoi_type = 1;
  P.cal = CalibrationData();
  P.cal.balance_factor = balance_factor;
  P.cal.coord_idx = coord_idx;
  P.cal.coord_value = coord_value;
  P.cal.coord_tol = coord_tol;
  P.cal.reaction_force_comp = reaction_force_comp;
  for (int k = 0; k < 3; ++k) P.cal.weights[k] = weights ? weights[k] : 1.;
  if (facet) P.cal.facet.assign(facet, facet + size_t(P.disc.n_elems) * 3);
  P.q_d = std::make_unique<Calibration<double>>(&P.cal);
  P.q_f = std::make_unique<Calibration<Fad>>(&P.cal);
}
// kind 2 reaction mismatch (plane + component [+ torque]), 3 load mismatch (facet [+ 2-D normal]),
// 4 surface mismatch (facet); facet [n_elems*3] local vertex ids of the side-set facet, -1 none
void orc_set_qoi_mismatch(void* p, int kind, int coord_idx, double coord_value, double coord_tol,
                          int reaction_force_comp, int compute_torque, int const* facet,
                          double const* normal_2d) {
  Problem& P = *static_cast<Problem*>(p);
  P.qoi_type = kind;
  P.cal = CalibrationData();
  P.cal.coord_idx = coord_idx; P.cal.coord_value = coord_value; P.cal.coord_tol = coord_tol;
  P.cal.reaction_force_comp = reaction_force_comp;
  P.cal.compute_torque = compute_torque != 0;
  if (facet) P.cal.facet.assign(facet, facet + size_t(P.disc.n_elems) * 3);
  if (normal_2d) { P.cal.normal_2d[0] = normal_2d[0]; P.cal.normal_2d[1] = normal_2d[1]; }
  if (kind == 2) {
    P.q_d = std::make_unique<ReactionMismatch<double>>(&P.cal);
    P.q_f = std::make_unique<ReactionMismatch<Fad>>(&P.cal);
  } else if (kind == 3) {
    P.q_d = std::make_unique<LoadMismatch<double>>(&P.cal);
    P.q_f = std::make_unique<LoadMismatch<Fad>>(&P.cal);
  } else {
    P.q_d = std::make_unique<SurfaceMismatch<double>>(&P.cal);
    P.q_f = std::make_unique<SurfaceMismatch<Fad>>(&P.cal);
  }
}
void orc_qoi_set_step(void* p, double dt, double total_time, double load_meas,
                      double const* measured) {
  Problem& P = *static_cast<Problem*>(p);
  P.cal.dt = dt; P.cal.total_time = total_time; P.cal.load_meas = load_meas;
  P.cal.measured = measured;
}
// out: [last_total_load, load_mismatch, J_disp, J_forc, area]
void orc_qoi_calibration_state(void* p, double* out) {
  Problem& P = *static_cast<Problem*>(p);
  out[0] = P.cal.last_total_load; out[1] = P.cal.load_mismatch;
  out[2] = P.cal.J_disp; out[3] = P.cal.J_forc; out[4] = P.cal.area;
}
double orc_qoi(void* p, double const* const* x, double const* const* x_prev, double const* xi,
               double const* xi_prev, int step) {
  return eval_qoi(*static_cast<Problem*>(p), x, x_prev, xi, xi_prev, step);
}

void orc_adjoint_jacobian(void* p, double const* const* x, double const* const* x_prev,
                          double const* xi, double const* xi_prev, double* g, double const* f,
                          double* const* LHS, double* const* RHS, int step) {
  eval_adjoint_jacobian(*static_cast<Problem*>(p), x, x_prev, xi, xi_prev, g, f, LHS, RHS, step);
}
void orc_adjoint_local(void* p, double const* const* x, double const* const* x_prev,
                       double const* xi, double const* xi_prev, double const* const* z,
                       double* phi, double* g, double* f) {
  solve_adjoint_local(*static_cast<Problem*>(p), x, x_prev, xi, xi_prev, z, phi, g, f);
}
// grad_ptr/grad_idx: per-es mapping from es-local active param to the global gradient slot
void orc_qoi_gradient(void* p, double const* const* x, double const* const* x_prev,
                      double const* xi, double const* xi_prev, double const* const* z,
                      double const* phi, int const* grad_ptr, int const* grad_idx, double* grad,
                      int n_grad, int step) {
  Problem& P = *static_cast<Problem*>(p);
  std::vector<std::vector<int>> gi(P.disc.n_es);
  for (int es = 0; es < P.disc.n_es; ++es)
    gi[es].assign(grad_idx + grad_ptr[es], grad_idx + grad_ptr[es + 1]);
  eval_qoi_gradient(P, x, x_prev, xi, xi_prev, z, phi, gi, grad, n_grad, step);
}

int orc_measured_residual(void* p, double const* const* x, double const* const* x_prev,
                          double* xi, double const* xi_prev, double* const* RHS) {
  return eval_measured_residual(*static_cast<Problem*>(p), x, x_prev, xi, xi_prev, RHS);
}
int orc_measured_residual_grad(void* p, double const* const* x, double const* const* x_prev,
                               double* xi, double const* xi_prev, double* const* RHS,
                               double* const* dR, double* local_sens) {
  return eval_measured_residual_and_grad(*static_cast<Problem*>(p), x, x_prev, xi, xi_prev, RHS,
                                         dR, local_sens);
}
void orc_vfm_adjoint_gradient(void* p, double const* const* x, double const* const* x_prev,
                              double const* xi, double const* xi_prev, double const* const* vf,
                              double* hist, double s, double* grad, int n_grad) {
  eval_vfm_adjoint_gradient(*static_cast<Problem*>(p), x, x_prev, xi, xi_prev, vf, hist, s, grad,
                            n_grad);
}

long long orc_flops_reset() {
#ifdef C8_ORACLE_COUNT
  long long v = g_flops; g_flops = 0; return v;
#else
  return -1;
#endif
}

// quadrature KAT hook (test/unit/quadrature.cpp.in:47-65): returns npts, fills xi[npts*3], w[npts]
int orc_quadrature(int dim, int order, double* xi, double* w) {
  auto q = quadrature(dim, order);
  for (size_t k = 0; k < q.size(); ++k) {
    for (int c = 0; c < 3; ++c) xi[k * 3 + c] = q[k].xi[c];
    w[k] = q[k].w;
  }
  return int(q.size());
}

}  // extern "C"
