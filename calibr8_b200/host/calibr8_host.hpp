// Host-side mirror of calibr8's step solvers and objective for the hot path, in C++, ABOVE the
// C ABI of include/c8b200.h (every device operation goes through a c8_* entry point).
//
//   Problem   <- State / Disc time + BC + QoI settings   src/state.cpp:33-46, src/disc.cpp:136-155
//   Primal    <- Primal::solve_at_step                    src/primal.cpp:31-208 (+ line_search.hpp)
//   Adjoint   <- Adjoint::solve_at_step + Adjoint_Objective::gradient
//                                                         src/adjoint.cpp:76-189, adjoint_objective.cpp:48-118
// The all-steps primal history (x[step], xi[step]) stays resident on the device for the reverse
// sweep, like the apf fields of Disc::primal(step) in the reference (src/disc.cpp:643-683).
#pragma once
#include <chrono>
#include <cmath>
#include <cstdio>
#include <stdexcept>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "c8b200.h"
#include "expr.hpp"
#include "residuals.hpp"

namespace c8host {

struct Dbc {  // "bc name: [resid_idx, eq, node_set_name, value]", src/dbcs.cpp:56-66
  int resid, eq;
  std::vector<int> nodes;
  Expr expr;
};

struct Tbc {  // "bc name: [resid_idx, side_set_name, x-val, y-val(, z-val)]", src/tbcs.cpp:28-35
  int resid;
  std::vector<int> side_nodes;     // [n_sides][dim] local node ids
  std::vector<Expr> exprs;         // one per dimension
  int n_sides = 0;
  int* d_side_nodes = nullptr;
  double* d_traction = nullptr;    // [n_sides][dim] at the side quadrature points, current step
};

struct SolverParams {
  int newton_max_iters = 15;           // "nonlinear max iters"
  double newton_abs_tol = 1e-8;        // "nonlinear absolute tol"
  double newton_rel_tol = 1e-8;        // "nonlinear relative tol"
  int gmres_restart = 100;
  int gmres_max_iters = 4000;
  double linear_tol = 1e-10;           // Belos "Convergence Tolerance" (relative)
  // line search, src/line_search.hpp:24-31
  double ls_c1 = 1e-4, ls_bmin = 0.5, ls_bmax = 0.9;
  int ls_max_evals = 4;
  bool reuse_accepted_assembly = true;  // do not re-assemble the state the line search just accepted
  bool print = false;
};

struct CalibrationQoi {
  bool enabled = false;
  double balance_factor = 1., coord_value = 0., coord_tol = 1e-12;
  int coord_idx = -1, reaction_force_comp = -1;
  double weights[3] = {1., 1., 1.};
  std::vector<double> load_data;           // measured load per step ("load input file")
  std::vector<double*> d_measured;         // per step (1..N) device [n_nodes][dim]
  signed char* d_facet = nullptr;          // side-set facets [n_elems][3] local vertex ids
  double area = 1.;
  bool compute_torque = false;             // reaction mismatch "compute torque"
  double normal_2d[2] = {0., 0.};          // load mismatch "2D surface normal"
  std::vector<double> total_load;          // per step: what "load out file" holds (load.dat)
};

class DevVec {  // RAII device array
 public:
  DevVec() {}
  explicit DevVec(size_t n) { resize(n); }
  ~DevVec() { if (p_) cudaFree(p_); }
  DevVec(const DevVec&) = delete;
  DevVec& operator=(const DevVec&) = delete;
  DevVec(DevVec&& o) noexcept : p_(o.p_), n_(o.n_) { o.p_ = nullptr; o.n_ = 0; }
  void resize(size_t n) {
    if (p_) cudaFree(p_);
    p_ = nullptr; n_ = n;
    if (n && cudaMalloc(&p_, n * sizeof(double)) != cudaSuccess) throw std::runtime_error("cudaMalloc");
    if (n) {  // null-stream memset is not ordered with the context's non-blocking stream
      cudaMemset(p_, 0, n * sizeof(double));
      cudaDeviceSynchronize();
    }
  }
  double* get() const { return p_; }
  size_t size() const { return n_; }
 private:
  double* p_ = nullptr;
  size_t n_ = 0;
};

class Problem {
 public:
  explicit Problem(c8_ctx* ctx);
  ~Problem();
  void check(int rc, const char* what) const;

  c8_ctx* ctx;
  int dim, nn, nb, nx, nxi, npar, n_elems, n_nodes, nnzb;
  int n_owned_nodes = 0, n_owned_elems = 0;  // partition (== n_nodes / n_elems on one GPU)
  long long n_dofs, xi_ld;
  std::vector<double> coords;  // [n_nodes][3] host copy for BC expressions
  int num_steps = 0;
  double step_size = 1.;
  std::vector<Dbc> dbcs;
  std::vector<Tbc> tbcs;
  SolverParams sp;
  int qoi_type = 0;  // c8_qoi.type: 0 average displacement, 1 calibration, 2 reaction / 3 load / 4 surface mismatch
  CalibrationQoi cal;

  // history (device)
  std::vector<DevVec> x, xi;   // [0..num_steps]
  std::vector<DevVec> z, phi;  // adjoint fields per step
  DevVec A, b, dx, Adx, saved_xi, work;
  DevVec adj_g, adj_f, adj_rhs, adj_grad;  // reverse-sweep work arrays (reused)
  // flattened Dirichlet dofs
  int n_dbc = 0;
  int* d_dbc_node = nullptr;
  int* d_dbc_eq = nullptr;
  double* d_dbc_val = nullptr;
  std::vector<double> h_dbc_val;
  int n_assemblies = 0, n_linear_iters = 0;
  // optional wall-clock profile (stream-synchronised around each phase when enabled):
  // 0 forward assembly (K1 + Dirichlet rows + norm), 1 linear solves, 2 K3 adjoint Jacobian,
  // 3 K4 adjoint local, 4 K5/K6 objective + gradient integrands, 5 line-search SpMV/dots
  bool profile = false;
  double t_phase[8] = {0, 0, 0, 0, 0, 0, 0, 0};

  double time(int step) const { return step * step_size; }
  void set_time(int n_steps, double dt);
  void add_dbc(int resid, int eq, const int* nodes, int n, const std::string& expr);
  void finalize_dbcs();
  void eval_dbc_values(double t);
  void add_tbc(int resid, const int* side_nodes, int n_sides, const std::vector<std::string>& exprs);
  void eval_tbc_values(double t);    // traction vectors at the side quadrature points for time t
  void apply_tbcs(double* R) const;  // apply_primal_tbcs, src/tbcs.cpp:88-98
  void allocate_history();
  double norm(const double* v) const;
  double dot(const double* a, const double* c) const;
  void axpy(double a, const double* xsrc, double* y) const;  // y += a x
};

struct PhaseTimer {  // adds the wall time of a scope to Problem::t_phase[id] when profiling
  PhaseTimer(Problem& p, int id);
  ~PhaseTimer();
  Problem& P;
  int id;
  double t0 = 0.;
};

class Primal {
 public:
  explicit Primal(Problem& p) : P(p) {}
  void solve_at_step(int step);
  double eval_qoi(int step);   // eval_qoi, src/evaluations.cpp:662-756
  double solve_all();          // Solver::solve, src/main_primal.cpp:221-243
 private:
  bool assemble(int step, double* R_norm);  // K1 + DBC; false = local solve failed
  Problem& P;
};

class Adjoint {
 public:
  explicit Adjoint(Problem& p) : P(p) {}
  // Adjoint_Objective::gradient's reverse sweep: d J / d p for EVERY model parameter of
  // element set 0..n_es-1 (the caller selects the active ones); grad [n_es][npar]
  void gradient(std::vector<double>& grad);
 private:
  Problem& P;
};

// Objective on canonical [-1, 1] parameters (the ROL::Objective<double> role of the reference,
// src/objective.hpp:15-60): value(p) / gradient(g, p), parameter scaling, re-use of the forward
// solve when the parameters did not change (param_diff).
class Objective {
 public:
  Objective(Problem& p, const std::vector<int>& active, const std::vector<double>& lower,
            const std::vector<double>& upper);
  virtual ~Objective() {}
  virtual double value(const std::vector<double>& p_canonical) = 0;
  virtual void gradient(std::vector<double>& g, const std::vector<double>& p_canonical) = 0;
  std::vector<double> transform_params(const std::vector<double>& p, bool scale_to_canonical) const;
  std::vector<double> transform_gradient(const std::vector<double>& g) const;
  std::vector<double> active_params() const;   // physical values of the active parameters
  int num_opt_params() const { return int(m_active.size()); }
 protected:
  bool param_diff(const std::vector<double>& p) const;
  void set_params_from_canonical(const std::vector<double>& p);
  Problem& P;
  std::vector<int> m_active;          // parameter indices (element set 0)
  std::vector<double> m_lower, m_upper, m_base, m_p_old;
  double m_J_old = 0.;
  const double m_difftol = 1.0e-15;
};

class AdjointObjective : public Objective {  // src/adjoint_objective.cpp:24-118
 public:
  using Objective::Objective;
  double value(const std::vector<double>& p) override;
  void gradient(std::vector<double>& g, const std::vector<double>& p) override;
};

}  // namespace c8host
