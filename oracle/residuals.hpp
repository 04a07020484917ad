// ORACLE (test infrastructure only).
//
// CPU restatement of calibr8's GlobalResidual<T> / LocalResidual<T> operator
// API for the hot path:
//   GlobalResidual   src/global_residual.{hpp,cpp}
//   Mechanics        src/mechanics.cpp:16-240
//   MechanicsPlaneStress  src/mechanics_plane_stress.cpp:16-95
//   LocalResidual    src/local_residual.{hpp,cpp}
//   Elastic / SmallJ2 / SmallHill / SmallHillPlaneStress / SmallHillPlaneStrain
//   HyperJ2 / HyperJ2PlaneStress / HyperJ2PlaneStrain   (files cited per class)
// T is double or orc::Fad (the SLFad<double,16> stand-in).
#pragma once
#include <memory>
#include <string>
#include <stdexcept>
#include <cstdio>
#include "disc.hpp"
#include "tensor.hpp"

namespace orc {

template <class T> class LocalResidual;

template <class T> inline void zero_derivs(T&) {}
template <> inline void zero_derivs<Fad>(Fad& x) {
  for (int k = 0; k < nmax_derivs; ++k) x.d[k] = 0.;
}
template <class T> inline void seed(T&, int, int) {}
template <> inline void seed<Fad>(Fad& x, int i, int n) { x.diff(i, n); }
template <class T> inline void unseed(T&) {}
template <> inline void unseed<Fad>(Fad& x) { double v = x.v; x = Fad(v); }

// ---------------------------------------------------------------------------
// GlobalResidual  (src/global_residual.cpp)
// ---------------------------------------------------------------------------
template <class T>
class GlobalResidual {
 public:
  virtual ~GlobalResidual() {}

  int num_residuals() const { return m_num_residuals; }
  int num_eqs(int i) const { return m_num_eqs[i]; }
  int var_type(int i) const { return m_var_types[i]; }
  int num_dims() const { return m_num_dims; }
  int num_nodes() const { return m_num_nodes; }
  int num_dofs() const { return m_num_dofs; }
  std::vector<int> const& ip_sets() const { return m_ip_sets; }

  // src/global_residual.cpp:21-23
  int dx_idx(int i, int node, int eq) const {
    return m_dx_offsets[i] + (node * m_num_eqs[i] + eq);
  }

  // src/global_residual.cpp:101-146
  void before_elems(Disc const& disc) {
    m_disc = &disc;
    m_num_dims = disc.dim;
    m_num_nodes = disc.nn;
    m_x_nodal.assign(m_num_residuals, {});
    m_R_nodal.assign(m_num_residuals, {});
    m_x_prev_nodal.assign(m_num_residuals, {});
    m_x.assign(m_num_residuals, {});
    m_x_prev.assign(m_num_residuals, {});
    m_grad_x.assign(m_num_residuals, {});
    m_grad_x_prev.assign(m_num_residuals, {});
    m_num_dofs = 0;
    m_dx_offsets.assign(m_num_residuals, 0);
    for (int i = 0; i < m_num_residuals; ++i) {
      int const neq = m_num_eqs[i];
      m_x_nodal[i].assign(m_num_nodes, std::vector<T>(neq, T(0.)));
      m_R_nodal[i].assign(m_num_nodes, std::vector<T>(neq, T(0.)));
      m_x_prev_nodal[i].assign(m_num_nodes, std::vector<T>(neq, T(0.)));
      m_x[i].assign(neq, T(0.));
      m_x_prev[i].assign(neq, T(0.));
      m_grad_x[i].assign(neq, std::vector<T>(m_num_dims, T(0.)));
      m_grad_x_prev[i].assign(neq, std::vector<T>(m_num_dims, T(0.)));
      m_dx_offsets[i] = m_num_dofs;
      m_num_dofs += neq * m_num_nodes;
    }
  }

  void set_elem(int elem) {
    m_elem = elem;
    m_geom.set(*m_disc, elem);
  }
  ElemGeom const& geom() const { return m_geom; }
  int elem() const { return m_elem; }

  // src/global_residual.cpp:153-178
  void zero_residual() {
    for (int i = 0; i < m_num_residuals; ++i)
      for (int n = 0; n < m_num_nodes; ++n)
        for (int eq = 0; eq < m_num_eqs[i]; ++eq) m_R_nodal[i][n][eq] = T(0.);
  }

  // src/global_residual.cpp:181-198 ; fields are [n_nodes * neq_i] per residual
  void gather(double const* const* x, double const* const* x_prev) {
    for (int i = 0; i < m_num_residuals; ++i) {
      int const neq = m_num_eqs[i];
      for (int n = 0; n < m_num_nodes; ++n) {
        int const node = m_disc->conn[size_t(m_elem) * m_num_nodes + n];
        for (int eq = 0; eq < neq; ++eq) {
          m_x_nodal[i][n][eq] = T(x[i][size_t(node) * neq + eq]);
          m_x_prev_nodal[i][n][eq] = T(x_prev[i][size_t(node) * neq + eq]);
        }
      }
    }
  }

  // src/global_residual.cpp:206-216
  int seed_wrt_x() {
    for (int i = 0; i < m_num_residuals; ++i)
      for (int n = 0; n < m_num_nodes; ++n)
        for (int eq = 0; eq < m_num_eqs[i]; ++eq)
          seed(m_x_nodal[i][n][eq], dx_idx(i, n, eq), m_num_dofs);
    return m_num_dofs;
  }
  // src/global_residual.cpp:226-238
  void unseed_wrt_x() {
    for (int i = 0; i < m_num_residuals; ++i)
      for (int n = 0; n < m_num_nodes; ++n)
        for (int eq = 0; eq < m_num_eqs[i]; ++eq) {
          unseed(m_x_nodal[i][n][eq]);
          zero_derivs(m_R_nodal[i][n][eq]);
        }
  }
  // src/global_residual.cpp:249-259
  int seed_wrt_x_prev() {
    for (int i = 0; i < m_num_residuals; ++i)
      for (int n = 0; n < m_num_nodes; ++n)
        for (int eq = 0; eq < m_num_eqs[i]; ++eq)
          seed(m_x_prev_nodal[i][n][eq], dx_idx(i, n, eq), m_num_dofs);
    return m_num_dofs;
  }
  void unseed_wrt_x_prev() {
    for (int i = 0; i < m_num_residuals; ++i)
      for (int n = 0; n < m_num_nodes; ++n)
        for (int eq = 0; eq < m_num_eqs[i]; ++eq) {
          unseed(m_x_prev_nodal[i][n][eq]);
          zero_derivs(m_R_nodal[i][n][eq]);
        }
  }

  // src/global_residual.cpp:288-332
  void interpolate(double const* iota) {
    m_geom.basis(iota, m_basis);
    for (int i = 0; i < m_num_residuals; ++i)
      for (int eq = 0; eq < m_num_eqs[i]; ++eq) {
        m_x[i][eq] = m_x_nodal[i][0][eq] * m_basis[0];
        m_x_prev[i][eq] = m_x_prev_nodal[i][0][eq] * m_basis[0];
        for (int n = 1; n < m_num_nodes; ++n) {
          m_x[i][eq] += m_x_nodal[i][n][eq] * m_basis[n];
          m_x_prev[i][eq] += m_x_prev_nodal[i][n][eq] * m_basis[n];
        }
      }
    for (int i = 0; i < m_num_residuals; ++i)
      for (int eq = 0; eq < m_num_eqs[i]; ++eq)
        for (int d = 0; d < m_num_dims; ++d) {
          m_grad_x[i][eq][d] = m_x_nodal[i][0][eq] * m_geom.grad[0][d];
          m_grad_x_prev[i][eq][d] = m_x_prev_nodal[i][0][eq] * m_geom.grad[0][d];
          for (int n = 1; n < m_num_nodes; ++n) {
            m_grad_x[i][eq][d] += m_x_nodal[i][n][eq] * m_geom.grad[n][d];
            m_grad_x_prev[i][eq][d] += m_x_prev_nodal[i][n][eq] * m_geom.grad[n][d];
          }
        }
    compute_kinematics();
    m_R_valid = false;   // src/global_residual.cpp:331
  }

  virtual void compute_kinematics() {}

  // accessors, src/global_residual.cpp:25-99
  T scalar_x(int i) const { return m_x[i][0]; }
  Vec<T> vector_x(int i) const {
    Vec<T> v(m_num_dims);
    for (int d = 0; d < m_num_dims; ++d) v(d) = m_x[i][d];
    return v;
  }
  Vec<T> grad_scalar_x(int i) const {
    Vec<T> v(m_num_dims);
    for (int d = 0; d < m_num_dims; ++d) v(d) = m_grad_x[i][0][d];
    return v;
  }
  Tensor<T> grad_vector_x(int i) const {
    Tensor<T> v(m_num_dims);
    for (int k = 0; k < m_num_dims; ++k)
      for (int l = 0; l < m_num_dims; ++l) v(k, l) = m_grad_x[i][k][l];
    return v;
  }
  Tensor<T> grad_vector_x_prev(int i) const {
    Tensor<T> v(m_num_dims);
    for (int k = 0; k < m_num_dims; ++k)
      for (int l = 0; l < m_num_dims; ++l) v(k, l) = m_grad_x_prev[i][k][l];
    return v;
  }
  T& R_nodal(int i, int n, int eq) { return m_R_nodal[i][n][eq]; }
  T const& x_nodal(int i, int n, int eq) const { return m_x_nodal[i][n][eq]; }
  double weight(int, int n, int) const { return m_basis[n]; }
  double grad_weight(int, int n, int, int d) const { return m_geom.grad[n][d]; }
  Tensor<T> const& cof_F() const { return m_cof_F; }
  T const& det_F() const { return m_det_F; }
  Tensor<T> const& F() const { return m_F; }
  Tensor<T> const& F_prev() const { return m_F_prev; }
  // src/global_residual.hpp:302-305: cached until the next interpolate (every re-seeding interpolates again)
  Tensor<T> const& R() const {
    if (!m_R_valid) { m_R = polar_rotation(m_F); m_R_valid = true; }
    return m_R;
  }

  // src/global_residual.cpp:373-414
  EVector eigen_residual() const {
    EVector R(m_num_dofs);
    for (int i = 0; i < m_num_residuals; ++i)
      for (int n = 0; n < m_num_nodes; ++n)
        for (int eq = 0; eq < m_num_eqs[i]; ++eq)
          R[dx_idx(i, n, eq)] = val(m_R_nodal[i][n][eq]);
    return R;
  }
  EMatrix eigen_jacobian(int nderivs) const {
    EMatrix J(m_num_dofs, nderivs);
    for (int i = 0; i < m_num_residuals; ++i)
      for (int n = 0; n < m_num_nodes; ++n)
        for (int eq = 0; eq < m_num_eqs[i]; ++eq)
          for (int j = 0; j < nderivs; ++j)
            J(dx_idx(i, n, eq), j) = dx(m_R_nodal[i][n][eq], j);
    return J;
  }
  // src/global_residual.cpp:422-438
  EVector gather_adjoint(double const* const* z) const {
    EVector zn(m_num_dofs);
    for (int i = 0; i < m_num_residuals; ++i)
      for (int n = 0; n < m_num_nodes; ++n) {
        int const node = m_disc->conn[size_t(m_elem) * m_num_nodes + n];
        for (int eq = 0; eq < m_num_eqs[i]; ++eq)
          zn[dx_idx(i, n, eq)] = z[i][size_t(node) * m_num_eqs[i] + eq];
      }
    return zn;
  }

  virtual void evaluate(LocalResidual<T>& local, double const* iota, double w,
                        double dv, int ip_set) = 0;

  void set_time_info(double t, double dt) { m_time = t; m_delta_t = dt; }

 protected:
  Disc const* m_disc = nullptr;
  int m_elem = -1;
  ElemGeom m_geom;
  double m_basis[4];
  int m_num_residuals = 0, m_num_dims = 0, m_num_nodes = 0, m_num_dofs = 0;
  std::vector<int> m_num_eqs, m_var_types, m_ip_sets, m_dx_offsets;
  std::vector<std::vector<std::vector<T>>> m_x_nodal, m_R_nodal, m_x_prev_nodal;
  std::vector<std::vector<T>> m_x, m_x_prev;
  std::vector<std::vector<std::vector<T>>> m_grad_x, m_grad_x_prev;
  Tensor<T> m_F, m_F_prev, m_cof_F;
  mutable Tensor<T> m_R;
  mutable bool m_R_valid = false;
  T m_det_F;
  double m_time = 0., m_delta_t = 0.;
};

// ---------------------------------------------------------------------------
// LocalResidual  (src/local_residual.cpp)
// ---------------------------------------------------------------------------
template <class T>
class LocalResidual {
 public:
  virtual ~LocalResidual() {}

  int num_residuals() const { return m_num_residuals; }
  int num_eqs(int i) const { return m_num_eqs[i]; }
  int var_type(int i) const { return m_var_types[i]; }
  int num_dofs() const { return m_num_dofs; }
  int num_params() const { return int(m_params.size()); }
  int z_stretch_idx() const { return m_z_stretch_idx; }
  T const& params(int p) const { return m_params[p]; }

  // total packed local dofs (usable before before_elems)
  int total_eqs() const {
    int s = 0;
    for (int i = 0; i < m_num_residuals; ++i) s += m_num_eqs[i];
    return s;
  }

  void set_param_values(std::vector<std::vector<double>> const& pv) {
    m_param_values = pv;
    m_params.assign(pv.empty() ? 0 : pv[0].size(), T(0.));
  }
  void set_active_indices(std::vector<std::vector<int>> const& a) { m_active_indices = a; }
  std::vector<std::vector<int>> const& active_indices() const { return m_active_indices; }
  void set_tolerances(int max_iters, double abs_tol, double rel_tol) {
    m_max_iters = max_iters; m_abs_tol = abs_tol; m_rel_tol = rel_tol;
  }

  // src/local_residual.cpp:76-102
  void before_elems(int es, Disc const& disc) {
    m_num_dims = disc.dim;
    m_xi.assign(m_num_residuals, {});
    m_xi_prev.assign(m_num_residuals, {});
    m_R.assign(m_num_residuals, {});
    m_num_dofs = 0;
    m_dxi_offsets.assign(m_num_residuals, 0);
    for (int i = 0; i < m_num_residuals; ++i) {
      m_xi[i].assign(m_num_eqs[i], T(0.));
      m_xi_prev[i].assign(m_num_eqs[i], T(0.));
      m_R[i].assign(m_num_eqs[i], T(0.));
      m_dxi_offsets[i] = m_num_dofs;
      m_num_dofs += m_num_eqs[i];
    }
    for (size_t p = 0; p < m_params.size(); ++p) m_params[p] = T(m_param_values[es][p]);
  }
  int dxi_idx(int i, int eq) const { return m_dxi_offsets[i] + eq; }

  // src/local_residual.cpp:108-119
  double norm_residual() const {
    double norm = 0.;
    for (int i = 0; i < m_num_residuals; ++i)
      for (int eq = 0; eq < m_num_eqs[i]; ++eq) {
        double const v = val(m_R[i][eq]);
        norm += v * v;
      }
    return std::sqrt(norm);
  }
  // src/local_residual.cpp:128-141
  EMatrix eigen_jacobian(int nderivs) const {
    EMatrix J(m_num_dofs, nderivs);
    for (int i = 0; i < m_num_residuals; ++i)
      for (int eq = 0; eq < m_num_eqs[i]; ++eq)
        for (int j = 0; j < nderivs; ++j) J(dxi_idx(i, eq), j) = dx(m_R[i][eq], j);
    return J;
  }
  EVector eigen_residual() const {
    EVector R(m_num_dofs);
    for (int i = 0; i < m_num_residuals; ++i)
      for (int eq = 0; eq < m_num_eqs[i]; ++eq) R[dxi_idx(i, eq)] = val(m_R[i][eq]);
    return R;
  }

  // packed accessors, src/local_residual.cpp:176-274
  T scalar_xi(int i) const { return m_xi[i][0]; }
  T scalar_xi_prev(int i) const { return m_xi_prev[i][0]; }
  Tensor<T> sym_tensor_xi(int i) const { return unpack_sym(m_xi[i]); }
  Tensor<T> sym_tensor_xi_prev(int i) const { return unpack_sym(m_xi_prev[i]); }

  // value-only setters, src/local_residual.cpp:276-360
  void set_scalar_xi(int i, T const& xi) { set_val(m_xi[i][0], val(xi)); }
  void set_sym_tensor_xi(int i, Tensor<T> const& xi) {
    if (m_num_dims == 2) {
      set_val(m_xi[i][0], val(xi(0, 0))); set_val(m_xi[i][1], val(xi(0, 1)));
      set_val(m_xi[i][2], val(xi(1, 1)));
    } else {
      set_val(m_xi[i][0], val(xi(0, 0))); set_val(m_xi[i][1], val(xi(0, 1)));
      set_val(m_xi[i][2], val(xi(0, 2))); set_val(m_xi[i][3], val(xi(1, 1)));
      set_val(m_xi[i][4], val(xi(1, 2))); set_val(m_xi[i][5], val(xi(2, 2)));
    }
  }
  // src/local_residual.cpp:414-540 (Newton update adds to values only)
  void add_to_xi(EVector const& dxi) {
    for (int i = 0; i < m_num_residuals; ++i)
      for (int eq = 0; eq < m_num_eqs[i]; ++eq)
        set_val(m_xi[i][eq], val(m_xi[i][eq]) + dxi[dxi_idx(i, eq)]);
  }

  void set_scalar_R(int i, T const& R) { m_R[i][0] = R; }
  // src/local_residual.cpp:564-579
  void set_sym_tensor_R(int i, Tensor<T> const& R) {
    if (m_num_dims == 2) {
      m_R[i][0] = R(0, 0); m_R[i][1] = R(0, 1); m_R[i][2] = R(1, 1);
    } else {
      m_R[i][0] = R(0, 0); m_R[i][1] = R(0, 1); m_R[i][2] = R(0, 2);
      m_R[i][3] = R(1, 1); m_R[i][4] = R(1, 2); m_R[i][5] = R(2, 2);
    }
  }

  // src/local_residual.cpp:598-617 ; xi fields are packed [n_elems * n_xi]
  void gather(int elem, double const* xi, double const* xi_prev) {
    for (int i = 0; i < m_num_residuals; ++i)
      for (int eq = 0; eq < m_num_eqs[i]; ++eq) {
        m_R[i][eq] = T(0.);
        m_xi[i][eq] = T(xi[size_t(elem) * m_num_dofs + dxi_idx(i, eq)]);
        m_xi_prev[i][eq] = T(xi_prev[size_t(elem) * m_num_dofs + dxi_idx(i, eq)]);
      }
  }
  // src/local_residual.cpp:623-631
  void scatter(int elem, double* xi) const {
    for (int i = 0; i < m_num_residuals; ++i)
      for (int eq = 0; eq < m_num_eqs[i]; ++eq)
        xi[size_t(elem) * m_num_dofs + dxi_idx(i, eq)] = val(m_xi[i][eq]);
  }

  // seeding, src/local_residual.cpp:702-819
  int seed_wrt_xi() {
    for (int i = 0; i < m_num_residuals; ++i)
      for (int eq = 0; eq < m_num_eqs[i]; ++eq) seed(m_xi[i][eq], dxi_idx(i, eq), m_num_dofs);
    return m_num_dofs;
  }
  void unseed_wrt_xi() {
    for (int i = 0; i < m_num_residuals; ++i)
      for (int eq = 0; eq < m_num_eqs[i]; ++eq) {
        unseed(m_xi[i][eq]);
        zero_derivs(m_R[i][eq]);
      }
  }
  int seed_wrt_xi_prev() {
    for (int i = 0; i < m_num_residuals; ++i)
      for (int eq = 0; eq < m_num_eqs[i]; ++eq)
        seed(m_xi_prev[i][eq], dxi_idx(i, eq), m_num_dofs);
    return m_num_dofs;
  }
  void unseed_wrt_xi_prev() {
    for (int i = 0; i < m_num_residuals; ++i)
      for (int eq = 0; eq < m_num_eqs[i]; ++eq) {
        unseed(m_xi_prev[i][eq]);
        zero_derivs(m_R[i][eq]);
      }
  }
  // src/local_residual.cpp:785-800
  void seed_wrt_x(EMatrix const& dxi_dx) { seed_chain(dxi_dx); }
  // src/local_residual.cpp:811-819
  int seed_wrt_params(int es) {
    int const np = int(m_active_indices[es].size());
    for (int p = 0; p < np; ++p) seed(m_params[m_active_indices[es][p]], p, np);
    return np;
  }
  void unseed_wrt_params(int es) {
    int const np = int(m_active_indices[es].size());
    for (int p = 0; p < np; ++p) unseed(m_params[m_active_indices[es][p]]);
    for (int i = 0; i < m_num_residuals; ++i)
      for (int eq = 0; eq < m_num_eqs[i]; ++eq) zero_derivs(m_R[i][eq]);
  }

  // initial conditions, src/local_residual.cpp:34-74
  virtual void init_variables_impl() = 0;
  void init_variables(Disc const& disc, double* xi) {
    before_elems(0, disc);
    for (int e = 0; e < disc.n_elems; ++e) {
      init_variables_impl();
      scatter(e, xi);
    }
  }

  // generic local Newton shared by all plastic models
  // (e.g. src/small_J2.cpp:136-172); returns path or -1
  int newton(GlobalResidual<T>& global) {
    int path = 0;
    int iter = 1;
    double R_norm_0 = 1.;
    bool converged = false;
    while ((iter <= m_max_iters) && (!converged)) {
      path = this->evaluate(global);
      double const R_norm = this->norm_residual();
      if (iter == 1) R_norm_0 = R_norm;
      double const R_norm_rel = R_norm / R_norm_0;
      if (debug_newton) std::printf("   local it %d path %d |R| %.6e rel %.3e\n", iter, path, R_norm, R_norm_rel);
      if ((R_norm_rel < m_rel_tol) || (R_norm < m_abs_tol)) {
        converged = true;
        break;
      }
      EMatrix const J = this->eigen_jacobian(this->m_num_dofs);
      EVector R = this->eigen_residual();
      for (auto& r : R) r = -r;
      EVector const dxi = full_piv_lu_solve(J, R);
      this->add_to_xi(dxi);
      iter++;
    }
    last_iters = iter;
    if ((iter > m_max_iters) && (!converged)) return -1;
    return path;
  }
  int last_iters = 0;
  bool debug_newton = false;

  virtual int solve_nonlinear(GlobalResidual<T>& global) = 0;
  virtual int evaluate(GlobalResidual<T>& global, bool force_path = false, int path = 0) = 0;
  virtual bool is_finite_deformation() = 0;
  virtual Tensor<T> cauchy(GlobalResidual<T>& global) = 0;
  virtual Tensor<T> dev_cauchy(GlobalResidual<T>& global) = 0;
  virtual T hydro_cauchy(GlobalResidual<T>& global) = 0;
  virtual T pressure_scale_factor() = 0;

 protected:
  static void set_val(double& x, double v) { x = v; }
  static void set_val(Fad& x, double v) { x.v = v; }
  void seed_chain(EMatrix const& dxi_dx);
  Tensor<T> unpack_sym(std::vector<T> const& p) const {
    Tensor<T> t(m_num_dims);
    if (m_num_dims == 2) {
      t(0, 0) = p[0]; t(0, 1) = p[1]; t(1, 0) = p[1]; t(1, 1) = p[2];
    } else {
      t(0, 0) = p[0]; t(0, 1) = p[1]; t(0, 2) = p[2];
      t(1, 0) = p[1]; t(1, 1) = p[3]; t(1, 2) = p[4];
      t(2, 0) = p[2]; t(2, 1) = p[4]; t(2, 2) = p[5];
    }
    return t;
  }

  int m_num_residuals = 0, m_num_dims = 0, m_num_dofs = 0;
  int m_z_stretch_idx = -1;
  std::vector<int> m_num_eqs, m_var_types, m_dxi_offsets;
  std::vector<std::vector<T>> m_xi, m_xi_prev, m_R;
  std::vector<T> m_params;
  std::vector<std::vector<double>> m_param_values;
  std::vector<std::vector<int>> m_active_indices;
  int m_max_iters = 0;
  double m_abs_tol = 0., m_rel_tol = 0.;
  enum { ELASTIC = 0, PLASTIC = 1 };
};

template <> inline void LocalResidual<double>::seed_chain(EMatrix const&) {}
template <> inline void LocalResidual<Fad>::seed_chain(EMatrix const& dxi_dx) {
  int const ng = dxi_dx.cols();
  for (int i = 0; i < m_num_residuals; ++i)
    for (int eq = 0; eq < m_num_eqs[i]; ++eq) {
      int const xi_idx = dxi_idx(i, eq);
      double const v = m_xi[i][eq].v;
      m_xi[i][eq].diff(0, ng);
      m_xi[i][eq].v = v;
      for (int k = 0; k < ng; ++k) m_xi[i][eq].d[k] = dxi_dx(xi_idx, k);
    }
}

// material_params.hpp:12-30
template <class T> T compute_mu(T const& E, T const& nu) { return E / (2. * (1. + nu)); }
template <class T> T compute_kappa(T const& E, T const& nu) { return E / (3. * (1. - 2. * (nu))); }
template <class T> T compute_lambda(T const& E, T const& nu) {
  return E * nu / ((1. + nu) * (1. - 2. * nu));
}

// ---------------------------------------------------------------------------
// Mechanics  (src/mechanics.cpp)
// ---------------------------------------------------------------------------
template <class T>
class Mechanics : public GlobalResidual<T> {
 public:
  Mechanics(int ndims, bool is_mixed, double stab_mult) {
    m_mixed = is_mixed;
    int const nr = is_mixed ? 2 : 1;
    this->m_num_residuals = nr;
    this->m_num_eqs.resize(nr);
    this->m_var_types.resize(nr);
    this->m_var_types[0] = VECTOR;
    this->m_num_eqs[0] = get_num_eqs(VECTOR, ndims);
    if (is_mixed) {
      this->m_var_types[1] = SCALAR;
      this->m_num_eqs[1] = 1;
      this->m_ip_sets = {1, 2};
      m_stabilization_multiplier = stab_mult;
    } else {
      this->m_ip_sets = {1};
    }
  }

  // src/mechanics.cpp:61-101
  void compute_kinematics() override {
    int const nd = this->m_num_dims;
    if (this->m_F.get_dimension() != nd) {
      this->m_F = Tensor<T>(nd);
      this->m_F_prev = Tensor<T>(nd);
      this->m_cof_F = Tensor<T>(nd);
    }
    for (int k = 0; k < nd; ++k) {
      for (int l = 0; l < nd; ++l) {
        this->m_F(k, l) = this->m_grad_x[0][k][l];
        this->m_F_prev(k, l) = this->m_grad_x_prev[0][k][l];
      }
      this->m_F(k, k) += T(1.0);
      this->m_F_prev(k, k) += T(1.0);
    }
    this->m_det_F = det(this->m_F);
    Tensor<T> const& F = this->m_F;
    Tensor<T>& C = this->m_cof_F;
    if (nd == 3) {
      C(0,0) =  F(1,1)*F(2,2) - F(1,2)*F(2,1);
      C(0,1) = -F(1,0)*F(2,2) + F(1,2)*F(2,0);
      C(0,2) =  F(1,0)*F(2,1) - F(1,1)*F(2,0);
      C(1,0) = -F(0,1)*F(2,2) + F(0,2)*F(2,1);
      C(1,1) =  F(0,0)*F(2,2) - F(0,2)*F(2,0);
      C(1,2) = -F(0,0)*F(2,1) + F(0,1)*F(2,0);
      C(2,0) =  F(0,1)*F(1,2) - F(0,2)*F(1,1);
      C(2,1) = -F(0,0)*F(1,2) + F(0,2)*F(1,0);
      C(2,2) =  F(0,0)*F(1,1) - F(0,1)*F(1,0);
    } else {
      C(0,0) =  F(1,1);
      C(0,1) = -F(1,0);
      C(1,0) = -F(0,1);
      C(1,1) =  F(0,0);
    }
  }

  // src/mechanics.cpp:115-145
  void evaluate_displacement(LocalResidual<T>& local, double w, double dv) {
    int const nd = this->m_num_dims, nn = this->m_num_nodes;
    Tensor<T> stress = local.cauchy(*this);
    if (local.is_finite_deformation()) stress = stress * this->cof_F();
    for (int n = 0; n < nn; ++n)
      for (int i = 0; i < nd; ++i)
        for (int j = 0; j < nd; ++j) {
          double const dbasis_dx = this->grad_weight(0, n, i, j);
          this->R_nodal(0, n, i) += stress(i, j) * dbasis_dx * w * dv;
        }
  }

  // src/mechanics.cpp:147-227
  void evaluate_mixed(LocalResidual<T>& local, double w, double dv, int ip_set) {
    int const nd = this->m_num_dims, nn = this->m_num_nodes;
    int const pressure_idx = 1;
    T const E = local.params(0);
    T const nu = local.params(1);
    T const mu = compute_mu(E, nu);
    T const p = this->scalar_x(pressure_idx);
    T pressure_scale_factor = local.pressure_scale_factor();
    if (ip_set == 0) {
      Vec<T> const grad_p = this->grad_scalar_x(pressure_idx);
      Tensor<T> const I = eye<T>(nd);
      T hydro_cauchy = local.hydro_cauchy(*this);
      for (int n = 0; n < nn; ++n) {
        double const basis = this->weight(pressure_idx, n, 0);
        this->R_nodal(pressure_idx, n, 0) -=
            hydro_cauchy / pressure_scale_factor * basis * w * dv;
      }
      double const h = this->m_geom.h;  // CURRENT mode, src/mechanics.cpp:103-113
      T const tau = m_stabilization_multiplier * 0.5 * h * h / mu;
      Tensor<T> stab_matrix = tau * I;
      if (local.is_finite_deformation()) {
        Tensor<T> const& cof_F = this->cof_F();
        stab_matrix = stab_matrix * (transpose(cof_F) * cof_F) / this->det_F();
      }
      for (int n = 0; n < nn; ++n)
        for (int i = 0; i < nd; ++i)
          for (int j = 0; j < nd; ++j) {
            double const dbasis_dx = this->grad_weight(pressure_idx, n, 0, i);
            this->R_nodal(pressure_idx, n, 0) -=
                stab_matrix(i, j) * grad_p(j) * dbasis_dx * w * dv;
          }
    } else if (ip_set == 1) {
      for (int n = 0; n < nn; ++n) {
        double const basis = this->weight(pressure_idx, n, 0);
        this->R_nodal(pressure_idx, n, 0) -= p / pressure_scale_factor * basis * w * dv;
      }
    } else {
      throw std::runtime_error("unimplemented ip set");
    }
  }

  // src/mechanics.cpp:229-240
  void evaluate(LocalResidual<T>& local, double const*, double w, double dv,
                int ip_set) override {
    if (ip_set == 0) evaluate_displacement(local, w, dv);
    if (m_mixed) evaluate_mixed(local, w, dv, ip_set);
  }

 private:
  bool m_mixed = true;
  double m_stabilization_multiplier = 1.;
};

// ---------------------------------------------------------------------------
// MechanicsPlaneStress  (src/mechanics_plane_stress.cpp)
// ---------------------------------------------------------------------------
template <class T>
class MechanicsPlaneStress : public GlobalResidual<T> {
 public:
  MechanicsPlaneStress(int ndims, double thickness) {
    m_thickness = thickness;
    this->m_num_residuals = 1;
    this->m_num_eqs = {get_num_eqs(VECTOR, ndims)};
    this->m_var_types = {VECTOR};
    this->m_ip_sets = {1};
  }
  // src/mechanics_plane_stress.cpp:46-95
  void evaluate(LocalResidual<T>& local, double const*, double w, double dv,
                int ip_set) override {
    int const nd = this->m_num_dims, nn = this->m_num_nodes;
    if (ip_set != 0) throw std::runtime_error("plane stress: ip_set != 0");
    Tensor<T> stress = local.cauchy(*this);
    if (local.is_finite_deformation()) {
      Tensor<T> const grad_u = this->grad_vector_x(0);
      Tensor<T> const I = eye<T>(nd);
      Tensor<T> const F = grad_u + I;
      Tensor<T> const F_inv = inverse(F);
      Tensor<T> const F_invT = transpose(F_inv);
      T const J = det(F);
      T const z_stretch = local.scalar_xi(local.z_stretch_idx());
      stress = (z_stretch * J) * stress * F_invT;
    }
    for (int n = 0; n < nn; ++n)
      for (int i = 0; i < nd; ++i)
        for (int j = 0; j < nd; ++j) {
          double const dbasis_dx = this->grad_weight(0, n, i, j);
          this->R_nodal(0, n, i) += stress(i, j) * dbasis_dx * w * m_thickness * dv;
        }
  }

 private:
  double m_thickness = 1.;
};

}  // namespace orc
