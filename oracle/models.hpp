// ORACLE (test infrastructure only).
//
// CPU restatement of calibr8's in-scope constitutive models:
//   Elastic               src/elastic.cpp:27-139
//   SmallJ2               src/small_J2.cpp:31-297
//   SmallHill             src/small_hill.cpp:38-324 + src/yield_functions.hpp:34-99
//   SmallHillPlaneStress  src/small_hill_plane_stress.cpp:36-327
//   SmallHillPlaneStrain  src/small_hill_plane_strain.cpp:36-331
//   HyperJ2               src/hyper_J2.cpp:40-360
//   HyperJ2PlaneStress    src/hyper_J2_plane_stress.cpp:42-411
//   HyperJ2PlaneStrain    src/hyper_J2_plane_strain.cpp:40-372
#pragma once
#include "residuals.hpp"

namespace orc {

// ---- src/yield_functions.hpp:8-99 ----
template <class T> Tensor<T> insert_2D_tensor_into_3D(Tensor<T> const& t2) {
  Tensor<T> t3 = zero<T>(3);
  for (int i = 0; i < 2; ++i) for (int j = 0; j < 2; ++j) t3(i, j) = t2(i, j);
  return t3;
}
template <class T> Tensor<T> extract_2D_tensor_from_3D(Tensor<T> const& t3) {
  Tensor<T> t2 = zero<T>(2);
  for (int i = 0; i < 2; ++i) for (int j = 0; j < 2; ++j) t2(i, j) = t3(i, j);
  return t2;
}
template <class T> struct HillParams { T p[6]; };
template <class T>
HillParams<T> compute_hill_params(T const& R00, T const& R11, T const& R22,
                                  T const& R01, T const& R02, T const& R12) {
  HillParams<T> h;
  h.p[0] = 0.5 * (pow(R11, -2) + pow(R22, -2) - pow(R00, -2));
  h.p[1] = 0.5 * (pow(R22, -2) + pow(R00, -2) - pow(R11, -2));
  h.p[2] = 0.5 * (pow(R00, -2) + pow(R11, -2) - pow(R22, -2));
  h.p[3] = 1.5 * pow(R12, -2);
  h.p[4] = 1.5 * pow(R02, -2);
  h.p[5] = 1.5 * pow(R01, -2);
  return h;
}
template <class T> T compute_hill_value(Tensor<T> const& TC, HillParams<T> const& hp) {
  T const F = hp.p[0], G = hp.p[1], H = hp.p[2], L = hp.p[3], M = hp.p[4], N = hp.p[5];
  T const hill = sqrt(F * pow(TC(1, 1) - TC(2, 2), 2) + G * pow(TC(2, 2) - TC(0, 0), 2) +
                      H * pow(TC(0, 0) - TC(1, 1), 2) +
                      2. * (L * pow(TC(1, 2), 2) + M * pow(TC(0, 2), 2) + N * pow(TC(0, 1), 2)));
  return hill;
}
template <class T>
Tensor<T> compute_hill_normal(Tensor<T> const& TC, HillParams<T> const& hp, T const& hill_value) {
  T const F = hp.p[0], G = hp.p[1], H = hp.p[2], L = hp.p[3], M = hp.p[4], N = hp.p[5];
  Tensor<T> n = zero<T>(3);
  n(0, 0) = (G + H) * TC(0, 0) - H * TC(1, 1) - G * TC(2, 2);
  n(1, 1) = (F + H) * TC(1, 1) - H * TC(0, 0) - F * TC(2, 2);
  n(2, 2) = (G + F) * TC(2, 2) - G * TC(0, 0) - F * TC(1, 1);
  n(0, 1) = N * TC(0, 1);
  n(0, 2) = M * TC(0, 2);
  n(1, 2) = L * TC(1, 2);
  n(1, 0) = n(0, 1);
  n(2, 0) = n(0, 2);
  n(2, 1) = n(1, 2);
  n /= hill_value;
  return n;
}

// ---------------------------------------------------------------------------
template <class T>
class Elastic : public LocalResidual<T> {
 public:
  explicit Elastic(int) {
    this->m_num_residuals = 1;
    this->m_num_eqs = {1};
    this->m_var_types = {SCALAR};
  }
  void init_variables_impl() override { this->set_scalar_xi(0, T(0.)); }
  int solve_nonlinear(GlobalResidual<T>&) override {  // src/elastic.cpp:76-80
    this->set_scalar_xi(0, T(0.));
    return 0;
  }
  int evaluate(GlobalResidual<T>&, bool, int) override { return 0; }  // :82-90
  bool is_finite_deformation() override { return false; }
  Tensor<T> cauchy(GlobalResidual<T>& g) override {  // :92-103
    T const p = g.scalar_x(1);
    Tensor<T> const I = eye<T>(g.num_dims());
    return this->dev_cauchy(g) - p * I;
  }
  Tensor<T> dev_cauchy(GlobalResidual<T>& g) override {  // :105-116
    Tensor<T> const I = eye<T>(g.num_dims());
    T const mu = compute_mu(this->m_params[0], this->m_params[1]);
    Tensor<T> const grad_u = g.grad_vector_x(0);
    Tensor<T> const eps = 0.5 * (grad_u + transpose(grad_u));
    Tensor<T> const dev_eps = eps - (trace(eps) / 3.) * I;
    return 2. * mu * dev_eps;
  }
  T hydro_cauchy(GlobalResidual<T>& g) override {  // :118-128
    T const E = this->m_params[0], nu = this->m_params[1];
    T const kappa = compute_kappa(E, nu);
    T const cte = this->m_params[2], delta_T = this->m_params[3];
    Tensor<T> const grad_u = g.grad_vector_x(0);
    Tensor<T> const eps = 0.5 * (grad_u + transpose(grad_u));
    return kappa * trace(eps) - cte * delta_T * E / (1. - 2. * nu);
  }
  T pressure_scale_factor() override {
    return compute_kappa(this->m_params[0], this->m_params[1]);
  }
};

// ---------------------------------------------------------------------------
template <class T>
class SmallJ2 : public LocalResidual<T> {
 public:
  explicit SmallJ2(int ndims) {
    this->m_num_residuals = 2;
    this->m_num_eqs = {get_num_eqs(SYM_TENSOR, ndims), 1};
    this->m_var_types = {SYM_TENSOR, SCALAR};
  }
  void init_variables_impl() override {  // src/small_J2.cpp:102-115
    this->set_scalar_xi(1, T(0.));
    this->set_sym_tensor_xi(0, zero<T>(this->m_num_dims));
  }
  int solve_nonlinear(GlobalResidual<T>& g) override {  // :121-173
    Tensor<T> const pstrain_old = this->sym_tensor_xi_prev(0);
    T const alpha_old = this->scalar_xi_prev(1);
    this->set_sym_tensor_xi(0, pstrain_old);
    this->set_scalar_xi(1, alpha_old);
    return this->newton(g);
  }
  int evaluate(GlobalResidual<T>& g, bool force_path, int path_in) override {  // :180-250
    int path = this->ELASTIC;
    double const sqrt_23 = std::sqrt(2. / 3.);
    double const sqrt_32 = std::sqrt(3. / 2.);
    T const E = this->m_params[0], nu = this->m_params[1];
    T const K = this->m_params[2], Y = this->m_params[3];
    T const mu = compute_mu(E, nu);
    Tensor<T> const pstrain_old = this->sym_tensor_xi_prev(0);
    T const alpha_old = this->scalar_xi_prev(1);
    Tensor<T> const pstrain = this->sym_tensor_xi(0);
    T const alpha = this->scalar_xi(1);
    Tensor<T> const s = this->dev_cauchy(g);
    T const s_mag = norm(s);
    Tensor<T> const n = s / s_mag;
    T const sigma_yield = Y + K * alpha;
    T const f = (s_mag - sqrt_23 * sigma_yield) / val(mu);
    Tensor<T> R_pstrain;
    T R_alpha;
    bool plastic;
    if (!force_path) plastic = (f > this->m_abs_tol || std::abs(val(f)) < this->m_abs_tol);
    else plastic = (path_in == this->PLASTIC);
    if (plastic) {
      T const dgam = sqrt_32 * (alpha - alpha_old);
      R_pstrain = pstrain - pstrain_old - dgam * n;
      R_alpha = f;
      path = this->PLASTIC;
    } else {
      R_pstrain = pstrain - pstrain_old;
      R_alpha = alpha - alpha_old;
      path = this->ELASTIC;
    }
    this->set_sym_tensor_R(0, R_pstrain);
    this->set_scalar_R(1, R_alpha);
    return path;
  }
  bool is_finite_deformation() override { return false; }
  Tensor<T> cauchy(GlobalResidual<T>& g) override {  // :252-263
    T const p = g.scalar_x(1);
    Tensor<T> const I = eye<T>(g.num_dims());
    return this->dev_cauchy(g) - p * I;
  }
  Tensor<T> dev_cauchy(GlobalResidual<T>& g) override {  // :265-277
    Tensor<T> const I = eye<T>(g.num_dims());
    T const mu = compute_mu(this->m_params[0], this->m_params[1]);
    Tensor<T> const pstrain = this->sym_tensor_xi(0);
    Tensor<T> const grad_u = g.grad_vector_x(0);
    Tensor<T> const eps = 0.5 * (grad_u + transpose(grad_u));
    Tensor<T> const dev_eps = eps - (trace(eps) / 3.) * I;
    return 2. * mu * (dev_eps - pstrain);
  }
  T hydro_cauchy(GlobalResidual<T>& g) override {  // :279-289
    T const E = this->m_params[0], nu = this->m_params[1];
    T const kappa = compute_kappa(E, nu);
    T const cte = this->m_params[4], delta_T = this->m_params[5];
    Tensor<T> const grad_u = g.grad_vector_x(0);
    Tensor<T> const eps = 0.5 * (grad_u + transpose(grad_u));
    return kappa * trace(eps) - cte * delta_T * E / (1. - 2. * nu);
  }
  T pressure_scale_factor() override {
    return compute_kappa(this->m_params[0], this->m_params[1]);
  }
};

// ---------------------------------------------------------------------------
template <class T>
class SmallHill : public LocalResidual<T> {
 public:
  explicit SmallHill(int ndims) {
    this->m_num_residuals = 2;
    this->m_num_eqs = {get_num_eqs(SYM_TENSOR, ndims), 1};
    this->m_var_types = {SYM_TENSOR, SCALAR};
  }
  void init_variables_impl() override {
    this->set_scalar_xi(1, T(0.));
    this->set_sym_tensor_xi(0, zero<T>(this->m_num_dims));
  }
  int solve_nonlinear(GlobalResidual<T>& g) override {  // src/small_hill.cpp:137-189
    this->set_sym_tensor_xi(0, this->sym_tensor_xi_prev(0));
    this->set_scalar_xi(1, this->scalar_xi_prev(1));
    return this->newton(g);
  }
  int evaluate(GlobalResidual<T>& g, bool force_path, int path_in) override {  // :196-276
    int path = this->ELASTIC;
    T const E = this->m_params[0], nu = this->m_params[1], Y = this->m_params[2];
    T const R00 = this->m_params[3], R11 = this->m_params[4], R22 = this->m_params[5];
    T const R01 = this->m_params[6], R02 = this->m_params[7], R12 = this->m_params[8];
    T const S = this->m_params[9], D = this->m_params[10];
    T const mu = compute_mu(E, nu);
    HillParams<T> const hp = compute_hill_params(R00, R11, R22, R01, R02, R12);
    Tensor<T> const pstrain_old = this->sym_tensor_xi_prev(0);
    T const alpha_old = this->scalar_xi_prev(1);
    Tensor<T> const pstrain = this->sym_tensor_xi(0);
    T const alpha = this->scalar_xi(1);
    Tensor<T> const s = this->dev_cauchy(g);
    T const hill = compute_hill_value(s, hp);
    T const sigma_yield = Y + S * (1. - exp(-D * alpha));
    T const f = (hill - sigma_yield) / val(mu);
    Tensor<T> R_pstrain;
    T R_alpha;
    bool plastic;
    if (!force_path) plastic = (f > this->m_abs_tol || std::abs(val(f)) < this->m_abs_tol);
    else plastic = (path_in == this->PLASTIC);
    if (plastic) {
      Tensor<T> const n = compute_hill_normal(s, hp, hill);
      T const dgam = alpha - alpha_old;
      R_pstrain = pstrain - pstrain_old - dgam * n;
      R_pstrain(2, 2) = trace(pstrain);
      R_alpha = f;
      path = this->PLASTIC;
    } else {
      R_pstrain = pstrain - pstrain_old;
      R_alpha = alpha - alpha_old;
      path = this->ELASTIC;
    }
    this->set_sym_tensor_R(0, R_pstrain);
    this->set_scalar_R(1, R_alpha);
    return path;
  }
  bool is_finite_deformation() override { return false; }
  Tensor<T> cauchy(GlobalResidual<T>& g) override {
    T const p = g.scalar_x(1);
    Tensor<T> const I = eye<T>(g.num_dims());
    return this->dev_cauchy(g) - p * I;
  }
  Tensor<T> dev_cauchy(GlobalResidual<T>& g) override {  // :291-303
    Tensor<T> const I = eye<T>(g.num_dims());
    T const mu = compute_mu(this->m_params[0], this->m_params[1]);
    Tensor<T> const pstrain = this->sym_tensor_xi(0);
    Tensor<T> const grad_u = g.grad_vector_x(0);
    Tensor<T> const eps = 0.5 * (grad_u + transpose(grad_u));
    Tensor<T> const dev_eps = eps - (trace(eps) / 3.) * I;
    return 2. * mu * (dev_eps - pstrain);
  }
  T hydro_cauchy(GlobalResidual<T>& g) override {  // :305-313
    T const kappa = compute_kappa(this->m_params[0], this->m_params[1]);
    Tensor<T> const grad_u = g.grad_vector_x(0);
    Tensor<T> const eps = 0.5 * (grad_u + transpose(grad_u));
    return kappa * trace(eps);
  }
  T pressure_scale_factor() override {
    return compute_kappa(this->m_params[0], this->m_params[1]);
  }
};

// ---------------------------------------------------------------------------
template <class T>
class SmallHillPlaneStress : public LocalResidual<T> {
 public:
  explicit SmallHillPlaneStress(int ndims) {
    this->m_num_residuals = 2;
    this->m_num_eqs = {get_num_eqs(SYM_TENSOR, ndims), 1};
    this->m_var_types = {SYM_TENSOR, SCALAR};
  }
  void init_variables_impl() override {
    this->set_scalar_xi(1, T(0.));
    this->set_sym_tensor_xi(0, zero<T>(this->m_num_dims));
  }
  int solve_nonlinear(GlobalResidual<T>& g) override {  // src/small_hill_plane_stress.cpp:135-187
    this->set_sym_tensor_xi(0, this->sym_tensor_xi_prev(0));
    this->set_scalar_xi(1, this->scalar_xi_prev(1));
    return this->newton(g);
  }
  int evaluate(GlobalResidual<T>& g, bool force_path, int path_in) override {  // :194-276
    int path = this->ELASTIC;
    T const E = this->m_params[0], nu = this->m_params[1], Y = this->m_params[2];
    T const S = this->m_params[3], D = this->m_params[4];
    T const R00 = this->m_params[5], R11 = this->m_params[6], R22 = this->m_params[7];
    T const R01 = this->m_params[8];
    T const mu = compute_mu(E, nu);
    T const R02 = 1., R12 = 1.;
    HillParams<T> const hp = compute_hill_params(R00, R11, R22, R01, R02, R12);
    Tensor<T> const pstrain_old = this->sym_tensor_xi_prev(0);
    T const alpha_old = this->scalar_xi_prev(1);
    Tensor<T> const pstrain = this->sym_tensor_xi(0);
    T const alpha = this->scalar_xi(1);
    Tensor<T> const sigma_2D = this->cauchy(g);
    Tensor<T> sigma_3D = insert_2D_tensor_into_3D(sigma_2D);
    T const hill = compute_hill_value(sigma_3D, hp);
    T const sigma_yield = Y + S * (1. - exp(-D * alpha));
    T const f = (hill - sigma_yield) / val(mu);
    Tensor<T> R_pstrain;
    T R_alpha;
    bool plastic;
    if (!force_path) plastic = (f > this->m_abs_tol || std::abs(val(f)) < this->m_abs_tol);
    else plastic = (path_in == this->PLASTIC);
    if (plastic) {
      Tensor<T> const n_3D = compute_hill_normal(sigma_3D, hp, hill);
      Tensor<T> const n_2D = extract_2D_tensor_from_3D(n_3D);
      T const dgam = alpha - alpha_old;
      R_pstrain = pstrain - pstrain_old - dgam * n_2D;
      R_alpha = f;
      path = this->PLASTIC;
    } else {
      R_pstrain = pstrain - pstrain_old;
      R_alpha = alpha - alpha_old;
      path = this->ELASTIC;
    }
    this->set_sym_tensor_R(0, R_pstrain);
    this->set_scalar_R(1, R_alpha);
    return path;
  }
  bool is_finite_deformation() override { return false; }
  Tensor<T> cauchy(GlobalResidual<T>& g) override {  // :278-294
    Tensor<T> const I = eye<T>(g.num_dims());
    T const E = this->m_params[0], nu = this->m_params[1];
    T const mu = compute_mu(E, nu);
    T const lambda = compute_lambda(E, nu);
    Tensor<T> const grad_u = g.grad_vector_x(0);
    Tensor<T> const epsilon = 0.5 * (grad_u + transpose(grad_u));
    Tensor<T> const pstrain = this->sym_tensor_xi(0);
    T const eps_zz = this->epsilon_zz(g);
    T const epsilon_kk = trace(epsilon) + eps_zz;
    Tensor<T> const sigma = lambda * epsilon_kk * I + 2. * mu * (epsilon - pstrain);
    return sigma;
  }
  Tensor<T> dev_cauchy(GlobalResidual<T>& g) override {
    Tensor<T> const I = eye<T>(g.num_dims());
    Tensor<T> c = this->cauchy(g);
    T h = this->hydro_cauchy(g);
    return c - h * I;
  }
  T hydro_cauchy(GlobalResidual<T>& g) override {
    Tensor<T> c = this->cauchy(g);
    return trace(c) / 3.;
  }
  T pressure_scale_factor() override { return T(0.); }
  T epsilon_zz(GlobalResidual<T>& g) {  // :314-327
    T const E = this->m_params[0], nu = this->m_params[1];
    T const mu = compute_mu(E, nu);
    T const lambda = compute_lambda(E, nu);
    Tensor<T> const grad_u = g.grad_vector_x(0);
    Tensor<T> const epsilon = 0.5 * (grad_u + transpose(grad_u));
    Tensor<T> const pstrain = this->sym_tensor_xi(0);
    T const eps_zz = -(lambda * trace(epsilon) + 2. * mu * trace(pstrain)) / (lambda + 2. * mu);
    return eps_zz;
  }
};

// ---------------------------------------------------------------------------
template <class T>
class SmallHillPlaneStrain : public LocalResidual<T> {
 public:
  explicit SmallHillPlaneStrain(int ndims) {
    this->m_num_residuals = 2;
    this->m_num_eqs = {get_num_eqs(SYM_TENSOR, ndims), 1};
    this->m_var_types = {SYM_TENSOR, SCALAR};
  }
  void init_variables_impl() override {
    this->set_scalar_xi(1, T(0.));
    this->set_sym_tensor_xi(0, zero<T>(this->m_num_dims));
  }
  int solve_nonlinear(GlobalResidual<T>& g) override {
    this->set_sym_tensor_xi(0, this->sym_tensor_xi_prev(0));
    this->set_scalar_xi(1, this->scalar_xi_prev(1));
    return this->newton(g);
  }
  int evaluate(GlobalResidual<T>& g, bool force_path, int path_in) override {  // src/small_hill_plane_strain.cpp:194-276
    int path = this->ELASTIC;
    T const E = this->m_params[0], nu = this->m_params[1], Y = this->m_params[2];
    T const S = this->m_params[3], D = this->m_params[4];
    T const R00 = this->m_params[5], R11 = this->m_params[6], R22 = this->m_params[7];
    T const R01 = this->m_params[8];
    T const mu = compute_mu(E, nu);
    T const R02 = 1., R12 = 1.;
    HillParams<T> const hp = compute_hill_params(R00, R11, R22, R01, R02, R12);
    Tensor<T> const pstrain_old = this->sym_tensor_xi_prev(0);
    T const alpha_old = this->scalar_xi_prev(1);
    Tensor<T> const pstrain = this->sym_tensor_xi(0);
    T const alpha = this->scalar_xi(1);
    Tensor<T> const s_2D = this->dev_cauchy(g);
    Tensor<T> const grad_u = g.grad_vector_x(0);
    Tensor<T> const epsilon = 0.5 * (grad_u + transpose(grad_u));
    T const s_zz = 2. * mu * (-trace(epsilon) / 3. + trace(pstrain));
    Tensor<T> s_3D = insert_2D_tensor_into_3D(s_2D);
    s_3D(2, 2) = s_zz;
    T const hill = compute_hill_value(s_3D, hp);
    T const sigma_yield = Y + S * (1. - exp(-D * alpha));
    T const f = (hill - sigma_yield) / val(mu);
    Tensor<T> R_pstrain;
    T R_alpha;
    bool plastic;
    if (!force_path) plastic = (f > this->m_abs_tol || std::abs(val(f)) < this->m_abs_tol);
    else plastic = (path_in == this->PLASTIC);
    if (plastic) {
      Tensor<T> const n_3D = compute_hill_normal(s_3D, hp, hill);
      Tensor<T> const n_2D = extract_2D_tensor_from_3D(n_3D);
      T const dgam = alpha - alpha_old;
      R_pstrain = pstrain - pstrain_old - dgam * n_2D;
      R_alpha = f;
      path = this->PLASTIC;
    } else {
      R_pstrain = pstrain - pstrain_old;
      R_alpha = alpha - alpha_old;
      path = this->ELASTIC;
    }
    this->set_sym_tensor_R(0, R_pstrain);
    this->set_scalar_R(1, R_alpha);
    return path;
  }
  bool is_finite_deformation() override { return false; }
  Tensor<T> cauchy(GlobalResidual<T>& g) override {
    T const p = g.scalar_x(1);
    Tensor<T> const I = eye<T>(g.num_dims());
    return this->dev_cauchy(g) - p * I;
  }
  Tensor<T> dev_cauchy(GlobalResidual<T>& g) override {
    Tensor<T> const I = eye<T>(g.num_dims());
    T const mu = compute_mu(this->m_params[0], this->m_params[1]);
    Tensor<T> const pstrain = this->sym_tensor_xi(0);
    Tensor<T> const grad_u = g.grad_vector_x(0);
    Tensor<T> const eps = 0.5 * (grad_u + transpose(grad_u));
    Tensor<T> const dev_eps = eps - (trace(eps) / 3.) * I;
    return 2. * mu * (dev_eps - pstrain);
  }
  T hydro_cauchy(GlobalResidual<T>& g) override {
    T const kappa = compute_kappa(this->m_params[0], this->m_params[1]);
    Tensor<T> const grad_u = g.grad_vector_x(0);
    Tensor<T> const eps = 0.5 * (grad_u + transpose(grad_u));
    return kappa * trace(eps);
  }
  T pressure_scale_factor() override {
    return compute_kappa(this->m_params[0], this->m_params[1]);
  }
};

// ---------------------------------------------------------------------------
// src/hyper_J2.cpp:136-154
template <class T>
Tensor<T> eval_be_bar(GlobalResidual<T>& g, Tensor<T> const& zeta, T const& Ie) {
  int const nd = g.num_dims();
  Tensor<T> const I = eye<T>(nd);
  Tensor<T> const grad_u = g.grad_vector_x(0);
  Tensor<T> const grad_u_prev = g.grad_vector_x_prev(0);
  Tensor<T> const F = grad_u + I;
  Tensor<T> const F_prev = grad_u_prev + I;
  Tensor<T> const rF = F * inverse(F_prev);
  T const det_rF = det(rF);
  T const det_rF_13 = cbrt(det_rF);
  Tensor<T> const rF_bar = rF / det_rF_13;
  Tensor<T> const rF_barT = transpose(rF_bar);
  Tensor<T> const be_bar = rF_bar * (zeta + Ie * I) * rF_barT;
  return be_bar;
}

template <class T>
class HyperJ2 : public LocalResidual<T> {
 public:
  explicit HyperJ2(int ndims) {
    this->m_num_residuals = 3;
    this->m_num_eqs = {get_num_eqs(SYM_TENSOR, ndims), 1, 1};
    this->m_var_types = {SYM_TENSOR, SCALAR, SCALAR};
  }
  void init_variables_impl() override {  // src/hyper_J2.cpp:118-134
    this->set_scalar_xi(1, T(1.0));
    this->set_scalar_xi(2, T(0.0));
    this->set_sym_tensor_xi(0, zero<T>(this->m_num_dims));
  }
  int solve_nonlinear(GlobalResidual<T>& g) override {  // :161-218
    Tensor<T> const zeta_old = this->sym_tensor_xi_prev(0);
    T const Ie_old = this->scalar_xi_prev(1);
    T const alpha_old = this->scalar_xi_prev(2);
    Tensor<T> const be_bar_trial = eval_be_bar(g, zeta_old, Ie_old);
    Tensor<T> const zeta = dev(be_bar_trial);
    T const Ie = trace(be_bar_trial) / 3.;
    this->set_sym_tensor_xi(0, zeta);
    this->set_scalar_xi(1, Ie);
    this->set_scalar_xi(2, alpha_old);
    return this->newton(g);
  }
  int evaluate(GlobalResidual<T>& g, bool force_path, int path_in) override {  // :225-314
    int path = this->ELASTIC;
    int const nd = this->m_num_dims;
    double const sqrt_23 = std::sqrt(2. / 3.);
    double const sqrt_32 = std::sqrt(3. / 2.);
    T const E = this->m_params[0], nu = this->m_params[1], Y = this->m_params[2];
    T const S = this->m_params[3], D = this->m_params[4], A = this->m_params[5];
    T const n = this->m_params[6], K = this->m_params[7];
    T const mu = compute_mu(E, nu);
    Tensor<T> const zeta_old = this->sym_tensor_xi_prev(0);
    T const Ie_old = this->scalar_xi_prev(1);
    T const alpha_old = this->scalar_xi_prev(2);
    Tensor<T> const zeta = this->sym_tensor_xi(0);
    T const Ie = this->scalar_xi(1);
    T const alpha = this->scalar_xi(2);
    Tensor<T> const I = eye<T>(nd);
    Tensor<T> const be_bar_trial = eval_be_bar(g, zeta_old, Ie_old);
    Tensor<T> const s = mu * zeta;
    T const s_mag = norm(s);
    double const power_law_offset = 1e-12;
    T const sigma_yield = Y + S * (1. - exp(-D * alpha)) +
                          A * pow(alpha + power_law_offset, n) + K * alpha;
    T const f = (s_mag - sqrt_23 * sigma_yield) / val(mu);
    Tensor<T> R_zeta;
    T R_Ie, R_alpha;
    bool plastic;
    if (!force_path) plastic = (f > this->m_abs_tol || std::abs(val(f)) < this->m_abs_tol);
    else plastic = (path_in == this->PLASTIC);
    if (plastic) {
      Tensor<T> const nn = s / s_mag;
      T const dgam = sqrt_32 * (alpha - alpha_old);
      R_zeta = zeta - dev(be_bar_trial) + 2. * dgam * Ie * nn;
      R_Ie = det(zeta + Ie * I) - 1.;
      R_alpha = f;
      path = this->PLASTIC;
    } else {
      R_zeta = zeta - dev(be_bar_trial);
      R_Ie = Ie - trace(be_bar_trial) / 3.;
      R_alpha = alpha - alpha_old;
      path = this->ELASTIC;
    }
    this->set_sym_tensor_R(0, R_zeta);
    this->set_scalar_R(1, R_Ie);
    this->set_scalar_R(2, R_alpha);
    return path;
  }
  bool is_finite_deformation() override { return true; }
  Tensor<T> cauchy(GlobalResidual<T>& g) override {  // :315-324
    T const p = g.scalar_x(1);
    Tensor<T> const I = eye<T>(g.num_dims());
    return this->dev_cauchy(g) - p * I;
  }
  Tensor<T> dev_cauchy(GlobalResidual<T>& g) override {  // :326-338
    T const mu = compute_mu(this->m_params[0], this->m_params[1]);
    Tensor<T> const I = eye<T>(g.num_dims());
    Tensor<T> const F = g.grad_vector_x(0) + I;
    Tensor<T> const zeta = this->sym_tensor_xi(0);
    T const J = det(F);
    return mu * zeta / J;
  }
  T hydro_cauchy(GlobalResidual<T>& g) override {  // :340-352
    T const kappa = compute_kappa(this->m_params[0], this->m_params[1]);
    Tensor<T> const I = eye<T>(g.num_dims());
    Tensor<T> const F = g.grad_vector_x(0) + I;
    T const J = det(F);
    return kappa / 2. * (J - 1. / J);
  }
  T pressure_scale_factor() override {
    return compute_kappa(this->m_params[0], this->m_params[1]);
  }
};

// ---------------------------------------------------------------------------
// src/hyper_J2_plane_stress.cpp:140-169
template <class T>
void eval_be_bar_plane_stress(GlobalResidual<T>& g, Tensor<T> const& zeta_2D, T const& Ie,
                              T const& lambda_z_prev, T const& lambda_z, T& J_2D,
                              Tensor<T>& be_bar) {
  Tensor<T> const I_2D = eye<T>(2);
  Tensor<T> const I = eye<T>(3);
  Tensor<T> const grad_u = g.grad_vector_x(0);
  Tensor<T> const grad_u_prev = g.grad_vector_x_prev(0);
  Tensor<T> const F_2D = grad_u + I_2D;
  J_2D = det(F_2D);
  Tensor<T> const F_prev_2D = grad_u_prev + I_2D;
  Tensor<T> F_3D = insert_2D_tensor_into_3D(F_2D);
  Tensor<T> F_prev_3D = insert_2D_tensor_into_3D(F_prev_2D);
  F_3D(2, 2) = lambda_z;
  F_prev_3D(2, 2) = lambda_z_prev;
  Tensor<T> const rF = F_3D * inverse(F_prev_3D);
  T const det_rF = det(rF);
  T const det_rF_13 = cbrt(det_rF);
  Tensor<T> const rF_bar = rF / det_rF_13;
  Tensor<T> const rF_barT = transpose(rF_bar);
  Tensor<T> zeta_3D = insert_2D_tensor_into_3D(zeta_2D);
  zeta_3D(2, 2) = -trace(zeta_2D);
  be_bar = rF_bar * (zeta_3D + Ie * I) * rF_barT;
}

template <class T>
class HyperJ2PlaneStress : public LocalResidual<T> {
 public:
  explicit HyperJ2PlaneStress(int ndims) {
    this->m_num_residuals = 4;
    this->m_num_eqs = {get_num_eqs(SYM_TENSOR, ndims), 1, 1, 1};
    this->m_var_types = {SYM_TENSOR, SCALAR, SCALAR, SCALAR};
    this->m_z_stretch_idx = 2;
  }
  void init_variables_impl() override {  // src/hyper_J2_plane_stress.cpp:117-137
    this->set_sym_tensor_xi(0, zero<T>(this->m_num_dims));
    this->set_scalar_xi(1, T(1.0));
    this->set_scalar_xi(2, T(1.0));
    this->set_scalar_xi(3, T(0.0));
  }
  int solve_nonlinear(GlobalResidual<T>& g) override {  // :176-241
    Tensor<T> const zeta_old = this->sym_tensor_xi_prev(0);
    T const Ie_old = this->scalar_xi_prev(1);
    T const lambda_z_old = this->scalar_xi_prev(2);
    T J_2D;
    Tensor<T> be_bar_trial;
    T const lambda_z = this->scalar_xi(2);  // NB: the gathered current-step value
    eval_be_bar_plane_stress(g, zeta_old, Ie_old, lambda_z_old, lambda_z, J_2D, be_bar_trial);
    T const Ie_trial = trace(be_bar_trial) / 3.;
    Tensor<T> const I = eye<T>(3);
    Tensor<T> const zeta_trial_3D = be_bar_trial - Ie_trial * I;
    Tensor<T> const zeta_trial_2D = extract_2D_tensor_from_3D(zeta_trial_3D);
    this->set_sym_tensor_xi(0, zeta_trial_2D);
    this->set_scalar_xi(1, Ie_trial);
    return this->newton(g);
  }
  int evaluate(GlobalResidual<T>& g, bool force_path, int path_in) override {  // :248-359
    int path = this->ELASTIC;
    double const sqrt_23 = std::sqrt(2. / 3.);
    double const sqrt_32 = std::sqrt(3. / 2.);
    T const E = this->m_params[0], nu = this->m_params[1], Y = this->m_params[2];
    T const S = this->m_params[3], D = this->m_params[4], A = this->m_params[5];
    T const n = this->m_params[6], K = this->m_params[7];
    T const mu = compute_mu(E, nu);
    T const kappa = compute_kappa(E, nu);
    Tensor<T> const zeta_old = this->sym_tensor_xi_prev(0);
    T const Ie_old = this->scalar_xi_prev(1);
    T const lambda_z_old = this->scalar_xi_prev(2);
    T const alpha_old = this->scalar_xi_prev(3);
    Tensor<T> const zeta = this->sym_tensor_xi(0);
    T const Ie = this->scalar_xi(1);
    T const lambda_z = this->scalar_xi(2);
    T const alpha = this->scalar_xi(3);
    Tensor<T> const I = eye<T>(3);
    T J_2D;
    Tensor<T> be_bar_trial;
    eval_be_bar_plane_stress(g, zeta_old, Ie_old, lambda_z_old, lambda_z, J_2D, be_bar_trial);
    T const Ie_trial = trace(be_bar_trial) / 3.;
    Tensor<T> const zeta_trial_3D = be_bar_trial - Ie_trial * I;
    Tensor<T> const zeta_trial_2D = extract_2D_tensor_from_3D(zeta_trial_3D);
    Tensor<T> zeta_3D = insert_2D_tensor_into_3D(zeta);
    T const zeta_zz = -trace(zeta);
    zeta_3D(2, 2) = zeta_zz;
    Tensor<T> const be_bar = zeta_3D + Ie * I;
    Tensor<T> const s = mu * zeta_3D;
    T const s_mag = norm(s);
    double const power_law_offset = 1e-12;
    T const sigma_yield = Y + S * (1. - exp(-D * alpha)) +
                          A * pow(alpha + power_law_offset, n) + K * alpha;
    T const f = (s_mag - sqrt_23 * sigma_yield) / val(mu);
    Tensor<T> R_zeta;
    T R_Ie, R_lambda_z, R_alpha;
    T const mat_factor = kappa / (2. * mu);
    R_lambda_z = lambda_z - sqrt((1. - zeta_zz / mat_factor) / pow(J_2D, 2));
    bool plastic;
    if (!force_path) plastic = (f > this->m_abs_tol || std::abs(val(f)) < this->m_abs_tol);
    else plastic = (path_in == this->PLASTIC);
    if (plastic) {
      Tensor<T> const n_2D = mu * zeta / s_mag;
      T const dgam = sqrt_32 * (alpha - alpha_old);
      R_zeta = zeta - zeta_trial_2D + 2. * dgam * Ie * n_2D;
      R_Ie = det(be_bar) - 1.;
      R_alpha = f;
      path = this->PLASTIC;
    } else {
      R_zeta = zeta - zeta_trial_2D;
      R_Ie = Ie - Ie_trial;
      R_alpha = alpha - alpha_old;
      path = this->ELASTIC;
    }
    this->set_sym_tensor_R(0, R_zeta);
    this->set_scalar_R(1, R_Ie);
    this->set_scalar_R(2, R_lambda_z);
    this->set_scalar_R(3, R_alpha);
    return path;
  }
  bool is_finite_deformation() override { return true; }
  Tensor<T> cauchy(GlobalResidual<T>& g) override {  // :361-375
    T const E = this->m_params[0], nu = this->m_params[1];
    T const mu = compute_mu(E, nu);
    T const kappa = compute_kappa(E, nu);
    Tensor<T> const I = eye<T>(g.num_dims());
    Tensor<T> const F = g.grad_vector_x(0) + I;
    T const lambda_z = this->scalar_xi(this->m_z_stretch_idx);
    T const J = det(F) * lambda_z;
    Tensor<T> const zeta = this->sym_tensor_xi(0);
    return mu * zeta / J + kappa / 2. * (J - 1. / J) * I;
  }
  Tensor<T> dev_cauchy(GlobalResidual<T>& g) override {
    T const mu = compute_mu(this->m_params[0], this->m_params[1]);
    Tensor<T> const I = eye<T>(g.num_dims());
    Tensor<T> const F = g.grad_vector_x(0) + I;
    T const lambda_z = this->scalar_xi(this->m_z_stretch_idx);
    T const J = det(F) * lambda_z;
    return mu * this->sym_tensor_xi(0) / J;
  }
  T hydro_cauchy(GlobalResidual<T>& g) override {
    T const kappa = compute_kappa(this->m_params[0], this->m_params[1]);
    Tensor<T> const I = eye<T>(g.num_dims());
    Tensor<T> const F = g.grad_vector_x(0) + I;
    T const lambda_z = this->scalar_xi(this->m_z_stretch_idx);
    T const J = det(F) * lambda_z;
    return kappa / 2. * (J - 1. / J);
  }
  T pressure_scale_factor() override { return T(0.); }
};

// ---------------------------------------------------------------------------
// src/hyper_J2_plane_strain.cpp:129-151
template <class T>
void eval_be_bar_plane_strain(GlobalResidual<T>& g, Tensor<T> const& zeta, T const& Ie,
                              Tensor<T>& be_bar) {
  int const nd = g.num_dims();
  Tensor<T> const I = eye<T>(nd);
  Tensor<T> const F = g.grad_vector_x(0) + I;
  Tensor<T> const F_prev = g.grad_vector_x_prev(0) + I;
  Tensor<T> const rF = F * inverse(F_prev);
  T const det_rF = det(rF);
  T const det_rF_13 = cbrt(det_rF);
  Tensor<T> const rF_bar = rF / det_rF_13;
  Tensor<T> const rF_barT = transpose(rF_bar);
  Tensor<T> const be_bar_2D = rF_bar * (zeta + Ie * I) * rF_barT;
  T const zeta_zz = -trace(zeta);
  T const be_bar_zz = (zeta_zz + Ie) / (det_rF_13 * det_rF_13);
  be_bar = insert_2D_tensor_into_3D(be_bar_2D);
  be_bar(2, 2) = be_bar_zz;
}

template <class T>
class HyperJ2PlaneStrain : public LocalResidual<T> {
 public:
  explicit HyperJ2PlaneStrain(int ndims) {
    this->m_num_residuals = 3;
    this->m_num_eqs = {get_num_eqs(SYM_TENSOR, ndims), 1, 1};
    this->m_var_types = {SYM_TENSOR, SCALAR, SCALAR};
  }
  void init_variables_impl() override {
    this->set_scalar_xi(1, T(1.));
    this->set_scalar_xi(2, T(0.));
    this->set_sym_tensor_xi(0, zero<T>(this->m_num_dims));
  }
  int solve_nonlinear(GlobalResidual<T>& g) override {  // src/hyper_J2_plane_strain.cpp:158-222
    Tensor<T> const zeta_old = this->sym_tensor_xi_prev(0);
    T const Ie_old = this->scalar_xi_prev(1);
    T const alpha_old = this->scalar_xi_prev(2);
    Tensor<T> be_bar_trial;
    eval_be_bar_plane_strain(g, zeta_old, Ie_old, be_bar_trial);
    T const Ie_trial = trace(be_bar_trial) / 3.;
    Tensor<T> const I = eye<T>(g.num_dims());
    Tensor<T> const zeta_trial = extract_2D_tensor_from_3D(be_bar_trial) - Ie_trial * I;
    this->set_sym_tensor_xi(0, zeta_trial);
    this->set_scalar_xi(1, Ie_trial);
    this->set_scalar_xi(2, alpha_old);
    return this->newton(g);
  }
  int evaluate(GlobalResidual<T>& g, bool force_path, int path_in) override {  // :229-318
    int path = this->ELASTIC;
    int const nd = this->m_num_dims;
    double const sqrt_23 = std::sqrt(2. / 3.);
    double const sqrt_32 = std::sqrt(3. / 2.);
    T const E = this->m_params[0], nu = this->m_params[1], K = this->m_params[2];
    T const Y = this->m_params[3], Y_inf = this->m_params[4], delta = this->m_params[5];
    T const mu = compute_mu(E, nu);
    Tensor<T> const zeta_old = this->sym_tensor_xi_prev(0);
    T const Ie_old = this->scalar_xi_prev(1);
    T const alpha_old = this->scalar_xi_prev(2);
    Tensor<T> const zeta = this->sym_tensor_xi(0);
    T const Ie = this->scalar_xi(1);
    T const alpha = this->scalar_xi(2);
    Tensor<T> const I = eye<T>(nd);
    Tensor<T> be_bar_trial;
    eval_be_bar_plane_strain(g, zeta_old, Ie_old, be_bar_trial);
    T const Ie_trial = trace(be_bar_trial) / 3.;
    Tensor<T> const zeta_trial = extract_2D_tensor_from_3D(be_bar_trial) - Ie_trial * I;
    Tensor<T> zeta_3D = insert_2D_tensor_into_3D(zeta);
    zeta_3D(2, 2) = -trace(zeta);
    Tensor<T> const I_3D = eye<T>(3);
    Tensor<T> be_bar_3D = zeta_3D + Ie * I_3D;
    Tensor<T> const s_3D = mu * zeta_3D;
    T const s_mag = norm(s_3D);
    T const sigma_yield = Y + K * alpha + (Y_inf - Y) * (1. - exp(-delta * alpha));
    T const f = (s_mag - sqrt_23 * sigma_yield) / val(mu);
    Tensor<T> R_zeta;
    T R_Ie, R_alpha;
    bool plastic;
    if (!force_path) plastic = (f > this->m_abs_tol || std::abs(val(f)) < this->m_abs_tol);
    else plastic = (path_in == this->PLASTIC);
    if (plastic) {
      Tensor<T> const n_2D = mu * zeta / s_mag;
      T const dgam = sqrt_32 * (alpha - alpha_old);
      R_zeta = zeta - zeta_trial + 2. * dgam * Ie * n_2D;
      R_Ie = det(be_bar_3D) - 1.;
      R_alpha = f;
      path = this->PLASTIC;
    } else {
      R_zeta = zeta - zeta_trial;
      R_Ie = Ie - Ie_trial;
      R_alpha = alpha - alpha_old;
      path = this->ELASTIC;
    }
    this->set_sym_tensor_R(0, R_zeta);
    this->set_scalar_R(1, R_Ie);
    this->set_scalar_R(2, R_alpha);
    return path;
  }
  bool is_finite_deformation() override { return true; }
  Tensor<T> cauchy(GlobalResidual<T>& g) override {
    T const p = g.scalar_x(1);
    Tensor<T> const I = eye<T>(g.num_dims());
    return this->dev_cauchy(g) - p * I;
  }
  Tensor<T> dev_cauchy(GlobalResidual<T>& g) override {
    T const mu = compute_mu(this->m_params[0], this->m_params[1]);
    Tensor<T> const I = eye<T>(g.num_dims());
    Tensor<T> const F = g.grad_vector_x(0) + I;
    T const J = det(F);
    return mu * this->sym_tensor_xi(0) / J;
  }
  T hydro_cauchy(GlobalResidual<T>& g) override {
    T const kappa = compute_kappa(this->m_params[0], this->m_params[1]);
    Tensor<T> const I = eye<T>(g.num_dims());
    Tensor<T> const F = g.grad_vector_x(0) + I;
    T const J = det(F);
    return kappa / 2. * (J - 1. / J);
  }
  T pressure_scale_factor() override {
    return compute_kappa(this->m_params[0], this->m_params[1]);
  }
};

// ---------------------------------------------------------------------------
// src/hypo_kinematics.hpp:10-17
template <class T>
Tensor<T> compute_unrotated_rate_of_deformation(Tensor<T> const& F, Tensor<T> const& F_prev,
                                                Tensor<T> const& R) {
  Tensor<T> const Finv = inverse(F);
  Tensor<T> const L = (F - F_prev) * Finv;
  Tensor<T> const D = 0.5 * (L + transpose(L));
  return transpose(R) * D * R;
}

// src/hypo_hill.cpp (3-D): unrotated Cauchy stress TC (sym) + alpha; Hill-48 yield on TC
template <class T>
class HypoHill : public LocalResidual<T> {
 public:
  explicit HypoHill(int ndims) {
    this->m_num_residuals = 2;
    this->m_num_eqs = {get_num_eqs(SYM_TENSOR, ndims), 1};
    this->m_var_types = {SYM_TENSOR, SCALAR};
  }
  void init_variables_impl() override {   // :127-139
    this->set_sym_tensor_xi(0, zero<T>(this->m_num_dims));
    this->set_scalar_xi(1, T(0.));
  }
  Tensor<T> eval_d(GlobalResidual<T>& g) {   // :141-147
    if (m_kinematics_cached) return m_d;
    return compute_unrotated_rate_of_deformation(g.F(), g.F_prev(), g.R());
  }
  int solve_nonlinear(GlobalResidual<T>& g) override {   // :155-226
    m_d = eval_d(g);
    m_kinematics_cached = true;
    {
      double const E = val(this->m_params[0]), nu = val(this->m_params[1]);
      double const lambda = compute_lambda(E, nu), mu = compute_mu(E, nu);
      Tensor<T> const I = eye<T>(this->m_num_dims);
      Tensor<T> const TC_old = this->sym_tensor_xi_prev(0);
      T const alpha_old = this->scalar_xi_prev(1);
      Tensor<T> const d = eval_d(g);
      Tensor<T> const TC = TC_old + (lambda * trace(d)) * I + (2. * mu) * d;
      this->set_sym_tensor_xi(0, TC);
      this->set_scalar_xi(1, alpha_old);
    }
    int const path = this->newton(g);
    m_kinematics_cached = false;
    return path;
  }
  int evaluate(GlobalResidual<T>& g, bool force_path, int path_in) override {   // :233-310
    int path = this->ELASTIC;
    T const E = this->m_params[0], nu = this->m_params[1], Y = this->m_params[2];
    T const R00 = this->m_params[3], R11 = this->m_params[4], R22 = this->m_params[5];
    T const R01 = this->m_params[6], R02 = this->m_params[7], R12 = this->m_params[8];
    T const S = this->m_params[9], D = this->m_params[10];
    T const lambda = compute_lambda(E, nu), mu = compute_mu(E, nu);
    HillParams<T> const hp = compute_hill_params(R00, R11, R22, R01, R02, R12);
    Tensor<T> const TC_old = this->sym_tensor_xi_prev(0);
    T const alpha_old = this->scalar_xi_prev(1);
    Tensor<T> const TC = this->sym_tensor_xi(0);
    T const alpha = this->scalar_xi(1);
    T const hill = compute_hill_value(TC, hp);
    T const sigma_yield = Y + S * (1. - exp(-D * alpha));
    T const f = (hill - sigma_yield) / val(mu);
    Tensor<T> const I = eye<T>(this->m_num_dims);
    Tensor<T> const d = eval_d(g);
    Tensor<T> R_TC = TC - TC_old - (lambda * trace(d)) * I - (2. * mu) * d;
    R_TC = R_TC / val(mu);
    T R_alpha;
    bool plastic;
    if (!force_path) plastic = (f > this->m_abs_tol || std::abs(val(f)) < this->m_abs_tol);
    else plastic = (path_in == this->PLASTIC);
    if (plastic) {
      T const dgam = alpha - alpha_old;
      Tensor<T> const n = compute_hill_normal(TC, hp, hill);
      R_TC = R_TC + ((2. * mu * dgam) * n) / val(mu);
      R_alpha = f;
      path = this->PLASTIC;
    } else {
      R_alpha = alpha - alpha_old;
      path = this->ELASTIC;
    }
    this->set_sym_tensor_R(0, R_TC);
    this->set_scalar_R(1, R_alpha);
    return path;
  }
  bool is_finite_deformation() override { return true; }
  Tensor<T> rotated_cauchy(GlobalResidual<T>& g) {   // :312-318
    Tensor<T> const TC = this->sym_tensor_xi(0);
    Tensor<T> const R = g.R();
    return R * TC * transpose(R);
  }
  Tensor<T> cauchy(GlobalResidual<T>& g) override {   // :320-329
    T const p = g.scalar_x(1);
    Tensor<T> const I = eye<T>(this->m_num_dims);
    return this->dev_cauchy(g) - p * I;
  }
  Tensor<T> dev_cauchy(GlobalResidual<T>& g) override { return dev(rotated_cauchy(g)); }   // :331-335
  T hydro_cauchy(GlobalResidual<T>& g) override { return trace(rotated_cauchy(g)) / 3.; }  // :337-340
  T pressure_scale_factor() override { return compute_kappa(this->m_params[0], this->m_params[1]); }

 private:
  Tensor<T> m_d;
  bool m_kinematics_cached = false;
};

// src/hypo_hill_plane_strain.cpp:135-384: xi = TC (2-D sym), alpha, TC_zz; mixed u-p mechanics
template <class T>
Tensor<T> hypo_eval_d_2D(GlobalResidual<T>& g) {   // src/hypo_hill_plane_strain.cpp:137-151
  Tensor<T> const I = eye<T>(g.num_dims());
  Tensor<T> const F = g.grad_vector_x(0) + I;
  Tensor<T> const F_prev = g.grad_vector_x_prev(0) + I;
  Tensor<T> const Finv = inverse(F);
  Tensor<T> const R = polar_rotation(F);
  Tensor<T> const L = (F - F_prev) * Finv;
  Tensor<T> const D = 0.5 * (L + transpose(L));
  return transpose(R) * D * R;
}

template <class T>
class HypoHillPlaneStrain : public LocalResidual<T> {
 public:
  explicit HypoHillPlaneStrain(int ndims) {
    this->m_num_residuals = 3;
    this->m_num_eqs = {get_num_eqs(SYM_TENSOR, ndims), 1, 1};
    this->m_var_types = {SYM_TENSOR, SCALAR, SCALAR};
  }
  void init_variables_impl() override {
    this->set_sym_tensor_xi(0, zero<T>(this->m_num_dims));
    this->set_scalar_xi(1, T(0.));
    this->set_scalar_xi(2, T(0.));
  }
  int solve_nonlinear(GlobalResidual<T>& g) override {   // :158-218
    double const E = val(this->m_params[0]), nu = val(this->m_params[1]);
    double const lambda = compute_lambda(E, nu), mu = compute_mu(E, nu);
    Tensor<T> const I = eye<T>(this->m_num_dims);
    Tensor<T> const TC_old = this->sym_tensor_xi_prev(0);
    T const alpha_old = this->scalar_xi_prev(1);
    T const TC_zz_old = this->scalar_xi_prev(2);
    Tensor<T> const d = hypo_eval_d_2D(g);
    Tensor<T> const TC = TC_old + (lambda * trace(d)) * I + (2. * mu) * d;
    T const TC_zz = TC_zz_old + lambda * trace(d);
    this->set_sym_tensor_xi(0, TC);
    this->set_scalar_xi(1, alpha_old);
    this->set_scalar_xi(2, TC_zz);
    return this->newton(g);
  }
  int evaluate(GlobalResidual<T>& g, bool force_path, int path_in) override {   // :225-325
    int path = this->ELASTIC;
    T const E = this->m_params[0], nu = this->m_params[1], Y = this->m_params[2];
    T const S = this->m_params[3], D = this->m_params[4];
    T const R00 = this->m_params[5], R11 = this->m_params[6], R22 = this->m_params[7], R01 = this->m_params[8];
    T const lambda = compute_lambda(E, nu), mu = compute_mu(E, nu);
    Tensor<T> const TC_old = this->sym_tensor_xi_prev(0);
    T const alpha_old = this->scalar_xi_prev(1);
    T const TC_zz_old = this->scalar_xi_prev(2);
    Tensor<T> const TC = this->sym_tensor_xi(0);
    T const alpha = this->scalar_xi(1);
    T const TC_zz = this->scalar_xi(2);
    Tensor<T> TC_3D = insert_2D_tensor_into_3D(TC);
    TC_3D(2, 2) = TC_zz;
    T const R02 = T(1.), R12 = T(1.);
    HillParams<T> const hp = compute_hill_params(R00, R11, R22, R01, R02, R12);
    T const phi = compute_hill_value(TC_3D, hp);
    T const sigma_yield = Y + S * (1. - exp(-D * alpha));
    T const f = (phi - sigma_yield) / val(mu);
    Tensor<T> const I = eye<T>(this->m_num_dims);
    Tensor<T> const d = hypo_eval_d_2D(g);
    Tensor<T> R_TC = TC - TC_old - (lambda * trace(d)) * I - (2. * mu) * d;
    T R_TC_zz = TC_zz - TC_zz_old - lambda * trace(d);
    T R_alpha;
    bool plastic;
    if (!force_path) plastic = (f > this->m_abs_tol || std::abs(val(f)) < this->m_abs_tol);
    else plastic = (path_in == this->PLASTIC);
    if (plastic) {
      Tensor<T> const n_3D = compute_hill_normal(TC_3D, hp, phi);
      Tensor<T> const n_2D = extract_2D_tensor_from_3D(n_3D);
      T const dgam = alpha - alpha_old;
      Tensor<T> const dp_2D = dgam * n_2D;
      T const dp_zz = -trace(dp_2D);
      R_TC = R_TC + (2. * mu) * dp_2D;
      R_alpha = f;
      R_TC_zz = R_TC_zz + 2. * mu * dp_zz;
      path = this->PLASTIC;
    } else {
      R_alpha = alpha - alpha_old;
      path = this->ELASTIC;
    }
    this->set_sym_tensor_R(0, R_TC);
    this->set_scalar_R(1, R_alpha);
    this->set_scalar_R(2, R_TC_zz);
    return path;
  }
  bool is_finite_deformation() override { return true; }
  Tensor<T> rotated_cauchy(GlobalResidual<T>& g) {   // :327-337
    Tensor<T> const I = eye<T>(this->m_num_dims);
    Tensor<T> const F = g.grad_vector_x(0) + I;
    Tensor<T> const TC = this->sym_tensor_xi(0);
    Tensor<T> const R = polar_rotation(F);
    return R * TC * transpose(R);
  }
  Tensor<T> cauchy(GlobalResidual<T>& g) override {   // :339-349
    T const p = g.scalar_x(1);
    Tensor<T> const I = eye<T>(this->m_num_dims);
    return this->dev_cauchy(g) - p * I;
  }
  Tensor<T> dev_cauchy(GlobalResidual<T>& g) override {   // :351-359
    Tensor<T> const RC = rotated_cauchy(g);
    Tensor<T> const I = eye<T>(this->m_num_dims);
    return RC - this->hydro_cauchy(g) * I;
  }
  T hydro_cauchy(GlobalResidual<T>& g) override {   // :361-366
    Tensor<T> const RC = rotated_cauchy(g);
    return (trace(RC) + this->scalar_xi(2)) / 3.;
  }
  T pressure_scale_factor() override { return compute_kappa(this->m_params[0], this->m_params[1]); }
};

// src/hypo_hill_plane_stress.cpp:138-408: xi = TC (2-D sym), alpha, lambda_z; single-field plane-stress
// mechanics; material frame Q (params 9-12).  NB the unforced plastic branch alone divides R_TC by val(mu)
// (:321); the elastic branch leaves it unscaled -- restated as written.
template <class T>
class HypoHillPlaneStress : public LocalResidual<T> {
 public:
  explicit HypoHillPlaneStress(int ndims) {
    this->m_num_residuals = 3;
    this->m_num_eqs = {get_num_eqs(SYM_TENSOR, ndims), 1, 1};
    this->m_var_types = {SYM_TENSOR, SCALAR, SCALAR};
    this->m_z_stretch_idx = 2;
  }
  void init_variables_impl() override {
    this->set_sym_tensor_xi(0, zero<T>(this->m_num_dims));
    this->set_scalar_xi(1, T(0.));
    this->set_scalar_xi(2, T(1.));
  }
  Tensor<T> compute_Q() {   // :155-163
    Tensor<T> Q = zero<T>(this->m_num_dims);
    Q(0, 0) = this->m_params[9]; Q(0, 1) = this->m_params[10];
    Q(1, 0) = this->m_params[11]; Q(1, 1) = this->m_params[12];
    return Q;
  }
  Tensor<T> eval_d(GlobalResidual<T>& g, Tensor<T> const& Q) {   // :165-179
    Tensor<T> const I = eye<T>(g.num_dims());
    Tensor<T> const F = g.grad_vector_x(0) + I;
    Tensor<T> const F_prev = g.grad_vector_x_prev(0) + I;
    Tensor<T> const Finv = inverse(F);
    Tensor<T> const R = polar_rotation(F);
    Tensor<T> const L = (F - F_prev) * Finv;
    Tensor<T> const D = 0.5 * (L + transpose(L));
    return transpose(Q) * transpose(R) * D * R * Q;
  }
  int solve_nonlinear(GlobalResidual<T>& g) override {   // :186-250
    double const E = val(this->m_params[0]), nu = val(this->m_params[1]);
    double const lambda = compute_lambda(E, nu), mu = compute_mu(E, nu);
    Tensor<T> const I = eye<T>(this->m_num_dims);
    Tensor<T> const TC_old = this->sym_tensor_xi_prev(0);
    T const alpha_old = this->scalar_xi_prev(1);
    T const lambda_z_old = this->scalar_xi_prev(2);
    Tensor<T> const Q = compute_Q();
    Tensor<T> const d = eval_d(g, Q);
    T const d_zz = -lambda * trace(d) / (lambda + 2. * mu);
    Tensor<T> const TC = TC_old + (lambda * (trace(d) + d_zz)) * I + (2. * mu) * d;
    T const lambda_z = lambda_z_old / (1. - d_zz);
    this->set_sym_tensor_xi(0, TC);
    this->set_scalar_xi(1, alpha_old);
    this->set_scalar_xi(2, lambda_z);
    return this->newton(g);
  }
  int evaluate(GlobalResidual<T>& g, bool force_path, int path_in) override {   // :257-364
    int path = this->ELASTIC;
    T const E = this->m_params[0], nu = this->m_params[1], Y = this->m_params[2];
    T const S = this->m_params[3], D = this->m_params[4];
    T const R00 = this->m_params[5], R11 = this->m_params[6], R22 = this->m_params[7], R01 = this->m_params[8];
    T const lambda = compute_lambda(E, nu), mu = compute_mu(E, nu);
    Tensor<T> const TC_old = this->sym_tensor_xi_prev(0);
    T const alpha_old = this->scalar_xi_prev(1);
    T const lambda_z_old = this->scalar_xi_prev(2);
    Tensor<T> const TC = this->sym_tensor_xi(0);
    T const alpha = this->scalar_xi(1);
    T const lambda_z = this->scalar_xi(2);
    Tensor<T> TC_3D = insert_2D_tensor_into_3D(TC);
    T const R02 = T(1.), R12 = T(1.);
    HillParams<T> const hp = compute_hill_params(R00, R11, R22, R01, R02, R12);
    T const phi = compute_hill_value(TC_3D, hp);
    T const sigma_yield = Y + S * (1. - exp(-D * alpha));
    T const f = (phi - sigma_yield) / val(mu);
    Tensor<T> const I = eye<T>(this->m_num_dims);
    Tensor<T> const Q = compute_Q();
    Tensor<T> const d = eval_d(g, Q);
    T const d_zz = -lambda * trace(d) / (lambda + 2. * mu);
    Tensor<T> R_TC = TC - TC_old - (lambda * (trace(d) + d_zz)) * I - (2. * mu) * d;
    T R_alpha, R_lambda_z;
    bool plastic;
    if (!force_path) plastic = (f > this->m_abs_tol || std::abs(val(f)) < this->m_abs_tol);
    else plastic = (path_in == this->PLASTIC);
    if (plastic) {
      Tensor<T> const n_3D = compute_hill_normal(TC_3D, hp, phi);
      Tensor<T> const n_2D = extract_2D_tensor_from_3D(n_3D);
      T const dgam = alpha - alpha_old;
      Tensor<T> const dp_2D = dgam * n_2D;
      T const dp_zz = -trace(dp_2D);
      T const corr_dp_zz = 2. * mu * dp_zz / (2. * mu + lambda);
      R_TC(0, 0) += 2. * mu * dp_2D(0, 0) - lambda * corr_dp_zz;
      R_TC(1, 1) += 2. * mu * dp_2D(1, 1) - lambda * corr_dp_zz;
      R_TC(0, 1) += 2. * mu * dp_2D(0, 1);
      if (!force_path) R_TC = R_TC / val(mu);   // :321 (the forced branch, :336-347, has no such line)
      R_alpha = f;
      R_lambda_z = lambda_z - lambda_z_old / (1. - (d_zz + corr_dp_zz));
      path = this->PLASTIC;
    } else {
      R_alpha = alpha - alpha_old;
      R_lambda_z = lambda_z - lambda_z_old / (1. - d_zz);
      path = this->ELASTIC;
    }
    this->set_sym_tensor_R(0, R_TC);
    this->set_scalar_R(1, R_alpha);
    this->set_scalar_R(2, R_lambda_z);
    return path;
  }
  bool is_finite_deformation() override { return true; }
  Tensor<T> rotated_cauchy(GlobalResidual<T>& g) {   // :366-377
    Tensor<T> const Q = compute_Q();
    Tensor<T> const I = eye<T>(this->m_num_dims);
    Tensor<T> const F = g.grad_vector_x(0) + I;
    Tensor<T> const TC = this->sym_tensor_xi(0);
    Tensor<T> const R = polar_rotation(F);
    return R * Q * TC * transpose(Q) * transpose(R);
  }
  Tensor<T> cauchy(GlobalResidual<T>& g) override { return rotated_cauchy(g); }   // :379-382
  Tensor<T> dev_cauchy(GlobalResidual<T>& g) override {   // :384-389
    Tensor<T> const I = eye<T>(this->m_num_dims);
    return rotated_cauchy(g) - this->hydro_cauchy(g) * I;
  }
  T hydro_cauchy(GlobalResidual<T>& g) override { return trace(rotated_cauchy(g)) / 3.; }   // :391-394
  T pressure_scale_factor() override { return T(0.); }
};

// ---- factories, src/local_residual.cpp:892-933, src/global_residual.cpp:619-630 ----
enum LocalType {
  L_ELASTIC = 0, L_SMALL_J2 = 1, L_SMALL_HILL = 2, L_SMALL_HILL_PLANE_STRESS = 3,
  L_HYPER_J2 = 4, L_HYPER_J2_PLANE_STRESS = 5, L_SMALL_HILL_PLANE_STRAIN = 6,
  L_HYPER_J2_PLANE_STRAIN = 7, L_HYPO_HILL = 8,
  L_HYPO_HILL_PLANE_STRAIN = 9, L_HYPO_HILL_PLANE_STRESS = 10
};
enum GlobalType { G_MECHANICS = 0, G_MECHANICS_PLANE_STRESS = 1 };

inline int local_num_params(int type) {
  switch (type) {
    case L_ELASTIC: return 4;
    case L_SMALL_J2: return 6;
    case L_SMALL_HILL: return 11;
    case L_SMALL_HILL_PLANE_STRESS: return 9;
    case L_SMALL_HILL_PLANE_STRAIN: return 9;
    case L_HYPER_J2: return 8;
    case L_HYPER_J2_PLANE_STRESS: return 8;
    case L_HYPER_J2_PLANE_STRAIN: return 6;
    case L_HYPO_HILL: return 11;
    case L_HYPO_HILL_PLANE_STRAIN: return 9;
    case L_HYPO_HILL_PLANE_STRESS: return 13;
  }
  return -1;
}

template <class T>
std::unique_ptr<LocalResidual<T>> create_local_residual(int type, int ndims) {
  switch (type) {
    case L_ELASTIC: return std::make_unique<Elastic<T>>(ndims);
    case L_SMALL_J2: return std::make_unique<SmallJ2<T>>(ndims);
    case L_SMALL_HILL: return std::make_unique<SmallHill<T>>(ndims);
    case L_SMALL_HILL_PLANE_STRESS: return std::make_unique<SmallHillPlaneStress<T>>(ndims);
    case L_SMALL_HILL_PLANE_STRAIN: return std::make_unique<SmallHillPlaneStrain<T>>(ndims);
    case L_HYPER_J2: return std::make_unique<HyperJ2<T>>(ndims);
    case L_HYPER_J2_PLANE_STRESS: return std::make_unique<HyperJ2PlaneStress<T>>(ndims);
    case L_HYPER_J2_PLANE_STRAIN: return std::make_unique<HyperJ2PlaneStrain<T>>(ndims);
    case L_HYPO_HILL: return std::make_unique<HypoHill<T>>(ndims);
    case L_HYPO_HILL_PLANE_STRAIN: return std::make_unique<HypoHillPlaneStrain<T>>(ndims);
    case L_HYPO_HILL_PLANE_STRESS: return std::make_unique<HypoHillPlaneStress<T>>(ndims);
  }
  throw std::runtime_error("unknown local residual type");
}

}  // namespace orc
