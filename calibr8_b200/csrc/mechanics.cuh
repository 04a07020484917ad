// Element geometry, interpolation and the mechanics global residuals
// (GlobalResidual<T> of the reference) as device templates.
//
//   MECH_MIXED         Mechanics<T>, mixed u-p, src/mechanics.cpp:61-240
//   MECH_PLANE_STRESS  MechanicsPlaneStress<T>, src/mechanics_plane_stress.cpp:46-95
//
// Element dofs are node-major interleaved here: dof(n, eq) = n*NB + eq with
// eq < DIM the displacement components and eq == DIM the pressure (mixed).
// (The reference orders residual-major, src/global_residual.cpp:21-23; the
// C-ABI export functions translate.)  Linear simplices only: basis gradients
// and dv are constant per element; the coupled point is the centroid.
#pragma once
#include "models.cuh"

namespace c8 {

enum MechType { MECH_MIXED = 0, MECH_PLANE_STRESS = 1 };

template <int DIM, int MECH>
struct MechTraits {
  static constexpr int NN = DIM + 1;
  static constexpr int NB = DIM + (MECH == MECH_MIXED ? 1 : 0);
  static constexpr int NX = NN * NB;
};

template <int DIM>
struct Geom {
  double gN[DIM + 1][DIM];  // dN_n/dX_j
  double dv;                // |det J| (tet) / 2*area (tri), apf::getDV
  double h;                 // sqrt(mean squared edge length), src/mechanics.cpp:103-113
};

template <int DIM>
C8_DI void geom_from_coords(const double (&X)[DIM + 1][DIM], Geom<DIM>& g);

template <int DIM>
C8_DI void load_geom(const double* __restrict__ coords, const int* nodes, Geom<DIM>& g) {
  double X[DIM + 1][DIM];
#pragma unroll
  for (int n = 0; n <= DIM; ++n)
#pragma unroll
    for (int k = 0; k < DIM; ++k) X[n][k] = __ldg(&coords[size_t(nodes[n]) * DIM + k]);
  geom_from_coords<DIM>(X, g);
}

template <int DIM>
C8_DI void geom_from_coords(const double (&X)[DIM + 1][DIM], Geom<DIM>& g) {
  Mat<double, DIM> J;  // J(a,k) = dx_k/dxi_a
#pragma unroll
  for (int a = 0; a < DIM; ++a)
#pragma unroll
    for (int k = 0; k < DIM; ++k) J(a, k) = X[a + 1][k] - X[0][k];
  const double d = det(J);
  g.dv = (DIM == 3) ? d : fabs(d);
  const Mat<double, DIM> Ji = inverse(J);
#pragma unroll
  for (int k = 0; k < DIM; ++k) g.gN[0][k] = 0.0;
#pragma unroll
  for (int a = 0; a < DIM; ++a)
#pragma unroll
    for (int k = 0; k < DIM; ++k) {
      g.gN[a + 1][k] = Ji(k, a);
      g.gN[0][k] -= Ji(k, a);
    }
  double s = 0.0;
#pragma unroll
  for (int a = 0; a <= DIM; ++a)
#pragma unroll
    for (int b = a + 1; b <= DIM; ++b)
#pragma unroll
      for (int k = 0; k < DIM; ++k) s += (X[a][k] - X[b][k]) * (X[a][k] - X[b][k]);
  g.h = sqrt(s / double((DIM + 1) * DIM / 2));
}

// value-only interpolation of grad u at the (constant-gradient) element
template <int DIM, int NB>
C8_DI Mat<double, DIM> grad_u_val(const double (&xn)[DIM + 1][NB], const Geom<DIM>& g) {
  Mat<double, DIM> r;
#pragma unroll
  for (int i = 0; i < DIM; ++i)
#pragma unroll
    for (int j = 0; j < DIM; ++j) {
      double s = xn[0][i] * g.gN[0][j];
#pragma unroll
      for (int n = 1; n <= DIM; ++n) s += xn[n][i] * g.gN[n][j];
      r(i, j) = s;
    }
  return r;
}

// Per-thread description of the derivative lanes it owns when the element
// dofs are the independent variables (seed_wrt_x / seed_wrt_x_prev).
template <int DIM, int NB, int L>
struct XLanes {
  double gsel[L][DIM];  // dN_{n_c}/dX_j of the lane's node (0 for padding lanes)
  double nsel[L];       // 1 for a real lane, 0 for padding
  int eq[L];            // equation index of the lane's dof (NB for padding -> matches nothing)
  int node[L];
  C8_DI void init(int base, const Geom<DIM>& g) {
#pragma unroll
    for (int s = 0; s < L; ++s) {
      const int c = base + s;
      const bool real = c < (DIM + 1) * NB;
      node[s] = c / NB;
      eq[s] = real ? c % NB : NB + 1;
      nsel[s] = real ? 1.0 : 0.0;
#pragma unroll
      for (int j = 0; j < DIM; ++j) {
        double v = 0.0;
#pragma unroll
        for (int n = 0; n <= DIM; ++n) v = pick(real && node[s] == n, g.gN[n][j], v);
        gsel[s][j] = v;
      }
    }
  }
};

// grad u as Dual<L> seeded w.r.t. the element dofs
template <int DIM, int NB, int L>
C8_DI Mat<Dual<L>, DIM> grad_u_seeded(const Mat<double, DIM>& gu, const XLanes<DIM, NB, L>& xl) {
  Mat<Dual<L>, DIM> r;
#pragma unroll
  for (int i = 0; i < DIM; ++i)
#pragma unroll
    for (int j = 0; j < DIM; ++j) {
      r(i, j).v = gu(i, j);
#pragma unroll
      for (int s = 0; s < L; ++s) r(i, j).d[s] = (xl.eq[s] == i) ? xl.gsel[s][j] : 0.0;
    }
  return r;
}

// -----------------------------------------------------------------------------
// Stress measure P with R_u[n,i] = sum_j P_ij dN_n/dX_j w dv  (both mechanics types)
//   mixed:         P = (dev sigma - p I) [* cof F if finite]         src/mechanics.cpp:115-145
//   plane stress:  P = sigma [* lambda_z J F^-T if finite] * thickness  src/mechanics_plane_stress.cpp:60-93
template <int DIM, int MECH, class Model, class TK, class TKP, class TX, class TP>
C8_DI Mat<prom3_t<TK, TX, TP>, DIM> first_pk(const Kin<DIM, TK, TKP>& k, const TK& p, const TX* xi,
                                             const TP* par, double thickness) {
  using R = prom3_t<TK, TX, TP>;
  if constexpr (MECH == MECH_MIXED) {
    Mat<R, DIM> sig = Model::dev_cauchy(k, xi, par);
#pragma unroll
    for (int i = 0; i < DIM; ++i) sig(i, i) = sig(i, i) - p;
    if constexpr (Model::FINITE) {
      const Mat<TK, DIM> F = add_diag(k.gu, 1.0);
      Mat<TK, DIM> cof;
      if constexpr (DIM == 3) cof = transpose(cofactor_T(F));
      else { cof(0, 0) = F(1, 1); cof(0, 1) = -F(1, 0); cof(1, 0) = -F(0, 1); cof(1, 1) = F(0, 0); }
      return mat_conv<R>(sig * cof);
    } else {
      return sig;
    }
  } else {
    Mat<R, DIM> sig = Model::cauchy(k, xi, par);
    if constexpr (Model::FINITE) {
      const Mat<TK, DIM> F = add_diag(k.gu, 1.0);
      const Mat<TK, DIM> FiT = transpose(inverse(F));
      const TK J = det(F);
      const auto zJ = xi[Model::Z_STRETCH] * J;
      return mat_conv<R>(scale(thickness, scale(zJ, sig) * FiT));
    } else {
      return mat_conv<R>(scale(thickness, sig));
    }
  }
}

// Pressure-row integrands of the mixed formulation at the coupled point,
// src/mechanics.cpp:147-204:  R_p[n] -= hp N_n w dv + sum_i sv_i dN_n/dX_i w dv
//   hp = hydro / pressure_scale ;  sv = S grad p, S = tau I or tau cofF^T cofF / det F,
//   tau = c_stab h^2 / (2 mu)
template <int DIM, class Model, class TK, class TKP, class TX, class TP>
C8_DI void pressure_terms(const Kin<DIM, TK, TKP>& k, const TK (&grad_p)[DIM], const TX* xi,
                          const TP* par, double h, double stab_mult, prom3_t<TK, TX, TP>& hp,
                          prom3_t<TK, TX, TP> (&sv)[DIM]) {
  using R = prom3_t<TK, TX, TP>;
  hp = conv<R>(Model::hydro(k, xi, par) / Model::pscale(par));
  const TP mu = mu_of(par[0], par[1]);
  const TP tau = stab_mult * 0.5 * h * h / mu;
  if constexpr (Model::FINITE) {
    const Mat<TK, DIM> F = add_diag(k.gu, 1.0);
    Mat<TK, DIM> cof;
    if constexpr (DIM == 3) cof = transpose(cofactor_T(F));
    else { cof(0, 0) = F(1, 1); cof(0, 1) = -F(1, 0); cof(1, 0) = -F(0, 1); cof(1, 1) = F(0, 0); }
    const TK dF = det(F);
    const auto S = scale(tau / dF, transpose(cof) * cof);
#pragma unroll
    for (int i = 0; i < DIM; ++i) {
      R s = conv<R>(S(i, 0) * grad_p[0]);
#pragma unroll
      for (int j = 1; j < DIM; ++j) s += S(i, j) * grad_p[j];
      sv[i] = s;
    }
  } else {
#pragma unroll
    for (int i = 0; i < DIM; ++i) sv[i] = conv<R>(tau * grad_p[i]);
  }
}

// quadrature of the pressure-mass ip set (order 2), apf rules
template <int DIM> struct Quad2;
template <> struct Quad2<3> {
  static constexpr int NPT = 4;
  static C8_DI void basis(int q, double* N) {
    const double a = 0.138196601125011, b = 0.585410196624969;
    // points (a,a,a),(b,a,a),(a,b,a),(a,a,b): N = (1-sum, xi0, xi1, xi2)
    N[0] = (q == 0) ? 1.0 - 3.0 * a : 1.0 - 2.0 * a - b;
    N[1] = (q == 1) ? b : a;
    N[2] = (q == 2) ? b : a;
    N[3] = (q == 3) ? b : a;
  }
  static C8_DI double weight() { return 1.0 / 24.0; }
};
template <> struct Quad2<2> {
  static constexpr int NPT = 3;
  static C8_DI void basis(int q, double* N) {
    const double a = 1.0 / 6.0, b = 2.0 / 3.0;
    // points (b,a),(a,b),(a,a)
    N[1] = (q == 0) ? b : a;
    N[2] = (q == 1) ? b : a;
    N[0] = 1.0 - N[1] - N[2];
  }
  static C8_DI double weight() { return 1.0 / 6.0; }
};
template <int DIM> C8_DI double quad1_weight() { return DIM == 3 ? 1.0 / 6.0 : 0.5; }

}  // namespace c8
