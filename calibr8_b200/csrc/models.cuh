// Constitutive models (LocalResidual<T> of the reference) as device-side static
// templates.  Every function is generic in the scalar type of each input group
//   TK  current kinematics (grad u)      TKP previous-step kinematics (grad u_prev)
//   TX  local state xi                   TXP previous local state xi_prev
//   TP  material parameters
// so that each AD pass only carries derivative lanes where it is seeded.
//
// Reference semantics restated (file:line of the reference):
//   Elastic               src/elastic.cpp:76-139
//   SmallJ2               src/small_J2.cpp:121-297
//   SmallHill             src/small_hill.cpp:137-324, src/yield_functions.hpp:34-99
//   SmallHillPlaneStress  src/small_hill_plane_stress.cpp:135-327
//   SmallHillPlaneStrain  src/small_hill_plane_strain.cpp:135-331
//   HyperJ2               src/hyper_J2.cpp:136-360
//   HyperJ2PlaneStress    src/hyper_J2_plane_stress.cpp:140-411
//   HyperJ2PlaneStrain    src/hyper_J2_plane_strain.cpp:129-372
//   HypoHill              src/hypo_hill.cpp:141-346, src/hypo_kinematics.hpp:10-17
//   HypoHillPlaneStrain   src/hypo_hill_plane_strain.cpp:137-379
//   HypoHillPlaneStress   src/hypo_hill_plane_stress.cpp:155-403
// Packed xi order = the reference's residual order (sym tensor first, then scalars).
#pragma once
#include "tensor.cuh"

namespace c8 {

enum LocalType {
  L_ELASTIC = 0, L_SMALL_J2 = 1, L_SMALL_HILL = 2, L_SMALL_HILL_PLANE_STRESS = 3,
  L_HYPER_J2 = 4, L_HYPER_J2_PLANE_STRESS = 5, L_SMALL_HILL_PLANE_STRAIN = 6,
  L_HYPER_J2_PLANE_STRAIN = 7, L_HYPO_HILL = 8, L_HYPO_HILL_PLANE_STRAIN = 9, L_HYPO_HILL_PLANE_STRESS = 10
};
enum { PATH_ELASTIC = 0, PATH_PLASTIC = 1 };

template <int DIM, class TK, class TKP>
struct Kin {
  Mat<TK, DIM> gu;    // grad u at the point
  Mat<TKP, DIM> gup;  // grad u_prev at the point
  // Optional cache of R = polar_rotation(I + grad u) for the hypoelastic models (the reference caches it per
  // interpolation, src/global_residual.hpp:302-305).  has_rot is a compile-time constant at every use (local
  // objects, everything inlined), so kernels that never set it carry no trace of the cache.
  Mat<TK, DIM> rot;
  bool has_rot = false;
};
// models that read the polar rotation declare NEEDS_ROTATION
template <class M, class = void> struct needs_rotation { static constexpr bool value = false; };
template <class M> struct needs_rotation<M, decltype(void(M::NEEDS_ROTATION))> {
  static constexpr bool value = M::NEEDS_ROTATION;
};

template <class A, class B, class C, class D, class E>
using prom5_t = prom_t<prom4_t<A, B, C, D>, E>;

// models whose local Newton starts at a closed-form predictor declare HAS_PREDICTOR and guess_r0()
template <class M, class = void> struct has_predictor { static constexpr bool value = false; };
template <class M> struct has_predictor<M, decltype(void(M::HAS_PREDICTOR))> {
  static constexpr bool value = M::HAS_PREDICTOR;
};

C8_DI bool is_plastic(double f, double abs_tol) { return (f > abs_tol) || (fabs(f) < abs_tol); }

// material_params.hpp:12-30
template <class T> C8_DI T mu_of(const T& E, const T& nu) { return E / (2.0 * (1.0 + nu)); }
template <class T> C8_DI T kappa_of(const T& E, const T& nu) { return E / (3.0 * (1.0 - 2.0 * nu)); }
template <class T> C8_DI T lambda_of(const T& E, const T& nu) {
  return E * nu / ((1.0 + nu) * (1.0 - 2.0 * nu));
}

template <class TK, int DIM> C8_DI Mat<TK, DIM> sym_grad(const Mat<TK, DIM>& gu) {
  Mat<TK, DIM> e;
#pragma unroll
  for (int i = 0; i < DIM; ++i)
#pragma unroll
    for (int j = 0; j < DIM; ++j) e.a[i][j] = 0.5 * (gu.a[i][j] + gu.a[j][i]);
  return e;
}
// eps - tr(eps)/3 I   (note: /3 also when DIM == 2, src/small_J2.cpp:275)
template <class TK, int DIM> C8_DI Mat<TK, DIM> dev3(const Mat<TK, DIM>& eps) {
  const TK th = trace(eps) / 3.0;
  Mat<TK, DIM> r = eps;
#pragma unroll
  for (int i = 0; i < DIM; ++i) r.a[i][i] = eps.a[i][i] - th;
  return r;
}

// ---- Hill-48, src/yield_functions.hpp:34-99 ---------------------------------
template <class T> struct Hill { T F, G, H, L, M, N; };
template <class T>
C8_DI Hill<T> hill_params(const T& R00, const T& R11, const T& R22, const T& R01, const T& R02,
                          const T& R12) {
  Hill<T> h;
  const T i00 = inv_sqr(R00), i11 = inv_sqr(R11), i22 = inv_sqr(R22);
  h.F = 0.5 * (i11 + i22 - i00);
  h.G = 0.5 * (i22 + i00 - i11);
  h.H = 0.5 * (i00 + i11 - i22);
  h.L = 1.5 * inv_sqr(R12);
  h.M = 1.5 * inv_sqr(R02);
  h.N = 1.5 * inv_sqr(R01);
  return h;
}
template <class TS, class TP>
C8_DI prom_t<TS, TP> hill_value(const Mat<TS, 3>& t, const Hill<TP>& h) {
  return dsqrt(h.F * sqr(t(1, 1) - t(2, 2)) + h.G * sqr(t(2, 2) - t(0, 0)) +
               h.H * sqr(t(0, 0) - t(1, 1)) +
               2.0 * (h.L * sqr(t(1, 2)) + h.M * sqr(t(0, 2)) + h.N * sqr(t(0, 1))));
}
template <class TS, class TP, class TH>
C8_DI Mat<prom3_t<TS, TP, TH>, 3> hill_normal(const Mat<TS, 3>& t, const Hill<TP>& h, const TH& hv) {
  using R = prom3_t<TS, TP, TH>;
  Mat<R, 3> n;
  const R ih = 1.0 / hv;
  n(0, 0) = ((h.G + h.H) * t(0, 0) - h.H * t(1, 1) - h.G * t(2, 2)) * ih;
  n(1, 1) = ((h.F + h.H) * t(1, 1) - h.H * t(0, 0) - h.F * t(2, 2)) * ih;
  n(2, 2) = ((h.G + h.F) * t(2, 2) - h.G * t(0, 0) - h.F * t(1, 1)) * ih;
  n(0, 1) = (h.N * t(0, 1)) * ih;
  n(0, 2) = (h.M * t(0, 2)) * ih;
  n(1, 2) = (h.L * t(1, 2)) * ih;
  n(1, 0) = n(0, 1); n(2, 0) = n(0, 2); n(2, 1) = n(1, 2);
  return n;
}
template <class T> C8_DI Mat<T, 3> embed3(const Mat<T, 2>& a) {
  Mat<T, 3> r = mat_zero<T, 3>();
  r(0, 0) = a(0, 0); r(0, 1) = a(0, 1); r(1, 0) = a(1, 0); r(1, 1) = a(1, 1);
  return r;
}
template <class T> C8_DI Mat<T, 3> embed3(const Mat<T, 3>& a) { return a; }
template <class T> C8_DI Mat<T, 2> extract2(const Mat<T, 3>& a) {
  Mat<T, 2> r;
  r(0, 0) = a(0, 0); r(0, 1) = a(0, 1); r(1, 0) = a(1, 0); r(1, 1) = a(1, 1);
  return r;
}

// Closed-form structure of the Hill return map (plain doubles), used as the INITIAL GUESS of the local
// Newton of the Hill models: with the associated flow n = M s / hill(s) (M the constant Hill matrix of
// hill_normal) the flow rule  s = s_tr - 2 mu dgam n(s)  and the yield condition  hill(s) = sigma_y  give
//   (I + c M) s = s_tr,   c = 2 mu dgam / sigma_y(alpha_old + dgam),
// a 3 x 3 solve for the normal components and three divisions for the shears, which leaves ONE scalar
// equation  g(dgam) = hill(s(dgam)) - sigma_y(alpha_old + dgam) = 0  (Newton, a few dozen flops per step)
// instead of 4-5 iterations of the 7 x 7 AD Newton.  sigma_y = Y + S (1 - exp(-D alpha)).  Returns false
// (and leaves the outputs untouched) when the scalar iteration does not converge: the caller then keeps
// the reference's starting point.  Only a guess: the reference's Newton confirms |C| < tol at it.
C8_DI bool hill_return_map(const Mat<double, 3>& s_tr, const Hill<double>& h, double mu, double Y,
                           double S, double D, double alpha_old, Mat<double, 3>& s_out, double& dgam_out) {
  auto sigma_y = [&](double a, double& dsy) {
    double v = Y; dsy = 0.0;
    if (S != 0.0) { const double ex = exp(-D * a); v += S * (1.0 - ex); dsy = S * D * ex; }
    return v;
  };
  Mat<double, 3> Mn;   // normal block of M on (s00, s11, s22)
  Mn(0, 0) = h.G + h.H; Mn(0, 1) = -h.H; Mn(0, 2) = -h.G;
  Mn(1, 0) = -h.H; Mn(1, 1) = h.F + h.H; Mn(1, 2) = -h.F;
  Mn(2, 0) = -h.G; Mn(2, 1) = -h.F; Mn(2, 2) = h.G + h.F;
  const double hill_tr = hill_value(s_tr, h);
  double dsy;
  const double sy0 = sigma_y(alpha_old, dsy);
  double dgam = (hill_tr - sy0) / (3.0 * mu + dsy);   // isotropic (J2) estimate
  if (!(dgam > 0.0)) dgam = 0.0;
  Mat<double, 3> s = s_tr;
  bool ok = false, polish = false;
#pragma unroll 1
  for (int it = 0; it < 40; ++it) {
    const double sy = sigma_y(alpha_old + dgam, dsy);
    const double c = 2.0 * mu * dgam / sy;
    Mat<double, 3> An = scale(c, Mn);
    An(0, 0) += 1.0; An(1, 1) += 1.0; An(2, 2) += 1.0;
    const Mat<double, 3> Ai = inverse(An);
    const double i01 = 1.0 / (1.0 + c * h.N), i02 = 1.0 / (1.0 + c * h.M), i12 = 1.0 / (1.0 + c * h.L);
    double sn[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) sn[i] = Ai(i, 0) * s_tr(0, 0) + Ai(i, 1) * s_tr(1, 1) + Ai(i, 2) * s_tr(2, 2);
    s(0, 0) = sn[0]; s(1, 1) = sn[1]; s(2, 2) = sn[2];
    s(0, 1) = s(1, 0) = s_tr(0, 1) * i01;
    s(0, 2) = s(2, 0) = s_tr(0, 2) * i02;
    s(1, 2) = s(2, 1) = s_tr(1, 2) * i12;
    const double hv = hill_value(s, h);
    const double g = hv - sy;
    if (fabs(g) < 1e-13 * mu) {
      if (polish) { ok = true; break; }
      polish = true;   // one more quadratically convergent step: the state then sits at rounding level
    }
    // dg/ddgam = dhill/ds : ds/dc * dc/ddgam - sigma_y',  ds/dc = -(I + c M)^-1 M s
    double ms[3], dsn[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) ms[i] = Mn(i, 0) * sn[0] + Mn(i, 1) * sn[1] + Mn(i, 2) * sn[2];
#pragma unroll
    for (int i = 0; i < 3; ++i) dsn[i] = -(Ai(i, 0) * ms[0] + Ai(i, 1) * ms[1] + Ai(i, 2) * ms[2]);
    const double ih = 1.0 / hv;
    double dh = (ms[0] * dsn[0] + ms[1] * dsn[1] + ms[2] * dsn[2]) * ih;   // n_ii = ms_i / hill
    dh += 2.0 * ih * (h.N * s(0, 1) * (-h.N * s(0, 1) * i01) + h.M * s(0, 2) * (-h.M * s(0, 2) * i02) +
                      h.L * s(1, 2) * (-h.L * s(1, 2) * i12));
    const double dc = 2.0 * mu * (sy - dgam * dsy) / (sy * sy);
    const double dg = dh * dc - dsy;
    if (!(dg < 0.0)) break;
    double nd = dgam - g / dg;
    if (!(nd > 0.0)) nd = 0.5 * dgam;   // stay on the loading side
    dgam = nd;
  }
  if (!ok) return false;
  s_out = s; dgam_out = dgam;
  return true;
}

// Plane-stress counterpart (small_hill_plane_stress): the in-plane stress obeys sigma = Cps : (eps - pstrain)
// with Cps = 2 mu I + lambda' 1 x 1, lambda' = 2 mu lambda / (lambda + 2 mu) (eps_zz eliminated), so
//   (I + c Cps M2) sigma = sigma_tr,  c = dgam / sigma_y(alpha_old + dgam),
// M2 the Hill matrix restricted to (00, 11, 01) at sigma_zz = 0: a 2 x 2 solve + one division per step.
C8_DI bool hill_return_map_plane_stress(const Mat<double, 2>& s_tr, const Hill<double>& h, double mu,
                                        double lambda, double Y, double S, double D, double alpha_old,
                                        Mat<double, 2>& s_out, double& dgam_out) {
  auto sigma_y = [&](double a, double& dsy) {
    double v = Y; dsy = 0.0;
    if (S != 0.0) { const double ex = exp(-D * a); v += S * (1.0 - ex); dsy = S * D * ex; }
    return v;
  };
  auto hill2 = [&](double a, double b, double c) {
    return sqrt(h.F * b * b + h.G * a * a + h.H * (a - b) * (a - b) + 2.0 * h.N * c * c);
  };
  const double lp = 2.0 * mu * lambda / (lambda + 2.0 * mu);
  const double m00 = h.G + h.H, m01 = -h.H, m11 = h.F + h.H;
  // CM = Cn Mn, Cn = [[2 mu + lp, lp], [lp, 2 mu + lp]]
  const double c00 = (2.0 * mu + lp) * m00 + lp * m01, c01 = (2.0 * mu + lp) * m01 + lp * m11;
  const double c10 = lp * m00 + (2.0 * mu + lp) * m01, c11 = lp * m01 + (2.0 * mu + lp) * m11;
  const double a_tr = s_tr(0, 0), b_tr = s_tr(1, 1), t_tr = s_tr(0, 1);
  double dsy;
  const double sy0 = sigma_y(alpha_old, dsy);
  double dgam = (hill2(a_tr, b_tr, t_tr) - sy0) / (3.0 * mu + dsy);
  if (!(dgam > 0.0)) dgam = 0.0;
  double sa = a_tr, sb = b_tr, st = t_tr;
  bool ok = false, polish = false;
#pragma unroll 1
  for (int it = 0; it < 40; ++it) {
    const double sy = sigma_y(alpha_old + dgam, dsy);
    const double c = dgam / sy;
    const double a00 = 1.0 + c * c00, a01 = c * c01, a10 = c * c10, a11 = 1.0 + c * c11;
    const double idet = 1.0 / (a00 * a11 - a01 * a10);
    const double i01 = 1.0 / (1.0 + c * 2.0 * mu * h.N);
    sa = (a11 * a_tr - a01 * b_tr) * idet;
    sb = (a00 * b_tr - a10 * a_tr) * idet;
    st = t_tr * i01;
    const double hv = hill2(sa, sb, st);
    const double g = hv - sy;
    if (fabs(g) < 1e-13 * mu) {
      if (polish) { ok = true; break; }
      polish = true;
    }
    const double ma = m00 * sa + m01 * sb, mb = m01 * sa + m11 * sb;     // M2 s (normal part)
    const double ca = c00 * sa + c01 * sb, cb = c10 * sa + c11 * sb;     // Cn M2 s
    const double da = -(a11 * ca - a01 * cb) * idet, db = -(a00 * cb - a10 * ca) * idet;   // ds/dc
    const double ih = 1.0 / hv;
    const double dh = (ma * da + mb * db) * ih + 2.0 * ih * h.N * st * (-2.0 * mu * h.N * st * i01);
    const double dc = (sy - dgam * dsy) / (sy * sy);
    const double dg = dh * dc - dsy;
    if (!(dg < 0.0)) break;
    double nd = dgam - g / dg;
    if (!(nd > 0.0)) nd = 0.5 * dgam;
    dgam = nd;
  }
  if (!ok) return false;
  s_out(0, 0) = sa; s_out(1, 1) = sb; s_out(0, 1) = s_out(1, 0) = st;
  dgam_out = dgam;
  return true;
}

// small-strain deviatoric stress 2 mu (dev3(eps) - pstrain), shared by several models
template <int DIM, class TK, class TX, class TP>
C8_DI Mat<prom3_t<TK, TX, TP>, DIM> small_dev_stress(const Mat<TK, DIM>& gu, const TX* xi,
                                                     const TP& E, const TP& nu) {
  const TP mu = mu_of(E, nu);
  const Mat<TK, DIM> de = dev3(sym_grad(gu));
  const Mat<TX, DIM> ps = unpack_sym<TX, DIM>(xi);
  return scale(2.0 * mu, de - ps);
}

// =============================================================================
template <int DIM>
struct Elastic {
  static constexpr int NXI = 1, NPAR = 4, TYPE = L_ELASTIC;
  static constexpr bool FINITE = false, HAS_NEWTON = false, PLANE_STRESS = false;
  // on the elastic branch C = xi - (a function of xi_prev and the kinematics): dC/dxi = I
  static constexpr bool ELASTIC_J_IDENTITY = false;
  static constexpr int Z_STRETCH = -1;
  static C8_DI void init(double* xi) { xi[0] = 0.0; }
  template <class K> static C8_DI void guess(const K&, const double*, const double*, double, double* xi) { xi[0] = 0.0; }
  template <class TK, class TKP, class TX, class TXP, class TP>
  static C8_DI int residual(const Kin<DIM, TK, TKP>&, const TX*, const TXP*, const TP*, double,
                            prom5_t<TK, TKP, TX, TXP, TP>* C) {
    C[0] = conv<prom5_t<TK, TKP, TX, TXP, TP>>(0.0);  // evaluate() leaves R untouched (== 0)
    return 0;
  }
  template <class TK, class TKP, class TX, class TP>
  static C8_DI Mat<prom3_t<TK, TX, TP>, DIM> dev_cauchy(const Kin<DIM, TK, TKP>& k, const TX*,
                                                        const TP* par) {
    using R = prom3_t<TK, TX, TP>;
    const TP mu = mu_of(par[0], par[1]);
    return mat_conv<R>(scale(2.0 * mu, dev3(sym_grad(k.gu))));
  }
  template <class TK, class TKP, class TX, class TP>
  static C8_DI prom3_t<TK, TX, TP> hydro(const Kin<DIM, TK, TKP>& k, const TX*, const TP* par) {
    const TP kappa = kappa_of(par[0], par[1]);
    return conv<prom3_t<TK, TX, TP>>(kappa * trace(sym_grad(k.gu)) -
                                     par[2] * par[3] * par[0] / (1.0 - 2.0 * par[1]));
  }
  template <class TP> static C8_DI TP pscale(const TP* par) { return kappa_of(par[0], par[1]); }
};

// =============================================================================
template <int DIM>
struct SmallJ2 {
  static constexpr int NS = SymIdx<DIM>::n;
  static constexpr int NXI = NS + 1, NPAR = 6, TYPE = L_SMALL_J2;
  static constexpr bool FINITE = false, HAS_NEWTON = true, PLANE_STRESS = false;
  // on the elastic branch C = xi - (a function of xi_prev and the kinematics): dC/dxi = I
  static constexpr bool ELASTIC_J_IDENTITY = true;
  static constexpr int Z_STRETCH = -1;
  static C8_DI void init(double* xi) {
#pragma unroll
    for (int i = 0; i < NXI; ++i) xi[i] = 0.0;
  }
  // Initial guess of the local Newton.  The reference starts from xi_prev ("pick an initial guess",
  // src/small_J2.cpp:126-135); with linear hardening the radial return is closed form, so a yielding
  // point starts AT the solution of the same residual (the Newton then only confirms |C| < tol):
  //   dgam = (|s_tr| - sqrt(2/3) sigma_y(alpha_old)) / (2 mu + 2/3 K),  n = s_tr / |s_tr|,
  //   pstrain = pstrain_old + dgam n,  alpha = alpha_old + sqrt(2/3) dgam.
  // HAS_PREDICTOR: guess_r0() also returns the residual norm at the REFERENCE's starting point
  // (xi_prev), which is the R_norm_0 of its relative convergence test (src/small_J2.cpp:147-151):
  // there R_pstrain = 0 (dgam = 0) and R_alpha = f, so the norm is |f|; -1 where no predictor was
  // applied (the Newton then starts at the reference's own point and measures R_norm_0 itself).
  static constexpr bool HAS_PREDICTOR = true;
  template <class K> static C8_DI void guess(const K& k, const double* xip, const double* par, double abs_tol,
                                             double* xi) {
    (void)guess_r0(k, xip, par, abs_tol, xi);
  }
  template <class K> static C8_DI double guess_r0(const K& k, const double* xip, const double* par,
                                                  double abs_tol, double* xi) {
#pragma unroll
    for (int i = 0; i < NXI; ++i) xi[i] = xip[i];
    const double sqrt_23 = 0.81649658092772603;
    const double mu = mu_of(par[0], par[1]);
    const Mat<double, DIM> s = small_dev_stress<DIM>(k.gu, xip, par[0], par[1]);
    const double s_mag = norm(s);
    const double f = (s_mag - sqrt_23 * (par[3] + par[2] * xip[NS])) / mu;
    double r0 = -1.0;
    if (is_plastic(f, abs_tol) && s_mag > 0.0) {
      r0 = fabs(f);
      const double dgam = mu * f / (2.0 * mu + (2.0 / 3.0) * par[2]);
      const double g = dgam / s_mag;
#pragma unroll
      for (int i = 0; i < DIM; ++i)
#pragma unroll
        for (int j = i; j < DIM; ++j) xi[SymIdx<DIM>::idx(i, j)] = xip[SymIdx<DIM>::idx(i, j)] + g * s.a[i][j];
      xi[NS] = xip[NS] + sqrt_23 * dgam;
    }
    return r0;
  }
  template <class TK, class TKP, class TX, class TXP, class TP>
  static C8_DI int residual(const Kin<DIM, TK, TKP>& k, const TX* xi, const TXP* xip,
                            const TP* par, double abs_tol, prom5_t<TK, TKP, TX, TXP, TP>* C) {
    using R = prom5_t<TK, TKP, TX, TXP, TP>;
    const double sqrt_23 = 0.81649658092772603, sqrt_32 = 1.2247448713915890;
    const TP mu = mu_of(par[0], par[1]);
    const auto s = small_dev_stress<DIM>(k.gu, xi, par[0], par[1]);
    const auto s_mag = norm(s);
    const auto sigma_yield = par[3] + par[2] * xi[NS];
    const auto f = (s_mag - sqrt_23 * sigma_yield) / val(mu);
    if (is_plastic(val(f), abs_tol)) {
      const auto dgam = sqrt_32 * (xi[NS] - xip[NS]);
      const auto g = dgam / s_mag;
#pragma unroll
      for (int i = 0; i < DIM; ++i)
#pragma unroll
        for (int j = i; j < DIM; ++j) {
          const int q = SymIdx<DIM>::idx(i, j);
          C[q] = conv<R>(xi[q] - xip[q] - g * s.a[i][j]);
        }
      C[NS] = conv<R>(f);
      return PATH_PLASTIC;
    }
#pragma unroll
    for (int q = 0; q < NXI; ++q) C[q] = conv<R>(xi[q] - xip[q]);
    return PATH_ELASTIC;
  }
  template <class TK, class TKP, class TX, class TP>
  static C8_DI Mat<prom3_t<TK, TX, TP>, DIM> dev_cauchy(const Kin<DIM, TK, TKP>& k, const TX* xi,
                                                        const TP* par) {
    return mat_conv<prom3_t<TK, TX, TP>>(small_dev_stress<DIM>(k.gu, xi, par[0], par[1]));
  }
  template <class TK, class TKP, class TX, class TP>
  static C8_DI prom3_t<TK, TX, TP> hydro(const Kin<DIM, TK, TKP>& k, const TX*, const TP* par) {
    const TP kappa = kappa_of(par[0], par[1]);
    return conv<prom3_t<TK, TX, TP>>(kappa * trace(sym_grad(k.gu)) -
                                     par[4] * par[5] * par[0] / (1.0 - 2.0 * par[1]));
  }
  template <class TP> static C8_DI TP pscale(const TP* par) { return kappa_of(par[0], par[1]); }
};

// =============================================================================
// SmallHill (3-D only: the Hill function reads the zz components)
template <int DIM>
struct SmallHill {
  static_assert(DIM == 3, "small_hill is a 3-D model");
  static constexpr int NS = 6, NXI = 7, NPAR = 11, TYPE = L_SMALL_HILL;
  static constexpr int K1_TEAM = 256;   // persistent K1: one 256-thread CTA per SM (forward.cuh, K1PBlock)
  static constexpr bool FINITE = false, HAS_NEWTON = true, PLANE_STRESS = false;
  // on the elastic branch C = xi - (a function of xi_prev and the kinematics): dC/dxi = I
  static constexpr bool ELASTIC_J_IDENTITY = true;
  static constexpr int Z_STRETCH = -1;
  static C8_DI void init(double* xi) {
#pragma unroll
    for (int i = 0; i < NXI; ++i) xi[i] = 0.0;
  }
  // The reference starts from xi_prev (src/small_hill.cpp:142-151).  A yielding point starts at the
  // solution of hill_return_map() instead (same flow rule and yield condition as residual() below), so
  // the AD Newton confirms |C| < tol in one evaluation; guess_r0 returns the residual norm at the
  // REFERENCE's starting point for its relative test: there R_pstrain = 0 (dgam = 0), the zz row holds
  // tr(pstrain_old) and R_alpha = f_trial.
  static constexpr bool HAS_PREDICTOR = true;
  template <class K> static C8_DI void guess(const K& k, const double* xip, const double* par, double abs_tol,
                                             double* xi) {
    (void)guess_r0(k, xip, par, abs_tol, xi);
  }
  template <class K> static C8_DI double guess_r0(const K& k, const double* xip, const double* par,
                                                  double abs_tol, double* xi) {
#pragma unroll
    for (int i = 0; i < NXI; ++i) xi[i] = xip[i];
    const double mu = mu_of(par[0], par[1]);
    const Hill<double> hp = hill_params(par[3], par[4], par[5], par[6], par[7], par[8]);
    const Mat<double, 3> s_tr = small_dev_stress<3>(k.gu, xip, par[0], par[1]);
    const double hill_tr = hill_value(s_tr, hp);
    const double f = (hill_tr - (par[2] + par[9] * (1.0 - exp(-par[10] * xip[NS])))) / mu;
    if (!(is_plastic(f, abs_tol) && hill_tr > 0.0)) return -1.0;
    const double tr_old = xip[0] + xip[3] + xip[5];
    const double r0 = sqrt(f * f + tr_old * tr_old);
    Mat<double, 3> s; double dgam;
    if (hill_return_map(s_tr, hp, mu, par[2], par[9], par[10], xip[NS], s, dgam)) {
      const Mat<double, 3> n = hill_normal(s, hp, hill_value(s, hp));
      xi[0] = xip[0] + dgam * n(0, 0); xi[1] = xip[1] + dgam * n(0, 1); xi[2] = xip[2] + dgam * n(0, 2);
      xi[3] = xip[3] + dgam * n(1, 1); xi[4] = xip[4] + dgam * n(1, 2);
      xi[5] = -(xi[0] + xi[3]);   // the zz row of the residual is tr(pstrain) = 0
      xi[NS] = xip[NS] + dgam;
    }
    return r0;
  }
  template <class TK, class TKP, class TX, class TXP, class TP>
  static C8_DI int residual(const Kin<3, TK, TKP>& k, const TX* xi, const TXP* xip, const TP* par,
                            double abs_tol, prom5_t<TK, TKP, TX, TXP, TP>* C) {
    using R = prom5_t<TK, TKP, TX, TXP, TP>;
    const TP mu = mu_of(par[0], par[1]);
    const Hill<TP> hp = hill_params(par[3], par[4], par[5], par[6], par[7], par[8]);
    const auto s = small_dev_stress<3>(k.gu, xi, par[0], par[1]);
    const auto hill = hill_value(s, hp);
    const auto sigma_yield = par[2] + par[9] * (1.0 - dexp(-par[10] * xi[NS]));
    const auto f = (hill - sigma_yield) / val(mu);
    if (is_plastic(val(f), abs_tol)) {
      const auto n = hill_normal(s, hp, hill);
      const auto dgam = xi[NS] - xip[NS];
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = i; j < 3; ++j) {
          const int q = SymIdx<3>::idx(i, j);
          C[q] = conv<R>(xi[q] - xip[q] - dgam * n.a[i][j]);
        }
      C[5] = conv<R>(xi[0] + xi[3] + xi[5]);  // R_pstrain(2,2) := tr(pstrain), src/small_hill.cpp:241
      C[NS] = conv<R>(f);
      return PATH_PLASTIC;
    }
#pragma unroll
    for (int q = 0; q < NXI; ++q) C[q] = conv<R>(xi[q] - xip[q]);
    return PATH_ELASTIC;
  }
  template <class TK, class TKP, class TX, class TP>
  static C8_DI Mat<prom3_t<TK, TX, TP>, 3> dev_cauchy(const Kin<3, TK, TKP>& k, const TX* xi,
                                                      const TP* par) {
    return mat_conv<prom3_t<TK, TX, TP>>(small_dev_stress<3>(k.gu, xi, par[0], par[1]));
  }
  template <class TK, class TKP, class TX, class TP>
  static C8_DI prom3_t<TK, TX, TP> hydro(const Kin<3, TK, TKP>& k, const TX*, const TP* par) {
    const TP kappa = kappa_of(par[0], par[1]);
    return conv<prom3_t<TK, TX, TP>>(kappa * trace(sym_grad(k.gu)));
  }
  template <class TP> static C8_DI TP pscale(const TP* par) { return kappa_of(par[0], par[1]); }
};

// =============================================================================
template <int DIM>
struct SmallHillPlaneStress {
  static_assert(DIM == 2, "plane stress is 2-D");
  static constexpr int NS = 3, NXI = 4, NPAR = 9, TYPE = L_SMALL_HILL_PLANE_STRESS;
  static constexpr bool FINITE = false, HAS_NEWTON = true, PLANE_STRESS = true;
  // on the elastic branch C = xi - (a function of xi_prev and the kinematics): dC/dxi = I
  static constexpr bool ELASTIC_J_IDENTITY = true;
  static constexpr int Z_STRETCH = -1;
  static C8_DI void init(double* xi) {
#pragma unroll
    for (int i = 0; i < NXI; ++i) xi[i] = 0.0;
  }
  // Reference start = xi_prev (src/small_hill_plane_stress.cpp:140-149); a yielding point starts at the
  // solution of hill_return_map_plane_stress() (see SmallHill); R_norm_0 = |f_trial| at the reference's start.
  static constexpr bool HAS_PREDICTOR = true;
  template <class K> static C8_DI void guess(const K& k, const double* xip, const double* par, double abs_tol,
                                             double* xi) {
    (void)guess_r0(k, xip, par, abs_tol, xi);
  }
  template <class K> static C8_DI double guess_r0(const K& k, const double* xip, const double* par,
                                                  double abs_tol, double* xi) {
#pragma unroll
    for (int i = 0; i < NXI; ++i) xi[i] = xip[i];
    const double mu = mu_of(par[0], par[1]), lambda = lambda_of(par[0], par[1]);
    const Hill<double> hp = hill_params(par[5], par[6], par[7], par[8], 1.0, 1.0);
    const Mat<double, 2> s_tr = cauchy(k, xip, par);
    const Mat<double, 3> s3 = embed3(s_tr);
    const double hill_tr = hill_value(s3, hp);
    const double f = (hill_tr - (par[2] + par[3] * (1.0 - exp(-par[4] * xip[NS])))) / mu;
    if (!(is_plastic(f, abs_tol) && hill_tr > 0.0)) return -1.0;
    Mat<double, 2> sg; double dgam;
    if (hill_return_map_plane_stress(s_tr, hp, mu, lambda, par[2], par[3], par[4], xip[NS], sg, dgam)) {
      const Mat<double, 3> g3 = embed3(sg);
      const Mat<double, 3> n = hill_normal(g3, hp, hill_value(g3, hp));
      xi[0] = xip[0] + dgam * n(0, 0); xi[1] = xip[1] + dgam * n(0, 1); xi[2] = xip[2] + dgam * n(1, 1);
      xi[NS] = xip[NS] + dgam;
    }
    return fabs(f);
  }
  // full in-plane Cauchy stress with eps_zz eliminated, src/small_hill_plane_stress.cpp:278-327
  template <class TK, class TKP, class TX, class TP>
  static C8_DI Mat<prom3_t<TK, TX, TP>, 2> cauchy(const Kin<2, TK, TKP>& k, const TX* xi,
                                                  const TP* par) {
    using R = prom3_t<TK, TX, TP>;
    const TP mu = mu_of(par[0], par[1]);
    const TP lambda = lambda_of(par[0], par[1]);
    const Mat<TK, 2> eps = sym_grad(k.gu);
    const Mat<TX, 2> ps = unpack_sym<TX, 2>(xi);
    const R eps_zz = -(lambda * trace(eps) + 2.0 * mu * trace(ps)) / (lambda + 2.0 * mu);
    const R eps_kk = trace(eps) + eps_zz;
    Mat<R, 2> sig = mat_conv<R>(scale(2.0 * mu, eps - ps));
    const R l = lambda * eps_kk;
    sig(0, 0) += l; sig(1, 1) += l;
    return sig;
  }
  template <class TK, class TKP, class TX, class TXP, class TP>
  static C8_DI int residual(const Kin<2, TK, TKP>& k, const TX* xi, const TXP* xip, const TP* par,
                            double abs_tol, prom5_t<TK, TKP, TX, TXP, TP>* C) {
    using R = prom5_t<TK, TKP, TX, TXP, TP>;
    const TP mu = mu_of(par[0], par[1]);
    const TP one = conv<TP>(1.0);
    const Hill<TP> hp = hill_params(par[5], par[6], par[7], par[8], one, one);
    const auto sig3 = embed3(cauchy(k, xi, par));
    const auto hill = hill_value(sig3, hp);
    const auto sigma_yield = par[2] + par[3] * (1.0 - dexp(-par[4] * xi[NS]));
    const auto f = (hill - sigma_yield) / val(mu);
    if (is_plastic(val(f), abs_tol)) {
      const auto n = hill_normal(sig3, hp, hill);
      const auto dgam = xi[NS] - xip[NS];
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = i; j < 2; ++j) {
          const int q = SymIdx<2>::idx(i, j);
          C[q] = conv<R>(xi[q] - xip[q] - dgam * n.a[i][j]);
        }
      C[NS] = conv<R>(f);
      return PATH_PLASTIC;
    }
#pragma unroll
    for (int q = 0; q < NXI; ++q) C[q] = conv<R>(xi[q] - xip[q]);
    return PATH_ELASTIC;
  }
  template <class TP> static C8_DI TP pscale(const TP*) { return conv<TP>(0.0); }
};

// =============================================================================
template <int DIM>
struct SmallHillPlaneStrain {
  static_assert(DIM == 2, "plane strain is 2-D");
  static constexpr int NS = 3, NXI = 4, NPAR = 9, TYPE = L_SMALL_HILL_PLANE_STRAIN;
  static constexpr bool FINITE = false, HAS_NEWTON = true, PLANE_STRESS = false;
  // on the elastic branch C = xi - (a function of xi_prev and the kinematics): dC/dxi = I
  static constexpr bool ELASTIC_J_IDENTITY = true;
  static constexpr int Z_STRETCH = -1;
  static C8_DI void init(double* xi) {
#pragma unroll
    for (int i = 0; i < NXI; ++i) xi[i] = 0.0;
  }
  // Reference start = xi_prev (src/small_hill_plane_strain.cpp:140-149).  The model is the 3-D one with
  // eps_zz = 0, pstrain_zz = -(p00 + p11) and no out-of-plane shear, so a yielding point starts at the
  // solution of the 3-D hill_return_map() (see SmallHill); R_norm_0 = |f_trial|.
  static constexpr bool HAS_PREDICTOR = true;
  template <class K> static C8_DI void guess(const K& k, const double* xip, const double* par, double abs_tol,
                                             double* xi) {
    (void)guess_r0(k, xip, par, abs_tol, xi);
  }
  template <class K> static C8_DI double guess_r0(const K& k, const double* xip, const double* par,
                                                  double abs_tol, double* xi) {
#pragma unroll
    for (int i = 0; i < NXI; ++i) xi[i] = xip[i];
    const double mu = mu_of(par[0], par[1]);
    const Hill<double> hp = hill_params(par[5], par[6], par[7], par[8], 1.0, 1.0);
    Mat<double, 3> s_tr = embed3(small_dev_stress<2>(k.gu, xip, par[0], par[1]));
    s_tr(2, 2) = 2.0 * mu * (-trace(sym_grad(k.gu)) / 3.0 + (xip[0] + xip[2]));
    const double hill_tr = hill_value(s_tr, hp);
    const double f = (hill_tr - (par[2] + par[3] * (1.0 - exp(-par[4] * xip[NS])))) / mu;
    if (!(is_plastic(f, abs_tol) && hill_tr > 0.0)) return -1.0;
    Mat<double, 3> sg; double dgam;
    if (hill_return_map(s_tr, hp, mu, par[2], par[3], par[4], xip[NS], sg, dgam)) {
      const Mat<double, 3> n = hill_normal(sg, hp, hill_value(sg, hp));
      xi[0] = xip[0] + dgam * n(0, 0); xi[1] = xip[1] + dgam * n(0, 1); xi[2] = xip[2] + dgam * n(1, 1);
      xi[NS] = xip[NS] + dgam;
    }
    return fabs(f);
  }
  template <class TK, class TKP, class TX, class TXP, class TP>
  static C8_DI int residual(const Kin<2, TK, TKP>& k, const TX* xi, const TXP* xip, const TP* par,
                            double abs_tol, prom5_t<TK, TKP, TX, TXP, TP>* C) {
    using R = prom5_t<TK, TKP, TX, TXP, TP>;
    const TP mu = mu_of(par[0], par[1]);
    const TP one = conv<TP>(1.0);
    const Hill<TP> hp = hill_params(par[5], par[6], par[7], par[8], one, one);
    const auto s2 = small_dev_stress<2>(k.gu, xi, par[0], par[1]);
    const Mat<TK, 2> eps = sym_grad(k.gu);
    auto s3 = embed3(s2);
    s3(2, 2) = 2.0 * mu * (-trace(eps) / 3.0 + (xi[0] + xi[2]));
    const auto hill = hill_value(s3, hp);
    const auto sigma_yield = par[2] + par[3] * (1.0 - dexp(-par[4] * xi[NS]));
    const auto f = (hill - sigma_yield) / val(mu);
    if (is_plastic(val(f), abs_tol)) {
      const auto n = hill_normal(s3, hp, hill);
      const auto dgam = xi[NS] - xip[NS];
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = i; j < 2; ++j) {
          const int q = SymIdx<2>::idx(i, j);
          C[q] = conv<R>(xi[q] - xip[q] - dgam * n.a[i][j]);
        }
      C[NS] = conv<R>(f);
      return PATH_PLASTIC;
    }
#pragma unroll
    for (int q = 0; q < NXI; ++q) C[q] = conv<R>(xi[q] - xip[q]);
    return PATH_ELASTIC;
  }
  template <class TK, class TKP, class TX, class TP>
  static C8_DI Mat<prom3_t<TK, TX, TP>, 2> dev_cauchy(const Kin<2, TK, TKP>& k, const TX* xi,
                                                      const TP* par) {
    return mat_conv<prom3_t<TK, TX, TP>>(small_dev_stress<2>(k.gu, xi, par[0], par[1]));
  }
  template <class TK, class TKP, class TX, class TP>
  static C8_DI prom3_t<TK, TX, TP> hydro(const Kin<2, TK, TKP>& k, const TX*, const TP* par) {
    const TP kappa = kappa_of(par[0], par[1]);
    return conv<prom3_t<TK, TX, TP>>(kappa * trace(sym_grad(k.gu)));
  }
  template <class TP> static C8_DI TP pscale(const TP* par) { return kappa_of(par[0], par[1]); }
};

// =============================================================================
// be_bar_trial = rF_bar (zeta_old + Ie_old I) rF_bar^T, src/hyper_J2.cpp:136-154
template <int DIM, class TK, class TKP, class TXP>
C8_DI Mat<prom3_t<TK, TKP, TXP>, DIM> be_bar_trial(const Kin<DIM, TK, TKP>& k,
                                                   const Mat<TXP, DIM>& zeta_old,
                                                   const TXP& Ie_old,
                                                   prom_t<TK, TKP>* det_rF_13_out = nullptr) {
  const Mat<TK, DIM> F = add_diag(k.gu, 1.0);
  const Mat<TKP, DIM> Fp = add_diag(k.gup, 1.0);
  const auto rF = F * inverse(Fp);
  const auto d13 = dcbrt(det(rF));
  if (det_rF_13_out) *det_rF_13_out = d13;
  const auto rFb = scale(1.0 / d13, rF);
  const Mat<TXP, DIM> core = add_diag(zeta_old, Ie_old);
  return rFb * core * transpose(rFb);
}

template <class TX, class TP>
C8_DI prom_t<TX, TP> hyper_yield(const TX& alpha, const TP& Y, const TP& S, const TP& D,
                                       const TP& A, const TP& n, const TP& K) {
  // sigma_y = Y + S (1 - exp(-D alpha)) + A (alpha + 1e-12)^n + K alpha, src/hyper_J2.cpp:260-262
  using R = prom_t<TX, TP>;
  if constexpr (is_dual<TP>::value) {
    return conv<R>(Y + S * (1.0 - dexp(-D * alpha)) + A * dpow(alpha + 1e-12, n) + K * alpha);
  } else {
    // Unseeded parameters: a saturation / power-law term whose coefficient is exactly zero
    // contributes exactly zero (for a finite exp / pow), so its transcendental is skipped -- a
    // parameter-uniform branch.  The sum keeps the reference's order, so non-zero terms are
    // unchanged.  (The reference's own hyper-J2 deck, test/primal/notch_hyper_J2.yaml.in:25-34,
    // has S = D = A = n = 0; pow() alone was ~1/6 of K1's instructions.)
    R sat = conv<R>(0.0), pw = conv<R>(0.0);
    if (S != 0.0) sat = conv<R>(S * (1.0 - dexp(-D * alpha)));
    if (A != 0.0) pw = conv<R>(A * dpow(alpha + 1e-12, n));
    return conv<R>(Y + sat + pw + K * alpha);
  }
}

template <int DIM>
struct HyperJ2 {
  static constexpr int NS = SymIdx<DIM>::n;
  static constexpr int NXI = NS + 2, NPAR = 8, TYPE = L_HYPER_J2;
  static constexpr bool FINITE = true, HAS_NEWTON = true, PLANE_STRESS = false;
  // on the elastic branch C = xi - (a function of xi_prev and the kinematics): dC/dxi = I
  static constexpr bool ELASTIC_J_IDENTITY = true;
  static constexpr int Z_STRETCH = -1;
  static C8_DI void init(double* xi) {
#pragma unroll
    for (int i = 0; i < NS; ++i) xi[i] = 0.0;
    xi[NS] = 1.0; xi[NS + 1] = 0.0;
  }
  // trial-state initial guess, src/hyper_J2.cpp:167-179.  For a yielding point with linear hardening
  // (S = A = 0) the flow rule and the yield condition of the SAME residual are solved in closed form
  // at Ie = Ie_trial (zeta stays parallel to dev be_bar_trial):
  //   dgam = mu f_tr / (2 mu Ie + 2/3 K),  |zeta| = |dev bt| - 2 Ie dgam,  alpha = alpha_old + sqrt(2/3) dgam
  // and Ie follows from det(zeta + Ie I) = 1 by a scalar Newton: the full local Newton then starts at
  // the solution (1 evaluation to confirm |C| < tol instead of 4-5 iterations).
  // HAS_PREDICTOR / guess_r0: as in SmallJ2.  At the reference's starting point (the trial state,
  // alpha = alpha_old) R_zeta = 0, R_Ie = det(be_bar_trial) - 1 = 0 up to rounding (be_bar_trial is a
  // unimodular transform of zeta_old + Ie_old I) and R_alpha = f, so R_norm_0 = |f_trial|.
  static constexpr bool HAS_PREDICTOR = true;
  template <class K> static C8_DI void guess(const K& k, const double* xip, const double* par, double abs_tol,
                                             double* xi) {
    (void)guess_r0(k, xip, par, abs_tol, xi);
  }
  template <class K> static C8_DI double guess_r0(const K& k, const double* xip, const double* par,
                                                  double abs_tol, double* xi) {
    double r0 = -1.0;
    const Mat<double, DIM> zo = unpack_sym<double, DIM>(xip);
    const Mat<double, DIM> bt = be_bar_trial<DIM>(k, zo, xip[NS]);
    Mat<double, DIM> z = dev(bt);
    double Ie = trace(bt) / 3.0;
    double alpha = xip[NS + 1];
    if (par[3] == 0.0 && par[5] == 0.0) {
      const double sqrt_23 = 0.81649658092772603;
      const double mu = mu_of(par[0], par[1]);
      const double z_mag = norm(z);
      const double f = (mu * z_mag - sqrt_23 * (par[2] + par[7] * alpha)) / mu;
      if (is_plastic(f, abs_tol) && z_mag > 0.0) {
        r0 = fabs(f);
        // zeta = nhat m(Ie), m = |dev bt| - 2 Ie dgam(Ie); Ie from the isochoric constraint
        // det(zeta + Ie I) = 1 by a scalar Newton in plain doubles (a few dozen flops per step,
        // against ~10^3 for one AD evaluation + 8x8 solve of the full local Newton)
        const Mat<double, DIM> nhat = scale(1.0 / z_mag, z);
        const double c0 = mu * f, kk = (2.0 / 3.0) * par[7];
        double m = 0.0, dgam = 0.0;
#pragma unroll 1
        for (int it = 0; it < 8; ++it) {
          const double den = 2.0 * mu * Ie + kk;
          dgam = c0 / den;
          m = z_mag - 2.0 * Ie * dgam;
          if constexpr (DIM == 3) {
            const Mat<double, DIM> Amat = add_diag(scale(m, nhat), Ie);
            const double g = det(Amat) - 1.0;
            if (fabs(g) < 1e-15) break;
            const double ddgam = -c0 * 2.0 * mu / (den * den);
            const double dm = -2.0 * dgam - 2.0 * Ie * ddgam;
            const Mat<double, DIM> adj = cofactor_T(Amat);
            double dg = 0.0;
#pragma unroll
            for (int i = 0; i < DIM; ++i)
#pragma unroll
              for (int j = 0; j < DIM; ++j) dg += adj(j, i) * (dm * nhat(i, j) + (i == j ? 1.0 : 0.0));
            Ie -= g / dg;
          } else {
            break;
          }
        }
        z = scale(m, nhat);
        alpha += sqrt_23 * dgam;
      }
    } else if constexpr (DIM == 3) {
      // Non-linear isotropic hardening (Voce and / or power law): the same radial-return structure,
      // zeta = nhat m with m = |dev bt| - 2 Ie dgam, leaves TWO scalar unknowns (dgam, Ie) and the two
      // scalar equations  g1 = mu m - sqrt(2/3) sigma_y(alpha_old + sqrt(2/3) dgam) = 0  (yield) and
      // g2 = det(m nhat + Ie I) - 1 = 0  (isochoric): a 2 x 2 Newton in plain doubles (one exp / pow
      // pair per step) instead of 4-5 iterations of the AD Newton with its 8 x 8 solve.  It starts
      // right of the root in dgam (linear hardening with the smallest slope K), where the convex g1
      // cannot stall on the infinite slope of the power law at alpha = 0.  Only an initial guess:
      // whatever it leaves, the local Newton of the reference finishes.
      const double sqrt_23 = 0.81649658092772603;
      const double mu = mu_of(par[0], par[1]);
      const double Y = par[2], S = par[3], D = par[4], A = par[5], n = par[6], K = par[7];
      auto sigma_y = [&](double a, double& dsy) {
        double v = Y + K * a;
        dsy = K;
        if (S != 0.0) { const double ex = exp(-D * a); v += S * (1.0 - ex); dsy += S * D * ex; }
        if (A != 0.0) { const double pw = pow(a + 1e-12, n); v += A * pw; dsy += A * n * pw / (a + 1e-12); }
        return v;
      };
      const double z_mag = norm(z);
      double dsy0;
      const double f = (mu * z_mag - sqrt_23 * sigma_y(alpha, dsy0)) / mu;
      if (is_plastic(f, abs_tol) && z_mag > 0.0) {
        r0 = fabs(f);
        const Mat<double, DIM> nhat = scale(1.0 / z_mag, z);
        const double kmin = K > 0.0 ? K : 0.0;
        double dgam = mu * f / (2.0 * mu * Ie + (2.0 / 3.0) * kmin);
        double m = z_mag - 2.0 * Ie * dgam;
        bool polish = false;   // one more (quadratically convergent) step after the test first passes:
                               // the state then sits at rounding level like the reference's last iterate
#pragma unroll 1
        for (int it = 0; it < 24; ++it) {
          double dsy;
          const double sy = sigma_y(alpha + sqrt_23 * dgam, dsy);
          m = z_mag - 2.0 * Ie * dgam;
          const Mat<double, DIM> Amat = add_diag(scale(m, nhat), Ie);
          const double g1 = mu * m - sqrt_23 * sy;
          const double g2 = det(Amat) - 1.0;
          if (fabs(g1) < 1e-12 * mu && fabs(g2) < 1e-12) {
            if (polish) break;
            polish = true;
          }
          const Mat<double, DIM> adj = cofactor_T(Amat);
          double dg2_dm = 0.0, tr_adj = 0.0;
#pragma unroll
          for (int i = 0; i < DIM; ++i) {
            tr_adj += adj(i, i);
#pragma unroll
            for (int j = 0; j < DIM; ++j) dg2_dm += adj(j, i) * nhat(i, j);
          }
          // Jacobian of (g1, g2) w.r.t. (dgam, Ie); dm/ddgam = -2 Ie, dm/dIe = -2 dgam
          const double a11 = -2.0 * mu * Ie - (2.0 / 3.0) * dsy, a12 = -2.0 * mu * dgam;
          const double a21 = -2.0 * Ie * dg2_dm, a22 = -2.0 * dgam * dg2_dm + tr_adj;
          const double dt = a11 * a22 - a12 * a21;
          if (dt == 0.0) break;
          double d_dgam = (-g1 * a22 + g2 * a12) / dt;
          const double d_Ie = (-a11 * g2 + a21 * g1) / dt;
          if (dgam + d_dgam < 0.0) d_dgam = -0.5 * dgam;   // stay on the loading side
          dgam += d_dgam;
          Ie += d_Ie;
        }
        m = z_mag - 2.0 * Ie * dgam;
        z = scale(m, nhat);
        alpha += sqrt_23 * dgam;
      }
    }
    pack_sym<double, DIM>(z, xi);
    xi[NS] = Ie;
    xi[NS + 1] = alpha;
    return r0;
  }
  template <class TK, class TKP, class TX, class TXP, class TP>
  static C8_DI int residual(const Kin<DIM, TK, TKP>& k, const TX* xi, const TXP* xip,
                            const TP* par, double abs_tol, prom5_t<TK, TKP, TX, TXP, TP>* C) {
    using R = prom5_t<TK, TKP, TX, TXP, TP>;
    const double sqrt_23 = 0.81649658092772603, sqrt_32 = 1.2247448713915890;
    const TP mu = mu_of(par[0], par[1]);
    const Mat<TXP, DIM> zo = unpack_sym<TXP, DIM>(xip);
    const auto bt = be_bar_trial<DIM>(k, zo, xip[NS]);
    const auto dbt = dev(bt);
    const Mat<TX, DIM> zeta = unpack_sym<TX, DIM>(xi);
    const TX Ie = xi[NS], alpha = xi[NS + 1];
    const auto s = scale(mu, zeta);
    const auto s_mag = norm(s);
    const auto sy = hyper_yield<TX, TP>(alpha, par[2], par[3], par[4], par[5], par[6], par[7]);
    const auto f = (s_mag - sqrt_23 * sy) / val(mu);
    if (is_plastic(val(f), abs_tol)) {
      const auto dgam = sqrt_32 * (alpha - xip[NS + 1]);
      const auto g = (2.0 * dgam * Ie) / s_mag;
#pragma unroll
      for (int i = 0; i < DIM; ++i)
#pragma unroll
        for (int j = i; j < DIM; ++j) {
          const int q = SymIdx<DIM>::idx(i, j);
          C[q] = conv<R>(zeta.a[i][j] - dbt.a[i][j] + g * s.a[i][j]);
        }
      C[NS] = conv<R>(det(add_diag(zeta, Ie)) - 1.0);
      C[NS + 1] = conv<R>(f);
      return PATH_PLASTIC;
    }
#pragma unroll
    for (int i = 0; i < DIM; ++i)
#pragma unroll
      for (int j = i; j < DIM; ++j) {
        const int q = SymIdx<DIM>::idx(i, j);
        C[q] = conv<R>(zeta.a[i][j] - dbt.a[i][j]);
      }
    C[NS] = conv<R>(Ie - trace(bt) / 3.0);
    C[NS + 1] = conv<R>(alpha - xip[NS + 1]);
    return PATH_ELASTIC;
  }
  template <class TK, class TKP, class TX, class TP>
  static C8_DI Mat<prom3_t<TK, TX, TP>, DIM> dev_cauchy(const Kin<DIM, TK, TKP>& k, const TX* xi,
                                                        const TP* par) {
    const TP mu = mu_of(par[0], par[1]);
    const TK J = det(add_diag(k.gu, 1.0));
    const Mat<TX, DIM> zeta = unpack_sym<TX, DIM>(xi);
    return mat_conv<prom3_t<TK, TX, TP>>(scale(mu / J, zeta));
  }
  template <class TK, class TKP, class TX, class TP>
  static C8_DI prom3_t<TK, TX, TP> hydro(const Kin<DIM, TK, TKP>& k, const TX*, const TP* par) {
    const TP kappa = kappa_of(par[0], par[1]);
    const TK J = det(add_diag(k.gu, 1.0));
    return conv<prom3_t<TK, TX, TP>>(kappa / 2.0 * (J - 1.0 / J));
  }
  template <class TP> static C8_DI TP pscale(const TP* par) { return kappa_of(par[0], par[1]); }
};

// =============================================================================
template <int DIM>
struct HyperJ2PlaneStrain {
  static_assert(DIM == 2, "plane strain is 2-D");
  static constexpr int NS = 3, NXI = 5, NPAR = 6, TYPE = L_HYPER_J2_PLANE_STRAIN;
  static constexpr bool FINITE = true, HAS_NEWTON = true, PLANE_STRESS = false;
  // on the elastic branch C = xi - (a function of xi_prev and the kinematics): dC/dxi = I
  static constexpr bool ELASTIC_J_IDENTITY = false;
  static constexpr int Z_STRETCH = -1;
  static C8_DI void init(double* xi) {
    xi[0] = xi[1] = xi[2] = 0.0; xi[3] = 1.0; xi[4] = 0.0;
  }
  // 3-D be_bar with zz = (-tr zeta_old + Ie_old)/cbrt(det rF)^2, src/hyper_J2_plane_strain.cpp:129-151
  template <class TK, class TKP, class TXP>
  static C8_DI void trial(const Kin<2, TK, TKP>& k, const TXP* xip,
                          Mat<prom3_t<TK, TKP, TXP>, 2>& zeta_trial, prom3_t<TK, TKP, TXP>& Ie_trial) {
    const Mat<TXP, 2> zo = unpack_sym<TXP, 2>(xip);
    prom_t<TK, TKP> d13;
    const auto b2 = be_bar_trial<2>(k, zo, xip[3], &d13);
    const auto bzz = (-(xip[0] + xip[2]) + xip[3]) / (d13 * d13);
    Ie_trial = (b2(0, 0) + b2(1, 1) + bzz) / 3.0;
    zeta_trial = b2;
    zeta_trial(0, 0) -= Ie_trial; zeta_trial(1, 1) -= Ie_trial;
  }
  template <class K> static C8_DI void guess(const K& k, const double* xip, const double* par, double abs_tol,
                                                  double* xi) {
    Mat<double, 2> zt; double Iet;
    trial(k, xip, zt, Iet);
    pack_sym<double, 2>(zt, xi);
    xi[3] = Iet; xi[4] = xip[4];
  }
  template <class TK, class TKP, class TX, class TXP, class TP>
  static C8_DI int residual(const Kin<2, TK, TKP>& k, const TX* xi, const TXP* xip, const TP* par,
                            double abs_tol, prom5_t<TK, TKP, TX, TXP, TP>* C) {
    using R = prom5_t<TK, TKP, TX, TXP, TP>;
    const double sqrt_23 = 0.81649658092772603, sqrt_32 = 1.2247448713915890;
    const TP mu = mu_of(par[0], par[1]);
    Mat<prom3_t<TK, TKP, TXP>, 2> zt; prom3_t<TK, TKP, TXP> Iet;
    trial(k, xip, zt, Iet);
    const Mat<TX, 2> zeta = unpack_sym<TX, 2>(xi);
    const TX Ie = xi[3], alpha = xi[4];
    Mat<TX, 3> z3 = embed3(zeta);
    z3(2, 2) = -(xi[0] + xi[2]);
    const auto s_mag = norm(scale(mu, z3));
    // sigma_y = Y + K alpha + (Y_inf - Y)(1 - exp(-delta alpha)), :269-270 ; params E,nu,K,Y,Y_inf,delta
    const auto sy = par[3] + par[2] * alpha + (par[4] - par[3]) * (1.0 - dexp(-par[5] * alpha));
    const auto f = (s_mag - sqrt_23 * sy) / val(mu);
    if (is_plastic(val(f), abs_tol)) {
      const auto dgam = sqrt_32 * (alpha - xip[4]);
      const auto g = (2.0 * dgam * Ie) * mu / s_mag;
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = i; j < 2; ++j) {
          const int q = SymIdx<2>::idx(i, j);
          C[q] = conv<R>(zeta.a[i][j] - zt.a[i][j] + g * zeta.a[i][j]);
        }
      C[3] = conv<R>(det(add_diag(z3, Ie)) - 1.0);
      C[4] = conv<R>(f);
      return PATH_PLASTIC;
    }
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = i; j < 2; ++j) {
        const int q = SymIdx<2>::idx(i, j);
        C[q] = conv<R>(zeta.a[i][j] - zt.a[i][j]);
      }
    C[3] = conv<R>(Ie - Iet);
    C[4] = conv<R>(alpha - xip[4]);
    return PATH_ELASTIC;
  }
  template <class TK, class TKP, class TX, class TP>
  static C8_DI Mat<prom3_t<TK, TX, TP>, 2> dev_cauchy(const Kin<2, TK, TKP>& k, const TX* xi,
                                                      const TP* par) {
    const TP mu = mu_of(par[0], par[1]);
    const TK J = det(add_diag(k.gu, 1.0));
    return mat_conv<prom3_t<TK, TX, TP>>(scale(mu / J, unpack_sym<TX, 2>(xi)));
  }
  template <class TK, class TKP, class TX, class TP>
  static C8_DI prom3_t<TK, TX, TP> hydro(const Kin<2, TK, TKP>& k, const TX*, const TP* par) {
    const TP kappa = kappa_of(par[0], par[1]);
    const TK J = det(add_diag(k.gu, 1.0));
    return conv<prom3_t<TK, TX, TP>>(kappa / 2.0 * (J - 1.0 / J));
  }
  template <class TP> static C8_DI TP pscale(const TP* par) { return kappa_of(par[0], par[1]); }
};

// =============================================================================
template <int DIM>
struct HyperJ2PlaneStress {
  static_assert(DIM == 2, "plane stress is 2-D");
  // xi = zeta(00,01,11), Ie, lambda_z, alpha
  static constexpr int NS = 3, NXI = 6, NPAR = 8, TYPE = L_HYPER_J2_PLANE_STRESS;
  static constexpr bool FINITE = true, HAS_NEWTON = true, PLANE_STRESS = true;
  // on the elastic branch C = xi - (a function of xi_prev and the kinematics): dC/dxi = I
  static constexpr bool ELASTIC_J_IDENTITY = false;
  static constexpr int Z_STRETCH = 4;  // packed index of lambda_z (residual index 2 in the reference)
  static C8_DI void init(double* xi) {
    xi[0] = xi[1] = xi[2] = 0.0; xi[3] = 1.0; xi[4] = 1.0; xi[5] = 0.0;
  }
  // src/hyper_J2_plane_stress.cpp:140-169 ; lambda_z is the CURRENT z stretch
  template <class TK, class TKP, class TX, class TXP>
  static C8_DI void trial(const Kin<2, TK, TKP>& k, const TXP* xip, const TX& lambda_z,
                          Mat<prom4_t<TK, TKP, TX, TXP>, 2>& zeta_trial,
                          prom4_t<TK, TKP, TX, TXP>& Ie_trial, TK& J2) {
    using R = prom4_t<TK, TKP, TX, TXP>;
    const Mat<TK, 2> F2 = add_diag(k.gu, 1.0);
    J2 = det(F2);
    const Mat<TKP, 2> Fp2 = add_diag(k.gup, 1.0);
    Mat<prom_t<TK, TX>, 3> F3 = mat_conv<prom_t<TK, TX>>(embed3(F2));
    F3(2, 2) = conv<prom_t<TK, TX>>(lambda_z);
    Mat<prom_t<TKP, TXP>, 3> Fp3 = mat_conv<prom_t<TKP, TXP>>(embed3(Fp2));
    Fp3(2, 2) = conv<prom_t<TKP, TXP>>(xip[4]);
    const auto rF = F3 * inverse(Fp3);
    const auto d13 = dcbrt(det(rF));
    const auto rFb = scale(1.0 / d13, rF);
    Mat<TXP, 3> core = embed3(unpack_sym<TXP, 2>(xip));
    core(2, 2) = -(xip[0] + xip[2]);
    const Mat<TXP, 3> corei = add_diag(core, xip[3]);
    const Mat<R, 3> bt = mat_conv<R>(rFb * corei * transpose(rFb));
    Ie_trial = trace(bt) / 3.0;
    zeta_trial = extract2(bt);
    zeta_trial(0, 0) -= Ie_trial; zeta_trial(1, 1) -= Ie_trial;
  }
  // initial guess: zeta, Ie from the trial state evaluated with the CURRENT-field lambda_z;
  // lambda_z and alpha keep the values gathered from the current xi field (:176-200)
  template <class K> static C8_DI void guess(const K& k, const double* xip, const double* par, double abs_tol,
                                                  double* xi) {
    Mat<double, 2> zt; double Iet, J2;
    trial(k, xip, xi[4], zt, Iet, J2);
    pack_sym<double, 2>(zt, xi);
    xi[3] = Iet;
  }
  template <class TK, class TKP, class TX, class TXP, class TP>
  static C8_DI int residual(const Kin<2, TK, TKP>& k, const TX* xi, const TXP* xip, const TP* par,
                            double abs_tol, prom5_t<TK, TKP, TX, TXP, TP>* C) {
    using R = prom5_t<TK, TKP, TX, TXP, TP>;
    const double sqrt_23 = 0.81649658092772603, sqrt_32 = 1.2247448713915890;
    const TP mu = mu_of(par[0], par[1]);
    const TP kappa = kappa_of(par[0], par[1]);
    const TX Ie = xi[3], lambda_z = xi[4], alpha = xi[5];
    Mat<prom4_t<TK, TKP, TX, TXP>, 2> zt; prom4_t<TK, TKP, TX, TXP> Iet; TK J2;
    trial(k, xip, lambda_z, zt, Iet, J2);
    const Mat<TX, 2> zeta = unpack_sym<TX, 2>(xi);
    Mat<TX, 3> z3 = embed3(zeta);
    const TX zeta_zz = -(xi[0] + xi[2]);
    z3(2, 2) = zeta_zz;
    const auto s_mag = norm(scale(mu, z3));
    const auto sy = hyper_yield<TX, TP>(alpha, par[2], par[3], par[4], par[5], par[6], par[7]);
    const auto f = (s_mag - sqrt_23 * sy) / val(mu);
    const TP mat_factor = kappa / (2.0 * mu);
    // R_lambda_z = lambda_z - sqrt((1 - zeta_zz / (kappa/2mu)) / J2^2), :310-312
    C[4] = conv<R>(lambda_z - dsqrt((1.0 - zeta_zz / mat_factor) / sqr(J2)));
    if (is_plastic(val(f), abs_tol)) {
      const auto dgam = sqrt_32 * (alpha - xip[5]);
      const auto g = (2.0 * dgam * Ie) * mu / s_mag;
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = i; j < 2; ++j) {
          const int q = SymIdx<2>::idx(i, j);
          C[q] = conv<R>(zeta.a[i][j] - zt.a[i][j] + g * zeta.a[i][j]);
        }
      C[3] = conv<R>(det(add_diag(z3, Ie)) - 1.0);
      C[5] = conv<R>(f);
      return PATH_PLASTIC;
    }
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = i; j < 2; ++j) {
        const int q = SymIdx<2>::idx(i, j);
        C[q] = conv<R>(zeta.a[i][j] - zt.a[i][j]);
      }
    C[3] = conv<R>(Ie - Iet);
    C[5] = conv<R>(alpha - xip[5]);
    return PATH_ELASTIC;
  }
  // sigma = mu zeta / J + kappa/2 (J - 1/J) I, J = det F2 * lambda_z, :361-375
  template <class TK, class TKP, class TX, class TP>
  static C8_DI Mat<prom3_t<TK, TX, TP>, 2> cauchy(const Kin<2, TK, TKP>& k, const TX* xi,
                                                  const TP* par) {
    using R = prom3_t<TK, TX, TP>;
    const TP mu = mu_of(par[0], par[1]);
    const TP kappa = kappa_of(par[0], par[1]);
    const auto J = det(add_diag(k.gu, 1.0)) * xi[4];
    Mat<R, 2> sig = mat_conv<R>(scale(mu / J, unpack_sym<TX, 2>(xi)));
    const R h = kappa / 2.0 * (J - 1.0 / J);
    sig(0, 0) += h; sig(1, 1) += h;
    return sig;
  }
  template <class TP> static C8_DI TP pscale(const TP*) { return conv<TP>(0.0); }
};

// =============================================================================
// Hypoelastic Hill plasticity in the unrotated configuration: the local state is the unrotated Cauchy
// stress TC (+ alpha), driven by the unrotated rate of deformation
//   d = R^T sym((F - F_prev) F^-1) R,   R = polar_rotation(F)          src/hypo_kinematics.hpp:10-17
// and the Cauchy stress handed to the mechanics residual is R TC R^T.
template <int DIM, class TK, class TKP>
C8_DI Mat<TK, DIM> kin_rotation(const Kin<DIM, TK, TKP>& k) {
  if (k.has_rot) return k.rot;
  return polar_rotation(add_diag(k.gu, 1.0));
}
// fill the cache (kernels call this once per AD pass for the models that need it)
template <int DIM, class TK, class TKP>
C8_DI void cache_rotation(Kin<DIM, TK, TKP>& k) {
  k.rot = polar_rotation(add_diag(k.gu, 1.0));
  k.has_rot = true;
}
template <int DIM, class TK, class TKP>
C8_DI Mat<prom_t<TK, TKP>, DIM> unrotated_rate(const Kin<DIM, TK, TKP>& k) {
  const Mat<TK, DIM> F = add_diag(k.gu, 1.0);
  const Mat<TKP, DIM> Fp = add_diag(k.gup, 1.0);
  const Mat<TK, DIM> Fi = inverse(F);
  const Mat<TK, DIM> R = kin_rotation(k);
  const auto Lv = (F - Fp) * Fi;
  const auto Dm = scale(0.5, Lv + transpose(Lv));
  return transpose(R) * Dm * R;
}
template <int DIM, class TK, class TKP, class TX>
C8_DI Mat<prom_t<TK, TX>, DIM> rotate_forward(const Kin<DIM, TK, TKP>& k, const Mat<TX, DIM>& TC) {
  const Mat<TK, DIM> R = kin_rotation(k);
  return R * TC * transpose(R);
}

template <int DIM>
struct HypoHill {
  static_assert(DIM == 3, "hypo_hill is a 3-D model");
  static constexpr int NS = 6, NXI = 7, NPAR = 11, TYPE = L_HYPO_HILL;
  static constexpr bool FINITE = true, HAS_NEWTON = true, PLANE_STRESS = false;
  static constexpr bool ELASTIC_J_IDENTITY = false;   // the TC rows are scaled by 1 / mu
  static constexpr int Z_STRETCH = -1;
  static constexpr bool NEEDS_ROTATION = true;
  static C8_DI void init(double* xi) {
#pragma unroll
    for (int i = 0; i < NXI; ++i) xi[i] = 0.0;
  }
  // elastic predictor, src/hypo_hill.cpp:163-177; a yielding point continues to the solution of
  // hill_return_map() (TC = TC_tr - 2 mu dgam n(TC): the same structure as the small-strain model, the
  // hydrostatic part of TC passes through M untouched).  R_norm_0 of the relative test is the norm at the
  // reference's start, the elastic predictor: R_TC = 0 there and R_alpha = f_trial.
  static constexpr bool HAS_PREDICTOR = true;
  template <class K> static C8_DI void guess(const K& k, const double* xip, const double* par, double abs_tol,
                                             double* xi) {
    (void)guess_r0(k, xip, par, abs_tol, xi);
  }
  template <class K> static C8_DI double guess_r0(const K& k, const double* xip, const double* par,
                                                  double abs_tol, double* xi) {
    const double lambda = lambda_of(par[0], par[1]), mu = mu_of(par[0], par[1]);
    const Mat<double, 3> d = unrotated_rate<3>(k);
    const double ltd = lambda * trace(d);
    Mat<double, 3> T_tr;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j)
        T_tr(i, j) = xip[SymIdx<3>::idx(i, j)] + (i == j ? ltd : 0.0) + 2.0 * mu * d.a[i][j];
    pack_sym<double, 3>(T_tr, xi);
    xi[NS] = xip[NS];
    const Hill<double> hp = hill_params(par[3], par[4], par[5], par[6], par[7], par[8]);
    const double hill_tr = hill_value(T_tr, hp);
    const double f = (hill_tr - (par[2] + par[9] * (1.0 - exp(-par[10] * xip[NS])))) / mu;
    if (!(is_plastic(f, abs_tol) && hill_tr > 0.0)) return -1.0;
    Mat<double, 3> T; double dgam;
    if (hill_return_map(T_tr, hp, mu, par[2], par[9], par[10], xip[NS], T, dgam)) {
      pack_sym<double, 3>(T, xi);
      xi[NS] = xip[NS] + dgam;
    }
    return fabs(f);
  }
  template <class TK, class TKP, class TX, class TXP, class TP>
  static C8_DI int residual(const Kin<3, TK, TKP>& k, const TX* xi, const TXP* xip, const TP* par,
                            double abs_tol, prom5_t<TK, TKP, TX, TXP, TP>* C) {
    using R = prom5_t<TK, TKP, TX, TXP, TP>;
    const TP mu = mu_of(par[0], par[1]), lambda = lambda_of(par[0], par[1]);
    const double imu = 1.0 / val(mu);
    const Hill<TP> hp = hill_params(par[3], par[4], par[5], par[6], par[7], par[8]);
    const Mat<TX, 3> TC = unpack_sym<TX, 3>(xi);
    const auto hill = hill_value(TC, hp);
    const auto sigma_yield = par[2] + par[9] * (1.0 - dexp(-par[10] * xi[NS]));
    const auto f = (hill - sigma_yield) * imu;
    const auto d = unrotated_rate<3>(k);
    const auto ltd = lambda * trace(d);
    const bool plastic = is_plastic(val(f), abs_tol);
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = i; j < 3; ++j) {
        const int q = SymIdx<3>::idx(i, j);
        R r = conv<R>(xi[q] - xip[q] - 2.0 * mu * d.a[i][j]);
        if (i == j) r -= ltd;
        C[q] = r * imu;
      }
    if (plastic) {
      const auto n = hill_normal(TC, hp, hill);
      const auto dgam = xi[NS] - xip[NS];
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = i; j < 3; ++j) {
          const int q = SymIdx<3>::idx(i, j);
          C[q] += conv<R>((2.0 * mu * dgam * n.a[i][j]) * imu);
        }
      C[NS] = conv<R>(f);
      return PATH_PLASTIC;
    }
    C[NS] = conv<R>(xi[NS] - xip[NS]);
    return PATH_ELASTIC;
  }
  template <class TK, class TKP, class TX, class TP>
  static C8_DI Mat<prom3_t<TK, TX, TP>, 3> dev_cauchy(const Kin<3, TK, TKP>& k, const TX* xi, const TP*) {
    return mat_conv<prom3_t<TK, TX, TP>>(dev(rotate_forward<3>(k, unpack_sym<TX, 3>(xi))));
  }
  template <class TK, class TKP, class TX, class TP>
  static C8_DI prom3_t<TK, TX, TP> hydro(const Kin<3, TK, TKP>& k, const TX* xi, const TP*) {
    return conv<prom3_t<TK, TX, TP>>(trace(rotate_forward<3>(k, unpack_sym<TX, 3>(xi))) / 3.0);
  }
  template <class TP> static C8_DI TP pscale(const TP* par) { return kappa_of(par[0], par[1]); }
};

// xi = TC(00,01,11), alpha, TC_zz ; params E, nu, Y, S, D, R00, R11, R22, R01 (R02 = R12 = 1)
template <int DIM>
struct HypoHillPlaneStrain {
  static_assert(DIM == 2, "plane strain is 2-D");
  static constexpr int NS = 3, NXI = 5, NPAR = 9, TYPE = L_HYPO_HILL_PLANE_STRAIN;
  static constexpr bool FINITE = true, HAS_NEWTON = true, PLANE_STRESS = false;
  static constexpr bool ELASTIC_J_IDENTITY = true;    // no 1 / mu scaling in this variant
  static constexpr int Z_STRETCH = -1;
  static constexpr bool NEEDS_ROTATION = true;
  static C8_DI void init(double* xi) {
#pragma unroll
    for (int i = 0; i < NXI; ++i) xi[i] = 0.0;
  }
  template <class K> static C8_DI void guess(const K& k, const double* xip, const double* par, double, double* xi) {
    const double lambda = lambda_of(par[0], par[1]), mu = mu_of(par[0], par[1]);
    const Mat<double, 2> d = unrotated_rate<2>(k);
    const double ltd = lambda * trace(d);
    xi[0] = xip[0] + ltd + 2.0 * mu * d(0, 0);
    xi[1] = xip[1] + 2.0 * mu * d(0, 1);
    xi[2] = xip[2] + ltd + 2.0 * mu * d(1, 1);
    xi[3] = xip[3];
    xi[4] = xip[4] + ltd;
  }
  template <class TK, class TKP, class TX, class TXP, class TP>
  static C8_DI int residual(const Kin<2, TK, TKP>& k, const TX* xi, const TXP* xip, const TP* par,
                            double abs_tol, prom5_t<TK, TKP, TX, TXP, TP>* C) {
    using R = prom5_t<TK, TKP, TX, TXP, TP>;
    const TP mu = mu_of(par[0], par[1]), lambda = lambda_of(par[0], par[1]);
    const TP one = conv<TP>(1.0);
    const Hill<TP> hp = hill_params(par[5], par[6], par[7], par[8], one, one);
    Mat<TX, 3> TC3 = embed3(unpack_sym<TX, 2>(xi));
    TC3(2, 2) = xi[4];
    const auto phi = hill_value(TC3, hp);
    const auto sigma_yield = par[2] + par[3] * (1.0 - dexp(-par[4] * xi[3]));
    const auto f = (phi - sigma_yield) / val(mu);
    const auto d = unrotated_rate<2>(k);
    const auto ltd = lambda * trace(d);
    C[0] = conv<R>(xi[0] - xip[0] - ltd - 2.0 * mu * d(0, 0));
    C[1] = conv<R>(xi[1] - xip[1] - 2.0 * mu * d(0, 1));
    C[2] = conv<R>(xi[2] - xip[2] - ltd - 2.0 * mu * d(1, 1));
    C[4] = conv<R>(xi[4] - xip[4] - ltd);
    if (is_plastic(val(f), abs_tol)) {
      const auto n = hill_normal(TC3, hp, phi);
      const auto dgam = xi[3] - xip[3];
      const auto tm = 2.0 * mu * dgam;
      C[0] += conv<R>(tm * n(0, 0));
      C[1] += conv<R>(tm * n(0, 1));
      C[2] += conv<R>(tm * n(1, 1));
      C[4] += conv<R>(tm * (-(n(0, 0) + n(1, 1))));
      C[3] = conv<R>(f);
      return PATH_PLASTIC;
    }
    C[3] = conv<R>(xi[3] - xip[3]);
    return PATH_ELASTIC;
  }
  // hydro = (tr(R TC R^T) + TC_zz) / 3 ; dev = R TC R^T - hydro I, src/hypo_hill_plane_strain.cpp:351-371
  template <class TK, class TKP, class TX, class TP>
  static C8_DI Mat<prom3_t<TK, TX, TP>, 2> dev_cauchy(const Kin<2, TK, TKP>& k, const TX* xi, const TP*) {
    using R = prom3_t<TK, TX, TP>;
    Mat<R, 2> rc = mat_conv<R>(rotate_forward<2>(k, unpack_sym<TX, 2>(xi)));
    const R h = (trace(rc) + xi[4]) / 3.0;
    rc(0, 0) -= h; rc(1, 1) -= h;
    return rc;
  }
  template <class TK, class TKP, class TX, class TP>
  static C8_DI prom3_t<TK, TX, TP> hydro(const Kin<2, TK, TKP>& k, const TX* xi, const TP*) {
    using R = prom3_t<TK, TX, TP>;
    const Mat<R, 2> rc = mat_conv<R>(rotate_forward<2>(k, unpack_sym<TX, 2>(xi)));
    return (trace(rc) + xi[4]) / 3.0;
  }
  template <class TP> static C8_DI TP pscale(const TP* par) { return kappa_of(par[0], par[1]); }
};

// xi = TC(00,01,11), alpha, lambda_z ; params E, nu, Y, S, D, R00, R11, R22, R01, Q00, Q01, Q10, Q11
template <int DIM>
struct HypoHillPlaneStress {
  static_assert(DIM == 2, "plane stress is 2-D");
  static constexpr int NS = 3, NXI = 5, NPAR = 13, TYPE = L_HYPO_HILL_PLANE_STRESS;
  static constexpr bool FINITE = true, HAS_NEWTON = true, PLANE_STRESS = true;
  static constexpr bool ELASTIC_J_IDENTITY = true;    // elastic branch: unscaled rows, C = xi - g(xi_prev, F)
  static constexpr int Z_STRETCH = 4;
  static constexpr bool NEEDS_ROTATION = true;
  static C8_DI void init(double* xi) { xi[0] = xi[1] = xi[2] = xi[3] = 0.0; xi[4] = 1.0; }
  template <class TP> static C8_DI Mat<TP, 2> frame(const TP* par) {
    Mat<TP, 2> Q;
    Q(0, 0) = par[9]; Q(0, 1) = par[10]; Q(1, 0) = par[11]; Q(1, 1) = par[12];
    return Q;
  }
  // d = Q^T R^T D R Q, src/hypo_hill_plane_stress.cpp:165-179
  template <class TK, class TKP, class TP>
  static C8_DI Mat<prom3_t<TK, TKP, TP>, 2> eval_d(const Kin<2, TK, TKP>& k, const TP* par) {
    const Mat<TP, 2> Q = frame(par);
    return mat_conv<prom3_t<TK, TKP, TP>>(transpose(Q) * unrotated_rate<2>(k) * Q);
  }
  template <class K> static C8_DI void guess(const K& k, const double* xip, const double* par, double, double* xi) {
    const double lambda = lambda_of(par[0], par[1]), mu = mu_of(par[0], par[1]);
    const Mat<double, 2> d = eval_d(k, par);
    const double d_zz = -lambda * trace(d) / (lambda + 2.0 * mu);
    const double l = lambda * (trace(d) + d_zz);
    xi[0] = xip[0] + l + 2.0 * mu * d(0, 0);
    xi[1] = xip[1] + 2.0 * mu * d(0, 1);
    xi[2] = xip[2] + l + 2.0 * mu * d(1, 1);
    xi[3] = xip[3];
    xi[4] = xip[4] / (1.0 - d_zz);
  }
  template <class TK, class TKP, class TX, class TXP, class TP>
  static C8_DI int residual(const Kin<2, TK, TKP>& k, const TX* xi, const TXP* xip, const TP* par,
                            double abs_tol, prom5_t<TK, TKP, TX, TXP, TP>* C) {
    using R = prom5_t<TK, TKP, TX, TXP, TP>;
    const TP mu = mu_of(par[0], par[1]), lambda = lambda_of(par[0], par[1]);
    const TP one = conv<TP>(1.0);
    const Hill<TP> hp = hill_params(par[5], par[6], par[7], par[8], one, one);
    const Mat<TX, 3> TC3 = embed3(unpack_sym<TX, 2>(xi));
    const auto phi = hill_value(TC3, hp);
    const auto sigma_yield = par[2] + par[3] * (1.0 - dexp(-par[4] * xi[3]));
    const auto f = (phi - sigma_yield) / val(mu);
    const auto d = eval_d(k, par);
    const auto d_zz = -lambda * trace(d) / (lambda + 2.0 * mu);
    const auto l = lambda * (trace(d) + d_zz);
    C[0] = conv<R>(xi[0] - xip[0] - l - 2.0 * mu * d(0, 0));
    C[1] = conv<R>(xi[1] - xip[1] - 2.0 * mu * d(0, 1));
    C[2] = conv<R>(xi[2] - xip[2] - l - 2.0 * mu * d(1, 1));
    if (is_plastic(val(f), abs_tol)) {
      const auto n = hill_normal(TC3, hp, phi);
      const auto dgam = xi[3] - xip[3];
      const auto dp_zz = -(dgam * n(0, 0) + dgam * n(1, 1));
      const auto corr = 2.0 * mu * dp_zz / (2.0 * mu + lambda);   // return-map correction of d_zz
      const double imu = 1.0 / val(mu);
      // the (unforced) plastic branch alone scales the TC rows by 1 / val(mu), :321
      C[0] = conv<R>((C[0] + 2.0 * mu * (dgam * n(0, 0)) - lambda * corr) * imu);
      C[1] = conv<R>((C[1] + 2.0 * mu * (dgam * n(0, 1))) * imu);
      C[2] = conv<R>((C[2] + 2.0 * mu * (dgam * n(1, 1)) - lambda * corr) * imu);
      C[3] = conv<R>(f);
      C[4] = conv<R>(xi[4] - xip[4] / (1.0 - (d_zz + corr)));
      return PATH_PLASTIC;
    }
    C[3] = conv<R>(xi[3] - xip[3]);
    C[4] = conv<R>(xi[4] - xip[4] / (1.0 - d_zz));
    return PATH_ELASTIC;
  }
  // sigma = R Q TC Q^T R^T, :366-382
  template <class TK, class TKP, class TX, class TP>
  static C8_DI Mat<prom3_t<TK, TX, TP>, 2> cauchy(const Kin<2, TK, TKP>& k, const TX* xi, const TP* par) {
    using R = prom3_t<TK, TX, TP>;
    const Mat<TP, 2> Q = frame(par);
    const auto core = Q * unpack_sym<TX, 2>(xi) * transpose(Q);
    return mat_conv<R>(rotate_forward<2>(k, core));
  }
  template <class TP> static C8_DI TP pscale(const TP*) { return conv<TP>(0.0); }
};

}  // namespace c8
