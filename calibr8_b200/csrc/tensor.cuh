// Fixed-size 2x2 / 3x3 tensors of double or Dual<L>, register-resident.
// Replaces the run-time-sized minitensor::Tensor<T> (reference: src/defines.hpp:33-37);
// only the operations the hot path uses: + - * scalar ops, transpose, trace,
// det, inverse (cofactor formulas), dev (A - tr A / N), Frobenius norm.
#pragma once
#include "dual.cuh"

namespace c8 {

template <class T, int N>
struct Mat {
  T a[N][N];
  C8_DI T& operator()(int i, int j) { return a[i][j]; }
  C8_DI const T& operator()(int i, int j) const { return a[i][j]; }
};

template <class T, int N> C8_DI Mat<T, N> mat_zero() {
  Mat<T, N> r;
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int j = 0; j < N; ++j) r.a[i][j] = conv<T>(0.0);
  return r;
}
template <class T, class S, int N> C8_DI Mat<T, N> mat_conv(const Mat<S, N>& x) {
  Mat<T, N> r;
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int j = 0; j < N; ++j) r.a[i][j] = conv<T>(x.a[i][j]);
  return r;
}

template <class A, class B, int N>
C8_DI Mat<prom_t<A, B>, N> operator+(const Mat<A, N>& x, const Mat<B, N>& y) {
  Mat<prom_t<A, B>, N> r;
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int j = 0; j < N; ++j) r.a[i][j] = x.a[i][j] + y.a[i][j];
  return r;
}
template <class A, class B, int N>
C8_DI Mat<prom_t<A, B>, N> operator-(const Mat<A, N>& x, const Mat<B, N>& y) {
  Mat<prom_t<A, B>, N> r;
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int j = 0; j < N; ++j) r.a[i][j] = x.a[i][j] - y.a[i][j];
  return r;
}
template <class A, class B, int N>
C8_DI Mat<prom_t<A, B>, N> operator*(const Mat<A, N>& x, const Mat<B, N>& y) {
  Mat<prom_t<A, B>, N> r;
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int j = 0; j < N; ++j) {
      prom_t<A, B> s = x.a[i][0] * y.a[0][j];
#pragma unroll
      for (int k = 1; k < N; ++k) s += x.a[i][k] * y.a[k][j];
      r.a[i][j] = s;
    }
  return r;
}
// scalar * tensor, tensor * scalar, tensor / scalar  (scalar: double or Dual)
template <class S, class B, int N>
C8_DI Mat<prom_t<S, B>, N> scale(const S& s, const Mat<B, N>& y) {
  Mat<prom_t<S, B>, N> r;
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int j = 0; j < N; ++j) r.a[i][j] = s * y.a[i][j];
  return r;
}
template <class B, int N> C8_DI auto operator*(double s, const Mat<B, N>& y) { return scale(s, y); }
template <int L, class B, int N> C8_DI auto operator*(const Dual<L>& s, const Mat<B, N>& y) { return scale(s, y); }
template <class S, class B, int N>
C8_DI Mat<prom_t<S, B>, N> divide(const Mat<B, N>& y, const S& s) {
  Mat<prom_t<S, B>, N> r;
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int j = 0; j < N; ++j) r.a[i][j] = y.a[i][j] / s;
  return r;
}
template <class B, int N> C8_DI auto operator/(const Mat<B, N>& y, double s) { return scale(1.0 / s, y); }
template <int L, class B, int N> C8_DI auto operator/(const Mat<B, N>& y, const Dual<L>& s) {
  return scale(1.0 / s, y);
}

template <class T, int N> C8_DI Mat<T, N> transpose(const Mat<T, N>& x) {
  Mat<T, N> r;
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int j = 0; j < N; ++j) r.a[i][j] = x.a[j][i];
  return r;
}
template <class T, int N> C8_DI T trace(const Mat<T, N>& x) {
  T s = x.a[0][0];
#pragma unroll
  for (int i = 1; i < N; ++i) s += x.a[i][i];
  return s;
}
template <class T> C8_DI T det(const Mat<T, 2>& A) { return A(0, 0) * A(1, 1) - A(1, 0) * A(0, 1); }
template <class T> C8_DI T det(const Mat<T, 3>& A) {
  return A(0, 0) * (A(1, 1) * A(2, 2) - A(1, 2) * A(2, 1)) -
         A(0, 1) * (A(1, 0) * A(2, 2) - A(1, 2) * A(2, 0)) +
         A(0, 2) * (A(1, 0) * A(2, 1) - A(1, 1) * A(2, 0));
}
template <class T> C8_DI Mat<T, 2> inverse(const Mat<T, 2>& A) {
  const T id = 1.0 / det(A);
  Mat<T, 2> r;
  r(0, 0) = A(1, 1) * id; r(0, 1) = -A(0, 1) * id;
  r(1, 0) = -A(1, 0) * id; r(1, 1) = A(0, 0) * id;
  return r;
}
template <class T> C8_DI Mat<T, 3> cofactor_T(const Mat<T, 3>& A) {  // adjugate
  Mat<T, 3> r;
  r(0, 0) = A(1, 1) * A(2, 2) - A(1, 2) * A(2, 1);
  r(0, 1) = A(0, 2) * A(2, 1) - A(0, 1) * A(2, 2);
  r(0, 2) = A(0, 1) * A(1, 2) - A(0, 2) * A(1, 1);
  r(1, 0) = A(1, 2) * A(2, 0) - A(1, 0) * A(2, 2);
  r(1, 1) = A(0, 0) * A(2, 2) - A(0, 2) * A(2, 0);
  r(1, 2) = A(0, 2) * A(1, 0) - A(0, 0) * A(1, 2);
  r(2, 0) = A(1, 0) * A(2, 1) - A(1, 1) * A(2, 0);
  r(2, 1) = A(0, 1) * A(2, 0) - A(0, 0) * A(2, 1);
  r(2, 2) = A(0, 0) * A(1, 1) - A(0, 1) * A(1, 0);
  return r;
}
template <class T> C8_DI Mat<T, 3> inverse(const Mat<T, 3>& A) {
  const Mat<T, 3> adj = cofactor_T(A);
  const T d = A(0, 0) * adj(0, 0) + A(0, 1) * adj(1, 0) + A(0, 2) * adj(2, 0);
  return scale(1.0 / d, adj);
}
// dev(A) = A - tr(A)/N I, N the tensor dimension (MiniTensor's definition)
template <class T, int N> C8_DI Mat<T, N> dev(const Mat<T, N>& A) {
  const T theta = trace(A) * (1.0 / N);
  Mat<T, N> r = A;
#pragma unroll
  for (int i = 0; i < N; ++i) r.a[i][i] = A.a[i][i] - theta;
  return r;
}
template <class T, int N> C8_DI T frob2(const Mat<T, N>& A) {
  T s = A.a[0][0] * A.a[0][0];
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int j = 0; j < N; ++j)
      if (i + j > 0) s += A.a[i][j] * A.a[i][j];
  return s;
}
template <class T, int N> C8_DI T norm(const Mat<T, N>& A) { return dsqrt(frob2(A)); }

template <class T, int N, class S> C8_DI Mat<T, N> add_diag(const Mat<T, N>& A, const S& s) {
  Mat<T, N> r = A;
#pragma unroll
  for (int i = 0; i < N; ++i) r.a[i][i] = A.a[i][i] + s;
  return r;
}

// ---- minitensor::polar_rotation (Trilinos MiniTensor, not vendored in the reference tree) ----------
// |x| and max(a, b) on the value part, derivatives of the selected operand (Sacado's semantics)
C8_DI double dabs(double a) { return fabs(a); }
template <int L> C8_DI Dual<L> dabs(const Dual<L>& a) { return a.v >= 0.0 ? a : -a; }
template <class T> C8_DI T tmax(const T& a, const T& b) { return val(a) >= val(b) ? a : b; }
// max absolute column sum / max absolute row sum
template <class T, int N> C8_DI T norm_1(const Mat<T, N>& A) {
  T best = dabs(A.a[0][0]);
#pragma unroll
  for (int i = 1; i < N; ++i) best += dabs(A.a[i][0]);
#pragma unroll
  for (int j = 1; j < N; ++j) {
    T s = dabs(A.a[0][j]);
#pragma unroll
    for (int i = 1; i < N; ++i) s += dabs(A.a[i][j]);
    best = tmax(best, s);
  }
  return best;
}
template <class T, int N> C8_DI T norm_infinity(const Mat<T, N>& A) {
  T best = dabs(A.a[0][0]);
#pragma unroll
  for (int j = 1; j < N; ++j) best += dabs(A.a[0][j]);
#pragma unroll
  for (int i = 1; i < N; ++i) {
    T s = dabs(A.a[i][0]);
#pragma unroll
    for (int j = 1; j < N; ++j) s += dabs(A.a[i][j]);
    best = tmax(best, s);
  }
  return best;
}
template <class T, int N> C8_DI double frob_val(const Mat<T, N>& A) {
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int j = 0; j < N; ++j) s += val(A.a[i][j]) * val(A.a[i][j]);
  return sqrt(s);
}
// Rotation R of the polar decomposition A = R U by Higham's scaled Newton iteration
//   X <- 1/2 (mu X + X^-T / mu),  mu = ((|Y|_1 |Y|_inf) / (|X|_1 |X|_inf))^(1/4),  Y = X^-1,
// scaling switched off once the relative change drops below 0.01, left when |Z - X|_F <= sqrt(sqrt(N) eps)
// or the change stops decreasing.  The reference evaluates it in its AD scalar (src/global_residual.hpp:
// 302-305, src/hypo_hill_plane_stress.cpp:174), so dR/dF is the derivative OF THE TRUNCATED ITERATION; the
// same iteration is therefore run here in Dual<L> (control flow on the value part, identical in every
// thread of a group), not replaced by the closed-form derivative of the exact rotation.
template <class T, int N> C8_DI Mat<T, N> polar_rotation(const Mat<T, N>& A) {
  bool scale = true;
  const double tol_scale = 0.01;
  const double sqrt_tol_conv = sqrt(sqrt(double(N)) * 2.220446049250313e-16);
  Mat<T, N> X = A;
  double gamma = 2.0;
#pragma unroll 1
  for (int it = 0; it < 128; ++it) {
    const Mat<T, N> Y = inverse(X);
    T mu = conv<T>(1.0);
    if (scale) mu = dsqrt(dsqrt((norm_1(Y) * norm_infinity(Y)) / (norm_1(X) * norm_infinity(X))));
    const T imu = 1.0 / mu;
    Mat<T, N> Z;
    double nD2 = 0.0, nZ2 = 0.0;
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
      for (int j = 0; j < N; ++j) {
        Z.a[i][j] = 0.5 * (mu * X.a[i][j] + Y.a[j][i] * imu);
        const double dz = val(Z.a[i][j]) - val(X.a[i][j]);
        nD2 += dz * dz;
        nZ2 += val(Z.a[i][j]) * val(Z.a[i][j]);
      }
    const double nD = sqrt(nD2), delta = nD / sqrt(nZ2);
    if (scale && delta < tol_scale) scale = false;
    const bool end_iter = nD <= sqrt_tol_conv || (delta > 0.5 * gamma && !scale);
    X = Z;
    gamma = delta;
    if (end_iter) break;
  }
  return X;
}

// packed symmetric storage order of the reference (src/local_residual.cpp:196-218):
// 3-D (00,01,02,11,12,22), 2-D (00,01,11)
template <int DIM> struct SymIdx;
template <> struct SymIdx<2> {
  static constexpr int n = 3;
  static C8_DI int idx(int i, int j) { return (i == j) ? (i == 0 ? 0 : 2) : 1; }
};
template <> struct SymIdx<3> {
  static constexpr int n = 6;
  static C8_DI int idx(int i, int j) {
    const int lo = i < j ? i : j, hi = i < j ? j : i;
    return lo == 0 ? hi : (lo == 1 ? 2 + hi : 5);
  }
};
template <class T, int DIM> C8_DI Mat<T, DIM> unpack_sym(const T* p) {
  Mat<T, DIM> r;
#pragma unroll
  for (int i = 0; i < DIM; ++i)
#pragma unroll
    for (int j = 0; j < DIM; ++j) r.a[i][j] = p[SymIdx<DIM>::idx(i, j)];
  return r;
}
template <class T, int DIM> C8_DI void pack_sym(const Mat<T, DIM>& A, T* p) {
#pragma unroll
  for (int i = 0; i < DIM; ++i)
#pragma unroll
    for (int j = i; j < DIM; ++j) p[SymIdx<DIM>::idx(i, j)] = A.a[i][j];
}

}  // namespace c8
