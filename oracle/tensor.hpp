// ORACLE (test infrastructure only).
//
// Small dense tensors of run-time dimension 2 or 3, restating the handful of
// MiniTensor operations calibr8's hot path uses (src/defines.hpp:33-37;
// call sites src/hyper_J2.cpp:11-15,147-152, src/small_J2.cpp:205).
// MiniTensor is an un-vendored Trilinos package; the published definitions
// restated here are: det/inverse by explicit cofactor formulas for N = 2, 3,
// trace, transpose, dev(A) = A - trace(A)/N I (N = tensor dimension),
// norm = Frobenius, eye, zero.
//
// Also a dense full-pivoting LU solve standing in for Eigen's
// `fullPivLu().solve()` (src/evaluations.cpp:112, src/small_J2.cpp:157).
#pragma once
#include <limits>
#include <vector>
#include <cassert>
#include "fad.hpp"

namespace orc {

template <class T>
struct Vec {
  int dim = 0;
  T a[3];
  Vec() {}
  explicit Vec(int d) : dim(d) { for (int i = 0; i < 3; ++i) a[i] = T(0.); }
  T& operator()(int i) { return a[i]; }
  T const& operator()(int i) const { return a[i]; }
  T& operator[](int i) { return a[i]; }
  T const& operator[](int i) const { return a[i]; }
};

template <class T>
struct Tensor {
  int dim = 0;
  T a[3][3];
  Tensor() {}
  explicit Tensor(int d) : dim(d) {
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) a[i][j] = T(0.);
  }
  int get_dimension() const { return dim; }
  T& operator()(int i, int j) { return a[i][j]; }
  T const& operator()(int i, int j) const { return a[i][j]; }
  Tensor& operator/=(T const& s) {
    for (int i = 0; i < dim; ++i) for (int j = 0; j < dim; ++j) a[i][j] = a[i][j] / s;
    return *this;
  }
};

template <class T> Tensor<T> zero(int d) { return Tensor<T>(d); }
template <class T> Tensor<T> eye(int d) {
  Tensor<T> r(d);
  for (int i = 0; i < d; ++i) r(i, i) = T(1.);
  return r;
}

template <class T> Tensor<T> operator+(Tensor<T> const& A, Tensor<T> const& B) {
  Tensor<T> r(A.dim);
  for (int i = 0; i < A.dim; ++i) for (int j = 0; j < A.dim; ++j) r(i, j) = A(i, j) + B(i, j);
  return r;
}
template <class T> Tensor<T> operator-(Tensor<T> const& A, Tensor<T> const& B) {
  Tensor<T> r(A.dim);
  for (int i = 0; i < A.dim; ++i) for (int j = 0; j < A.dim; ++j) r(i, j) = A(i, j) - B(i, j);
  return r;
}
template <class T> Tensor<T> operator*(Tensor<T> const& A, Tensor<T> const& B) {
  Tensor<T> r(A.dim);
  for (int i = 0; i < A.dim; ++i)
    for (int j = 0; j < A.dim; ++j) {
      T s = A(i, 0) * B(0, j);
      for (int k = 1; k < A.dim; ++k) s += A(i, k) * B(k, j);
      r(i, j) = s;
    }
  return r;
}
template <class T, class S> Tensor<T> scale(S const& s, Tensor<T> const& A) {
  Tensor<T> r(A.dim);
  for (int i = 0; i < A.dim; ++i) for (int j = 0; j < A.dim; ++j) r(i, j) = s * A(i, j);
  return r;
}
template <class T> Tensor<T> operator*(T const& s, Tensor<T> const& A) { return scale(s, A); }
template <class T> Tensor<T> operator*(Tensor<T> const& A, T const& s) { return scale(s, A); }
inline Tensor<Fad> operator*(double s, Tensor<Fad> const& A) { return scale(s, A); }
template <class T> Tensor<T> operator/(Tensor<T> const& A, T const& s) {
  Tensor<T> r(A.dim);
  for (int i = 0; i < A.dim; ++i) for (int j = 0; j < A.dim; ++j) r(i, j) = A(i, j) / s;
  return r;
}
inline Tensor<Fad> operator/(Tensor<Fad> const& A, double s) {
  Tensor<Fad> r(A.dim);
  for (int i = 0; i < A.dim; ++i) for (int j = 0; j < A.dim; ++j) r(i, j) = A(i, j) / s;
  return r;
}

template <class T> Tensor<T> transpose(Tensor<T> const& A) {
  Tensor<T> r(A.dim);
  for (int i = 0; i < A.dim; ++i) for (int j = 0; j < A.dim; ++j) r(i, j) = A(j, i);
  return r;
}
template <class T> T trace(Tensor<T> const& A) {
  T s = A(0, 0);
  for (int i = 1; i < A.dim; ++i) s += A(i, i);
  return s;
}
template <class T> T det(Tensor<T> const& A) {
  if (A.dim == 2) return A(0, 0) * A(1, 1) - A(1, 0) * A(0, 1);
  return -A(0, 2) * A(1, 1) * A(2, 0) + A(0, 1) * A(1, 2) * A(2, 0) +
         A(0, 2) * A(1, 0) * A(2, 1) - A(0, 0) * A(1, 2) * A(2, 1) -
         A(0, 1) * A(1, 0) * A(2, 2) + A(0, 0) * A(1, 1) * A(2, 2);
}
template <class T> Tensor<T> inverse(Tensor<T> const& A) {
  T const d = det(A);
  Tensor<T> r(A.dim);
  if (A.dim == 2) {
    r(0, 0) = A(1, 1); r(0, 1) = -A(0, 1);
    r(1, 0) = -A(1, 0); r(1, 1) = A(0, 0);
  } else {
    r(0, 0) = -A(1, 2) * A(2, 1) + A(1, 1) * A(2, 2);
    r(0, 1) = A(0, 2) * A(2, 1) - A(0, 1) * A(2, 2);
    r(0, 2) = -A(0, 2) * A(1, 1) + A(0, 1) * A(1, 2);
    r(1, 0) = A(1, 2) * A(2, 0) - A(1, 0) * A(2, 2);
    r(1, 1) = -A(0, 2) * A(2, 0) + A(0, 0) * A(2, 2);
    r(1, 2) = A(0, 2) * A(1, 0) - A(0, 0) * A(1, 2);
    r(2, 0) = -A(1, 1) * A(2, 0) + A(1, 0) * A(2, 1);
    r(2, 1) = A(0, 1) * A(2, 0) - A(0, 0) * A(2, 1);
    r(2, 2) = -A(0, 1) * A(1, 0) + A(0, 0) * A(1, 1);
  }
  return r / d;
}
template <class T> Tensor<T> dev(Tensor<T> const& A) {
  T const theta = trace(A) / double(A.dim);
  Tensor<T> r = A;
  for (int i = 0; i < A.dim; ++i) r(i, i) = A(i, i) - theta;
  return r;
}
template <class T> T norm(Tensor<T> const& A) {
  T s = A(0, 0) * A(0, 0);
  for (int i = 0; i < A.dim; ++i)
    for (int j = 0; j < A.dim; ++j)
      if (i + j > 0) s += A(i, j) * A(i, j);
  return sqrt(s);
}

// comparison on the value part (Sacado's max(a, b) returns the operand with the larger value, with its
// derivatives)
inline double fmax_val(double a, double b) { return a > b ? a : b; }
template <class T> T tmax(T const& a, T const& b) { return val(a) >= val(b) ? a : b; }
// MiniTensor norm_1 (max absolute column sum) and norm_infinity (max absolute row sum)
template <class T> T norm_1(Tensor<T> const& A) {
  T best = T(0.);
  for (int j = 0; j < A.dim; ++j) {
    T s = abs(A(0, j));
    for (int i = 1; i < A.dim; ++i) s += abs(A(i, j));
    best = (j == 0) ? s : tmax(best, s);
  }
  return best;
}
template <class T> T norm_infinity(Tensor<T> const& A) {
  T best = T(0.);
  for (int i = 0; i < A.dim; ++i) {
    T s = abs(A(i, 0));
    for (int j = 1; j < A.dim; ++j) s += abs(A(i, j));
    best = (i == 0) ? s : tmax(best, s);
  }
  return best;
}

// minitensor::polar_rotation (Trilinos MiniTensor, MiniTensor_LinearAlgebra.t.h; pinned with Trilinos at
// 33f21298..., pkg/trilinos/package.cmake:31-32; NOT vendored in the reference tree, restated from its
// published algorithm): the rotation R of the polar decomposition A = R U by Higham's scaled Newton
// iteration  X <- 1/2 (mu X + X^-T / mu),  mu = ((|Y|_1 |Y|_inf) / (|X|_1 |X|_inf))^(1/4), Y = X^-1,  with
// the scaling switched off once the relative change falls below 0.01 and the loop left when
// |Z - X|_F <= sqrt(sqrt(N) eps) or the change stops decreasing.  Evaluated in the AD scalar by the
// reference (src/global_residual.hpp:302-305), so dR/dF is the derivative OF THIS ITERATION: it lags the
// value by one step and is accurate to about |Z - X| of the last step (<= 2e-8), not to rounding.
template <class T> Tensor<T> polar_rotation(Tensor<T> const& A) {
  int const dimension = A.dim;
  bool scale = true;
  double const tol_scale = 0.01;
  double const tol_conv = std::sqrt(double(dimension)) * std::numeric_limits<double>::epsilon();
  Tensor<T> X = A;
  double gamma = 2.0;
  int const max_iter = 128;
  int num_iter = 0;
  while (num_iter < max_iter) {
    Tensor<T> const Y = inverse(X);
    T mu = T(1.0);
    if (scale) {
      mu = (norm_1(Y) * norm_infinity(Y)) / (norm_1(X) * norm_infinity(X));
      mu = sqrt(sqrt(mu));
    }
    Tensor<T> Z(dimension);
    Tensor<T> const YT = transpose(Y);
    for (int i = 0; i < dimension; ++i)
      for (int j = 0; j < dimension; ++j) Z(i, j) = 0.5 * (mu * X(i, j) + YT(i, j) / mu);
    Tensor<T> const D = Z - X;
    double const nD = val(norm(D));
    double const delta = nD / val(norm(Z));
    if (scale && delta < tol_scale) scale = false;
    bool const end_iter = nD <= std::sqrt(tol_conv) || (delta > 0.5 * gamma && !scale);
    X = Z;
    gamma = delta;
    if (end_iter) break;
    num_iter++;
  }
  return X;
}

// ---------------------------------------------------------------------------
// dense matrices (row-major) for the per-point solves
struct EMatrix {
  int r = 0, c = 0;
  std::vector<double> a;
  EMatrix() {}
  EMatrix(int r_, int c_) : r(r_), c(c_), a(size_t(r_) * c_, 0.) {}
  double& operator()(int i, int j) { return a[size_t(i) * c + j]; }
  double operator()(int i, int j) const { return a[size_t(i) * c + j]; }
  int rows() const { return r; }
  int cols() const { return c; }
  EMatrix transpose() const {
    EMatrix t(c, r);
    for (int i = 0; i < r; ++i) for (int j = 0; j < c; ++j) t(j, i) = (*this)(i, j);
    return t;
  }
};
using EVector = std::vector<double>;

inline EMatrix neg(EMatrix const& A) {
  EMatrix B = A;
  for (auto& v : B.a) v = -v;
  return B;
}
inline EMatrix matmul(EMatrix const& A, EMatrix const& B) {
  EMatrix C(A.r, B.c);
  for (int i = 0; i < A.r; ++i)
    for (int j = 0; j < B.c; ++j) {
      double s = 0.;
      for (int k = 0; k < A.c; ++k) s += A(i, k) * B(k, j);
      C(i, j) = s;
    }
  C8_FLOPS(2LL * A.r * B.c * A.c);
  return C;
}
inline EVector matvec(EMatrix const& A, EVector const& x) {
  EVector y(A.r, 0.);
  for (int i = 0; i < A.r; ++i) {
    double s = 0.;
    for (int k = 0; k < A.c; ++k) s += A(i, k) * x[k];
    y[i] = s;
  }
  C8_FLOPS(2LL * A.r * A.c);
  return y;
}

// Full-pivoting LU solve of A X = B (A n x n, B n x m) -- the role of
// Eigen::FullPivLU::solve at src/evaluations.cpp:112.
inline EMatrix full_piv_lu_solve(EMatrix A, EMatrix B) {
  // Eigen::FullPivLU semantics (Eigen/src/LU/FullPivLU.h, computeInPlace + _solve_impl): the
  // elimination stops at an exactly zero corner; the solve uses the `rank` leading pivots (those above
  // maxpivot * epsilon * n) and sets the remaining unknowns to zero -- so a zero matrix (the elastic
  // model's dC/dxi, src/elastic.cpp) solves to zero instead of NaN.
  int const n = A.r, m = B.c;
  std::vector<int> colperm(n);
  for (int i = 0; i < n; ++i) colperm[i] = i;
  int nonzero_pivots = n;
  double maxpivot = 0.;
  for (int k = 0; k < n; ++k) {
    int pr = k, pc = k;
    double best = -1.;
    for (int i = k; i < n; ++i)
      for (int j = k; j < n; ++j) {
        double const v = std::abs(A(i, j));
        if (v > best) { best = v; pr = i; pc = j; }
      }
    if (best == 0.) { nonzero_pivots = k; break; }
    if (best > maxpivot) maxpivot = best;
    if (pr != k) {
      for (int j = 0; j < n; ++j) std::swap(A(k, j), A(pr, j));
      for (int j = 0; j < m; ++j) std::swap(B(k, j), B(pr, j));
    }
    if (pc != k) {
      for (int i = 0; i < n; ++i) std::swap(A(i, k), A(i, pc));
      std::swap(colperm[k], colperm[pc]);
    }
    double const piv = A(k, k);
    for (int i = k + 1; i < n; ++i) {
      double const l = A(i, k) / piv;
      A(i, k) = l;
      for (int j = k + 1; j < n; ++j) A(i, j) -= l * A(k, j);
      for (int j = 0; j < m; ++j) B(i, j) -= l * B(k, j);
    }
    C8_FLOPS((long long)(n - k - 1) * (1 + 2 * (n - k - 1) + 2 * m));
  }
  int rank = 0;
  double const thresh = maxpivot * (std::numeric_limits<double>::epsilon() * n);
  for (int i = 0; i < nonzero_pivots; ++i) rank += (std::abs(A(i, i)) > thresh);
  EMatrix Y(n, m);
  for (int j = 0; j < m; ++j) {
    for (int i = rank; i < n; ++i) Y(i, j) = 0.;
    for (int i = rank - 1; i >= 0; --i) {
      double s = B(i, j);
      for (int k = i + 1; k < rank; ++k) s -= A(i, k) * Y(k, j);
      Y(i, j) = s / A(i, i);
    }
  }
  C8_FLOPS((long long)m * n * (n + 1));
  EMatrix X(n, m);
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < m; ++j) X(colperm[i], j) = Y(i, j);
  return X;
}
inline EVector full_piv_lu_solve(EMatrix const& A, EVector const& b) {
  EMatrix B(A.r, 1);
  for (int i = 0; i < A.r; ++i) B(i, 0) = b[i];
  EMatrix X = full_piv_lu_solve(A, B);
  EVector x(A.r);
  for (int i = 0; i < A.r; ++i) x[i] = X(i, 0);
  return x;
}

}  // namespace orc
