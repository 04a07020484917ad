// ORACLE (test infrastructure only -- never imported by the product path).
//
// Forward-mode AD scalar with a static maximum width of 16, restating the
// semantics of Sacado::Fad::SLFad<double,16> as calibr8 uses it
// (/root/reference/source/calibr8/src/defines.hpp:23-26).  Sacado itself is an
// un-vendored Trilinos dependency (pin: package/trilinos/package.cmake:31-32);
// what is restated here is its published behaviour:
//   * a value plus `n` derivative components, n == 0 meaning "constant";
//   * binary ops produce max(n_a, n_b) components, a missing operand
//     derivative reads as 0;
//   * diff(i, n): resize to n, zero, set component i to 1;
//   * comparisons act on the value only;
//   * pow(a,b): d = (b' log a + b a'/a) a^b, with a == 0 giving derivative 0.
//
// Compile with -DC8_ORACLE_COUNT to count one flop per scalar +,-,*,/ and per
// transcendental on every value and every derivative lane.
#pragma once
#include <cmath>
#include <algorithm>

namespace orc {

#ifdef C8_ORACLE_COUNT
extern thread_local long long g_flops;
#define C8_FLOPS(n) (::orc::g_flops += (n))
#else
#define C8_FLOPS(n) ((void)0)
#endif

static constexpr int nmax_derivs = 16;

struct Fad {
  double v = 0.;
  int n = 0;
  double d[nmax_derivs];

  Fad() : v(0.), n(0) { for (int i = 0; i < nmax_derivs; ++i) d[i] = 0.; }
  Fad(double x) : v(x), n(0) { for (int i = 0; i < nmax_derivs; ++i) d[i] = 0.; }

  double& val() { return v; }
  double val() const { return v; }
  int size() const { return n; }
  double dx(int i) const { return (i < n) ? d[i] : 0.; }
  double& fastAccessDx(int i) { return d[i]; }
  double fastAccessDx(int i) const { return d[i]; }

  // SLFad::diff
  void diff(int i, int nd) {
    n = nd;
    for (int k = 0; k < nmax_derivs; ++k) d[k] = 0.;
    d[i] = 1.;
  }
  // make sure storage up to nd is addressable (Sacado resize semantics)
  void resize(int nd) {
    for (int k = n; k < nd; ++k) d[k] = 0.;
    n = nd;
  }

  Fad& operator+=(Fad const& b);
  Fad& operator-=(Fad const& b);
  Fad& operator*=(Fad const& b);
  Fad& operator/=(Fad const& b);
  Fad& operator+=(double b) { v += b; C8_FLOPS(1); return *this; }
  Fad& operator-=(double b) { v -= b; C8_FLOPS(1); return *this; }
  Fad& operator*=(double b) {
    v *= b; for (int i = 0; i < n; ++i) d[i] *= b; C8_FLOPS(1 + n); return *this;
  }
  Fad& operator/=(double b) {
    v /= b; for (int i = 0; i < n; ++i) d[i] /= b; C8_FLOPS(1 + n); return *this;
  }
};

inline Fad operator-(Fad const& a) {
  Fad r; r.v = -a.v; r.n = a.n;
  for (int i = 0; i < a.n; ++i) r.d[i] = -a.d[i];
  return r;
}

inline Fad operator+(Fad const& a, Fad const& b) {
  Fad r; r.v = a.v + b.v; r.n = std::max(a.n, b.n);
  for (int i = 0; i < r.n; ++i) r.d[i] = a.dx(i) + b.dx(i);
  C8_FLOPS(1 + r.n);
  return r;
}
inline Fad operator+(Fad const& a, double b) { Fad r = a; r.v = a.v + b; C8_FLOPS(1); return r; }
inline Fad operator+(double a, Fad const& b) { Fad r = b; r.v = a + b.v; C8_FLOPS(1); return r; }

inline Fad operator-(Fad const& a, Fad const& b) {
  Fad r; r.v = a.v - b.v; r.n = std::max(a.n, b.n);
  for (int i = 0; i < r.n; ++i) r.d[i] = a.dx(i) - b.dx(i);
  C8_FLOPS(1 + r.n);
  return r;
}
inline Fad operator-(Fad const& a, double b) { Fad r = a; r.v = a.v - b; C8_FLOPS(1); return r; }
inline Fad operator-(double a, Fad const& b) {
  Fad r; r.v = a - b.v; r.n = b.n;
  for (int i = 0; i < b.n; ++i) r.d[i] = -b.d[i];
  C8_FLOPS(1);
  return r;
}

inline Fad operator*(Fad const& a, Fad const& b) {
  Fad r; r.v = a.v * b.v; r.n = std::max(a.n, b.n);
  for (int i = 0; i < r.n; ++i) r.d[i] = a.dx(i) * b.v + a.v * b.dx(i);
  C8_FLOPS(1 + 3 * r.n);
  return r;
}
inline Fad operator*(Fad const& a, double b) {
  Fad r; r.v = a.v * b; r.n = a.n;
  for (int i = 0; i < a.n; ++i) r.d[i] = a.d[i] * b;
  C8_FLOPS(1 + r.n);
  return r;
}
inline Fad operator*(double a, Fad const& b) { return b * a; }

inline Fad operator/(Fad const& a, Fad const& b) {
  Fad r; r.v = a.v / b.v; r.n = std::max(a.n, b.n);
  double const b2 = b.v * b.v;
  for (int i = 0; i < r.n; ++i) r.d[i] = (a.dx(i) * b.v - a.v * b.dx(i)) / b2;
  C8_FLOPS(2 + 4 * r.n);
  return r;
}
inline Fad operator/(Fad const& a, double b) {
  Fad r; r.v = a.v / b; r.n = a.n;
  for (int i = 0; i < a.n; ++i) r.d[i] = a.d[i] / b;
  C8_FLOPS(1 + r.n);
  return r;
}
inline Fad operator/(double a, Fad const& b) {
  Fad r; r.v = a / b.v; r.n = b.n;
  double const b2 = b.v * b.v;
  for (int i = 0; i < b.n; ++i) r.d[i] = -a * b.d[i] / b2;
  C8_FLOPS(2 + 3 * r.n);
  return r;
}

inline Fad& Fad::operator+=(Fad const& b) { *this = *this + b; return *this; }
inline Fad& Fad::operator-=(Fad const& b) { *this = *this - b; return *this; }
inline Fad& Fad::operator*=(Fad const& b) { *this = *this * b; return *this; }
inline Fad& Fad::operator/=(Fad const& b) { *this = *this / b; return *this; }

inline bool operator>(Fad const& a, double b) { return a.v > b; }
inline bool operator<(Fad const& a, double b) { return a.v < b; }
inline bool operator>(Fad const& a, Fad const& b) { return a.v > b.v; }
inline bool operator<(Fad const& a, Fad const& b) { return a.v < b.v; }

inline Fad sqrt(Fad const& a) {
  Fad r; r.v = std::sqrt(a.v); r.n = a.n;
  for (int i = 0; i < a.n; ++i) r.d[i] = a.d[i] / (2. * r.v);
  C8_FLOPS(1 + 2 * r.n);
  return r;
}
inline Fad exp(Fad const& a) {
  Fad r; r.v = std::exp(a.v); r.n = a.n;
  for (int i = 0; i < a.n; ++i) r.d[i] = r.v * a.d[i];
  C8_FLOPS(1 + r.n);
  return r;
}
inline Fad log(Fad const& a) {
  Fad r; r.v = std::log(a.v); r.n = a.n;
  for (int i = 0; i < a.n; ++i) r.d[i] = a.d[i] / a.v;
  C8_FLOPS(1 + r.n);
  return r;
}
inline Fad cbrt(Fad const& a) {
  Fad r; r.v = std::cbrt(a.v); r.n = a.n;
  for (int i = 0; i < a.n; ++i) r.d[i] = a.d[i] / (3. * r.v * r.v);
  C8_FLOPS(1 + 3 * r.n);
  return r;
}
inline Fad abs(Fad const& a) {
  if (a.v >= 0.) return a;
  return -a;
}
inline Fad pow(Fad const& a, double b) {
  Fad r; r.v = std::pow(a.v, b); r.n = a.n;
  if (b == 1.) {
    for (int i = 0; i < a.n; ++i) r.d[i] = a.d[i];
  } else if (a.v == 0.) {
    for (int i = 0; i < a.n; ++i) r.d[i] = 0.;
  } else {
    for (int i = 0; i < a.n; ++i) r.d[i] = b * a.d[i] / a.v * r.v;
  }
  C8_FLOPS(1 + 3 * r.n);
  return r;
}
inline Fad pow(Fad const& a, int b) { return pow(a, double(b)); }
inline Fad pow(Fad const& a, Fad const& b) {
  // size-aware like Sacado's PowerOp: an operand with no derivative array
  // (an unseeded parameter) contributes no log()/division term at all.
  Fad r; r.v = std::pow(a.v, b.v); r.n = std::max(a.n, b.n);
  if (a.n > 0 && b.n > 0) {
    if (a.v == 0.) {
      for (int i = 0; i < r.n; ++i) r.d[i] = 0.;
    } else {
      double const la = std::log(a.v);
      for (int i = 0; i < r.n; ++i)
        r.d[i] = (b.dx(i) * la + b.v * a.dx(i) / a.v) * r.v;
    }
  } else if (a.n > 0) {
    if (b.v == 1.) {
      for (int i = 0; i < r.n; ++i) r.d[i] = a.d[i];
    } else if (a.v == 0.) {
      for (int i = 0; i < r.n; ++i) r.d[i] = 0.;
    } else {
      for (int i = 0; i < r.n; ++i) r.d[i] = b.v * a.d[i] / a.v * r.v;
    }
  } else if (b.n > 0) {
    if (a.v == 0.) {
      for (int i = 0; i < r.n; ++i) r.d[i] = 0.;
    } else {
      double const la = std::log(a.v);
      for (int i = 0; i < r.n; ++i) r.d[i] = b.d[i] * la * r.v;
    }
  }
  C8_FLOPS(2 + 5 * r.n);
  return r;
}

// ---- the `val()` / `dx()` helpers of src/fad.hpp:14-44 ----
inline double val(double x) { return x; }
inline double val(Fad const& x) { return x.v; }
inline double dx(double, int) { return 0.; }
inline double dx(Fad const& x, int i) { return x.d[i]; }

// double versions so templates resolve uniformly
inline double cbrt(double a) { C8_FLOPS(1); return std::cbrt(a); }
using std::sqrt; using std::exp; using std::pow; using std::abs; using std::log;

}  // namespace orc
