// Lane-split fixed-width forward-mode dual numbers for sm_100a fp64 kernels.
//
// Replaces Sacado::Fad::SLFad<double,16> (reference: src/defines.hpp:23-26).
// A derivative array of width W is split across the G threads that cooperate
// on one quadrature point: thread t of the group holds the value (replicated)
// and the L = ceil(W/G) derivative lanes [t*L, (t+1)*L).  All arithmetic is
// thread-local and register-resident; lanes only meet in the small dense
// solves (group shuffles, see groupsolve.cuh).
//
// Unseeded quantities stay plain `double`, so a pass only pays derivative
// arithmetic on what it actually differentiates (the reference pays the full
// 16-wide array on every operation).
#pragma once
#include <cuda_runtime.h>
#include <math.h>

#define C8_DI __device__ __forceinline__

namespace c8 {

// c ? a : b through an explicit selp.  Written as a C++ select inside an unrolled
// "pick entry n of a register array" chain, the compiler turns the chain into a
// dynamically indexed local-memory array and drags the whole array out of registers.
C8_DI double pick(int c, double a, double b) {
  double r;
  asm("{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %3, 0;\n\tselp.f64 %0, %1, %2, p;\n\t}"
      : "=d"(r) : "d"(a), "d"(b), "r"(c));
  return r;
}

C8_DI int picki(int c, int a, int b) {
  int r;
  asm("{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %3, 0;\n\tselp.s32 %0, %1, %2, p;\n\t}"
      : "=r"(r) : "r"(a), "r"(b), "r"(c));
  return r;
}

// fp64 reduction to global memory without a return value (RED.E.ADD.F64)
C8_DI void red_add(double* addr, double v) {
  asm volatile("red.global.add.f64 [%0], %1;" :: "l"(addr), "d"(v) : "memory");
}

template <int L>
struct Dual {
  double v;
  double d[L];
};

template <class T> struct is_dual { static constexpr bool value = false; };
template <int L> struct is_dual<Dual<L>> { static constexpr bool value = true; };

// value part
C8_DI double val(double x) { return x; }
template <int L> C8_DI double val(const Dual<L>& x) { return x.v; }

template <int L> C8_DI Dual<L> make_dual(double v) {
  Dual<L> r; r.v = v;
#pragma unroll
  for (int i = 0; i < L; ++i) r.d[i] = 0.0;
  return r;
}
// seed derivative index k of a width-(G*L) array; `base` = first lane owned by this thread
template <int L> C8_DI Dual<L> seeded(double v, int k, int base) {
  Dual<L> r; r.v = v;
#pragma unroll
  for (int i = 0; i < L; ++i) r.d[i] = (k == base + i) ? 1.0 : 0.0;
  return r;
}

// promotion of mixed double / Dual arithmetic
template <class A, class B> struct Prom { using type = double; };
template <int L> struct Prom<Dual<L>, double> { using type = Dual<L>; };
template <int L> struct Prom<double, Dual<L>> { using type = Dual<L>; };
template <int L> struct Prom<Dual<L>, Dual<L>> { using type = Dual<L>; };
template <class A, class B> using prom_t = typename Prom<A, B>::type;
template <class A, class B, class C> using prom3_t = prom_t<prom_t<A, B>, C>;
template <class A, class B, class C, class D> using prom4_t = prom_t<prom3_t<A, B, C>, D>;

template <class T> C8_DI T lift(double x);
template <> C8_DI double lift<double>(double x) { return x; }
template <class T> C8_DI T lift_dual(double x) {
  T r; r.v = x;
#pragma unroll
  for (int i = 0; i < int(sizeof(r.d) / sizeof(double)); ++i) r.d[i] = 0.0;
  return r;
}
template <class T, class S> struct Conv;
template <> struct Conv<double, double> { static C8_DI double f(double x) { return x; } };
template <int L> struct Conv<Dual<L>, double> {
  static C8_DI Dual<L> f(double x) { return make_dual<L>(x); }
};
template <int L> struct Conv<Dual<L>, Dual<L>> {
  static C8_DI Dual<L> f(const Dual<L>& x) { return x; }
};
// convert S -> T (T is S or a promotion of it)
template <class T, class S> C8_DI T conv(const S& x) { return Conv<T, S>::f(x); }

// ---------------------------------------------------------------- add / sub
template <int L> C8_DI Dual<L> operator+(const Dual<L>& a, const Dual<L>& b) {
  Dual<L> r; r.v = a.v + b.v;
#pragma unroll
  for (int i = 0; i < L; ++i) r.d[i] = a.d[i] + b.d[i];
  return r;
}
template <int L> C8_DI Dual<L> operator+(const Dual<L>& a, double b) { Dual<L> r = a; r.v = a.v + b; return r; }
template <int L> C8_DI Dual<L> operator+(double a, const Dual<L>& b) { Dual<L> r = b; r.v = a + b.v; return r; }
template <int L> C8_DI Dual<L> operator-(const Dual<L>& a, const Dual<L>& b) {
  Dual<L> r; r.v = a.v - b.v;
#pragma unroll
  for (int i = 0; i < L; ++i) r.d[i] = a.d[i] - b.d[i];
  return r;
}
template <int L> C8_DI Dual<L> operator-(const Dual<L>& a, double b) { Dual<L> r = a; r.v = a.v - b; return r; }
template <int L> C8_DI Dual<L> operator-(double a, const Dual<L>& b) {
  Dual<L> r; r.v = a - b.v;
#pragma unroll
  for (int i = 0; i < L; ++i) r.d[i] = -b.d[i];
  return r;
}
template <int L> C8_DI Dual<L> operator-(const Dual<L>& a) {
  Dual<L> r; r.v = -a.v;
#pragma unroll
  for (int i = 0; i < L; ++i) r.d[i] = -a.d[i];
  return r;
}
template <int L> C8_DI Dual<L>& operator+=(Dual<L>& a, const Dual<L>& b) { a = a + b; return a; }
template <int L> C8_DI Dual<L>& operator+=(Dual<L>& a, double b) { a.v += b; return a; }
template <int L> C8_DI Dual<L>& operator-=(Dual<L>& a, const Dual<L>& b) { a = a - b; return a; }
template <int L> C8_DI Dual<L>& operator-=(Dual<L>& a, double b) { a.v -= b; return a; }

// ---------------------------------------------------------------- mul / div
template <int L> C8_DI Dual<L> operator*(const Dual<L>& a, const Dual<L>& b) {
  Dual<L> r; r.v = a.v * b.v;
#pragma unroll
  for (int i = 0; i < L; ++i) r.d[i] = fma(a.d[i], b.v, a.v * b.d[i]);
  return r;
}
template <int L> C8_DI Dual<L> operator*(const Dual<L>& a, double b) {
  Dual<L> r; r.v = a.v * b;
#pragma unroll
  for (int i = 0; i < L; ++i) r.d[i] = a.d[i] * b;
  return r;
}
template <int L> C8_DI Dual<L> operator*(double a, const Dual<L>& b) { return b * a; }
template <int L> C8_DI Dual<L> operator/(const Dual<L>& a, const Dual<L>& b) {
  Dual<L> r;
  const double inv = 1.0 / b.v;
  r.v = a.v * inv;
#pragma unroll
  for (int i = 0; i < L; ++i) r.d[i] = fma(-r.v, b.d[i], a.d[i]) * inv;
  return r;
}
template <int L> C8_DI Dual<L> operator/(const Dual<L>& a, double b) {
  const double inv = 1.0 / b;
  return a * inv;
}
template <int L> C8_DI Dual<L> operator/(double a, const Dual<L>& b) {
  Dual<L> r;
  const double inv = 1.0 / b.v;
  r.v = a * inv;
  const double s = -r.v * inv;
#pragma unroll
  for (int i = 0; i < L; ++i) r.d[i] = s * b.d[i];
  return r;
}
template <int L> C8_DI Dual<L>& operator*=(Dual<L>& a, const Dual<L>& b) { a = a * b; return a; }
template <int L> C8_DI Dual<L>& operator*=(Dual<L>& a, double b) { a = a * b; return a; }
template <int L> C8_DI Dual<L>& operator/=(Dual<L>& a, const Dual<L>& b) { a = a / b; return a; }
template <int L> C8_DI Dual<L>& operator/=(Dual<L>& a, double b) { a = a / b; return a; }

// ---------------------------------------------------------------- functions
C8_DI double dsqrt(double a) { return sqrt(a); }
template <int L> C8_DI Dual<L> dsqrt(const Dual<L>& a) {
  Dual<L> r; r.v = sqrt(a.v);
  const double s = 0.5 / r.v;
#pragma unroll
  for (int i = 0; i < L; ++i) r.d[i] = a.d[i] * s;
  return r;
}
C8_DI double dexp(double a) { return exp(a); }
template <int L> C8_DI Dual<L> dexp(const Dual<L>& a) {
  Dual<L> r; r.v = exp(a.v);
#pragma unroll
  for (int i = 0; i < L; ++i) r.d[i] = r.v * a.d[i];
  return r;
}
C8_DI double dcbrt(double a) { return cbrt(a); }
template <int L> C8_DI Dual<L> dcbrt(const Dual<L>& a) {
  Dual<L> r; r.v = cbrt(a.v);
  const double s = 1.0 / (3.0 * r.v * r.v);
#pragma unroll
  for (int i = 0; i < L; ++i) r.d[i] = a.d[i] * s;
  return r;
}
// x^2 and x^-2 (the only constant-exponent powers on the path, src/yield_functions.hpp:38-66)
C8_DI double sqr(double a) { return a * a; }
template <int L> C8_DI Dual<L> sqr(const Dual<L>& a) {
  Dual<L> r; r.v = a.v * a.v;
  const double s = 2.0 * a.v;
#pragma unroll
  for (int i = 0; i < L; ++i) r.d[i] = s * a.d[i];
  return r;
}
C8_DI double inv_sqr(double a) { return 1.0 / (a * a); }
template <int L> C8_DI Dual<L> inv_sqr(const Dual<L>& a) {
  Dual<L> r;
  const double inv = 1.0 / a.v;
  r.v = inv * inv;
  const double s = -2.0 * r.v * inv;
#pragma unroll
  for (int i = 0; i < L; ++i) r.d[i] = s * a.d[i];
  return r;
}
// a^b: an unseeded (double) operand contributes no log()/division term, which is how the
// reference's AD type treats a constant operand (e.g. the unseeded exponent `n` of
// src/hyper_J2.cpp:260-262 with alpha + 1e-12 possibly negative during a Newton step).
C8_DI double dpow(double a, double b) { return pow(a, b); }
template <int L> C8_DI Dual<L> dpow(const Dual<L>& a, double b) {
  Dual<L> r; r.v = pow(a.v, b);
  // branch-free: s = 1 (b == 1), 0 (a == 0), b/a * a^b otherwise
  double s = pick(a.v == 0.0, 0.0, b / a.v * r.v);
  s = pick(b == 1.0, 1.0, s);
#pragma unroll
  for (int i = 0; i < L; ++i) r.d[i] = s * a.d[i];
  return r;
}
template <int L> C8_DI Dual<L> dpow(double a, const Dual<L>& b) {
  Dual<L> r; r.v = pow(a, b.v);
  const double s = (a == 0.0) ? 0.0 : log(a) * r.v;
#pragma unroll
  for (int i = 0; i < L; ++i) r.d[i] = s * b.d[i];
  return r;
}
template <int L> C8_DI Dual<L> dpow(const Dual<L>& a, const Dual<L>& b) {
  Dual<L> r; r.v = pow(a.v, b.v);
  if (a.v == 0.0) {
#pragma unroll
    for (int i = 0; i < L; ++i) r.d[i] = 0.0;
  } else {
    const double la = log(a.v), s = b.v / a.v;
#pragma unroll
    for (int i = 0; i < L; ++i) r.d[i] = (b.d[i] * la + s * a.d[i]) * r.v;
  }
  return r;
}

}  // namespace c8
