// Batched small dense solves held in registers, one per quadrature point, with
// the columns distributed over the G threads of the point's thread group.
//
// Replaces Eigen's heap-allocated `fullPivLu().solve()` of the reference
// (src/evaluations.cpp:112,456,624; src/small_J2.cpp:157).  Gauss-Jordan with
// row pivoting: the pivot column is broadcast with group shuffles, every thread
// then eliminates in the columns it owns (its slice of J, its slice of the
// right-hand sides, and the replicated single rhs).  Row pivoting instead of
// full pivoting changes rounding only.
#pragma once
#include "dual.cuh"

namespace c8 {

// conditional swap through explicit selp: written as C++ selects the compiler rewrites the
// unrolled swap chain into a dynamically indexed local-memory array.
C8_DI void cswap(int sw, double& a, double& b) {
  double ta, tb;
  asm("{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %4, 0;\n\tselp.f64 %0, %3, %2, p;\n\t"
      "selp.f64 %1, %2, %3, p;\n\t}"
      : "=&d"(ta), "=&d"(tb)
      : "d"(a), "d"(b), "r"(sw));
  a = ta; b = tb;
}

template <int G> C8_DI double group_bcast(unsigned mask, double v, int src) {
  return __shfl_sync(mask, v, src, G);
}
// Same, for call sites reached by ALL 32 lanes of the warp together (warp-uniform control flow):
// with a compile-time full mask the shuffle is two plain SHFL.IDX; a run-time sub-warp mask makes
// the compiler wrap every shuffle in a WARPSYNC / reconvergence sequence (about a quarter of K1's
// static code before this, see profiles/README.md).
template <int G> C8_DI double group_bcast_full(double v, int src) {
  return __shfl_sync(0xffffffffu, v, src, G);
}

// J (N x N): column c lives in thread c / LJ, slot c % LJ  -> Jc[row][slot]
// B (N x G*LB): thread owns LB columns                     -> Bc[row][slot]
// b (N): replicated on every thread of the group
// On return Bc = J^-1 B and b = J^-1 b (J is destroyed).
// FULL: the call site is warp-uniform (all 32 lanes arrive together) -> full-mask shuffles.
template <int N, int LJ, int LB, int G, bool FULL = false>
C8_DI void group_gauss_jordan(double (&Jc)[N][LJ], double (&Bc)[N][LB > 0 ? LB : 1],
                              double (&b)[N], unsigned mask) {
#pragma unroll
  for (int k = 0; k < N; ++k) {
    const int owner = k / LJ, slot = k % LJ;
    double col[N];
#pragma unroll
    for (int i = 0; i < N; ++i)
      col[i] = FULL ? group_bcast_full<G>(Jc[i][slot], owner) : group_bcast<G>(mask, Jc[i][slot], owner);
    int p = k;
    double best = fabs(col[k]);
#pragma unroll
    for (int i = k + 1; i < N; ++i) {
      const double a = fabs(col[i]);
      if (a > best) { best = a; p = i; }
    }
    if (__builtin_expect(p != k, 0)) {  // group-uniform (col is identical on all threads of the group)
#pragma unroll
      for (int i = k + 1; i < N; ++i) {
        const int sw = (i == p) ? 1 : 0;
        cswap(sw, col[k], col[i]);
        cswap(sw, b[k], b[i]);
#pragma unroll
        for (int s = 0; s < LJ; ++s) cswap(sw, Jc[k][s], Jc[i][s]);
#pragma unroll
        for (int s = 0; s < LB; ++s) cswap(sw, Bc[k][s], Bc[i][s]);
      }
    }
    const double inv = 1.0 / col[k];
    b[k] *= inv;
#pragma unroll
    for (int s = 0; s < LJ; ++s) Jc[k][s] *= inv;
#pragma unroll
    for (int s = 0; s < LB; ++s) Bc[k][s] *= inv;
#pragma unroll
    for (int i = 0; i < N; ++i) {
      if (i == k) continue;
      const double m = col[i];
      b[i] = fma(-m, b[k], b[i]);
#pragma unroll
      for (int s = 0; s < LJ; ++s) Jc[i][s] = fma(-m, Jc[k][s], Jc[i][s]);
#pragma unroll
      for (int s = 0; s < LB; ++s) Bc[i][s] = fma(-m, Bc[k][s], Bc[i][s]);
    }
  }
}

}  // namespace c8
