// K1 / K2: the forward element loop of the reference,
//   eval_forward_jacobian   src/evaluations.cpp:12-154
//   eval_global_residual    src/evaluations.cpp:156-259
// One thread GROUP of G threads per element (= per coupled quadrature point;
// linear simplices carry one).  The AD seeding choreography of the reference
// (Appendix B of SURVEY.md) becomes three register-resident passes:
//   P1  local Newton, xi seeded, width NXI split over the group   -> dC/dxi
//   P2  C re-evaluated with the element dofs seeded, width NX     -> dC/dx
//       dxi/dx = -(dC/dxi)^-1 dC/dx   (group Gauss-Jordan, groupsolve.cuh)
//   P3  global residual with x seeded and xi carrying dxi/dx      -> R_e, dR/dx
// followed by the pressure-mass ip set and ONE scatter per element (the
// reference scatters the full block once per integration point, 5x per tet).
#pragma once
#include "args.h"
#include "groupsolve.cuh"
#include "mechanics.cuh"

namespace c8 {

template <int DIM, int MECH, class Model_, int G_>
struct Cfg {
  using Model = Model_;
  using MT = MechTraits<DIM, MECH>;
  static constexpr int D = DIM, M = MECH, G = G_;
  static constexpr int NN = MT::NN, NB = MT::NB, NX = MT::NX;
  static constexpr int NXI = Model::NXI, NPAR = Model::NPAR;
  static constexpr int LX = (NX + G - 1) / G;
  static constexpr int LXI = (NXI + G - 1) / G;
  static_assert(32 % G == 0, "group must divide the warp");
  // position of element dof (n, eq) in the reference's residual-major order
  static C8_DI int ref_dof(int n, int eq) { return eq < DIM ? n * DIM + eq : NN * DIM + n; }
};

template <class C>
struct Elem {
  int nodes[C::NN];
  Geom<C::D> g;
  double xn[C::NN][C::NB];
  double xpn[C::NN][C::NB];
  double par[C::NPAR];
  double xip[C::NXI];
};

template <class C>
C8_DI void load_elem(const MeshArgs& m, const ModelArgs& md, const double* __restrict__ x,
                     const double* __restrict__ x_prev, const double* __restrict__ xi_prev,
                     long long xi_ld, int e, Elem<C>& E) {
#pragma unroll
  for (int n = 0; n < C::NN; ++n) E.nodes[n] = __ldg(&m.conn[size_t(e) * C::NN + n]);
  load_geom<C::D>(m.coords, E.nodes, E.g);
#pragma unroll
  for (int n = 0; n < C::NN; ++n)
#pragma unroll
    for (int q = 0; q < C::NB; ++q) {
      E.xn[n][q] = __ldg(&x[size_t(E.nodes[n]) * C::NB + q]);
      E.xpn[n][q] = x_prev ? __ldg(&x_prev[size_t(E.nodes[n]) * C::NB + q]) : 0.0;
    }
  const int es = m.elem_es ? __ldg(&m.elem_es[e]) : 0;
#pragma unroll
  for (int q = 0; q < C::NPAR; ++q) E.par[q] = __ldg(&md.params[es * md.npar + q]);
#pragma unroll
  for (int q = 0; q < C::NXI; ++q) E.xip[q] = __ldg(&xi_prev[size_t(q) * xi_ld + e]);
}

// The same record written ONCE per thread group into shared memory (E lives in __shared__): the G
// threads split the loads -- thread n < NN fetches node n's id and nodal rows, the parameters and
// xi_prev are strided over the group, the last thread computes the geometry -- instead of every
// thread holding a private copy (~65 doubles in 3-D) in registers.  Callers __syncwarp() afterwards
// (a group never spans two warps).
template <class C>
C8_DI void load_elem_shared(const MeshArgs& m, const ModelArgs& md, const double* __restrict__ x,
                            const double* __restrict__ x_prev, const double* __restrict__ xi_prev,
                            long long xi_ld, int e, int t, Elem<C>& E) {
  static_assert(C::G >= 2, "needs at least two threads per group");
  for (int n = t; n < C::NN; n += C::G) {
    const int nd = __ldg(&m.conn[size_t(e) * C::NN + n]);
    E.nodes[n] = nd;
#pragma unroll
    for (int q = 0; q < C::NB; ++q) {
      E.xn[n][q] = __ldg(&x[size_t(nd) * C::NB + q]);
      E.xpn[n][q] = x_prev ? __ldg(&x_prev[size_t(nd) * C::NB + q]) : 0.0;
    }
  }
  const int es = m.elem_es ? __ldg(&m.elem_es[e]) : 0;
  for (int q = t; q < C::NPAR; q += C::G) E.par[q] = __ldg(&md.params[es * md.npar + q]);
  for (int q = t; q < C::NXI; q += C::G) E.xip[q] = __ldg(&xi_prev[size_t(q) * xi_ld + e]);
  if (t == C::G - 1) {
    int nodes[C::NN];
#pragma unroll
    for (int n = 0; n < C::NN; ++n) nodes[n] = __ldg(&m.conn[size_t(e) * C::NN + n]);
    load_geom<C::D>(m.coords, nodes, E.g);
  }
}

// The local Newton of LocalResidual::solve_nonlinear (e.g. src/small_J2.cpp:121-173).
// xi: in = values gathered from the current xi field, out = converged state.
// Cd: the last residual evaluation with xi seeded (value + this thread's dC/dxi columns).
// Returns the branch (0/1) or -1 when not converged within max_iters.
//
// WARP-SYNCHRONOUS: every thread of a warp must call this (inactive groups pass active = false);
// the body is executed by all 32 lanes together so that the in-group solves can use full-mask
// shuffles (profiles/README.md).
template <class C>
C8_DI int local_newton(const Kin<C::D, double, double>& k0, const Elem<C>& E, const ModelArgs& md,
                       double (&xi)[C::NXI], Dual<C::LXI> (&Cd)[C::NXI], unsigned mask, int t,
                       bool active) {
  using Model = typename C::Model;
  constexpr int NXI = C::NXI, LXI = C::LXI;
  if constexpr (!Model::HAS_NEWTON) {
    Model::guess(k0, E.xip, E.par, md.abs_tol, xi);
#pragma unroll
    for (int q = 0; q < NXI; ++q) Cd[q] = make_dual<LXI>(0.0);
    return 0;
  } else {
    int path = 0, iter = 1;
    double R_norm_0 = 1.0;
    bool converged = false;
    // R_norm_0 of the relative test is the residual norm at the REFERENCE's starting point
    // (src/small_J2.cpp:147-151).  Where the Newton starts at a closed-form predictor instead, the
    // model returns that norm with the guess (it is |f| of the trial state, which the predictor
    // needs anyway), so a deck converges by rel_tol exactly when it does in the reference.
    bool have_r0 = false;
    if constexpr (has_predictor<Model>::value) {
      const double r0 = Model::guess_r0(k0, E.xip, E.par, md.abs_tol, xi);
      have_r0 = r0 >= 0.0;
      R_norm_0 = have_r0 ? r0 : 1.0;
    } else {
      Model::guess(k0, E.xip, E.par, md.abs_tol, xi);
    }
    // WARP-UNIFORM body: every lane evaluates the residual and takes part in the solve (full-mask
    // shuffles); lanes that are done re-evaluate at their unchanged xi, which reproduces the same
    // Cd / path, and simply do not apply the update.
    while (true) {
      const bool work = active && (iter <= md.max_iters) && !converged;
      // warp-uniform loop: a warp leaves when its own quadrature points are done (the phase
      // barriers of the caller re-align the CTA); -DC8_K1_BLOCK_NEWTON keeps the whole CTA together
#ifdef C8_K1_BLOCK_NEWTON
      if (!__syncthreads_or(work)) break;
#else
      if (!__any_sync(0xffffffffu, work)) break;
#endif
      Dual<LXI> xs[NXI];
#pragma unroll
      for (int q = 0; q < NXI; ++q) xs[q] = seeded<LXI>(xi[q], q, t * LXI);
      path = Model::residual(k0, xs, E.xip, E.par, md.abs_tol, Cd);
      double nrm = 0.0;
#pragma unroll
      for (int q = 0; q < NXI; ++q) nrm += Cd[q].v * Cd[q].v;
      const double R_norm = sqrt(nrm);
      if (work && iter == 1 && !have_r0) R_norm_0 = R_norm;
      const double R_norm_rel = R_norm / R_norm_0;
      const bool conv_now = (R_norm_rel < md.rel_tol) || (R_norm < md.abs_tol);
      const bool update = work && !conv_now;
      if (work && conv_now) converged = true;
      if (__any_sync(0xffffffffu, update)) {
        double Jc[NXI][LXI], rhs[NXI], dummy[NXI][1];
#pragma unroll
        for (int q = 0; q < NXI; ++q) {
          rhs[q] = -Cd[q].v;
#pragma unroll
          for (int s = 0; s < LXI; ++s) Jc[q][s] = Cd[q].d[s];
        }
        group_gauss_jordan<NXI, LXI, 0, C::G, true>(Jc, dummy, rhs, mask);
#pragma unroll
        for (int q = 0; q < NXI; ++q) xi[q] = pick(update, xi[q] + rhs[q], xi[q]);
      }
      if (update) ++iter;
    }
    if (active && !converged) return -1;
    return path;
  }
}

// dxi/dx = -(dC/dxi)^-1 dC/dx for the columns this thread owns (Bc in: dC/dx, out: dxi/dx)
// path: the branch of this point (used at warp-uniform call sites only): when every point of the warp
// is elastic and the model's elastic Jacobian is the identity, -(I)^-1 B = -B needs no solve (the
// reference's LU of an identity matrix is exact, so the result is bit-identical).
template <class C, int LB, bool FULL = false>
C8_DI void local_sensitivity(const Dual<C::LXI> (&Cd)[C::NXI], double (&Bc)[C::NXI][LB],
                             unsigned mask, int path = -1) {
  if constexpr (C::Model::HAS_NEWTON && C::Model::ELASTIC_J_IDENTITY && FULL) {
    if (__all_sync(0xffffffffu, path == PATH_ELASTIC)) {
#pragma unroll
      for (int q = 0; q < C::NXI; ++q)
#pragma unroll
        for (int s = 0; s < LB; ++s) Bc[q][s] = -Bc[q][s];
      return;
    }
  }
  if constexpr (!C::Model::HAS_NEWTON) {
    // Elastic: C == 0 identically, the reference's LU of a zero matrix returns 0
#pragma unroll
    for (int q = 0; q < C::NXI; ++q)
#pragma unroll
      for (int s = 0; s < LB; ++s) Bc[q][s] = 0.0;
  } else {
    double Jc[C::NXI][C::LXI], dummy[C::NXI];
#pragma unroll
    for (int q = 0; q < C::NXI; ++q) {
      dummy[q] = 0.0;
#pragma unroll
      for (int s = 0; s < C::LXI; ++s) Jc[q][s] = Cd[q].d[s];
#pragma unroll
      for (int s = 0; s < LB; ++s) Bc[q][s] = -Bc[q][s];
    }
    group_gauss_jordan<C::NXI, C::LXI, LB, C::G, FULL>(Jc, Bc, dummy, mask);
  }
}

// Seeded kinematics at the coupled point: grad u (Dual), p, grad p  (x seeded)
template <class C>
struct SeededX {
  static constexpr int L = C::LX;
  XLanes<C::D, C::NB, L> xl;
  Mat<Dual<L>, C::D> gu;
  Dual<L> p;
  Dual<L> gp[C::D];
  C8_DI void init(const Elem<C>& E, const Mat<double, C::D>& gu0, int t) {
    xl.init(t * L, E.g);
    gu = grad_u_seeded<C::D, C::NB, L>(gu0, xl);
    if constexpr (C::M == MECH_MIXED) {
      const double Nc = 1.0 / C::NN;
      double pv = 0.0;
#pragma unroll
      for (int n = 0; n < C::NN; ++n) pv += E.xn[n][C::D] * Nc;
      p.v = pv;
#pragma unroll
      for (int s = 0; s < L; ++s) p.d[s] = (xl.eq[s] == C::D) ? Nc : 0.0;
#pragma unroll
      for (int j = 0; j < C::D; ++j) {
        double gv = 0.0;
#pragma unroll
        for (int n = 0; n < C::NN; ++n) gv += E.xn[n][C::D] * E.g.gN[n][j];
        gp[j].v = gv;
#pragma unroll
        for (int s = 0; s < L; ++s) gp[j].d[s] = (xl.eq[s] == C::D) ? xl.gsel[s][j] : 0.0;
      }
    } else {
      p = make_dual<L>(0.0);
    }
  }
};

// Element row output (value + this thread's derivative lanes).
//
// Two-phase assembly into the fixed BSR pattern, replacing Tpetra's sumIntoLocalValues of the
// reference (src/global_residual.cpp:556-586):
//   phase 1 (here)        every thread group writes its element matrix row by row into a scratch
//                         [n_elems+1][NX][NX] with plain vector stores (one base pointer, immediate
//                         offsets; thread groups that must not contribute write to the sink slot);
//   phase 2 (k_bsr_gather) every BSR block sums the element-matrix sub-blocks listed in the
//                         precomputed gather plan, in element order -> bit-reproducible sums, no
//                         fp64 atomics (256 per tet were the top stall of K1, profiles/README.md),
//                         rows of ghost nodes (another part's rows) are simply not gathered.
// The residual vector keeps reductions (16 per tet): each thread first collects the entries it
// owns (values are replicated over the group), then adds them once in finish().
// FAST: b is given and no element-level output is wanted (the production case, branch-free).
// STAGE (persistent K1, vector-store layouts only): the element matrix rows go to a shared-memory slot of
// the CTA's output buffer instead (st.shared), and the whole 2 KB element matrix leaves with ONE bulk
// asynchronous copy (cp.async.bulk) after the tile.  ncu on the direct path: the 32 STG.128 of a thread
// funnel through the same 4 data registers, and the instruction that next overwrites them waits for the
// store unit to have read them -- 15 % of all warp-stall samples of the kernel sat on that dependency.
template <class C, bool FAST = false, bool STAGE = false>
struct Scatter {
  const FwdArgs& a;
  const Elem<C>& E;
  const XLanes<C::D, C::NB, C::LX>& xl;
  int e, t;
  bool on;  // false: compute but store nothing (padding groups, failed local solves)
  unsigned stage_addr = 0;  // STAGE: shared-memory address of this thread's columns in its element slot
  double* em = nullptr;    // this thread's columns of the element matrix: em[row*NX + s]
  static constexpr int NBACC = (C::NX + C::G - 1) / C::G;
  double bacc[NBACC];
  C8_DI void init() {
    constexpr int NX = C::NX;
    if (a.vals != nullptr)   // slot of the element in the scratch; the slot after the last one is the sink
      em = a.emat + size_t(on ? e - a.elem_begin : a.elem_end - a.elem_begin) * NX * NX + t * C::LX;
#pragma unroll
    for (int j = 0; j < NBACC; ++j) bacc[j] = 0.0;
  }
  // after the last row(): one reduction per kept residual entry
  C8_DI void finish() const {
    constexpr int NB = C::NB, NN = C::NN, NX = C::NX;
    if (FAST || a.b != nullptr) {
#pragma unroll
      for (int j = 0; j < NBACC; ++j) {
        const int dof = t + j * C::G;
        const int dofc = dof < NX ? dof : NX - 1;
        const int n = dofc / NB, eq = dofc - n * NB;
        int gn = 0;
#pragma unroll
        for (int n2 = 0; n2 < NN; ++n2) gn = picki(n == n2, E.nodes[n2], gn);
        const bool keep = on && dof < NX && gn < a.mesh.n_row_nodes;
        red_add(&a.b[size_t(gn) * NB + eq], pick(keep, bacc[j], 0.0));
      }
    }
  }
  C8_DI void row(int n, int eq, const Dual<C::LX>& r) {
    constexpr int NB = C::NB, NX = C::NX, LX = C::LX;
    const int row_dof = n * NB + eq;
    bacc[row_dof / C::G] = pick((row_dof % C::G) == t, r.v, bacc[row_dof / C::G]);
    if constexpr (STAGE) {
      static_assert(LX % 2 == 0 && C::G * LX == NX, "staged output needs the vector-store layout");
      const unsigned dst = stage_addr + unsigned(row_dof * NX * sizeof(double));
#pragma unroll
      for (int s = 0; s < LX; s += 2)
        asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(dst + unsigned(s * sizeof(double))), "d"(r.d[s]),
                     "d"(r.d[s + 1]) : "memory");
    } else if (FAST || a.vals != nullptr) {
      double* dst = em + row_dof * NX;
      if constexpr (LX % 2 == 0 && C::G * LX == NX) {
#pragma unroll
        for (int s = 0; s < LX; s += 2)
          *reinterpret_cast<double2*>(dst + s) = make_double2(r.d[s], r.d[s + 1]);
      } else {
#pragma unroll
        for (int s = 0; s < LX; ++s)
          if (t * LX + s < NX) dst[s] = r.d[s];
      }
    }
    if constexpr (!FAST) {
      if (on && a.elem_R && (row_dof % C::G) == t) a.elem_R[size_t(e) * NX + C::ref_dof(n, eq)] = r.v;
#pragma unroll
      for (int s = 0; s < LX; ++s)
        if (on && a.elem_J && xl.nsel[s] != 0.0)
          a.elem_J[(size_t(e) * NX + C::ref_dof(n, eq)) * NX + C::ref_dof(xl.node[s], xl.eq[s])] = r.d[s];
    }
  }
};

// Phase 2 of the assembly: vals[block] = sum over the gather plan of the element sub-blocks
// (TRANSPOSE: the transposed sub-block of the transposed pair -> dtotal^T for the adjoint).
// Rows of nodes >= n_row_nodes are left untouched.  Non-transposed blocks with an even NB are
// gathered two entries per thread (16-byte loads / stores); the additions stay in plan order, so
// the sum is bit-reproducible.
template <int NB, int NN, bool TRANSPOSE>
__global__ void k_bsr_gather(const int* __restrict__ gptr, const int* __restrict__ gsrc,
                             const double* __restrict__ emat, double* __restrict__ vals,
                             int n_row_blocks) {
  constexpr int BB = NB * NB, NX = NB * NN;
  constexpr bool VEC2 = (NB % 2 == 0);
  constexpr int EPT = VEC2 ? 2 : 1;                 // entries per thread
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= (long long)n_row_blocks * (BB / EPT)) return;
  const int blk = int(i / (BB / EPT)), ent = int(i % (BB / EPT)) * EPT;
  // the pair of entries a thread owns is contiguous in the ELEMENT matrix: (r, c), (r, c+1) of the
  // output block when not transposing; (r, c), (r+1, c) when transposing
  const int r = TRANSPOSE ? ent % NB : ent / NB, c = TRANSPOSE ? ent / NB : ent % NB;
  auto src = [&](int k) -> const double* {
    const int q = __ldg(&gsrc[k]);           // e*NN*NN + na*NN + nb : block (node na, node nb)
    const int e = q / (NN * NN), rem = q - e * (NN * NN);
    const int na = rem / NN, nb = rem - na * NN;
    const double* m = emat + size_t(e) * NX * NX;
    return TRANSPOSE ? &m[(nb * NB + c) * NX + na * NB + r] : &m[(na * NB + r) * NX + nb * NB + c];
  };
  int k = gptr[blk];
  const int k1 = gptr[blk + 1];
  if constexpr (VEC2) {
    double s0 = 0.0, s1 = 0.0;
    for (; k + 2 <= k1; k += 2) {
      const double2 v0 = __ldg(reinterpret_cast<const double2*>(src(k)));
      const double2 v1 = __ldg(reinterpret_cast<const double2*>(src(k + 1)));
      s0 += v0.x; s1 += v0.y; s0 += v1.x; s1 += v1.y;
    }
    for (; k < k1; ++k) {
      const double2 v = __ldg(reinterpret_cast<const double2*>(src(k)));
      s0 += v.x; s1 += v.y;
    }
    if constexpr (TRANSPOSE) {
      vals[size_t(blk) * BB + r * NB + c] = s0;
      vals[size_t(blk) * BB + (r + 1) * NB + c] = s1;
    } else {
      *reinterpret_cast<double2*>(vals + size_t(blk) * BB + ent) = make_double2(s0, s1);
    }
  } else {
    double s = 0.0;
    for (; k < k1; ++k) s += __ldg(src(k));
    vals[size_t(blk) * BB + r * NB + c] = s;
  }
}

// Phase 2 for ONE CHUNK of elements [e0, e0 + chunk): every plan entry is a (block, contribution
// range) pair of the chunk; the first chunk that touches a block overwrites it, later ones accumulate
// (chunks run in element order and the contributions inside an entry are in element order, so the sum
// has the same order as the one-pass gather: bit-identical values).  emat holds the chunk only.
template <int NB, int NN>
__global__ void k_bsr_gather_chunk(const unsigned* __restrict__ cblk, const int* __restrict__ ck,
                                   const int* __restrict__ gsrc, const double* __restrict__ emat,
                                   double* __restrict__ vals, int entry0, int n_entries, int e0) {
  constexpr int BB = NB * NB, NX = NB * NN;
  static_assert(NB % 2 == 0, "chunked gather: even block size");
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= (long long)n_entries * (BB / 2)) return;
  const int en = entry0 + int(i / (BB / 2)), ent = int(i % (BB / 2)) * 2;
  const int r = ent / NB, c = ent % NB;
  const unsigned bw = __ldg(&cblk[en]);
  const int blk = int(bw & 0x7fffffffu);
  const bool first = (bw >> 31) != 0u;
  double2* dst = reinterpret_cast<double2*>(vals + size_t(blk) * BB + ent);
  double s0 = 0.0, s1 = 0.0;
  if (!first) { const double2 v = *dst; s0 = v.x; s1 = v.y; }
  const int k1 = __ldg(&ck[2 * en + 1]);
  for (int k = __ldg(&ck[2 * en]); k < k1; ++k) {
    const int q = __ldg(&gsrc[k]);
    const int e = q / (NN * NN), rem = q - e * (NN * NN);
    const int na = rem / NN, nb = rem - na * NN;
    const double2 v = *reinterpret_cast<const double2*>(emat + size_t(e - e0) * NX * NX + (na * NB + r) * NX + nb * NB + c);
    s0 += v.x; s1 += v.y;
  }
  *dst = make_double2(s0, s1);
}

// Optional per-phase cycle counters (tuning builds only: -DC8_K1_PHASE_CLOCKS)
#ifdef C8_K1_PHASE_CLOCKS
static __device__ unsigned long long g_k1_phase_clocks[8];
#define C8_PHASE_MARK(k)                                                        \
  do {                                                                          \
    const long long now_ = clock64();                                           \
    if ((threadIdx.x & 31) == 0) atomicAdd(&g_k1_phase_clocks[k], (unsigned long long)(now_ - tphase_)); \
    tphase_ = now_;                                                             \
  } while (0)
#define C8_PHASE_START() long long tphase_ = clock64()
#else
#define C8_PHASE_MARK(k) do {} while (0)
#define C8_PHASE_START() do {} while (0)
#endif

// CTA-wide barriers between the phases kept the warps of a CTA close together in the long straight-line
// program of the one-tile-per-CTA kernel (instruction-cache locality; measured then: hyper-J2 3.03 vs
// 3.57 ms / 1M tets with / without).  The persistent kernel already meets twice per tile, and there the five
// phase barriers cost more than they give (2.034 vs 2.011 ms), so they are OFF by default;
// C8_K1_SYNC_MASK selects which stay (bit k = k-th barrier in program order, finite-strain models only).
// They must be off when an SM holds more than one persistent CTA (the 128-thread teams of the small-strain models).
#ifndef C8_K1_SYNC_MASK
#define C8_K1_SYNC_MASK 0
#endif
#define C8_PHASE_SYNC_K(k) do { if constexpr (C::Model::FINITE && ((C8_K1_SYNC_MASK >> (k)) & 1)) __syncthreads(); } while (0)

#ifndef C8_K1_BLOCK
#define C8_K1_BLOCK 256
#endif
#ifndef C8_K1_MINB
#define C8_K1_MINB 1
#endif

// Everything of K1 after the element record E and the gathered current-field xi are in registers:
// kinematics, the three AD passes, the element rows and their stores.  Shared by the one-tile-per-CTA
// kernel and the persistent, prefetching one below.
// STAGE / stage_addr: see Scatter.  Returns true when the element's matrix was produced (in range, local
// solve converged).
template <class C, bool FAST, bool STAGE = false>
C8_DI bool k1_element(const FwdArgs& a, const Elem<C>& E, double (&xi)[C::NXI], int e, int t, bool in_range,
                      unsigned mask, unsigned stage_addr = 0) {
  using Model = typename C::Model;
  constexpr int D = C::D, NN = C::NN, NB = C::NB, NXI = C::NXI, LX = C::LX, G = C::G;
  C8_PHASE_START();
  Kin<D, double, double> k0;
  k0.gu = grad_u_val<D, NB>(E.xn, E.g);
  k0.gup = grad_u_val<D, NB>(E.xpn, E.g);
  if constexpr (needs_rotation<Model>::value) cache_rotation(k0);   // one polar rotation for the whole local Newton
  C8_PHASE_MARK(0);

  // ---- P1: local Newton (block-synchronous) ---------------------------------
  Dual<C::LXI> Cd[NXI];
  const int path = local_newton<C>(k0, E, a.model, xi, Cd, mask, t, in_range);
  const bool ok = in_range && path >= 0;
  if (in_range && path < 0) {
    // the reference aborts the assembly (src/evaluations.cpp:95-97): report, scatter nothing
    if (t == 0) atomicAdd(a.n_failed, 1);
    if (a.path && t == 0) a.path[e] = -1;
    // the element contributes nothing: clear its scratch slot, so that the gather does not sum stale
    // data of an earlier call into A when the caller ignores the status
    if (a.vals != nullptr) {
      double* em = a.emat + size_t(e - a.elem_begin) * C::NX * C::NX + t * LX;
      for (int r = 0; r < C::NX; ++r)
        for (int s = 0; s < LX; ++s)
          if (t * LX + s < C::NX) em[r * C::NX + s] = 0.0;
    }
  }
  if (ok) {
#pragma unroll
    for (int q = 0; q < NXI; ++q)
      if (q % G == t) a.xi[size_t(q) * a.xi_ld + e] = xi[q];
    if (a.path && t == 0) a.path[e] = (signed char)path;
  }
  C8_PHASE_SYNC_K(0);
  C8_PHASE_MARK(1);

  // ---- P2: dC/dx, then dxi/dx ---------------------------------------------
  SeededX<C> sx;
  sx.init(E, k0.gu, t);
  Kin<D, Dual<LX>, double> k2;
  k2.gu = sx.gu;
  k2.gup = k0.gup;
  Dual<LX> xid[NXI];
  // With one thread per node (G == NN, LX == NB) in the mixed formulation, lane LX-1 of every thread
  // is the node's PRESSURE dof, and the constitutive residual does not depend on the pressure: P2 and
  // the sensitivity solve run on the LX-1 displacement lanes only, dxi/dp = 0 exactly.
  constexpr bool DROP_P_LANE = (C::M == MECH_MIXED) && (G == NN) && (LX == NB);
  constexpr int L2 = DROP_P_LANE ? LX - 1 : LX;
  {
    Kin<D, Dual<L2>, double> k2n;
    k2n.gup = k0.gup;
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
      for (int j = 0; j < D; ++j) {
        k2n.gu(i, j).v = k2.gu(i, j).v;
#pragma unroll
        for (int s = 0; s < L2; ++s) k2n.gu(i, j).d[s] = k2.gu(i, j).d[s];
      }
    if constexpr (needs_rotation<Model>::value) {
      // ONE rotation under AD per point: R depends on the displacement dofs only, so the width-L2 result
      // also serves P3 (its extra pressure lane carries a zero derivative)
      cache_rotation(k2n);
#pragma unroll
      for (int i = 0; i < D; ++i)
#pragma unroll
        for (int j = 0; j < D; ++j) {
          k2.rot(i, j).v = k2n.rot(i, j).v;
#pragma unroll
          for (int s = 0; s < L2; ++s) k2.rot(i, j).d[s] = k2n.rot(i, j).d[s];
#pragma unroll
          for (int s = L2; s < LX; ++s) k2.rot(i, j).d[s] = 0.0;
        }
      k2.has_rot = true;
    }
    Dual<L2> C2[NXI];
    Model::residual(k2n, xi, E.xip, E.par, a.model.abs_tol, C2);
    C8_PHASE_SYNC_K(1);
    double Bc[NXI][L2];
#pragma unroll
    for (int q = 0; q < NXI; ++q)
#pragma unroll
      for (int s = 0; s < L2; ++s) Bc[q][s] = C2[q].d[s];
    local_sensitivity<C, L2, true>(Cd, Bc, mask, in_range ? path : PATH_ELASTIC);
#pragma unroll
    for (int q = 0; q < NXI; ++q) {
      xid[q].v = xi[q];
#pragma unroll
      for (int s = 0; s < L2; ++s) xid[q].d[s] = Bc[q][s];
#pragma unroll
      for (int s = L2; s < LX; ++s) xid[q].d[s] = 0.0;
    }
  }
  C8_PHASE_SYNC_K(2);
  C8_PHASE_MARK(2);

  // ---- P3: element residual and total Jacobian, scattered row by row ------
  const double wdv = quad1_weight<D>() * E.g.dv;
  Scatter<C, FAST, STAGE> sc{a, E, sx.xl, e, t, ok};
  sc.stage_addr = stage_addr;
  sc.init();
  {
    const Mat<Dual<LX>, D> P = first_pk<D, C::M, Model>(k2, sx.p, xid, E.par, a.model.thickness);
    C8_PHASE_SYNC_K(3);
#pragma unroll
    for (int n = 0; n < NN; ++n)
#pragma unroll
      for (int i = 0; i < D; ++i) {
        Dual<LX> r = P(i, 0) * (E.g.gN[n][0] * wdv);
#pragma unroll
        for (int j = 1; j < D; ++j) r += P(i, j) * (E.g.gN[n][j] * wdv);
        sc.row(n, i, r);
      }
  }
  C8_PHASE_MARK(3);
  if constexpr (C::M == MECH_MIXED) {
    C8_PHASE_SYNC_K(4);
    Dual<LX> Rp[NN];
    {
      Dual<LX> hp, sv[D];
      pressure_terms<D, Model>(k2, sx.gp, xid, E.par, E.g.h, a.model.stab_mult, hp, sv);
      const double Nc = 1.0 / NN;
#pragma unroll
      for (int n = 0; n < NN; ++n) {
        Dual<LX> r = hp * (Nc * wdv);
#pragma unroll
        for (int i = 0; i < D; ++i) r += sv[i] * (E.g.gN[n][i] * wdv);
        Rp[n] = -r;
      }
    }
    // pressure-mass ip set (quadrature order 2), src/mechanics.cpp:206-215
    const double ipk = 1.0 / Model::pscale(E.par);
#pragma unroll
    for (int q = 0; q < Quad2<D>::NPT; ++q) {
      double N[NN];
      Quad2<D>::basis(q, N);
      Dual<LX> pq;
      pq.v = 0.0;
#pragma unroll
      for (int n = 0; n < NN; ++n) pq.v += E.xn[n][D] * N[n];
#pragma unroll
      for (int s = 0; s < LX; ++s) {
        double Ns = 0.0;
#pragma unroll
        for (int n = 0; n < NN; ++n) Ns = pick(sx.xl.node[s] == n, N[n], Ns);
        pq.d[s] = (sx.xl.eq[s] == D) ? Ns : 0.0;
      }
      const double wq = Quad2<D>::weight() * E.g.dv;
      const Dual<LX> pk = pq * ipk;
#pragma unroll
      for (int n = 0; n < NN; ++n) Rp[n] -= pk * (N[n] * wq);
    }
#pragma unroll
    for (int n = 0; n < NN; ++n) sc.row(n, D, Rp[n]);
  }
  sc.finish();
  C8_PHASE_MARK(4);
  return ok;
}

template <class C, bool FAST>
__global__ void __launch_bounds__(C8_K1_BLOCK, C8_K1_MINB) k_forward_jacobian(const FwdArgs a) {
  constexpr int NXI = C::NXI, G = C::G;
  const int gid = blockIdx.x * blockDim.x + threadIdx.x;
  const int t = gid % G;
  const bool in_range = a.elem_begin + (gid / G) < a.elem_end;
  // out-of-range groups recompute the last element (no stores) so that every thread of the CTA
  // reaches the phase barriers of k1_element
  const int e = in_range ? a.elem_begin + gid / G : a.elem_end - 1;
  const unsigned lane = threadIdx.x & 31u;
  const unsigned mask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (lane / G * G));
#ifdef C8_K1_SMEM_ELEM
  // element record in shared memory, one copy per group (see load_elem_shared)
  __shared__ Elem<C> sE[C8_K1_BLOCK / G];
  load_elem_shared<C>(a.mesh, a.model, a.x, a.x_prev, a.xi_prev, a.xi_ld, e, t, sE[threadIdx.x / G]);
  __syncwarp();
  const Elem<C>& E = sE[threadIdx.x / G];
#else
  Elem<C> E;
  load_elem<C>(a.mesh, a.model, a.x, a.x_prev, a.xi_prev, a.xi_ld, e, E);
#endif
  double xi[NXI];
#pragma unroll
  for (int q = 0; q < NXI; ++q) xi[q] = a.xi[size_t(q) * a.xi_ld + e];
  k1_element<C, FAST>(a, E, xi, e, t, in_range, mask);
}

// ---- persistent K1: one CTA per SM loops over tiles of C8_K1_BLOCK / G elements and PREFETCHES -------------
// At 255 registers one CTA of 8 warps is resident per SM, all of them in the same phase (phase barriers),
// so the dependent gathers at the start of a tile -- connectivity -> nodal coords / x / x_prev, plus the
// xi_prev / xi rows -- are exposed latency (the "load" phase: 14 % of the warp cycles of the one-tile
// kernel).  Here tile k+1's record is copied into shared memory with cp.async while tile k computes
// (nodal rows gathered through the connectivity that was itself prefetched one tile earlier), and a tile
// starts by reading its record from shared memory into registers.
C8_DI void cp_async16(void* smem, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
C8_DI void cp_async8(void* smem, const void* gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
C8_DI void cp_async4(void* smem, const void* gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
C8_DI void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
C8_DI void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// row stride (in doubles) >= n with stride % 4 == 2: the 32 / G groups of a warp then read the same entry of
// their own rows from distinct bank pairs, and rows stay 16-byte aligned for the vector copies
constexpr int k1_row_stride(int n) { return n + ((2 - n % 4) + 4) % 4; }

// Threads per CTA of the persistent K1 = the team that walks the tile list together.  Per model: the
// finite-strain programs (~140 KB of straight-line code) want ONE 256-thread CTA per SM in lockstep (two
// 128-thread CTAs drift apart and thrash the instruction caches: hyper-J2 2.37 against 2.01 ms), the
// small-strain programs are about half as long and run faster as TWO independent 128-thread CTAs per SM,
// each hiding the other's barrier and latency stalls (small-J2 1.35 against 1.47 ms, elastic 0.94 against 1.00).
// A model overrides the default with K1_TEAM (3-D small-Hill with its return-map predictor: 1.88 at 256 against
// 1.91 at 2 x 128).  C8_K1_PBLOCK forces one size for every model.
template <class M, class = void> struct k1_team_of { static constexpr int value = M::FINITE ? 256 : 128; };
template <class M> struct k1_team_of<M, decltype(void(M::K1_TEAM))> { static constexpr int value = M::K1_TEAM; };
template <class C>
struct K1PBlock {
#ifdef C8_K1_PBLOCK
  static constexpr int value = C8_K1_PBLOCK;
#else
  static constexpr int value = k1_team_of<typename C::Model>::value;
#endif
};

template <class C, int TEAM>
struct alignas(16) K1Stage {
  static constexpr int EPB = TEAM / C::G;   // elements per tile
  static constexpr int SX = k1_row_stride(C::NN * C::NB), SC = k1_row_stride(C::NN * C::D);
  double xn[EPB][SX];      // [element][node * NB + eq]
  double xpn[EPB][SX];
  double xip[C::NXI][EPB];
  double xi[C::NXI][EPB];
  double X[EPB][SC];       // [element][node * D + k]
  int nodes[EPB][C::NN];
};
// element matrices of a tile staged for the bulk stores: one slot per element, 16 bytes of padding per slot
// so that the 16-byte row stores of the two thread groups of a quarter warp fall into different banks
template <class C, int TEAM>
struct K1Out {
  static constexpr bool STAGED = (C::LX % 2 == 0) && (C::G * C::LX == C::NX) && ((C::NX * C::NX * 8) % 16 == 0);
  static constexpr int SLOT = C::NX * C::NX * 8 + 16;
  static constexpr int BYTES = STAGED ? K1Stage<C, TEAM>::EPB * SLOT : 16;
};
template <class C, int TEAM, bool WITH_OUT = true>
struct alignas(16) K1Smem {
  K1Stage<C, TEAM> stage[2];
  int conn[2][K1Stage<C, TEAM>::EPB][C::NN];
  alignas(16) unsigned char out[WITH_OUT ? K1Out<C, TEAM>::BYTES : 16];
  int out_ok[K1Stage<C, TEAM>::EPB];
  int fetch[2];
};
template <class C>
constexpr int k1_smem_bytes() { return int(sizeof(K1Smem<C, K1PBlock<C>::value>)); }

// The asynchronous copies of a tile's element records (shared by the persistent K1 and K3): connectivity
// of a tile into conn[buf], then -- once that has landed -- the nodal rows it names and the tile's local-state
// rows into stage[st].
template <class C, int TEAM, class SMEM>
struct TilePrefetch {
  static constexpr int D = C::D, NN = C::NN, NB = C::NB, NXI = C::NXI;
  static constexpr int EPB = K1Stage<C, TEAM>::EPB;
  SMEM& S;
  const int* conn;
  const double *coords, *x, *x_prev, *xi_prev, *xi;
  long long xi_ld;
  int elem_begin, elem_end, tid;
  // 16-byte copies need 16-byte aligned arrays (any cudaMalloc'd array is; a caller's offset pointer may not be)
  C8_DI bool aligned16() const {
    return ((reinterpret_cast<size_t>(x) | reinterpret_cast<size_t>(x_prev) | reinterpret_cast<size_t>(xi_prev) |
             reinterpret_cast<size_t>(xi)) & 15) == 0;
  }
  // element of slot j of tile `tile` (clamped: padding slots repeat the last element, no stores)
  C8_DI int elem_of(int tile, int j) const {
    const int e = elem_begin + tile * EPB + j;
    return e < elem_end ? e : elem_end - 1;
  }
  C8_DI void issue_conn(int tile, int buf) const {
    for (int i = tid; i < EPB * NN; i += TEAM) {
      const int j = i / NN, n = i - j * NN;
      cp_async4(&S.conn[buf][j][n], &conn[size_t(elem_of(tile, j)) * NN + n]);
    }
  }
  C8_DI void issue_record(int tile, int buf, int st) const {
    K1Stage<C, TEAM>& T = S.stage[st];
    const bool have_xp = x_prev != nullptr;
    const bool a16 = aligned16();
    for (int i = tid; i < EPB * NN; i += TEAM) {
      const int j = i / NN, n = i - j * NN;
      const int nd = S.conn[buf][j][n];
      T.nodes[j][n] = nd;
#pragma unroll
      for (int k = 0; k < D; ++k) cp_async8(&T.X[j][n * D + k], &coords[size_t(nd) * D + k]);
      if (NB % 2 == 0 && a16) {
#pragma unroll
        for (int q = 0; q + 1 < NB; q += 2) {
          cp_async16(&T.xn[j][n * NB + q], &x[size_t(nd) * NB + q]);
          if (have_xp) cp_async16(&T.xpn[j][n * NB + q], &x_prev[size_t(nd) * NB + q]);
        }
      } else {
#pragma unroll
        for (int q = 0; q < NB; ++q) {
          cp_async8(&T.xn[j][n * NB + q], &x[size_t(nd) * NB + q]);
          if (have_xp) cp_async8(&T.xpn[j][n * NB + q], &x_prev[size_t(nd) * NB + q]);
        }
      }
    }
    // local state rows: EPB consecutive elements per component (a full tile is 16-byte aligned: elem_begin
    // and xi_ld are multiples of 32); the ragged last tile goes element by element
    const int e0 = elem_begin + tile * EPB;
    if (e0 + EPB <= elem_end && (e0 % 2) == 0 && a16) {
      for (int i = tid; i < NXI * (EPB / 2); i += TEAM) {
        const int q = i / (EPB / 2), j = (i - q * (EPB / 2)) * 2;
        cp_async16(&T.xip[q][j], &xi_prev[size_t(q) * xi_ld + e0 + j]);
        cp_async16(&T.xi[q][j], &xi[size_t(q) * xi_ld + e0 + j]);
      }
    } else {
      for (int i = tid; i < NXI * EPB; i += TEAM) {
        const int q = i / EPB, j = i - q * EPB;
        const int e = elem_of(tile, j);
        cp_async8(&T.xip[q][j], &xi_prev[size_t(q) * xi_ld + e]);
        cp_async8(&T.xi[q][j], &xi[size_t(q) * xi_ld + e]);
      }
    }
  }
};

// The tile loop of a persistent team: tiles are handed out dynamically (the cost of a tile depends on how
// many of its points yield) -- the first three of a team are static, every further one comes from `counter`;
// thread 0 of the team fetches it one iteration ahead of its first use (the connectivity prefetch runs two
// tiles ahead).  before(k): runs ahead of the barrier that opens iteration k (e.g. wait for the bulk stores
// of the previous tile); body(tile, st): the tile's record is in stage[st].
template <class C, int TEAM, class PF, class SMEM, class BEFORE, class BODY>
C8_DI void persistent_tile_loop(const PF& pf, SMEM& S, int* counter, int n_tiles, int team_id, int n_teams, int tid,
                                BEFORE before, BODY body) {
  auto team_sync = [] { __syncthreads(); };   // a team is a whole CTA
  int tile = team_id, next = team_id + n_teams, next2 = team_id + 2 * n_teams;
  if (tile >= n_tiles) return;   // team-uniform (a CTA-wide team leaves as a whole)
  if (tid == 0) S.fetch[0] = 3 * n_teams + atomicAdd(counter, 1);
  // prologue: connectivity of the first tile, then its record and the connectivity of the second
  pf.issue_conn(tile, 0);
  cp_async_commit();
  cp_async_wait_all();
  team_sync();
  pf.issue_record(tile, 0, 0);
  if (next < n_tiles) pf.issue_conn(next, 1);
  cp_async_commit();
#pragma unroll 1
  for (int k = 0; tile < n_tiles; ++k) {
    const int st = k & 1;
    cp_async_wait_all();
    before(k);
    team_sync();   // record of this tile and connectivity of the next are in shared memory; every
                   // thread is done with the other stage (previous tile)
    const int fetched = S.fetch[st];
    if (tid == 0)   // once a team has seen the end of the tile list it stops drawing (the list is handed out in order)
      S.fetch[st ^ 1] = (next2 < n_tiles && fetched < n_tiles) ? 3 * n_teams + atomicAdd(counter, 1) : n_tiles;
    if (next < n_tiles) {
      pf.issue_record(next, st ^ 1, st ^ 1);
      // conn buffer `st` held THIS tile's connectivity, consumed when its record was issued
      if (next2 < n_tiles) pf.issue_conn(next2, st);
    }
    cp_async_commit();
    body(tile, st);
    tile = next; next = next2; next2 = fetched;
  }
}

template <class C, bool FAST>
__global__ void __launch_bounds__(K1PBlock<C>::value, 256 / K1PBlock<C>::value) k_forward_jacobian_persistent(const FwdArgs a) {
  constexpr int D = C::D, NN = C::NN, NB = C::NB, NXI = C::NXI, G = C::G;
  constexpr int TEAM = K1PBlock<C>::value;
  static_assert(C8_K1_SYNC_MASK == 0 || TEAM == 256, "phase barriers: every CTA of an SM must be one team");
  constexpr int EPB = K1Stage<C, TEAM>::EPB;
  constexpr bool STAGED = K1Out<C, TEAM>::STAGED && FAST;
  using SMEM = K1Smem<C, TEAM>;
  extern __shared__ __align__(16) unsigned char k1_smem_raw[];
  const int tid = threadIdx.x;
  SMEM& S = *reinterpret_cast<SMEM*>(k1_smem_raw);
  auto team_sync = [] { __syncthreads(); };
  const int n_teams = gridDim.x;
  const int team_id = blockIdx.x;
  const int t = tid % G, gl = tid / G;
  const unsigned lane = threadIdx.x & 31u;
  const unsigned mask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (lane / G * G));
  const int n_range = a.elem_end - a.elem_begin;
  const int n_tiles = (n_range + EPB - 1) / EPB;
  const int last = a.elem_end - 1;
  const bool have_xp = a.x_prev != nullptr;
  const TilePrefetch<C, TEAM, SMEM> pf{S, a.mesh.conn, a.mesh.coords, a.x, a.x_prev, a.xi_prev, a.xi, a.xi_ld,
                                      a.elem_begin, a.elem_end, tid};
  auto before = [&](int) __attribute__((always_inline)) {
    if constexpr (STAGED) {   // the bulk stores of the previous tile have read their shared-memory slots
      if (tid < EPB) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
  };
  auto body = [&](int tile, int st) __attribute__((always_inline)) {
    const K1Stage<C, TEAM>& T = S.stage[st];
    const int slot = tile * EPB + gl;
    const bool in_range = slot < n_range;
    const int e = in_range ? a.elem_begin + slot : last;
    Elem<C> E;
    {
      double X[NN][D];
#pragma unroll
      for (int n = 0; n < NN; ++n) {
        E.nodes[n] = T.nodes[gl][n];
#pragma unroll
        for (int c = 0; c < D; ++c) X[n][c] = T.X[gl][n * D + c];
#pragma unroll
        for (int q = 0; q < NB; ++q) {
          E.xn[n][q] = T.xn[gl][n * NB + q];
          E.xpn[n][q] = have_xp ? T.xpn[gl][n * NB + q] : 0.0;
        }
      }
      geom_from_coords<D>(X, E.g);
    }
    const int es = a.mesh.elem_es ? __ldg(&a.mesh.elem_es[e]) : 0;
#pragma unroll
    for (int q = 0; q < C::NPAR; ++q) E.par[q] = __ldg(&a.model.params[es * a.model.npar + q]);
    double xi[NXI];
#pragma unroll
    for (int q = 0; q < NXI; ++q) { E.xip[q] = T.xip[q][gl]; xi[q] = T.xi[q][gl]; }
    if constexpr (STAGED) {
      const unsigned slot_addr = (unsigned)__cvta_generic_to_shared(S.out) + unsigned(gl * K1Out<C, TEAM>::SLOT);
      const bool ok = k1_element<C, FAST, true>(a, E, xi, e, t, in_range, mask,
                                                slot_addr + unsigned(t * C::LX * sizeof(double)));
      if (t == 0) S.out_ok[gl] = ok ? 1 : 0;
      // generic-proxy writes of the slots -> visible to the async proxy, then one bulk store per element
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      team_sync();
      if (tid < EPB && S.out_ok[tid]) {
        const int ee = a.elem_begin + tile * EPB + tid;     // ok implies in range
        double* dst = a.emat + size_t(ee - a.elem_begin) * C::NX * C::NX;
        const unsigned src = (unsigned)__cvta_generic_to_shared(S.out) + unsigned(tid * K1Out<C, TEAM>::SLOT);
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src),
                     "n"(C::NX * C::NX * 8) : "memory");
      }
      if (tid < EPB) asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    } else {
      k1_element<C, FAST>(a, E, xi, e, t, in_range, mask);
    }
  };
  persistent_tile_loop<C, TEAM>(pf, S, a.n_failed + 1, n_tiles, team_id, n_teams, tid, before, body);
  if constexpr (STAGED) {
    if (tid < EPB) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
}

// K2: residual only (eval_global_residual): no Newton, xi given, T = double
template <class C>
__global__ void __launch_bounds__(128) k_global_residual(const FwdArgs a) {
  using Model = typename C::Model;
  constexpr int D = C::D, NN = C::NN, NB = C::NB, NXI = C::NXI;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= a.mesh.n_elems) return;
  Elem<C> E;
  load_elem<C>(a.mesh, a.model, a.x, a.x_prev, a.xi_prev, a.xi_ld, e, E);
  double xi[NXI];
#pragma unroll
  for (int q = 0; q < NXI; ++q) xi[q] = a.xi[size_t(q) * a.xi_ld + e];
  Kin<D, double, double> k0;
  k0.gu = grad_u_val<D, NB>(E.xn, E.g);
  k0.gup = grad_u_val<D, NB>(E.xpn, E.g);
  if constexpr (needs_rotation<Model>::value) cache_rotation(k0);
  const double wdv = quad1_weight<D>() * E.g.dv;
  double p = 0.0, gp[D];
  if constexpr (C::M == MECH_MIXED) {
#pragma unroll
    for (int n = 0; n < NN; ++n) p += E.xn[n][D] * (1.0 / NN);
#pragma unroll
    for (int j = 0; j < D; ++j) {
      gp[j] = 0.0;
#pragma unroll
      for (int n = 0; n < NN; ++n) gp[j] += E.xn[n][D] * E.g.gN[n][j];
    }
  }
  const Mat<double, D> P = first_pk<D, C::M, Model>(k0, p, xi, E.par, a.model.thickness);
#pragma unroll
  for (int n = 0; n < NN; ++n)
#pragma unroll
    for (int i = 0; i < D; ++i) {
      double r = 0.0;
#pragma unroll
      for (int j = 0; j < D; ++j) r += P(i, j) * (E.g.gN[n][j] * wdv);
      if (E.nodes[n] < a.mesh.n_row_nodes) atomicAdd(&a.b[size_t(E.nodes[n]) * NB + i], r);
    }
  if constexpr (C::M == MECH_MIXED) {
    double hp, sv[D], Rp[NN];
    pressure_terms<D, Model>(k0, gp, xi, E.par, E.g.h, a.model.stab_mult, hp, sv);
#pragma unroll
    for (int n = 0; n < NN; ++n) {
      double r = hp * ((1.0 / NN) * wdv);
#pragma unroll
      for (int i = 0; i < D; ++i) r += sv[i] * (E.g.gN[n][i] * wdv);
      Rp[n] = -r;
    }
    const double ipk = 1.0 / Model::pscale(E.par);
#pragma unroll
    for (int q = 0; q < Quad2<D>::NPT; ++q) {
      double N[NN];
      Quad2<D>::basis(q, N);
      double pq = 0.0;
#pragma unroll
      for (int n = 0; n < NN; ++n) pq += E.xn[n][D] * N[n];
      const double wq = Quad2<D>::weight() * E.g.dv;
#pragma unroll
      for (int n = 0; n < NN; ++n) Rp[n] -= pq * ipk * (N[n] * wq);
    }
#pragma unroll
    for (int n = 0; n < NN; ++n)
      if (E.nodes[n] < a.mesh.n_row_nodes) atomicAdd(&a.b[size_t(E.nodes[n]) * NB + D], Rp[n]);
  }
}

}  // namespace c8
