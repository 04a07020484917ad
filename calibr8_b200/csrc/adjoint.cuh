// K3-K6: the reverse-in-time adjoint element loops of the reference,
//   K3 eval_adjoint_jacobian  src/evaluations.cpp:349-526  (A^T, rhs, g -= dJ/dxi)
//   K4 solve_adjoint_local    src/evaluations.cpp:528-659  (phi, f, g)
//   K5 eval_qoi / preprocess_qoi  src/evaluations.cpp:662-756, 261-347
//   K6 eval_qoi_gradient      src/evaluations.cpp:758-925  (dC/dp^T phi + dJ/dp + dR/dp^T z)
// Same thread-group / lane-split AD design as forward.cuh.  History arrays g [NXI][ld],
// f [NX][ld], phi [NXI][ld] are structure-of-arrays like xi.
#pragma once
#include "forward.cuh"
#include "qoi.cuh"

// minimum resident CTAs per SM asked of ptxas for K3 / K4 / K6 (tuning: tools/build_variant.sh -DC8_K6_MINB=3).
// K6 at 4 (128 registers, 300 B of spills): 0.520 -> 0.422 ms at 1 M hyper-J2 tets; K4 gains nothing from 3.
#ifndef C8_K3_MINB
#define C8_K3_MINB 1
#endif
#ifndef C8_K4_MINB
#define C8_K4_MINB 1
#endif
#ifndef C8_K6_MINB
#define C8_K6_MINB 4
#endif

namespace c8 {


template <class C> C8_DI unsigned group_mask() {
  const unsigned lane = threadIdx.x & 31u;
  return (C::G == 32) ? 0xffffffffu : (((1u << C::G) - 1u) << (lane / C::G * C::G));
}

// nodal displacements as Dual<L> seeded w.r.t. the element dofs (for QoI derivatives)
template <class C, int L>
C8_DI void seeded_nodal_u(const Elem<C>& E, const XLanes<C::D, C::NB, L>& xl,
                          Dual<L> (&un)[C::NN][C::D]) {
#pragma unroll
  for (int n = 0; n < C::NN; ++n)
#pragma unroll
    for (int i = 0; i < C::D; ++i) {
      un[n][i].v = E.xn[n][i];
#pragma unroll
      for (int s = 0; s < L; ++s)
        un[n][i].d[s] = (xl.nsel[s] != 0.0 && xl.node[s] == n && xl.eq[s] == i) ? 1.0 : 0.0;
    }
}

template <class C>
C8_DI void load_measured(const QoiArgs& q, const Elem<C>& E, double (&um)[C::NN][C::D],
                         double (&X)[C::NN][C::D], const double* coords) {
#pragma unroll
  for (int n = 0; n < C::NN; ++n)
#pragma unroll
    for (int i = 0; i < C::D; ++i) {
      um[n][i] = q.measured ? __ldg(&q.measured[size_t(E.nodes[n]) * C::D + i]) : 0.0;
      X[n][i] = __ldg(&coords[size_t(E.nodes[n]) * C::D + i]);
    }
}

// ---------------------------------------------------------------------------------------
// K3
// Register diet (round 2; the first version kept 255 registers and a 2.2 KB spill frame, 7.5 GB of
// local-memory traffic per launch at 1 M tets):
//   * the element record (connectivity, geometry, nodal x / x_prev, parameters, xi_prev: ~65 doubles
//     that every thread of the group held redundantly) lives in shared memory, one copy per group;
//   * the passes are ordered so that each one's temporaries die before the next starts: QoI
//     xi-derivatives (needs the xi-seeded state) -> sensitivity solve -> (dxi/dx)^T g and the
//     right-hand side -> element matrix rows last.
// everything of K3 after the element record E (in shared memory) and the stored xi are at hand
template <class C>
C8_DI void k3_element(const AdjArgs& a, const Elem<C>& E, const double (&xi)[C::NXI], int e, int t, bool in_range,
                      unsigned mask) {
  using Model = typename C::Model;
  constexpr int D = C::D, NN = C::NN, NB = C::NB, NX = C::NX, NXI = C::NXI, LX = C::LX,
                LXI = C::LXI, G = C::G;
  Kin<D, double, double> k0;
  k0.gu = grad_u_val<D, NB>(E.xn, E.g);
  k0.gup = grad_u_val<D, NB>(E.xpn, E.g);
  if constexpr (needs_rotation<Model>::value) cache_rotation(k0);
  const double wdv = quad1_weight<D>() * E.g.dv;
  const bool calib = a.qoi.type != QOI_AVG_DISP;
  const int nmask = a.qoi.has_node_load() ? load_node_mask<D>(a.qoi, a.mesh.coords, E.nodes) : 0;
  const signed char* fv = a.qoi.facet ? &a.qoi.facet[size_t(e) * 3] : nullptr;
  const bool fload = a.qoi.has_face_load() && fv && fv[0] >= 0;   // load mismatch: facet on the side set
  const double coef = a.qoi.balance_factor * a.qoi.dt_over_T * a.qoi.load_mismatch;
  double fN[D], fwdv = 0.0;
#pragma unroll
  for (int k = 0; k < D; ++k) fN[k] = 0.0;
  if (fload) facet_normal<D>(a.qoi, a.mesh.coords, E.nodes, fv, fN, fwdv);

  // ---- xi seeded: dC/dxi at the stored state (no Newton, src/evaluations.cpp:442-446) and the QoI's
  // dJ/dxi (:474-478): g -= dJ/dxi
  double gq[NXI];
#pragma unroll
  for (int q = 0; q < NXI; ++q) gq[q] = a.g[size_t(q) * a.xi_ld + e];
  Dual<LXI> Cd[NXI];
  int path;
  {
    Dual<LXI> xs[NXI];
#pragma unroll
    for (int q = 0; q < NXI; ++q) xs[q] = seeded<LXI>(xi[q], q, t * LXI);
    path = Model::residual(k0, xs, E.xip, E.par, a.model.abs_tol, Cd);
    if (nmask || fload) {   // load term: coordinate-plane nodes or side-set facet (group-uniform branch)
      double p0 = 0.0;
      if constexpr (C::M == MECH_MIXED) {
        if (nmask) {
#pragma unroll
          for (int n = 0; n < NN; ++n) p0 += E.xn[n][D] * (1.0 / NN);
        } else {
          p0 = facet_pressure<D, NB>(E.xn, fv);
        }
      }
      const Mat<Dual<LXI>, D> Pxi = first_pk<D, C::M, Model>(k0, p0, xs, E.par, nmask ? a.model.thickness : 1.0);
      const Dual<LXI> load = nmask ? calibration_load<D>(a.qoi, Pxi, E.g, wdv, nmask, a.mesh.coords, E.nodes)
                                   : face_normal_load<D>(Pxi, fN, fwdv);
#pragma unroll
      for (int q = 0; q < NXI; ++q) {
        const double v = group_bcast<G>(mask, load.d[q % LXI], q / LXI);
        gq[q] -= coef * v;
      }
    }
  }
  if (in_range) {
#pragma unroll
    for (int q = 0; q < NXI; ++q)
      if (q % G == t) a.g[size_t(q) * a.xi_ld + e] = gq[q];
  }

  // ---- x seeded: dC/dx, dxi/dx = -(dC/dxi)^-1 dC/dx
  SeededX<C> sx;
  sx.init(E, k0.gu, t);
  Kin<D, Dual<LX>, double> k2;
  k2.gu = sx.gu;
  k2.gup = k0.gup;
  Dual<LX> xid[NXI];
  {
    // as in K1 (forward.cuh): the pressure lane carries no constitutive derivative, and an
    // all-elastic warp needs no solve
    constexpr bool DROP_P_LANE = (C::M == MECH_MIXED) && (G == NN) && (LX == NB);
    constexpr int L2 = DROP_P_LANE ? LX - 1 : LX;
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
    if constexpr (needs_rotation<Model>::value) {   // one rotation under AD for this pass and the next (see K1)
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
      for (int s = 0; s < LX; ++s) xid[q].d[s] = s < L2 ? Bc[q][s < L2 ? s : 0] : 0.0;
    }
  }

  // ---- rhs = -dJ/dx + f + dxi/dx^T g (:481-488); dJ/dx with x seeded and xi NOT seeded (:468-472)
  {
    double r[LX];
#pragma unroll
    for (int s = 0; s < LX; ++s) {
      double v = 0.0;
#pragma unroll
      for (int q = 0; q < NXI; ++q) v = fma(xid[q].d[s], gq[q], v);
      r[s] = v;
    }
    if (!calib) {
#pragma unroll
      for (int s = 0; s < LX; ++s) r[s] -= (sx.xl.eq[s] < D) ? (1.0 / NN) * wdv / D : 0.0;
    } else {
      const bool in_obj = a.qoi.has_disp() && (D == 2 ? a.qoi.type == QOI_CALIBRATION : (fv && fv[0] >= 0));
      if (in_obj) {
        Dual<LX> un[NN][D];
        seeded_nodal_u<C, LX>(E, sx.xl, un);
        double um[NN][D], X[NN][D];
        load_measured<C>(a.qoi, E, um, X, a.mesh.coords);
        const Dual<LX> mm = calibration_disp_mismatch<D, Dual<LX>>(a.qoi, un, um, X, E.g.dv, (D == 3) ? fv : nullptr);
#pragma unroll
        for (int s = 0; s < LX; ++s) r[s] -= mm.d[s];
      }
      if (nmask) {
        const Mat<Dual<LX>, D> Px = first_pk<D, C::M, Model>(k2, sx.p, xi, E.par, a.model.thickness);
        const Dual<LX> load = calibration_load<D>(a.qoi, Px, E.g, wdv, nmask, a.mesh.coords, E.nodes);
#pragma unroll
        for (int s = 0; s < LX; ++s) r[s] -= coef * load.d[s];
      }
      if (fload) {
        // the pressure at the facet centroid, seeded w.r.t. the pressure dofs of the facet's nodes
        Dual<LX> pf = make_dual<LX>(facet_pressure<D, NB>(E.xn, fv));
#pragma unroll
        for (int s = 0; s < LX; ++s) {
          const int nd = sx.xl.node[s];
          const bool in_f = fv[0] == nd || fv[1] == nd || (D == 3 && fv[2] == nd);
          pf.d[s] = (sx.xl.eq[s] == D && in_f) ? 1.0 / D : 0.0;
        }
        const Mat<Dual<LX>, D> Px = first_pk<D, C::M, Model>(k2, pf, xi, E.par, 1.0);
        const Dual<LX> load = face_normal_load<D>(Px, fN, fwdv);
#pragma unroll
        for (int s = 0; s < LX; ++s) r[s] -= coef * load.d[s];
      }
    }
    if (in_range) {
#pragma unroll
      for (int s = 0; s < LX; ++s) {
        if (sx.xl.nsel[s] == 0.0) continue;
        const int c = t * LX + s;
        const double v = r[s] + __ldg(&a.f[size_t(c) * a.xi_ld + e]);
        if (E.nodes[sx.xl.node[s]] < a.mesh.n_row_nodes)
          atomicAdd(&a.b[size_t(E.nodes[sx.xl.node[s]]) * NB + sx.xl.eq[s]], v);
      }
    }
  }

  // ---- total Jacobian -> element-matrix scratch; the gather pass transposes (matrix only)
  FwdArgs fa{};
  fa.mesh = a.mesh;
  fa.vals = a.vals;
  fa.emat = a.emat;
  fa.elem_begin = 0; fa.elem_end = a.mesh.n_elems;   // one pass over the whole mesh (full scratch)
  Scatter<C, false> sc{fa, E, sx.xl, e, t, in_range};  // element matrix only; the rhs was added above
  sc.init();
  {
    const Mat<Dual<LX>, D> P = first_pk<D, C::M, Model>(k2, sx.p, xid, E.par, a.model.thickness);
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
  if constexpr (C::M == MECH_MIXED) {
    Dual<LX> Rp[NN];
    {
      Dual<LX> hp, sv[D];
      pressure_terms<D, Model>(k2, sx.gp, xid, E.par, E.g.h, a.model.stab_mult, hp, sv);
#pragma unroll
      for (int n = 0; n < NN; ++n) {
        Dual<LX> r = hp * ((1.0 / NN) * wdv);
#pragma unroll
        for (int i = 0; i < D; ++i) r += sv[i] * (E.g.gN[n][i] * wdv);
        Rp[n] = -r;
      }
    }
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
}

template <class C>
__global__ void __launch_bounds__(128, C8_K3_MINB) k_adjoint_jacobian(const AdjArgs a) {
  constexpr int NXI = C::NXI, G = C::G;
  const int gid = blockIdx.x * blockDim.x + threadIdx.x;
  const int t = gid % G;
  // padding groups recompute the last element and store nothing, so that every lane of a warp
  // reaches the in-group solves (full-mask shuffles, see groupsolve.cuh)
  const bool in_range = (gid / G) < a.mesh.n_elems;
  const int e = in_range ? gid / G : a.mesh.n_elems - 1;
  const unsigned mask = group_mask<C>();
  __shared__ Elem<C> sE[128 / G];
  load_elem_shared<C>(a.mesh, a.model, a.x, a.x_prev, a.xi_prev, a.xi_ld, e, t, sE[threadIdx.x / G]);
  __syncwarp();   // a group never spans two warps
  const Elem<C>& E = sE[threadIdx.x / G];
  double xi[NXI];
#pragma unroll
  for (int q = 0; q < NXI; ++q) xi[q] = __ldg(&a.xi[size_t(q) * a.xi_ld + e]);
  k3_element<C>(a, E, xi, e, t, in_range, mask);
}

// Persistent K3, the K1 treatment (forward.cuh): one CTA per SM walks the tile list and prefetches the next
// tile's element records with cp.async; a tile starts by building its Elem records in shared memory from the
// prefetched stage (the threads of a group split the copy, the last one computes the geometry).  ncu on the
// one-tile kernel: 30 % of the stall samples long-scoreboard on that load, 18 % no-instruction with two
// 128-thread CTAs per SM drifting apart in the program.
template <class C>
struct alignas(16) K3Smem {
  K1Smem<C, C8_K1_BLOCK, false> pf;
  Elem<C> sE[C8_K1_BLOCK / C::G];
};

template <class C>
__global__ void __launch_bounds__(C8_K1_BLOCK, 1) k_adjoint_jacobian_persistent(const AdjArgs a) {
  constexpr int D = C::D, NN = C::NN, NB = C::NB, NXI = C::NXI, G = C::G;
  constexpr int TEAM = C8_K1_BLOCK;
  constexpr int EPB = K1Stage<C, TEAM>::EPB;
  using SMEM = K1Smem<C, TEAM, false>;
  extern __shared__ __align__(16) unsigned char k3_smem_raw[];
  K3Smem<C>& S3 = *reinterpret_cast<K3Smem<C>*>(k3_smem_raw);
  SMEM& S = S3.pf;
  const int tid = threadIdx.x;
  const int t = tid % G, gl = tid / G;
  const unsigned mask = group_mask<C>();
  const int n_range = a.mesh.n_elems;
  const int n_tiles = (n_range + EPB - 1) / EPB;
  const bool have_xp = a.x_prev != nullptr;
  const TilePrefetch<C, TEAM, SMEM> pf{S, a.mesh.conn, a.mesh.coords, a.x, a.x_prev, a.xi_prev, a.xi, a.xi_ld,
                                      0, a.mesh.n_elems, tid};
  auto before = [](int) {};
  auto body = [&](int tile, int st) __attribute__((always_inline)) {
    const K1Stage<C, TEAM>& T = S.stage[st];
    const int slot = tile * EPB + gl;
    const bool in_range = slot < n_range;
    const int e = in_range ? slot : n_range - 1;
    Elem<C>& Ew = S3.sE[gl];
    for (int n = t; n < NN; n += G) {
      Ew.nodes[n] = T.nodes[gl][n];
#pragma unroll
      for (int q = 0; q < NB; ++q) {
        Ew.xn[n][q] = T.xn[gl][n * NB + q];
        Ew.xpn[n][q] = have_xp ? T.xpn[gl][n * NB + q] : 0.0;
      }
    }
    const int es = a.mesh.elem_es ? __ldg(&a.mesh.elem_es[e]) : 0;
    for (int q = t; q < C::NPAR; q += G) Ew.par[q] = __ldg(&a.model.params[es * a.model.npar + q]);
    for (int q = t; q < NXI; q += G) Ew.xip[q] = T.xip[q][gl];
    if (t == G - 1) {
      double X[NN][D];
#pragma unroll
      for (int n = 0; n < NN; ++n)
#pragma unroll
        for (int c = 0; c < D; ++c) X[n][c] = T.X[gl][n * D + c];
      geom_from_coords<D>(X, Ew.g);
    }
    double xi[NXI];
#pragma unroll
    for (int q = 0; q < NXI; ++q) xi[q] = T.xi[q][gl];
    __syncwarp();   // a group never spans two warps
    k3_element<C>(a, S3.sE[gl], xi, e, t, in_range, mask);
  };
  persistent_tile_loop<C, TEAM>(pf, S, a.tile_counter, n_tiles, (int)blockIdx.x, (int)gridDim.x, tid, before, body);
}

// ---------------------------------------------------------------------------------------
// K4
template <class C>
__global__ void __launch_bounds__(128, C8_K4_MINB) k_adjoint_local(const AdjArgs a) {
  using Model = typename C::Model;
  constexpr int D = C::D, NN = C::NN, NB = C::NB, NXI = C::NXI, LX = C::LX, LXI = C::LXI,
                G = C::G;
  const int gid = blockIdx.x * blockDim.x + threadIdx.x;
  const int t = gid % G;
  const bool in_range = (gid / G) < a.mesh.n_elems;   // padding groups: compute, store nothing
  const int e = in_range ? gid / G : a.mesh.n_elems - 1;
  const unsigned mask = group_mask<C>();

  Elem<C> E;
  load_elem<C>(a.mesh, a.model, a.x, a.x_prev, a.xi_prev, a.xi_ld, e, E);
  double xi[NXI];
#pragma unroll
  for (int q = 0; q < NXI; ++q) xi[q] = __ldg(&a.xi[size_t(q) * a.xi_ld + e]);
  Kin<D, double, double> k0;
  k0.gu = grad_u_val<D, NB>(E.xn, E.g);
  k0.gup = grad_u_val<D, NB>(E.xpn, E.g);
  if constexpr (needs_rotation<Model>::value) cache_rotation(k0);
  const double wdv = quad1_weight<D>() * E.g.dv;
  double z[NN][NB];
#pragma unroll
  for (int n = 0; n < NN; ++n)
#pragma unroll
    for (int q = 0; q < NB; ++q) z[n][q] = __ldg(&a.z[size_t(E.nodes[n]) * NB + q]);

  // xi seeded: dC/dxi and dR/dxi (ip set 0 only, :613-621)
  Dual<LXI> xs[NXI], Cd[NXI];
#pragma unroll
  for (int q = 0; q < NXI; ++q) xs[q] = seeded<LXI>(xi[q], q, t * LXI);
  Model::residual(k0, xs, E.xip, E.par, a.model.abs_tol, Cd);
  double rl[LXI];  // (dR/dxi^T z) for this thread's xi lanes
#pragma unroll
  for (int s = 0; s < LXI; ++s) rl[s] = 0.0;
  {
    double p0 = 0.0, gp0[D];
    if constexpr (C::M == MECH_MIXED) {
#pragma unroll
      for (int n = 0; n < NN; ++n) p0 += E.xn[n][D] * (1.0 / NN);
#pragma unroll
      for (int j = 0; j < D; ++j) {
        gp0[j] = 0.0;
#pragma unroll
        for (int n = 0; n < NN; ++n) gp0[j] += E.xn[n][D] * E.g.gN[n][j];
      }
    }
    const Mat<Dual<LXI>, D> P = first_pk<D, C::M, Model>(k0, p0, xs, E.par, a.model.thickness);
#pragma unroll
    for (int n = 0; n < NN; ++n)
#pragma unroll
      for (int i = 0; i < D; ++i) {
        Dual<LXI> r = P(i, 0) * (E.g.gN[n][0] * wdv);
#pragma unroll
        for (int j = 1; j < D; ++j) r += P(i, j) * (E.g.gN[n][j] * wdv);
#pragma unroll
        for (int s = 0; s < LXI; ++s) rl[s] = fma(r.d[s], z[n][i], rl[s]);
      }
    if constexpr (C::M == MECH_MIXED) {
      Dual<LXI> hp, sv[D];
      pressure_terms<D, Model>(k0, gp0, xs, E.par, E.g.h, a.model.stab_mult, hp, sv);
#pragma unroll
      for (int n = 0; n < NN; ++n) {
        Dual<LXI> r = hp * ((1.0 / NN) * wdv);
#pragma unroll
        for (int i = 0; i < D; ++i) r += sv[i] * (E.g.gN[n][i] * wdv);
#pragma unroll
        for (int s = 0; s < LXI; ++s) rl[s] = fma(-r.d[s], z[n][D], rl[s]);
      }
    }
  }
  // rhs = g - dR/dxi^T z (replicated), J^T by an in-group transpose, phi = J^-T rhs
  double rhs[NXI], JT[NXI][LXI], dummy[NXI][1];
#pragma unroll
  for (int q = 0; q < NXI; ++q) {
    const double v = group_bcast_full<G>(rl[q % LXI], q / LXI);
    rhs[q] = __ldg(&a.g[size_t(q) * a.xi_ld + e]) - v;
  }
  // thread owns columns c = t*LXI+s of J (Cd[i].d[s] = J[i][c]); it needs columns c of J^T,
  // i.e. JT[i][s] = J[c][i]: broadcast every J[r][i] and keep those with r == c
#pragma unroll
  for (int i = 0; i < NXI; ++i) {
#pragma unroll
    for (int s = 0; s < LXI; ++s) JT[i][s] = 0.0;
  }
#pragma unroll
  for (int r = 0; r < NXI; ++r)
#pragma unroll
    for (int i = 0; i < NXI; ++i) {
      const double v = group_bcast_full<G>(Cd[r].d[i % LXI], i / LXI);  // J[r][i]
#pragma unroll
      for (int s = 0; s < LXI; ++s) JT[i][s] = pick(t * LXI + s == r, v, JT[i][s]);
    }
  if constexpr (Model::HAS_NEWTON) {
    group_gauss_jordan<NXI, LXI, 0, G, true>(JT, dummy, rhs, mask);  // rhs := phi
  } else {
    // Elastic: C == 0 identically, dC/dxi is the zero matrix and the reference's rank-revealing LU
    // (Eigen::FullPivLU, src/evaluations.cpp:624) returns phi = 0 -- as in local_sensitivity()
#pragma unroll
    for (int q = 0; q < NXI; ++q) rhs[q] = 0.0;
  }
  if (!in_range) return;  // no shuffles below
#pragma unroll
  for (int q = 0; q < NXI; ++q)
    if (q % G == t) a.phi[size_t(q) * a.xi_ld + e] = rhs[q];

  // f = -dC/dx_prev^T phi (x_prev seeded; non-zero only for finite-strain models), :626-632
  {
    XLanes<D, NB, LX> xl;
    xl.init(t * LX, E.g);
    Kin<D, double, Dual<LX>> kp;
    kp.gu = k0.gu;
    kp.gup = grad_u_seeded<D, NB, LX>(k0.gup, xl);
    Dual<LX> Cp[NXI];
    Model::residual(kp, xi, E.xip, E.par, a.model.abs_tol, Cp);
#pragma unroll
    for (int s = 0; s < LX; ++s) {
      if (xl.nsel[s] == 0.0) continue;
      double v = 0.0;
#pragma unroll
      for (int q = 0; q < NXI; ++q) v = fma(Cp[q].d[s], rhs[q], v);
      a.f[size_t(t * LX + s) * a.xi_ld + e] = -v;
    }
  }
  // g = -dC/dxi_prev^T phi (xi_prev seeded), :634-641
  {
    Dual<LXI> xps[NXI], Cq[NXI];
#pragma unroll
    for (int q = 0; q < NXI; ++q) xps[q] = seeded<LXI>(E.xip[q], q, t * LXI);
    Model::residual(k0, xi, xps, E.par, a.model.abs_tol, Cq);
#pragma unroll
    for (int s = 0; s < LXI; ++s) {
      const int c = t * LXI + s;
      if (c >= NXI) continue;
      double v = 0.0;
#pragma unroll
      for (int q = 0; q < NXI; ++q) v = fma(Cq[q].d[s], rhs[q], v);
      a.g[size_t(c) * a.xi_ld + e] = -v;
    }
  }
}

// ---------------------------------------------------------------------------------------
// K6: every parameter of the model is seeded (lane k = parameter k); the host picks the
// active ones (LocalResidual::seed_wrt_params seeds only those, src/local_residual.cpp:811-819;
// lanes are independent so the active lanes are identical).
template <class C>
__global__ void __launch_bounds__(128, C8_K6_MINB) k_qoi_gradient(const AdjArgs a) {
  using Model = typename C::Model;
  constexpr int D = C::D, NN = C::NN, NB = C::NB, NXI = C::NXI, NPAR = C::NPAR, G = C::G;
  constexpr int LP = (NPAR + G - 1) / G;
  using TP = Dual<LP>;
  __shared__ double sacc[128 / G][G * LP];  // per group in the block
  const int gid = blockIdx.x * blockDim.x + threadIdx.x;
  const int e = gid / G, t = gid % G;
  const int gl = threadIdx.x / G;
  double acc[LP];
#pragma unroll
  for (int s = 0; s < LP; ++s) acc[s] = 0.0;
  int es = 0;
  if (e < a.mesh.n_elems) {
    Elem<C> E;
    load_elem<C>(a.mesh, a.model, a.x, a.x_prev, a.xi_prev, a.xi_ld, e, E);
    es = a.mesh.elem_es ? a.mesh.elem_es[e] : 0;
    double xi[NXI], phi[NXI];
#pragma unroll
    for (int q = 0; q < NXI; ++q) {
      xi[q] = __ldg(&a.xi[size_t(q) * a.xi_ld + e]);
      phi[q] = __ldg(&a.phi[size_t(q) * a.xi_ld + e]);
    }
    Kin<D, double, double> k0;
    k0.gu = grad_u_val<D, NB>(E.xn, E.g);
    k0.gup = grad_u_val<D, NB>(E.xpn, E.g);
    if constexpr (needs_rotation<Model>::value) cache_rotation(k0);
    const double wdv = quad1_weight<D>() * E.g.dv;
    double z[NN][NB];
#pragma unroll
    for (int n = 0; n < NN; ++n)
#pragma unroll
      for (int q = 0; q < NB; ++q) z[n][q] = __ldg(&a.z[size_t(E.nodes[n]) * NB + q]);
    TP par[NPAR];
#pragma unroll
    for (int k = 0; k < NPAR; ++k) par[k] = seeded<LP>(E.par[k], k, t * LP);
    // dC/dp^T phi
    {
      TP Cp[NXI];
      Model::residual(k0, xi, E.xip, par, a.model.abs_tol, Cp);
#pragma unroll
      for (int q = 0; q < NXI; ++q)
#pragma unroll
        for (int s = 0; s < LP; ++s) acc[s] = fma(Cp[q].d[s], phi[q], acc[s]);
    }
    // dR/dp^T z (all ip sets) and dJ/dp (calibration load term)
    double p0 = 0.0, gp0[D];
    if constexpr (C::M == MECH_MIXED) {
#pragma unroll
      for (int n = 0; n < NN; ++n) p0 += E.xn[n][D] * (1.0 / NN);
#pragma unroll
      for (int j = 0; j < D; ++j) {
        gp0[j] = 0.0;
#pragma unroll
        for (int n = 0; n < NN; ++n) gp0[j] += E.xn[n][D] * E.g.gN[n][j];
      }
    }
    {
      const Mat<TP, D> P = first_pk<D, C::M, Model>(k0, p0, xi, par, a.model.thickness);
#pragma unroll
      for (int n = 0; n < NN; ++n)
#pragma unroll
        for (int i = 0; i < D; ++i) {
          TP r = P(i, 0) * (E.g.gN[n][0] * wdv);
#pragma unroll
          for (int j = 1; j < D; ++j) r += P(i, j) * (E.g.gN[n][j] * wdv);
#pragma unroll
          for (int s = 0; s < LP; ++s) acc[s] = fma(r.d[s], z[n][i], acc[s]);
        }
      const double coef = a.qoi.balance_factor * a.qoi.dt_over_T * a.qoi.load_mismatch;
      if (a.qoi.has_node_load()) {
        const int nmask = load_node_mask<D>(a.qoi, a.mesh.coords, E.nodes);
        if (nmask) {
          const TP load = calibration_load<D>(a.qoi, P, E.g, wdv, nmask, a.mesh.coords, E.nodes);
#pragma unroll
          for (int s = 0; s < LP; ++s) acc[s] += coef * load.d[s];
        }
      }
      if (a.qoi.has_face_load() && a.qoi.facet && a.qoi.facet[size_t(e) * 3] >= 0) {
        const signed char* fv = &a.qoi.facet[size_t(e) * 3];
        double fN[D], fwdv;
        facet_normal<D>(a.qoi, a.mesh.coords, E.nodes, fv, fN, fwdv);
        const Mat<TP, D> Pf = first_pk<D, C::M, Model>(k0, facet_pressure<D, NB>(E.xn, fv), xi, par, 1.0);
        const TP load = face_normal_load<D>(Pf, fN, fwdv);
#pragma unroll
        for (int s = 0; s < LP; ++s) acc[s] += coef * load.d[s];
      }
    }
    if constexpr (C::M == MECH_MIXED) {
      TP hp, sv[D];
      pressure_terms<D, Model>(k0, gp0, xi, par, E.g.h, a.model.stab_mult, hp, sv);
#pragma unroll
      for (int n = 0; n < NN; ++n) {
        TP r = hp * ((1.0 / NN) * wdv);
#pragma unroll
        for (int i = 0; i < D; ++i) r += sv[i] * (E.g.gN[n][i] * wdv);
#pragma unroll
        for (int s = 0; s < LP; ++s) acc[s] = fma(-r.d[s], z[n][D], acc[s]);
      }
      const TP ipk = 1.0 / Model::pscale(par);
#pragma unroll
      for (int q = 0; q < Quad2<D>::NPT; ++q) {
        double N[NN];
        Quad2<D>::basis(q, N);
        double pq = 0.0;
#pragma unroll
        for (int n = 0; n < NN; ++n) pq += E.xn[n][D] * N[n];
        const double wq = Quad2<D>::weight() * E.g.dv;
        const TP pk = pq * ipk;
#pragma unroll
        for (int n = 0; n < NN; ++n) {
          const TP r = pk * (N[n] * wq);
#pragma unroll
          for (int s = 0; s < LP; ++s) acc[s] = fma(-r.d[s], z[n][D], acc[s]);
        }
      }
    }
  }
  // block reduction per element set is not needed when the block is single-set (the common
  // case); otherwise fall back to direct global atomics
  const bool multi_set = a.mesh.elem_es != nullptr;
  if (multi_set) {
    if (e < a.mesh.n_elems) {
#pragma unroll
      for (int s = 0; s < LP; ++s) {
        const int k = t * LP + s;
        if (k < NPAR) atomicAdd(&a.grad[es * NPAR + k], acc[s]);
      }
    }
    return;
  }
#pragma unroll
  for (int s = 0; s < LP; ++s) sacc[gl][t * LP + s] = acc[s];
  __syncthreads();
  if (threadIdx.x < G * LP) {
    double v = 0.0;
    for (int g2 = 0; g2 < 128 / G; ++g2) v += sacc[g2][threadIdx.x];
    if (threadIdx.x < NPAR) atomicAdd(&a.grad[threadIdx.x], v);
  }
}

// ---------------------------------------------------------------------------------------
// K5: QoI value (and the calibration preprocess pass: total load on the coordinate plane).
// T = double; one thread per element; block reduction then one atomic per block.
//   mode 0: scalars[0] += sum of per-point QoI values   (eval_qoi main loop)
//   mode 1: scalars[1] += load on the plane             (preprocess_qoi / Calibration::preprocess)
template <class C>
__global__ void __launch_bounds__(128) k_qoi_value(const AdjArgs a, int mode) {
  using Model = typename C::Model;
  constexpr int D = C::D, NN = C::NN, NB = C::NB, NXI = C::NXI;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  double v = 0.0;
  if (e < a.mesh.n_elems) {
    Elem<C> E;
    load_elem<C>(a.mesh, a.model, a.x, a.x_prev, a.xi_prev, a.xi_ld, e, E);
    const double wdv = quad1_weight<D>() * E.g.dv;
    if (mode == 0 && a.qoi.type == QOI_AVG_DISP) {
      double s = 0.0;
#pragma unroll
      for (int i = 0; i < D; ++i) {
        double u = 0.0;
#pragma unroll
        for (int n = 0; n < NN; ++n) u += E.xn[n][i] * (1.0 / NN);
        s += u * wdv;
      }
      v = s / D;
    } else if (mode == 0) {
      const signed char* fv0 = a.qoi.facet ? &a.qoi.facet[size_t(e) * 3] : nullptr;
      const bool in_obj = a.qoi.has_disp() && (D == 2 ? a.qoi.type == QOI_CALIBRATION : (fv0 && fv0[0] >= 0));
      if (in_obj) {
        double un[NN][D], um[NN][D], X[NN][D];
#pragma unroll
        for (int n = 0; n < NN; ++n)
#pragma unroll
          for (int i = 0; i < D; ++i) un[n][i] = E.xn[n][i];
        load_measured<C>(a.qoi, E, um, X, a.mesh.coords);
        v = calibration_disp_mismatch<D, double>(a.qoi, un, um, X, E.g.dv, (D == 3) ? fv0 : nullptr);
      }
    } else {
      const int nmask = a.qoi.has_node_load() ? load_node_mask<D>(a.qoi, a.mesh.coords, E.nodes) : 0;
      const signed char* fv = a.qoi.facet ? &a.qoi.facet[size_t(e) * 3] : nullptr;
      const bool fload = a.qoi.has_face_load() && fv && fv[0] >= 0;
      if (nmask || fload) {
        double xi[NXI];
#pragma unroll
        for (int q = 0; q < NXI; ++q) xi[q] = __ldg(&a.xi[size_t(q) * a.xi_ld + e]);
        Kin<D, double, double> k0;
        k0.gu = grad_u_val<D, NB>(E.xn, E.g);
        k0.gup = grad_u_val<D, NB>(E.xpn, E.g);
        if constexpr (needs_rotation<Model>::value) cache_rotation(k0);
        if (nmask) {
          double p0 = 0.0;
          if constexpr (C::M == MECH_MIXED) {
#pragma unroll
            for (int n = 0; n < NN; ++n) p0 += E.xn[n][D] * (1.0 / NN);
          }
          const Mat<double, D> P = first_pk<D, C::M, Model>(k0, p0, xi, E.par, a.model.thickness);
          v = calibration_load<D>(a.qoi, P, E.g, wdv, nmask, a.mesh.coords, E.nodes);
        } else {
          double fN[D], fwdv;
          facet_normal<D>(a.qoi, a.mesh.coords, E.nodes, fv, fN, fwdv);
          const Mat<double, D> P = first_pk<D, C::M, Model>(k0, facet_pressure<D, NB>(E.xn, fv), xi, E.par, 1.0);
          v = face_normal_load<D>(P, fN, fwdv);
        }
      }
    }
  }
  __shared__ double sh[4];
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) atomicAdd(&a.scalars[mode], sh[0] + sh[1] + sh[2] + sh[3]);
}

}  // namespace c8
