// K7 / K8: the virtual-fields-method element loops of the reference (single-residual
// mechanics only, as the reference asserts: src/evaluations.cpp:1800,1911,2059),
//   K7 eval_measured_residual / eval_measured_residual_and_grad  src/evaluations.cpp:1750-1973
//   K8 eval_vfm_adjoint_gradient                                   src/evaluations.cpp:1975-2143
// Per-point sensitivity state: local_sens [NXI*NPAR][ld] (dxi/dp for EVERY model parameter,
// lane p = parameter p), history h [NXI][ld].  dR [NPAR][n_dofs] node-interleaved rows.
#pragma once
#include "adjoint.cuh"

namespace c8 {


template <class C>
__global__ void __launch_bounds__(128) k_vfm_forward(const VfmArgs a) {
  using Model = typename C::Model;
  constexpr int D = C::D, NN = C::NN, NB = C::NB, NX = C::NX, NXI = C::NXI, NPAR = C::NPAR,
                LXI = C::LXI, G = C::G;
  constexpr int LP = (NPAR + G - 1) / G;
  const int gid = blockIdx.x * blockDim.x + threadIdx.x;
  const int t = gid % G;
  const bool in_range = (gid / G) < a.mesh.n_elems;
  const int e = in_range ? gid / G : a.mesh.n_elems - 1;
  const unsigned mask = group_mask<C>();

  Elem<C> E;
  load_elem<C>(a.mesh, a.model, a.x, a.x_prev, a.xi_prev, a.xi_ld, e, E);
  double xi[NXI];
#pragma unroll
  for (int q = 0; q < NXI; ++q) xi[q] = a.xi[size_t(q) * a.xi_ld + e];
  Kin<D, double, double> k0;
  k0.gu = grad_u_val<D, NB>(E.xn, E.g);
  k0.gup = grad_u_val<D, NB>(E.xpn, E.g);
  if constexpr (needs_rotation<Model>::value) cache_rotation(k0);
  Dual<LXI> Cd[NXI];
  const int path = local_newton<C>(k0, E, a.model, xi, Cd, mask, t, in_range);
  if (!in_range) return;
  if (path < 0) {
    if (t == 0) atomicAdd(a.n_failed, 1);
    // the reference ignores the status here (:1822) and carries on with the last iterate
  }
#pragma unroll
  for (int q = 0; q < NXI; ++q)
    if (q % G == t) a.xi[size_t(q) * a.xi_ld + e] = xi[q];

  const double wdv = quad1_weight<D>() * E.g.dv;
  // element residual (values)
  {
    const Mat<double, D> P = first_pk<D, C::M, Model>(k0, 0.0, xi, E.par, a.model.thickness);
#pragma unroll
    for (int n = 0; n < NN; ++n)
#pragma unroll
      for (int i = 0; i < D; ++i) {
        double r = 0.0;
#pragma unroll
        for (int j = 0; j < D; ++j) r += P(i, j) * (E.g.gN[n][j] * wdv);
        if ((n * NB + i) % G == t && E.nodes[n] < a.mesh.n_row_nodes)
          atomicAdd(&a.b[size_t(E.nodes[n]) * NB + i], r);
      }
  }
  if (!a.dR) return;

  // ---- forward sensitivities, :1921-1957 ---------------------------------------------
  // dR/dxi (rows a, xi lanes) gathered to every thread
  double dR_dxi[NX][NXI];
  {
    Dual<LXI> xs[NXI];
#pragma unroll
    for (int q = 0; q < NXI; ++q) xs[q] = seeded<LXI>(xi[q], q, t * LXI);
    const Mat<Dual<LXI>, D> P = first_pk<D, C::M, Model>(k0, 0.0, xs, E.par, a.model.thickness);
#pragma unroll
    for (int n = 0; n < NN; ++n)
#pragma unroll
      for (int i = 0; i < D; ++i) {
        Dual<LXI> r = P(i, 0) * (E.g.gN[n][0] * wdv);
#pragma unroll
        for (int j = 1; j < D; ++j) r += P(i, j) * (E.g.gN[n][j] * wdv);
#pragma unroll
        for (int q = 0; q < NXI; ++q)
          dR_dxi[n * NB + i][q] = group_bcast<G>(mask, r.d[q % LXI], q / LXI);
      }
  }
  // dC/dxi_prev gathered to every thread
  double dC_dxip[NXI][NXI];
  {
    Dual<LXI> xps[NXI], Cq[NXI];
#pragma unroll
    for (int q = 0; q < NXI; ++q) xps[q] = seeded<LXI>(E.xip[q], q, t * LXI);
    Model::residual(k0, xi, xps, E.par, a.model.abs_tol, Cq);
#pragma unroll
    for (int i = 0; i < NXI; ++i)
#pragma unroll
      for (int q = 0; q < NXI; ++q)
        dC_dxip[i][q] = group_bcast<G>(mask, Cq[i].d[q % LXI], q / LXI);
  }
  // parameters seeded: dC/dp, dR/dp for this thread's parameter lanes
  using TP = Dual<LP>;
  TP par[NPAR];
#pragma unroll
  for (int k = 0; k < NPAR; ++k) par[k] = seeded<LP>(E.par[k], k, t * LP);
  double rhs[NXI][LP];
  {
    TP Cp[NXI];
    Model::residual(k0, xi, E.xip, par, a.model.abs_tol, Cp);
#pragma unroll
    for (int i = 0; i < NXI; ++i)
#pragma unroll
      for (int s = 0; s < LP; ++s) {
        const int p = t * LP + s;
        double v = -Cp[i].d[s];
        if (p < NPAR) {
#pragma unroll
          for (int q = 0; q < NXI; ++q)
            v -= dC_dxip[i][q] * a.local_sens[size_t(q * NPAR + p) * a.xi_ld + e];
        }
        rhs[i][s] = v;
      }
  }
  // dxi/dp = (dC/dxi)^-1 rhs
  {
    double Jc[NXI][LXI], dummy[NXI];
#pragma unroll
    for (int q = 0; q < NXI; ++q) {
      dummy[q] = 0.0;
#pragma unroll
      for (int s = 0; s < LXI; ++s) Jc[q][s] = Cd[q].d[s];
    }
    group_gauss_jordan<NXI, LXI, LP, G>(Jc, rhs, dummy, mask);
  }
#pragma unroll
  for (int s = 0; s < LP; ++s) {
    const int p = t * LP + s;
    if (p >= NPAR) continue;
#pragma unroll
    for (int q = 0; q < NXI; ++q) a.local_sens[size_t(q * NPAR + p) * a.xi_ld + e] = rhs[q][s];
  }
  // dR/dp_total = dR/dxi dxi/dp + dR/dp, scattered into the multivector
  {
    const Mat<TP, D> P = first_pk<D, C::M, Model>(k0, 0.0, xi, par, a.model.thickness);
    const size_t n_dofs = size_t(a.mesh.n_nodes) * NB;
#pragma unroll
    for (int n = 0; n < NN; ++n)
#pragma unroll
      for (int i = 0; i < D; ++i) {
        TP r = P(i, 0) * (E.g.gN[n][0] * wdv);
#pragma unroll
        for (int j = 1; j < D; ++j) r += P(i, j) * (E.g.gN[n][j] * wdv);
#pragma unroll
        for (int s = 0; s < LP; ++s) {
          const int p = t * LP + s;
          if (p >= NPAR) continue;
          double v = r.d[s];
#pragma unroll
          for (int q = 0; q < NXI; ++q) v = fma(dR_dxi[n * NB + i][q], rhs[q][s], v);
          if (E.nodes[n] < a.mesh.n_row_nodes)
            atomicAdd(&a.dR[size_t(p) * n_dofs + size_t(E.nodes[n]) * NB + i], v);
        }
      }
  }
}

template <class C>
__global__ void __launch_bounds__(128) k_vfm_adjoint(const VfmArgs a) {
  using Model = typename C::Model;
  constexpr int D = C::D, NN = C::NN, NB = C::NB, NXI = C::NXI, NPAR = C::NPAR, LXI = C::LXI,
                G = C::G;
  constexpr int LP = (NPAR + G - 1) / G;
  using TP = Dual<LP>;
  __shared__ double sacc[128 / G][G * LP];
  const int gid = blockIdx.x * blockDim.x + threadIdx.x;
  const int t = gid % G, gl = threadIdx.x / G;
  const bool in_range = (gid / G) < a.mesh.n_elems;
  const int e = in_range ? gid / G : a.mesh.n_elems - 1;
  const unsigned mask = group_mask<C>();
  double acc[LP];
#pragma unroll
  for (int s = 0; s < LP; ++s) acc[s] = 0.0;
  {
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
    double w[NN][NB];
#pragma unroll
    for (int n = 0; n < NN; ++n)
#pragma unroll
      for (int q = 0; q < NB; ++q) w[n][q] = __ldg(&a.w[size_t(E.nodes[n]) * NB + q]);
    // xi seeded: dC/dxi, dR/dxi^T w
    Dual<LXI> xs[NXI], Cd[NXI];
#pragma unroll
    for (int q = 0; q < NXI; ++q) xs[q] = seeded<LXI>(xi[q], q, t * LXI);
    Model::residual(k0, xs, E.xip, E.par, a.model.abs_tol, Cd);
    double rl[LXI];
#pragma unroll
    for (int s = 0; s < LXI; ++s) rl[s] = 0.0;
    {
      const Mat<Dual<LXI>, D> P = first_pk<D, C::M, Model>(k0, 0.0, xs, E.par, a.model.thickness);
#pragma unroll
      for (int n = 0; n < NN; ++n)
#pragma unroll
        for (int i = 0; i < D; ++i) {
          Dual<LXI> r = P(i, 0) * (E.g.gN[n][0] * wdv);
#pragma unroll
          for (int j = 1; j < D; ++j) r += P(i, j) * (E.g.gN[n][j] * wdv);
#pragma unroll
          for (int s = 0; s < LXI; ++s) rl[s] = fma(r.d[s], w[n][i], rl[s]);
        }
    }
    // rhs = s * (-dR/dxi^T w) - h ; phi = (dC/dxi)^-T rhs, :2099-2103
    double rhs[NXI], JT[NXI][LXI], dummy[NXI][1];
#pragma unroll
    for (int q = 0; q < NXI; ++q) {
      const double v = group_bcast_full<G>(rl[q % LXI], q / LXI);
      rhs[q] = a.s * -v - a.hist[size_t(q) * a.xi_ld + e];
    }
#pragma unroll
    for (int i = 0; i < NXI; ++i)
#pragma unroll
      for (int s = 0; s < LXI; ++s) JT[i][s] = 0.0;
#pragma unroll
    for (int r = 0; r < NXI; ++r)
#pragma unroll
      for (int i = 0; i < NXI; ++i) {
        const double v = group_bcast_full<G>(Cd[r].d[i % LXI], i / LXI);
#pragma unroll
        for (int s = 0; s < LXI; ++s) JT[i][s] = pick(t * LXI + s == r, v, JT[i][s]);
      }
    group_gauss_jordan<NXI, LXI, 0, G, true>(JT, dummy, rhs, mask);  // rhs := phi
    // h <- dC/dxi_prev^T phi
    {
      Dual<LXI> xps[NXI], Cq[NXI];
#pragma unroll
      for (int q = 0; q < NXI; ++q) xps[q] = seeded<LXI>(E.xip[q], q, t * LXI);
      Model::residual(k0, xi, xps, E.par, a.model.abs_tol, Cq);
#pragma unroll
      for (int s = 0; s < LXI; ++s) {
        const int c = t * LXI + s;
        if (c >= NXI || !in_range) continue;
        double v = 0.0;
#pragma unroll
        for (int q = 0; q < NXI; ++q) v = fma(Cq[q].d[s], rhs[q], v);
        a.hist[size_t(c) * a.xi_ld + e] = v;
      }
    }
    // grad += s * w^T dR/dp + phi^T dC/dp
    TP par[NPAR];
#pragma unroll
    for (int k = 0; k < NPAR; ++k) par[k] = seeded<LP>(E.par[k], k, t * LP);
    {
      TP Cp[NXI];
      Model::residual(k0, xi, E.xip, par, a.model.abs_tol, Cp);
#pragma unroll
      for (int q = 0; q < NXI; ++q)
#pragma unroll
        for (int s = 0; s < LP; ++s) acc[s] = fma(Cp[q].d[s], rhs[q], acc[s]);
    }
    {
      const Mat<TP, D> P = first_pk<D, C::M, Model>(k0, 0.0, xi, par, a.model.thickness);
#pragma unroll
      for (int n = 0; n < NN; ++n)
#pragma unroll
        for (int i = 0; i < D; ++i) {
          TP r = P(i, 0) * (E.g.gN[n][0] * wdv);
#pragma unroll
          for (int j = 1; j < D; ++j) r += P(i, j) * (E.g.gN[n][j] * wdv);
#pragma unroll
          for (int s = 0; s < LP; ++s) acc[s] = fma(a.s * w[n][i], r.d[s], acc[s]);
        }
    }
  }
#pragma unroll
  for (int s = 0; s < LP; ++s) sacc[gl][t * LP + s] = in_range ? acc[s] : 0.0;
  __syncthreads();
  if (threadIdx.x < G * LP) {
    double v = 0.0;
    for (int g2 = 0; g2 < 128 / G; ++g2) v += sacc[g2][threadIdx.x];
    if (threadIdx.x < NPAR) atomicAdd(&a.grad[threadIdx.x], v);
  }
}

}  // namespace c8
