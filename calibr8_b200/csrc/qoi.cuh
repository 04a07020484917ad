// Per-point objective integrands (QoI<T> of the reference) as device templates.
//   AvgDisp      src/avg_disp.cpp:15-33
//   Calibration  src/calibration.cpp:162-223 (2-D element mismatch), :225-303 (3-D surface
//                mismatch), :305-346 (coordinate-plane load), :414-478 (evaluate)
//   ReactionMismatch  src/reaction_mismatch.cpp:58-129 (coordinate-plane load or torque)
//   LoadMismatch      src/load_mismatch.cpp:79-171 (normal load N.P.N over a side-set facet)
//   SurfaceMismatch   src/surface_mismatch.cpp:32-117 (the calibration surface integrand with unit
//                     weights and no normalisation: the host passes weights 2, inv_area 1, dt/T 1)
#pragma once
#include "adj_args.h"
#include "mechanics.cuh"

namespace c8 {


// which element nodes lie on the load plane (QoI::setup_coord_based_node_mapping, src/qoi.cpp:159-198)
template <int DIM>
C8_DI int load_node_mask(const QoiArgs& q, const double* __restrict__ coords, const int* nodes) {
  int m = 0;
#pragma unroll
  for (int n = 0; n <= DIM; ++n) {
    const double x = __ldg(&coords[size_t(nodes[n]) * DIM + q.coord_idx]);
    if (fabs(x - q.coord_value) < q.coord_tol) m |= (1 << n);
  }
  return m;
}

// displacement at a reference point with basis N: value type T per nodal dof
template <int DIM, int NB, class T>
C8_DI void interp_u(const T (&un)[DIM + 1][DIM], const double* N, T* u) {
#pragma unroll
  for (int i = 0; i < DIM; ++i) {
    T s = un[0][i] * N[0];
#pragma unroll
    for (int n = 1; n <= DIM; ++n) s += un[n][i] * N[n];
    u[i] = s;
  }
}

// Calibration displacement-mismatch integrand of one element; un = nodal displacements (T).
// 2-D: order-2 rule over the element; 3-D: order-2 triangle rule over the facet on the side set.
template <int DIM, class T>
C8_DI T calibration_disp_mismatch(const QoiArgs& q, const T (&un)[DIM + 1][DIM],
                                  const double (&um)[DIM + 1][DIM], const double (&X)[DIM + 1][DIM],
                                  double elem_dv, const signed char* fv) {
  T mismatch = conv<T>(0.0);
  if constexpr (DIM == 2) {
#pragma unroll
    for (int p = 0; p < 3; ++p) {
      double N[3];
      Quad2<2>::basis(p, N);
      T u[2];
      interp_u<2, 2, T>(un, N, u);
      T qv = conv<T>(0.0);
#pragma unroll
      for (int d = 0; d < 2; ++d) {
        const double m = um[0][d] * N[0] + um[1][d] * N[1] + um[2][d] * N[2];
        const T diff = u[d] - m;
        qv += q.weights[d] * diff * diff;
      }
      mismatch += 0.5 * qv * (Quad2<2>::weight() * elem_dv) * q.inv_area * q.dt_over_T;
    }
  } else {
    // facet geometry
    double a[3], b[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      double x0 = 0, x1 = 0, x2 = 0;
#pragma unroll
      for (int n = 0; n < 4; ++n) {
        x0 = pick(fv[0] == n, X[n][k], x0);
        x1 = pick(fv[1] == n, X[n][k], x1);
        x2 = pick(fv[2] == n, X[n][k], x2);
      }
      a[k] = x1 - x0; b[k] = x2 - x0;
    }
    const double cx = a[1] * b[2] - a[2] * b[1], cy = a[2] * b[0] - a[0] * b[2],
                 cz = a[0] * b[1] - a[1] * b[0];
    const double fdv = sqrt(cx * cx + cy * cy + cz * cz);
#pragma unroll
    for (int p = 0; p < 3; ++p) {
      double Nf[3];
      Quad2<2>::basis(p, Nf);
      double N[4];
#pragma unroll
      for (int n = 0; n < 4; ++n)
        N[n] = (fv[0] == n ? Nf[0] : 0.0) + (fv[1] == n ? Nf[1] : 0.0) + (fv[2] == n ? Nf[2] : 0.0);
      T u[3];
      interp_u<3, 3, T>(un, N, u);
      T qv = conv<T>(0.0);
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        const double m = um[0][d] * N[0] + um[1][d] * N[1] + um[2][d] * N[2] + um[3][d] * N[3];
        const T diff = u[d] - m;
        qv += q.weights[d] * diff * diff;
      }
      mismatch += 0.5 * qv * (Quad2<2>::weight() * fdv) * q.inv_area * q.dt_over_T;
    }
  }
  return mismatch;
}

// sum over the element's nodes on the load plane of R_u[node, comp] with
// R_u[n,i] = sum_j P_ij dN_n/dX_j w dv  (Calibration::compute_load, ReactionMismatch::compute_load);
// compute_torque: the moment of the nodal force about axis `comp` instead (r = node position)
template <int DIM, class T>
C8_DI T calibration_load(const QoiArgs& q, const Mat<T, DIM>& P, const Geom<DIM>& g, double wdv,
                         int node_mask, const double* __restrict__ coords = nullptr,
                         const int* nodes = nullptr) {
  T load = conv<T>(0.0);
  const int c = q.reaction_force_comp;
#pragma unroll
  for (int n = 0; n <= DIM; ++n) {
    if (!(node_mask & (1 << n))) continue;
    if (q.compute_torque) {
      T F[DIM];
#pragma unroll
      for (int i = 0; i < DIM; ++i) {
        F[i] = P(i, 0) * (g.gN[n][0] * wdv);
#pragma unroll
        for (int j = 1; j < DIM; ++j) F[i] += P(i, j) * (g.gN[n][j] * wdv);
      }
      double r[3] = {0.0, 0.0, 0.0};
#pragma unroll
      for (int k = 0; k < DIM; ++k) r[k] = __ldg(&coords[size_t(nodes[n]) * DIM + k]);
      if constexpr (DIM == 3) {
        if (c == 2) load += r[0] * F[1] - r[1] * F[0];
        else if (c == 0) load += r[1] * F[2] - r[2] * F[1];
        else load += r[2] * F[0] - r[0] * F[2];
      } else {
        load += r[0] * F[1] - r[1] * F[0];   // the only moment of a plane problem (axis 2)
      }
    } else {
#pragma unroll
      for (int i = 0; i < DIM; ++i) {
        if (i != c) continue;
#pragma unroll
        for (int j = 0; j < DIM; ++j) load += P(i, j) * (g.gN[n][j] * wdv);
      }
    }
  }
  return load;
}

// the facet of the element on the side set: vertex coordinates -> outward unit normal N and w dv of
// the one-point facet rule (3-D: area, 2-D: length; the 2-D normal is the deck's "2D surface normal")
template <int DIM>
C8_DI void facet_normal(const QoiArgs& q, const double* __restrict__ coords, const int* nodes,
                        const signed char* fv, double (&N)[DIM], double& wdv) {
  double X[DIM + 1][DIM];
#pragma unroll
  for (int n = 0; n <= DIM; ++n)
#pragma unroll
    for (int k = 0; k < DIM; ++k) X[n][k] = __ldg(&coords[size_t(nodes[n]) * DIM + k]);
  double xa[DIM], xb[DIM], xc[DIM], xo[DIM];   // facet vertices and the opposite vertex
#pragma unroll
  for (int k = 0; k < DIM; ++k) {
    xa[k] = xb[k] = xc[k] = xo[k] = 0.0;
#pragma unroll
    for (int n = 0; n <= DIM; ++n) {
      xa[k] = pick(fv[0] == n, X[n][k], xa[k]);
      xb[k] = pick(fv[1] == n, X[n][k], xb[k]);
      if (DIM == 3) xc[k] = pick(fv[2] == n, X[n][k], xc[k]);
      const bool in_f = fv[0] == n || fv[1] == n || (DIM == 3 && fv[2] == n);
      xo[k] = pick(!in_f, X[n][k], xo[k]);
    }
  }
  if constexpr (DIM == 3) {
    const double u[3] = {xb[0] - xa[0], xb[1] - xa[1], xb[2] - xa[2]};
    const double v[3] = {xc[0] - xa[0], xc[1] - xa[1], xc[2] - xa[2]};
    double n[3] = {u[1] * v[2] - u[2] * v[1], u[2] * v[0] - u[0] * v[2], u[0] * v[1] - u[1] * v[0]};
    const double len = sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
    const double d = n[0] * (xo[0] - xa[0]) + n[1] * (xo[1] - xa[1]) + n[2] * (xo[2] - xa[2]);
    const double sgn = d > 0.0 ? -1.0 : 1.0;
#pragma unroll
    for (int k = 0; k < 3; ++k) N[k] = sgn * n[k] / len;
    wdv = 0.5 * len;
  } else {
    N[0] = q.normal_2d[0]; N[1] = q.normal_2d[1];
    wdv = sqrt((xb[0] - xa[0]) * (xb[0] - xa[0]) + (xb[1] - xa[1]) * (xb[1] - xa[1]));
  }
}

// LoadMismatch::compute_load: N . P . N w dv at the facet centroid, P = J sigma F^-T (times the z
// stretch in plane stress) with the FULL Cauchy stress, i.e. the pressure interpolated at the facet
// centroid (p_face = mean of the facet nodes' pressure dofs).  P comes from first_pk with unit thickness.
template <int DIM, class T>
C8_DI T face_normal_load(const Mat<T, DIM>& P, const double (&N)[DIM], double wdv) {
  T load = conv<T>(0.0);
#pragma unroll
  for (int i = 0; i < DIM; ++i)
#pragma unroll
    for (int j = 0; j < DIM; ++j) load += (N[i] * N[j]) * P(i, j);
  return load * wdv;
}

// mean pressure of the facet nodes (value) and its weight per element node (1/DIM on the facet)
template <int DIM, int NB>
C8_DI double facet_pressure(const double (&xn)[DIM + 1][NB], const signed char* fv) {
  double p = 0.0;
  if constexpr (NB > DIM) {
#pragma unroll
    for (int n = 0; n <= DIM; ++n) {
      const bool in_f = fv[0] == n || fv[1] == n || (DIM == 3 && fv[2] == n);
      p += in_f ? xn[n][DIM] * (1.0 / DIM) : 0.0;
    }
  }
  return p;
}

}  // namespace c8
