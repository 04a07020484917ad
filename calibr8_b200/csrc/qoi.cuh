// Per-point objective integrands (QoI<T> of the reference) as device templates.
//   AvgDisp      src/avg_disp.cpp:15-33
//   Calibration  src/calibration.cpp:162-223 (2-D element mismatch), :225-303 (3-D surface
//                mismatch), :305-346 (coordinate-plane load), :414-478 (evaluate)
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
// R_u[n,i] = sum_j P_ij dN_n/dX_j w dv  (Calibration::compute_load)
template <int DIM, class T>
C8_DI T calibration_load(const QoiArgs& q, const Mat<T, DIM>& P, const Geom<DIM>& g, double wdv,
                         int node_mask) {
  T load = conv<T>(0.0);
#pragma unroll
  for (int n = 0; n <= DIM; ++n) {
    if (!(node_mask & (1 << n))) continue;
#pragma unroll
    for (int i = 0; i < DIM; ++i) {
      if (i != q.reaction_force_comp) continue;
#pragma unroll
      for (int j = 0; j < DIM; ++j) load += P(i, j) * (g.gN[n][j] * wdv);
    }
  }
  return load;
}

}  // namespace c8
