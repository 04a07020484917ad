// ORACLE (test infrastructure only).
//
// Flat-array stand-in for calibr8's Disc (src/disc.hpp:72-483) plus the
// pieces of SCOREC apf the hot path leans on (un-vendored; pin
// package/scorec/package.cmake:2-3): linear simplex shape functions, the
// order-1 / order-2 Gauss rules on triangles and tets, the differential
// volume and the node numbering -> CSR graph construction of
// src/disc.cpp:263-269,334-484.
#pragma once
#include <vector>
#include <cmath>
#include <algorithm>
#include <cstdint>

namespace orc {

enum { SCALAR = 0, VECTOR = 1, SYM_TENSOR = 2, TENSOR = 3 };

// src/fields.cpp:11-18
inline int get_num_eqs(int type, int ndims) {
  if (type == SCALAR) return 1;
  if (type == VECTOR) return ndims;
  if (type == SYM_TENSOR) return (ndims + 1) * ndims / 2;
  if (type == TENSOR) return ndims * ndims;
  return -1;
}

struct Disc {
  int dim = 0;       // mesh dimension
  int nn = 0;        // nodes per element (dim + 1)
  int n_elems = 0;
  int n_nodes = 0;
  int n_es = 1;
  std::vector<int> conn;      // [n_elems * nn]
  std::vector<double> coords; // [n_nodes * 3]
  std::vector<int> elem_es;   // [n_elems]
  // elements of each set in mesh order (src/disc.cpp:486-499)
  std::vector<std::vector<int>> es_elems;

  void finalize() {
    es_elems.assign(n_es, {});
    for (int e = 0; e < n_elems; ++e) es_elems[elem_es[e]].push_back(e);
  }
  double const* X(int node) const { return &coords[size_t(node) * 3]; }
};

// ---- quadrature on the reference simplex (apf getIntPoint/getIntWeight) ----
struct QPoint { double xi[3]; double w; };

inline std::vector<QPoint> quadrature(int dim, int order) {
  std::vector<QPoint> q;
  if (dim == 3) {
    if (order <= 1) {
      q.push_back({{0.25, 0.25, 0.25}, 1. / 6.});
    } else {
      double const a = 0.138196601125011, b = 0.585410196624969;
      q.push_back({{a, a, a}, 1. / 24.});
      q.push_back({{b, a, a}, 1. / 24.});
      q.push_back({{a, b, a}, 1. / 24.});
      q.push_back({{a, a, b}, 1. / 24.});
    }
  } else {
    if (order <= 1) {
      q.push_back({{1. / 3., 1. / 3., 0.}, 0.5});
    } else {
      q.push_back({{2. / 3., 1. / 6., 0.}, 1. / 6.});
      q.push_back({{1. / 6., 2. / 3., 0.}, 1. / 6.});
      q.push_back({{1. / 6., 1. / 6., 0.}, 1. / 6.});
    }
  }
  return q;
}

// Geometry of one straight-sided simplex element: basis values at a point,
// constant global basis gradients, differential volume, mean edge size.
struct ElemGeom {
  int dim, nn;
  double x[4][3];
  double grad[4][3]; // dN_n / dX_j
  double dv;         // |det J| (tet) or 2*area (tri in 2D)
  double h;          // sqrt(mean squared edge length), src/mechanics.cpp:103-113

  void set(Disc const& d, int elem) {
    dim = d.dim; nn = d.nn;
    for (int n = 0; n < nn; ++n) {
      double const* p = d.X(d.conn[size_t(elem) * nn + n]);
      for (int k = 0; k < 3; ++k) x[n][k] = p[k];
    }
    // Jacobian rows: J[a][k] = d x_k / d xi_a = x[a+1][k] - x[0][k]
    double J[3][3] = {{0}};
    for (int a = 0; a < dim; ++a)
      for (int k = 0; k < dim; ++k) J[a][k] = x[a + 1][k] - x[0][k];
    double Jinv[3][3] = {{0}};
    if (dim == 3) {
      double const det =
          J[0][0] * (J[1][1] * J[2][2] - J[1][2] * J[2][1]) -
          J[0][1] * (J[1][0] * J[2][2] - J[1][2] * J[2][0]) +
          J[0][2] * (J[1][0] * J[2][1] - J[1][1] * J[2][0]);
      dv = det;
      Jinv[0][0] = (J[1][1] * J[2][2] - J[1][2] * J[2][1]) / det;
      Jinv[0][1] = (J[0][2] * J[2][1] - J[0][1] * J[2][2]) / det;
      Jinv[0][2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) / det;
      Jinv[1][0] = (J[1][2] * J[2][0] - J[1][0] * J[2][2]) / det;
      Jinv[1][1] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) / det;
      Jinv[1][2] = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) / det;
      Jinv[2][0] = (J[1][0] * J[2][1] - J[1][1] * J[2][0]) / det;
      Jinv[2][1] = (J[0][1] * J[2][0] - J[0][0] * J[2][1]) / det;
      Jinv[2][2] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) / det;
    } else {
      double const det = J[0][0] * J[1][1] - J[0][1] * J[1][0];
      dv = std::abs(det);
      Jinv[0][0] = J[1][1] / det;
      Jinv[0][1] = -J[0][1] / det;
      Jinv[1][0] = -J[1][0] / det;
      Jinv[1][1] = J[0][0] / det;
    }
    // reference gradients: N0 = 1 - sum xi, N_a+1 = xi_a ;  grad_X N = Jinv * grad_xi N
    for (int k = 0; k < 3; ++k) grad[0][k] = 0.;
    for (int a = 0; a < dim; ++a) {
      for (int k = 0; k < dim; ++k) {
        grad[a + 1][k] = Jinv[k][a];
        grad[0][k] -= Jinv[k][a];
      }
      for (int k = dim; k < 3; ++k) grad[a + 1][k] = 0.;
    }
    // element size
    double s = 0.; int ne = 0;
    for (int a = 0; a < nn; ++a)
      for (int b = a + 1; b < nn; ++b) {
        double l2 = 0.;
        for (int k = 0; k < 3; ++k) l2 += (x[a][k] - x[b][k]) * (x[a][k] - x[b][k]);
        s += l2; ++ne;
      }
    h = std::sqrt(s / ne);
  }
  void basis(double const* xi, double* N) const {
    double s = 0.;
    for (int a = 0; a < dim; ++a) { N[a + 1] = xi[a]; s += xi[a]; }
    N[0] = 1. - s;
  }
};

// ---- node graph and per-block CSR pattern (src/disc.cpp:356-387) ----
struct CsrGraph {
  int n_rows = 0;
  std::vector<int> rowptr;
  std::vector<int> colind;
};

inline CsrGraph node_graph(Disc const& d) {
  std::vector<std::vector<int>> adj(d.n_nodes);
  for (int e = 0; e < d.n_elems; ++e)
    for (int a = 0; a < d.nn; ++a)
      for (int b = 0; b < d.nn; ++b)
        adj[d.conn[size_t(e) * d.nn + a]].push_back(d.conn[size_t(e) * d.nn + b]);
  CsrGraph g;
  g.n_rows = d.n_nodes;
  g.rowptr.assign(d.n_nodes + 1, 0);
  for (int n = 0; n < d.n_nodes; ++n) {
    auto& v = adj[n];
    std::sort(v.begin(), v.end());
    v.erase(std::unique(v.begin(), v.end()), v.end());
    g.rowptr[n + 1] = g.rowptr[n] + int(v.size());
  }
  g.colind.reserve(g.rowptr.back());
  for (int n = 0; n < d.n_nodes; ++n)
    for (int c : adj[n]) g.colind.push_back(c);
  return g;
}

// block (i,j) graph: row = node*neq_i + eq_i, col = node*neq_j + eq_j, sorted
inline CsrGraph block_graph(CsrGraph const& ng, int neq_i, int neq_j) {
  CsrGraph g;
  g.n_rows = ng.n_rows * neq_i;
  g.rowptr.assign(g.n_rows + 1, 0);
  for (int n = 0; n < ng.n_rows; ++n) {
    int const cnt = (ng.rowptr[n + 1] - ng.rowptr[n]) * neq_j;
    for (int eq = 0; eq < neq_i; ++eq) g.rowptr[n * neq_i + eq + 1] = cnt;
  }
  for (int r = 0; r < g.n_rows; ++r) g.rowptr[r + 1] += g.rowptr[r];
  g.colind.resize(g.rowptr.back());
  for (int n = 0; n < ng.n_rows; ++n)
    for (int eq = 0; eq < neq_i; ++eq) {
      int o = g.rowptr[n * neq_i + eq];
      for (int k = ng.rowptr[n]; k < ng.rowptr[n + 1]; ++k)
        for (int ej = 0; ej < neq_j; ++ej) g.colind[o++] = ng.colind[k] * neq_j + ej;
    }
  return g;
}

}  // namespace orc
