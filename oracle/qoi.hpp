// ORACLE (test infrastructure only).
//
// CPU restatement of the QoI operator API (src/qoi.{hpp,cpp}) and the in-scope QoIs:
//   AvgDisp           src/avg_disp.cpp:15-33
//   Calibration       src/calibration.cpp:56-478  (2-D disp mismatch over elements,
//                     3-D surface mismatch over a side set, coordinate-plane load)
//   ReactionMismatch  src/reaction_mismatch.cpp:42-212 (coordinate-plane load or torque)
//   LoadMismatch      src/load_mismatch.cpp:40-259 (normal load N.P.N over a side set)
//   SurfaceMismatch   src/surface_mismatch.cpp:19-117 (|u - u_meas|^2 over a side set)
#pragma once
#include "residuals.hpp"

namespace orc {

template <class T>
class QoI {
 public:
  virtual ~QoI() {}
  virtual void before_elems(Disc const& disc, int step) {
    m_disc = &disc; m_num_dims = disc.dim; m_step = step;
  }
  void set_elem(int elem) { m_elem = elem; }
  virtual void preprocess(int, int, GlobalResidual<T>&, LocalResidual<T>&,
                          double const*, double, double) {}
  virtual void preprocess_finalize(int) {}
  virtual void evaluate(int es, int elem, GlobalResidual<T>& global,
                        LocalResidual<T>& local, double const* iota, double w, double dv) = 0;
  virtual void postprocess(double&) {}
  void scatter(double& J) { J += val(value_pt); }  // src/qoi.cpp scatter
  // src/qoi.cpp:225-233
  EVector eigen_dvector(int nderivs) const {
    EVector dJ(nderivs, 0.);
    for (int i = 0; i < nderivs; ++i) dJ[i] = dx(value_pt, i);
    return dJ;
  }
  void initialize_value_pt() { value_pt = T(0.); }
  T value_pt;

 protected:
  Disc const* m_disc = nullptr;
  int m_num_dims = 0, m_step = 0, m_elem = -1;
};

template <class T>
class AvgDisp : public QoI<T> {
 public:
  void evaluate(int, int, GlobalResidual<T>& global, LocalResidual<T>&, double const*,
                double w, double dv) override {
    this->initialize_value_pt();
    Vec<T> const u = global.vector_x(0);
    for (int i = 0; i < this->m_num_dims; ++i) this->value_pt += u[i] * w * dv;
    this->value_pt /= double(this->m_num_dims);
  }
};

// Shared (type-independent) calibration settings and per-step data.
struct CalibrationData {
  double balance_factor = 1.;
  int coord_idx = -1;
  double coord_value = 0.;
  double coord_tol = 1e-12;
  int reaction_force_comp = -1;
  double weights[3] = {1., 1., 1.};
  // 3-D: local vertex ids of the facet on the displacement side set, or -1
  std::vector<int> facet;          // [n_elems * 3]
  bool initd = false;
  double area = 1.;
  std::vector<int> mapping_disp;   // [n_elems]  (-1: not in objective)
  std::vector<std::vector<int>> mapping_load;  // [n_elems] node ids or {-1}
  // per step
  double dt = 1., total_time = 1.;
  double load_meas = 0.;
  double const* measured = nullptr;  // [n_nodes * 3]
  // state shared between the double and Fad instances (m_total_load etc.)
  double total_load = 0.;
  double last_total_load = 0.;  // what `load out file` would hold for the step
  double load_mismatch = 0.;
  int n_ranks = 1;
  // last step's J split (objective out file columns)
  double J_disp = 0., J_forc = 0.;
  // reaction / load mismatch
  bool compute_torque = false;         // ReactionMismatch "compute torque"
  double normal_2d[2] = {0., 0.};      // LoadMismatch "2D surface normal"
};

template <class T>
class Calibration : public QoI<T> {
 public:
  explicit Calibration(CalibrationData* d) : D(d) {}

  // src/calibration.cpp:56-160
  void before_elems(Disc const& disc, int step) override {
    QoI<T>::before_elems(disc, step);
    if (!D->initd) {
      int const nd = disc.dim;
      D->area = 0.;
      D->mapping_disp.assign(disc.n_elems, -1);
      ElemGeom g;
      if (nd == 2) {
        for (int e = 0; e < disc.n_elems; ++e) {
          D->mapping_disp[e] = 1;
          g.set(disc, e);
          D->area += 0.5 * g.dv;
        }
      } else {
        for (int e = 0; e < disc.n_elems; ++e) {
          if (D->facet.empty() || D->facet[size_t(e) * 3] < 0) continue;
          D->mapping_disp[e] = 1;
          double dv = face_dv(disc, e);
          D->area += 0.5 * dv;
        }
      }
      // src/qoi.cpp:159-198 coordinate-plane node mapping
      D->mapping_load.assign(disc.n_elems, {});
      for (int e = 0; e < disc.n_elems; ++e) {
        std::vector<int> ids;
        for (int n = 0; n < disc.nn; ++n) {
          double const* x = disc.X(disc.conn[size_t(e) * disc.nn + n]);
          if (std::abs(x[D->coord_idx] - D->coord_value) < D->coord_tol) ids.push_back(n);
        }
        if (ids.empty()) ids.push_back(-1);
        D->mapping_load[e] = ids;
      }
      D->initd = true;
    }
  }

  double face_dv(Disc const& disc, int e) const {
    double const* a = disc.X(disc.conn[size_t(e) * disc.nn + D->facet[size_t(e) * 3 + 0]]);
    double const* b = disc.X(disc.conn[size_t(e) * disc.nn + D->facet[size_t(e) * 3 + 1]]);
    double const* c = disc.X(disc.conn[size_t(e) * disc.nn + D->facet[size_t(e) * 3 + 2]]);
    double u[3], v[3];
    for (int k = 0; k < 3; ++k) { u[k] = b[k] - a[k]; v[k] = c[k] - a[k]; }
    double const cx = u[1] * v[2] - u[2] * v[1];
    double const cy = u[2] * v[0] - u[0] * v[2];
    double const cz = u[0] * v[1] - u[1] * v[0];
    return std::sqrt(cx * cx + cy * cy + cz * cz);
  }

  // measured displacement interpolated with the element basis
  void interp_measured(int elem, double const* N, double* u_meas) const {
    Disc const& d = *this->m_disc;
    for (int k = 0; k < 3; ++k) u_meas[k] = 0.;
    for (int n = 0; n < d.nn; ++n) {
      int const node = d.conn[size_t(elem) * d.nn + n];
      for (int k = 0; k < 3; ++k) u_meas[k] += D->measured[size_t(node) * 3 + k] * N[n];
    }
  }

  // src/calibration.cpp:162-223
  T compute_disp_mismatch(int elem, GlobalResidual<T>& global, double const* iota_input) {
    T mismatch = T(0.);
    int const nd = this->m_num_dims;
    auto const qps = quadrature(nd, 2);
    for (auto const& qp : qps) {
      double const w = qp.w;
      double const dv = global.geom().dv;
      global.interpolate(qp.xi);
      Vec<T> const u_fem = global.vector_x(0);
      double N[4], u_meas[3];
      global.geom().basis(qp.xi, N);
      interp_measured(elem, N, u_meas);
      T qoi = 0.;
      for (int d = 0; d < nd; ++d)
        qoi += D->weights[d] * (u_fem[d] - u_meas[d]) * (u_fem[d] - u_meas[d]);
      mismatch += 0.5 * qoi * w * dv / D->area * D->dt / D->total_time;
    }
    global.interpolate(iota_input);
    return mismatch;
  }

  // src/calibration.cpp:225-303
  T compute_surface_mismatch(int elem, GlobalResidual<T>& global, double const* iota_input) {
    T mismatch = T(0.);
    int const nd = this->m_num_dims;
    int const* fv = &D->facet[size_t(elem) * 3];
    static double const ref[4][3] = {{0, 0, 0}, {1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    double const dv = face_dv(*this->m_disc, elem);
    auto const qps = quadrature(2, 2);
    for (auto const& qp : qps) {
      double const w = qp.w;
      double const Nf[3] = {1. - qp.xi[0] - qp.xi[1], qp.xi[0], qp.xi[1]};
      double iota_elem[3] = {0., 0., 0.};
      for (int k = 0; k < 3; ++k)
        for (int c = 0; c < 3; ++c) iota_elem[c] += Nf[k] * ref[fv[k]][c];
      global.interpolate(iota_elem);
      Vec<T> const u_fem = global.vector_x(0);
      double N[4], u_meas[3];
      global.geom().basis(iota_elem, N);
      interp_measured(elem, N, u_meas);
      T qoi = 0.;
      for (int d = 0; d < nd; ++d)
        qoi += D->weights[d] * (u_fem[d] - u_meas[d]) * (u_fem[d] - u_meas[d]);
      mismatch += 0.5 * qoi * w * dv / D->area * D->dt / D->total_time;
    }
    global.interpolate(iota_input);
    return mismatch;
  }

  // src/calibration.cpp:305-346
  T compute_load(int elem, GlobalResidual<T>& global, LocalResidual<T>& local,
                 double const* iota, double w, double dv) {
    T load_pt = T(0.);
    std::vector<int> const& node_ids = D->mapping_load[elem];
    global.zero_residual();
    global.evaluate(local, iota, w, dv, 0);
    for (size_t i = 0; i < node_ids.size(); ++i)
      load_pt += global.R_nodal(0, node_ids[i], D->reaction_force_comp);
    global.zero_residual();
    return load_pt;
  }

  // src/calibration.cpp:395-412
  void preprocess(int, int elem, GlobalResidual<T>& global, LocalResidual<T>& local,
                  double const* iota, double w, double dv) override {
    if (D->mapping_load[elem][0] < 0) return;
    T load = compute_load(elem, global, local, iota, w, dv);
    D->total_load += val(load);
  }
  // src/calibration.cpp:348-371
  void preprocess_finalize(int) override {
    D->load_mismatch = D->total_load - D->load_meas;
    D->last_total_load = D->total_load;
    D->total_load = 0.;
  }
  // src/calibration.cpp:373-393
  void postprocess(double& J) override {
    D->J_disp = J;
    D->J_forc = 0.5 * D->balance_factor * D->dt / D->total_time *
                std::pow(D->load_mismatch, 2);
    J += D->J_forc;
    J /= D->n_ranks;
  }

  void evaluate(int, int elem, GlobalResidual<T>& global, LocalResidual<T>& local,
                double const* iota_input, double w, double dv) override;

 private:
  CalibrationData* D;
};

// src/calibration.cpp:414-438
template <>
inline void Calibration<double>::evaluate(int, int elem, GlobalResidual<double>& global,
                                          LocalResidual<double>&, double const* iota_input,
                                          double, double) {
  this->initialize_value_pt();
  if (D->mapping_disp[elem] < 0) return;
  if (this->m_num_dims == 2) this->value_pt = compute_disp_mismatch(elem, global, iota_input);
  else this->value_pt = compute_surface_mismatch(elem, global, iota_input);
}
// src/calibration.cpp:440-478
template <>
inline void Calibration<Fad>::evaluate(int, int elem, GlobalResidual<Fad>& global,
                                       LocalResidual<Fad>& local, double const* iota_input,
                                       double w, double dv) {
  this->initialize_value_pt();
  int const facet_id_disp = D->mapping_disp[elem];
  int const node_id_load = D->mapping_load[elem][0];
  if ((facet_id_disp < 0) && (node_id_load < 0)) return;
  if (facet_id_disp > -1) {
    Fad mismatch;
    if (this->m_num_dims == 2) mismatch = compute_disp_mismatch(elem, global, iota_input);
    else mismatch = compute_surface_mismatch(elem, global, iota_input);
    this->value_pt += mismatch;
  }
  if (node_id_load > -1) {
    Fad load = compute_load(elem, global, local, iota_input, w, dv);
    this->value_pt += D->balance_factor * D->dt / D->total_time * D->load_mismatch * load;
  }
}

// ---------------------------------------------------------------------------
// src/reaction_mismatch.cpp.  Shares CalibrationData (plane, component, per-step measured load).
template <class T>
class ReactionMismatch : public QoI<T> {
 public:
  explicit ReactionMismatch(CalibrationData* d) : D(d) {}
  void before_elems(Disc const& disc, int step) override {   // :42-56, src/qoi.cpp:159-198
    QoI<T>::before_elems(disc, step);
    if (!D->initd) {
      D->mapping_load.assign(disc.n_elems, {});
      for (int e = 0; e < disc.n_elems; ++e) {
        std::vector<int> ids;
        for (int n = 0; n < disc.nn; ++n) {
          double const* x = disc.X(disc.conn[size_t(e) * disc.nn + n]);
          if (std::abs(x[D->coord_idx] - D->coord_value) < D->coord_tol) ids.push_back(n);
        }
        if (ids.empty()) ids.push_back(-1);
        D->mapping_load[e] = ids;
      }
      D->initd = true;
    }
  }
  // :58-103 and compute_torque :105-129
  T compute_load(int elem, GlobalResidual<T>& global, LocalResidual<T>& local, double const* iota,
                 double w, double dv) {
    T load_pt = T(0.);
    std::vector<int> const& node_ids = D->mapping_load[elem];
    global.zero_residual();
    global.evaluate(local, iota, w, dv, 0);
    int const c = D->reaction_force_comp;
    for (size_t i = 0; i < node_ids.size(); ++i) {
      int const n = node_ids[i];
      if (D->compute_torque) {
        double const* r = this->m_disc->X(this->m_disc->conn[size_t(elem) * this->m_disc->nn + n]);
        if (c == 2) load_pt += r[0] * global.R_nodal(0, n, 1) - r[1] * global.R_nodal(0, n, 0);
        else if (c == 0) load_pt += r[1] * global.R_nodal(0, n, 2) - r[2] * global.R_nodal(0, n, 1);
        else load_pt += r[2] * global.R_nodal(0, n, 0) - r[0] * global.R_nodal(0, n, 2);
      } else {
        load_pt += global.R_nodal(0, n, c);
      }
    }
    global.zero_residual();
    return load_pt;
  }
  void preprocess(int, int elem, GlobalResidual<T>& global, LocalResidual<T>& local,
                  double const* iota, double w, double dv) override {   // :155-175
    if (D->mapping_load[elem][0] < 0) return;
    D->total_load += val(compute_load(elem, global, local, iota, w, dv));
  }
  void preprocess_finalize(int) override {   // :131-153 (the "load out file" line is last_total_load)
    D->load_mismatch = D->total_load - D->load_meas;
    D->last_total_load = D->total_load;
    D->total_load = 0.;
  }
  void postprocess(double& J) override {     // :150-153
    D->J_disp = J;
    D->J_forc = 0.5 * std::pow(D->load_mismatch, 2) / D->n_ranks;
    J += D->J_forc;
  }
  void evaluate(int, int elem, GlobalResidual<T>& global, LocalResidual<T>& local,
                double const* iota, double w, double dv) override;   // :177-212
 private:
  CalibrationData* D;
};
template <>
inline void ReactionMismatch<double>::evaluate(int, int, GlobalResidual<double>&, LocalResidual<double>&,
                                               double const*, double, double) {
  this->initialize_value_pt();
}
template <>
inline void ReactionMismatch<Fad>::evaluate(int, int elem, GlobalResidual<Fad>& global,
                                            LocalResidual<Fad>& local, double const* iota, double w,
                                            double dv) {
  this->initialize_value_pt();
  if (D->mapping_load[elem][0] < 0) return;
  Fad load = compute_load(elem, global, local, iota, w, dv);
  this->value_pt = D->load_mismatch * load;
}

// outward unit normal of the facet (local vertex ids fv) of a tet: away from the opposite vertex
// (the role of ree::computeFaceOutwardNormal, src/load_mismatch.cpp:149)
inline void facet_outward_normal(ElemGeom const& g, int const* fv, double* N, double* area2) {
  double const* a = g.x[fv[0]]; double const* b = g.x[fv[1]]; double const* c = g.x[fv[2]];
  double u[3], v[3];
  for (int k = 0; k < 3; ++k) { u[k] = b[k] - a[k]; v[k] = c[k] - a[k]; }
  double n[3] = {u[1] * v[2] - u[2] * v[1], u[2] * v[0] - u[0] * v[2], u[0] * v[1] - u[1] * v[0]};
  double const len = std::sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
  int opp = 0;
  for (int k = 0; k < 4; ++k) if (k != fv[0] && k != fv[1] && k != fv[2]) opp = k;
  double d = 0.;
  for (int k = 0; k < 3; ++k) d += n[k] * (g.x[opp][k] - a[k]);
  double const sgn = d > 0. ? -1. : 1.;
  for (int k = 0; k < 3; ++k) N[k] = sgn * n[k] / len;
  *area2 = len;
}

// ---------------------------------------------------------------------------
// src/load_mismatch.cpp: the facet of the element on the side set is D->facet (3 local vertex ids
// in 3-D, 2 in 2-D with the third -1).
template <class T>
class LoadMismatch : public QoI<T> {
 public:
  explicit LoadMismatch(CalibrationData* d) : D(d) {}
  bool on_side(int elem) const { return !D->facet.empty() && D->facet[size_t(elem) * 3] >= 0; }
  // :79-171: one-point rule on the facet, P = J sigma F^-T (z stretch for plane stress), N.P.N w dv
  T compute_load(int elem, GlobalResidual<T>& global, LocalResidual<T>& local, double const* iota_input) {
    int const nd = this->m_num_dims;
    int const* fv = &D->facet[size_t(elem) * 3];
    static double const ref[4][3] = {{0, 0, 0}, {1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    double iota_elem[3] = {0., 0., 0.};
    for (int k = 0; k < nd; ++k)
      for (int c = 0; c < 3; ++c) iota_elem[c] += ref[fv[k]][c] / double(nd);   // facet centroid
    global.interpolate(iota_elem);
    Tensor<T> const grad_u = global.grad_vector_x(0);
    Tensor<T> stress = local.cauchy(global);
    if (local.is_finite_deformation()) {
      Tensor<T> const I = eye<T>(nd);
      Tensor<T> const F = grad_u + I;
      Tensor<T> const F_invT = transpose(inverse(F));
      T const J = det(F);
      stress = J * stress * F_invT;
      if (local.z_stretch_idx() > -1) stress = local.scalar_xi(local.z_stretch_idx()) * stress;
    }
    double N[3] = {0., 0., 0.}, wdv;
    ElemGeom const& g = global.geom();
    if (nd == 3) {
      double area2;
      facet_outward_normal(g, fv, N, &area2);
      wdv = 0.5 * area2;                       // w = 1/2, dv = |a x b|
    } else {
      N[0] = D->normal_2d[0]; N[1] = D->normal_2d[1];
      double l2 = 0.;
      for (int k = 0; k < 2; ++k) l2 += (g.x[fv[1]][k] - g.x[fv[0]][k]) * (g.x[fv[1]][k] - g.x[fv[0]][k]);
      wdv = std::sqrt(l2);                     // w = 2 on [-1, 1], dv = length / 2
    }
    T load_pt = T(0.);
    for (int i = 0; i < nd; ++i)
      for (int j = 0; j < nd; ++j) load_pt += N[i] * stress(i, j) * N[j];
    load_pt *= wdv;
    global.interpolate(iota_input);
    return load_pt;
  }
  void preprocess(int, int elem, GlobalResidual<T>& global, LocalResidual<T>& local,
                  double const* iota, double, double) override {   // :203-222
    if (!on_side(elem)) return;
    D->total_load += val(compute_load(elem, global, local, iota));
  }
  void preprocess_finalize(int) override {
    D->load_mismatch = D->total_load - D->load_meas;
    D->last_total_load = D->total_load;
    D->total_load = 0.;
  }
  void postprocess(double& J) override {
    D->J_disp = J;
    D->J_forc = 0.5 * std::pow(D->load_mismatch, 2) / D->n_ranks;
    J += D->J_forc;
  }
  void evaluate(int, int elem, GlobalResidual<T>& global, LocalResidual<T>& local,
                double const* iota, double, double) override;
 private:
  CalibrationData* D;
};
template <>
inline void LoadMismatch<double>::evaluate(int, int, GlobalResidual<double>&, LocalResidual<double>&,
                                           double const*, double, double) {
  this->initialize_value_pt();
}
template <>
inline void LoadMismatch<Fad>::evaluate(int, int elem, GlobalResidual<Fad>& global,
                                        LocalResidual<Fad>& local, double const* iota, double, double) {
  this->initialize_value_pt();
  if (!on_side(elem)) return;
  Fad load = compute_load(elem, global, local, iota);
  this->value_pt = D->load_mismatch * load;
}

// ---------------------------------------------------------------------------
// src/surface_mismatch.cpp:32-117 (3-D): sum over the order-2 facet rule of |u - u_meas|^2 w dv
template <class T>
class SurfaceMismatch : public QoI<T> {
 public:
  explicit SurfaceMismatch(CalibrationData* d) : D(d) {}
  void evaluate(int, int elem, GlobalResidual<T>& global, LocalResidual<T>&, double const* iota_input,
                double, double) override {
    this->initialize_value_pt();
    if (D->facet.empty() || D->facet[size_t(elem) * 3] < 0) return;
    int const* fv = &D->facet[size_t(elem) * 3];
    static double const ref[4][3] = {{0, 0, 0}, {1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    ElemGeom const& g = global.geom();
    double N3[3], area2;
    facet_outward_normal(g, fv, N3, &area2);
    auto const qps = quadrature(2, 2);
    Disc const& d = *this->m_disc;
    for (auto const& qp : qps) {
      double const Nf[3] = {1. - qp.xi[0] - qp.xi[1], qp.xi[0], qp.xi[1]};
      double iota_elem[3] = {0., 0., 0.};
      for (int k = 0; k < 3; ++k)
        for (int c = 0; c < 3; ++c) iota_elem[c] += Nf[k] * ref[fv[k]][c];
      global.interpolate(iota_elem);
      Vec<T> const u_fem = global.vector_x(0);
      double N[4], u_meas[3] = {0., 0., 0.};
      g.basis(iota_elem, N);
      for (int n = 0; n < d.nn; ++n) {
        int const node = d.conn[size_t(elem) * d.nn + n];
        for (int k = 0; k < 3; ++k) u_meas[k] += D->measured[size_t(node) * 3 + k] * N[n];
      }
      T const qoi = (u_fem[0] - u_meas[0]) * (u_fem[0] - u_meas[0]) +
                    (u_fem[1] - u_meas[1]) * (u_fem[1] - u_meas[1]) +
                    (u_fem[2] - u_meas[2]) * (u_fem[2] - u_meas[2]);
      this->value_pt += qoi * qp.w * area2;
    }
    global.interpolate(iota_input);
  }
 private:
  CalibrationData* D;
};

}  // namespace orc
