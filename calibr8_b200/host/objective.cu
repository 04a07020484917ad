// Objectives on canonical [-1, 1] parameters, restating
//   Objective (transform_params / transform_gradient / param_diff)   src/objective.cpp:41-137
//   Adjoint_Objective::value / gradient                               src/adjoint_objective.cpp:24-118
//   FS_VFM_Objective, Adjoint_VFM_Objective                           src/forward_sens_vfm_objective.cpp,
//                                                                     src/adjoint_sens_vfm_objective.cpp
// The reference derives these from ROL::Objective<double>; here the same value(p) / gradient(g, p)
// pair is a plain C++ interface (and a C API for ctypes), so any optimiser can drive it
// (calibr8_b200/cli.py `inverse` uses L-BFGS-B on the canonical box like the reference's ROL set-up).
#include <cstring>
#include <memory>

#include "calibr8_host.hpp"

namespace c8host {

Objective::Objective(Problem& p, const std::vector<int>& active, const std::vector<double>& lower,
                     const std::vector<double>& upper)
    : P(p), m_active(active), m_lower(lower), m_upper(upper), m_p_old(active.size(), 2.0) {
  if (active.size() != lower.size() || active.size() != upper.size())
    throw std::runtime_error("objective: active / bounds size mismatch");
  for (int k : active)
    if (k < 0 || k >= P.npar) throw std::runtime_error("objective: active parameter index out of range");
  m_base.resize(P.npar);
  std::vector<double> all(size_t(P.npar) * 64);
  P.check(c8_get_params(P.ctx, all.data()), "c8_get_params");
  for (int k = 0; k < P.npar; ++k) m_base[k] = all[k];
}

std::vector<double> Objective::transform_params(const std::vector<double>& p, bool to_canonical) const {
  std::vector<double> out(p.size());
  for (size_t i = 0; i < p.size(); ++i) {
    const double span = 0.5 * (m_upper[i] - m_lower[i]), mean = 0.5 * (m_upper[i] + m_lower[i]);
    if (to_canonical) {
      out[i] = (p[i] - mean) / span;
      if (out[i] < -1.) out[i] = -1.;
      if (out[i] > 1.) out[i] = 1.;
    } else {
      out[i] = span * p[i] + mean;
    }
  }
  return out;
}

std::vector<double> Objective::transform_gradient(const std::vector<double>& g) const {
  std::vector<double> out(g.size());
  for (size_t i = 0; i < g.size(); ++i) out[i] = 0.5 * (m_upper[i] - m_lower[i]) * g[i];
  return out;
}

std::vector<double> Objective::active_params() const {
  std::vector<double> all(size_t(P.npar) * 64), out(m_active.size());
  P.check(c8_get_params(P.ctx, all.data()), "c8_get_params");
  for (size_t i = 0; i < m_active.size(); ++i) out[i] = all[m_active[i]];
  return out;
}

bool Objective::param_diff(const std::vector<double>& p) const {  // src/objective.cpp:125-137
  for (size_t i = 0; i < p.size(); ++i)
    if (std::fabs(p[i] - m_p_old[i]) > m_difftol) return true;
  return false;
}

void Objective::set_params_from_canonical(const std::vector<double>& p) {
  const std::vector<double> phys = transform_params(p, false);
  std::vector<double> all = m_base;
  for (size_t i = 0; i < m_active.size(); ++i) all[m_active[i]] = phys[i];
  P.check(c8_set_params(P.ctx, all.data()), "c8_set_params");
}

// ---- Adjoint_Objective -----------------------------------------------------------------------
double AdjointObjective::value(const std::vector<double>& p) {
  if (param_diff(p)) {
    set_params_from_canonical(p);
    Primal pr(P);
    m_J_old = pr.solve_all();
    m_p_old = p;
  }
  return m_J_old;
}

void AdjointObjective::gradient(std::vector<double>& g, const std::vector<double>& p) {
  value(p);  // forward solve only when the parameters changed (the history is reused otherwise)
  Adjoint ad(P);
  std::vector<double> all;
  ad.gradient(all);
  std::vector<double> act(m_active.size());
  for (size_t i = 0; i < m_active.size(); ++i) act[i] = all[m_active[i]];
  g = transform_gradient(act);
}

}  // namespace c8host

using namespace c8host;

struct c8h_objective {
  std::unique_ptr<Objective> obj;
  std::string err;
};
struct c8h_problem;
c8host::Problem& c8h_problem_ref(c8h_problem* h);  // host.cu
c8host::Objective* c8h_make_vfm_objective(c8host::Problem& P, int mode, const std::vector<int>& active,
                                          const std::vector<double>& lo, const std::vector<double>& hi,
                                          const double* measured, const double* w, const double* loads,
                                          double scale, double thickness);  // vfm_host.cu

extern "C" {

// type 0: adjoint (PDE-constrained) objective with the problem's QoI; 1: forward-sensitivity VFM;
// 2: adjoint-sensitivity VFM (measured [num_steps][n_nodes][dim], w [n_nodes][dim], loads [num_steps])
c8h_objective* c8h_objective_create(c8h_problem* h, int type, int n_active, const int32_t* active,
                                    const double* lower, const double* upper, const double* measured,
                                    const double* w, const double* loads, double obj_scale_factor,
                                    double thickness) {
  auto* o = new c8h_objective;
  try {
    Problem& P = c8h_problem_ref(h);
    std::vector<int> act(active, active + n_active);
    std::vector<double> lo(lower, lower + n_active), hi(upper, upper + n_active);
    if (type == 0) o->obj.reset(new AdjointObjective(P, act, lo, hi));
    else o->obj.reset(c8h_make_vfm_objective(P, type == 1 ? 0 : 1, act, lo, hi, measured, w, loads,
                                             obj_scale_factor, thickness));
  } catch (const std::exception& ex) {
    o->err = ex.what();
  }
  return o;
}
void c8h_objective_destroy(c8h_objective* o) { delete o; }
const char* c8h_objective_error(c8h_objective* o) { return o->err.c_str(); }

#define C8O_TRY(o, body)                                        \
  if (!(o)->obj) return -1;                                      \
  try { body; return 0; }                                        \
  catch (const std::exception& ex) { (o)->err = ex.what(); return -1; }

int c8h_objective_value(c8h_objective* o, const double* p, int n, double* J) {
  C8O_TRY(o, { *J = o->obj->value(std::vector<double>(p, p + n)); });
}
int c8h_objective_gradient(c8h_objective* o, const double* p, int n, double* g) {
  C8O_TRY(o, {
    std::vector<double> gv;
    o->obj->gradient(gv, std::vector<double>(p, p + n));
    std::memcpy(g, gv.data(), n * sizeof(double));
  });
}
// to_canonical != 0: physical -> canonical (clamped to [-1, 1]); else canonical -> physical
int c8h_objective_transform(c8h_objective* o, const double* in, int n, int to_canonical, double* out) {
  C8O_TRY(o, {
    const std::vector<double> r = o->obj->transform_params(std::vector<double>(in, in + n), to_canonical != 0);
    std::memcpy(out, r.data(), n * sizeof(double));
  });
}
int c8h_objective_active_params(c8h_objective* o, double* out, int n) {
  C8O_TRY(o, {
    const std::vector<double> r = o->obj->active_params();
    std::memcpy(out, r.data(), n * sizeof(double));
  });
}

}  // extern "C"
