// Virtual-fields-method objectives above the C ABI, restating
//   VirtualPower::compute_at_step / _forward_sens / _adjoint   src/virtual_power.cpp:109-203
//   FS_VFM_Objective::gradient                                 src/forward_sens_vfm_objective.cpp:66-115
//   Adjoint_VFM_Objective::gradient                            src/adjoint_sens_vfm_objective.cpp:67-124
// (single problem, parameters unscaled; the canonical [-1,1] scaling of src/objective.cpp:41-61 is
// applied by the caller).
#include "calibr8_host.hpp"

namespace c8host {

class VirtualPower {
 public:
  // measured: [num_steps][n_nodes][dim] host; w: [n_nodes][dim] host virtual field
  VirtualPower(Problem& p, const double* measured, const double* w) : P(p) {
    const size_t nd = size_t(P.n_nodes) * P.dim;
    meas.emplace_back(nd);  // step 0: zero field
    std::vector<double> zero(nd, 0.0);
    for (int s = 1; s <= P.num_steps; ++s) {
      meas.emplace_back(nd);
      P.check(c8_pack_x(P.ctx, measured + size_t(s - 1) * nd, nullptr, meas[s].get()), "c8_pack_x");
    }
    vf.resize(nd);
    P.check(c8_pack_x(P.ctx, w, nullptr, vf.get()), "c8_pack_x");
    P.allocate_history();  // xi[0] = initial state
  }
  // eval_measured_residual + R . w
  double compute_at_step(int step) {
    start_step(step);
    int nf = 0;
    P.check(c8_vfm_forward(P.ctx, meas[step].get(), meas[step - 1].get(), P.xi[step - 1].get(),
                           P.xi[step].get(), P.b.get(), nullptr, nullptr, &nf), "c8_vfm_forward");
    return P.dot(P.b.get(), vf.get());
  }
  void compute_at_step_forward_sens(int step, double& ivp, std::vector<double>& grad) {
    cudaStream_t s = (cudaStream_t)c8_get_stream(P.ctx);
    if (step == 1) {
      local_sens.resize(size_t(P.xi_ld) * P.nxi * P.npar);
      dR.resize(size_t(P.npar) * P.n_dofs);
    }
    start_step(step);
    cudaMemsetAsync(dR.get(), 0, dR.size() * sizeof(double), s);
    int nf = 0;
    P.check(c8_vfm_forward(P.ctx, meas[step].get(), meas[step - 1].get(), P.xi[step - 1].get(),
                           P.xi[step].get(), P.b.get(), dR.get(), local_sens.get(), &nf),
            "c8_vfm_forward");
    ivp = P.dot(P.b.get(), vf.get());
    grad.assign(P.npar, 0.0);
    for (int p = 0; p < P.npar; ++p) grad[p] = P.dot(dR.get() + size_t(p) * P.n_dofs, vf.get());
  }
  void compute_at_step_adjoint(int step, double scaled_mismatch, std::vector<double>& grad) {
    cudaStream_t s = (cudaStream_t)c8_get_stream(P.ctx);
    if (step == P.num_steps) hist.resize(size_t(P.xi_ld) * P.nxi);
    cudaMemsetAsync(P.work.get(), 0, 64 * sizeof(double), s);
    P.check(c8_vfm_adjoint(P.ctx, meas[step].get(), meas[step - 1].get(), P.xi[step].get(),
                           P.xi[step - 1].get(), vf.get(), scaled_mismatch, hist.get(),
                           P.work.get()), "c8_vfm_adjoint");
    P.check(c8_allreduce(P.ctx, P.work.get(), P.npar), "c8_allreduce");
    grad.assign(P.npar, 0.0);
    cudaMemcpyAsync(grad.data(), P.work.get(), P.npar * sizeof(double), cudaMemcpyDeviceToHost, s);
    cudaStreamSynchronize(s);
  }

 private:
  void start_step(int step) {
    cudaStream_t s = (cudaStream_t)c8_get_stream(P.ctx);
    // create_primal(step, use_measured): local state starts as a copy of step-1
    cudaMemcpyAsync(P.xi[step].get(), P.xi[step - 1].get(), size_t(P.xi_ld) * P.nxi * sizeof(double),
                    cudaMemcpyDeviceToDevice, s);
    cudaMemsetAsync(P.b.get(), 0, P.n_dofs * sizeof(double), s);
  }
  Problem& P;
  std::vector<DevVec> meas;
  DevVec vf, local_sens, dR, hist;
};

}  // namespace c8host

namespace c8host {

// mode 0: forward-sensitivity objective (FS_VFM), mode 1: adjoint-sensitivity objective.
// load_data [num_steps] = external virtual power per step; grad [npar] (all model parameters)
void vfm_evaluate(Problem& P, int mode, const double* measured_host, const double* w_host,
                  const double* load_data, double obj_scale_factor, double thickness, double& J,
                  std::vector<double>& grad) {
  VirtualPower vp(P, measured_host, w_host);
  const int N = P.num_steps;
  const double total_time = P.time(N) - P.time(0), dt = P.step_size;
  std::vector<double> gs;
  grad.assign(P.npar, 0.0);
  J = 0.0;
  if (mode == 0) {
    for (int step = 1; step <= N; ++step) {
      double ivp;
      vp.compute_at_step_forward_sens(step, ivp, gs);
      const double mismatch = thickness * ivp - load_data[step - 1];
      J += 0.5 * obj_scale_factor * dt / total_time * mismatch * mismatch;
      for (int p = 0; p < P.npar; ++p) grad[p] += gs[p] * mismatch * obj_scale_factor * dt / total_time;
    }
  } else {
    std::vector<double> ivp(N);
    for (int step = 1; step <= N; ++step) ivp[step - 1] = vp.compute_at_step(step);
    for (int step = N; step > 0; --step) {
      const double mismatch = ivp[step - 1] * thickness - load_data[step - 1];
      const double scaled = mismatch * obj_scale_factor * dt / total_time;
      J += 0.5 * mismatch * scaled;
      vp.compute_at_step_adjoint(step, scaled, gs);
      for (int p = 0; p < P.npar; ++p) grad[p] += gs[p];
    }
  }
}

// FS_VFM_Objective / Adjoint_VFM_Objective on canonical parameters
class VfmObjective : public Objective {
 public:
  VfmObjective(Problem& p, int mode, const std::vector<int>& active, const std::vector<double>& lo,
               const std::vector<double>& hi, const double* measured, const double* w,
               const double* loads, double scale, double thickness)
      : Objective(p, active, lo, hi), m_mode(mode), m_scale(scale), m_thickness(thickness) {
    const size_t nd = size_t(P.n_nodes) * P.dim;
    m_measured.assign(measured, measured + size_t(P.num_steps) * nd);
    m_w.assign(w, w + nd);
    m_loads.assign(loads, loads + P.num_steps);
  }
  double value(const std::vector<double>& p) override { evaluate(p); return m_J_old; }
  void gradient(std::vector<double>& g, const std::vector<double>& p) override {
    evaluate(p);
    std::vector<double> act(m_active.size());
    for (size_t i = 0; i < m_active.size(); ++i) act[i] = m_grad[m_active[i]];
    g = transform_gradient(act);
  }
 private:
  void evaluate(const std::vector<double>& p) {
    if (!param_diff(p)) return;
    set_params_from_canonical(p);
    vfm_evaluate(P, m_mode, m_measured.data(), m_w.data(), m_loads.data(), m_scale, m_thickness, m_J_old,
                 m_grad);
    m_p_old = p;
  }
  int m_mode;
  double m_scale, m_thickness;
  std::vector<double> m_measured, m_w, m_loads, m_grad;
};

}  // namespace c8host

using namespace c8host;

c8host::Objective* c8h_make_vfm_objective(c8host::Problem& P, int mode, const std::vector<int>& active,
                                          const std::vector<double>& lo, const std::vector<double>& hi,
                                          const double* measured, const double* w, const double* loads,
                                          double scale, double thickness) {
  return new VfmObjective(P, mode, active, lo, hi, measured, w, loads, scale, thickness);
}

extern "C" {

int c8h_vfm_objective(c8h_problem* h, int mode, const double* measured_host, const double* w_host,
                      const double* load_data, double obj_scale_factor, double thickness,
                      double* J_out, double* grad_out) {
  try {
    Problem& P = h->P;
    std::vector<double> grad;
    double J = 0.0;
    vfm_evaluate(P, mode, measured_host, w_host, load_data, obj_scale_factor, thickness, J, grad);
    *J_out = J;
    for (int p = 0; p < P.npar; ++p) grad_out[p] = grad[p];
    return 0;
  } catch (const std::exception& ex) {
    h->err = ex.what();
    return -1;
  }
}

}  // extern "C"
