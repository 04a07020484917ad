"""ORACLE (test infrastructure only) -- VFM objectives above the oracle's evaluation loops.

Restates VirtualPower (src/virtual_power.cpp:109-203), FS_VFM_Objective::gradient
(src/forward_sens_vfm_objective.cpp:66-115) and Adjoint_VFM_Objective::gradient
(src/adjoint_sens_vfm_objective.cpp:67-124) for a single problem, unscaled parameters."""
import numpy as np


def _fields(orc, measured):
    """measured [num_steps, n_nodes, dim] -> per-step x lists, step 0 = zero"""
    xs = [[np.zeros(orc.n_nodes * orc.dim)]]
    for m in measured:
        xs.append([np.ascontiguousarray(m.reshape(-1))])
    return xs


def vfm_objective(orc, mode, measured, w, load_data, n_active, obj_scale_factor=1.0, thickness=1.0,
                  step_size=1.0):
    N = len(load_data)
    xs = _fields(orc, measured)
    wv = [np.ascontiguousarray(w.reshape(-1))]
    total_time = N * step_size
    dt = step_size
    xi = [orc.init_xi()]
    J, grad = 0.0, np.zeros(n_active)
    if mode == "forward":
        sens = np.zeros((orc.n_elems, orc.n_xi * n_active))
        for step in range(1, N + 1):
            st, b, dR, xi_s = orc.measured_residual_grad(xs[step], xs[step - 1], xi[step - 1],
                                                        xi[step - 1], sens, n_active)
            xi.append(xi_s)
            ivp = float(b[0] @ wv[0])
            gs = dR[0] @ wv[0]
            mismatch = thickness * ivp - load_data[step - 1]
            J += 0.5 * obj_scale_factor * dt / total_time * mismatch ** 2
            grad += gs * mismatch * obj_scale_factor * dt / total_time
    else:
        ivp = []
        for step in range(1, N + 1):
            st, b, xi_s = orc.measured_residual(xs[step], xs[step - 1], xi[step - 1], xi[step - 1])
            xi.append(xi_s)
            ivp.append(float(b[0] @ wv[0]))
        hist = np.zeros((orc.n_elems, orc.n_xi))
        for step in range(N, 0, -1):
            mismatch = ivp[step - 1] * thickness - load_data[step - 1]
            scaled = mismatch * obj_scale_factor * dt / total_time
            J += 0.5 * mismatch * scaled
            grad += orc.vfm_adjoint_gradient(xs[step], xs[step - 1], xi[step], xi[step - 1], wv,
                                             hist, scaled, n_active)
    return J, grad
