"""Virtual-fields-method entry points (ctypes): the device kernels c8_vfm_forward / c8_vfm_adjoint
and the C++ host objective c8h_vfm_objective (calibr8_b200/host/vfm_host.cu)."""
import ctypes as C

import numpy as np

from .capi import SYMBOLS, HOST_SYMBOLS, _dp, _hp

SYMBOLS += ["c8_vfm_forward", "c8_vfm_adjoint"]
HOST_SYMBOLS += ["c8h_vfm_objective"]


def vfm_forward(ctx, x_meas, x_meas_prev, xi_prev, xi, b, dR=None, local_sens=None):
    nf = C.c_int(0)
    ctx._check(ctx.lib.c8_vfm_forward(ctx.h, _dp(x_meas), _dp(x_meas_prev), _dp(xi_prev), _dp(xi),
                                      _dp(b), _dp(dR), _dp(local_sens), C.byref(nf)))
    return nf.value


def vfm_adjoint(ctx, x_meas, x_meas_prev, xi, xi_prev, w, s, hist, grad):
    ctx._check(ctx.lib.c8_vfm_adjoint(ctx.h, _dp(x_meas), _dp(x_meas_prev), _dp(xi), _dp(xi_prev),
                                      _dp(w), C.c_double(s), _dp(hist), _dp(grad)))


def vfm_objective(host_problem, mode, measured, w, load_data, obj_scale_factor=1.0, thickness=1.0):
    """mode 'forward' (FS_VFM) or 'adjoint' (Adjoint_VFM).  measured [num_steps, n_nodes, dim],
    w [n_nodes, dim].  Returns (J, grad[npar])."""
    hp = host_problem
    measured = np.ascontiguousarray(measured, dtype=np.float64)
    w = np.ascontiguousarray(w, dtype=np.float64)
    load_data = np.ascontiguousarray(load_data, dtype=np.float64)
    J = C.c_double(0)
    g = np.zeros(hp.ctx.npar)
    hp._check(hp.lib.c8h_vfm_objective(hp.h, 0 if mode == "forward" else 1, _hp(measured), _hp(w),
                                       _hp(load_data), C.c_double(obj_scale_factor),
                                       C.c_double(thickness), C.byref(J), _hp(g)))
    return J.value, g
