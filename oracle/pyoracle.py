"""ORACLE (test infrastructure only) -- ctypes front-end of oracle/liboracle.so.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module.  The product package
(calibr8_b200/) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np
import scipy.sparse as sp

_HERE = os.path.dirname(os.path.abspath(__file__))

LOCAL_TYPES = {
    "elastic": 0, "small_J2": 1, "small_hill": 2, "small_hill_plane_stress": 3,
    "hyper_J2": 4, "hyper_J2_plane_stress": 5, "small_hill_plane_strain": 6,
    "hyper_J2_plane_strain": 7, "hypo_hill": 8, "hypo_hill_plane_strain": 9,
    "hypo_hill_plane_stress": 10,
}
# parameter name order per model (init_params of each src/<model>.cpp)
PARAM_NAMES = {
    "elastic": ["E", "nu", "cte", "delta_T"],
    "small_J2": ["E", "nu", "K", "Y", "cte", "delta_T"],
    "small_hill": ["E", "nu", "Y", "R00", "R11", "R22", "R01", "R02", "R12", "S", "D"],
    "small_hill_plane_stress": ["E", "nu", "Y", "S", "D", "R00", "R11", "R22", "R01"],
    "small_hill_plane_strain": ["E", "nu", "Y", "S", "D", "R00", "R11", "R22", "R01"],
    "hyper_J2": ["E", "nu", "Y", "S", "D", "A", "n", "K"],
    "hyper_J2_plane_stress": ["E", "nu", "Y", "S", "D", "A", "n", "K"],
    "hyper_J2_plane_strain": ["E", "nu", "K", "Y", "Y_inf", "delta"],
    "hypo_hill": ["E", "nu", "Y", "R00", "R11", "R22", "R01", "R02", "R12", "S", "D"],
    "hypo_hill_plane_strain": ["E", "nu", "Y", "S", "D", "R00", "R11", "R22", "R01"],
    "hypo_hill_plane_stress": ["E", "nu", "Y", "S", "D", "R00", "R11", "R22", "R01", "Q00", "Q01", "Q10", "Q11"],
}
GLOBAL_TYPES = {"mechanics": 0, "mechanics_plane_stress": 1}


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    if force or not os.path.exists(so) or not os.path.exists(os.path.join(_HERE, "liboracle_count.so")):
        subprocess.check_call(["make", "-C", _HERE, "-j2"], stdout=subprocess.DEVNULL)
    return so


_libs = {}


def _lib(count=False):
    key = "count" if count else "plain"
    if key not in _libs:
        build()
        name = "liboracle_count.so" if count else "liboracle.so"
        lib = C.CDLL(os.path.join(_HERE, name))
        lib.orc_create.restype = C.c_void_p
        lib.orc_qoi.restype = C.c_double
        lib.orc_flops_reset.restype = C.c_longlong
        _libs[key] = lib
    return _libs[key]


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _pp(arrs):
    """list of arrays -> double** (keeps a reference to the arrays)."""
    n = len(arrs)
    T = C.c_void_p * max(n, 1)
    return T(*[a.ctypes.data for a in arrs])


class Oracle:
    """One calibr8 'State' worth of residual objects on a flat mesh."""

    def __init__(self, dim, conn, coords, elem_set=None, n_es=1, *, global_type="mechanics",
                 mixed=True, stab_mult=1.0, thickness=1.0, local_type="elastic", params=None,
                 max_iters=0, abs_tol=0.0, rel_tol=0.0, active=None, count_flops=False):
        self.lib = _lib(count_flops)
        self.dim = dim
        self.conn = np.ascontiguousarray(conn, dtype=np.int32)
        self.coords = np.ascontiguousarray(coords, dtype=np.float64)
        assert self.coords.shape[1] == 3
        self.n_elems, self.nn = self.conn.shape
        self.n_nodes = self.coords.shape[0]
        self.n_es = n_es
        self.elem_set = None if elem_set is None else np.ascontiguousarray(elem_set, dtype=np.int32)
        self.h = C.c_void_p(self.lib.orc_create(dim, self.n_elems, self.n_nodes, _p(self.conn),
                                                _p(self.coords), _p(self.elem_set), n_es))
        self.lib.orc_set_global(self.h, GLOBAL_TYPES[global_type], int(mixed),
                                C.c_double(stab_mult), C.c_double(thickness))
        self.local_type = local_type
        self.param_names = PARAM_NAMES[local_type]
        self.max_iters, self.abs_tol, self.rel_tol = max_iters, abs_tol, rel_tol
        self.active = active  # list per es of param indices
        self.set_params(params)
        info = (C.c_int * 7)()
        self.lib.orc_info(self.h, info)
        self.num_resid, neq0, neq1, self.n_xi, self.n_par, self.n_x, _ = list(info)
        self.neq = [neq0, neq1][: self.num_resid]
        self._graphs = {}

    def __del__(self):
        try:
            self.lib.orc_destroy(self.h)
        except Exception:
            pass

    # -- parameters -----------------------------------------------------
    def set_params(self, params):
        """params: [n_es, n_par] array (or list of dicts name->value per es)."""
        names = self.param_names
        if params is None:
            params = np.zeros((self.n_es, len(names)))
        if isinstance(params, (list, tuple)) and isinstance(params[0], dict):
            params = np.array([[d[k] for k in names] for d in params], dtype=np.float64)
        self.params = np.ascontiguousarray(np.atleast_2d(params), dtype=np.float64)
        assert self.params.shape == (self.n_es, len(names))
        if self.active is not None:
            ptr = np.zeros(self.n_es + 1, dtype=np.int32)
            for es in range(self.n_es):
                ptr[es + 1] = ptr[es] + len(self.active[es])
            idx = np.array([i for a in self.active for i in a], dtype=np.int32)
            self._act = (ptr, idx)
            aptr, aidx = _p(ptr), _p(idx)
        else:
            aptr = aidx = None
        self.lib.orc_set_local(self.h, LOCAL_TYPES[self.local_type], self.max_iters,
                               C.c_double(self.abs_tol), C.c_double(self.rel_tol),
                               _p(self.params), aptr, aidx)

    def set_time(self, t, dt):
        self.lib.orc_set_time(self.h, C.c_double(t), C.c_double(dt))

    # -- graphs -----------------------------------------------------------
    def graph(self, i, j):
        if (i, j) not in self._graphs:
            nnz = self.lib.orc_graph_nnz(self.h, i, j)
            nrows = self.lib.orc_graph_rows(self.h, i, j)
            rowptr = np.zeros(nrows + 1, dtype=np.int32)
            colind = np.zeros(nnz, dtype=np.int32)
            self.lib.orc_graph(self.h, i, j, _p(rowptr), _p(colind))
            self._graphs[(i, j)] = (rowptr, colind)
        return self._graphs[(i, j)]

    def node_graph(self):
        nnz = self.lib.orc_node_graph_nnz(self.h)
        rowptr = np.zeros(self.n_nodes + 1, dtype=np.int32)
        colind = np.zeros(nnz, dtype=np.int32)
        self.lib.orc_node_graph(self.h, _p(rowptr), _p(colind))
        return rowptr, colind

    def scatter_offsets(self, i, j):
        n = self.n_elems * (self.nn * self.neq[i]) * (self.nn * self.neq[j])
        out = np.zeros(n, dtype=np.int32)
        self.lib.orc_scatter_offsets(self.h, i, j, _p(out))
        return out

    def zeros_A(self):
        return [np.zeros(len(self.graph(i, j)[1])) for i in range(self.num_resid)
                for j in range(self.num_resid)]

    def zeros_b(self):
        return [np.zeros(self.n_nodes * self.neq[i]) for i in range(self.num_resid)]

    def csr(self, vals, i, j):
        rowptr, colind = self.graph(i, j)
        return sp.csr_matrix((vals, colind, rowptr),
                             shape=(self.n_nodes * self.neq[i], self.n_nodes * self.neq[j]))

    def bmat(self, A):
        nr = self.num_resid
        return sp.bmat([[self.csr(A[i * nr + j], i, j) for j in range(nr)] for i in range(nr)],
                       format="csr")

    # -- state ------------------------------------------------------------
    def init_xi(self):
        xi = np.zeros((self.n_elems, self.n_xi))
        self.lib.orc_init_xi(self.h, _p(xi))
        return xi

    def zeros_x(self):
        return [np.zeros(self.n_nodes * self.neq[i]) for i in range(self.num_resid)]

    # -- evaluations ------------------------------------------------------
    def forward_jacobian(self, x, x_prev, xi, xi_prev, *, assemble=True, element_out=False):
        """xi is used as the initial state and overwritten (copy made). Returns dict."""
        xi = np.ascontiguousarray(xi, dtype=np.float64).copy()
        xi_prev = np.ascontiguousarray(xi_prev, dtype=np.float64)
        A = self.zeros_A() if assemble else None
        b = self.zeros_b() if assemble else None
        ed = np.zeros((self.n_elems, self.n_x, self.n_x)) if element_out else None
        er = np.zeros((self.n_elems, self.n_x)) if element_out else None
        path = np.full(self.n_elems, -1, dtype=np.int32)
        iters = np.zeros(self.n_elems, dtype=np.int32)
        status = self.lib.orc_forward_jacobian(
            self.h, _pp(x), _pp(x_prev), _p(xi), _p(xi_prev),
            _pp(A) if assemble else None, _pp(b) if assemble else None,
            _p(ed), _p(er), _p(path), _p(iters))
        return dict(status=status, A=A, b=b, xi=xi, elem_dtotal=ed, elem_R=er, path=path,
                    iters=iters)

    def forward_jacobian_range(self, x, x_prev, xi, xi_prev, e0, e1, ed=None, er=None):
        """In-place element-range evaluation (cpu baseline threads); no global scatter."""
        return self.lib.orc_forward_jacobian_range(self.h, _pp(x), _pp(x_prev), _p(xi),
                                                   _p(xi_prev), _p(ed), _p(er), e0, e1)

    def global_residual(self, x, x_prev, xi, xi_prev):
        b = self.zeros_b()
        self.lib.orc_global_residual(self.h, _pp(x), _pp(x_prev), _p(xi), _p(xi_prev), _pp(b))
        return b

    # -- QoI ----------------------------------------------------------------
    def set_qoi_avg_disp(self):
        self.lib.orc_set_qoi_avg_disp(self.h)

    def set_qoi_calibration(self, *, balance_factor, coord_idx, coord_value, coord_tol=1e-12,
                            reaction_force_comp, weights=(1., 1., 1.), facet=None):
        w = np.zeros(3); w[: len(weights)] = weights
        self._facet = None if facet is None else np.ascontiguousarray(facet, dtype=np.int32)
        self.lib.orc_set_qoi_calibration(self.h, C.c_double(balance_factor), coord_idx,
                                         C.c_double(coord_value), C.c_double(coord_tol),
                                         reaction_force_comp, _p(w), _p(self._facet))

    def set_qoi_mismatch(self, kind, *, coord_idx=0, coord_value=0.0, coord_tol=1e-12, reaction_force_comp=0,
                         compute_torque=False, facet=None, normal_2d=None):
        """kind: 'reaction' (coordinate plane, component / torque), 'load' (facet [n_elems][3] local vertex
        ids of the side-set facet, 2-D: normal_2d), 'surface' (facet; measured field per step)"""
        k = {"reaction": 2, "load": 3, "surface": 4}[kind]
        self._facet = None if facet is None else np.ascontiguousarray(facet, dtype=np.int32)
        n2 = None if normal_2d is None else np.ascontiguousarray(normal_2d, dtype=np.float64)
        self.lib.orc_set_qoi_mismatch(self.h, k, coord_idx, C.c_double(coord_value), C.c_double(coord_tol),
                                      reaction_force_comp, int(compute_torque), _p(self._facet), _p(n2))

    def qoi_set_step(self, dt, total_time, load_meas, measured):
        self._measured = None if measured is None else np.ascontiguousarray(measured, dtype=np.float64)
        self.lib.orc_qoi_set_step(self.h, C.c_double(dt), C.c_double(total_time),
                                  C.c_double(load_meas), _p(self._measured))

    def calibration_state(self):
        out = np.zeros(5)
        self.lib.orc_qoi_calibration_state(self.h, _p(out))
        return dict(total_load=out[0], load_mismatch=out[1], J_disp=out[2], J_forc=out[3],
                    area=out[4])

    def qoi(self, x, x_prev, xi, xi_prev, step):
        return self.lib.orc_qoi(self.h, _pp(x), _pp(x_prev), _p(xi), _p(xi_prev), step)

    # -- adjoint --------------------------------------------------------------
    def adjoint_jacobian(self, x, x_prev, xi, xi_prev, g, f, step):
        """g is updated in place (g -= dJ/dxi). Returns (A^T blocks, rhs)."""
        A = self.zeros_A(); b = self.zeros_b()
        self.lib.orc_adjoint_jacobian(self.h, _pp(x), _pp(x_prev), _p(xi), _p(xi_prev), _p(g),
                                      _p(f), _pp(A), _pp(b), step)
        return A, b

    def adjoint_local(self, x, x_prev, xi, xi_prev, z, g, f):
        phi = np.zeros((self.n_elems, self.n_xi))
        self.lib.orc_adjoint_local(self.h, _pp(x), _pp(x_prev), _p(xi), _p(xi_prev), _pp(z),
                                   _p(phi), _p(g), _p(f))
        return phi

    def qoi_gradient(self, x, x_prev, xi, xi_prev, z, phi, grad_indices, n_grad, step):
        ptr = np.zeros(self.n_es + 1, dtype=np.int32)
        for es in range(self.n_es):
            ptr[es + 1] = ptr[es] + len(grad_indices[es])
        idx = np.array([i for a in grad_indices for i in a], dtype=np.int32)
        grad = np.zeros(n_grad)
        self.lib.orc_qoi_gradient(self.h, _pp(x), _pp(x_prev), _p(xi), _p(xi_prev), _pp(z),
                                  _p(phi), _p(ptr), _p(idx), _p(grad), n_grad, step)
        return grad

    # -- VFM ----------------------------------------------------------------
    def measured_residual(self, x, x_prev, xi, xi_prev):
        xi = xi.copy(); b = self.zeros_b()
        st = self.lib.orc_measured_residual(self.h, _pp(x), _pp(x_prev), _p(xi), _p(xi_prev), _pp(b))
        return st, b, xi

    def measured_residual_grad(self, x, x_prev, xi, xi_prev, local_sens, n_p):
        xi = xi.copy(); b = self.zeros_b()
        dR = [np.zeros((n_p, self.n_nodes * self.neq[i])) for i in range(self.num_resid)]
        st = self.lib.orc_measured_residual_grad(self.h, _pp(x), _pp(x_prev), _p(xi), _p(xi_prev),
                                                 _pp(b), _pp(dR), _p(local_sens))
        return st, b, dR, xi

    def vfm_adjoint_gradient(self, x, x_prev, xi, xi_prev, vf, hist, s, n_grad):
        grad = np.zeros(n_grad)
        self.lib.orc_vfm_adjoint_gradient(self.h, _pp(x), _pp(x_prev), _p(xi), _p(xi_prev),
                                          _pp(vf), _p(hist), C.c_double(s), _p(grad), n_grad)
        return grad

    def flops_reset(self):
        return self.lib.orc_flops_reset()


def quadrature(dim, order):
    lib = _lib()
    xi = np.zeros((8, 3)); w = np.zeros(8)
    n = lib.orc_quadrature(dim, order, _p(xi), _p(w))
    return xi[:n].copy(), w[:n].copy()
