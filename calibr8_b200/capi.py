"""ctypes binding of calibr8_b200/lib/libc8b200.so (the C ABI of include/c8b200.h).

This is the only way Python reaches the product: there is no CPU fallback.  If the
shared library is missing or no CUDA device is present, construction fails loudly.
PyTorch is used by callers only to own device memory (``tensor.data_ptr()``) and streams.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# C8B200_LIB selects an alternative build of the same library (kernel-tuning variants)
LIB_PATH = os.environ.get("C8B200_LIB") or os.path.join(_HERE, "lib", "libc8b200.so")

GLOBAL_TYPES = {"mechanics": 0, "mechanics_plane_stress": 1}
LOCAL_TYPES = {
    "elastic": 0, "small_J2": 1, "small_hill": 2, "small_hill_plane_stress": 3,
    "hyper_J2": 4, "hyper_J2_plane_stress": 5, "small_hill_plane_strain": 6,
    "hyper_J2_plane_strain": 7, "hypo_hill": 8, "hypo_hill_plane_strain": 9, "hypo_hill_plane_stress": 10,
}
# parameter order per model = LocalResidual::init_params of each reference model file
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

# every symbol include/c8b200.h declares (checked by tests/test_capi_symbols.py)
SYMBOLS = [
    "c8_create", "c8_destroy", "c8_last_error", "c8_version", "c8_set_mesh", "c8_set_model",
    "c8_set_params", "c8_get_params", "c8_info", "c8_bsr_pattern", "c8_bsr_pattern_dev", "c8_csr_block_size",
    "c8_csr_block_pattern", "c8_csr_block_values", "c8_pack_x", "c8_unpack_x", "c8_pack_xi",
    "c8_unpack_xi", "c8_init_xi", "c8_forward_jacobian", "c8_forward_jacobian_elem",
    "c8_global_residual", "c8_forward_jacobian_host", "c8_resident_matrix", "c8_set_stream",
    "c8_synchronize", "c8_state_set_prev", "c8_state_forward_jacobian", "c8_state_get_xi",
    "c8_state_ptrs", "c8_bench_dfma", "c8_bench_copy", "c8_set_assembly_chunk",
]

_lib = None


class C8Error(RuntimeError):
    pass


def load_library():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise C8Error(f"{LIB_PATH} is missing: build it with `python -c 'import "
                          f"__graft_entry__ as g; g.build()'` (there is no CPU fallback)")
        lib = C.CDLL(LIB_PATH)
        lib.c8_create.restype = C.c_void_p
        lib.c8_create.argtypes = [C.c_int]
        lib.c8_last_error.restype = C.c_char_p
        lib.c8_last_error.argtypes = [C.c_void_p]
        lib.c8_version.restype = C.c_char_p
        lib.c8_destroy.argtypes = [C.c_void_p]
        _lib = lib
    return _lib


def _hp(a):
    """host numpy array -> void*"""
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _dp(t):
    """device pointer of a torch tensor (or int / None)"""
    if t is None:
        return None
    if isinstance(t, int):
        return C.c_void_p(t)
    return C.c_void_p(t.data_ptr())


class Context:
    """One GPU's resident discretisation + model (the role of calibr8's State/Disc for this path)."""

    def __init__(self, device=0):
        self.lib = load_library()
        self.h = self.lib.c8_create(device)
        if not self.h:
            raise C8Error("c8_create failed: no CUDA device (this library has no CPU path)")
        self.h = C.c_void_p(self.h)
        self.device = device
        self._keep = []

    def close(self):
        if getattr(self, "h", None):
            self.lib.c8_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, allow_local_fail=False):
        if rc == 0 or (allow_local_fail and rc == -1):
            return rc
        raise C8Error(f"c8 error {rc}: {self.lib.c8_last_error(self.h).decode()}")

    # ---- set-up ------------------------------------------------------------------
    def set_assembly_chunk(self, chunk_elems):
        """before set_mesh: elements per chunk of the two-phase assembly (0 one pass, -1 default)"""
        self._check(self.lib.c8_set_assembly_chunk(self.h, int(chunk_elems)))

    def set_mesh(self, dim, conn, coords, elem_set=None, n_elem_sets=1):
        conn = np.ascontiguousarray(conn, dtype=np.int32)
        coords = np.ascontiguousarray(coords, dtype=np.float64)
        assert coords.shape[1] == 3 and conn.shape[1] == dim + 1
        es = None if elem_set is None else np.ascontiguousarray(elem_set, dtype=np.int32)
        self.dim, self.n_elems, self.n_nodes = dim, conn.shape[0], coords.shape[0]
        self.n_es = n_elem_sets
        self._check(self.lib.c8_set_mesh(self.h, dim, self.n_elems, self.n_nodes, _hp(conn),
                                         _hp(coords), _hp(es), n_elem_sets))

    def set_model(self, global_type, local_type, params, *, max_iters=0, abs_tol=0.0, rel_tol=0.0,
                  stab_mult=1.0, thickness=1.0):
        self.global_type, self.local_type = global_type, local_type
        self.param_names = PARAM_NAMES[local_type]
        p = self._params_array(params)
        self._check(self.lib.c8_set_model(self.h, GLOBAL_TYPES[global_type], LOCAL_TYPES[local_type],
                                          _hp(p), max_iters, C.c_double(abs_tol),
                                          C.c_double(rel_tol), C.c_double(stab_mult),
                                          C.c_double(thickness)))
        out = (C.c_int64 * 12)()
        self._check(self.lib.c8_info(self.h, out))
        (_, self.nn, self.nb, self.nx, self.nxi, self.npar, _, _, self.nnzb, self.n_dofs,
         self.group, self.xi_ld) = [int(v) for v in out]
        self.num_resid = 2 if self.nb > self.dim else 1
        self.neq = [self.dim, 1][: self.num_resid]

    def _params_array(self, params):
        names = self.param_names
        if isinstance(params, dict):
            params = [params]
        if isinstance(params, (list, tuple)) and isinstance(params[0], dict):
            params = [[d[k] for k in names] for d in params]
        p = np.ascontiguousarray(np.atleast_2d(np.asarray(params, dtype=np.float64)))
        assert p.shape == (self.n_es, len(names)), p.shape
        return p

    def set_params(self, params):
        self._check(self.lib.c8_set_params(self.h, _hp(self._params_array(params))))

    # ---- patterns ------------------------------------------------------------------
    def bsr_pattern(self):
        rowptr = np.zeros(self.n_nodes + 1, dtype=np.int32)
        colind = np.zeros(self.nnzb, dtype=np.int32)
        self._check(self.lib.c8_bsr_pattern(self.h, _hp(rowptr), _hp(colind)))
        return rowptr, colind

    def csr_block_pattern(self, i, j):
        nr, nnz = C.c_int64(), C.c_int64()
        self._check(self.lib.c8_csr_block_size(self.h, i, j, C.byref(nr), C.byref(nnz)))
        rowptr = np.zeros(nr.value + 1, dtype=np.int32)
        colind = np.zeros(nnz.value, dtype=np.int32)
        self._check(self.lib.c8_csr_block_pattern(self.h, i, j, _hp(rowptr), _hp(colind)))
        return rowptr, colind

    def csr_block_values(self, i, j, A_dev):
        nr, nnz = C.c_int64(), C.c_int64()
        self._check(self.lib.c8_csr_block_size(self.h, i, j, C.byref(nr), C.byref(nnz)))
        vals = np.zeros(nnz.value)
        self._check(self.lib.c8_csr_block_values(self.h, i, j, _dp(A_dev), _hp(vals)))
        return vals

    # ---- layout helpers ------------------------------------------------------------
    def pack_x(self, u, p, x_dev):
        u = np.ascontiguousarray(u, dtype=np.float64)
        p = None if p is None else np.ascontiguousarray(p, dtype=np.float64)
        self._check(self.lib.c8_pack_x(self.h, _hp(u), _hp(p), _dp(x_dev)))

    def unpack_x(self, x_dev):
        u = np.zeros(self.n_nodes * self.dim)
        p = np.zeros(self.n_nodes) if self.num_resid == 2 else None
        self._check(self.lib.c8_unpack_x(self.h, _dp(x_dev), _hp(u), _hp(p)))
        return [u, p][: self.num_resid]

    def pack_xi(self, xi_aos, xi_dev):
        xi_aos = np.ascontiguousarray(xi_aos, dtype=np.float64)
        self._check(self.lib.c8_pack_xi(self.h, _hp(xi_aos), _dp(xi_dev)))

    def unpack_xi(self, xi_dev):
        xi = np.zeros((self.n_elems, self.nxi))
        self._check(self.lib.c8_unpack_xi(self.h, _dp(xi_dev), _hp(xi)))
        return xi

    def init_xi(self, xi_dev):
        self._check(self.lib.c8_init_xi(self.h, _dp(xi_dev)))

    # ---- device allocation helpers (torch owns the memory) ---------------------------
    def alloc(self, kind):
        import torch
        dev = torch.device("cuda", self.device)
        n = {"x": self.n_dofs, "b": self.n_dofs, "xi": self.xi_ld * self.nxi,
             "A": self.nnzb * self.nb * self.nb, "path": self.n_elems,
             "elem_J": self.n_elems * self.nx * self.nx, "elem_R": self.n_elems * self.nx}[kind]
        dt = torch.int8 if kind == "path" else torch.float64
        t = torch.zeros(n, dtype=dt, device=dev)
        # the library launches on its own non-blocking stream unless set_stream() binds torch's:
        # make sure the fill has landed before the tensor is handed to it
        torch.cuda.current_stream(dev).synchronize()
        return t

    # ---- hot path ----------------------------------------------------------------------
    def forward_jacobian(self, x, x_prev, xi_prev, xi, A=None, b=None, path=None, elem_J=None,
                         elem_R=None, check=True):
        """Device tensors in/out.  Returns the number of failed local solves."""
        nf = C.c_int(0)
        if elem_J is not None or elem_R is not None:
            rc = self.lib.c8_forward_jacobian_elem(self.h, _dp(x), _dp(x_prev), _dp(xi_prev), _dp(xi),
                                                   _dp(A), _dp(b), _dp(path), _dp(elem_J),
                                                   _dp(elem_R), C.byref(nf))
        else:
            rc = self.lib.c8_forward_jacobian(self.h, _dp(x), _dp(x_prev), _dp(xi_prev), _dp(xi),
                                              _dp(A), _dp(b), _dp(path),
                                              C.byref(nf) if check else None)
        self._check(rc, allow_local_fail=True)
        return nf.value

    def global_residual(self, x, x_prev, xi, xi_prev, b):
        self._check(self.lib.c8_global_residual(self.h, _dp(x), _dp(x_prev), _dp(xi), _dp(xi_prev),
                                                _dp(b)))

    def forward_jacobian_host(self, u, p, u_prev, p_prev, xi_prev, xi):
        """Host numpy arrays in the reference's layout; returns (n_failed, xi, [b_u, b_p])."""
        xi = np.ascontiguousarray(xi, dtype=np.float64).copy()
        bu = np.zeros(self.n_nodes * self.dim)
        bp = np.zeros(self.n_nodes) if self.num_resid == 2 else None
        nf = C.c_int(0)
        c = np.ascontiguousarray
        rc = self.lib.c8_forward_jacobian_host(self.h, _hp(c(u)), _hp(None if p is None else c(p)),
                                               _hp(c(u_prev)),
                                               _hp(None if p_prev is None else c(p_prev)),
                                               _hp(c(xi_prev)), _hp(xi), _hp(bu), _hp(bp),
                                               C.byref(nf))
        self._check(rc, allow_local_fail=True)
        return nf.value, xi, [bu, bp][: self.num_resid]

    # ---- resident step state (host buffers in/out; what a calibr8 caller does per Newton iteration)
    def state_set_prev(self, u_prev, p_prev, xi_prev):
        c = np.ascontiguousarray
        self._check(self.lib.c8_state_set_prev(self.h, _hp(c(u_prev)),
                                               _hp(None if p_prev is None else c(p_prev)),
                                               _hp(c(xi_prev))))

    def state_forward_jacobian(self, u, p, b_u, b_p):
        """u, p, b_u, b_p: C-contiguous float64 host arrays (pinned for full copy speed)."""
        nf = C.c_int(0)
        rc = self.lib.c8_state_forward_jacobian(self.h, _hp(u), _hp(p), _hp(b_u), _hp(b_p),
                                                C.byref(nf))
        self._check(rc, allow_local_fail=True)
        return nf.value

    def state_get_xi(self):
        xi = np.zeros((self.n_elems, self.nxi))
        self._check(self.lib.c8_state_get_xi(self.h, _hp(xi)))
        return xi

    def state_ptrs(self):
        ps = [C.c_void_p() for _ in range(6)]
        self._check(self.lib.c8_state_ptrs(self.h, *[C.byref(p) for p in ps]))
        return dict(zip(["x", "x_prev", "xi", "xi_prev", "A", "b"], [p.value for p in ps]))

    def bench_dfma(self, iters=4096):
        v = C.c_double(0)
        self._check(self.lib.c8_bench_dfma(self.h, iters, C.byref(v)))
        return v.value

    def bench_copy(self):
        v = C.c_double(0)
        self._check(self.lib.c8_bench_copy(self.h, C.byref(v)))
        return v.value

    def resident_matrix_ptr(self):
        p = C.c_void_p()
        self._check(self.lib.c8_resident_matrix(self.h, C.byref(p)))
        return p.value

    def set_stream(self, cuda_stream_ptr):
        self._check(self.lib.c8_set_stream(self.h, C.c_void_p(cuda_stream_ptr)))

    def synchronize(self):
        self._check(self.lib.c8_synchronize(self.h))


# ======================================================================================
# adjoint / QoI / linear-algebra entry points and the C++ host step solvers (c8h_*)
SYMBOLS += [
    "c8_adjoint_jacobian", "c8_adjoint_local", "c8_qoi_value", "c8_qoi_gradient", "c8_spmv",
    "c8_dot", "c8_axpby", "c8_apply_dbc", "c8_apply_tbc", "c8_gmres", "c8_linalg_release", "c8_get_coords",
    "c8_get_conn", "c8_get_stream",
]
HOST_SYMBOLS = [
    "c8h_create", "c8h_destroy", "c8h_last_error", "c8h_set_time", "c8h_add_dbc", "c8h_add_tbc",
    "c8h_finalize_dbcs", "c8h_set_solver", "c8h_set_line_search", "c8h_set_qoi_avg_disp", "c8h_set_qoi_calibration",
    "c8h_set_qoi_mismatch", "c8h_get_loads", "c8h_primal_solve", "c8h_adjoint_gradient", "c8h_get_step", "c8h_get_adjoint_step", "c8h_stats",
    "c8h_profile", "c8h_eval_expr", "c8h_describe_residuals",
]


class C8Qoi(C.Structure):
    """struct c8_qoi of include/c8b200.h"""
    _fields_ = [("type", C.c_int), ("weights", C.c_double * 3), ("balance_factor", C.c_double),
                ("dt_over_T", C.c_double), ("inv_area", C.c_double), ("load_mismatch", C.c_double),
                ("coord_idx", C.c_int), ("coord_value", C.c_double), ("coord_tol", C.c_double),
                ("reaction_force_comp", C.c_int), ("measured_dev", C.c_void_p),
                ("facet_dev", C.c_void_p), ("compute_torque", C.c_int), ("normal_2d", C.c_double * 2)]


QOI_TYPES = {"avg_disp": 0, "calibration": 1, "reaction_mismatch": 2, "load_mismatch": 3, "surface_mismatch": 4}


def make_qoi(kind="avg_disp", **kw):
    q = C8Qoi()
    q.type = QOI_TYPES[kind]
    w = kw.get("weights", (1., 1., 1.))
    for k in range(3):
        q.weights[k] = w[k] if k < len(w) else 1.0
    q.balance_factor = kw.get("balance_factor", 1.0)
    q.dt_over_T = kw.get("dt_over_T", 1.0)
    q.inv_area = kw.get("inv_area", 1.0)
    q.load_mismatch = kw.get("load_mismatch", 0.0)
    q.coord_idx = kw.get("coord_idx", 0)
    q.coord_value = kw.get("coord_value", 0.0)
    q.coord_tol = kw.get("coord_tol", 1e-12)
    q.reaction_force_comp = kw.get("reaction_force_comp", 0)
    m = kw.get("measured")
    q.measured_dev = None if m is None else m.data_ptr()
    f = kw.get("facet")
    q.facet_dev = None if f is None else f.data_ptr()
    q.compute_torque = int(bool(kw.get("compute_torque", False)))
    n2 = kw.get("normal_2d", (0., 0.))
    q.normal_2d[0], q.normal_2d[1] = float(n2[0]), float(n2[1])
    return q


def _ctx_adjoint_jacobian(self, qoi, x, x_prev, xi, xi_prev, g, f, AT, rhs):
    self._check(self.lib.c8_adjoint_jacobian(self.h, C.byref(qoi) if qoi is not None else None,
                                             _dp(x), _dp(x_prev), _dp(xi), _dp(xi_prev), _dp(g),
                                             _dp(f), _dp(AT), _dp(rhs)))


def _ctx_adjoint_local(self, x, x_prev, xi, xi_prev, z, phi, g, f):
    self._check(self.lib.c8_adjoint_local(self.h, _dp(x), _dp(x_prev), _dp(xi), _dp(xi_prev),
                                          _dp(z), _dp(phi), _dp(g), _dp(f)))


def _ctx_qoi_value(self, qoi, x, x_prev, xi, xi_prev, mode, scalars):
    self._check(self.lib.c8_qoi_value(self.h, C.byref(qoi) if qoi is not None else None, _dp(x),
                                      _dp(x_prev), _dp(xi), _dp(xi_prev), mode, _dp(scalars)))


def _ctx_qoi_gradient(self, qoi, x, x_prev, xi, xi_prev, z, phi, grad):
    self._check(self.lib.c8_qoi_gradient(self.h, C.byref(qoi) if qoi is not None else None, _dp(x),
                                         _dp(x_prev), _dp(xi), _dp(xi_prev), _dp(z), _dp(phi),
                                         _dp(grad)))


def _ctx_spmv(self, A, x, y):
    self._check(self.lib.c8_spmv(self.h, _dp(A), _dp(x), _dp(y)))


def _ctx_dot(self, x, y):
    v = C.c_double(0)
    self._check(self.lib.c8_dot(self.h, _dp(x), _dp(y), C.byref(v)))
    return v.value


def _ctx_apply_dbc(self, A, R, x, nodes, eqs, vals, is_adjoint=False):
    self._check(self.lib.c8_apply_dbc(self.h, _dp(A), _dp(R), _dp(x), _dp(nodes), _dp(eqs),
                                      _dp(vals), int(nodes.numel()), int(is_adjoint)))


def _ctx_apply_tbc(self, R, side_nodes, traction):
    """side_nodes int32 device [n_sides][dim], traction float64 device [n_sides][dim]"""
    self._check(self.lib.c8_apply_tbc(self.h, _dp(R), _dp(side_nodes), _dp(traction),
                                      int(side_nodes.numel()) // self.dim))


def _ctx_gmres(self, A, b, x, restart=100, max_iters=5000, rel_tol=1e-10, abs_tol=0.0):
    info = (C.c_double * 3)()
    rc = self.lib.c8_gmres(self.h, _dp(A), _dp(b), _dp(x), restart, max_iters, C.c_double(rel_tol),
                           C.c_double(abs_tol), info)
    if rc not in (0, -4):
        self._check(rc)
    return dict(converged=(rc == 0), iters=int(info[0]), resid=info[1], resid0=info[2])


Context.adjoint_jacobian = _ctx_adjoint_jacobian
Context.adjoint_local = _ctx_adjoint_local
Context.qoi_value = _ctx_qoi_value
Context.qoi_gradient = _ctx_qoi_gradient
Context.spmv = _ctx_spmv
Context.dot = _ctx_dot
Context.apply_dbc = _ctx_apply_dbc
Context.apply_tbc = _ctx_apply_tbc
Context.gmres = _ctx_gmres


class HostProblem:
    """The C++ host step solvers (calibr8_b200/host): Primal / Adjoint above the C ABI."""

    def __init__(self, ctx: Context):
        self.ctx = ctx
        self.lib = ctx.lib
        self.lib.c8h_create.restype = C.c_void_p
        self.lib.c8h_last_error.restype = C.c_char_p
        self.lib.c8h_last_error.argtypes = [C.c_void_p]
        h = self.lib.c8h_create(ctx.h)
        if not h:
            raise C8Error("c8h_create failed")
        self.h = C.c_void_p(h)

    def close(self):
        if getattr(self, "h", None):
            self.lib.c8h_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise C8Error("c8h: " + self.lib.c8h_last_error(self.h).decode())

    def set_time(self, num_steps, step_size=1.0):
        self.num_steps = num_steps
        self._check(self.lib.c8h_set_time(self.h, num_steps, C.c_double(step_size)))

    def add_dbc(self, resid, eq, nodes, expr):
        nodes = np.ascontiguousarray(nodes, dtype=np.int32)
        self._check(self.lib.c8h_add_dbc(self.h, resid, eq, _hp(nodes), int(nodes.size),
                                         str(expr).encode()))

    def finalize_dbcs(self):
        self._check(self.lib.c8h_finalize_dbcs(self.h))

    def add_tbc(self, resid, side_nodes, exprs):
        """traction bc [resid, side set, x-val, y-val(, z-val)]: side_nodes [n_sides][dim] node ids"""
        sn = np.ascontiguousarray(side_nodes, dtype=np.int32).reshape(-1, self.ctx.dim)
        ex = ";".join(str(e) for e in list(exprs)[: self.ctx.dim])
        self._check(self.lib.c8h_add_tbc(self.h, resid, _hp(sn), int(sn.shape[0]), ex.encode()))

    def set_solver(self, newton_max_iters=15, abs_tol=1e-8, rel_tol=1e-8, gmres_restart=100,
                   gmres_max_iters=4000, linear_tol=1e-10, verbose=False):
        self._check(self.lib.c8h_set_solver(self.h, newton_max_iters, C.c_double(abs_tol),
                                            C.c_double(rel_tol), gmres_restart, gmres_max_iters,
                                            C.c_double(linear_tol), int(verbose)))

    def set_line_search(self, sufficient_decrease=1e-4, min_backtrack=0.5, max_backtrack=0.9, max_evals=4):
        """the deck's `line search` sublist of the global residual (src/line_search.hpp:33-49)"""
        self._check(self.lib.c8h_set_line_search(self.h, C.c_double(sufficient_decrease), C.c_double(min_backtrack),
                                                 C.c_double(max_backtrack), int(max_evals)))

    def set_qoi_avg_disp(self):
        self._check(self.lib.c8h_set_qoi_avg_disp(self.h))

    def set_qoi_calibration(self, *, balance_factor, coord_idx, coord_value, coord_tol=1e-12,
                            reaction_force_comp, weights, measured, load_data, area, facet=None):
        w = np.ones(3); w[: len(weights)] = weights
        measured = np.ascontiguousarray(measured, dtype=np.float64)
        load_data = np.ascontiguousarray(load_data, dtype=np.float64)
        fct = None if facet is None else np.ascontiguousarray(facet, dtype=np.int8)
        self._check(self.lib.c8h_set_qoi_calibration(
            self.h, C.c_double(balance_factor), coord_idx, C.c_double(coord_value),
            C.c_double(coord_tol), reaction_force_comp, _hp(w), _hp(measured), _hp(load_data),
            _hp(fct), C.c_double(area)))

    def set_qoi_mismatch(self, kind, *, coord_idx=0, coord_value=0.0, coord_tol=1e-12, reaction_force_comp=0,
                         compute_torque=False, facet=None, normal_2d=None, measured=None, load_data=None):
        """'reaction mismatch' / 'load mismatch' / 'surface mismatch' (src/qoi.cpp:272-287).  load_data None =
        the reference's "load out file" mode (mismatch against zero; loads() returns the file's lines)."""
        k = QOI_TYPES[kind.replace(" ", "_")]
        fct = None if facet is None else np.ascontiguousarray(facet, dtype=np.int8)
        n2 = None if normal_2d is None else np.ascontiguousarray(normal_2d, dtype=np.float64)
        m = None if measured is None else np.ascontiguousarray(measured, dtype=np.float64)
        ld = None if load_data is None else np.ascontiguousarray(load_data, dtype=np.float64)
        self._check(self.lib.c8h_set_qoi_mismatch(self.h, k, coord_idx, C.c_double(coord_value), C.c_double(coord_tol),
                                                  reaction_force_comp, int(compute_torque), _hp(fct), _hp(n2),
                                                  _hp(m), _hp(ld)))

    def loads(self):
        out = np.zeros(self.num_steps)
        self._check(self.lib.c8h_get_loads(self.h, _hp(out)))
        return out

    def primal_solve(self):
        J = C.c_double(0)
        self._check(self.lib.c8h_primal_solve(self.h, C.byref(J)))
        return J.value

    def adjoint_gradient(self):
        g = np.zeros(self.ctx.npar)
        self._check(self.lib.c8h_adjoint_gradient(self.h, _hp(g)))
        return g

    def get_step(self, step):
        c = self.ctx
        u = np.zeros(c.n_nodes * c.dim)
        p = np.zeros(c.n_nodes) if c.num_resid == 2 else None
        xi = np.zeros((c.n_elems, c.nxi))
        self._check(self.lib.c8h_get_step(self.h, step, _hp(u), _hp(p), _hp(xi)))
        return [u, p][: c.num_resid], xi

    def get_adjoint_step(self, step):
        c = self.ctx
        zu = np.zeros(c.n_nodes * c.dim)
        zp = np.zeros(c.n_nodes) if c.num_resid == 2 else None
        phi = np.zeros((c.n_elems, c.nxi))
        self._check(self.lib.c8h_get_adjoint_step(self.h, step, _hp(zu), _hp(zp), _hp(phi)))
        return [zu, zp][: c.num_resid], phi

    def profile(self, enable=True):
        """wall-clock seconds per phase (assembly, linear solves, K3, K4, K5/K6, line-search ops)"""
        out = (C.c_double * 8)()
        self.lib.c8h_profile(self.h, int(enable), out)
        names = ["assemble", "linear_solve", "adjoint_jacobian", "adjoint_local", "qoi_gradient", "line_search"]
        return {n: out[k] for k, n in enumerate(names)}

    def stats(self):
        a, b = C.c_int(0), C.c_int(0)
        self.lib.c8h_stats(self.h, C.byref(a), C.byref(b))
        return dict(assemblies=a.value, linear_iters=b.value)


# ======================================================================================
# partition / communication (one Context per GPU; calibr8_b200/partition.py builds the plan)
SYMBOLS += [
    "c8_set_partition", "c8_get_partition", "c8_set_comm", "c8_set_halo_plan", "c8_nccl_unique_id",
    "c8_nccl_init", "c8_set_comm_host", "c8_set_comm_rank", "c8_halo", "c8_halo_nb", "c8_allreduce", "c8_comm_stats",
    "c8_comm_release", "c8_comm_p2p_active",
]
_EXCHANGE_FN = C.CFUNCTYPE(None, C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_int)
_HOST_ALLREDUCE_FN = C.CFUNCTYPE(None, C.c_void_p, C.POINTER(C.c_double), C.c_int)


def _ctx_set_partition(self, part):
    """part: calibr8_b200.partition.Part whose local mesh was given to set_mesh."""
    assert part.n_nodes == self.n_nodes and part.n_elems == self.n_elems
    self.part = part
    self._check(self.lib.c8_set_partition(self.h, part.n_owned_nodes, part.n_owned_elems))
    i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)
    nbr, sp, sn, rp = i32(part.nbr_rank), i32(part.send_ptr), i32(part.send_nodes), i32(part.recv_ptr)
    self._check(self.lib.c8_set_halo_plan(self.h, int(nbr.size), _hp(nbr), _hp(sp), _hp(sn), _hp(rp)))


def _ctx_nccl_init(self, rank, nranks, broadcast_bytes):
    """broadcast_bytes(b: bytes | None) -> bytes : distributes rank 0's 128-byte id"""
    buf = C.create_string_buffer(128)
    if rank == 0:
        rc = self.lib.c8_nccl_unique_id(buf)
        if rc != 0:
            raise C8Error("c8_nccl_unique_id failed (libnccl.so.2 not loadable)")
    raw = broadcast_bytes(buf.raw if rank == 0 else None)
    self._check(self.lib.c8_nccl_init(self.h, C.create_string_buffer(raw, 128), rank, nranks))


def _ctx_set_comm_host(self, exchanger):
    """exchanger: object with .exchange(send[n_send][nb], nb) -> recv[n_ghost][nb] and
    .allreduce(buf) -> summed buf (e.g. partition.HostExchange)"""
    part = self.part
    n_send, n_recv = int(part.send_ptr[-1]), int(part.recv_ptr[-1])

    def ex(_user, send_p, recv_p, nb):
        send = np.ctypeslib.as_array(send_p, shape=(max(n_send, 1) * nb,))[: n_send * nb]
        recv = exchanger.exchange(send.reshape(n_send, nb), nb)
        if n_recv:
            np.ctypeslib.as_array(recv_p, shape=(n_recv * nb,))[:] = np.asarray(recv).reshape(-1)

    def ar(_user, buf_p, n):
        a = np.ctypeslib.as_array(buf_p, shape=(n,))
        a[:] = exchanger.allreduce(a.copy())

    self._cb = (_EXCHANGE_FN(ex), _HOST_ALLREDUCE_FN(ar))   # keep alive
    self._check(self.lib.c8_set_comm_host(self.h, self._cb[0], self._cb[1], None))
    if getattr(exchanger, "world", 1) > 1:   # lets the multigrid preconditioner span the parts
        self._check(self.lib.c8_set_comm_rank(self.h, int(exchanger.rank), int(exchanger.world)))


def _ctx_halo(self, vec, nb=None):
    if nb is None:
        self._check(self.lib.c8_halo(self.h, _dp(vec)))
    else:
        self._check(self.lib.c8_halo_nb(self.h, _dp(vec), nb))


def _ctx_allreduce(self, buf):
    self._check(self.lib.c8_allreduce(self.h, _dp(buf), int(buf.numel())))


def _ctx_comm_stats(self):
    out = (C.c_int64 * 3)()
    self.lib.c8_comm_stats(self.h, out)
    return dict(halo_calls=int(out[0]), allreduce_calls=int(out[1]), halo_bytes=int(out[2]))


def _ctx_p2p_active(self):
    v = int(self.lib.c8_comm_p2p_active(self.h))
    return dict(halo=bool(v & 1), allreduce=bool(v & 2))


Context.p2p_active = _ctx_p2p_active
Context.set_partition = _ctx_set_partition
Context.nccl_init = _ctx_nccl_init
Context.set_comm_host = _ctx_set_comm_host
Context.halo = _ctx_halo
Context.allreduce = _ctx_allreduce
Context.comm_stats = _ctx_comm_stats


SYMBOLS += ["c8_set_preconditioner", "c8_preconditioner_info", "c8_linalg_invalidate"]


def _ctx_set_preconditioner(self, kind="amg", nu_pre=2, nu_post=2, omega=0.8, over_correction=1.6,
                            coarsest_max_nodes=40, max_aggregate_size=8, coarse_aggregate_size=12, coarse_nu=0,
                            distributed=True, replicate_max_nodes=30000):
    """right preconditioner of gmres(): 'amg' (aggregation multigrid, default) or 'block_jacobi'.
    distributed: on a partitioned run the hierarchy spans the parts (False: each part's owned block);
    replicate_max_nodes: a level with at most this many nodes globally is replicated on every part"""
    opts = np.array([nu_pre, nu_post, omega, over_correction, coarsest_max_nodes, max_aggregate_size,
                     coarse_aggregate_size, coarse_nu, 1.0 if distributed else 0.0, replicate_max_nodes],
                    dtype=np.float64)
    self._check(self.lib.c8_set_preconditioner(self.h, {"block_jacobi": 0, "amg": 1}[kind], _hp(opts),
                                               int(opts.size)))


def _ctx_preconditioner_info(self):
    out = (C.c_double * 16)()
    self._check(self.lib.c8_preconditioner_info(self.h, out, 16))
    nl = int(out[0])
    return dict(levels=nl, operator_complexity=out[1], nodes=[int(out[2 + l]) for l in range(nl)])


Context.set_preconditioner = _ctx_set_preconditioner
Context.preconditioner_info = _ctx_preconditioner_info


def eval_expr(expr, coords, t=0.0):
    """Evaluate an expression in x, y, z, t (the grammar of calibr8_b200/host/expr.hpp, the role of the
    reference's Pamgen RTC strings) at the rows of coords [n, 3] -- host only, no GPU needed."""
    lib = load_library()
    xyz = np.ascontiguousarray(coords, dtype=np.float64).reshape(-1, 3)
    out = np.zeros(xyz.shape[0])
    err = C.create_string_buffer(256)
    rc = lib.c8h_eval_expr(str(expr).encode(), _hp(xyz), int(xyz.shape[0]), C.c_double(t), _hp(out), err, 256)
    if rc != 0:
        raise C8Error("expression: " + err.value.decode())
    return out


# ======================================================================================
# Objective on canonical [-1, 1] parameters (calibr8_b200/host/objective.cu)
HOST_SYMBOLS += ["c8h_objective_create", "c8h_objective_destroy", "c8h_objective_error",
                 "c8h_objective_value", "c8h_objective_gradient", "c8h_objective_transform",
                 "c8h_objective_active_params", "c8h_vfm_objective"]


class Objective:
    """value(p) / gradient(p) of a calibration objective on canonical parameters -- the role of the
    reference's ROL::Objective subclasses (Adjoint_Objective, FS_VFM_Objective, Adjoint_VFM_Objective)."""
    TYPES = {"adjoint": 0, "pdeco": 0, "fs_vfm": 1, "vfm": 1, "adjoint_vfm": 2}

    def __init__(self, host_problem, kind, active, lower, upper, *, measured=None, w=None, loads=None,
                 obj_scale_factor=1.0, thickness=1.0):
        self.hp, self.lib = host_problem, host_problem.lib
        self.n = len(active)
        act = np.ascontiguousarray(active, dtype=np.int32)
        lo = np.ascontiguousarray(lower, dtype=np.float64)
        hi = np.ascontiguousarray(upper, dtype=np.float64)
        c = lambda a: None if a is None else np.ascontiguousarray(a, dtype=np.float64)
        self._keep = (c(measured), c(w), c(loads))
        self.lib.c8h_objective_create.restype = C.c_void_p
        self.lib.c8h_objective_error.restype = C.c_char_p
        self.lib.c8h_objective_error.argtypes = [C.c_void_p]
        h = self.lib.c8h_objective_create(host_problem.h, self.TYPES[kind.lower()], self.n, _hp(act), _hp(lo),
                                          _hp(hi), _hp(self._keep[0]), _hp(self._keep[1]), _hp(self._keep[2]),
                                          C.c_double(obj_scale_factor), C.c_double(thickness))
        self.h = C.c_void_p(h)
        err = self.lib.c8h_objective_error(self.h).decode()
        if err:
            raise C8Error("objective: " + err)

    def _check(self, rc):
        if rc != 0:
            raise C8Error("objective: " + self.lib.c8h_objective_error(self.h).decode())

    def value(self, p):
        p = np.ascontiguousarray(p, dtype=np.float64)
        J = C.c_double(0)
        self._check(self.lib.c8h_objective_value(self.h, _hp(p), self.n, C.byref(J)))
        return J.value

    def gradient(self, p):
        p = np.ascontiguousarray(p, dtype=np.float64)
        g = np.zeros(self.n)
        self._check(self.lib.c8h_objective_gradient(self.h, _hp(p), self.n, _hp(g)))
        return g

    def to_canonical(self, phys):
        out = np.zeros(self.n)
        self._check(self.lib.c8h_objective_transform(self.h, _hp(np.ascontiguousarray(phys, dtype=np.float64)),
                                                     self.n, 1, _hp(out)))
        return out

    def to_physical(self, can):
        out = np.zeros(self.n)
        self._check(self.lib.c8h_objective_transform(self.h, _hp(np.ascontiguousarray(can, dtype=np.float64)),
                                                     self.n, 0, _hp(out)))
        return out

    def active_params(self):
        out = np.zeros(self.n)
        self._check(self.lib.c8h_objective_active_params(self.h, _hp(out), self.n))
        return out

    def close(self):
        if getattr(self, "h", None):
            self.lib.c8h_objective_destroy(self.h)
            self.h = None


def describe_residuals(local_type, global_type, ndims):
    """create_local_residual / create_global_residual metadata of the host layer (names, variable
    types, equation counts, parameter order) -- host only, no GPU needed."""
    import json
    lib = load_library()
    buf = C.create_string_buffer(2048)
    rc = lib.c8h_describe_residuals(local_type.encode(), global_type.encode(), int(ndims), buf, 2048)
    if rc != 0:
        raise C8Error(buf.value.decode())
    return json.loads(buf.value.decode())
