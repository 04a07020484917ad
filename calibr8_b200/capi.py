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
LIB_PATH = os.path.join(_HERE, "lib", "libc8b200.so")

GLOBAL_TYPES = {"mechanics": 0, "mechanics_plane_stress": 1}
LOCAL_TYPES = {
    "elastic": 0, "small_J2": 1, "small_hill": 2, "small_hill_plane_stress": 3,
    "hyper_J2": 4, "hyper_J2_plane_stress": 5, "small_hill_plane_strain": 6,
    "hyper_J2_plane_strain": 7,
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
}

# every symbol include/c8b200.h declares (checked by tests/test_capi_symbols.py)
SYMBOLS = [
    "c8_create", "c8_destroy", "c8_last_error", "c8_version", "c8_set_mesh", "c8_set_model",
    "c8_set_params", "c8_info", "c8_bsr_pattern", "c8_bsr_pattern_dev", "c8_csr_block_size",
    "c8_csr_block_pattern", "c8_csr_block_values", "c8_pack_x", "c8_unpack_x", "c8_pack_xi",
    "c8_unpack_xi", "c8_init_xi", "c8_forward_jacobian", "c8_forward_jacobian_elem",
    "c8_global_residual", "c8_forward_jacobian_host", "c8_resident_matrix", "c8_set_stream",
    "c8_synchronize", "c8_state_set_prev", "c8_state_forward_jacobian", "c8_state_get_xi",
    "c8_state_ptrs", "c8_bench_dfma", "c8_bench_copy",
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
        return torch.zeros(n, dtype=dt, device=dev)

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
