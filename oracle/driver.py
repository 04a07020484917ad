"""ORACLE (test infrastructure only) -- step solvers above the oracle's evaluation loops.

Restates, with scipy's sparse direct solver in place of Belos/Teko/MueLu:
  Primal::solve_at_step     src/primal.cpp:31-208  (+ src/line_search.hpp:56-135)
  apply_expression_primal_dbcs  src/dbcs.cpp:28-121
  apply_primal_tbcs         src/tbcs.cpp:17-98
  Adjoint::solve_at_step    src/adjoint.cpp:76-189
  Adjoint_Objective value/gradient  src/adjoint_objective.cpp:22-118
  VirtualPower / VFM objectives     src/virtual_power.cpp:109-203,
                                    src/adjoint_sens_vfm_objective.cpp, forward_sens_vfm_objective.cpp
The linear solver only affects results through the Newton tolerance.
"""
from __future__ import annotations

import math

import numpy as np
import scipy.sparse.linalg as spla

_SAFE = {k: getattr(math, k) for k in ("sin", "cos", "tan", "atan", "exp", "sqrt", "pi", "pow",
                                       "fabs", "log")}
_SAFE["abs"] = abs


def eval_expr(expr, x, y, z, t):
    """String expression in x,y,z,t (the role of Pamgen RTC, src/control.cpp:104-120)."""
    if isinstance(expr, (int, float)):
        return float(expr)
    return float(eval(str(expr), {"__builtins__": {}}, dict(_SAFE, x=x, y=y, z=z, t=t)))


class Dbc:
    def __init__(self, resid, eq, nodes, expr):
        self.resid, self.eq, self.nodes, self.expr = resid, eq, np.asarray(nodes), expr


def apply_primal_dbcs(orc, dbcs, A, R, x, t, is_adjoint=False):
    """src/dbcs.cpp:56-119 on the per-block CSR value arrays A[i*nr+j] and vectors R[i]."""
    nr = orc.num_resid
    for bc in dbcs:
        i, eq = bc.resid, bc.eq
        neq = orc.neq[i]
        for node in bc.nodes:
            row = node * neq + eq
            sol = x[i][row]
            X = orc.coords[node]
            v = eval_expr(bc.expr, X[0], X[1], X[2], t)
            for j in range(nr):
                rowptr, colind = orc.graph(i, j)
                lo, hi = rowptr[row], rowptr[row + 1]
                vals = A[i * nr + j]
                if i == j:
                    cols = colind[lo:hi]
                    k = lo + int(np.searchsorted(cols, row))
                    diag = vals[k]
                    vals[lo:hi] = 0.0
                    vals[k] = diag
                    R[i][row] = 0.0 if is_adjoint else diag * (sol - v)
                else:
                    vals[lo:hi] = 0.0


class Tbc:
    """"bc name: [resid_idx, side_set_name, x-val, y-val(, z-val)]" (src/tbcs.cpp:28-35):
    sides = [n_sides][dim] node ids of each side of the set, exprs = one expression per dimension"""

    def __init__(self, resid, sides, exprs):
        self.resid, self.sides, self.exprs = resid, np.asarray(sides), list(exprs)


def apply_primal_tbcs(orc, tbcs, R, t):
    """src/tbcs.cpp:17-86: R[n,d] -= T_d(x_q, t) N_n(x_q) w dv over the side quadrature of the local
    variables' order (getIPFitShape(dim, 1): one point at the side centroid, where every linear side
    basis function is 1/dim and w dv = the side's area (3-D, w = 1/2, dv = |a x b|) or length (2-D,
    w = 2 on [-1, 1], dv = length / 2)."""
    dim = orc.dim
    for bc in tbcs:
        neq = orc.neq[bc.resid]
        for side in bc.sides:
            X = orc.coords[side]
            xq = X.mean(axis=0)
            if dim == 3:
                wdv = 0.5 * float(np.linalg.norm(np.cross(X[1] - X[0], X[2] - X[0])))
            else:
                wdv = float(np.linalg.norm(X[1] - X[0]))
            T = [eval_expr(e, xq[0], xq[1], xq[2], t) for e in bc.exprs[:dim]]
            for node in side:
                for d in range(dim):
                    R[bc.resid][node * neq + d] -= T[d] * (1.0 / dim) * wdv


def norm_b(R):
    return math.sqrt(sum(float(r @ r) for r in R))


def _split(orc, v):
    out, o = [], 0
    for i in range(orc.num_resid):
        n = orc.n_nodes * orc.neq[i]
        out.append(v[o:o + n].copy()); o += n
    return out


def _cubic_min(phi_0, dphi_0, a, phi, slope_a):  # src/line_search.hpp:56-67
    d1 = dphi_0 + slope_a - 3. * (phi_0 - phi) / (0. - a)
    rad = d1 * d1 - dphi_0 * slope_a
    if rad < 0.:
        return 0.5 * a
    d2 = math.sqrt(rad)
    denom = slope_a - dphi_0 + 2. * d2
    if denom == 0.:
        return 0.5 * a
    return a - a * (slope_a + d2 - d1) / denom


def line_search(phi_0, dphi_0, evalf, c1=1e-4, bmin=0.5, bmax=0.9, max_evals=4):
    armijo = c1 * dphi_0
    alpha, best_alpha, best_phi, any_ok = 1., 1., float("inf"), False
    for _ in range(max_evals):
        ok, phi, slope = evalf(alpha)
        if not ok:
            alpha *= 0.5
            continue
        any_ok = True
        if phi < best_phi:
            best_phi, best_alpha = phi, alpha
        if phi <= phi_0 + alpha * armijo:
            return alpha, True
        am = _cubic_min(phi_0, dphi_0, alpha, phi, slope)
        alpha = min(max(am, bmin * alpha), bmax * alpha)
    return best_alpha, any_ok


class Primal:
    """Forward load-step solver; keeps the all-steps history (x[step], xi[step])."""

    def __init__(self, orc, dbcs, num_steps, step_size=1.0, *, max_iters=15, abs_tol=1e-8,
                 rel_tol=1e-8, tbc=None, verbose=False):
        self.orc, self.dbcs = orc, dbcs
        self.num_steps, self.step_size = num_steps, step_size
        self.max_iters, self.abs_tol, self.rel_tol = max_iters, abs_tol, rel_tol
        self.verbose = verbose
        self.tbc = tbc  # optional callable(R, t) adding traction terms, or a list of Tbc
        self.reset()

    def reset(self):
        self.x = [self.orc.zeros_x()]
        self.xi = [self.orc.init_xi()]
        self.n_jac_evals = 0

    def time(self, step):
        return step * self.step_size

    def _assemble(self, step, x, xi_start):
        o = self.orc
        t = self.time(step)
        o.set_time(t, self.step_size)
        r = o.forward_jacobian(x, self.x[step - 1], xi_start, self.xi[step - 1])
        self.n_jac_evals += 1
        if r["status"] != 0:
            return None
        A, R = r["A"], r["b"]
        if self.tbc:
            if callable(self.tbc):
                self.tbc(R, t)
            else:
                apply_primal_tbcs(o, self.tbc, R, t)
        apply_primal_dbcs(o, self.dbcs, A, R, x, t)
        return A, R, r["xi"]

    def solve_at_step(self, step):
        o = self.orc
        x = [v.copy() for v in self.x[step - 1]]
        xi = self.xi[step - 1].copy()  # create_primal copies step-1 (src/disc.cpp:643-683)
        it, converged, r0 = 1, False, 1.0
        while it <= self.max_iters and not converged:
            res = self._assemble(step, x, xi)
            if res is None:
                raise RuntimeError(f"primal step {step}: local solve failed at the base point")
            A, R, xi = res
            rn = norm_b(R)
            if it == 1:
                r0 = rn
            rel = rn / r0 if r0 != 0 else float("nan")
            if self.verbose:
                print(f"  step {step} it {it} |R|={rn:.3e} rel={rel:.3e}")
            if rn < self.abs_tol or rel < self.rel_tol:
                converged = True
                break
            K = o.bmat(A).tocsc()
            dx = _split(o, spla.spsolve(K, -np.concatenate(R)))
            x = [a + b for a, b in zip(x, dx)]
            psi_0 = 0.5 * rn * rn
            dpsi_0 = -2. * psi_0
            saved_xi = xi.copy()
            state = dict(alpha_applied=1.0, x=x, xi=xi)

            def evalf(alpha):
                xt = [a + (alpha - state["alpha_applied"]) * b for a, b in zip(state["x"], dx)]
                state["x"], state["alpha_applied"] = xt, alpha
                res2 = self._assemble(step, xt, saved_xi)
                if res2 is None:
                    return False, 0., 0.
                A2, R2, xi2 = res2
                state["xi"] = xi2
                ra = norm_b(R2)
                Adx = o.bmat(A2) @ np.concatenate(dx)
                return True, 0.5 * ra * ra, float(np.concatenate(R2) @ Adx)

            alpha, ok = line_search(psi_0, dpsi_0, evalf)
            if not ok:
                raise RuntimeError(f"primal step {step}: line search could not assemble")
            x = [a + (alpha - state["alpha_applied"]) * b for a, b in zip(state["x"], dx)]
            xi = state["xi"]
            it += 1
        if not converged:
            raise RuntimeError(f"Newton's method failed in {self.max_iters} iterations")
        if len(self.x) > step:
            self.x[step], self.xi[step] = x, xi
        else:
            self.x.append(x); self.xi.append(xi)

    def solve(self, qoi_step_setup=None):
        """All steps; returns summed QoI J (src/main_primal.cpp:221-243)."""
        J = 0.
        self.J_steps = []
        for step in range(1, self.num_steps + 1):
            self.solve_at_step(step)
            if qoi_step_setup:
                qoi_step_setup(step)
            self.orc.set_time(self.time(step), self.step_size)
            Js = self.orc.qoi(self.x[step], self.x[step - 1], self.xi[step], self.xi[step - 1], step)
            self.J_steps.append(Js)
            J += Js
        return J


class Adjoint:
    """Reverse-in-time adjoint sweep + parameter gradient (src/adjoint.cpp, adjoint_objective.cpp:48-118)."""

    def __init__(self, primal, *, max_iters=15, abs_tol=1e-8, rel_tol=1e-8):
        self.p = primal
        self.max_iters, self.abs_tol, self.rel_tol = max_iters, abs_tol, rel_tol

    def gradient(self, grad_indices, n_grad, qoi_step_setup=None):
        p, o = self.p, self.p.orc
        N = p.num_steps
        g = np.zeros((o.n_elems, o.n_xi))
        f = np.zeros((o.n_elems, o.n_x))
        grad = np.zeros(n_grad)
        self.z, self.phi = {}, {}
        for step in range(N, 0, -1):
            if qoi_step_setup:
                qoi_step_setup(step)
            o.set_time(p.time(step), p.step_size)
            x, xp, xi, xip = p.x[step], p.x[step - 1], p.xi[step], p.xi[step - 1]
            AT, rhs = o.adjoint_jacobian(x, xp, xi, xip, g, f, step)
            z = o.zeros_x()
            apply_primal_dbcs(o, p.dbcs, AT, rhs, z, 0., is_adjoint=True)
            K = o.bmat(AT).tocsc()
            lu = spla.splu(K)
            it, r0 = 1, 1.0
            rhs_v = np.concatenate(rhs)
            while True:  # iterative refinement loop, src/adjoint.cpp:113-180
                dx = lu.solve(rhs_v)
                z = [a + b for a, b in zip(z, _split(o, dx))]
                rhs_v = rhs_v - K @ dx
                rn = float(np.linalg.norm(rhs_v))
                if it == 1:
                    r0 = rn
                if rn < self.abs_tol or (r0 > 0 and rn / r0 < self.rel_tol):
                    break
                it += 1
                if it > self.max_iters:
                    raise RuntimeError("adjoint refinement failed")
            phi = o.adjoint_local(x, xp, xi, xip, z, g, f)
            grad += o.qoi_gradient(x, xp, xi, xip, z, phi, grad_indices, n_grad, step)
            self.z[step], self.phi[step] = z, phi
        return grad
