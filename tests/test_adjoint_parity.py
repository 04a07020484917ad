"""K3 / K4 / K5 / K6 / K7 / K8 entry-level parity: every reverse-sweep / QoI / virtual-fields call of the
C ABI against the CPU oracle's restatement of the same reference function, on the same seeded inputs,
entry by entry at 1e-10 relative (BASELINE.json north_star: residual / Jacobian entries 1e-10).

    c8_adjoint_jacobian  <-> eval_adjoint_jacobian      src/evaluations.cpp:349-526  (A^T per CSR row, rhs, g)
    c8_adjoint_local     <-> solve_adjoint_local        src/evaluations.cpp:528-659  (phi, f, g)
    c8_qoi_value         <-> eval_qoi / preprocess_qoi  src/evaluations.cpp:662-756, 261-347
    c8_qoi_gradient      <-> eval_qoi_gradient          src/evaluations.cpp:758-925  (every parameter)
    c8_vfm_forward       <-> eval_measured_residual_and_grad  :1847-1973 (b, dR, local_sens)
    c8_vfm_adjoint       <-> eval_vfm_adjoint_gradient  :1975-2143 (hist, grad)

The random adjoint histories g, f and the nodal adjoint z are O(1) fields, so every derivative block
of the kernels is exercised (the end-to-end gradient tests of test_solve_parity.py only see the
combination a converged sweep produces)."""
import numpy as np
import pytest

from parity_common import (COMBOS, LOCAL_TOL, make_context, make_mesh, rel_err_blockwise, rel_err_state,
                           rel_err_rows, synthetic_fields, xlist)

TOL = 1e-10
pytestmark = pytest.mark.gpu


# ---- layout helpers (reference layouts <-> the library's HBM layouts) -------------------------
def ref_dof(ctx, c):
    """element dof c = n*NB + eq of the library -> position in the reference's residual-major order
    (src/global_residual.cpp:21-23)"""
    n, eq = divmod(c, ctx.nb)
    return n * ctx.dim + eq if eq < ctx.dim else ctx.nn * ctx.dim + n


def pack_elem_dofs(ctx, f_ref):
    """[n_elems][nx] (reference dof order) -> device [nx][xi_ld] (node-interleaved dofs)"""
    import torch
    out = np.zeros((ctx.nx, ctx.xi_ld))
    for c in range(ctx.nx):
        out[c, : ctx.n_elems] = f_ref[:, ref_dof(ctx, c)]
    return torch.from_numpy(out).cuda().reshape(-1)


def unpack_elem_dofs(ctx, f_dev):
    a = f_dev.cpu().numpy().reshape(ctx.nx, ctx.xi_ld)
    out = np.zeros((ctx.n_elems, ctx.nx))
    for c in range(ctx.nx):
        out[:, ref_dof(ctx, c)] = a[c, : ctx.n_elems]
    return out


def xi_tensor(ctx, aos):
    t = ctx.alloc("xi")
    ctx.pack_xi(aos, t)
    return t


def x_tensor(ctx, parts):
    t = ctx.alloc("x")
    ctx.pack_x(parts[0], parts[1] if len(parts) > 1 else None, t)
    return t


def facets_on_plane(mesh, coord_idx, value, tol=1e-12):
    """[n_elems][3] local vertex ids of the tet face lying on a coordinate plane, -1 where none
    (the side-set facet of Calibration::compute_surface_mismatch, src/calibration.cpp:225-303)"""
    on = np.abs(mesh.coords[mesh.conn][:, :, coord_idx] - value) < tol      # [n_elems][4]
    fac = np.full((mesh.n_elems, 3), -1, dtype=np.int32)
    for e in np.nonzero(on.sum(axis=1) == 3)[0]:
        fac[e] = np.nonzero(on[e])[0]
    return fac


class Pair:
    """One synthetic two-step state evaluated by both sides (as test_forward_parity.run_pair)."""

    def __init__(self, name, size="small", all_active=True, params=None, mesh=None, amp=None):
        import torch
        from oracle.pyoracle import Oracle, PARAM_NAMES
        dim, gtype, ltype, par0, amp0 = COMBOS[name]
        params = params or par0
        self.mesh = mesh = mesh or make_mesh(dim, size)
        mixed = gtype == "mechanics"
        (u1, p1), (u2, p2) = synthetic_fields(mesh, amp or amp0, mixed)
        self.npar = len(PARAM_NAMES[ltype])
        self.orc = o = Oracle(mesh.dim, mesh.conn, mesh.coords, global_type=gtype, local_type=ltype,
                              params=[params], active=[list(range(self.npar))] if all_active else None,
                              **LOCAL_TOL)
        self.ctx = c = make_context(mesh, gtype, ltype, params)
        xi0 = o.init_xi()
        rA = o.forward_jacobian(xlist(u1, p1), o.zeros_x(), xi0, xi0, assemble=False)
        rB = o.forward_jacobian(xlist(u2, p2), xlist(u1, p1), rA["xi"], rA["xi"], assemble=False)
        assert rA["status"] == 0 and rB["status"] == 0
        self.plastic = float(rB["path"].mean())
        self.x, self.xp, self.xi, self.xip = xlist(u2, p2), xlist(u1, p1), rB["xi"], rA["xi"]
        self.dx, self.dxp = x_tensor(c, self.x), x_tensor(c, self.xp)
        self.dxi, self.dxip = xi_tensor(c, self.xi), xi_tensor(c, self.xip)
        rng = np.random.RandomState(3)
        self.g = rng.uniform(-1., 1., size=(o.n_elems, o.n_xi))
        self.f = rng.uniform(-1., 1., size=(o.n_elems, o.n_x))
        self.z = [rng.uniform(-1., 1., size=o.n_nodes * q) for q in o.neq]
        torch.cuda.synchronize()

    # QoI on both sides; returns the library's c8_qoi struct (load mismatch filled by the K5 pass)
    def set_qoi(self, kind):
        import torch
        from calibr8_b200.capi import make_qoi
        o, c, mesh = self.orc, self.ctx, self.mesh
        if kind == "avg_disp":
            o.set_qoi_avg_disp()
            return make_qoi("avg_disp")
        if kind != "calibration":
            return self._set_mismatch_qoi(kind)
        rng = np.random.RandomState(11)
        # "measured" displacement: the current one perturbed, so that the mismatch is non-zero
        u = self.x[0].reshape(-1, mesh.dim)
        meas = u * (1.0 + 0.05 * rng.uniform(-1., 1., size=u.shape))
        m3 = np.zeros((mesh.n_nodes, 3)); m3[:, : mesh.dim] = meas
        w = (1e8, 2e8, 3e8)[: mesh.dim]      # the shipped decks weigh the displacement mismatch by 1e8
        dt, T, load_meas, bf = 1.0, 4.0, 0.01, 1e2
        self.fac = None
        if mesh.dim == 3:
            self.fac = facets_on_plane(mesh, 2, 1.0)    # displacement mismatch on the zmax face
            assert (self.fac[:, 0] >= 0).sum() > 0
        o.set_qoi_calibration(balance_factor=bf, coord_idx=1, coord_value=1.0, reaction_force_comp=1,
                              weights=w, facet=self.fac)
        o.qoi_set_step(dt, T, load_meas, m3)
        self.J_orc = o.qoi(self.x, self.xp, self.xi, self.xip, 1)
        st = o.calibration_state()
        self._keep = (torch.from_numpy(np.ascontiguousarray(meas)).cuda(),
                      None if self.fac is None else torch.from_numpy(self.fac.astype(np.int8)).cuda())
        q = make_qoi("calibration", weights=w, balance_factor=bf, dt_over_T=dt / T,
                     inv_area=1.0 / st["area"], coord_idx=1, coord_value=1.0, coord_tol=1e-12,
                     reaction_force_comp=1, measured=self._keep[0], facet=self._keep[1])
        # K5, both modes: total load on the plane (preprocess_qoi), then the objective value
        sc = torch.zeros(8, dtype=torch.float64, device="cuda")
        torch.cuda.synchronize()
        c.qoi_value(q, self.dx, self.dxp, self.dxi, self.dxip, 1, sc)
        c.synchronize()
        total = float(sc[1].item())
        assert abs(total - st["total_load"]) <= TOL * abs(st["total_load"]), (total, st)
        q.load_mismatch = total - load_meas
        c.qoi_value(q, self.dx, self.dxp, self.dxi, self.dxip, 0, sc)
        c.synchronize()
        J_disp = float(sc[0].item())
        assert st["J_disp"] > 1e-3 * st["J_forc"] > 0     # both terms matter in this state
        assert abs(J_disp - st["J_disp"]) <= TOL * st["J_disp"], (J_disp, st)
        J = J_disp + 0.5 * bf * dt / T * q.load_mismatch ** 2
        assert abs(J - self.J_orc) <= TOL * abs(self.J_orc), (J, self.J_orc)
        return q


    def _set_mismatch_qoi(self, kind):
        """reaction mismatch (plane load / torque), load mismatch (normal load over the ymax facets),
        surface mismatch (zmax facets) -- src/reaction_mismatch.cpp, load_mismatch.cpp, surface_mismatch.cpp"""
        import torch
        from calibr8_b200.capi import make_qoi
        o, c, mesh = self.orc, self.ctx, self.mesh
        dim = mesh.dim
        rng = np.random.RandomState(12)
        u = self.x[0].reshape(-1, dim)
        meas = u * (1.0 + 0.05 * rng.uniform(-1., 1., size=u.shape))
        m3 = np.zeros((mesh.n_nodes, 3)); m3[:, :dim] = meas
        load_meas = 0.01
        fac = None
        if kind in ("load", "surface"):
            plane = 1 if kind == "load" else 2
            on = np.abs(mesh.coords[mesh.conn][:, :, plane] - 1.0) < 1e-12
            fac = np.full((mesh.n_elems, 3), -1, dtype=np.int32)
            for e in np.nonzero(on.sum(axis=1) == dim)[0]:
                fac[e, :dim] = np.nonzero(on[e])[0]
            assert (fac[:, 0] >= 0).sum() > 0
        torque = kind == "reaction_torque"
        comp = 2 if torque else 1
        okind = {"reaction": "reaction", "reaction_torque": "reaction", "load": "load", "surface": "surface"}[kind]
        o.set_qoi_mismatch(okind, coord_idx=1, coord_value=1.0, reaction_force_comp=comp, compute_torque=torque,
                           facet=fac, normal_2d=(0., 1.))
        o.qoi_set_step(1.0, 4.0, load_meas, m3)
        self.J_orc = o.qoi(self.x, self.xp, self.xi, self.xip, 1)
        st = o.calibration_state()
        self._keep = (torch.from_numpy(np.ascontiguousarray(meas)).cuda(),
                      None if fac is None else torch.from_numpy(fac.astype(np.int8)).cuda())
        ctype = {"reaction": "reaction_mismatch", "reaction_torque": "reaction_mismatch", "load": "load_mismatch",
                 "surface": "surface_mismatch"}[kind]
        if kind == "surface":     # the host's mapping of the surface integrand (host.cu make_qoi)
            q = make_qoi(ctype, weights=(2., 2., 2.), balance_factor=0.0, dt_over_T=1.0, inv_area=1.0,
                         measured=self._keep[0], facet=self._keep[1])
        else:
            q = make_qoi(ctype, balance_factor=1.0, dt_over_T=1.0, inv_area=1.0, coord_idx=1, coord_value=1.0,
                         coord_tol=1e-12, reaction_force_comp=comp, compute_torque=torque, facet=self._keep[1],
                         normal_2d=(0., 1.))
        sc = torch.zeros(8, dtype=torch.float64, device="cuda")
        torch.cuda.synchronize()
        if kind != "surface":
            c.qoi_value(q, self.dx, self.dxp, self.dxi, self.dxip, 1, sc)
            c.synchronize()
            total = float(sc[1].item())
            assert abs(st["total_load"]) > 0
            assert abs(total - st["total_load"]) <= TOL * abs(st["total_load"]), (total, st)
            q.load_mismatch = total - load_meas
        c.qoi_value(q, self.dx, self.dxp, self.dxi, self.dxip, 0, sc)
        c.synchronize()
        J = float(sc[0].item()) + (0.5 * q.load_mismatch ** 2 if kind != "surface" else 0.0)
        assert self.J_orc > 0 and abs(J - self.J_orc) <= TOL * abs(self.J_orc), (J, self.J_orc)
        return q


ADJ_CASES = [("3d_small_J2", "avg_disp"), ("3d_small_hill", "calibration"), ("3d_hyper_J2", "avg_disp"),
             ("3d_hyper_J2", "calibration"), ("3d_elastic", "avg_disp"),
             ("2d_small_hill_plane_stress", "calibration"), ("2d_hyper_J2_plane_stress", "avg_disp"),
             ("2d_small_J2", "avg_disp"), ("2d_hyper_J2_plane_strain", "calibration"),
             # reaction / load / surface mismatch QoIs
             ("3d_small_hill", "reaction"), ("3d_hyper_J2", "reaction_torque"), ("3d_hyper_J2", "load"),
             ("3d_small_J2", "load"), ("2d_hyper_J2_plane_stress", "load"), ("2d_small_hill_plane_strain", "load"),
             ("3d_small_J2", "surface"), ("2d_small_hill_plane_stress", "reaction"),
             # finite-strain Hill (AD through the polar-rotation iteration)
             ("3d_hypo_hill", "calibration"), ("2d_hypo_hill_plane_strain", "avg_disp"),
             ("2d_hypo_hill_plane_stress", "load")]


@pytest.mark.parametrize("name,qoi", ADJ_CASES)
def test_adjoint_entry_points(name, qoi):
    import torch
    P = Pair(name)
    o, c = P.orc, P.ctx
    assert 0 < P.plastic < 1 or name.endswith("elastic")
    q = P.set_qoi(qoi)
    nr = o.num_resid

    # ---- K3: A^T, rhs, g -= dJ/dxi -------------------------------------------------------------
    g_o = P.g.copy()
    AT_o, rhs_o = o.adjoint_jacobian(P.x, P.xp, P.xi, P.xip, g_o, P.f, 1)
    g_d, f_d = xi_tensor(c, P.g), pack_elem_dofs(c, P.f)
    AT, rhs = c.alloc("A"), c.alloc("b")
    c.adjoint_jacobian(q, P.dx, P.dxp, P.dxi, P.dxip, g_d, f_d, AT, rhs)
    c.synchronize()
    for i in range(nr):
        for j in range(nr):
            rp, _ = o.graph(i, j)
            vals = c.csr_block_values(i, j, AT)
            assert rel_err_rows(vals, AT_o[i * nr + j], rp) < TOL, (i, j)
    rhs_h = c.unpack_x(rhs)
    for i in range(nr):
        assert np.abs(rhs_h[i] - rhs_o[i]).max() < TOL * np.abs(rhs_o[i]).max(), i
    assert rel_err_blockwise(c.unpack_xi(g_d), g_o, 0) < TOL
    if qoi not in ("avg_disp", "surface"):
        assert np.abs(g_o - P.g).max() > 0 or o.n_xi == 1   # the load term did reach g

    # ---- K4: phi, f, g of the previous step -----------------------------------------------------
    g2_o, f2_o = g_o.copy(), P.f.copy()
    phi_o = o.adjoint_local(P.x, P.xp, P.xi, P.xip, P.z, g2_o, f2_o)
    z_d, phi_d = x_tensor(c, P.z), c.alloc("xi")
    c.adjoint_local(P.dx, P.dxp, P.dxi, P.dxip, z_d, phi_d, g_d, f_d)
    c.synchronize()
    assert rel_err_blockwise(c.unpack_xi(phi_d), phi_o, 0) < TOL
    g_scale = max(np.abs(g2_o).max(), 1e-300)
    assert np.abs(c.unpack_xi(g_d) - g2_o).max() < TOL * g_scale
    f_scale = np.abs(f2_o).max()
    f_h = unpack_elem_dofs(c, f_d)
    if f_scale > 0:     # non-zero only for the finite-strain models (dC/dx_prev)
        assert np.abs(f_h - f2_o).max() < TOL * f_scale
    else:
        assert np.abs(f_h).max() == 0.0

    # ---- K6: gradient w.r.t. every parameter of the model ---------------------------------------
    grad_o = o.qoi_gradient(P.x, P.xp, P.xi, P.xip, P.z, phi_o, [list(range(P.npar))], P.npar, 1)
    grad = torch.zeros(P.npar, dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    c.qoi_gradient(q, P.dx, P.dxp, P.dxi, P.dxip, z_d, phi_d, grad)
    c.synchronize()
    gh = grad.cpu().numpy()
    # the sum over the elements cancels, so the absolute error is bounded against the largest entry
    # (1e-10) and the leading entries also relative to themselves (1e-8, the north-star gradient bar)
    assert np.abs(gh - grad_o).max() < TOL * np.abs(grad_o).max(), (gh, grad_o)
    nz = np.abs(grad_o) > 1e-2 * np.abs(grad_o).max()
    assert (np.abs(gh - grad_o)[nz] < 1e-8 * np.abs(grad_o)[nz]).all(), (gh, grad_o)
    c.close()


@pytest.mark.parametrize("name", ["2d_small_hill_plane_stress", "2d_hyper_J2_plane_stress",
                                  "2d_hypo_hill_plane_stress"])
def test_vfm_entry_points(name):
    """K7 / K8 with the forward sensitivities carried over TWO steps (so dC/dxi_prev . dxi/dp_prev is
    exercised) and a random history h."""
    import torch
    from calibr8_b200.vfm import vfm_adjoint, vfm_forward
    P = Pair(name)
    o, c, npar = P.orc, P.ctx, P.npar
    n, nxi = o.n_elems, o.n_xi
    zero = o.zeros_x()
    xi0 = o.init_xi()
    # oracle: step 1 (from rest) then step 2, sensitivities w.r.t. every parameter
    ls_o = np.zeros((n, nxi, npar))
    st1, b1_o, dR1_o, xi1_o = o.measured_residual_grad(P.xp, zero, xi0, xi0, ls_o, npar)
    ls1_o = ls_o.copy()
    st2, b2_o, dR2_o, xi2_o = o.measured_residual_grad(P.x, P.xp, xi1_o, xi1_o, ls_o, npar)
    assert st1 == 0 and st2 == 0
    # CUDA
    dzero, dxi0 = c.alloc("x"), c.alloc("xi")
    c.init_xi(dxi0)
    xi1, xi2 = c.alloc("xi"), c.alloc("xi")
    c.init_xi(xi1)
    ls = torch.zeros(nxi * npar * c.xi_ld, dtype=torch.float64, device="cuda")
    b1, b2 = c.alloc("b"), c.alloc("b")
    dR1 = torch.zeros(npar * c.n_dofs, dtype=torch.float64, device="cuda")
    dR2 = torch.zeros_like(dR1)
    torch.cuda.synchronize()
    assert vfm_forward(c, P.dxp, dzero, dxi0, xi1, b1, dR1, ls) == 0
    c.synchronize()

    def ls_host():
        a = ls.cpu().numpy().reshape(nxi, npar, c.xi_ld)[:, :, :n]
        return np.transpose(a, (2, 0, 1))
    s1 = np.abs(ls1_o).max()
    assert np.abs(ls_host() - ls1_o).max() < TOL * s1
    xi2.copy_(xi1)
    assert vfm_forward(c, P.dx, P.dxp, xi1, xi2, b2, dR2, ls) == 0
    c.synchronize()
    assert rel_err_blockwise(c.unpack_xi(xi2), xi2_o, 0) < TOL
    for b_d, b_o in ((b1, b1_o), (b2, b2_o)):
        bh = c.unpack_x(b_d)[0]
        assert np.abs(bh - b_o[0]).max() < TOL * np.abs(b_o[0]).max()
    for dR_d, dR_o in ((dR1, dR1_o), (dR2, dR2_o)):
        dh = dR_d.cpu().numpy().reshape(npar, c.n_dofs)
        for p in range(npar):
            sc = np.abs(dR_o[0][p]).max()
            if sc > 0:
                assert np.abs(dh[p] - dR_o[0][p]).max() < TOL * sc, p
            else:
                assert np.abs(dh[p]).max() == 0.0, p
    assert np.abs(ls_host() - ls_o).max() < TOL * np.abs(ls_o).max()

    # K8 with a random history and virtual field
    rng = np.random.RandomState(5)
    hist = rng.uniform(-1., 1., size=(n, nxi))
    vf = [rng.uniform(-1., 1., size=o.n_nodes * o.neq[0])]
    s = 0.37
    hist_o = hist.copy()
    grad_o = o.vfm_adjoint_gradient(P.x, P.xp, xi2_o, xi1_o, vf, hist_o, s, npar)
    h_d, w_d = xi_tensor(c, hist), x_tensor(c, vf)
    grad = torch.zeros(npar, dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    vfm_adjoint(c, P.dx, P.dxp, xi2, xi1, w_d, s, h_d, grad)
    c.synchronize()
    assert np.abs(c.unpack_xi(h_d) - hist_o).max() < TOL * np.abs(hist_o).max()
    gh = grad.cpu().numpy()
    assert np.abs(gh - grad_o).max() < TOL * np.abs(grad_o).max(), (gh, grad_o)
    c.close()


def test_bench_parameters_fast_path():
    """bench.py's own hyper-J2 parameter set (S = D = A = n = 0: the closed-form return-map predictor of
    models.cuh and the skipped exp / pow terms) on a slice of bench.py's own workload mesh and state:
    local state, branch, element Jacobian and residual entry-wise against the oracle at 1e-10."""
    import torch
    import bench
    from oracle.pyoracle import Oracle
    from calibr8_b200.capi import Context
    mesh = bench.workload_mesh(10)          # same generator, geometry and state law as the 56-cell run
    (u1, p1), (u2, p2) = bench.workload_fields(mesh)
    o = Oracle(mesh.dim, mesh.conn, mesh.coords, global_type="mechanics", local_type="hyper_J2",
               params=[bench.PARAMS], **bench.LOCAL)
    xi0 = o.init_xi()
    rA = o.forward_jacobian([u1, p1], o.zeros_x(), xi0, xi0, assemble=False)
    rB = o.forward_jacobian([u2, p2], [u1, p1], rA["xi"], rA["xi"], element_out=True)
    assert rA["status"] == 0 and rB["status"] == 0
    assert 0.2 < rB["path"].mean() < 0.8
    assert rB["iters"][rB["path"] == 1].min() >= 2      # the reference iterates; the predictor does not
    c = Context(0)
    c.set_mesh(mesh.dim, mesh.conn, mesh.coords)
    c.set_model("mechanics", "hyper_J2", bench.PARAMS, **bench.LOCAL)
    c.set_stream(torch.cuda.current_stream().cuda_stream)
    x, xp = c.alloc("x"), c.alloc("x")
    c.pack_x(u2, p2, x); c.pack_x(u1, p1, xp)
    xi, xip = xi_tensor(c, rA["xi"]), xi_tensor(c, rA["xi"])
    A, b, path = c.alloc("A"), c.alloc("b"), c.alloc("path")
    eJ, eR = c.alloc("elem_J"), c.alloc("elem_R")
    assert c.forward_jacobian(x, xp, xip, xi, A, b, path, eJ, eR) == 0
    torch.cuda.synchronize()
    assert (path.cpu().numpy().astype(np.int32) == rB["path"]).all()
    assert rel_err_state(c.unpack_xi(xi), rB["xi"]) < TOL
    n, nx = c.n_elems, c.nx
    assert rel_err_blockwise(eJ.cpu().numpy().reshape(n, nx, nx), rB["elem_dtotal"], 0) < TOL
    assert rel_err_blockwise(eR.cpu().numpy().reshape(n, nx), rB["elem_R"], 0) < TOL
    # the production (FAST, branch-free) instantiation of the kernel gives the same matrix
    A2, b2 = c.alloc("A"), c.alloc("b")
    xi.copy_(xip)
    assert c.forward_jacobian(x, xp, xip, xi, A2, b2, None) == 0
    torch.cuda.synchronize()
    # (two instantiations of the same arithmetic -- one tile per CTA with direct stores vs the persistent kernel
    # with staged bulk stores: the compiler contracts a few multiply-adds differently, so rounding level, not bits)
    scale = A.abs().max().item()
    assert (A - A2).abs().max().item() < 1e-14 * scale
    # ... and the production path is bit-reproducible from call to call (deterministic gather order, dynamic
    # tile scheduling notwithstanding)
    A3, b3 = c.alloc("A"), c.alloc("b")
    xi.copy_(xip)
    assert c.forward_jacobian(x, xp, xip, xi, A3, b3, None) == 0
    torch.cuda.synchronize()
    assert torch.equal(A2, A3)
    for i in range(2):
        for j in range(2):
            rp, _ = o.graph(i, j)
            assert rel_err_rows(c.csr_block_values(i, j, A2), rB["A"][i * 2 + j], rp) < TOL
    c.close()


def test_local_newton_relative_tolerance():
    """a deck whose abs_tol lies below the residual floor of the converged state (hyper-J2:
    det(zeta + Ie I) - 1 ~ 2e-16) converges by rel_tol in the reference, because its R_norm_0 is the
    residual at ITS starting point (src/small_J2.cpp:147-151, src/hyper_J2.cpp:189-193).  The kernels
    start the Newton at the closed-form predictor for these two models and must still converge the same
    way (not run to max_iters and report C8_ERR_LOCAL_SOLVE).  abs_tol stays above the floor of the
    yield function itself: the same abs_tol selects the branch (|f| < abs_tol), and below ~1e-18 the
    reference's own Newton flips between the branches (status -1 from the oracle as well)."""
    import torch
    from oracle.pyoracle import Oracle
    from calibr8_b200.capi import Context
    for name in ("3d_small_J2", "3d_hyper_J2"):
        dim, gtype, ltype, par, amp = COMBOS[name]
        if ltype == "hyper_J2":
            par = dict(par, S=0., D=0., A=0., n=0.)     # linear hardening: the predictor path
        mesh = make_mesh(dim)
        (u1, p1), (u2, p2) = synthetic_fields(mesh, amp, True)
        tol = dict(max_iters=30, abs_tol=1e-17, rel_tol=1e-8)
        o = Oracle(mesh.dim, mesh.conn, mesh.coords, global_type=gtype, local_type=ltype, params=[par], **tol)
        xi0 = o.init_xi()
        rA = o.forward_jacobian(xlist(u1, p1), o.zeros_x(), xi0, xi0, assemble=False)
        rB = o.forward_jacobian(xlist(u2, p2), xlist(u1, p1), rA["xi"], rA["xi"], assemble=False)
        assert rA["status"] == 0 and rB["status"] == 0 and rB["path"].sum() > 0
        c = Context(0)
        c.set_mesh(mesh.dim, mesh.conn, mesh.coords)
        c.set_model(gtype, ltype, par, **tol)
        c.set_stream(torch.cuda.current_stream().cuda_stream)
        x, xp, x0 = c.alloc("x"), c.alloc("x"), c.alloc("x")
        c.pack_x(u1, p1, xp); c.pack_x(u2, p2, x)
        xi0_d, xi1, xi2 = c.alloc("xi"), c.alloc("xi"), c.alloc("xi")
        c.init_xi(xi0_d); c.init_xi(xi1)
        path = c.alloc("path")
        assert c.forward_jacobian(xp, x0, xi0_d, xi1, None, c.alloc("b"), path) == 0, name
        xi2.copy_(xi1)
        assert c.forward_jacobian(x, xp, xi1, xi2, None, c.alloc("b"), path) == 0, name
        torch.cuda.synchronize()
        assert (path.cpu().numpy().astype(np.int32) == rB["path"]).all()
        # both sides stop at rel_tol = 1e-8 of their first residual: the states agree to that level
        assert rel_err_blockwise(c.unpack_xi(xi2), rB["xi"], 0) < 1e-7
        c.close()


@pytest.mark.parametrize("name", ["3d_hyper_J2", "2d_small_hill_plane_stress", "2d_small_J2"])
def test_traction_bc_kernel(name):
    """c8_apply_tbc <-> apply_primal_tbc (src/tbcs.cpp:17-86): R[n, d] -= T_d N_n w dv over the one-point
    side quadrature, against the oracle's restatement (oracle/driver.py) on every boundary side of the
    mesh with a position- and time-dependent traction, entry-wise at 1e-10."""
    import torch
    from calibr8_b200 import capi
    from oracle.driver import Tbc, apply_primal_tbcs
    P = Pair(name)
    o, c, mesh = P.orc, P.ctx, P.mesh
    dim = mesh.dim
    # every boundary side: faces (3-D) / edges (2-D) that belong to exactly one element
    import itertools
    cnt = {}
    for e in range(mesh.n_elems):
        for f in itertools.combinations(sorted(int(v) for v in mesh.conn[e]), dim):
            cnt[f] = cnt.get(f, 0) + 1
    sides = np.array([f for f, k in cnt.items() if k == 1], dtype=np.int32)
    assert sides.shape[0] > 10
    exprs = ["0.3 * t + x * y", "sin(3.0 * x) - 0.5 * t * z", "0.1 + y * y"][:dim]
    t = 1.7
    R_o = o.zeros_b()
    apply_primal_tbcs(o, [Tbc(0, sides, exprs)], R_o, t)
    cent = mesh.coords[sides].mean(axis=1)
    trac = np.stack([capi.eval_expr(ex, cent, t) for ex in exprs], axis=1)
    R = c.alloc("b")
    sn = torch.from_numpy(sides).cuda()
    tr = torch.from_numpy(np.ascontiguousarray(trac)).cuda()
    torch.cuda.synchronize()
    c.apply_tbc(R, sn, tr)
    c.synchronize()
    Rh = c.unpack_x(R)
    assert np.abs(Rh[0] - R_o[0]).max() < TOL * np.abs(R_o[0]).max()
    if len(Rh) > 1:
        assert np.abs(Rh[1]).max() == 0.0     # the pressure rows see no traction
    c.close()
