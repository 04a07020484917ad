"""K1 / K2 parity: CUDA eval_forward_jacobian / eval_global_residual through the C ABI
vs the CPU oracle on the same seeded inputs (residual/Jacobian entries to 1e-10 relative,
CSR pattern bit-exact).  North-star tolerance: 1e-10 relative on entries (BASELINE.json)."""
import numpy as np
import pytest

from parity_common import (COMBOS, make_context, make_mesh, make_oracle, rel_err_blockwise, rel_err_state,
                           rel_err_rows, synthetic_fields, xlist)

TOL = 1e-10  # BASELINE.json north_star: residual/Jacobian entries within 1e-10 relative

pytestmark = pytest.mark.gpu


def rotate_fields(mesh, fields, angles):
    """superimpose a finite rigid rotation about z (one angle per state): u -> (Q - I) X + Q u"""
    out = []
    for (u, p), th in zip(fields, angles):
        c, s_ = np.cos(th), np.sin(th)
        Q = np.eye(mesh.dim)
        Q[0, 0], Q[0, 1], Q[1, 0], Q[1, 1] = c, -s_, s_, c
        X = mesh.coords[:, :mesh.dim]
        ur = X @ (Q - np.eye(mesh.dim)).T + u.reshape(-1, mesh.dim) @ Q.T
        out.append((ur.reshape(-1).copy(), p))
    return out


def run_pair(name, size="small", rotation=None):
    import torch
    dim, gtype, ltype, params, amp = COMBOS[name]
    mesh = make_mesh(dim, size)
    mixed = gtype == "mechanics"
    (u1, p1), (u2, p2) = synthetic_fields(mesh, amp, mixed)
    x_start = None
    if rotation is not None:
        # the history starts from a stress-free rigidly rotated configuration (rotation[0]), then the two states
        (u0, p0), (u1, p1), (u2, p2) = rotate_fields(mesh, [(0.0 * u1, None if p1 is None else 0.0 * p1),
                                                            (u1, p1), (u2, p2)], rotation)
        x_start = xlist(u0, p0)
    orc = make_oracle(mesh, gtype, ltype, params)
    ctx = make_context(mesh, gtype, ltype, params)
    zero = orc.zeros_x() if x_start is None else x_start
    xi0 = orc.init_xi()
    # step A (oracle): a first state so that step B starts from a non-trivial history
    rA = orc.forward_jacobian(xlist(u1, p1), zero, xi0, xi0, assemble=False)
    assert rA["status"] == 0
    xi1 = rA["xi"]
    # step B, oracle
    rB = orc.forward_jacobian(xlist(u2, p2), xlist(u1, p1), xi1, xi1, element_out=True)
    assert rB["status"] == 0
    # step B, CUDA
    x = ctx.alloc("x"); xp = ctx.alloc("x"); xi = ctx.alloc("xi"); xip = ctx.alloc("xi")
    A = ctx.alloc("A"); b = ctx.alloc("b"); path = ctx.alloc("path")
    eJ = ctx.alloc("elem_J"); eR = ctx.alloc("elem_R")
    ctx.pack_x(u2, p2, x); ctx.pack_x(u1, p1, xp)
    ctx.pack_xi(xi1, xi); ctx.pack_xi(xi1, xip)
    nf = ctx.forward_jacobian(x, xp, xip, xi, A, b, path, eJ, eR)
    torch.cuda.synchronize()
    return dict(mesh=mesh, orc=orc, ctx=ctx, rB=rB, nf=nf, x=x, xp=xp, xi=xi, xip=xip, A=A, b=b,
                path=path, eJ=eJ, eR=eR, xi1=xi1, fields=((u1, p1), (u2, p2)))


@pytest.mark.parametrize("name", list(COMBOS))
def test_forward_jacobian_parity(name):
    r = run_pair(name)
    orc, ctx, rB = r["orc"], r["ctx"], r["rB"]
    assert r["nf"] == 0
    # branch per element is identical, and the state exercises both branches for plastic models
    path = r["path"].cpu().numpy().astype(np.int32)
    assert (path == rB["path"]).all()
    if COMBOS[name][2] != "elastic":
        assert 0 < path.sum() < path.size, "synthetic state should mix elastic and plastic points"
    # local state
    xi = ctx.unpack_xi(r["xi"])
    assert rel_err_state(xi, rB["xi"]) < TOL
    # element Jacobians / residuals in the reference's dof order
    n, nx = ctx.n_elems, ctx.nx
    eJ = r["eJ"].cpu().numpy().reshape(n, nx, nx)
    eR = r["eR"].cpu().numpy().reshape(n, nx)
    assert rel_err_blockwise(eJ, rB["elem_dtotal"], 0) < TOL
    assert rel_err_blockwise(eR, rB["elem_R"], 0) < TOL
    # assembled system: pattern bit-exact, values to TOL per row
    nr = orc.num_resid
    for i in range(nr):
        for j in range(nr):
            rp_o, ci_o = orc.graph(i, j)
            rp_c, ci_c = ctx.csr_block_pattern(i, j)
            assert np.array_equal(rp_o, rp_c) and np.array_equal(ci_o, ci_c)
            vals = ctx.csr_block_values(i, j, r["A"])
            assert rel_err_rows(vals, rB["A"][i * nr + j], rp_o) < TOL
    bs = ctx.unpack_x(r["b"])
    for i in range(nr):
        scale = np.abs(rB["b"][i]).max()
        assert np.abs(bs[i] - rB["b"][i]).max() < TOL * scale
    ctx.close()


@pytest.mark.parametrize("name", ["3d_small_J2", "3d_hyper_J2", "2d_small_hill_plane_stress",
                                  "2d_hyper_J2_plane_stress"])
def test_global_residual_parity(name):
    import torch
    r = run_pair(name)
    orc, ctx, rB = r["orc"], r["ctx"], r["rB"]
    (u1, p1), (u2, p2) = r["fields"]
    b_o = orc.global_residual(xlist(u2, p2), xlist(u1, p1), rB["xi"], r["xi1"])
    b = ctx.alloc("b")
    ctx.global_residual(r["x"], r["xp"], r["xi"], r["xip"], b)
    torch.cuda.synchronize()
    bs = ctx.unpack_x(b)
    for i in range(orc.num_resid):
        assert np.abs(bs[i] - b_o[i]).max() < TOL * np.abs(b_o[i]).max()
    # and the residual-only kernel agrees with the Jacobian kernel's residual
    bj = ctx.unpack_x(r["b"])
    for i in range(orc.num_resid):
        assert np.abs(bs[i] - bj[i]).max() < TOL * np.abs(bj[i]).max()
    ctx.close()


@pytest.mark.parametrize("name", ["3d_hypo_hill", "2d_hypo_hill_plane_strain", "2d_hypo_hill_plane_stress",
                                  "3d_hyper_J2"])
def test_finite_rotation(name):
    """The finite-strain models under a finite rigid rotation superimposed on the states (0.29 rad stress-free
    start, then 0.30 and 0.31 rad about z): the hypo models then run minitensor::polar_rotation far from the
    identity (scaled Newton steps, more iterations), in the AD scalar -- the device iteration must follow the
    oracle's step for step."""
    r = run_pair(name, rotation=(0.29, 0.30, 0.31))
    rB, ctx = r["rB"], r["ctx"]
    assert r["nf"] == 0
    path = r["path"].cpu().numpy().astype(np.int32)
    assert (path == rB["path"]).all()
    n, nx = ctx.n_elems, ctx.nx
    eJ = r["eJ"].cpu().numpy().reshape(n, nx, nx)
    eR = r["eR"].cpu().numpy().reshape(n, nx)
    assert rel_err_blockwise(eJ, rB["elem_dtotal"], 0) < TOL
    assert rel_err_blockwise(eR, rB["elem_R"], 0) < TOL
    assert rel_err_state(ctx.unpack_xi(r["xi"]), rB["xi"]) < TOL
    ctx.close()


@pytest.mark.parametrize("name", ["3d_small_hill", "3d_hyper_J2", "2d_small_hill_plane_stress"])
def test_forward_jacobian_reference_mesh(name):
    """Same parity on the reference's own regression meshes (notch / notch2D)."""
    r = run_pair(name, size="ref")
    rB, ctx = r["rB"], r["ctx"]
    assert r["nf"] == 0
    n, nx = ctx.n_elems, ctx.nx
    eJ = r["eJ"].cpu().numpy().reshape(n, nx, nx)
    assert rel_err_blockwise(eJ, rB["elem_dtotal"], 0) < TOL
    assert rel_err_state(ctx.unpack_xi(r["xi"]), rB["xi"]) < TOL
    ctx.close()


def test_host_buffer_entry_point():
    """c8_forward_jacobian_host (reference layouts, copies inside) == device-pointer path."""
    name = "3d_small_J2"
    r = run_pair(name)
    ctx, rB = r["ctx"], r["rB"]
    (u1, p1), (u2, p2) = r["fields"]
    nf, xi, bs = ctx.forward_jacobian_host(u2, p2, u1, p1, r["xi1"], r["xi1"])
    assert nf == 0
    assert rel_err_state(xi, rB["xi"]) < TOL
    for i in range(2):
        assert np.abs(bs[i] - rB["b"][i]).max() < TOL * np.abs(rB["b"][i]).max()
    ctx.close()


@pytest.mark.parametrize("name", ["3d_hyper_J2", "2d_small_hill_plane_stress"])
def test_resident_state_host_call(name):
    """c8_state_forward_jacobian (bench.py's e2e call: host nodal iterate in, host residual + status
    out, matrix resident; the copies back overlap the BSR gather) against the oracle, twice in a row."""
    r = run_pair(name)
    orc, ctx, rB = r["orc"], r["ctx"], r["rB"]
    (u1, p1), (u2, p2) = r["fields"]
    ctx.state_set_prev(u1, p1, r["xi1"])
    nr = orc.num_resid
    for rep in range(2):
        bu = np.zeros(ctx.n_nodes * ctx.dim)
        bp = np.zeros(ctx.n_nodes) if nr > 1 else None
        nf = ctx.state_forward_jacobian(np.ascontiguousarray(u2), None if p2 is None else np.ascontiguousarray(p2),
                                        bu, bp)
        assert nf == 0
        bs = [bu] if nr == 1 else [bu, bp]
        for i in range(nr):
            assert np.abs(bs[i] - rB["b"][i]).max() < TOL * np.abs(rB["b"][i]).max()
        assert rel_err_blockwise(ctx.state_get_xi(), rB["xi"], 0) < TOL
        # the resident matrix is complete when the call returns
        A = int(ctx.state_ptrs()["A"])   # device pointer of the resident BSR values
        for i in range(nr):
            for j in range(nr):
                rp_o, _ = orc.graph(i, j)
                vals = ctx.csr_block_values(i, j, A)
                assert rel_err_rows(vals, rB["A"][i * nr + j], rp_o) < TOL
    ctx.close()


def test_local_solve_failure_is_reported():
    """max_iters too small for a plastic step -> status -1 like the reference (evaluations.cpp:95-97)."""
    import torch
    from calibr8_b200.capi import Context
    # a model whose local Newton starts at the reference's own point (no return-map predictor): several iterations
    dim, gtype, ltype, params, amp = COMBOS["2d_hyper_J2_plane_strain"]
    mesh = make_mesh(dim)
    (u1, p1), (u2, p2) = synthetic_fields(mesh, amp, True)
    ctx = Context(0)
    ctx.set_mesh(mesh.dim, mesh.conn, mesh.coords)
    ctx.set_model(gtype, ltype, params, max_iters=1, abs_tol=1e-12, rel_tol=1e-12)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    x = ctx.alloc("x"); xp = ctx.alloc("x"); xi = ctx.alloc("xi"); xip = ctx.alloc("xi")
    ctx.pack_x(u2, p2, x)
    ctx.init_xi(xi); ctx.init_xi(xip)
    nf = ctx.forward_jacobian(x, xp, xip, xi, None, ctx.alloc("b"))
    assert nf > 0
    ctx.close()


def test_unaligned_device_arrays_give_the_same_matrix():
    """The persistent kernel prefetches with 16-byte asynchronous copies when the caller's arrays allow it; arrays
    that are only 8-byte aligned (an offset pointer into a larger buffer) must take the 8-byte path and give
    the same bits."""
    import torch
    r = run_pair("3d_hyper_J2")
    ctx = r["ctx"]
    def shifted(t):
        big = torch.zeros(t.numel() + 1, dtype=t.dtype, device=t.device)
        v = big[1:]
        v.copy_(t)
        assert v.data_ptr() % 16 == 8
        return v
    x, xp, xip = shifted(r["x"]), shifted(r["xp"]), shifted(r["xip"])
    xi0 = r["xip"].clone()
    A1, b1 = ctx.alloc("A"), ctx.alloc("b")
    assert ctx.forward_jacobian(r["x"], r["xp"], r["xip"], xi0, A1, b1, None) == 0
    xi2 = shifted(r["xip"])
    A2, b2 = ctx.alloc("A"), ctx.alloc("b")
    assert ctx.forward_jacobian(x, xp, xip, xi2, A2, b2, None) == 0
    torch.cuda.synchronize()
    assert torch.equal(A1, A2) and torch.equal(xi0, xi2)
    assert (b1 - b2).abs().max().item() <= 1e-13 * b1.abs().max().item()   # red.add order is not fixed
    ctx.close()


def test_failed_local_solves_leave_no_stale_matrix_entries():
    """Production path (persistent element kernel, staged bulk stores) with a matrix: points whose local
    Newton fails contribute NOTHING -- their scratch slots are cleared, so the gather cannot sum an earlier
    call's element matrices into A -- and the count is the same as on the element-output path."""
    import torch
    from calibr8_b200.capi import Context
    dim, gtype, ltype, params, amp = COMBOS["3d_hyper_J2"]     # general hardening: the predictor is only a guess
    mesh = make_mesh(dim)
    (u1, p1), (u2, p2) = synthetic_fields(mesh, amp, True)
    ctx = Context(0)
    ctx.set_mesh(mesh.dim, mesh.conn, mesh.coords)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    # a converged assembly first fills every scratch slot with real element matrices
    ctx.set_model(gtype, ltype, params, max_iters=60, abs_tol=1e-14, rel_tol=1e-14)
    x = ctx.alloc("x"); xp = ctx.alloc("x"); xi = ctx.alloc("xi"); xip = ctx.alloc("xi")
    ctx.pack_x(u2, p2, x)
    A, A0, b, path = ctx.alloc("A"), ctx.alloc("A"), ctx.alloc("b"), ctx.alloc("path")
    ctx.init_xi(xi); ctx.init_xi(xip)
    assert ctx.forward_jacobian(x, xp, xip, xi, A0, b, path) == 0
    # now a tolerance no iterate can meet: every yielding point fails, the elastic ones still converge
    ctx.set_model(gtype, ltype, params, max_iters=2, abs_tol=1e-300, rel_tol=1e-300)
    ctx.init_xi(xi); ctx.init_xi(xip); b.zero_()
    nf = ctx.forward_jacobian(x, xp, xip, xi, A, b, path)
    torch.cuda.synchronize()
    failed = path.cpu().numpy() == -1
    assert nf == int(failed.sum()) and 0 < nf
    eJ = ctx.alloc("elem_J")
    ctx.init_xi(xi)
    nf2 = ctx.forward_jacobian(x, xp, xip, xi, None, None, path, eJ, None)
    assert nf2 == nf
    # the assembled matrix equals the sum of the converged elements' matrices only: rebuild it on the host
    n, nx = ctx.n_elems, ctx.nx
    eJh = eJ.cpu().numpy().reshape(n, nx, nx)
    eJh[failed] = 0.0
    ok = ~failed
    assert np.abs(eJh[ok]).max() > 0
    # compare through the action on a vector: A v == sum_e P_e^T J_e P_e v (reference dof order of elem_J)
    v = np.random.RandomState(3).randn(mesh.n_nodes, ctx.nb)
    vd = ctx.alloc("x"); vd.copy_(torch.from_numpy(v.reshape(-1)).cuda()); y = ctx.alloc("x")
    ctx.spmv(A, vd, y); torch.cuda.synchronize()
    yh = y.cpu().numpy().reshape(mesh.n_nodes, ctx.nb)
    nn = mesh.conn.shape[1]
    ref_dof = [(a, q) for a in range(nn) for q in range(dim)] + [(a, dim) for a in range(nn)]
    yref = np.zeros_like(yh)
    for e in np.nonzero(ok)[0]:
        ve = np.array([v[mesh.conn[e, a], q] for a, q in ref_dof])
        re = eJh[e] @ ve
        for k, (a, q) in enumerate(ref_dof):
            yref[mesh.conn[e, a], q] += re[k]
    assert np.abs(yh - yref).max() < 1e-10 * np.abs(yref).max()
    ctx.close()


def test_two_element_sets_with_different_materials():
    """Two element sets with their own material parameters (`materials: {es: {...}}` of the reference
    decks, LocalResidual::init_params per element set): K1 parity against the oracle."""
    import torch
    from calibr8_b200.capi import Context
    from oracle.pyoracle import Oracle
    dim, gtype, ltype, params, amp = COMBOS["3d_small_J2"]
    mesh = make_mesh(dim)
    cx = mesh.coords[mesh.conn].mean(axis=1)[:, 0]
    es = (cx > 0.5).astype(np.int32)
    assert 0 < es.sum() < es.size
    pars = [dict(params), dict(params, Y=3.0, K=40.0, E=800.0)]
    (u1, p1), (u2, p2) = synthetic_fields(mesh, amp, True)
    orc = Oracle(mesh.dim, mesh.conn, mesh.coords, es, 2, global_type=gtype, local_type=ltype,
                 params=pars, max_iters=60, abs_tol=1e-12, rel_tol=1e-12)
    xi0 = orc.init_xi()
    rA = orc.forward_jacobian(xlist(u1, p1), orc.zeros_x(), xi0, xi0, assemble=False)
    rB = orc.forward_jacobian(xlist(u2, p2), xlist(u1, p1), rA["xi"], rA["xi"], element_out=True)
    ctx = Context(0)
    ctx.set_mesh(mesh.dim, mesh.conn, mesh.coords, es, 2)
    ctx.set_model(gtype, ltype, pars, max_iters=60, abs_tol=1e-12, rel_tol=1e-12)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    x, xp, xi, xip = ctx.alloc("x"), ctx.alloc("x"), ctx.alloc("xi"), ctx.alloc("xi")
    A, b, path, eJ = ctx.alloc("A"), ctx.alloc("b"), ctx.alloc("path"), ctx.alloc("elem_J")
    ctx.pack_x(u2, p2, x); ctx.pack_x(u1, p1, xp)
    ctx.pack_xi(rA["xi"], xi); ctx.pack_xi(rA["xi"], xip)
    assert ctx.forward_jacobian(x, xp, xip, xi, A, b, path, eJ, None) == 0
    torch.cuda.synchronize()
    assert (path.cpu().numpy().astype(np.int32) == rB["path"]).all()
    n, nx = ctx.n_elems, ctx.nx
    assert rel_err_blockwise(eJ.cpu().numpy().reshape(n, nx, nx), rB["elem_dtotal"], 0) < TOL
    assert rel_err_state(ctx.unpack_xi(xi), rB["xi"]) < TOL
    # the two sets really behave differently
    assert rB["path"][es == 0].mean() != rB["path"][es == 1].mean()
    ctx.close()


@pytest.mark.parametrize("n_elems", [1, 3, 9])
def test_tiny_and_ragged_meshes(n_elems):
    """Edge cases of the launch geometry: fewer elements than one warp's worth of thread groups,
    element counts that are not a multiple of the groups per CTA."""
    import torch
    from calibr8_b200.capi import Context
    from oracle.pyoracle import Oracle
    dim, gtype, ltype, params, amp = COMBOS["3d_hyper_J2"]
    full = make_mesh(dim)
    conn = full.conn[:n_elems]
    used = np.unique(conn)
    remap = -np.ones(full.n_nodes, dtype=np.int64); remap[used] = np.arange(used.size)
    conn = remap[conn].astype(np.int32)
    coords = np.ascontiguousarray(full.coords[used])
    rng = np.random.RandomState(3)
    u1 = rng.uniform(-1, 1, size=coords.shape[0] * 3) * amp
    u2 = u1 * 1.3 + rng.uniform(-1, 1, size=u1.size) * 0.2 * amp
    p1 = rng.uniform(-1, 1, size=coords.shape[0]); p2 = rng.uniform(-1, 1, size=coords.shape[0])
    orc = Oracle(3, conn, coords, global_type=gtype, local_type=ltype, params=[params], max_iters=60,
                 abs_tol=1e-12, rel_tol=1e-12)
    xi0 = orc.init_xi()
    rA = orc.forward_jacobian([u1, p1], orc.zeros_x(), xi0, xi0, assemble=False)
    rB = orc.forward_jacobian([u2, p2], [u1, p1], rA["xi"], rA["xi"])
    ctx = Context(0)
    ctx.set_mesh(3, conn, coords)
    ctx.set_model(gtype, ltype, params, max_iters=60, abs_tol=1e-12, rel_tol=1e-12)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    x, xp, xi, xip = ctx.alloc("x"), ctx.alloc("x"), ctx.alloc("xi"), ctx.alloc("xi")
    A, b = ctx.alloc("A"), ctx.alloc("b")
    ctx.pack_x(u2, p2, x); ctx.pack_x(u1, p1, xp)
    ctx.pack_xi(rA["xi"], xi); ctx.pack_xi(rA["xi"], xip)
    assert ctx.forward_jacobian(x, xp, xip, xi, A, b) == 0
    torch.cuda.synchronize()
    bu, bp = ctx.unpack_x(b)
    scale = max(np.abs(rB["b"][0]).max(), 1e-300)
    assert np.abs(bu - rB["b"][0]).max() < TOL * scale
    for i in range(2):
        for j in range(2):
            rp, ci = ctx.csr_block_pattern(i, j)
            assert np.array_equal(ci, orc.graph(i, j)[1])
            v = ctx.csr_block_values(i, j, A)
            ref = rB["A"][i * 2 + j]
            assert np.abs(v - ref).max() < TOL * max(np.abs(ref).max(), 1e-300)
    ctx.close()


@pytest.mark.parametrize("name,owned_frac", [("3d_hyper_J2", 1.0), ("3d_small_hill", 1.0), ("2d_small_hill_plane_stress", 1.0),
                                             ("3d_small_J2", 0.6)])
def test_chunked_assembly_is_bit_identical(name, owned_frac):
    """The chunked two-phase assembly (element kernel + gather per chunk of elements, the chunk's element
    matrices staying in L2) gives the SAME BITS as one pass over the whole mesh: chunks run in element
    order and every block keeps its summation order.  Also with a partition's row filter (only the rows
    of the leading `owned` nodes are assembled)."""
    import torch
    from calibr8_b200.capi import Context
    from parity_common import LOCAL_TOL
    dim, gtype, ltype, params, amp = COMBOS[name]
    mesh = make_mesh(dim, "large")
    (u1, p1), (u2, p2) = synthetic_fields(mesh, amp, gtype == "mechanics")
    out = {}
    for chunk in (0, 96, 400):
        ctx = Context(0)
        ctx.set_assembly_chunk(chunk)
        ctx.set_mesh(mesh.dim, mesh.conn, mesh.coords)
        ctx.set_model(gtype, ltype, params, **LOCAL_TOL)
        ctx.set_stream(torch.cuda.current_stream().cuda_stream)
        if owned_frac < 1.0:   # a row filter as a partition sets it (no halo plan needed for the assembly)
            n_own = int(mesh.n_nodes * owned_frac)
            assert ctx.lib.c8_set_partition(ctx.h, n_own, mesh.n_elems) == 0
        x, xp, x0 = ctx.alloc("x"), ctx.alloc("x"), ctx.alloc("x")
        xi0, xip, xi = ctx.alloc("xi"), ctx.alloc("xi"), ctx.alloc("xi")
        ctx.pack_x(u2, p2, x); ctx.pack_x(u1, p1, xp); ctx.init_xi(xi0); ctx.init_xi(xip)
        assert ctx.forward_jacobian(xp, x0, xi0, xip, None, ctx.alloc("b")) == 0
        res = []
        for rep in range(2):      # twice: the second pass overwrites the first one's matrix
            A, b = ctx.alloc("A"), ctx.alloc("b")
            A.fill_(float("nan")); torch.cuda.synchronize()
            xi.copy_(xip)
            assert ctx.forward_jacobian(x, xp, xip, xi, A, b) == 0
            torch.cuda.synchronize()
            res.append((A.clone(), b.clone()))
        n_rows = ctx.n_nodes if owned_frac == 1.0 else int(mesh.n_nodes * owned_frac)
        rowptr, _ = ctx.bsr_pattern()
        nvals = int(rowptr[n_rows]) * ctx.nb * ctx.nb
        assert torch.equal(res[0][0][:nvals], res[1][0][:nvals])
        assert not torch.isnan(res[0][0][:nvals]).any()
        if owned_frac < 1.0:      # rows of the other nodes are left untouched
            assert torch.isnan(res[0][0][nvals:]).all()
        out[chunk] = (res[0][0][:nvals].clone(), res[0][1].clone())
        ctx.close()
    assert mesh.n_elems > 2 * 400
    for chunk in (96, 400):
        assert torch.equal(out[chunk][0], out[0][0]), chunk     # bit-identical matrix
        assert torch.allclose(out[chunk][1], out[0][1], rtol=1e-13, atol=1e-16)   # residual: atomics order
