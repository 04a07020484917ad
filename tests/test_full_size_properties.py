"""Size-independent properties at BASELINE.json's full size (1M tets on one GPU), where the CPU
oracle is too slow to be the checker:

* K1 + gather + K9: the assembled Jacobian applied to a direction equals the central difference of
  the assembled residual (with the local state re-solved at each perturbed point) -- checks the
  element kernel, the two-phase assembly, the BSR SpMV and K2's consistency at full size;
* the assembly is bit-reproducible (deterministic gather plan) and a checksum of the BSR values is
  invariant under a permutation of the element order;
* adjoint gradient vs a central finite difference of the objective (the reference's own acceptance
  method, src/main_inverse.cpp:126-140) on a 190k-tet box.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

PAR_J2 = dict(E=1000., nu=.25, K=100., Y=2., cte=0., delta_T=0.)
# linear hardening (the reference's notch_hyper_J2 deck): a power-law term A (alpha + 1e-12)^n has an
# unbounded second derivative at the onset of yielding, which defeats a finite-difference check
PAR_HYPER = dict(E=1000., nu=.25, Y=10., S=0., D=0., A=0., n=0., K=100.)
LOCAL = dict(max_iters=500, abs_tol=1e-12, rel_tol=1e-12)


def _state(ctx, mesh, amp, seed=0):
    import bench
    (u1, p1), (u2, p2) = bench.workload_fields(mesh, seed=seed)
    s = amp / bench.AMP
    x, xp, x0 = ctx.alloc("x"), ctx.alloc("x"), ctx.alloc("x")
    xi0, xip = ctx.alloc("xi"), ctx.alloc("xi")
    ctx.pack_x(u2 * s, p2, x); ctx.pack_x(u1 * s, p1, xp)
    ctx.init_xi(xi0); ctx.init_xi(xip)
    b = ctx.alloc("b")
    assert ctx.forward_jacobian(xp, x0, xi0, xip, None, b) == 0
    return x, xp, xip


@pytest.mark.parametrize("ltype,params,amp", [("hyper_J2", PAR_HYPER, 2.2e-3), ("small_J2", PAR_J2, 0.45e-3)])
def test_jacobian_is_derivative_of_residual_1M(ltype, params, amp):
    import torch
    import bench
    from calibr8_b200.capi import Context
    mesh = bench.workload_mesh(56)
    ctx = Context(0)
    ctx.set_mesh(3, mesh.conn, mesh.coords)
    ctx.set_model("mechanics", ltype, params, **LOCAL)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)   # torch's copies and the kernels in order
    assert ctx.n_elems > 1_000_000
    x, xp, xip = _state(ctx, mesh, amp)
    xi, A, b, path = ctx.alloc("xi"), ctx.alloc("A"), ctx.alloc("b"), ctx.alloc("path")
    xi.copy_(xip)
    assert ctx.forward_jacobian(x, xp, xip, xi, A, b, path) == 0
    plastic = float(path.float().mean())
    assert 0.2 < plastic < 0.8
    # bit-reproducible assembly
    A2, b2, xi2 = ctx.alloc("A"), ctx.alloc("b"), ctx.alloc("xi")
    xi2.copy_(xip)
    assert ctx.forward_jacobian(x, xp, xip, xi2, A2, b2) == 0
    assert torch.equal(A, A2) and torch.equal(xi, xi2)
    # J v vs central difference of the residual along a smooth direction v
    g = torch.Generator(device="cpu").manual_seed(5)
    coords = torch.from_numpy(mesh.coords)
    v = torch.zeros(ctx.n_nodes, 4, dtype=torch.float64)
    for k in range(3):
        v[:, k] = torch.sin(2.0 * coords[:, (k + 1) % 3] + 0.3 * k) * torch.cos(1.5 * coords[:, k])
    v[:, 3] = torch.cos(coords.sum(dim=1))
    v = v.reshape(-1).cuda()
    Jv = ctx.alloc("x")
    ctx.spmv(A, v, Jv)
    eps = 1e-6 * float(x.abs().max())   # well above the noise of the local Newton tolerance
    R, paths = [], []
    for sgn in (+1.0, -1.0):
        xe = x + sgn * eps * v
        xie, be, pe = ctx.alloc("xi"), ctx.alloc("b"), ctx.alloc("path")
        xie.copy_(xip)
        assert ctx.forward_jacobian(xe, xp, xip, xie, None, be, pe) == 0
        R.append(be); paths.append(pe)
    fd = (R[0] - R[1]) / (2 * eps)
    torch.cuda.synchronize()
    # quadrature points that change branch between x - eps v and x + eps v have a kink there: the
    # rows of their nodes are excluded (and must be a tiny fraction)
    switched = ((paths[0] != path) | (paths[1] != path)).cpu().numpy()
    bad_nodes = np.unique(mesh.conn[switched].ravel())
    keep = torch.ones(ctx.n_nodes, 4, dtype=torch.bool)
    keep[torch.from_numpy(bad_nodes).long()] = False
    keep = keep.reshape(-1).cuda()
    assert bad_nodes.size < 0.03 * ctx.n_nodes, bad_nodes.size
    err = float(((Jv - fd).abs() * keep).max() / Jv.abs().max())
    assert err < 1e-5, (err, bad_nodes.size)
    # K2 (residual only, state given) reproduces K1's residual
    b3 = ctx.alloc("b")
    ctx.global_residual(x, xp, xi, xip, b3)
    torch.cuda.synchronize()
    assert float((b3 - b).abs().max()) <= 1e-12 * float(b.abs().max())
    ctx.close()


def test_assembly_checksum_is_element_order_invariant():
    import torch
    import bench
    from calibr8_b200.capi import Context
    mesh = bench.workload_mesh(24)
    sums = []
    for perm_seed in (None, 3):
        conn = mesh.conn
        if perm_seed is not None:
            conn = conn[np.random.RandomState(perm_seed).permutation(conn.shape[0])]
        ctx = Context(0)
        ctx.set_mesh(3, conn, mesh.coords)
        ctx.set_model("mechanics", "small_J2", PAR_J2, **LOCAL)
        ctx.set_stream(torch.cuda.current_stream().cuda_stream)
        x, xp, xip = _state(ctx, mesh, 0.45e-3)
        xi, A, b = ctx.alloc("xi"), ctx.alloc("A"), ctx.alloc("b")
        xi.copy_(xip)
        assert ctx.forward_jacobian(x, xp, xip, xi, A, b) == 0
        torch.cuda.synchronize()
        sums.append((A.double().sum().item(), A.abs().sum().item(), b.abs().sum().item(),
                     ctx.bsr_pattern()[1].copy(), A.cpu().numpy().copy()))
        ctx.close()
    # same BSR pattern (it depends on the node graph only) and the same values up to summation order
    assert np.array_equal(sums[0][3], sums[1][3])
    scale = np.abs(sums[0][4]).max()
    assert np.abs(sums[0][4] - sums[1][4]).max() < 1e-12 * scale
    assert abs(sums[0][1] - sums[1][1]) < 1e-10 * sums[0][1]


def test_adjoint_gradient_vs_finite_difference_190k():
    from calibr8_b200 import meshgen
    from calibr8_b200.capi import Context, HostProblem
    mesh = meshgen.box_tets(32, notch_radius=0.2)
    names = ["E", "nu", "K", "Y", "cte", "delta_T"]

    def solve(params, adjoint):
        ctx = Context(0)
        ctx.set_mesh(3, mesh.conn, mesh.coords)
        ctx.set_model("mechanics", "small_J2", params, **LOCAL)
        hp = HostProblem(ctx)
        hp.set_time(2, 1.0)
        hp.add_dbc(0, 0, mesh.node_sets["xmin"], "0.0")
        hp.add_dbc(0, 1, mesh.node_sets["ymin"], "0.0")
        hp.add_dbc(0, 2, mesh.node_sets["zmin"], "0.0")
        hp.add_dbc(0, 1, mesh.node_sets["ymax"], "0.002 * t")
        hp.finalize_dbcs()
        hp.set_solver(20, 1e-11, 1e-11, gmres_restart=100, gmres_max_iters=5000, linear_tol=1e-11)
        hp.set_qoi_avg_disp()
        J = hp.primal_solve()
        g = hp.adjoint_gradient() if adjoint else None
        hp.close(); ctx.close()
        return J, g

    J0, g = solve(PAR_J2, True)
    d = np.array([0.3, 0.0, 0.2, 1.0, 0.0, 0.0])     # direction in (E, nu, K, Y) space, scaled below
    scale = np.array([PAR_J2[k] for k in names])
    d = d * scale
    h = 1e-5
    Jp, _ = solve({k: PAR_J2[k] + h * d[i] for i, k in enumerate(names)}, False)
    Jm, _ = solve({k: PAR_J2[k] - h * d[i] for i, k in enumerate(names)}, False)
    fd = (Jp - Jm) / (2 * h)
    ad = float(g @ d)
    assert abs(ad - fd) < 2e-5 * abs(fd), (ad, fd)
