"""End-to-end parity on the GPU: the C++ host Primal / Adjoint (calibr8_b200/host) driving the CUDA
kernels through the C ABI vs
  * the reference's own regression constants (test/primal/*.yaml.in `regression: QoI`), and
  * the CPU oracle's forward solve / adjoint gradient on the same decks (north-star: objective and
    adjoint gradient within 1e-8 relative)."""
import numpy as np
import pytest

from conftest import load_mesh

pytestmark = pytest.mark.gpu

J_TOL = 1e-8     # BASELINE.json north_star: objective within 1e-8 relative
G_TOL = 1e-8     # ... and its adjoint gradient


def gpu_problem(d, mesh, params=None, qoi="avg_disp", lin_tol=1e-12):
    import torch
    from calibr8_b200.capi import Context, HostProblem
    ctx = Context(0)
    ctx.set_mesh(mesh.dim, mesh.conn, mesh.coords)
    ctx.set_model(d["global_type"], d["local_type"], params or d["params"],
                  max_iters=max(d["local_max_iters"], 1), abs_tol=d["local_tol"], rel_tol=d["local_tol"])
    hp = HostProblem(ctx)
    hp.set_time(d["num_steps"], 1.0)
    for r, e, s, v in d["dbcs"]:
        hp.add_dbc(r, e, mesh.node_sets[s], v)
    hp.finalize_dbcs()
    hp.set_solver(d["global_max_iters"], d["global_tol"], d["global_tol"], gmres_restart=200,
                  gmres_max_iters=20000, linear_tol=lin_tol)
    if qoi == "avg_disp":
        hp.set_qoi_avg_disp()
    return ctx, hp


def oracle_problem(d, mesh, params=None, active=None):
    from oracle.driver import Dbc, Primal
    from oracle.pyoracle import Oracle
    o = Oracle(mesh.dim, mesh.conn, mesh.coords, global_type=d["global_type"],
               local_type=d["local_type"], params=[params or d["params"]],
               max_iters=d["local_max_iters"], abs_tol=d["local_tol"], rel_tol=d["local_tol"],
               active=active)
    bcs = [Dbc(r, e, mesh.node_sets[s], v) for r, e, s, v in d["dbcs"]]
    p = Primal(o, bcs, d["num_steps"], 1.0, max_iters=d["global_max_iters"],
               abs_tol=d["global_tol"], rel_tol=d["global_tol"])
    return o, p


DECKS = ["cube_elastic", "cube_hyper_J2", "notch_small_J2", "notch_hyper_J2", "notch2D_small_J2",
         "notch2D_small_J2_plane_strain", "notch2D_small_J2_plane_stress",
         "notch2D_hyper_J2_plane_stress", "notch2D_hyper_J2_plane_strain"]


@pytest.mark.parametrize("name", DECKS)
def test_forward_regression_on_gpu(golden, name):
    """The reference's regression decks solved on the GPU: J vs the golden constant (the
    reference's own tolerance) and vs the oracle's J (1e-8)."""
    d = golden["decks"][name]
    mesh = load_mesh(d["mesh"])
    ctx, hp = gpu_problem(d, mesh)
    J = hp.primal_solve()
    assert abs((J - d["J"]) / d["J"]) < d["rel_tol"], (J, d["J"])
    o, p = oracle_problem(d, mesh)
    o.set_qoi_avg_disp()
    J_o = p.solve()
    assert abs(J - J_o) / abs(J_o) < J_TOL, (J, J_o)
    # final state agrees too
    xs, xi = hp.get_step(d["num_steps"])
    scale = np.abs(p.x[-1][0]).max()
    assert np.abs(xs[0] - p.x[-1][0]).max() < 1e-7 * scale
    hp.close(); ctx.close()


@pytest.mark.parametrize("name,active", [
    ("notch2D_small_J2", ["E", "nu", "K", "Y"]),                 # test/adjoint/notch2D_small_J2_adjoint_check
    ("notch2D_small_J2_plane_stress", ["Y", "S", "D"]),          # the shipped example's parameters
    ("cube_hyper_J2", ["E", "nu", "Y", "K"]),
])
def test_adjoint_gradient_vs_oracle(golden, name, active):
    """Avg-displacement objective: adjoint gradient on the GPU vs the oracle's adjoint gradient."""
    from oracle.driver import Adjoint
    from oracle.pyoracle import PARAM_NAMES
    d = dict(golden["decks"][name])
    d["num_steps"] = min(d["num_steps"], 8)
    mesh = load_mesh(d["mesh"])
    names = PARAM_NAMES[d["local_type"]]
    act = [names.index(a) for a in active]
    ctx, hp = gpu_problem(d, mesh)
    J = hp.primal_solve()
    g_gpu = hp.adjoint_gradient()[act]
    o, p = oracle_problem(d, mesh, active=[act])
    o.set_qoi_avg_disp()
    J_o = p.solve()
    g_o = Adjoint(p, max_iters=d["global_max_iters"], abs_tol=1e-14, rel_tol=1e-12).gradient(
        [list(range(len(act)))], len(act))
    assert abs(J - J_o) / abs(J_o) < J_TOL
    assert np.abs(g_gpu - g_o).max() < G_TOL * np.abs(g_o).max(), (g_gpu, g_o)
    hp.close(); ctx.close()
