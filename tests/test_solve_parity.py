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
    for tb in d.get("tbcs", []):
        hp.add_tbc(tb[0], mesh.side_sets[tb[1]], tb[2:])
    hp.set_solver(d["global_max_iters"], d["global_tol"], d["global_tol"], gmres_restart=200,
                  gmres_max_iters=20000, linear_tol=lin_tol)
    if qoi == "avg_disp":
        hp.set_qoi_avg_disp()
    return ctx, hp


def oracle_problem(d, mesh, params=None, active=None):
    from oracle.driver import Dbc, Primal, Tbc
    from oracle.pyoracle import Oracle
    o = Oracle(mesh.dim, mesh.conn, mesh.coords, global_type=d["global_type"],
               local_type=d["local_type"], params=[params or d["params"]],
               max_iters=d["local_max_iters"], abs_tol=d["local_tol"], rel_tol=d["local_tol"],
               active=active)
    bcs = [Dbc(r, e, mesh.node_sets[s], v) for r, e, s, v in d["dbcs"]]
    tbcs = [Tbc(t[0], mesh.side_sets[t[1]], t[2:]) for t in d.get("tbcs", [])]
    p = Primal(o, bcs, d["num_steps"], 1.0, max_iters=d["global_max_iters"],
               abs_tol=d["global_tol"], rel_tol=d["global_tol"], tbc=tbcs or None)
    return o, p


DECKS = ["cube_elastic", "cube_hyper_J2", "notch_small_J2", "notch_hyper_J2", "notch2D_small_J2",
         "notch2D_small_J2_plane_strain", "notch2D_small_J2_plane_stress",
         "notch2D_hyper_J2_plane_stress", "notch2D_hyper_J2_plane_strain",
         "cube_hyperelasticity_traction",       # traction bcs (src/tbcs.cpp:17-98)
         "notch_hypo_J2", "notch2D_hypo_J2_plane_strain", "notch2D_hypo_J2_plane_stress"]


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
    ("notch2D_hypo_J2_plane_stress", ["Y", "S", "D", "R11"]),
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


def _calibration_setup(mesh):
    """examples/synthetic_calibration: forward/notch2D_small_J2_plane_stress.yaml:14-52 and
    inverse_pdeco/pdeco_notch2D_small_J2_plane_stress.yaml:25-68."""
    truth = dict(E=1000., nu=.25, Y=2., S=10., D=50., R00=1., R11=1., R22=1., R01=1.)
    start = dict(truth, Y=2.2, S=8., D=60.)
    deck = dict(global_type="mechanics_plane_stress", local_type="small_hill_plane_stress",
                params=truth, dbcs=[[0, 0, "xmin", "0.0"], [0, 1, "ymin", "0.0"], [0, 1, "ymax", "0.01 * t"]],
                num_steps=4, global_max_iters=30, global_tol=1e-12, local_max_iters=20, local_tol=1e-12)
    X = mesh.coords[mesh.conn]
    area = 0.5 * np.abs((X[:, 1, 0] - X[:, 0, 0]) * (X[:, 2, 1] - X[:, 0, 1]) -
                        (X[:, 1, 1] - X[:, 0, 1]) * (X[:, 2, 0] - X[:, 0, 0])).sum()
    qoi = dict(balance_factor=1e2, coord_idx=1, coord_value=1.0, reaction_force_comp=1,
               weights=(1e8, 1e8))
    return truth, start, deck, area, qoi


def test_calibration_objective_and_gradient(golden):
    """BASELINE configs[0]: the shipped synthetic calibration (2-D plane-stress Hill, calibration
    QoI = displacement mismatch + load mismatch): objective and adjoint gradient w.r.t. Y, S, D at
    the inverse deck's starting point, GPU vs oracle."""
    from oracle.driver import Adjoint
    from oracle.pyoracle import PARAM_NAMES
    mesh = load_mesh("notch2D")
    truth, start, deck, area, qoi = _calibration_setup(mesh)
    N = deck["num_steps"]
    names = PARAM_NAMES[deck["local_type"]]
    act = [names.index(a) for a in ("Y", "S", "D")]
    # ---- synthetic data from the oracle's forward run at the true parameters
    o, p = oracle_problem(deck, mesh)
    o.set_qoi_calibration(**qoi)
    zero_meas = np.zeros((mesh.n_nodes, 3))
    loads = []

    def setup_truth(step):
        o.qoi_set_step(1.0, float(N), 0.0, zero_meas)
    p.solve(setup_truth)
    measured = []
    for step in range(1, N + 1):
        o.qoi_set_step(1.0, float(N), 0.0, zero_meas)
        o.qoi(p.x[step], p.x[step - 1], p.xi[step], p.xi[step - 1], step)
        loads.append(o.calibration_state()["total_load"])
        m3 = np.zeros((mesh.n_nodes, 3)); m3[:, :2] = p.x[step][0].reshape(-1, 2)
        measured.append(m3)
    assert abs(o.calibration_state()["area"] - area) < 1e-12
    # ---- oracle objective + gradient at the starting parameters
    d2 = dict(deck, params=start)
    o2, p2 = oracle_problem(d2, mesh, active=[act])
    o2.set_qoi_calibration(**qoi)

    def setup(step):
        o2.qoi_set_step(1.0, float(N), loads[step - 1], measured[step - 1])
    J_o = p2.solve(setup)
    g_o = Adjoint(p2, max_iters=30, abs_tol=1e-14, rel_tol=1e-12).gradient(
        [list(range(3))], 3, setup)
    assert J_o > 0
    # ---- GPU
    ctx, hp = gpu_problem(d2, mesh, qoi=None)
    meas2 = np.stack([m[:, :2] for m in measured])
    hp.set_qoi_calibration(measured=meas2, load_data=loads, area=area, **qoi)
    J = hp.primal_solve()
    g = hp.adjoint_gradient()[act]
    assert abs(J - J_o) / abs(J_o) < J_TOL, (J, J_o)
    assert np.abs(g - g_o).max() < G_TOL * np.abs(g_o).max(), (g, g_o)
    hp.close(); ctx.close()


def test_calibration_3d_hill_objective_and_gradient():
    """BASELINE configs[2] in miniature: 3-D Hill anisotropic plasticity, `calibration` QoI with the
    displacement mismatch on the zmax face (Calibration::compute_surface_mismatch,
    src/calibration.cpp:225-303) and the reaction on the ymax plane, adjoint gradient w.r.t. the 8
    active parameters Y, S, D, R00, R11, R01, R02, R12 -- GPU vs oracle."""
    from calibr8_b200 import meshgen
    from oracle.driver import Adjoint
    from oracle.pyoracle import PARAM_NAMES
    from test_adjoint_parity import facets_on_plane
    from parity_common import HILL3D
    mesh = meshgen.box_tets(4, notch_radius=0.3)
    truth = dict(HILL3D)
    start = dict(truth, Y=2.2, S=8., D=2.5, R00=1.05, R11=0.95, R01=1.05, R02=1.0, R12=1.0)
    N = 4
    deck = dict(global_type="mechanics", local_type="small_hill", params=truth,
                dbcs=[[0, 0, "xmin", "0.0"], [0, 1, "ymin", "0.0"], [0, 2, "zmin", "0.0"],
                      [0, 1, "ymax", "0.0012 * t"]],
                num_steps=N, global_max_iters=30, global_tol=1e-11, local_max_iters=60, local_tol=1e-12)
    names = PARAM_NAMES["small_hill"]
    act = [names.index(a) for a in ("Y", "S", "D", "R00", "R11", "R01", "R02", "R12")]
    fac = facets_on_plane(mesh, 2, 1.0)
    qoi = dict(balance_factor=1e2, coord_idx=1, coord_value=1.0, reaction_force_comp=1,
               weights=(1e8, 1e8, 1e8))
    # synthetic data at the true parameters (oracle)
    o, p = oracle_problem(deck, mesh)
    o.set_qoi_calibration(facet=fac, **qoi)
    zero_meas = np.zeros((mesh.n_nodes, 3))
    p.solve(lambda step: o.qoi_set_step(1.0, float(N), 0.0, zero_meas))
    assert p.xi[N][:, 6].max() > 1e-4, "the truth run must yield"
    loads, measured = [], []
    for step in range(1, N + 1):
        o.qoi_set_step(1.0, float(N), 0.0, zero_meas)
        o.qoi(p.x[step], p.x[step - 1], p.xi[step], p.xi[step - 1], step)
        loads.append(o.calibration_state()["total_load"])
        measured.append(p.x[step][0].reshape(-1, 3).copy())
    area = o.calibration_state()["area"]
    # oracle objective + gradient at the starting parameters
    d2 = dict(deck, params=start)
    o2, p2 = oracle_problem(d2, mesh, active=[act])
    o2.set_qoi_calibration(facet=fac, **qoi)

    def setup(step):
        o2.qoi_set_step(1.0, float(N), loads[step - 1], measured[step - 1])
    J_o = p2.solve(setup)
    g_o = Adjoint(p2, max_iters=30, abs_tol=1e-14, rel_tol=1e-12).gradient([list(range(len(act)))], len(act), setup)
    assert J_o > 0 and (np.abs(g_o) > 0).all()
    # GPU
    ctx, hp = gpu_problem(d2, mesh, qoi=None)
    hp.set_qoi_calibration(measured=np.stack(measured), load_data=loads, area=area, facet=fac, **qoi)
    J = hp.primal_solve()
    g = hp.adjoint_gradient()[act]
    assert abs(J - J_o) / abs(J_o) < J_TOL, (J, J_o)
    assert np.abs(g - g_o).max() < G_TOL * np.abs(g_o).max(), (g, g_o)
    hp.close(); ctx.close()


@pytest.mark.parametrize("kind", ["reaction mismatch", "load mismatch", "surface mismatch"])
def test_mismatch_qoi_objective_and_gradient(kind):
    """The reaction / load / surface mismatch objectives end to end (forward solves, two-pass total
    load, adjoint gradient) on the GPU vs the oracle: src/reaction_mismatch.cpp, load_mismatch.cpp,
    surface_mismatch.cpp; "load out file" lines (load.dat) included."""
    from calibr8_b200 import meshgen
    from oracle.driver import Adjoint
    from oracle.pyoracle import PARAM_NAMES
    mesh = meshgen.box_tets(3, notch_radius=0.3)
    params = dict(E=1000., nu=.25, K=100., Y=2., cte=0., delta_T=0.)
    N = 3
    deck = dict(global_type="mechanics", local_type="small_J2", params=params,
                dbcs=[[0, 0, "xmin", "0.0"], [0, 1, "ymin", "0.0"], [0, 2, "zmin", "0.0"], [0, 1, "ymax", "0.0015 * t"]],
                num_steps=N, global_max_iters=30, global_tol=1e-11, local_max_iters=60, local_tol=1e-13)
    names = PARAM_NAMES["small_J2"]
    act = [names.index(a) for a in ("E", "nu", "K", "Y")]
    rng = np.random.RandomState(4)
    meas = [1e-3 * rng.uniform(-1, 1, size=(mesh.n_nodes, 3)) for _ in range(N)]
    load_meas = [0.3, 0.5, 0.6]

    def facets(plane):
        on = np.abs(mesh.coords[mesh.conn][:, :, plane] - 1.0) < 1e-12
        fac = np.full((mesh.n_elems, 3), -1, dtype=np.int32)
        for e in np.nonzero(on.sum(axis=1) == 3)[0]:
            fac[e] = np.nonzero(on[e])[0]
        return fac
    okw = {"reaction mismatch": dict(coord_idx=1, coord_value=1.0, reaction_force_comp=1),
           "load mismatch": dict(facet=facets(1)), "surface mismatch": dict(facet=facets(2))}[kind]
    o, p = oracle_problem(deck, mesh, active=[act])
    o.set_qoi_mismatch(kind.split()[0], **okw)
    setup = lambda step: o.qoi_set_step(1.0, float(N), load_meas[step - 1], meas[step - 1])
    J_o = p.solve(setup)
    g_o = Adjoint(p, max_iters=30, abs_tol=1e-14, rel_tol=1e-12).gradient([list(range(4))], 4, setup)
    loads_o = []
    for step in range(1, N + 1):
        setup(step)
        o.qoi(p.x[step], p.x[step - 1], p.xi[step], p.xi[step - 1], step)
        loads_o.append(o.calibration_state()["total_load"])
    assert J_o > 0 and p.xi[-1][:, -1].max() > 0
    ctx, hp = gpu_problem(deck, mesh, qoi=None)
    hp.set_qoi_mismatch(kind, measured=np.stack(meas) if kind == "surface mismatch" else None,
                        load_data=None if kind == "surface mismatch" else load_meas, **okw)
    J = hp.primal_solve()
    g = hp.adjoint_gradient()[act]
    assert abs(J - J_o) / abs(J_o) < J_TOL, (J, J_o)
    assert np.abs(g - g_o).max() < G_TOL * np.abs(g_o).max(), (g, g_o)
    if kind != "surface mismatch":
        assert np.abs(hp.loads() - np.array(loads_o)).max() < 1e-9 * np.abs(loads_o).max()
    hp.close(); ctx.close()
