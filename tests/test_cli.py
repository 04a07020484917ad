"""The `primal` / `objective` command-line front end (calibr8_b200/cli.py): deck parsing on CPU,
and on the GPU the shipped synthetic-calibration flow (BASELINE configs[0]) end to end through the
text-file process boundary of the reference (objective_value.txt / objective_gradient.txt, load.dat)."""
import os
import textwrap

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MESH = os.path.join(ROOT, "tests", "golden", "mesh_notch2D.npz")

COMMON = """
  discretization:
    mesh file: '{mesh}'
    num steps: 3
    step size: 1.
  residuals:
    global residual:
      type: 'mechanics_plane_stress'
      nonlinear max iters: 30
      nonlinear absolute tol: 1.e-12
      nonlinear relative tol: 1.e-12
      print convergence: false
    local residual:
      type: 'small_hill_plane_stress'
      nonlinear max iters: 20
      nonlinear absolute tol: 1.e-12
      nonlinear relative tol: 1.e-12
      materials:
        body: {{E: 1000., nu: 0.25, Y: {Y}, S: {S}, D: {D}, R00: 1., R11: 1., R22: 1., R01: 1.}}
  dirichlet bcs:
    expression:
      bc 1: [0, 0, xmin, 0.0]
      bc 2: [0, 1, ymin, 0.0]
      bc 3: [0, 1, ymax, 0.01 * t]
"""
FORWARD = "fwd:\n  problem:\n    name: fwd\n    write synthetic: true\n" + COMMON + """
  quantity of interest:
    type: 'reaction mismatch'
    coordinate index: 1
    coordinate value: 1.
    load out file: "load.dat"
    reaction force component: 1
"""
PDECO = "pdeco:\n  problem:\n    name: pdeco\n" + COMMON + """
  quantity of interest:
    type: 'calibration'
    coordinate index: 1
    coordinate value: 1.
    load input file: "load.dat"
    reaction force component: 1
    displacement weights: [1e8, 1e8]
    balance factor: 1e2
  inverse:
    objective type: "adjoint"
    materials:
      body:
        Y: [1., 3.]
        S: [5., 15.]
        D: [40., 80.]
"""
VFM = "vfm:\n  problem:\n    name: vfm\n" + COMMON + """
  inverse:
    objective type: "{otype}"
    objective scale factor: 1e2
    thickness: 1.
    load input file: "load.dat"
    materials:
      body:
        Y: [1., 3.]
        S: [5., 15.]
        D: [40., 80.]
  virtual fields:
    w_x: 'cos(4. * atan(1.) * (y - 0.5) / 1.) * x'
    w_y: '(y * (2 * (y - 0.5) + 1.) / (2 * 1.))'
"""


def _write(path, text):
    with open(path, "w") as f:
        f.write(textwrap.dedent(text))


def test_deck_parsing_and_expression_evaluator(tmp_path):
    from calibr8_b200 import cli, meshio
    from calibr8_b200.capi import eval_expr
    p = tmp_path / "pdeco.yaml"
    _write(p, PDECO.format(mesh=MESH, Y=2.2, S=8., D=60.))
    name, deck = cli.load_deck(str(p))
    assert name == "pdeco" and deck["discretization"]["num steps"] == 3
    mesh = meshio.load_npz(MESH)
    act = cli._active(deck, mesh, "small_hill_plane_stress")
    assert act == [(0, 2), (0, 3), (0, 4)]                       # Y, S, D in parameter order
    mats = cli._materials(deck["residuals"]["local residual"], mesh)
    assert mats[0]["Y"] == 2.2 and list(mats[0]) == ["E", "nu", "Y", "S", "D", "R00", "R11", "R22", "R01"]
    xyz = np.array([[0.5, 1.0, 0.0], [0.25, 0.5, 0.0]])
    assert np.allclose(eval_expr("(y * (2 * (y - 0.5) + 1.) / (2 * 1.))", xyz), [1.0, 0.25])
    assert np.allclose(eval_expr("cos(4. * atan(1.) * (y - 0.5) / 1.) * x", xyz), [0.0, 0.25], atol=1e-15)
    assert np.allclose(eval_expr("0.01 * t", xyz, t=3.0), 0.03)
    with pytest.raises(Exception):
        eval_expr("nosuch(x)", xyz)


@pytest.mark.gpu
def test_synthetic_calibration_flow_through_text_files(tmp_path, monkeypatch):
    from calibr8_b200 import cli
    monkeypatch.chdir(tmp_path)
    _write("fwd.yaml", FORWARD.format(mesh=MESH, Y=2., S=10., D=50.))
    cli.main(["primal", "fwd.yaml"])
    loads = np.loadtxt("load.dat")
    assert loads.shape == (3,) and (loads > 0).all() and loads[2] > loads[0]
    assert os.path.exists("fwd_synthetic/measured.npz")
    synth = os.path.join(str(tmp_path), "fwd_synthetic") + "/"
    # objective at the truth: zero; at the inverse deck's starting point: positive, with a gradient
    _write("truth.yaml", PDECO.format(mesh=synth, Y=2., S=10., D=50.))
    cli.main(["objective", "truth.yaml", "false", "truth"])
    assert abs(float(open("objective_value_truth.txt").read())) < 1e-12
    _write("start.yaml", PDECO.format(mesh=synth, Y=2.2, S=8., D=60.))
    cli.main(["objective", "start.yaml", "true"])
    J = float(open("objective_value.txt").read())
    g = np.loadtxt("objective_gradient.txt")
    assert J > 0 and g.shape == (3,)
    assert len(open("objective_value.txt").read().strip().split("e")[0]) >= 19      # %.17e
    # directional finite difference through the same text boundary
    d = np.array([0.2, -1.0, 5.0])
    h = 1e-5
    Js = []
    for sgn in (+1, -1):
        _write("fd.yaml", PDECO.format(mesh=synth, Y=2.2 + sgn * h * d[0], S=8. + sgn * h * d[1], D=60. + sgn * h * d[2]))
        cli.main(["objective", "fd.yaml", "false", "fd"])
        Js.append(float(open("objective_value_fd.txt").read()))
    fd = (Js[0] - Js[1]) / (2 * h)
    assert abs(g @ d - fd) < 1e-5 * abs(fd), (g @ d, fd)
    # virtual-fields objectives on the same data: zero at the truth; forward-sensitivity and adjoint
    # gradients agree
    out = {}
    for otype in ("FS_VFM", "Adjoint_VFM"):
        _write("vfm.yaml", VFM.format(mesh=synth, Y=2.2, S=8., D=60., otype=otype))
        cli.main(["objective", "vfm.yaml", "true", otype])
        out[otype] = (float(open(f"objective_value_{otype}.txt").read()), np.loadtxt(f"objective_gradient_{otype}.txt"))
    assert out["FS_VFM"][0] > 0
    assert abs(out["FS_VFM"][0] - out["Adjoint_VFM"][0]) < 1e-10 * out["FS_VFM"][0]
    assert np.abs(out["FS_VFM"][1] - out["Adjoint_VFM"][1]).max() < 1e-8 * np.abs(out["FS_VFM"][1]).max()
    _write("vfm0.yaml", VFM.format(mesh=synth, Y=2., S=10., D=50., otype="FS_VFM"))
    cli.main(["objective", "vfm0.yaml", "false", "vfm0"])
    assert float(open("objective_value_vfm0.txt").read()) < 1e-6 * out["FS_VFM"][0]


@pytest.mark.gpu
def test_inverse_recovers_the_true_parameters(tmp_path, monkeypatch):
    """BASELINE configs[0] end to end: the shipped synthetic calibration (forward run at the true
    Y, S, D = 2, 10, 50 -> synthetic data -> `inverse` from 2.2, 8, 60 within the deck's bounds)
    recovers the truth, with the PDE-constrained (adjoint) objective."""
    from calibr8_b200 import cli
    from calibr8_b200.capi import Objective
    monkeypatch.chdir(tmp_path)
    _write("fwd.yaml", FORWARD.format(mesh=MESH, Y=2., S=10., D=50.))
    cli.main(["primal", "fwd.yaml"])
    synth = os.path.join(str(tmp_path), "fwd_synthetic") + "/"
    deck = PDECO.format(mesh=synth, Y=2.2, S=8., D=60.) + """    iteration limit: 60
    gradient tolerance: 1e-10
"""
    _write("inv.yaml", deck)
    params, J, nevals = cli.run_inverse("inv.yaml")
    assert J < 1e-9, J
    assert abs(params["Y"] - 2.) < 1e-4 and abs(params["S"] - 10.) < 2e-2 and abs(params["D"] - 50.) < 0.2, params
    got = np.loadtxt("inverse_result.txt")
    assert np.allclose(got, [params["Y"], params["S"], params["D"]])
    assert os.path.exists("ROL_out.txt") and nevals < 200
    # canonical scaling of the Objective (src/objective.cpp:41-61): bounds Y [1,3], S [5,15], D [40,80]
    name, d = cli.load_deck("inv.yaml")
    ctx, hp, obj, names, p0 = cli._setup_objective(d, str(tmp_path))
    assert names == ["Y", "S", "D"]
    assert np.allclose(p0, [(2.2 - 2.) / 1., (8. - 10.) / 5., (60. - 60.) / 20.])
    assert np.allclose(obj.to_physical([1., -1., 0.5]), [3., 5., 70.])
    assert np.allclose(obj.to_canonical([10., 0., 60.]), [1., -1., 0.])          # clamped
    # canonical gradient = span * physical gradient, and value() re-uses the cached forward solve
    g_can = obj.gradient(p0)
    _write("start.yaml", PDECO.format(mesh=synth, Y=2.2, S=8., D=60.))
    Jp, g_phys = cli.run_objective("start.yaml", True)
    assert abs(obj.value(p0) - Jp) < 1e-10 * Jp
    assert np.allclose(g_can, g_phys * np.array([1., 5., 20.]), rtol=1e-7)
    obj.close(); hp.close(); ctx.close()
