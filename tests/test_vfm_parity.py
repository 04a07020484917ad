"""VFM (virtual fields method) parity, BASELINE configs[3] in miniature: forward-sensitivity and
adjoint-sensitivity objectives + gradients w.r.t. Y, S, D on the reference's notch2D mesh with
the shipped example's virtual fields (examples/.../inverse_vfm/...yaml:59-61), GPU vs oracle."""
import numpy as np
import pytest

from conftest import load_mesh
from test_solve_parity import _calibration_setup, gpu_problem, oracle_problem

pytestmark = pytest.mark.gpu


def _virtual_field(mesh):
    x, y = mesh.coords[:, 0], mesh.coords[:, 1]
    wx = np.cos(4. * np.arctan(1.) * (y - 0.5) / 1.) * x
    wy = (y * (2 * (y - 0.5) + 1.) / (2 * 1.))
    return np.stack([wx, wy], axis=1)


def _synthetic(mesh):
    truth, start, deck, area, qoi = _calibration_setup(mesh)
    o, p = oracle_problem(deck, mesh)
    o.set_qoi_avg_disp()
    p.solve()
    N = deck["num_steps"]
    measured = np.stack([p.x[s][0].reshape(-1, 2) for s in range(1, N + 1)])
    w = _virtual_field(mesh)
    # external virtual power from the true internal force: P_ext = R_int(u_true) . w
    loads = []
    for s in range(1, N + 1):
        b = o.global_residual(p.x[s], p.x[s - 1], p.xi[s], p.xi[s - 1])
        loads.append(float(b[0] @ w.reshape(-1)))
    return start, deck, measured, w, np.array(loads)


@pytest.mark.parametrize("mode", ["forward", "adjoint"])
def test_vfm_objective_and_gradient(mode):
    from calibr8_b200.vfm import vfm_objective as gpu_vfm
    from oracle.pyoracle import PARAM_NAMES
    from oracle.vfm_driver import vfm_objective as orc_vfm
    mesh = load_mesh("notch2D")
    start, deck, measured, w, loads = _synthetic(mesh)
    names = PARAM_NAMES[deck["local_type"]]
    act = [names.index(a) for a in ("Y", "S", "D")]
    d2 = dict(deck, params=start)
    o2, _ = oracle_problem(d2, mesh, active=[act])
    J_o, g_o = orc_vfm(o2, mode, measured, w, loads, 3, obj_scale_factor=1e2)
    assert J_o > 0 and np.abs(g_o).max() > 0
    ctx, hp = gpu_problem(d2, mesh, qoi=None)
    J, g = gpu_vfm(hp, mode, measured, w, loads, obj_scale_factor=1e2)
    g = g[act]
    assert abs(J - J_o) / abs(J_o) < 1e-8, (J, J_o)
    assert np.abs(g - g_o).max() < 1e-8 * np.abs(g_o).max(), (g, g_o)
    hp.close(); ctx.close()
