"""Pins the oracle's derivative paths with the reference's own acceptance method: adjoint /
sensitivity gradients vs finite differences of the objective (src/main_inverse.cpp:126-140 uses
ROL::checkGradient; test/adjoint and test/vfm decks).  CPU only."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, load_mesh
from oracle.driver import Adjoint, Dbc, Primal
from oracle.pyoracle import PARAM_NAMES, Oracle

G = json.load(open(os.path.join(GOLDEN, "golden.json")))


def _solve(d, mesh, params, act, nsteps, tol=1e-12):
    o = Oracle(mesh.dim, mesh.conn, mesh.coords, global_type=d["global_type"],
               local_type=d["local_type"], params=[params], max_iters=d["local_max_iters"],
               abs_tol=d["local_tol"], rel_tol=d["local_tol"], active=[act])
    o.set_qoi_avg_disp()
    p = Primal(o, [Dbc(r, e, mesh.node_sets[s], v) for r, e, s, v in d["dbcs"]], nsteps, 1.0,
               max_iters=d["global_max_iters"], abs_tol=tol, rel_tol=tol)
    return o, p, p.solve()


@pytest.mark.parametrize("name,active,nsteps", [
    ("notch2D_small_J2", ["E", "nu", "K", "Y"], 8),        # test/adjoint/notch2D_small_J2_adjoint_check
    ("cube_hyper_J2", ["E", "nu", "Y", "K"], 6),
])
def test_adjoint_gradient_matches_finite_differences(name, active, nsteps):
    d = G["decks"][name]
    mesh = load_mesh(d["mesh"])
    names = PARAM_NAMES[d["local_type"]]
    act = [names.index(a) for a in active]
    o, p, J = _solve(d, mesh, d["params"], act, nsteps)
    g = Adjoint(p, abs_tol=1e-14, rel_tol=1e-12).gradient([list(range(len(act)))], len(act))
    for k, a in enumerate(active):
        pp, pm = dict(d["params"]), dict(d["params"])
        h = 1e-6 * max(abs(pp[a]), 1.0)
        pp[a] += h; pm[a] -= h
        fd = (_solve(d, mesh, pp, act, nsteps)[2] - _solve(d, mesh, pm, act, nsteps)[2]) / (2 * h)
        assert abs(g[k] - fd) < 2e-6 * max(abs(fd), np.abs(g).max() * 1e-3), (a, g[k], fd)


def test_fd_drop_like_the_reference_gradient_check():
    """The reference's regression metric: log10(max err / min err) over 13 FD steps of the
    directional derivative (src/main_inverse.cpp:127-140); its golden for the 2-D small_J2 adjoint
    deck is 7.74 +- 10% (golden.json fd_drops).  A correct gradient shows a drop of many decades."""
    d = G["decks"]["notch2D_small_J2"]
    mesh = load_mesh(d["mesh"])
    names = PARAM_NAMES[d["local_type"]]
    active = ["E", "nu", "K", "Y"]
    act = [names.index(a) for a in active]
    nsteps = 8
    o, p, J0 = _solve(d, mesh, d["params"], act, nsteps)
    g = Adjoint(p, abs_tol=1e-14, rel_tol=1e-12).gradient([list(range(4))], 4)
    # direction in physical parameters, scaled like the canonical [-1,1] variables
    span = np.array([abs(d["params"][a]) * 0.1 + 1e-3 for a in active])
    ddir = 0.1 * span
    errs = []
    for k in range(0, 9):
        eps = 10.0 ** (-k)
        pp = dict(d["params"])
        for a, dv in zip(active, ddir):
            pp[a] += eps * dv
        Jp = _solve(d, mesh, pp, act, nsteps)[2]
        errs.append(abs((Jp - J0) / eps - g @ ddir))
    errs = np.array(errs)
    drop = np.log10(errs.max() / errs.min())
    assert drop > 4.0, (drop, errs)
