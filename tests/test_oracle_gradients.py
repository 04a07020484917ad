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
    # finite-strain Hill: the adjoint runs through d/dx and d/dx_prev of the AD'd polar-rotation iteration
    ("notch2D_hypo_J2_plane_stress", ["E", "Y", "S", "R11"], 4),
    ("notch2D_hypo_J2_plane_strain", ["nu", "Y", "D", "R01"], 4),
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


def _facets(mesh, coord_idx, value, tol=1e-12):
    """[n_elems][3] local vertex ids of the element facet on a coordinate plane (-1: none)"""
    on = np.abs(mesh.coords[mesh.conn][:, :, coord_idx] - value) < tol
    fac = np.full((mesh.n_elems, 3), -1, dtype=np.int32)
    for e in np.nonzero(on.sum(axis=1) == mesh.dim)[0]:
        fac[e, : mesh.dim] = np.nonzero(on[e])[0]
    return fac


@pytest.mark.parametrize("kind,dim", [("reaction", 3), ("reaction_torque", 3), ("load", 3), ("load", 2),
                                      ("surface", 3)])
def test_mismatch_qois_adjoint_gradient_matches_finite_differences(kind, dim):
    """The reaction / load / surface mismatch QoIs (src/reaction_mismatch.cpp, load_mismatch.cpp,
    surface_mismatch.cpp) restated in oracle/qoi.hpp: adjoint gradient vs central differences of the
    objective (the reference's own acceptance method for its derivative paths)."""
    from calibr8_b200 import meshgen
    if dim == 3:
        mesh = meshgen.box_tets(3, notch_radius=0.3)
        gtype, ltype = "mechanics", "small_J2"
        params = dict(E=1000., nu=.25, K=100., Y=2., cte=0., delta_T=0.)
        dbcs = [[0, 0, "xmin", "0.0"], [0, 1, "ymin", "0.0"], [0, 2, "zmin", "0.0"], [0, 1, "ymax", "0.0015 * t"]]
        active = ["E", "K", "Y"]
    else:
        mesh = meshgen.square_tris(6, notch_radius=0.3)
        gtype, ltype = "mechanics_plane_stress", "small_hill_plane_stress"
        params = dict(E=1000., nu=.25, Y=2., S=10., D=20., R00=1., R11=.9, R22=1.1, R01=.95)
        dbcs = [[0, 0, "xmin", "0.0"], [0, 1, "ymin", "0.0"], [0, 1, "ymax", "0.002 * t"]]
        active = ["E", "Y", "S"]
    names = PARAM_NAMES[ltype]
    act = [names.index(a) for a in active]
    nsteps = 3
    rng = np.random.RandomState(4)
    meas = [np.zeros((mesh.n_nodes, 3)) for _ in range(nsteps)]
    for m in meas:
        m[:, :dim] = 1e-3 * rng.uniform(-1, 1, size=(mesh.n_nodes, dim))
    load_meas = [0.3, 0.5, 0.6]

    def run(par, want_grad):
        o = Oracle(mesh.dim, mesh.conn, mesh.coords, global_type=gtype, local_type=ltype, params=[par],
                   max_iters=60, abs_tol=1e-13, rel_tol=1e-13, active=[act])
        if kind.startswith("reaction"):
            o.set_qoi_mismatch("reaction", coord_idx=1, coord_value=1.0, reaction_force_comp=1 if kind == "reaction" else 2,
                               compute_torque=(kind == "reaction_torque"))
        elif kind == "load":
            o.set_qoi_mismatch("load", facet=_facets(mesh, 1, 1.0), normal_2d=(0., 1.))
        else:
            o.set_qoi_mismatch("surface", facet=_facets(mesh, 2, 1.0))
        p = Primal(o, [Dbc(r, e, mesh.node_sets[s], v) for r, e, s, v in dbcs], nsteps, 1.0, max_iters=30,
                   abs_tol=1e-12, rel_tol=1e-12)
        setup = lambda step: o.qoi_set_step(1.0, float(nsteps), load_meas[step - 1], meas[step - 1])
        J = p.solve(setup)
        g = None
        if want_grad:
            g = Adjoint(p, max_iters=30, abs_tol=1e-14, rel_tol=1e-12).gradient([list(range(len(act)))], len(act), setup)
        return J, g, p

    J, g, p = run(params, True)
    assert J > 0 and p.xi[-1][:, -1].max() > 0, "the state must be plastic and the objective non-trivial"
    for k, a in enumerate(active):
        pp, pm = dict(params), dict(params)
        h = 1e-6 * max(abs(pp[a]), 1.0)
        pp[a] += h; pm[a] -= h
        fd = (run(pp, False)[0] - run(pm, False)[0]) / (2 * h)
        assert abs(g[k] - fd) < 5e-6 * max(abs(fd), np.abs(g).max() * 1e-3), (kind, a, g[k], fd)
