"""Process-boundary front end: the reference's `primal` and
`objective` executables on top of the B200 hot path.

    python -m calibr8_b200.cli primal    deck.yaml
    python -m calibr8_b200.cli objective deck.yaml true|false [label]
    python -m calibr8_b200.cli inverse   deck.yaml          (src/main_inverse.cpp: the in-process optimiser)

Same YAML deck schema as the reference (src/main_primal.cpp, src/main_objective.cpp:512-560):
`problem`, `discretization`, `residuals`, `dirichlet bcs: expression`, `quantity of interest`,
`traction bcs`, `inverse`, `virtual fields`; QoI types `average displacement`, `calibration`,
`reaction mismatch`, `load mismatch`, `surface mismatch`.  Same text outputs, one `%.17e` per line:
`objective_value[_label].txt`, `objective_gradient[_label].txt` (src/main_objective.cpp:199-219),
the QoI side file `load out file` (src/reaction_mismatch.cpp:137-147).  The unmodified Python
drivers of the reference (py/calibr8/util/driver_support.py) only need the executable names mapped.

Differences, stated: the mesh may be a flat `.npz` (calibr8_b200.meshio) besides `.smb`+`.dmg`+assoc;
`write synthetic: true` stores the measured displacement history as `<name>_synthetic/measured.npz`
next to a copy of the mesh instead of SMB field tags; `linear algebra` is read for the convergence
tolerance only (the solver is the library's AMG-GMRES).  No CPU path: needs the CUDA library.
"""
from __future__ import annotations

import os
import sys

import numpy as np

from . import meshio
from .capi import PARAM_NAMES, Context, HostProblem, eval_expr


def load_deck(path):
    import yaml
    with open(path) as f:
        d = yaml.safe_load(f)
    assert isinstance(d, dict) and len(d) == 1, "a deck has one top-level sublist"
    name, body = next(iter(d.items()))
    return name, body


def _resolve(base, p):
    return p if os.path.isabs(p) else os.path.normpath(os.path.join(base, p))


def load_mesh(disc, base):
    """-> (Mesh, measured [num_steps, n_nodes, dim] or None)"""
    mf = _resolve(base, disc["mesh file"])
    if os.path.isdir(mf) or mf.endswith("/"):
        z = np.load(os.path.join(mf, "measured.npz"))
        return meshio.load_npz(os.path.join(mf, "mesh.npz")), z["measured"]
    if mf.endswith(".npz"):
        return meshio.load_npz(mf), None
    return meshio.load_calibr8_mesh(mf, _resolve(base, disc["geom file"]),
                                    _resolve(base, disc["assoc file"])), None


def _materials(local, mesh, override=None):
    names = PARAM_NAMES[local["type"]]
    out = []
    for es in mesh.elem_set_names:
        m = dict(local["materials"][es])
        if override and es in override:
            m.update(override[es])
        out.append({k: float(m[k]) for k in names})
    return out


def build(deck, base, override=None, device=0):
    disc = deck["discretization"]
    mesh, measured = load_mesh(disc, base)
    res = deck["residuals"]
    g, l = res["global residual"], res["local residual"]
    ctx = Context(device)
    ctx.set_mesh(mesh.dim, mesh.conn, mesh.coords, mesh.elem_set, len(mesh.elem_set_names))
    gtype = g["type"]
    if gtype == "mechanics" and not bool(g.get("mixed formulation", True)):
        # displacement-only mechanics (src/mechanics.cpp:18-54): usable with a local residual whose cauchy()
        # is the full stress -- small_hill_plane_stress -- where it IS mechanics_plane_stress with unit
        # thickness (host/residuals.hpp: create_global_residual(..., mixed = false))
        if l["type"] != "small_hill_plane_stress":
            raise SystemExit("mixed formulation: false needs a local residual whose cauchy() does not read the "
                             "pressure field (small_hill_plane_stress)")
        gtype, g = "mechanics_plane_stress", dict(g, thickness=1.0)
    ctx.set_model(gtype, l["type"], _materials(l, mesh, override),
                  max_iters=int(l.get("nonlinear max iters", 0)),
                  abs_tol=float(l.get("nonlinear absolute tol", 0.0)),
                  rel_tol=float(l.get("nonlinear relative tol", 0.0)),
                  stab_mult=float(g.get("stabilization multiplier", 1.0)),
                  thickness=float(g.get("thickness", 1.0)))
    hp = HostProblem(ctx)
    nsteps = int(disc["num steps"])
    hp.set_time(nsteps, float(disc.get("step size", 1.0)))
    for _, bc in (deck.get("dirichlet bcs", {}).get("expression", {}) or {}).items():
        hp.add_dbc(int(bc[0]), int(bc[1]), mesh.node_sets[str(bc[2])], str(bc[3]))
    hp.finalize_dbcs()
    # traction bcs, "bc name: [resid_idx, side_set_name, x-val, y-val(, z-val)]" (src/tbcs.cpp:28-35)
    for _, bc in (deck.get("traction bcs") or {}).items():
        hp.add_tbc(int(bc[0]), mesh.side_sets[str(bc[1])], [str(v) for v in bc[2:2 + mesh.dim]])
    lin_tol = 1e-10
    try:
        la = deck["linear algebra"]
        st = la["Linear Solver Types"]["Belos"]["Solver Types"]
        lin_tol = float(next(iter(st.values()))["Convergence Tolerance"])
    except Exception:
        pass
    hp.set_solver(int(g.get("nonlinear max iters", 15)), float(g.get("nonlinear absolute tol", 1e-8)),
                  float(g.get("nonlinear relative tol", 1e-8)), gmres_restart=200, gmres_max_iters=20000,
                  linear_tol=max(lin_tol, 1e-13), verbose=bool(g.get("print convergence", False)))
    ls = g.get("line search") or {}
    hp.set_line_search(float(ls.get("sufficient decrease", 1e-4)), float(ls.get("min backtrack factor", 0.5)),
                       float(ls.get("max backtrack factor", 0.9)), int(ls.get("max evals", 4)))
    return ctx, hp, mesh, measured


def _area(mesh):
    X = mesh.coords[mesh.conn]
    return float(0.5 * np.abs((X[:, 1, 0] - X[:, 0, 0]) * (X[:, 2, 1] - X[:, 0, 1]) -
                              (X[:, 1, 1] - X[:, 0, 1]) * (X[:, 2, 0] - X[:, 0, 0])).sum())


def set_qoi(hp, mesh, q, base, nsteps, measured):
    """create_qoi (src/qoi.cpp:261-289) for the QoIs on the path.  A QoI with a `load out file` runs in
    write mode (mismatch against zero); the file's lines come back through hp.loads()."""
    t = q["type"]
    loads = None
    if q.get("load input file"):
        loads = np.loadtxt(_resolve(base, q["load input file"])).reshape(-1)[:nsteps]
    if t == "average displacement":
        hp.set_qoi_avg_disp()
    elif t == "calibration":
        assert measured is not None, "the calibration objective needs a *_synthetic/ mesh directory"
        w = [float(v) for v in q.get("displacement weights", [1.0] * mesh.dim)]
        facet = meshio.side_set_facets(mesh, str(q["side set"])) if mesh.dim == 3 else None
        if mesh.dim == 3:
            X = mesh.coords[mesh.conn]
            on = facet[:, 0] >= 0
            idx = np.nonzero(on)[0]
            a_, b_, c_ = (X[idx, facet[idx, k]] for k in range(3))
            area = float(0.5 * np.linalg.norm(np.cross(b_ - a_, c_ - a_), axis=1).sum())
        else:
            area = _area(mesh)
        hp.set_qoi_calibration(balance_factor=float(q.get("balance factor", 1.0)),
                               coord_idx=int(q["coordinate index"]), coord_value=float(q["coordinate value"]),
                               coord_tol=float(q.get("coordinate tolerance", 1e-12)),
                               reaction_force_comp=int(q["reaction force component"]), weights=w,
                               measured=measured, load_data=loads if loads is not None else np.zeros(nsteps),
                               area=area, facet=facet)
    elif t == "reaction mismatch":
        hp.set_qoi_mismatch(t, coord_idx=int(q["coordinate index"]), coord_value=float(q["coordinate value"]),
                            coord_tol=float(q.get("coordinate tolerance", 1e-12)),
                            reaction_force_comp=int(q["reaction force component"]),
                            compute_torque=bool(q.get("compute torque", False)), load_data=loads)
    elif t == "load mismatch":
        hp.set_qoi_mismatch(t, facet=meshio.side_set_facets(mesh, str(q["side set"])),
                            normal_2d=q.get("2D surface normal"), load_data=loads)
    elif t == "surface mismatch":
        assert measured is not None, "the surface mismatch objective needs a *_synthetic/ mesh directory"
        hp.set_qoi_mismatch(t, facet=meshio.side_set_facets(mesh, str(q["side set"])), measured=measured)
    else:
        raise SystemExit(f"quantity of interest '{t}' is not on this path")


def _write_lines(path, values):
    with open(path, "w") as f:
        for v in values:
            f.write("%.17e\n" % v)


def run_primal(deck_path):
    name, deck = load_deck(deck_path)
    base = os.path.dirname(os.path.abspath(deck_path))
    ctx, hp, mesh, _ = build(deck, base)
    q = deck.get("quantity of interest", {"type": "average displacement"})
    nsteps = int(deck["discretization"]["num steps"])
    set_qoi(hp, mesh, q, base, nsteps, None)
    J = hp.primal_solve()
    print("QoI (%s): %.17e" % (q["type"], J))
    if q.get("load out file"):      # written by the QoI's preprocess pass (src/reaction_mismatch.cpp:137-147)
        _write_lines(_resolve(os.getcwd(), q["load out file"]), hp.loads())
    if deck.get("problem", {}).get("write synthetic", False):
        out = os.path.join(os.getcwd(), deck["problem"]["name"] + "_synthetic")
        os.makedirs(out, exist_ok=True)
        meshio.save_npz(mesh, os.path.join(out, "mesh.npz"))
        meas = np.stack([hp.get_step(s)[0][0].reshape(-1, mesh.dim) for s in range(1, nsteps + 1)])
        np.savez(os.path.join(out, "measured.npz"), measured=meas)
    hp.close(); ctx.close()
    return J


def _active(deck, mesh, local_type):
    """setup_opt_params, src/main_objective.cpp: active (elem set, parameter) pairs in order"""
    names = PARAM_NAMES[local_type]
    inv = deck["inverse"]["materials"]
    act = []
    for es_i, es in enumerate(mesh.elem_set_names):
        for k, nm in enumerate(names):
            if nm in (inv.get(es) or {}):
                act.append((es_i, k))
    return act


def run_objective(deck_path, gradient, label=""):
    name, deck = load_deck(deck_path)
    base = os.path.dirname(os.path.abspath(deck_path))
    ctx, hp, mesh, measured = build(deck, base)
    inv = deck["inverse"]
    otype = str(inv["objective type"]).lower()
    nsteps = int(deck["discretization"]["num steps"])
    ltype = deck["residuals"]["local residual"]["type"]
    act = _active(deck, mesh, ltype)
    assert all(e == 0 for e, _ in act) or ctx.n_es > 1
    if otype in ("adjoint", "pdeco", "femu"):
        set_qoi(hp, mesh, deck["quantity of interest"], base, nsteps, measured)
        J = hp.primal_solve()
        g = hp.adjoint_gradient() if gradient else None
    elif otype in ("vfm", "fs_vfm", "adjoint_vfm"):
        from .vfm import vfm_objective
        assert measured is not None, "the VFM objective needs a *_synthetic/ mesh directory"
        vf = deck["virtual fields"]
        w = np.stack([eval_expr(vf["w_x"], mesh.coords), eval_expr(vf["w_y"], mesh.coords)], axis=1)
        loads = np.loadtxt(_resolve(base, inv["load input file"])).reshape(-1)[:nsteps]
        # load.dat is the external virtual power directly: the shipped virtual field is (0, 1) on the
        # loaded edge; "internal power scale factor" multiplies the internal power like the thickness
        scale = float(inv.get("internal power scale factor", 1.0))
        J, g = vfm_objective(hp, "forward" if otype in ("vfm", "fs_vfm") else "adjoint", measured, w,
                             loads, obj_scale_factor=float(inv.get("objective scale factor", 1.0)),
                             thickness=float(inv.get("thickness", 1.0)) * scale)
    else:
        raise SystemExit(f"objective type '{otype}' is not implemented")
    lab = ("_" + label) if label else ""
    _write_lines("objective_value%s.txt" % lab, [J])
    if gradient:
        gall = np.asarray(g).reshape(-1)
        _write_lines("objective_gradient%s.txt" % lab, [gall[k] for _, k in act])
    hp.close(); ctx.close()
    return J, (None if not gradient else np.array([np.asarray(g).reshape(-1)[k] for _, k in act]))


def _setup_objective(deck, base):
    """-> (ctx, hp, Objective, names of the active parameters, start point in canonical coordinates)"""
    from .capi import Objective
    ctx, hp, mesh, measured = build(deck, base)
    inv = deck["inverse"]
    otype = str(inv["objective type"]).lower()
    nsteps = int(deck["discretization"]["num steps"])
    ltype = deck["residuals"]["local residual"]["type"]
    act = _active(deck, mesh, ltype)
    assert all(e == 0 for e, _ in act), "one element set"
    names = [PARAM_NAMES[ltype][k] for _, k in act]
    bounds = inv["materials"][mesh.elem_set_names[0]]
    lo = [float(bounds[n][0]) for n in names]
    hi = [float(bounds[n][1]) for n in names]
    kw = {}
    if otype in ("adjoint", "pdeco", "femu"):
        set_qoi(hp, mesh, deck["quantity of interest"], base, nsteps, measured)
        kind = "adjoint"
    else:
        vf = deck["virtual fields"]
        kw = dict(measured=measured,
                  w=np.stack([eval_expr(vf["w_x"], mesh.coords), eval_expr(vf["w_y"], mesh.coords)], axis=1),
                  loads=np.loadtxt(_resolve(base, inv["load input file"])).reshape(-1)[:nsteps],
                  obj_scale_factor=float(inv.get("objective scale factor", 1.0)),
                  thickness=float(inv.get("thickness", 1.0)) * float(inv.get("internal power scale factor", 1.0)))
        kind = "adjoint_vfm" if otype == "adjoint_vfm" else "fs_vfm"
    obj = Objective(hp, kind, [k for _, k in act], lo, hi, **kw)
    p0 = obj.to_canonical(obj.active_params())
    return ctx, hp, obj, names, p0


def run_inverse(deck_path):
    """The `inverse` executable (src/main_inverse.cpp:21-162): bound-constrained quasi-Newton on the
    canonical box [-1, 1]^n from the deck's material values; writes the iterate log to ROL_out.txt
    and the calibrated physical parameters to inverse_result.txt (one `%.17e` per line)."""
    from scipy.optimize import minimize
    name, deck = load_deck(deck_path)
    base = os.path.dirname(os.path.abspath(deck_path))
    ctx, hp, obj, names, p0 = _setup_objective(deck, base)
    inv = deck["inverse"]
    log = []

    def fun(p):
        J, g = obj.value(p), obj.gradient(p)
        log.append((J, float(np.abs(g).max()), obj.to_physical(p)))
        return J, g

    res = minimize(fun, p0, jac=True, method="L-BFGS-B", bounds=[(-1.0, 1.0)] * len(p0),
                   options=dict(maxiter=int(inv.get("iteration limit", 50)),
                                gtol=float(inv.get("gradient tolerance", 1e-8)), ftol=1e-30,
                                maxls=int(inv.get("max line search evals", 20))))
    phys = obj.to_physical(res.x)
    with open("ROL_out.txt", "w") as f:
        f.write("# eval  objective  |grad|_inf  " + "  ".join(names) + "\n")
        for k, (J, gn, pp) in enumerate(log):
            f.write("%4d  %.12e  %.6e  %s\n" % (k, J, gn, "  ".join("%.10e" % v for v in pp)))
    _write_lines("inverse_result.txt", phys)
    for n, v in zip(names, phys):
        print("%s = %.12e" % (n, v))
    obj.close(); hp.close(); ctx.close()
    return dict(zip(names, phys)), res.fun, len(log)


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    if len(argv) >= 2 and argv[0] == "primal":
        run_primal(argv[1])
    elif len(argv) >= 3 and argv[0] == "objective":
        run_objective(argv[1], argv[2] == "true", argv[3] if len(argv) > 3 else "")
    elif len(argv) >= 2 and argv[0] == "inverse":
        run_inverse(argv[1])
    else:
        raise SystemExit(__doc__)


if __name__ == "__main__":
    main()
