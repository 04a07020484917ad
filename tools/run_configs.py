#!/usr/bin/env python
"""BASELINE.json configs[2] and [3] at full size (bounded number of steps): timings for
profiles/README.md.  cfg3 = Hill anisotropic plasticity, 3-D 1M tets, cyclic load, adjoint gradient
over 8 parameters; cfg4 = virtual-fields objective on a 2-D plane-stress 4M-triangle mesh."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from calibr8_b200 import meshgen
from calibr8_b200.capi import Context, HostProblem
from calibr8_b200.vfm import vfm_objective, vfm_forward, vfm_adjoint

which = sys.argv[1] if len(sys.argv) > 1 else "both"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 4

if which in ("cfg3", "both"):
    HILL = dict(E=1000., nu=.25, Y=2., R00=1., R11=.9, R22=1.1, R01=1., R02=.95, R12=1.05, S=10., D=2.)
    mesh = meshgen.box_tets(int(os.environ.get("NCELLS", "56")), notch_radius=0.2)
    ctx = Context(0); ctx.set_mesh(3, mesh.conn, mesh.coords)
    ctx.set_model("mechanics", "small_hill", HILL, max_iters=500, abs_tol=1e-12, rel_tol=1e-12)
    hp = HostProblem(ctx); hp.set_time(steps, 1.0)
    for r, e, s, v in [(0, 0, "xmin", "0.0"), (0, 1, "ymin", "0.0"), (0, 2, "zmin", "0.0"),
                       # triangular cyclic path, amplitude 0.004, period 12.5 steps (SURVEY 8(d) cfg 3)
                       (0, 1, "ymax", "0.004 * (2/3.14159265358979) * asin(sin(2*3.14159265358979*t/12.5))" )]:
        hp.add_dbc(r, e, mesh.node_sets[s], v)
    hp.finalize_dbcs()
    hp.set_solver(int(os.environ.get("NEWTON", "40")), 1e-8, 1e-8, gmres_restart=100, gmres_max_iters=20000, linear_tol=float(os.environ.get("LINTOL", "1e-8")),
                  verbose=bool(os.environ.get("VERBOSE")))
    hp.set_qoi_avg_disp()
    if os.environ.get("LS_EVALS"):      # the deck's "line search: max evals" (src/line_search.hpp:33-49)
        hp.set_line_search(max_evals=int(os.environ["LS_EVALS"]), min_backtrack=float(os.environ.get("LS_BMIN", "0.5")))
    for rep in range(1 if os.environ.get("ONCE") else 2):
        torch.cuda.synchronize(); t0 = time.perf_counter(); J = hp.primal_solve()
        torch.cuda.synchronize(); t1 = time.perf_counter(); g = hp.adjoint_gradient()
        torch.cuda.synchronize(); t2 = time.perf_counter()
    st = hp.stats()
    print(f"cfg3 small_hill {mesh.n_elems} tets, {steps} cyclic steps: forward {1e3*(t1-t0)/steps:.1f} ms/step, "
          f"adjoint {1e3*(t2-t1)/steps:.1f} ms/step; J={J:.10e}; grad (8 active of 11) "
          f"{np.array2string(g[[2,3,4,6,7,8,9,10]], precision=4)}; totals {st}", flush=True)
    hp.close(); ctx.close()

if which in ("cfg4", "both"):
    HILL2D = dict(E=1000., nu=.25, Y=2., S=10., D=50., R00=1., R11=1., R22=1., R01=1.)
    mesh = meshgen.square_tris(1414, notch_radius=0.0)
    ctx = Context(0); ctx.set_mesh(2, mesh.conn, mesh.coords)
    ctx.set_model("mechanics_plane_stress", "small_hill_plane_stress", HILL2D, max_iters=20, abs_tol=1e-12, rel_tol=1e-12)
    hp = HostProblem(ctx); hp.set_time(steps, 1.0); hp.finalize_dbcs(); hp.set_qoi_avg_disp()
    X, Y = mesh.coords[:, 0], mesh.coords[:, 1]
    w = np.stack([np.cos(np.pi * (Y - 0.5)) * X, Y * (2 * (Y - 0.5) + 1) / 2], axis=1)
    # synthetic full-field "measured" displacement: uniaxial stretch + a smooth perturbation, growing in time
    base = np.stack([-0.25 * 0.004 * X + 2e-4 * np.sin(3 * Y) * X, 0.004 * Y + 2e-4 * np.sin(2 * X) * Y], axis=1)
    measured = np.stack([base * (s / steps) * 2.0 for s in range(1, steps + 1)])
    loads = np.linspace(0.5, 2.0, steps)
    for mode in ("forward", "adjoint"):
        for rep in range(2):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            J, g = vfm_objective(hp, mode, measured, w, loads, obj_scale_factor=1e2)
            torch.cuda.synchronize(); t1 = time.perf_counter()
        print(f"cfg4 VFM {mode:8s} {mesh.n_elems} tris, {steps} steps: {1e3*(t1-t0)/steps:.1f} ms/step "
              f"({mesh.n_elems*steps/(t1-t0)/1e6:.1f} M QP-steps/s incl. host copies of the measured fields); "
              f"J={J:.10e} grad(Y,S,D)={g[[2,3,4]]}", flush=True)
    hp.close(); ctx.close()
