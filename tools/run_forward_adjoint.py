#!/usr/bin/env python
"""Forward load-step solve + adjoint gradient on a synthetic notched box (BASELINE configs[1]/[4]
shape) through the C++ host solvers: wall time per load step, assemblies, Krylov iterations."""
import argparse, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from calibr8_b200 import meshgen
from calibr8_b200.capi import Context, HostProblem

ap = argparse.ArgumentParser()
ap.add_argument("--cells", type=int, default=16)
ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--model", default="small_J2")
ap.add_argument("--disp", type=float, default=0.001, help="u_y(ymax) per step")
ap.add_argument("--restart", type=int, default=100)
ap.add_argument("--lin-tol", type=float, default=1e-8)
ap.add_argument("--no-adjoint", action="store_true")
ap.add_argument("--verbose", action="store_true")
ap.add_argument("--pc", default="amg")
ap.add_argument("--profile", action="store_true")
ap.add_argument("--nu", type=int, default=2)
ap.add_argument("--omega", type=float, default=0.7)
ap.add_argument("--oc", type=float, default=1.6)
ap.add_argument("--agg", type=int, default=8)
ap.add_argument("--cagg", type=int, default=8)
ap.add_argument("--cmax", type=int, default=40)
ap.add_argument("--cnu", type=int, default=0)
a = ap.parse_args()

PAR = {"small_J2": dict(E=1000., nu=.25, K=100., Y=2., cte=0., delta_T=0.),
       "hyper_J2": dict(E=1000., nu=.25, Y=10., S=0., D=0., A=0., n=0., K=100.),
       "small_hill": dict(E=1000., nu=.25, Y=2., R00=1., R11=.9, R22=1.1, R01=1., R02=.95, R12=1.05, S=10., D=2.),
       "elastic": dict(E=1000., nu=.25, cte=0., delta_T=0.)}
t0 = time.time()
mesh = meshgen.box_tets(a.cells, notch_radius=0.2)
ctx = Context(0)
ctx.set_mesh(3, mesh.conn, mesh.coords)
ctx.set_model("mechanics", a.model, PAR[a.model], max_iters=500, abs_tol=1e-12, rel_tol=1e-12)
ctx.set_preconditioner(a.pc, nu_pre=a.nu, nu_post=a.nu, omega=a.omega, over_correction=a.oc, max_aggregate_size=a.agg, coarse_aggregate_size=a.cagg, coarsest_max_nodes=a.cmax, coarse_nu=a.cnu)
hp = HostProblem(ctx)
hp.set_time(a.steps, 1.0)
hp.add_dbc(0, 0, mesh.node_sets["xmin"], "0.0")
hp.add_dbc(0, 1, mesh.node_sets["ymin"], "0.0")
hp.add_dbc(0, 2, mesh.node_sets["zmin"], "0.0")
hp.add_dbc(0, 1, mesh.node_sets["ymax"], f"{a.disp} * t")
hp.finalize_dbcs()
hp.set_solver(15, 1e-8, 1e-8, gmres_restart=a.restart, gmres_max_iters=20000, linear_tol=a.lin_tol,
              verbose=a.verbose)
hp.set_qoi_avg_disp()
print(f"mesh {mesh.n_elems} tets {mesh.n_nodes} nodes, setup {time.time()-t0:.1f}s", flush=True)
if a.profile: hp.profile(True)
torch.cuda.synchronize(); t1 = time.time()
J = hp.primal_solve()
torch.cuda.synchronize(); t2 = time.time()
s1 = hp.stats()
print("preconditioner", a.pc, ctx.preconditioner_info())
print(f"forward: J={J:.12e} {t2-t1:.2f}s = {(t2-t1)/a.steps*1e3:.1f} ms/step  assemblies {s1['assemblies']} "
      f"krylov its {s1['linear_iters']}", flush=True)
if a.profile: print("profile forward", {k: round(v, 4) for k, v in hp.profile(True).items()})
if not a.no_adjoint:
    g = hp.adjoint_gradient()
    torch.cuda.synchronize(); t3 = time.time()
    s2 = hp.stats()
    print(f"adjoint: {t3-t2:.2f}s = {(t3-t2)/a.steps*1e3:.1f} ms/step krylov its {s2['linear_iters']-s1['linear_iters']} grad {g}", flush=True)
    print(f"forward+adjoint gradient wall-time/load step: {(t3-t1)/a.steps*1e3:.1f} ms")
if not a.no_adjoint:
    for rep in range(3):
        t4 = time.time(); g = hp.adjoint_gradient(); torch.cuda.synchronize()
        print(f"adjoint (call {rep+2}): {time.time()-t4:.2f}s")
    t4 = time.time(); J = hp.primal_solve(); torch.cuda.synchronize()
    print(f"forward (call 2): {time.time()-t4:.2f}s")
if a.profile: print("profile total", {k: round(v, 4) for k, v in hp.profile(True).items()})
