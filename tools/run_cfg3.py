#!/usr/bin/env python
"""BASELINE.json configs[2] at spec: 3-D Hill anisotropic plasticity (small_hill) on the 1M-tet notched
box, 50-step triangular cyclic load path u_y(ymax) = 0.004 tri(t / 12.5), `calibration` objective
(displacement mismatch on the zmax face + reaction at ymax), adjoint gradient over the 8 active
parameters Y, S, D, R00, R11, R01, R02, R12 (SURVEY.md 8(d) cfg 3).

    python tools/run_cfg3.py [--cells 56] [--steps 50]

Synthetic data: displacement history and plane loads of a forward run at the true parameters; the
objective and its gradient are then evaluated at perturbed starting parameters.  The global Newton uses the
deck option `line search: max evals: 8` (src/line_search.hpp:33-49): with the default 4 the reference's
Newton stagnates at the first load reversal on this mesh (profiles/README.md).
Prints one JSON line."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from calibr8_b200 import meshgen
from calibr8_b200.capi import Context, HostProblem, PARAM_NAMES

ap = argparse.ArgumentParser()
ap.add_argument("--cells", type=int, default=56)
ap.add_argument("--steps", type=int, default=50)
ap.add_argument("--ls-evals", type=int, default=8)
ap.add_argument("--newton", type=int, default=60)
a = ap.parse_args()
TRUTH = dict(E=1000., nu=.25, Y=2., R00=1., R11=.9, R22=1.1, R01=1., R02=.95, R12=1.05, S=10., D=2.)
START = dict(TRUTH, Y=2.2, S=8., D=2.5, R00=1.05, R11=.95, R01=1.05, R02=1., R12=1.)
ACTIVE = ["Y", "S", "D", "R00", "R11", "R01", "R02", "R12"]
N = a.steps
mesh = meshgen.box_tets(a.cells, notch_radius=0.2)
on = np.abs(mesh.coords[mesh.conn][:, :, 2] - 1.0) < 1e-12          # zmax facets
fac = np.full((mesh.n_elems, 3), -1, dtype=np.int8)
idx = np.nonzero(on.sum(axis=1) == 3)[0]
fac[idx] = np.stack([np.nonzero(r)[0] for r in on[idx]])
X = mesh.coords[mesh.conn[idx]]
fa, fb, fc = (X[np.arange(len(idx)), fac[idx, k]] for k in range(3))
area = float(0.5 * np.linalg.norm(np.cross(fb - fa, fc - fa), axis=1).sum())
ctx = Context(0); ctx.set_mesh(3, mesh.conn, mesh.coords)
ctx.set_model("mechanics", "small_hill", TRUTH, max_iters=500, abs_tol=1e-12, rel_tol=1e-12)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); ctx.set_stream(stream.cuda_stream)
hp = HostProblem(ctx); hp.set_time(N, 1.0)
PATH = "0.004 * (2/3.14159265358979) * asin(sin(2*3.14159265358979*t/12.5))"
for r, e, s, v in [(0, 0, "xmin", "0.0"), (0, 1, "ymin", "0.0"), (0, 2, "zmin", "0.0"), (0, 1, "ymax", PATH)]:
    hp.add_dbc(r, e, mesh.node_sets[s], v)
hp.finalize_dbcs()
hp.set_solver(a.newton, 1e-8, 1e-8, gmres_restart=100, gmres_max_iters=20000, linear_tol=1e-8)
hp.set_line_search(max_evals=a.ls_evals)
qoi = dict(balance_factor=1e2, coord_idx=1, coord_value=1.0, reaction_force_comp=1, weights=(1e8, 1e8, 1e8))
# ---- synthetic data at the true parameters ("load out file" mode: mismatch against zero)
hp.set_qoi_calibration(measured=np.zeros((N, mesh.n_nodes, 3)), load_data=np.zeros(N), area=area, facet=fac, **qoi)
torch.cuda.synchronize(); t0 = time.perf_counter()
hp.primal_solve()
torch.cuda.synchronize(); t_truth = time.perf_counter() - t0
loads = hp.loads()
measured = np.stack([hp.get_step(s)[0][0].reshape(-1, 3) for s in range(1, N + 1)])
alpha_max = float(hp.get_step(N)[1][:, 6].max())
# ---- objective + adjoint gradient at the starting parameters
ctx.set_params(START)
hp.set_qoi_calibration(measured=measured, load_data=loads, area=area, facet=fac, **qoi)
hp.profile(True); p0 = hp.profile(True); s0 = hp.stats()
torch.cuda.synchronize(); t0 = time.perf_counter()
J = hp.primal_solve()
torch.cuda.synchronize(); t1 = time.perf_counter()
g = hp.adjoint_gradient()
torch.cuda.synchronize(); t2 = time.perf_counter()
s1 = hp.stats(); p1 = hp.profile(True)
names = PARAM_NAMES["small_hill"]
out = dict(config="BASELINE configs[2]: 3-D small_hill, cyclic 50-step path, calibration objective, 8 active parameters",
           n_elems=mesh.n_elems, n_nodes=mesh.n_nodes, load_steps=N, zmax_facets=int(len(idx)),
           forward_ms_per_load_step=(t1 - t0) / N * 1e3, adjoint_ms_per_load_step=(t2 - t1) / N * 1e3,
           ms_per_load_step=(t2 - t0) / N * 1e3, truth_forward_s=t_truth,
           assemblies=s1["assemblies"] - s0["assemblies"], krylov_iterations=s1["linear_iters"] - s0["linear_iters"],
           objective=J, gradient={n: float(g[names.index(n)]) for n in ACTIVE}, alpha_max_truth=alpha_max,
           load_min_max=[float(loads.min()), float(loads.max())],
           phase_seconds={k: p1[k] - p0[k] for k in p1}, line_search_max_evals=a.ls_evals,
           preconditioner=ctx.preconditioner_info())
print(json.dumps(out), flush=True)
hp.close(); ctx.close()
