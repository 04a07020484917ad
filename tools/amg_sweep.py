#!/usr/bin/env python
"""Time-to-solution sweep of the multigrid options on the forward solve of the bench problem (hyper-J2 notch,
1.02 M tets): LOAD_STEPS load steps of `Primal::solve_at_step` per option set, the second of two runs timed
(the first allocates / builds the hierarchy).  Prints wall time, assemblies and Krylov iterations per set.
Not a bench value (see bench.py)."""
import itertools, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from calibr8_b200.capi import Context, HostProblem

STEPS = int(os.environ.get("LOAD_STEPS", "12"))
mesh = bench.workload_mesh(int(os.environ.get("NCELLS", "56")))
ctx = Context(0)
ctx.set_mesh(3, mesh.conn, mesh.coords)
ctx.set_model("mechanics", "hyper_J2", bench.PARAMS, **bench.LOCAL)
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
ctx.set_stream(stream.cuda_stream)
ns = mesh.node_sets
hp = HostProblem(ctx)
hp.add_dbc(0, 0, ns["xmin"], "0.0"); hp.add_dbc(0, 1, ns["ymin"], "0.0"); hp.add_dbc(0, 2, ns["zmin"], "0.0")
hp.add_dbc(0, 1, ns["ymax"], "%g * t" % (0.02 / STEPS))      # the same 2 % total stretch in STEPS steps
hp.finalize_dbcs()
hp.set_solver(15, 1e-8, 1e-8, gmres_restart=100, gmres_max_iters=int(os.environ.get("GMRES_MAX", "3000")), linear_tol=bench.LINEAR_TOL)
hp.set_qoi_avg_disp()
hp.set_time(STEPS, 1.0)

BASE = dict(nu_pre=2, nu_post=2, omega=0.8, over_correction=1.6, coarsest_max_nodes=40, max_aggregate_size=8,
            coarse_aggregate_size=12, coarse_nu=0)
sets = [("base", {})]
for a, b in [(1, 1), (1, 2), (2, 1), (3, 3), (2, 3), (3, 2)]:
    sets.append((f"V({a},{b})", dict(nu_pre=a, nu_post=b)))
for w in (0.6, 0.8, 0.9):
    sets.append((f"omega {w}", dict(omega=w)))
for oc in (1.3, 1.45, 1.8, 2.0):
    sets.append((f"over {oc}", dict(over_correction=oc)))
for cn in (1, 3):
    sets.append((f"coarse_nu {cn}", dict(coarse_nu=cn)))
for m in (6, 12, 16):
    sets.append((f"agg {m}", dict(max_aggregate_size=m)))
for m in (12, 16, 0):
    sets.append((f"coarse agg {m}", dict(coarse_aggregate_size=m)))
for m in (100, 600):
    sets.append((f"coarsest {m}", dict(coarsest_max_nodes=m)))
only = os.environ.get("ONLY")
if only:
    sets = [s for s in sets if s[0] in only.split(";")]
extra = os.environ.get("EXTRA")          # JSON list of [name, {opts}]
if extra:
    sets += [(n, o) for n, o in json.loads(extra)]

for name, kw in sets:
    opts = dict(BASE, **kw)
    try:
        ctx.set_preconditioner("amg", **opts)
        res = None
        for rep in range(2):
            s0 = hp.stats()
            torch.cuda.synchronize(); t0 = time.perf_counter()
            J = hp.primal_solve()
            torch.cuda.synchronize(); t1 = time.perf_counter()
            s1 = hp.stats()
            res = (t1 - t0, s1["assemblies"] - s0["assemblies"], s1["linear_iters"] - s0["linear_iters"], J)
        info = ctx.preconditioner_info()
        print(f"{name:16s} {res[0]*1e3/STEPS:8.1f} ms/load step  assemblies {res[1]:4d}  krylov {res[2]:6d}  "
              f"({res[2]/max(res[1],1):5.1f}/solve)  levels {info['levels']} oc {info['operator_complexity']:.2f}  J {res[3]:.10e}",
              flush=True)
    except Exception as e:
        print(f"{name:16s} FAILED: {e}", flush=True)
