#!/usr/bin/env python
"""ms per GMRES iteration on the bench matrix (fixed iteration count) for each preconditioner."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from calibr8_b200.capi import Context
mesh = bench.workload_mesh(int(os.environ.get("NCELLS", "56")))
(u1, p1), (u2, p2) = bench.workload_fields(mesh)
ctx = Context(0); ctx.set_mesh(3, mesh.conn, mesh.coords)
ctx.set_model("mechanics", "hyper_J2", bench.PARAMS, **bench.LOCAL)
x, xp, x0 = ctx.alloc("x"), ctx.alloc("x"), ctx.alloc("x")
xi0, xip, xi = ctx.alloc("xi"), ctx.alloc("xi"), ctx.alloc("xi")
A, b = ctx.alloc("A"), ctx.alloc("b")
ctx.pack_x(u2, p2, x); ctx.pack_x(u1, p1, xp); ctx.init_xi(xi0); ctx.init_xi(xip)
ctx.forward_jacobian(xp, x0, xi0, xip, None, b); xi.copy_(xip); torch.cuda.synchronize()
ctx.forward_jacobian(x, xp, xip, xi, A, b)
rhs = torch.randn(ctx.n_dofs, dtype=torch.float64, device="cuda"); sol = torch.zeros_like(rhs)
CASES = [("block_jacobi", {}), ("amg", {}), ("amg", dict(nu_pre=1, nu_post=1))]
if os.environ.get("ONLY_AMG"): CASES = [("amg", {})]
for pc, kw in CASES:
    ctx.set_preconditioner(pc, **kw)
    for restart in ((25,) if os.environ.get("ONLY_AMG") else (25, 100)):
        for rep in range(2):   # the first call builds the hierarchy and captures the iteration graphs
            sol.zero_(); torch.cuda.synchronize(); t0 = time.perf_counter()
            info = ctx.gmres(A, rhs, sol, restart=restart, max_iters=int(os.environ.get("ITS", "200")), rel_tol=1e-30)
            ctx.synchronize(); t1 = time.perf_counter()
        print(f"{pc:13s} {kw} restart {restart:3d}: {1e3*(t1-t0)/info['iters']:.3f} ms/it ({info['iters']} its, resid {info['resid']:.6e}, "
              f"|sol| {float(sol.norm()):.12e})", flush=True)
