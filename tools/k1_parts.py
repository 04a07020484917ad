#!/usr/bin/env python
"""K1 time of every part of an N-way partition of the bench mesh, one part after the other on ONE GPU
(what each rank of `bench.py --gpus N` runs), with the element-count partition and with a cost-weighted
one (weight = 1 + BETA * branch of the previous state).  Shows how much of the strong-scaling loss of K1
is load imbalance from the plastic zone.  Not a bench value."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from calibr8_b200 import partition
from calibr8_b200.capi import Context

N = int(os.environ.get("PARTS", "8"))
BETAS = [float(v) for v in os.environ.get("BETAS", "0,0.5,1.0").split(",")]
mesh = bench.workload_mesh()
(u1, p1), (u2, p2) = bench.workload_fields(mesh)


def k1(conn, coords, f, reps=8):
    a1, q1, a2, q2 = f
    ctx = Context(0)
    ctx.set_mesh(3, conn, coords)
    ctx.set_model("mechanics", "hyper_J2", bench.PARAMS, **bench.LOCAL)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    x, xp, x0 = ctx.alloc("x"), ctx.alloc("x"), ctx.alloc("x")
    xi0, xip, xi = ctx.alloc("xi"), ctx.alloc("xi"), ctx.alloc("xi")
    A, b, path = ctx.alloc("A"), ctx.alloc("b"), ctx.alloc("path")
    ctx.pack_x(a2, q2, x); ctx.pack_x(a1, q1, xp)
    ctx.init_xi(xi0); ctx.init_xi(xip)
    assert ctx.forward_jacobian(xp, x0, xi0, xip, None, b, path) == 0
    path_prev = path.cpu().numpy().astype(np.int8).copy()
    ts = []
    for k in range(reps + 3):
        b.zero_(); xi.copy_(xip)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ctx.forward_jacobian(x, xp, xip, xi, A, b, path, check=False); e1.record()
        torch.cuda.synchronize()
        if k >= 3: ts.append(e0.elapsed_time(e1))
    pl = float(path.float().mean())
    n = ctx.n_elems
    ctx.close()
    return float(np.median(ts)), pl, n, path_prev


t1, pl1, n1, path_prev = k1(mesh.conn, mesh.coords, (u1, p1, u2, p2))
print(f"one part: {n1} elems, plastic {pl1:.3f} (previous state {path_prev.mean():.3f}), K1 {t1:.4f} ms", flush=True)
for beta in BETAS:
    w = None if beta == 0 else 1.0 + beta * path_prev.astype(np.float64)
    ep, parts = partition.partition_mesh(mesh, N, weights=w)
    rows = []
    for part in parts:
        loc = lambda a, nc: part.localize_nodal(a, nc)
        t, pl, n, _ = k1(part.conn, part.coords, (loc(u1, 3), loc(p1, 1), loc(u2, 3), loc(p2, 1)))
        rows.append((t, pl, n, part.n_owned_elems))
    tmax = max(r[0] for r in rows)
    print(f"beta {beta}: max {tmax:.4f} ms -> {n1 / tmax / 1e3:.0f} M QP/s = {t1 / tmax:.2f}x of one part; per part "
          + " ".join(f"[{r[2]} el ({r[3]} owned) pl {r[1]:.2f} {r[0]:.3f} ms]" for r in rows), flush=True)
