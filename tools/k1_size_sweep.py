#!/usr/bin/env python
"""K1 time against the mesh size, timed (a) right after an idle gap (sync + 0.3 s sleep, what a bench
that starts its clock sampler before the timed loop does) and (b) after a 50 ms untimed spin-up of the
same step.  Shows how much of the small-mesh 'fixed cost' is the GPU leaving its idle clocks.
Not a bench value (see bench.py)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from calibr8_b200.capi import Context


def run(ncells, steps=20):
    mesh = bench.workload_mesh(ncells)
    (u1, p1), (u2, p2) = bench.workload_fields(mesh)
    ctx = Context(0)
    ctx.set_mesh(mesh.dim, mesh.conn, mesh.coords)
    ctx.set_model("mechanics", "hyper_J2", bench.PARAMS, **bench.LOCAL)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    x, xp, x0 = ctx.alloc("x"), ctx.alloc("x"), ctx.alloc("x")
    xi0, xip, xi = ctx.alloc("xi"), ctx.alloc("xi"), ctx.alloc("xi")
    A, b, path = ctx.alloc("A"), ctx.alloc("b"), ctx.alloc("path")
    ctx.pack_x(u2, p2, x); ctx.pack_x(u1, p1, xp)
    ctx.init_xi(xi0); ctx.init_xi(xip)
    assert ctx.forward_jacobian(xp, x0, xi0, xip, None, b) == 0

    def step():
        b.zero_(); xi.copy_(xip)
        ctx.forward_jacobian(x, xp, xip, xi, A, b, path, check=False)

    def timed():
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps

    out = {}
    for mode in ("idle", "spun"):
        ts = []
        for rep in range(5):
            for _ in range(3):
                step()
            torch.cuda.synchronize()
            time.sleep(0.3)
            if mode == "spun":
                t0 = time.perf_counter()
                while time.perf_counter() - t0 < 0.05:
                    step()
            ts.append(timed())
        out[mode] = float(np.median(ts))
    print(f"ncells {ncells:3d} n_elems {ctx.n_elems:8d}  after idle {out['idle']:.4f} ms/step   after spin-up {out['spun']:.4f} ms/step   "
          f"({ctx.n_elems / out['spun'] / 1e3:.1f} M QP/s)", flush=True)
    ctx.close()


for nc in [int(v) for v in os.environ.get("NCELLS", "14,20,28,40,56").split(",")]:
    run(nc)
