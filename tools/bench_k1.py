#!/usr/bin/env python
"""Kernel-tuning helper: K1 (eval_forward_jacobian) time on the bench workload for the library
selected by C8B200_LIB.  Prints one line per model.  Not a bench value (see bench.py)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from calibr8_b200.capi import Context

def run(ltype, params, amp_scale, mesh, reps=10):
    (u1, p1), (u2, p2) = bench.workload_fields(mesh)
    u1, u2 = u1 * amp_scale, u2 * amp_scale
    ctx = Context(0)
    ctx.set_mesh(mesh.dim, mesh.conn, mesh.coords)
    ctx.set_model("mechanics", ltype, params, **bench.LOCAL)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    x, xp, x0 = ctx.alloc("x"), ctx.alloc("x"), ctx.alloc("x")
    xi0, xip, xi = ctx.alloc("xi"), ctx.alloc("xi"), ctx.alloc("xi")
    A, b, path = ctx.alloc("A"), ctx.alloc("b"), ctx.alloc("path")
    ctx.pack_x(u2, p2, x); ctx.pack_x(u1, p1, xp)
    ctx.init_xi(xi0); ctx.init_xi(xip)
    assert ctx.forward_jacobian(xp, x0, xi0, xip, None, b) == 0
    ts = []
    for k in range(reps + 3):
        b.zero_(); xi.copy_(xip)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ctx.forward_jacobian(x, xp, xip, xi, A, b, path, check=False)
        e1.record(); torch.cuda.synchronize()
        if k >= 3: ts.append(e0.elapsed_time(e1))
    nf = ctx.forward_jacobian(x, xp, xip, xi, None, None, path)
    pl = float(path.float().mean())
    chk = float(A.abs().sum()), float(b.abs().sum())
    print(f"{os.environ.get('C8B200_LIB','default').split('/')[-1]:28s} {ltype:10s} n={ctx.n_elems} "
          f"K1 {np.median(ts):7.3f} ms (min {min(ts):.3f}) {ctx.n_elems/np.median(ts)/1e3:7.1f} MQP/s plastic {pl:.2f} "
          f"nf {nf} sumA {chk[0]:.10e} sumb {chk[1]:.10e}", flush=True)
    ctx.close()

mesh = bench.workload_mesh(int(os.environ.get("NCELLS", "56")))
which = os.environ.get("MODELS", "hyper_J2,small_J2").split(",")
if "hyper_J2" in which:
    run("hyper_J2", bench.PARAMS, 1.0, mesh)
if "small_J2" in which:
    run("small_J2", dict(E=1000., nu=.25, K=100., Y=2., cte=0., delta_T=0.), 0.2, mesh)
if "small_hill" in which:
    run("small_hill", dict(E=1000., nu=.25, Y=2., R00=1., R11=.9, R22=1.1, R01=1., R02=.95, R12=1.05, S=10., D=2.), 0.2, mesh)
if "hypo_hill" in which:
    run("hypo_hill", dict(E=1000., nu=.25, Y=2., R00=1., R11=.9, R22=1.1, R01=1., R02=.95, R12=1.05, S=10., D=2.), 0.2, mesh)
if "elastic" in which:
    run("elastic", dict(E=1000., nu=.25, cte=0., delta_T=0.), 1.0, mesh)
