#!/usr/bin/env python
"""K1 on the 2-D plane-stress 4M-triangle mesh (BASELINE configs[3] shape) and K7/K8 kernel times."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from calibr8_b200 import meshgen
from calibr8_b200.capi import Context
from calibr8_b200.vfm import vfm_forward, vfm_adjoint
HILL2D = dict(E=1000., nu=.25, Y=2., S=10., D=50., R00=1., R11=1., R22=1., R01=1.)
mesh = meshgen.square_tris(int(os.environ.get("NCELLS", "1414")))
ctx = Context(0); ctx.set_mesh(2, mesh.conn, mesh.coords)
ctx.set_model("mechanics_plane_stress", "small_hill_plane_stress", HILL2D, max_iters=20, abs_tol=1e-12, rel_tol=1e-12)
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
X, Y = mesh.coords[:, 0], mesh.coords[:, 1]
base = np.stack([-0.25 * 0.004 * X + 2e-4 * np.sin(3 * Y) * X, 0.004 * Y + 2e-4 * np.sin(2 * X) * Y], axis=1).reshape(-1)
x, xp, x0 = ctx.alloc("x"), ctx.alloc("x"), ctx.alloc("x")
xi0, xip, xi = ctx.alloc("xi"), ctx.alloc("xi"), ctx.alloc("xi")
A, b, path = ctx.alloc("A"), ctx.alloc("b"), ctx.alloc("path")
ctx.pack_x(base * 1.0, None, x); ctx.pack_x(base * 0.6, None, xp)
ctx.init_xi(xi0); ctx.init_xi(xip)
assert ctx.forward_jacobian(xp, x0, xi0, xip, None, b) == 0
def timeit(fn, reps=8):
    ts = []
    for k in range(reps + 2):
        b.zero_(); xi.copy_(xip)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        if k >= 2: ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))
t = timeit(lambda: ctx.forward_jacobian(x, xp, xip, xi, A, b, path, check=False))
print(f"K1 2-D plane stress small_hill: {ctx.n_elems} tris {t:.3f} ms -> {ctx.n_elems/t/1e3:.0f} M QP/s, plastic {float(path.float().mean()):.2f}")
w = ctx.alloc("x"); w.copy_(torch.randn_like(w))
hist = ctx.alloc("xi"); grad = torch.zeros(64, dtype=torch.float64, device="cuda")
t7 = timeit(lambda: vfm_forward(ctx, x, xp, xip, xi, b))
t8 = timeit(lambda: vfm_adjoint(ctx, x, xp, xi, xip, w, 1.0, hist, grad))
print(f"K7 vfm_forward (no sens): {t7:.3f} ms ({ctx.n_elems/t7/1e3:.0f} M QP/s); K8 vfm_adjoint: {t8:.3f} ms ({ctx.n_elems/t8/1e3:.0f} M QP/s)")
