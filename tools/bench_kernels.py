#!/usr/bin/env python
"""Per-kernel times of the whole hot path (K1..K8) on the bench workload (1M hyper-J2 tets) and on
the 4M-triangle plane-stress mesh, with the reference algorithm's flop count per quadrature point from
the op-counting oracle (CPU, small sample of the same state) -> fraction of the measured DFMA peak.
Output: one JSON object (also written to gpurun_out/kernels.json when that directory exists)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from calibr8_b200 import meshgen
from calibr8_b200.capi import Context, make_qoi
from calibr8_b200.vfm import vfm_forward, vfm_adjoint


def timeit(fn, pre=lambda: None, reps=6):
    ts = []
    for k in range(reps + 2):
        pre(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        if k >= 2: ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


def ref_flops_3d():
    """flops per QP of the reference algorithm for K1..K6 on the bench state (counted by
    tests/golden/make_flop_counts.py with the op-counting oracle build)"""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    return json.load(open(os.path.join(root, "tests", "golden", "flop_counts.json")))


res = {}
# ---------------------------------------------------------------- 3-D, 1M hyper-J2 tets
mesh = bench.workload_mesh(int(os.environ.get("NCELLS", "56")))
(u1, p1), (u2, p2) = bench.workload_fields(mesh)
ctx = Context(0); ctx.set_mesh(3, mesh.conn, mesh.coords)
ctx.set_model("mechanics", "hyper_J2", bench.PARAMS, **bench.LOCAL)
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
peak = ctx.bench_dfma(4096)
x, xp, x0 = ctx.alloc("x"), ctx.alloc("x"), ctx.alloc("x")
xi0, xip, xi = ctx.alloc("xi"), ctx.alloc("xi"), ctx.alloc("xi")
A, b, path = ctx.alloc("A"), ctx.alloc("b"), ctx.alloc("path")
ctx.pack_x(u2, p2, x); ctx.pack_x(u1, p1, xp); ctx.init_xi(xi0); ctx.init_xi(xip)
assert ctx.forward_jacobian(xp, x0, xi0, xip, None, b) == 0
n = ctx.n_elems
t = {}
t["K1"] = timeit(lambda: ctx.forward_jacobian(x, xp, xip, xi, A, b, path, check=False), lambda: (b.zero_(), xi.copy_(xip)))
t["K2"] = timeit(lambda: ctx.global_residual(x, xp, xi, xip, b), lambda: b.zero_())
g = ctx.alloc("xi"); f = torch.zeros(ctx.xi_ld * ctx.nx, dtype=torch.float64, device="cuda")
rhs = ctx.alloc("b"); z = ctx.alloc("x"); z.copy_(torch.randn_like(z)); phi = ctx.alloc("xi")
q = make_qoi("avg_disp")
t["K3"] = timeit(lambda: ctx.adjoint_jacobian(q, x, xp, xi, xip, g, f, A, rhs), lambda: (rhs.zero_(), g.zero_()))
t["K4"] = timeit(lambda: ctx.adjoint_local(x, xp, xi, xip, z, phi, g, f))
sc = torch.zeros(2, dtype=torch.float64, device="cuda")
t["K5"] = timeit(lambda: ctx.qoi_value(q, x, xp, xi, xip, 0, sc), lambda: sc.zero_())
grad = torch.zeros(64, dtype=torch.float64, device="cuda")
t["K6"] = timeit(lambda: ctx.qoi_gradient(q, x, xp, xi, xip, z, phi, grad), lambda: grad.zero_())
fl = ref_flops_3d()
res["mesh_3d"] = {"n_elems": n, "model": "hyper_J2 mixed u-p", "dfma_peak_tflops": peak}
names = {"K1": "eval_forward_jacobian (+gather)", "K2": "eval_global_residual", "K3": "eval_adjoint_jacobian (+transposed gather)",
         "K4": "solve_adjoint_local", "K5": "eval_qoi", "K6": "eval_qoi_gradient"}
for k in ["K1", "K2", "K3", "K4", "K5", "K6"]:
    tf = fl[k] * n / (t[k] * 1e-3) * 1e-12
    res[k] = {"reference_fn": names[k], "ms": t[k], "M_qp_per_s": n / t[k] / 1e3, "ref_flops_per_qp": fl[k],
              "tflops_ref_count": tf, "frac_of_dfma_peak": tf / peak}
ctx.close()
# ---------------------------------------------------------------- 2-D, 4M plane-stress triangles
HILL2D = dict(E=1000., nu=.25, Y=2., S=10., D=50., R00=1., R11=1., R22=1., R01=1.)
m2 = meshgen.square_tris(int(os.environ.get("NCELLS2D", "1414")))
c2 = Context(0); c2.set_mesh(2, m2.conn, m2.coords)
c2.set_model("mechanics_plane_stress", "small_hill_plane_stress", HILL2D, max_iters=20, abs_tol=1e-12, rel_tol=1e-12)
c2.set_stream(torch.cuda.current_stream().cuda_stream)
X, Y = m2.coords[:, 0], m2.coords[:, 1]
base = np.stack([-0.25 * 0.004 * X + 2e-4 * np.sin(3 * Y) * X, 0.004 * Y + 2e-4 * np.sin(2 * X) * Y], axis=1).reshape(-1)
x, xp, x0 = c2.alloc("x"), c2.alloc("x"), c2.alloc("x")
xi0, xip, xi = c2.alloc("xi"), c2.alloc("xi"), c2.alloc("xi")
b = c2.alloc("b"); A = c2.alloc("A")
c2.pack_x(base, None, x); c2.pack_x(base * 0.6, None, xp); c2.init_xi(xi0); c2.init_xi(xip)
assert c2.forward_jacobian(xp, x0, xi0, xip, None, b) == 0
w = c2.alloc("x"); w.copy_(torch.randn_like(w)); hist = c2.alloc("xi"); grad = torch.zeros(64, dtype=torch.float64, device="cuda")
n2 = c2.n_elems
t2 = {"K1_2d": timeit(lambda: c2.forward_jacobian(x, xp, xip, xi, A, b, None, check=False), lambda: (b.zero_(), xi.copy_(xip))),
      "K7": timeit(lambda: vfm_forward(c2, x, xp, xip, xi, b), lambda: (b.zero_(), xi.copy_(xip))),
      "K8": timeit(lambda: vfm_adjoint(c2, x, xp, xi, xip, w, 1.0, hist, grad), lambda: grad.zero_())}
res["mesh_2d"] = {"n_elems": n2, "model": "small_hill_plane_stress"}
for k, v in t2.items():
    res[k] = {"ms": v, "M_qp_per_s": n2 / v / 1e3}
c2.close()
print(json.dumps(res, indent=1))
if os.path.isdir("gpurun_out"):
    json.dump(res, open("gpurun_out/kernels.json", "w"), indent=1)
