#!/usr/bin/env python
"""torchrun helper: latency of the library's NCCL halo copy / allreduce on a partitioned box."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import bench
from calibr8_b200 import partition
from calibr8_b200.capi import Context

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
torch.cuda.set_device(lr)
mesh = bench.workload_mesh(int(os.environ.get("NCELLS", "56")))
_, part = partition.partition_mesh(mesh, world, rank=rank)
ctx = Context(lr)
ctx.set_mesh(3, part.conn, part.coords)
ctx.set_model("mechanics", "hyper_J2", bench.PARAMS, **bench.LOCAL)
ctx.set_partition(part)
def bcast(raw):
    t = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if raw is not None: t.copy_(torch.tensor(list(raw), dtype=torch.uint8))
    dist.broadcast(t, 0); return bytes(t.cpu().tolist())
ctx.nccl_init(rank, world, bcast)
x = torch.ones(ctx.n_dofs, dtype=torch.float64, device="cuda")
buf = torch.ones(64, dtype=torch.float64, device="cuda")
for name, fn in [("allreduce(1)", lambda: ctx.allreduce(buf[:1])), ("allreduce(64)", lambda: ctx.allreduce(buf)),
                 ("halo", lambda: ctx.halo(x))]:
    for _ in range(20): fn()
    ctx.synchronize(); dist.barrier(); t0 = time.perf_counter()
    for _ in range(500): fn()
    ctx.synchronize(); t1 = time.perf_counter()
    if rank == 0: print(f"{name:14s} {1e6*(t1-t0)/500:8.1f} us/call  (ghost nodes {part.n_nodes-part.n_owned_nodes})", flush=True)
# K1 + gather and a GMRES solve in the partitioned context
import numpy as np
(u1, p1), (u2, p2) = bench.workload_fields(mesh)
x, xp, x0 = ctx.alloc("x"), ctx.alloc("x"), ctx.alloc("x")
xi0, xip, xi = ctx.alloc("xi"), ctx.alloc("xi"), ctx.alloc("xi")
A, b = ctx.alloc("A"), ctx.alloc("b")
ctx.pack_x(part.localize_nodal(u2, 3), part.localize_nodal(p2, 1), x)
ctx.pack_x(part.localize_nodal(u1, 3), part.localize_nodal(p1, 1), xp)
ctx.init_xi(xi0); ctx.init_xi(xip)
ctx.forward_jacobian(xp, x0, xi0, xip, None, b)
for k in range(6):
    b.zero_(); xi.copy_(xip); ctx.synchronize(); torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
    nf = ctx.forward_jacobian(x, xp, xip, xi, A, b)
    ctx.synchronize(); t1 = time.perf_counter()
    if rank == 0: print(f"K1 partitioned (n_elems local {ctx.n_elems}): {1e3*(t1-t0):.2f} ms nf={nf}", flush=True)
rhs = torch.randn(ctx.n_dofs, dtype=torch.float64, device="cuda"); sol = torch.zeros_like(rhs)
for k in range(3):
    sol.zero_(); ctx.synchronize(); torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
    info = ctx.gmres(A, rhs, sol, restart=100, max_iters=300, rel_tol=1e-8)
    ctx.synchronize(); t1 = time.perf_counter()
    if rank == 0: print(f"gmres: {info} {1e3*(t1-t0):.1f} ms -> {1e3*(t1-t0)/max(info['iters'],1):.3f} ms/it", flush=True)
ctx.close(); dist.destroy_process_group()
