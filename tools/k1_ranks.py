#!/usr/bin/env python
"""Diagnostic for `bench.py --gpus N`: every rank times K1 on its part of the bench mesh (a) with direct
launches and nothing else running, (b) with one nvidia-smi poller per rank (what the bench's clock sampler
did), (c) with one poller on rank 0 only, (d) replaying the step from a CUDA graph.  Launch with torchrun.
Prints one line per mode with the per-rank medians.  Not a bench value."""
import os, subprocess, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import bench
from calibr8_b200 import partition
from calibr8_b200.capi import Context

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
lr = int(os.environ.get("LOCAL_RANK", "0"))
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
mesh = bench.workload_mesh()
(u1, p1), (u2, p2) = bench.workload_fields(mesh)
ctx = Context(lr)
if world > 1:
    _, part = partition.partition_mesh(mesh, world, rank=rank)
    ctx.set_mesh(3, part.conn, part.coords)
    loc = lambda a, nc: part.localize_nodal(a, nc)
    u1, p1, u2, p2 = loc(u1, 3), loc(p1, 1), loc(u2, 3), loc(p2, 1)
else:
    ctx.set_mesh(3, mesh.conn, mesh.coords)
ctx.set_model("mechanics", "hyper_J2", bench.PARAMS, **bench.LOCAL)
stream = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(stream)
ctx.set_stream(stream.cuda_stream)
x, xp, x0 = ctx.alloc("x"), ctx.alloc("x"), ctx.alloc("x")
xi0, xip, xi = ctx.alloc("xi"), ctx.alloc("xi"), ctx.alloc("xi")
A, b, path = ctx.alloc("A"), ctx.alloc("b"), ctx.alloc("path")
ctx.pack_x(u2, p2, x); ctx.pack_x(u1, p1, xp); ctx.init_xi(xi0); ctx.init_xi(xip)
assert ctx.forward_jacobian(xp, x0, xi0, xip, None, b) == 0


def step():
    b.zero_(); xi.copy_(xip)
    ctx.forward_jacobian(x, xp, xip, xi, A, b, path, check=False)


def measure(fn, steps=20):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    t0 = time.perf_counter()
    ev[0].record()
    for k in range(steps):
        fn(); ev[k + 1].record()
    t_cpu = (time.perf_counter() - t0) / steps * 1e3
    torch.cuda.synchronize()
    per = [ev[k].elapsed_time(ev[k + 1]) for k in range(steps)]
    return float(np.median(per)), ev[0].elapsed_time(ev[-1]) / steps, t_cpu


def report(tag, r):
    t = torch.tensor(list(r), dtype=torch.float64, device=dev)
    if world > 1:
        out = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(out, t)
    else:
        out = [t]
    if rank == 0:
        m = torch.stack(out).cpu().numpy()
        print(f"{tag:32s} median step ms per rank {np.round(m[:, 0], 3).tolist()}  mean(max) {m[:, 1].max():.3f}  cpu enqueue ms/step (max) {m[:, 2].max():.3f}", flush=True)


def poller(on):
    if not on:
        return None
    return subprocess.Popen(["nvidia-smi", "--query-gpu=index,clocks.sm,clocks.max.sm,power.draw", "--format=csv,noheader,nounits",
                             "-lms", "100", "-i", str(lr)], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)


if rank == 0:
    print(f"world {world}, host cores {os.cpu_count()}, local elems rank0 {ctx.n_elems}", flush=True)
report("direct, no poller", measure(step))
p = poller(True); time.sleep(0.5)
report("direct, poller on every rank", measure(step))
p.terminate(); p.wait()
if world > 1:
    dist.barrier()
p = poller(rank == 0); time.sleep(0.5)
report("direct, poller on rank 0", measure(step))
if p: p.terminate(); p.wait()
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g, stream=stream):
    step()
torch.cuda.synchronize()
report("graph replay, no poller", measure(g.replay))
p = poller(True); time.sleep(0.5)
report("graph replay, poller every rank", measure(g.replay))
p.terminate(); p.wait()
report("direct, no poller (again)", measure(step, steps=200))
ctx.close()
if world > 1:
    dist.destroy_process_group()
