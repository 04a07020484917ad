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
ctx.close(); dist.destroy_process_group()
