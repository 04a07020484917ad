#!/usr/bin/env python
"""BASELINE.json configs[4]: 3-D small-strain J2 plasticity on a partitioned notched box, forward
load steps + adjoint gradient (average-displacement objective), strong scaling over the GPUs of one box.

    python tools/run_cfg5.py --cells 110 --load-steps 2                      (one GPU)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29512 tools/run_cfg5.py --cells 110 --load-steps 2      (N parts, NCCL)

cells = 110 is the 8M-tet mesh of the config (SURVEY 8(d) cfg 5); every rank generates the mesh and
keeps its own part (RCB).  Prints one JSON line on rank 0: ms per load step (max over ranks), Krylov
iterations, assemblies, hierarchy, communication counts.
"""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
from calibr8_b200 import meshgen, partition
from calibr8_b200.capi import Context, HostProblem

ap = argparse.ArgumentParser()
ap.add_argument("--cells", type=int, default=110)
ap.add_argument("--load-steps", type=int, default=2)
ap.add_argument("--model", default="small_J2")
ap.add_argument("--local-amg", action="store_true", help="hierarchy on each part's owned block (for comparison)")
ap.add_argument("--replicate", type=int, default=30000)
ap.add_argument("--lin-tol", type=float, default=1e-8, help="GMRES relative tolerance (the reference decks use 1e-6)")
a = ap.parse_args()
PAR = {"small_J2": dict(E=1000., nu=.25, K=100., Y=2., cte=0., delta_T=0.),   # test/adjoint/notch2D_small_J2_adjoint_check.yaml.in:26-33
       "hyper_J2": dict(E=1000., nu=.25, Y=10., S=0., D=0., A=0., n=0., K=100.)}
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
local_rank = int(os.environ.get("LOCAL_RANK", "0"))
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
torch.cuda.set_device(local_rank)
dev = torch.device("cuda", local_rank)
t0 = time.time()
mesh = meshgen.box_tets(a.cells, notch_radius=0.2)
ctx = Context(local_rank)
if world > 1:
    _, part = partition.partition_mesh(mesh, world, rank=rank)
    ctx.set_mesh(3, part.conn, part.coords); node_sets = part.node_sets
else:
    part = None
    ctx.set_mesh(3, mesh.conn, mesh.coords); node_sets = mesh.node_sets
ctx.set_model("mechanics", a.model, PAR[a.model], max_iters=500, abs_tol=1e-12, rel_tol=1e-12)
stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream); ctx.set_stream(stream.cuda_stream)
if world > 1:
    ctx.set_partition(part)

    def bcast(raw):
        t = torch.zeros(128, dtype=torch.uint8, device=dev)
        if raw is not None:
            t.copy_(torch.tensor(list(raw), dtype=torch.uint8))
        dist.broadcast(t, 0)
        return bytes(t.cpu().tolist())
    ctx.nccl_init(rank, world, bcast)
ctx.set_preconditioner("amg", distributed=not a.local_amg, replicate_max_nodes=a.replicate)
hp = HostProblem(ctx)
hp.set_time(a.load_steps, 1.0)
hp.add_dbc(0, 0, node_sets["xmin"], "0.0")
hp.add_dbc(0, 1, node_sets["ymin"], "0.0")
hp.add_dbc(0, 2, node_sets["zmin"], "0.0")
hp.add_dbc(0, 1, node_sets["ymax"], "0.001 * t")
hp.finalize_dbcs()
hp.set_solver(15, 1e-8, 1e-8, gmres_restart=100, gmres_max_iters=20000, linear_tol=a.lin_tol)
hp.set_qoi_avg_disp()
setup_s = time.time() - t0
out = {}
for rep in range(2):     # the first pass builds the hierarchy and the iteration graphs
    s0 = hp.stats()
    if world > 1: dist.barrier()
    torch.cuda.synchronize(); t1 = time.perf_counter()
    J = hp.primal_solve()
    torch.cuda.synchronize(); t2 = time.perf_counter()
    g = hp.adjoint_gradient()
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    t3 = time.perf_counter()
    s1 = hp.stats()
    tt = torch.tensor([t3 - t1, t2 - t1, t3 - t2], dtype=torch.float64, device=dev)
    if world > 1: dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    tot, fwd, adj = [float(v) for v in tt.tolist()]
    out = dict(config="3D small_J2 notched box, forward + adjoint gradient", n_gpus=world, cells=a.cells,
               n_elems=mesh.n_elems, n_nodes=mesh.n_nodes, load_steps=a.load_steps,
               ms_per_load_step=tot / a.load_steps * 1e3, forward_ms=fwd / a.load_steps * 1e3,
               adjoint_ms=adj / a.load_steps * 1e3, krylov_iterations=s1["linear_iters"] - s0["linear_iters"],
               assemblies=s1["assemblies"] - s0["assemblies"], objective=J, gradient=[float(v) for v in g],
               preconditioner=ctx.preconditioner_info(), amg="owned block" if a.local_amg else "across the parts", linear_tol=a.lin_tol,
               comm_rank0=ctx.comm_stats(), p2p_rank0=ctx.p2p_active(), setup_s=setup_s, pass_index=rep)
    if part is not None:
        out["partition_rank0"] = dict(owned_elems=part.n_owned_elems, halo_elems=part.n_elems - part.n_owned_elems,
                                      ghost_nodes=part.n_nodes - part.n_owned_nodes, neighbours=int(part.nbr_rank.size))
if rank == 0:
    print(json.dumps(out), flush=True)
hp.close(); ctx.close()
if world > 1:
    dist.destroy_process_group()
