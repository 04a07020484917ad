#!/usr/bin/env python
"""Krylov iterations and time of the partitioned forward + adjoint solve for several part counts on
ONE GPU (host-staged transport, the parts are processes sharing cuda:0): shows what the multigrid
hierarchy across the parts buys over a hierarchy on each part's owned block.

    python tools/partitioned_iterations.py --cells 24 --parts 1 2 4 8
"""
import argparse, json, os, socket, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PAR = {"small_J2": dict(E=1000., nu=.25, K=100., Y=2., cte=0., delta_T=0.),
       "hyper_J2": dict(E=1000., nu=.25, Y=10., S=0., D=0., A=0., n=0., K=100.)}


def worker(rank, world, port, a, distributed, q):
    import torch
    import torch.distributed as dist
    from calibr8_b200 import meshgen, partition
    from calibr8_b200.capi import Context, HostProblem
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    if world > 1:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.cuda.set_device(0)
    mesh = meshgen.box_tets(a.cells, notch_radius=0.2)
    ctx = Context(0)
    if world > 1:
        _, part = partition.partition_mesh(mesh, world, rank=rank)
        ctx.set_mesh(3, part.conn, part.coords)
        node_sets = part.node_sets
    else:
        ctx.set_mesh(3, mesh.conn, mesh.coords)
        node_sets = mesh.node_sets
    ctx.set_model("mechanics", a.model, PAR[a.model], max_iters=500, abs_tol=1e-12, rel_tol=1e-12)
    if world > 1:
        ctx.set_partition(part)
        ctx.set_comm_host(partition.HostExchange(part))
    ctx.set_preconditioner("amg", distributed=distributed, replicate_max_nodes=a.replicate)
    hp = HostProblem(ctx)
    hp.set_time(a.steps, 1.0)
    hp.add_dbc(0, 0, node_sets["xmin"], "0.0")
    hp.add_dbc(0, 1, node_sets["ymin"], "0.0")
    hp.add_dbc(0, 2, node_sets["zmin"], "0.0")
    hp.add_dbc(0, 1, node_sets["ymax"], "0.001 * t")
    hp.finalize_dbcs()
    hp.set_solver(15, 1e-8, 1e-8, gmres_restart=100, gmres_max_iters=20000, linear_tol=1e-8)
    hp.set_qoi_avg_disp()
    torch.cuda.synchronize(); t0 = time.time()
    J = hp.primal_solve()
    g = hp.adjoint_gradient()
    torch.cuda.synchronize(); t1 = time.time()
    s = hp.stats()
    if rank == 0:
        q.put(dict(parts=world, distributed=bool(distributed), J=J, grad=[float(v) for v in g], seconds=t1 - t0,
                   krylov_iterations=s["linear_iters"], assemblies=s["assemblies"],
                   preconditioner=ctx.preconditioner_info(), comm=ctx.comm_stats(), n_elems=mesh.n_elems))
    hp.close(); ctx.close()
    if world > 1:
        dist.barrier(); dist.destroy_process_group()


def main():
    import torch.multiprocessing as mp
    ap = argparse.ArgumentParser()
    ap.add_argument("--cells", type=int, default=24)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--model", default="small_J2")
    ap.add_argument("--parts", type=int, nargs="+", default=[1, 2, 4, 8])
    ap.add_argument("--replicate", type=int, default=30000)
    ap.add_argument("--modes", nargs="+", default=["dist", "local"])
    a = ap.parse_args()
    mpc = mp.get_context("spawn")
    for world in a.parts:
        for mode in (a.modes if world > 1 else ["dist"]):
            q = mpc.Queue()
            s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
            procs = [mpc.Process(target=worker, args=(r, world, port, a, mode == "dist", q)) for r in range(world)]
            for p in procs: p.start()
            res = q.get(timeout=600)
            for p in procs: p.join(timeout=120)
            print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
