"""Partitioned runs on the GPU: forward solve + adjoint gradient over N parts vs one part.

* host-staged transport (c8_set_comm_host + gloo): N processes share cuda:0 -- runs on a 1-GPU box
  and covers the owned/ghost row filtering, halo plan, distributed GMRES and the reductions;
* NCCL transport (c8_nccl_init): one process per GPU, needs >= 2 GPUs (`gpurun --gpus 2`).
"""
import json
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _deck(case):
    golden = json.load(open(os.path.join(ROOT, "tests", "golden", "golden.json")))
    if case == "notch_small_J2":
        d = dict(golden["decks"]["notch_small_J2"])
        d["num_steps"] = 2
        return d, "notch", "avg_disp"
    # the shipped synthetic calibration (examples/synthetic_calibration), 2-D plane-stress Hill
    truth = dict(E=1000., nu=.25, Y=2., S=10., D=50., R00=1., R11=1., R22=1., R01=1.)
    d = dict(global_type="mechanics_plane_stress", local_type="small_hill_plane_stress",
             params=dict(truth, Y=2.2, S=8., D=60.), truth=truth,
             dbcs=[[0, 0, "xmin", "0.0"], [0, 1, "ymin", "0.0"], [0, 1, "ymax", "0.01 * t"]],
             num_steps=3, global_max_iters=30, global_tol=1e-12, local_max_iters=20, local_tol=1e-12)
    return d, "notch2D", "calibration"


def _solve(case, mesh, part, ctx_setup, measured=None, area=None, params=None):
    """build the (local) problem, run primal + adjoint; returns J, grad, per-step fields"""
    import torch
    from calibr8_b200.capi import Context, HostProblem
    d, _, qoi = _deck(case)
    dev = torch.cuda.current_device()
    ctx = Context(dev)
    conn, coords = (mesh.conn, mesh.coords) if part is None else (part.conn, part.coords)
    node_sets = mesh.node_sets if part is None else part.node_sets
    ctx.set_mesh(mesh.dim, conn, coords)
    ctx.set_model(d["global_type"], d["local_type"], params or d["params"],
                  max_iters=max(d["local_max_iters"], 1), abs_tol=d["local_tol"], rel_tol=d["local_tol"])
    if part is not None:
        ctx.set_partition(part)
        ctx_setup(ctx)
    hp = HostProblem(ctx)
    hp.set_time(d["num_steps"], 1.0)
    for r, e, s, v in d["dbcs"]:
        hp.add_dbc(r, e, node_sets[s], v)
    hp.finalize_dbcs()
    hp.set_solver(d["global_max_iters"], d["global_tol"], d["global_tol"], gmres_restart=200,
                  gmres_max_iters=20000, linear_tol=1e-12)
    if qoi == "avg_disp":
        hp.set_qoi_avg_disp()
    else:
        m = measured if part is None else np.stack([part.localize_nodal(s, 2).reshape(-1, 2) for s in measured])
        hp.set_qoi_calibration(balance_factor=1e2, coord_idx=1, coord_value=1.0, reaction_force_comp=1,
                               weights=(1e8, 1e8), measured=m, load_data=np.zeros(d["num_steps"]),
                               area=area)
    J = hp.primal_solve()
    g = hp.adjoint_gradient()
    u_last = hp.get_step(d["num_steps"])[0][0]
    stats = ctx.comm_stats()
    stats["krylov_iterations"] = hp.stats()["linear_iters"]
    stats["amg_levels"] = ctx.preconditioner_info()["levels"]
    stats["p2p"] = ctx.p2p_active()
    hp.close(); ctx.close()
    return J, g, u_last, stats


def _measured(case, mesh):
    """calibration data: displacement history of a one-part run at the true parameters"""
    import torch
    from calibr8_b200.capi import Context, HostProblem
    d, _, qoi = _deck(case)
    if qoi != "calibration":
        return None, None
    ctx = Context(torch.cuda.current_device())
    ctx.set_mesh(mesh.dim, mesh.conn, mesh.coords)
    ctx.set_model(d["global_type"], d["local_type"], d["truth"], max_iters=d["local_max_iters"],
                  abs_tol=d["local_tol"], rel_tol=d["local_tol"])
    hp = HostProblem(ctx)
    hp.set_time(d["num_steps"], 1.0)
    for r, e, s, v in d["dbcs"]:
        hp.add_dbc(r, e, mesh.node_sets[s], v)
    hp.finalize_dbcs()
    hp.set_solver(d["global_max_iters"], d["global_tol"], d["global_tol"], gmres_restart=200,
                  gmres_max_iters=20000, linear_tol=1e-12)
    hp.set_qoi_avg_disp()
    hp.primal_solve()
    meas = np.stack([hp.get_step(s)[0][0].reshape(-1, 2) for s in range(1, d["num_steps"] + 1)])
    X = mesh.coords[mesh.conn]
    area = 0.5 * np.abs((X[:, 1, 0] - X[:, 0, 0]) * (X[:, 2, 1] - X[:, 0, 1]) -
                        (X[:, 1, 1] - X[:, 0, 1]) * (X[:, 2, 0] - X[:, 0, 0])).sum()
    hp.close(); ctx.close()
    return meas, float(area)


def _elem_part(mesh_name, how):
    """None (recursive coordinate bisection) or the reference's own two-part SCOREC split of the mesh"""
    if how != "reference":
        return None
    return np.load(os.path.join(ROOT, "tests", "golden", f"partition_{mesh_name}_2p.npy")).astype(np.int32)


def _worker(rank, world, port, transport, case, q, how="rcb"):
    import sys
    if transport == "nccl_p2p":     # experimental NVLink push halo (csrc/comm.cu), same checks
        os.environ["C8_P2P"] = "1"
        transport = "nccl"
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    from calibr8_b200 import partition
    from conftest import load_mesh
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.cuda.set_device(rank if transport == "nccl" else 0)
        _, mesh_name, _ = _deck(case)
        mesh = load_mesh(mesh_name)
        measured, area = _measured(case, mesh)
        _, part = partition.partition_mesh(mesh, world, rank=rank, elem_part=_elem_part(mesh_name, how))

        def setup(ctx):
            if transport == "host":
                ctx.set_comm_host(partition.HostExchange(part))
            else:
                def bcast(b):
                    t = torch.zeros(128, dtype=torch.uint8) if b is None else torch.tensor(list(b), dtype=torch.uint8)
                    dist.broadcast(t, 0)
                    return bytes(t.tolist())
                ctx.nccl_init(rank, world, bcast)

        J, g, u_last, stats = _solve(case, mesh, part, setup, measured, area)
        u_owned = u_last.reshape(-1, mesh.dim)[: part.n_owned_nodes]
        q.put((rank, J, g, part.node_gid[: part.n_owned_nodes], u_owned, stats))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def _run_parts(world, transport, case, how="rcb"):
    import torch.multiprocessing as mp
    mpc = mp.get_context("spawn")
    q = mpc.Queue()
    port = _free_port()
    procs = [mpc.Process(target=_worker, args=(r, world, port, transport, case, q, how)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    return sorted(res, key=lambda r: r[0])


def _check(world, transport, case, how="rcb"):
    import torch
    from conftest import load_mesh
    torch.cuda.set_device(0)
    d, mesh_name, _ = _deck(case)
    mesh = load_mesh(mesh_name)
    measured, area = _measured(case, mesh)
    J1, g1, u1, st1 = _solve(case, mesh, None, None, measured, area)
    res = _run_parts(world, transport, case, how)
    u = np.zeros((mesh.n_nodes, mesh.dim))
    for rank, J, g, gid, u_owned, stats in res:
        assert abs(J - J1) <= 1e-8 * abs(J1), (rank, J, J1)                       # objective, 1e-8 relative
        assert np.abs(g - g1).max() <= 1e-8 * np.abs(g1).max(), (rank, g, g1)     # adjoint gradient
        assert stats["halo_calls"] > 0 and stats["allreduce_calls"] > 0
        if transport == "nccl_p2p":   # a silent fall-back to ncclSend/Recv must not pass as the push halo
            assert stats["p2p"]["halo"], stats
        # the multigrid hierarchy spans the parts: the Krylov iteration count stays that of one part
        # (a hierarchy on each part's owned block needs 1.4x (2 parts) to 2x (8 parts) as many)
        assert stats["krylov_iterations"] <= 1.25 * st1["krylov_iterations"] + 10, (stats, st1)
        u[gid] = u_owned
    assert np.abs(u - u1.reshape(-1, mesh.dim)).max() <= 1e-8 * np.abs(u1).max()


@pytest.mark.parametrize("case", ["notch_small_J2", "calibration2D"])
def test_two_parts_host_staged_one_gpu(case):
    _check(2, "host", case)


@pytest.mark.parametrize("case", ["notch_small_J2", "calibration2D"])
def test_reference_partition_host_staged_one_gpu(case):
    """element ownership = the reference's own offline two-part split of its regression meshes
    (test/mesh/notch/notch_2p{0,1}.smb, SCOREC split -> ParMETIS), fixtures tests/golden/partition_*_2p.npy"""
    _check(2, "host", case, how="reference")


def test_three_parts_host_staged_one_gpu():
    _check(3, "host", "calibration2D")


@pytest.mark.parametrize("case", ["notch_small_J2", "calibration2D"])
def test_two_parts_nccl(case):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    _check(2, "nccl", case)


@pytest.mark.skipif(os.environ.get("C8_TEST_P2P") != "1",
                    reason="experimental NVLink push halo: set C8_TEST_P2P=1 on a box with >= 2 GPUs")
def test_two_parts_nccl_p2p_push_halo():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    _check(2, "nccl_p2p", "notch_small_J2")
