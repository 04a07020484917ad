"""Partition / halo-plan host logic (CPU only; the N>1 path at world size 2 over gloo).

The design under test: every part evaluates its owned elements plus the halo
elements touching an owned node, and therefore holds the complete rows of its owned nodes -- no
matrix/residual export; only ghost entries of nodal vectors (halo copy) and scalars travel.
"""
import os
import socket

import numpy as np
import pytest

from calibr8_b200 import meshgen, partition
from parity_common import COMBOS, make_oracle, synthetic_fields, xlist


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _mesh_and_partition(kind, n_parts):
    """structured synthetic boxes, or the reference's own UNSTRUCTURED regression meshes (gmsh) with
    RCB or with the reference's own two-part SCOREC split (ParMETIS) as the element ownership"""
    import os
    from conftest import GOLDEN, load_mesh
    if kind == "box3d":
        return meshgen.box_tets(6, notch_radius=0.3), None
    if kind == "box2d":
        return meshgen.square_tris(12, notch_radius=0.3), None
    name, how = kind.split(":")
    mesh = load_mesh(name)
    ep = np.load(os.path.join(GOLDEN, f"partition_{name}_2p.npy")).astype(np.int32) if how == "reference" else None
    return mesh, ep


@pytest.mark.parametrize("kind,n_parts", [("box3d", 2), ("box3d", 3), ("box3d", 8), ("box2d", 2), ("box2d", 4),
                                          ("notch:rcb", 2), ("notch:rcb", 4), ("notch2D:rcb", 4),
                                          ("notch:reference", 2), ("notch2D:reference", 2), ("cube:reference", 2)])
def test_partition_invariants(kind, n_parts):
    mesh, ep = _mesh_and_partition(kind, n_parts)
    elem_part, parts = partition.partition_mesh(mesh, n_parts, elem_part=ep)
    counts = np.bincount(elem_part, minlength=n_parts)
    assert counts.sum() == mesh.n_elems and counts.max() - counts.min() <= n_parts
    if ep is not None:
        assert np.array_equal(elem_part, ep)      # the reference's ownership is used as given
    owned_nodes = np.concatenate([p.node_gid[: p.n_owned_nodes] for p in parts])
    assert np.array_equal(np.sort(owned_nodes), np.arange(mesh.n_nodes))      # each node owned once
    owned_elems = np.concatenate([p.elem_gid[: p.n_owned_elems] for p in parts])
    assert np.array_equal(np.sort(owned_elems), np.arange(mesh.n_elems))      # each element owned once
    for p in parts:
        # local connectivity reproduces the global one
        assert np.array_equal(p.node_gid[p.conn], np.asarray(mesh.conn)[p.elem_gid])
        # every element touching an owned node is local (complete owned rows)
        owner_is_me = np.zeros(mesh.n_nodes, dtype=bool)
        owner_is_me[p.node_gid[: p.n_owned_nodes]] = True
        touching = np.nonzero(owner_is_me[np.asarray(mesh.conn)].any(axis=1))[0]
        assert set(touching) <= set(p.elem_gid)
        assert p.recv_ptr[-1] == p.n_nodes - p.n_owned_nodes
        assert (p.send_nodes < p.n_owned_nodes).all()
        # plan symmetry: what r sends to q is exactly q's ghost range from r, in the same order
        for k, q in enumerate(p.nbr_rank):
            Q = parts[q]
            kq = list(Q.nbr_rank).index(p.rank)
            sent = p.node_gid[p.send_nodes[p.send_ptr[k]: p.send_ptr[k + 1]]]
            got = Q.node_gid[Q.n_owned_nodes + Q.recv_ptr[kq]: Q.n_owned_nodes + Q.recv_ptr[kq + 1]]
            assert np.array_equal(sent, got)
    assert partition.edge_cut(mesh, elem_part) > 0


def test_weighted_rcb_balances_the_cost():
    """rcb(weights=): the cuts balance the summed per-element cost (the role of ParMETIS vertex weights) and
    the partition stays a valid one (owned nodes / elements disjoint and complete)."""
    mesh = meshgen.box_tets(8, notch_radius=0.3)
    cent = mesh.coords[mesh.conn].mean(axis=1)
    w = 1.0 + 3.0 * (cent[:, 1] > 0.6)            # a "plastic zone" four times as expensive
    for n_parts in (2, 3, 8):
        ep, parts = partition.partition_mesh(mesh, n_parts, weights=w)
        cost = np.array([w[ep == k].sum() for k in range(n_parts)])
        assert cost.max() / cost.mean() < 1.03, cost
        ep0, _ = partition.partition_mesh(mesh, n_parts)
        cost0 = np.array([w[ep0 == k].sum() for k in range(n_parts)])
        assert cost.max() < cost0.max()           # better balanced than the element-count cut
        owned = np.concatenate([p.elem_gid[: p.n_owned_elems] for p in parts])
        assert np.array_equal(np.sort(owned), np.arange(mesh.n_elems))


def test_rcb_is_exact_on_structured_boxes():
    mesh = meshgen.box_tets(8)
    elem_part, _ = partition.partition_mesh(mesh, 2, rank=0)
    assert np.bincount(elem_part).tolist() == [mesh.n_elems // 2, mesh.n_elems // 2]
    # a planar cut through a structured n^3 Kuhn mesh crosses 2 n^2 triangular faces
    assert partition.edge_cut(mesh, elem_part) == 2 * 8 * 8


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        name = "3d_small_J2"
        dim, gtype, ltype, params, amp = COMBOS[name]
        mesh = meshgen.box_tets(4, notch_radius=0.3)
        _, part = partition.partition_mesh(mesh, world, rank=rank)
        ex = partition.HostExchange(part)
        # 1. halo copy: owned entries known, ghosts filled by the exchange
        nb = 4
        f_global = np.arange(mesh.n_nodes * nb, dtype=np.float64).reshape(-1, nb) * 0.5 + 1.0
        f_local = np.full((part.n_nodes, nb), np.nan)
        f_local[: part.n_owned_nodes] = f_global[part.node_gid[: part.n_owned_nodes]]
        recv = ex.exchange(f_local[part.send_nodes], nb)
        f_local[part.n_owned_nodes:] = recv
        ok_halo = bool(np.array_equal(f_local, f_global[part.node_gid]))
        # 2. allreduce
        s = ex.allreduce(np.array([float(part.n_owned_elems), float(part.n_owned_nodes)]))
        ok_sum = (s[0] == mesh.n_elems) and (s[1] == mesh.n_nodes)
        # 3. complete owned rows without export: the oracle on the local mesh (owned + halo elements)
        #    reproduces the serial oracle's residual on the owned rows
        (u1, p1), (u2, p2) = synthetic_fields(mesh, amp, True)
        serial = make_oracle(mesh, gtype, ltype, params)
        xi0 = serial.init_xi()
        rS = serial.forward_jacobian(xlist(u2, p2), xlist(u1, p1), xi0, xi0)

        class LM:
            pass
        lm = LM(); lm.dim = 3; lm.conn = part.conn; lm.coords = part.coords
        local = make_oracle(lm, gtype, ltype, params)
        lx = [part.localize_nodal(u2, 3), part.localize_nodal(p2, 1)]
        lxp = [part.localize_nodal(u1, 3), part.localize_nodal(p1, 1)]
        lxi0 = local.init_xi()
        rL = local.forward_jacobian(lx, lxp, lxi0, lxi0)
        no = part.n_owned_nodes
        bu_ref = rS["b"][0].reshape(-1, 3)[part.node_gid[:no]]
        bp_ref = rS["b"][1][part.node_gid[:no]]
        err = max(np.abs(rL["b"][0].reshape(-1, 3)[:no] - bu_ref).max(),
                  np.abs(rL["b"][1][:no] - bp_ref).max()) / np.abs(bu_ref).max()
        # and the owned elements' local state equals the serial state
        err_xi = np.abs(rL["xi"][: part.n_owned_elems] - rS["xi"][part.elem_gid[: part.n_owned_elems]]).max()
        q.put((rank, ok_halo, bool(ok_sum), float(err), float(err_xi)))
    finally:
        dist.destroy_process_group()


def test_halo_exchange_and_owned_rows_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok_halo, ok_sum, err, err_xi in res:
        assert ok_halo, f"rank {rank}: ghost values differ from the owners'"
        assert ok_sum
        assert err < 1e-12, f"rank {rank}: owned rows differ from the serial assembly ({err:.2e})"
        assert err_xi < 1e-14
