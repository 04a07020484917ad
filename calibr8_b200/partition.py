"""Element partition of a simplex mesh over the GPUs of one box, and the owned/ghost halo plan.

Replaces, for this path, what the reference gets from an offline SCOREC ``split`` (Zoltan ->
ParMETIS, test/mesh/notch/Makefile:24-25) plus the OWNED/GHOST maps of ``Disc::compute_*_maps``
(src/disc.cpp:271-314).  ParMETIS is not in this image and the reference pins no partition, so the
partitioner is a recursive coordinate bisection of the element centroids (exact slabs/bricks on the
structured synthetic boxes); everything else is deterministic bookkeeping:

* node owner      = lowest part id among the elements touching the node
* local nodes     = owned nodes (ascending global id), then ghosts sorted by (owner, global id)
* local elements  = owned elements (ascending global id), then the halo elements: elements of other
                    parts that touch an owned node.  A part evaluates its halo elements redundantly so
                    that the rows of its owned nodes are complete without any matrix/residual export.
* halo plan       = per neighbour: owned local nodes to send (in the receiver's ghost order) and the
                    contiguous ghost range to receive into (c8_set_halo_plan, include/c8b200.h)

Pure numpy: runs without a GPU (the CPU ``gloo`` tests exercise it at world size 2).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np


def rcb(centroids: np.ndarray, n_parts: int, weights: np.ndarray | None = None) -> np.ndarray:
    """Recursive coordinate bisection.  Returns part id per element; part sizes differ by <= 1
    element per cut level (non-power-of-two counts are split proportionally).  weights: per-element
    cost estimate (the vertex weights a ParMETIS run would be given); the cuts then balance the summed
    weight instead of the element count."""
    part = np.zeros(centroids.shape[0], dtype=np.int32)
    w = None if weights is None else np.asarray(weights, dtype=np.float64)

    def split(idx, p0, np_):
        if np_ == 1:
            part[idx] = p0
            return
        c = centroids[idx]
        axis = int(np.argmax(c.max(axis=0) - c.min(axis=0)))
        left_parts = np_ // 2
        # stable order on (coordinate, element id) makes the cut deterministic on structured grids
        order = np.lexsort((idx, c[:, axis]))
        if w is None:
            n_left = int(round(len(idx) * left_parts / np_))
        else:
            cw = np.cumsum(w[idx[order]])
            n_left = int(np.searchsorted(cw, cw[-1] * left_parts / np_, side="left")) + 1
            n_left = min(max(n_left, left_parts), len(idx) - (np_ - left_parts))
        split(idx[order[:n_left]], p0, left_parts)
        split(idx[order[n_left:]], p0 + left_parts, np_ - left_parts)

    split(np.arange(centroids.shape[0]), 0, n_parts)
    return part


@dataclass
class Part:
    rank: int
    n_parts: int
    dim: int
    conn: np.ndarray            # [n_local_elems][nn] local node ids
    coords: np.ndarray          # [n_local_nodes][3]
    elem_set: np.ndarray
    n_owned_nodes: int
    n_owned_elems: int
    node_gid: np.ndarray        # local node -> global node
    elem_gid: np.ndarray        # local elem -> global elem
    nbr_rank: np.ndarray
    send_ptr: np.ndarray
    send_nodes: np.ndarray      # owned local ids, concatenated per neighbour
    recv_ptr: np.ndarray        # ghost offsets (relative to n_owned_nodes), per neighbour
    node_sets: dict = field(default_factory=dict)   # name -> local node ids (owned AND ghost)

    @property
    def n_nodes(self):
        return self.coords.shape[0]

    @property
    def n_elems(self):
        return self.conn.shape[0]

    def localize_nodal(self, a: np.ndarray, ncomp: int) -> np.ndarray:
        """global nodal array [n_global_nodes*ncomp] -> this part's [n_nodes*ncomp]"""
        return np.ascontiguousarray(np.asarray(a).reshape(-1, ncomp)[self.node_gid]).reshape(-1)

    def localize_elem(self, a: np.ndarray) -> np.ndarray:
        return np.ascontiguousarray(np.asarray(a)[self.elem_gid])


def node_owners(conn: np.ndarray, elem_part: np.ndarray, n_nodes: int) -> np.ndarray:
    owner = np.full(n_nodes, np.iinfo(np.int32).max, dtype=np.int32)
    np.minimum.at(owner, conn.ravel(), np.repeat(elem_part, conn.shape[1]))
    return owner


def build_part(mesh, elem_part: np.ndarray, rank: int, n_parts: int, owner=None) -> Part:
    conn, coords = np.asarray(mesh.conn), np.asarray(mesh.coords)
    n_nodes = coords.shape[0]
    if owner is None:
        owner = node_owners(conn, elem_part, n_nodes)
    owned_nodes = np.nonzero(owner == rank)[0]
    is_owned_node = owner == rank
    owned_elems = np.nonzero(elem_part == rank)[0]
    touches = is_owned_node[conn].any(axis=1)
    halo_elems = np.nonzero(touches & (elem_part != rank))[0]
    elem_gid = np.concatenate([owned_elems, halo_elems])
    used = np.unique(conn[elem_gid].ravel())
    ghosts = used[owner[used] != rank]
    ghosts = ghosts[np.lexsort((ghosts, owner[ghosts]))]
    node_gid = np.concatenate([owned_nodes, ghosts])
    g2l = np.full(n_nodes, -1, dtype=np.int64)
    g2l[node_gid] = np.arange(node_gid.size)
    lconn = g2l[conn[elem_gid]].astype(np.int32)
    assert (lconn >= 0).all()
    # receive plan: ghosts grouped by owner
    g_owner = owner[ghosts]
    recv_ranks, recv_counts = np.unique(g_owner, return_counts=True)
    # send plan: nodes I own that part q holds as ghosts = my owned nodes used by q's local elements
    # (q's local elements = its owned elements + elements touching a node q owns)
    send_lists = {}
    for q in range(n_parts):
        if q == rank:
            continue
        q_local = (elem_part == q) | (owner[conn] == q).any(axis=1)
        q_used = np.unique(conn[q_local].ravel())
        mine = q_used[owner[q_used] == rank]            # ascending global id == q's ghost order
        if mine.size:
            send_lists[q] = g2l[mine].astype(np.int32)
    nbrs = sorted(set(send_lists) | set(int(r) for r in recv_ranks))
    send_ptr, recv_ptr, send_nodes = [0], [0], []
    rc = dict(zip((int(r) for r in recv_ranks), (int(c) for c in recv_counts)))
    for q in nbrs:
        s = send_lists.get(q, np.zeros(0, dtype=np.int32))
        send_nodes.append(s)
        send_ptr.append(send_ptr[-1] + s.size)
        recv_ptr.append(recv_ptr[-1] + rc.get(q, 0))
    node_sets = {}
    for name, ids in getattr(mesh, "node_sets", {}).items():
        l = g2l[np.asarray(ids)]
        node_sets[name] = l[l >= 0].astype(np.int32)
    es = np.asarray(mesh.elem_set) if getattr(mesh, "elem_set", None) is not None else np.zeros(conn.shape[0], np.int32)
    return Part(rank=rank, n_parts=n_parts, dim=mesh.dim, conn=lconn,
                coords=np.ascontiguousarray(coords[node_gid]), elem_set=es[elem_gid].astype(np.int32),
                n_owned_nodes=int(owned_nodes.size), n_owned_elems=int(owned_elems.size),
                node_gid=node_gid, elem_gid=elem_gid,
                nbr_rank=np.asarray(nbrs, dtype=np.int32), send_ptr=np.asarray(send_ptr, dtype=np.int32),
                send_nodes=(np.concatenate(send_nodes) if send_nodes else np.zeros(0, np.int32)).astype(np.int32),
                recv_ptr=np.asarray(recv_ptr, dtype=np.int32), node_sets=node_sets)


def partition_mesh(mesh, n_parts: int, rank: int | None = None, elem_part=None, weights=None):
    """Element partition; returns (elem_part, [Part...]) or (elem_part, Part) for one rank.
    elem_part: the caller's own element -> part map (the hook for METIS / ParMETIS output or for the
    reference's offline SCOREC `split`, see meshio.reference_partition); None: recursive coordinate
    bisection of the element centroids (works on unstructured meshes too, exact on structured boxes),
    balancing the element count or, with `weights`, a per-element cost estimate."""
    conn, coords = np.asarray(mesh.conn), np.asarray(mesh.coords)
    if elem_part is None:
        cent = coords[conn].mean(axis=1)[:, : mesh.dim]
        elem_part = rcb(cent, n_parts, weights)
    else:
        elem_part = np.ascontiguousarray(elem_part, dtype=np.int32)
        if elem_part.shape != (conn.shape[0],) or elem_part.min() < 0 or elem_part.max() >= n_parts:
            raise ValueError("elem_part must map every element to a part in [0, n_parts)")
    owner = node_owners(conn, elem_part, coords.shape[0])
    if rank is not None:
        return elem_part, build_part(mesh, elem_part, rank, n_parts, owner)
    return elem_part, [build_part(mesh, elem_part, r, n_parts, owner) for r in range(n_parts)]


def edge_cut(mesh, elem_part: np.ndarray) -> int:
    """number of element faces shared between two parts (so that scaling numbers are interpretable)"""
    conn = np.asarray(mesh.conn)
    nn = conn.shape[1]
    faces, owner = [], []
    for k in range(nn):
        f = np.sort(np.delete(conn, k, axis=1), axis=1)
        faces.append(f)
        owner.append(elem_part)
    faces = np.concatenate(faces)
    owner = np.concatenate(owner)
    order = np.lexsort(faces.T[::-1])
    faces, owner = faces[order], owner[order]
    same = (faces[1:] == faces[:-1]).all(axis=1)
    return int((same & (owner[1:] != owner[:-1])).sum())


# ---- host exchange over torch.distributed (gloo on CPU, or any backend with host tensors) ----
class HostExchange:
    """The two host callbacks of c8_set_comm_host, implemented with torch.distributed point-to-point
    and all_reduce on host tensors -- what an MPI rank of the reference would do with PCU/MPI."""

    def __init__(self, part: Part, group=None):
        import torch.distributed as dist
        self.part, self.dist, self.group = part, dist, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)

    def exchange(self, send: np.ndarray, nb: int) -> np.ndarray:
        """send: [n_send][nb] packed in send_ptr order; returns [n_ghost][nb] in recv_ptr order"""
        import torch
        p, dist = self.part, self.dist
        send = np.ascontiguousarray(send, dtype=np.float64).reshape(-1, nb)
        recv = np.zeros((int(p.recv_ptr[-1]), nb))
        ops, keep = [], []
        for k, q in enumerate(p.nbr_rank):
            s0, s1, r0, r1 = p.send_ptr[k], p.send_ptr[k + 1], p.recv_ptr[k], p.recv_ptr[k + 1]
            if s1 > s0:
                t = torch.from_numpy(np.ascontiguousarray(send[s0:s1]))
                keep.append(t)
                ops.append(dist.P2POp(dist.isend, t, int(q), group=self.group))
            if r1 > r0:
                t = torch.from_numpy(recv[r0:r1])
                keep.append(t)
                ops.append(dist.P2POp(dist.irecv, t, int(q), group=self.group))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        return recv

    def allreduce(self, buf: np.ndarray) -> np.ndarray:
        import torch
        t = torch.from_numpy(np.ascontiguousarray(buf, dtype=np.float64))
        self.dist.all_reduce(t, group=self.group)
        return t.numpy()
