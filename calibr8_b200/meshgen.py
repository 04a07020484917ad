"""Deterministic synthetic meshes of the sizes BASELINE.json names.

* ``box_tets(n)``    unit cube, n cells per side, each cell Kuhn-split into 6 tets
                     (n=55 -> 998 250 tets / 175 616 nodes, n=110 -> 7 986 000 tets), optional
                     quarter-circle notch of radius r at the z-axis edge removed cell-wise
                     (the geometry of the reference's test/mesh/notch/notch.cpp:7-27).
* ``square_tris(n)`` unit square, n x n cells x 2 triangles (n=1414 -> 3 998 792 tris).

Returned as ``calibr8_b200.meshio.Mesh`` with coordinate-plane node sets
(xmin, xmax, ymin, ymax[, zmin, zmax]) like the reference's assoc files.
"""
from __future__ import annotations

import itertools

import numpy as np

from .meshio import Mesh


def _node_sets(coords, dim, tol=1e-12):
    ns = {}
    for k, name in enumerate("xyz"[:dim]):
        lo, hi = coords[:, k].min(), coords[:, k].max()
        ns[name + "min"] = np.nonzero(np.abs(coords[:, k] - lo) < tol)[0].astype(np.int32)
        ns[name + "max"] = np.nonzero(np.abs(coords[:, k] - hi) < tol)[0].astype(np.int32)
    return ns


def _compact(coords, conn):
    used = np.zeros(coords.shape[0], dtype=bool)
    used[conn.ravel()] = True
    new_id = np.cumsum(used) - 1
    return np.ascontiguousarray(coords[used]), new_id[conn].astype(np.int32)


def box_tets(n, notch_radius=0.0, lengths=(1.0, 1.0, 1.0)) -> Mesh:
    n1 = n + 1
    g = np.arange(n1, dtype=np.float64) / n
    # node id = i + n1*(j + n1*k): x fastest
    Z, Y, X = np.meshgrid(g * lengths[2], g * lengths[1], g * lengths[0], indexing="ij")
    coords = np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1)
    k, j, i = np.meshgrid(np.arange(n), np.arange(n), np.arange(n), indexing="ij")
    base = (i + n1 * (j + n1 * k)).ravel().astype(np.int64)
    stride = np.array([1, n1, n1 * n1], dtype=np.int64)
    tets = []
    for perm in itertools.permutations(range(3)):
        v0 = base
        v1 = v0 + stride[perm[0]]
        v2 = v1 + stride[perm[1]]
        v3 = v2 + stride[perm[2]]
        # orientation = sign of the permutation
        sign = np.linalg.det(np.eye(3)[list(perm)])
        t = np.stack([v0, v1, v2, v3], axis=1) if sign > 0 else np.stack([v0, v2, v1, v3], axis=1)
        tets.append(t)
    # cell-major ordering: the 6 tets of a cell are consecutive (locality for the scatter)
    conn = np.stack(tets, axis=1).reshape(-1, 4)
    if notch_radius > 0.0:
        cx = (i.ravel() + 0.5) / n * lengths[0]
        cy = (j.ravel() + 0.5) / n * lengths[1]
        keep = (cx * cx + cy * cy) >= notch_radius * notch_radius
        conn = conn.reshape(-1, 6, 4)[keep].reshape(-1, 4)
    coords, conn = _compact(coords, conn)
    return Mesh(dim=3, coords=coords, conn=conn, elem_set=np.zeros(conn.shape[0], dtype=np.int32),
                elem_set_names=["body"], node_sets=_node_sets(coords, 3), side_sets={})


def square_tris(n, notch_radius=0.0) -> Mesh:
    n1 = n + 1
    g = np.arange(n1, dtype=np.float64) / n
    Y, X = np.meshgrid(g, g, indexing="ij")
    coords = np.stack([X.ravel(), Y.ravel(), np.zeros(n1 * n1)], axis=1)
    j, i = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
    v00 = (i + n1 * j).ravel().astype(np.int64)
    v10, v01, v11 = v00 + 1, v00 + n1, v00 + n1 + 1
    conn = np.stack([np.stack([v00, v10, v11], axis=1), np.stack([v00, v11, v01], axis=1)],
                    axis=1).reshape(-1, 3)
    if notch_radius > 0.0:
        cx = (i.ravel() + 0.5) / n
        cy = (j.ravel() + 0.5) / n
        keep = (cx * cx + cy * cy) >= notch_radius * notch_radius
        conn = conn.reshape(-1, 2, 3)[keep].reshape(-1, 3)
    coords, conn = _compact(coords, conn)
    return Mesh(dim=2, coords=coords, conn=conn, elem_set=np.zeros(conn.shape[0], dtype=np.int32),
                elem_set_names=["body"], node_sets=_node_sets(coords, 2), side_sets={})


def smooth_field(mesh: Mesh, amplitude, seed=0):
    """A deterministic smooth displacement field (a few low Fourier modes + a uniaxial stretch)
    used as synthetic nodal input for kernel benchmarks and parity tests."""
    rng = np.random.RandomState(seed)
    x = mesh.coords
    u = np.zeros((mesh.n_nodes, mesh.dim))
    u[:, 1] += amplitude * x[:, 1]
    for _ in range(4):
        kvec = rng.randint(1, 4, size=3) * np.pi
        ph = rng.uniform(0, 2 * np.pi)
        a = rng.uniform(-1, 1, size=mesh.dim) * amplitude * 0.35
        s = np.sin(x @ kvec + ph)
        u += s[:, None] * a[None, :]
    return u
