"""On-disk mesh formats at the boundary of the hot path.

* ``read_smb``   SCOREC MDS ``.smb`` (big-endian, version 5), the mesh format
                 calibr8 loads at ``src/disc.cpp:31-39``.
* ``read_dmg``   ASCII ``.dmg`` geometric model (model topology only).
* ``read_assoc`` set-association text file parsed at ``src/disc.cpp:56-100``.
* ``load_calibr8_mesh``  the three together -> flat arrays + named sets, i.e.
                 what ``Disc`` derives at ``src/disc.cpp:486-538``.

Only single-part meshes (``<name>0.smb``) are read.  Element vertex order is
rebuilt from the downward chain and tets are re-oriented to positive volume;
results of the finite-element path do not depend on the local vertex order.
"""
from __future__ import annotations

import dataclasses
import struct

import numpy as np

# MDS type order in the .smb header
_TYPES = ("vtx", "edge", "tri", "quad", "hex", "prism", "pyramid", "tet")
_DOWN = {"edge": 2, "tri": 3, "quad": 4, "hex": 6, "prism": 5, "pyramid": 5, "tet": 4}


@dataclasses.dataclass
class Mesh:
    dim: int
    coords: np.ndarray          # [n_nodes, 3] float64
    conn: np.ndarray            # [n_elems, dim+1] int32
    elem_set: np.ndarray        # [n_elems] int32 (index into elem_set_names)
    elem_set_names: list
    node_sets: dict             # name -> int32 node ids (sorted)
    side_sets: dict             # name -> [n_sides, dim] int32 vertex ids of each side

    @property
    def n_nodes(self):
        return self.coords.shape[0]

    @property
    def n_elems(self):
        return self.conn.shape[0]


def read_smb(path):
    """Return dict(dim, coords, edges, tris, tets (vertex lists), class_{vtx,edge,tri,tet})."""
    buf = open(path, "rb").read()
    magic, version, dim, nparts = struct.unpack(">4I", buf[:16])
    if magic != 0:
        raise ValueError(f"{path}: unsupported smb header magic={magic}")
    counts = dict(zip(_TYPES, struct.unpack(">8I", buf[16:48])))
    off = 48
    down = {}
    for t in _TYPES[1:]:
        n = counts[t] * _DOWN[t]
        down[t] = np.frombuffer(buf, dtype=">u4", count=n, offset=off).astype(np.int64)
        down[t] = down[t].reshape(counts[t], _DOWN[t])
        off += 4 * n
    nv = counts["vtx"]
    xyz = np.frombuffer(buf, dtype=">f8", count=nv * 3, offset=off).astype(np.float64).reshape(nv, 3)
    off += 8 * nv * 3
    off += 8 * nv * 2  # parametric coordinates
    # remote-copy links of a part of a partitioned mesh (<name>_<n>p<part>.smb): number of peer parts,
    # then per peer (peer id, count, `count` local vertex ids in the order both sides share)
    (n_peers,) = struct.unpack(">I", buf[off:off + 4])
    off += 4
    remotes = {}
    for _ in range(n_peers):
        peer, cnt = struct.unpack(">2I", buf[off:off + 8])
        off += 8
        remotes[int(peer)] = np.frombuffer(buf, dtype=">u4", count=cnt, offset=off).astype(np.int64)
        off += 4 * cnt
    if n_peers and nparts == 1:
        raise ValueError("remote copies present in a serial mesh")
    cls = {}
    for t in _TYPES:
        n = counts[t]
        c = np.frombuffer(buf, dtype=">u4", count=2 * n, offset=off).astype(np.int64).reshape(n, 2)
        off += 8 * n
        cls[t] = c  # (model_tag, model_dim)
    edges = down["edge"]
    # tri vertices: v0 = e0 ∩ e2, v1 = e0 ∩ e1, v2 = e1 ∩ e2
    tris_e = down["tri"]
    tris = np.zeros((counts["tri"], 3), dtype=np.int64)
    for k in range(counts["tri"]):
        e0, e1, e2 = (set(edges[e]) for e in tris_e[k])
        tris[k] = [(e0 & e2).pop(), (e0 & e1).pop(), (e1 & e2).pop()]
    tets_f = down["tet"]
    tets = np.zeros((counts["tet"], 4), dtype=np.int64)
    for k in range(counts["tet"]):
        f0 = list(tris[tets_f[k, 0]])
        rest = set(tris[tets_f[k, 1]]) - set(f0)
        tets[k] = f0 + [rest.pop()]
    if counts["tet"]:
        a = xyz[tets[:, 1]] - xyz[tets[:, 0]]
        b = xyz[tets[:, 2]] - xyz[tets[:, 0]]
        c = xyz[tets[:, 3]] - xyz[tets[:, 0]]
        vol = np.einsum("ij,ij->i", a, np.cross(b, c))
        neg = vol < 0
        tets[neg] = tets[neg][:, [0, 2, 1, 3]]
    return dict(dim=int(dim), nparts=int(nparts), remotes=remotes, coords=xyz, edges=edges, tris=tris, tets=tets,
                tris_e=tris_e, tets_f=tets_f,
                cls_vtx=cls["vtx"], cls_edge=cls["edge"], cls_tri=cls["tri"], cls_tet=cls["tet"])


def read_dmg(path):
    """Model topology: closure[(dim, tag)] = set of (dim', tag') of all bounding entities."""
    tok = open(path).read().split()
    it = iter(tok)
    nr, nf, ne, nv = (int(next(it)) for _ in range(4))
    for _ in range(6):
        next(it)
    verts = []
    for _ in range(nv):
        verts.append(int(next(it)))
        next(it); next(it); next(it)
    edge_v = {}
    for _ in range(ne):
        tag = int(next(it))
        edge_v[tag] = (int(next(it)), int(next(it)))
    face_e = {}
    for _ in range(nf):
        tag = int(next(it)); nloops = int(next(it))
        es = []
        for _ in range(nloops):
            n = int(next(it))
            for _ in range(n):
                es.append(int(next(it))); next(it)
        face_e[tag] = es
    region_f = {}
    for _ in range(nr):
        tag = int(next(it)); nshells = int(next(it))
        fs = []
        for _ in range(nshells):
            n = int(next(it))
            for _ in range(n):
                fs.append(int(next(it))); next(it)
        region_f[tag] = fs
    closure = {}
    for v in verts:
        closure[(0, v)] = {(0, v)}
    for e, vs in edge_v.items():
        closure[(1, e)] = {(1, e)} | {(0, v) for v in vs if v >= 0}
    for f, es in face_e.items():
        c = {(2, f)}
        for e in es:
            c |= closure[(1, e)]
        closure[(2, f)] = c
    for r, fs in region_f.items():
        c = {(3, r)}
        for f in fs:
            c |= closure[(2, f)]
        closure[(3, r)] = c
    return closure


def read_assoc(path):
    """-> list of (kind, name, [(model_dim, model_tag), ...])  (src/disc.cpp:56-100)."""
    lines = [ln.split() for ln in open(path).read().splitlines() if ln.strip()]
    out = []
    i = 0
    while i < len(lines):
        kind = lines[i][0] + " " + lines[i][1]
        name = lines[i][2]
        n = int(lines[i][3])
        ents = [(int(lines[i + 1 + k][0]), int(lines[i + 1 + k][1])) for k in range(n)]
        out.append((kind, name, ents))
        i += 1 + n
    return out


def reference_partition(mesh: Mesh, part_paths) -> np.ndarray:
    """Element ownership of a mesh the reference partitioned offline (SCOREC `split`: Zoltan ->
    ParMETIS, test/mesh/*/Makefile): part_paths = the `<name>_<n>p<k>.smb` files in part order.
    Elements are matched to the serial mesh through their vertex coordinates.  -> elem_part [n_elems]"""
    dim = mesh.dim
    key = lambda X: tuple(sorted(tuple(np.round(x, 12)) for x in X))
    where = {key(mesh.coords[mesh.conn[e]]): e for e in range(mesh.n_elems)}
    elem_part = np.full(mesh.n_elems, -1, dtype=np.int32)
    for k, path in enumerate(part_paths):
        smb = read_smb(path)
        conn = smb["tets"] if dim == 3 else smb["tris"]
        for c in conn:
            e = where[key(smb["coords"][c])]
            assert elem_part[e] < 0, "an element belongs to exactly one part"
            elem_part[e] = k
    assert (elem_part >= 0).all(), "every element of the serial mesh is in a part"
    return elem_part


def load_calibr8_mesh(smb_path, dmg_path, assoc_path) -> Mesh:
    smb = read_smb(smb_path)
    closure = read_dmg(dmg_path)
    assoc = read_assoc(assoc_path)
    dim = smb["dim"]
    conn = smb["tets"] if dim == 3 else smb["tris"]
    elem_cls = smb["cls_tet"] if dim == 3 else smb["cls_tri"]
    side_vtx = smb["tris"] if dim == 3 else smb["edges"]
    side_cls = smb["cls_tri"] if dim == 3 else smb["cls_edge"]
    elem_set_names, node_sets, side_sets = [], {}, {}
    elem_set = np.full(conn.shape[0], -1, dtype=np.int32)
    for kind, name, ents in assoc:
        if kind == "elem set":
            idx = len(elem_set_names)
            elem_set_names.append(name)
            for (mdim, mtag) in ents:
                elem_set[(elem_cls[:, 1] == mdim) & (elem_cls[:, 0] == mtag)] = idx
        elif kind == "node set":
            # all nodes classified on the closure of the listed entities (src/disc.cpp:519-538)
            members = set()
            for ent in ents:
                members |= closure[ent]
            cv = smb["cls_vtx"]
            mask = np.array([(int(d), int(t)) in members for t, d in cv], dtype=bool)
            node_sets[name] = np.nonzero(mask)[0].astype(np.int32)
        elif kind == "side set":
            mask = np.zeros(side_vtx.shape[0], dtype=bool)
            for (mdim, mtag) in ents:
                mask |= (side_cls[:, 1] == mdim) & (side_cls[:, 0] == mtag)
            side_sets[name] = side_vtx[mask].astype(np.int32)
    if (elem_set < 0).any():
        raise ValueError("elements without an element set")
    return Mesh(dim=dim, coords=np.ascontiguousarray(smb["coords"]),
                conn=np.ascontiguousarray(conn.astype(np.int32)), elem_set=elem_set,
                elem_set_names=elem_set_names, node_sets=node_sets, side_sets=side_sets)


def side_set_facets(mesh: Mesh, name: str) -> np.ndarray:
    """[n_elems][3] int8: local vertex ids of the element facet that lies on side set `name`, -1 where
    none (2-D: two ids).  The mapping QoI::setup_side_set_mapping builds (src/qoi.cpp, load_mismatch.cpp:48-76)."""
    dim = mesh.dim
    sides = {tuple(sorted(int(v) for v in s)) for s in np.asarray(mesh.side_sets[name])}
    fac = np.full((mesh.n_elems, 3), -1, dtype=np.int8)
    import itertools
    combos = list(itertools.combinations(range(dim + 1), dim))
    conn = np.asarray(mesh.conn)
    node_on = np.zeros(mesh.n_nodes, dtype=bool)
    for s in sides:
        node_on[list(s)] = True
    cand = np.nonzero(node_on[conn].sum(axis=1) >= dim)[0]
    for e in cand:
        for c in combos:
            if tuple(sorted(int(conn[e, k]) for k in c)) in sides:
                fac[e, :dim] = c
    return fac


def save_npz(mesh: Mesh, path):
    d = dict(dim=mesh.dim, coords=mesh.coords, conn=mesh.conn, elem_set=mesh.elem_set,
             elem_set_names=np.array(mesh.elem_set_names))
    for k, v in mesh.node_sets.items():
        d["ns_" + k] = v
    for k, v in mesh.side_sets.items():
        d["ss_" + k] = v
    np.savez_compressed(path, **d)


def load_npz(path) -> Mesh:
    z = np.load(path)
    ns = {k[3:]: z[k] for k in z.files if k.startswith("ns_")}
    ss = {k[3:]: z[k] for k in z.files if k.startswith("ss_")}
    return Mesh(dim=int(z["dim"]), coords=z["coords"], conn=z["conn"], elem_set=z["elem_set"],
                elem_set_names=[str(s) for s in z["elem_set_names"]], node_sets=ns, side_sets=ss)
