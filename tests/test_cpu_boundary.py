"""CPU-only checks of the drop-in boundary and host logic (no compute calls without a GPU):
the C-ABI library loads, exports every symbol include/c8b200.h declares, and refuses to run
without a CUDA device; mesh readers / generators behave."""
import ctypes
import os
import re

import numpy as np

from conftest import ROOT, load_mesh


def _declared_symbols():
    txt = open(os.path.join(ROOT, "include", "c8b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(c8_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    from calibr8_b200 import capi, vfm  # noqa: F401  (vfm extends the symbol lists)
    lib = capi.load_library()
    declared = _declared_symbols()
    assert len(declared) >= 40
    missing = [s for s in declared if not hasattr(lib, s)]
    assert not missing, missing
    # the Python binding's own list is a subset of the header
    assert set(capi.SYMBOLS) <= set(declared), set(capi.SYMBOLS) - set(declared)
    assert not [s for s in capi.HOST_SYMBOLS if not hasattr(lib, s)]


def test_no_cpu_fallback():
    """Without a CUDA device the product path must fail loudly (c8_create returns NULL)."""
    import torch
    from calibr8_b200 import capi
    if torch.cuda.is_available():
        return
    lib = capi.load_library()
    assert lib.c8_create(0) is None
    try:
        capi.Context(0)
    except capi.C8Error as ex:
        assert "no CUDA device" in str(ex)
    else:
        raise AssertionError("Context() must raise without a GPU")


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under calibr8_b200/ may reference it."""
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "calibr8_b200")):
        if "lib" in dirpath.split(os.sep):
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp")):
                txt = open(os.path.join(dirpath, f)).read()
                if re.search(r"^\s*(from|import)\s+oracle|liboracle|#include\s+\"[^\"]*oracle", txt, re.M):
                    bad.append(f)
    assert not bad, bad


def test_reference_meshes_round_trip():
    for name, (dim, nv, ne) in {"cube": (3, 14, 24), "notch2D": (2, 252, 447),
                                "notch": (3, 546, 1550)}.items():
        m = load_mesh(name)  # counts of SURVEY.md section 4 / test/unit/disc.cpp.in:15-29
        assert (m.dim, m.n_nodes, m.n_elems) == (dim, nv, ne)
        assert m.conn.min() == 0 and m.conn.max() == nv - 1
        for k, v in m.node_sets.items():
            axis = "xyz".index(k[0])
            assert np.ptp(m.coords[v, axis]) < 1e-12, k


def test_mesh_generators():
    from calibr8_b200 import meshgen
    m = meshgen.box_tets(5, notch_radius=0.25)
    X = m.coords[m.conn]
    vol = np.linalg.det(np.stack([X[:, 1] - X[:, 0], X[:, 2] - X[:, 0], X[:, 3] - X[:, 0]], 1)) / 6
    assert vol.min() > 0
    n_cells = m.n_elems // 6
    assert abs(vol.sum() - n_cells / 125.0) < 1e-12
    assert set(m.node_sets) == {"xmin", "xmax", "ymin", "ymax", "zmin", "zmax"}
    m2 = meshgen.square_tris(7)
    assert m2.n_elems == 98 and m2.n_nodes == 64
    # the 1M-tet configuration of BASELINE.json has exactly this many tets before the notch
    assert 55 ** 3 * 6 == 998250


def test_residual_factories_match_the_reference_metadata():
    """create_local_residual / create_global_residual (host C++): names, variable types, equation counts
    and parameter order as the reference's constructors set them (e.g. src/small_J2.cpp:36-60,
    src/hyper_J2_plane_stress.cpp:43-63), and the Python tables agree with them."""
    import pytest
    from calibr8_b200 import capi
    cases = {("small_J2", "mechanics", 3): (["pstrain", "alpha"], [6, 1]),
             ("small_hill", "mechanics", 3): (["pstrain", "alpha"], [6, 1]),
             ("hyper_J2", "mechanics", 3): (["zeta", "Ie", "alpha"], [6, 1, 1]),
             ("elastic", "mechanics", 2): (["dummy"], [1]),
             ("small_hill_plane_stress", "mechanics_plane_stress", 2): (["pstrain", "alpha"], [3, 1]),
             ("hyper_J2_plane_stress", "mechanics_plane_stress", 2): (["zeta", "Ie", "lambda_z", "alpha"], [3, 1, 1, 1]),
             ("hyper_J2_plane_strain", "mechanics", 2): (["zeta", "Ie", "alpha"], [3, 1, 1]),
             ("hypo_hill", "mechanics", 3): (["TC", "alpha"], [6, 1]),
             ("hypo_hill_plane_strain", "mechanics", 2): (["TC", "alpha", "TC_zz"], [3, 1, 1]),
             ("hypo_hill_plane_stress", "mechanics_plane_stress", 2): (["TC", "alpha", "lambda_z"], [3, 1, 1])}
    for (lt, gt, nd), (names, neq) in cases.items():
        d = capi.describe_residuals(lt, gt, nd)
        assert d["local"]["resid_names"] == names and d["local"]["num_eqs"] == neq
        assert d["local"]["param_names"] == capi.PARAM_NAMES[lt]
        assert d["local"]["c8_type"] == capi.LOCAL_TYPES[lt] and d["global"]["c8_type"] == capi.GLOBAL_TYPES[gt]
        assert d["global"]["resid_names"] == (["u", "p"] if gt == "mechanics" else ["u"])
        assert d["global"]["num_ip_sets"] == (2 if gt == "mechanics" else 1)
    assert capi.describe_residuals("hyper_J2_plane_stress", "mechanics_plane_stress", 2)["local"]["z_stretch_idx"] == 2
    assert capi.describe_residuals("hypo_hill_plane_stress", "mechanics_plane_stress", 2)["local"]["z_stretch_idx"] == 2
    with pytest.raises(capi.C8Error):
        capi.describe_residuals("hypo_barlat", "mechanics", 3)        # out of the hot-path scope
    with pytest.raises(capi.C8Error):
        capi.describe_residuals("small_hill_plane_stress", "mechanics_plane_stress", 3)
