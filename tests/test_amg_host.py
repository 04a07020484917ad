"""CPU check of the partitioned multigrid hierarchy (calibr8_b200/csrc/amg_host.hpp): simulated
parts (threads) build the hierarchy of a structured node graph; tests/amg_host_check.cpp verifies
the halo plans, ghost aggregates and Galerkin lists of every level against the global product."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def checker(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("amg") / "amg_host_check")
    subprocess.check_call(["g++", "-std=c++20", "-O2", "-pthread", "-I", os.path.join(ROOT, "calibr8_b200", "csrc"),
                           os.path.join(ROOT, "tests", "amg_host_check.cpp"), "-o", exe])
    return exe


# nx ny nz  px py pz  replicate_max_nodes
CASES = [
    ("serial", "8 8 8 1 1 1 1000"),
    ("two parts, coarse level replicated at once", "12 10 9 2 1 1 1000"),
    ("eight parts, replicated at once", "12 10 9 2 2 2 1000"),
    ("eight parts, one distributed coarse level", "16 16 16 2 2 2 100"),
    ("six ragged parts, two distributed coarse levels", "20 18 16 3 2 1 50"),
    ("eight parts, never replicated", "24 24 24 2 2 2 20"),
    ("twelve tiny parts", "6 5 4 3 2 2 10"),
]


@pytest.mark.parametrize("name,args", CASES, ids=[c[0] for c in CASES])
def test_partitioned_hierarchy_matches_global_galerkin(checker, name, args):
    out = subprocess.run([checker] + args.split(), capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    assert out.stdout.startswith("ok "), out.stdout
