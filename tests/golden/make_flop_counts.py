"""Reference-algorithm flop counts per quadrature point (op-counting build of the oracle, 16-wide AD on
every operation like Sacado SLFad<double,16>) for the kernels of the path on the bench state
(BASELINE configs[1]: hyper-J2, 54 % plastic points).  Written to tests/golden/flop_counts.json, which
bench.py and tools/bench_kernels.py read -- the product and its benchmark never call the oracle for this.

    python tests/golden/make_flop_counts.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.join(HERE, "..", "..")
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from oracle.pyoracle import Oracle  # noqa: E402


def main(n_cells=6):
    mesh = bench.workload_mesh(n_cells)
    (u1, p1), (u2, p2) = bench.workload_fields(mesh)
    o = Oracle(3, mesh.conn, mesh.coords, global_type="mechanics", local_type="hyper_J2", params=[bench.PARAMS],
               count_flops=True, active=[[0, 1, 2, 7]], **bench.LOCAL)
    o.set_qoi_avg_disp()
    xi0 = o.init_xi()
    rA = o.forward_jacobian([u1, p1], o.zeros_x(), xi0, xi0, assemble=False)
    n = mesh.n_elems
    out = {"state": "bench.workload_fields on bench.workload_mesh(%d): %d tets" % (n_cells, n),
           "counting": "+ - * / and sqrt/exp/pow/cbrt as 1 each, per AD lane and value, 16 lanes (SURVEY.md 8(d))"}
    o.flops_reset(); rB = o.forward_jacobian([u2, p2], [u1, p1], rA["xi"], rA["xi"]); out["K1"] = o.flops_reset() / n
    out["plastic_fraction"] = float(rB["path"].mean())
    o.flops_reset(); o.global_residual([u2, p2], [u1, p1], rB["xi"], rA["xi"]); out["K2"] = o.flops_reset() / n
    g = np.zeros((n, o.n_xi)); f = np.zeros((n, o.n_x))
    o.flops_reset(); o.adjoint_jacobian([u2, p2], [u1, p1], rB["xi"], rA["xi"], g, f, 1); out["K3"] = o.flops_reset() / n
    z = [np.random.RandomState(1).randn(mesh.n_nodes * 3), np.random.RandomState(2).randn(mesh.n_nodes)]
    o.flops_reset(); phi = o.adjoint_local([u2, p2], [u1, p1], rB["xi"], rA["xi"], z, g, f); out["K4"] = o.flops_reset() / n
    o.flops_reset(); o.qoi([u2, p2], [u1, p1], rB["xi"], rA["xi"], 1); out["K5"] = o.flops_reset() / n
    o.flops_reset(); o.qoi_gradient([u2, p2], [u1, p1], rB["xi"], rA["xi"], z, phi, [[0, 1, 2, 3]], 4, 1)
    out["K6"] = o.flops_reset() / n
    # K1 on the general-path states of bench.py (iterated local Newton, exp / pow in the yield law)
    for key, (ltype, params, amp_scale) in bench.GENERAL_PATHS.items():
        og = Oracle(3, mesh.conn, mesh.coords, global_type="mechanics", local_type=ltype, params=[params],
                    count_flops=True, **bench.LOCAL)
        x0g = og.init_xi()
        a1, a2 = [u1 * amp_scale, p1], [u2 * amp_scale, p2]
        rAg = og.forward_jacobian(a1, og.zeros_x(), x0g, x0g, assemble=False)
        og.flops_reset(); rBg = og.forward_jacobian(a2, a1, rAg["xi"], rAg["xi"])
        out["K1_" + key] = og.flops_reset() / n
        out["plastic_fraction_" + key] = float(rBg["path"].mean())
        out["newton_iters_plastic_" + key] = float(rBg["iters"][rBg["path"] == 1].mean())
    json.dump(out, open(os.path.join(HERE, "flop_counts.json"), "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
