#!/usr/bin/env python
"""Bisection of the two reference goldens the oracle reproduces only loosely:
notch2D_small_J2 (+1.157e-5) and cube_hyper_J2 (-9.02e-8).  CPU only; test infrastructure.

    python tools/golden_bisect.py tolerance   # global Newton tolerance 1e-8 ... 1e-3
    python tools/golden_bisect.py inexact     # linear solves stopped at the deck's Belos tolerance 1e-6
    python tools/golden_bisect.py variants    # other readings of the 2-D quirks of src/small_J2.cpp:205,275
                                              # (patched COPY of oracle/ under scratch/, never the repo's)
Results of the round-2 run are quoted in tests/test_oracle_golden.py and DESIGN.md section 2.
"""
import json
import os
import shutil
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DECKS = ("notch2D_small_J2", "cube_hyper_J2")


def run(name, oracle_root=ROOT, global_tol=None, residual_tol=None, seed=0):
    code = f"""
import sys, json, numpy as np
sys.path.insert(0, {os.path.join(ROOT, 'tests')!r}); sys.path.insert(0, {ROOT!r}); sys.path.insert(0, {oracle_root!r})
from conftest import load_mesh
import oracle.driver as drv
from oracle.driver import Dbc, Primal
from oracle.pyoracle import Oracle
import scipy.sparse.linalg as spla
d = json.load(open({os.path.join(ROOT, 'tests', 'golden', 'golden.json')!r}))['decks'][{name!r}]
m = load_mesh(d['mesh'])
o = Oracle(m.dim, m.conn, m.coords, global_type=d['global_type'], local_type=d['local_type'], params=[d['params']],
           max_iters=d['local_max_iters'], abs_tol=d['local_tol'], rel_tol=d['local_tol'])
o.set_qoi_avg_disp()
rt = {residual_tol!r}
if rt:
    rng = np.random.RandomState({seed})
    class Inexact:
        splu = spla.splu
        @staticmethod
        def spsolve(K, b):   # stops at |b - K x| = rt |b|, random residual direction
            r = rng.standard_normal(b.shape); r *= rt * np.linalg.norm(b) / np.linalg.norm(r)
            return spla.spsolve(K, b - r)
    drv.spla = Inexact
gt = {global_tol!r} or d['global_tol']
p = Primal(o, [Dbc(r, e, m.node_sets[s], v) for r, e, s, v in d['dbcs']], d['num_steps'], 1.0,
           max_iters=d['global_max_iters'], abs_tol=gt, rel_tol=gt)
J = p.solve()
print(json.dumps(dict(J=J, rel=(J - d['J']) / d['J'])))
"""
    out = subprocess.check_output([sys.executable, "-c", code], env=dict(os.environ))
    return json.loads(out.decode().strip().splitlines()[-1])


def variant_copy():
    dst = os.path.join(ROOT, "scratch", "oracle_variants")
    shutil.rmtree(dst, ignore_errors=True)
    shutil.copytree(os.path.join(ROOT, "oracle"), os.path.join(dst, "oracle"),
                    ignore=shutil.ignore_patterns("*.so", "__pycache__"))
    p = os.path.join(dst, "oracle", "models.hpp")
    s = open(p).read()
    i, j = s.index("class SmallJ2 : public LocalResidual<T>"), s.index("class SmallHill : public LocalResidual<T>")
    blk = s[i:j]
    blk = blk.replace("    T const s_mag = norm(s);\n    Tensor<T> const n = s / s_mag;", """    static int const V = getenv("ORC_V") ? atoi(getenv("ORC_V")) : 0;
    T s_mag = norm(s);
    if (V == 1 && this->m_num_dims == 2) {  // Frobenius norm of the 3-D deviator (s_zz = -(s_xx + s_yy))
      T const szz = -(s(0,0) + s(1,1));
      s_mag = sqrt(s(0,0)*s(0,0) + 2.*s(0,1)*s(0,1) + s(1,1)*s(1,1) + szz*szz);
    }
    Tensor<T> const n = s / s_mag;""")
    blk = blk.replace("    Tensor<T> const dev_eps = eps - (trace(eps) / 3.) * I;", """    static int const V = getenv("ORC_V") ? atoi(getenv("ORC_V")) : 0;
    Tensor<T> const dev_eps = eps - (trace(eps) / ((V == 2) ? double(g.num_dims()) : 3.)) * I;""")
    s = (s[:i] + blk + s[j:]).replace("#pragma once", "#pragma once\n#include <cstdlib>", 1)
    open(p, "w").write(s)
    subprocess.check_call(["make", "-C", os.path.join(dst, "oracle"), "-j4"], stdout=subprocess.DEVNULL)
    return dst


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 else "tolerance"
    if mode == "tolerance":
        for name in DECKS:
            for gt in (1e-8, 1e-6, 1e-5, 1e-4, 1e-3):
                print(name, "global Newton tol", gt, run(name, global_tol=gt))
    elif mode == "inexact":
        for name in DECKS:
            for seed in range(4):
                print(name, "Belos-tolerance solves, seed", seed, run(name, residual_tol=1e-6, seed=seed))
    elif mode == "variants":
        root = variant_copy()
        for v, what in ((0, "literal (2x2 norm, trace/3)"), (1, "norm with s_zz"), (2, "trace/2")):
            os.environ["ORC_V"] = str(v)
            print("notch2D_small_J2", what, run("notch2D_small_J2", oracle_root=root))


if __name__ == "__main__":
    main()
