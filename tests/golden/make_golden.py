"""Generates tests/golden/*.npz mesh fixtures and golden.json from the reference tree.

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden.py
The meshes are the reference's own regression meshes (test/mesh/{cube,notch,notch2D}),
read with calibr8_b200.meshio and stored as flat arrays; golden.json holds the
`regression: QoI` constants of test/primal/*.yaml.in with the deck settings that
produce them (file:line cited per entry) and the FD-drop constants of the
adjoint / VFM gradient-check decks.
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from calibr8_b200 import meshio  # noqa: E402

REF = "/root/reference/source/calibr8/test"

B3 = [[0, 0, "xmin", "0.0"], [0, 1, "ymin", "0.0"], [0, 2, "zmin", "0.0"]]
B2 = [[0, 0, "xmin", "0.0"], [0, 1, "ymin", "0.0"]]
HJ2 = dict(E=1000., nu=.25, K=100., Y=10., S=0., D=0., A=0., n=0.)
HILL2D = dict(E=1000., nu=.25, Y=2., S=10., D=2., R00=1., R11=1., R22=1., R01=1.)

DECKS = {
    "cube_elastic": dict(
        src="test/primal/cube_elastic.yaml.in:5-41", mesh="cube", global_type="mechanics",
        local_type="elastic", params=dict(E=1000., nu=.25, cte=1e-3, delta_T=10.), dbcs=B3,
        num_steps=1, global_max_iters=15, global_tol=1e-8, local_max_iters=0, local_tol=0.,
        J=5.00000000000000184e-3, rel_tol=1e-6),
    "cube_hyper_J2": dict(
        src="test/primal/cube_hyper_J2.yaml.in:5-49", mesh="cube", global_type="mechanics",
        local_type="hyper_J2", params=HJ2, dbcs=B3 + [[0, 1, "ymax", "0.01 * t"]],
        num_steps=10, global_max_iters=15, global_tol=1e-8, local_max_iters=30, local_tol=1e-12,
        J=1.57817536611772440e-02, rel_tol=1e-4),
    "notch_small_J2": dict(
        src="test/primal/notch_small_J2.yaml.in:5-52", mesh="notch", global_type="mechanics",
        local_type="small_hill",
        params=dict(E=1000., nu=.25, Y=2., S=10., D=2., R00=1., R11=1., R22=1., R01=1., R02=1., R12=1.),
        dbcs=B3 + [[0, 1, "ymax", "0.001 * t"]],
        num_steps=4, global_max_iters=15, global_tol=1e-8, local_max_iters=500, local_tol=1e-12,
        J=1.4622046563394649e-04, rel_tol=1e-4),
    "notch_hyper_J2": dict(
        src="test/primal/notch_hyper_J2.yaml.in:5-49", mesh="notch", global_type="mechanics",
        local_type="hyper_J2", params=HJ2, dbcs=B3 + [[0, 1, "ymax", "0.005 * t"]],
        num_steps=4, global_max_iters=15, global_tol=1e-8, local_max_iters=500, local_tol=1e-12,
        J=7.0080671510235862e-04, rel_tol=1e-4),
    "notch2D_small_J2": dict(
        src="test/primal/notch2D_small_J2.yaml.in:5-46", mesh="notch2D", global_type="mechanics",
        local_type="small_J2", params=dict(E=1000., nu=.25, K=100., Y=10., cte=0., delta_T=0.),
        dbcs=B2 + [[0, 1, "ymax", "0.001 * t"]],
        num_steps=8, global_max_iters=15, global_tol=1e-8, local_max_iters=500, local_tol=1e-12,
        J=6.55208497250819866e-03, rel_tol=1e-4),
    "notch2D_small_J2_plane_strain": dict(
        src="test/primal/notch2D_small_J2_plane_strain.yaml.in:5-49", mesh="notch2D",
        global_type="mechanics", local_type="small_hill_plane_strain", params=HILL2D,
        dbcs=B2 + [[0, 1, "ymax", "0.001 * t"]],
        num_steps=4, global_max_iters=30, global_tol=1e-8, local_max_iters=500, local_tol=1e-12,
        J=1.7664579853744898e-03, rel_tol=1e-4),
    "notch2D_small_J2_plane_stress": dict(
        src="test/primal/notch2D_small_J2_plane_stress.yaml.in:5-49", mesh="notch2D",
        global_type="mechanics_plane_stress", local_type="small_hill_plane_stress", params=HILL2D,
        dbcs=B2 + [[0, 1, "ymax", "0.001 * t"]],
        num_steps=4, global_max_iters=30, global_tol=1e-8, local_max_iters=500, local_tol=1e-12,
        J=2.2831790025047405e-03, rel_tol=1e-4),
    "notch2D_hyper_J2_plane_stress": dict(
        src="test/primal/notch2D_hyper_J2_plane_stress.yaml.in:5-48", mesh="notch2D",
        global_type="mechanics_plane_stress", local_type="hyper_J2_plane_stress",
        params=dict(E=1000., nu=.25, Y=2., S=10., D=2., A=0., n=0., K=0.),
        dbcs=B2 + [[0, 1, "ymax", "0.005 * t"]],
        num_steps=5, global_max_iters=30, global_tol=1e-8, local_max_iters=500, local_tol=1e-12,
        J=1.7493199283412385e-02, rel_tol=1e-4),
    "notch2D_hyper_J2_plane_strain": dict(
        src="test/primal/notch2D_hyper_J2_plane_strain.yaml.in:5-46", mesh="notch2D",
        global_type="mechanics", local_type="hyper_J2_plane_strain",
        params=dict(E=1000., nu=.25, K=100., Y=10., Y_inf=0., delta=0.),
        dbcs=B2 + [[0, 1, "ymax", "0.001 * t"]],
        num_steps=8, global_max_iters=15, global_tol=1e-8, local_max_iters=500, local_tol=1e-12,
        J=6.5626182813091150e-03, rel_tol=1e-4),
    # finite-strain Hill through the unrotated rate of deformation (src/hypo_hill.cpp, minitensor::polar_rotation)
    "notch_hypo_J2": dict(
        src="test/primal/notch_hypo_J2.yaml.in:5-52", mesh="notch", global_type="mechanics", local_type="hypo_hill",
        params=dict(E=1000., nu=.25, Y=2., S=10., D=2., R00=1., R11=1., R22=1., R01=1., R02=1., R12=1.),
        dbcs=B3 + [[0, 1, "ymax", "0.005 * t"]],
        num_steps=4, global_max_iters=15, global_tol=1e-8, local_max_iters=500, local_tol=1e-12,
        J=7.5441386985803955e-04, rel_tol=1e-4),
    "notch2D_hypo_J2_plane_strain": dict(
        src="test/primal/notch2D_hypo_J2_plane_strain.yaml.in:5-48", mesh="notch2D", global_type="mechanics",
        local_type="hypo_hill_plane_strain", params=HILL2D, dbcs=B2 + [[0, 1, "ymax", "0.005 * t"]],
        num_steps=4, global_max_iters=30, global_tol=1e-8, local_max_iters=500, local_tol=1e-12,
        J=7.10226176768509899e-03, rel_tol=1e-4),
    "notch2D_hypo_J2_plane_stress": dict(
        src="test/primal/notch2D_hypo_J2_plane_stress.yaml.in:5-52", mesh="notch2D",
        global_type="mechanics_plane_stress", local_type="hypo_hill_plane_stress",
        params=dict(HILL2D, Q00=1., Q01=0., Q10=0., Q11=1.), dbcs=B2 + [[0, 1, "ymax", "0.005 * t"]],
        num_steps=4, global_max_iters=30, global_tol=1e-8, local_max_iters=500, local_tol=1e-12,
        J=1.1852379652063684e-02, rel_tol=1e-4),
    # traction boundary condition (src/tbcs.cpp:17-98): [resid, side set, x-val, y-val, z-val]
    "cube_hyperelasticity_traction": dict(
        src="test/primal/cube_hyperelasticity_traction.yaml.in:5-51", mesh="cube", global_type="mechanics",
        local_type="hyper_J2", params=dict(E=1000., nu=.25, K=100., Y=100000., S=0., D=0., A=0., n=0.),
        dbcs=[[0, 0, "ymin", "0.0"], [0, 1, "ymin", "0.0"], [0, 2, "ymin", "0.0"]],
        tbcs=[[0, "ymax", "0.", "0.1 * t", "0."]],
        num_steps=4, global_max_iters=10, global_tol=1e-8, local_max_iters=30, local_tol=1e-12,
        J=1.61757374785081228e-04, rel_tol=1e-4),
}

FD_DROPS = {
    "notch2D_small_J2_adjoint_check": dict(
        src="test/adjoint/notch2D_small_J2_adjoint_check.yaml.in:38-40", log10_drop=7.7384790056517998, tol=1e-1),
    "vfm_sens_notch2D_small_J2_plane_stress": dict(
        src="test/vfm/vfm_adjoint_sens_notch2D_small_J2_plane_stress.yaml.in:37-39", log10_drop=7.6799236451528792, tol=1e-1),
}


def main():
    for name in ("cube", "notch", "notch2D"):
        m = meshio.load_calibr8_mesh(f"{REF}/mesh/{name}/{name}0.smb", f"{REF}/mesh/{name}/{name}.dmg",
                                     f"{REF}/mesh/{name}/{name}.txt")
        meshio.save_npz(m, os.path.join(HERE, f"mesh_{name}.npz"))
        # the reference's OWN two-part partition of the mesh (SCOREC split -> ParMETIS, test/mesh/*/Makefile):
        # element ownership as a fixture, for the partition / halo-plan / partitioned-solve tests
        ep = meshio.reference_partition(m, [f"{REF}/mesh/{name}/{name}_2p{k}.smb" for k in range(2)])
        import numpy as np
        np.save(os.path.join(HERE, f"partition_{name}_2p.npy"), ep.astype(np.int8))
        print(name, m.n_nodes, m.n_elems, "reference 2-part split:", np.bincount(ep).tolist())
    json.dump(dict(decks=DECKS, fd_drops=FD_DROPS), open(os.path.join(HERE, "golden.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
