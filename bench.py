#!/usr/bin/env python
"""bench.py -- the driver's measurement contract for the calibr8 hot path on B200.

    python bench.py --gpus N --steps K --warmup W            (our arm: CUDA through the C ABI)
    python bench.py --impl reference --gpus N --steps K ...  (CPU restatement of the reference)

Workload (BASELINE.json configs[1]): 3-D finite-strain hyper-J2, mixed u-p mechanics, on a
synthetic notched unit box of ~1M linear tets (56 cells/side, Kuhn split, quarter-circle notch
r=0.2 removed).  A "step" is one pass of eval_forward_jacobian (local Newton per quadrature
point + element residual/Jacobian + scatter into the BSR matrix) over the whole mesh at a
synthetic state that puts roughly half the points on the plastic branch.

metric  = QP residual+Jacobian evals/s (fp64)  [= elements / second; one coupled QP per element]
value   = kernel-path throughput with inputs resident in HBM (memset A,b + K1 per step)
e2e     = the same through c8_state_forward_jacobian with HOST nodal buffers
          (H2D of the Newton iterate, assembly, D2H of the residual + status every step)
N > 1   = STRONG scaling: the same mesh is split into N parts (recursive coordinate bisection, one rank
          per GPU); a part evaluates its owned + halo elements and only OWNED points are counted.
          K1 itself has no data-path collective; the forward+adjoint leg (same mesh, partitioned) uses
          NCCL halo copies and allreduces.  parity_vs_n1 compares every N > 1 run with the one-GPU
          values of tests/golden/bench_n1.json.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "QP residual+Jacobian evals/s (fp64)"
UNIT = "QP evals/s"
PARAMS = dict(E=1000., nu=.25, Y=10., S=0., D=0., A=0., n=0., K=100.)   # test/primal/notch_hyper_J2.yaml.in:25-34
LOCAL = dict(max_iters=500, abs_tol=1e-12, rel_tol=1e-12)             # same deck :20-24
LINEAR_TOL = 1e-6   # Belos "Convergence Tolerance" of the same deck (:59); Newton tolerances 1e-8 (:16-18)
# general-path rows: the iterated local Newton with exp / pow in the yield law, i.e. what a
# Voce / power-law calibration runs, and the Hill model of BASELINE configs[2] (anisotropic flow direction:
# its return map reduces to one scalar equation, models.cuh hill_return_map);
# name -> (local residual, parameters, displacement scale of the synthetic state)
GENERAL_PATHS = {
    "hyper_J2_general": ("hyper_J2", dict(E=1000., nu=.25, Y=10., S=10., D=2., A=1., n=.5, K=100.), 1.0),
    "small_hill": ("small_hill", dict(E=1000., nu=.25, Y=2., R00=1., R11=.9, R22=1.1, R01=1., R02=.95, R12=1.05,
                                      S=10., D=2.), 0.2),
    # finite-strain Hill in the unrotated frame (AD through the polar-rotation iteration), SURVEY 8(f) rank 4
    "hypo_hill": ("hypo_hill", dict(E=1000., nu=.25, Y=2., R00=1., R11=.9, R22=1.1, R01=1., R02=.95, R12=1.05,
                                    S=10., D=2.), 0.2),
}
AMP = 2.2e-3
N_CELLS = 56
NOTCH = 0.2


def workload_mesh(n_cells=N_CELLS):
    from calibr8_b200 import meshgen
    return meshgen.box_tets(n_cells, notch_radius=NOTCH)


def workload_fields(mesh, seed=0):
    from calibr8_b200 import meshgen
    rng = np.random.RandomState(seed + 7)
    base = meshgen.smooth_field(mesh, AMP, seed=seed)
    u1 = (1.45 * base).reshape(-1)
    u2 = (1.8 * base + meshgen.smooth_field(mesh, 0.2 * AMP, seed=seed + 1)).reshape(-1)
    p1 = rng.uniform(-1., 1., size=mesh.n_nodes)
    p2 = rng.uniform(-1., 1., size=mesh.n_nodes)
    return (u1, p1), (u2, p2)


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip().split(", "))

    def wait_first_sample(self, timeout):
        """block until the poller has printed its first line (it is past its start-up) or `timeout` s"""
        t0 = time.perf_counter()
        while self.proc and not self.rows and time.perf_counter() - t0 < timeout:
            time.sleep(0.02)

    def stop(self):
        if not self.proc:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        time.sleep(0.15)
        self.proc.terminate()
        try:                       # make sure the poller is gone before anything else is timed
            self.proc.wait(timeout=3)
        except Exception:
            self.proc.kill()
            self.proc.wait()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for k, nm in enumerate(names):
                    if r[4 + k].strip().lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        return dict(sm_mhz=float(np.median(sm)) if sm else None,
                    sm_max_mhz=float(np.max(mx)) if mx else None, samples=len(sm),
                    reasons=sorted(reasons))


# --------------------------------------------------------------------------------------
def cpu_sample(n_threads, n_cells_sample=10, reps=1):
    """Oracle (CPU restatement of the reference algorithm) on a bounded sample of the same
    workload: each thread assembles its own n_cells_sample^3*6-tet notched box at the same
    state (full eval_forward_jacobian incl. scatter into its private per-block CSR)."""
    from oracle.pyoracle import Oracle
    mesh = workload_mesh(n_cells_sample)
    (u1, p1), (u2, p2) = workload_fields(mesh)
    orcs = []
    for _ in range(n_threads):
        o = Oracle(mesh.dim, mesh.conn, mesh.coords, global_type="mechanics", local_type="hyper_J2",
                   params=[PARAMS], **LOCAL)
        xi0 = o.init_xi()
        rA = o.forward_jacobian([u1, p1], o.zeros_x(), xi0, xi0, assemble=False)
        orcs.append((o, rA["xi"]))
    times = []
    plastic = None
    for _ in range(reps):
        res = [None] * n_threads

        def work(k):
            o, xi1 = orcs[k]
            res[k] = o.forward_jacobian([u2, p2], [u1, p1], xi1, xi1)

        ths = [threading.Thread(target=work, args=(k,)) for k in range(n_threads)]
        t0 = time.perf_counter()
        for t in ths:
            t.start()
        for t in ths:
            t.join()
        times.append(time.perf_counter() - t0)
        plastic = float(res[0]["path"].mean())
    t = min(times)
    return dict(evals_per_s=mesh.n_elems * n_threads / t, seconds=t, n_elems=mesh.n_elems,
                plastic_fraction=plastic)


def flops_per_qp_reference():
    """Algorithmic flops of the reference algorithm (16-wide AD on every operation) per QP on this
    workload's state: counted once with the op-counting oracle build by tests/golden/make_flop_counts.py
     and read here as a constant."""
    with open(os.path.join(ROOT, "tests", "golden", "flop_counts.json")) as f:
        return float(json.load(f)["K1"])


def calibration_step(ctx, mesh, part, load_steps, world, dev):
    """forward + adjoint gradient wall time per load step (BASELINE.json metric, second half) for
    BASELINE configs[1]: 3-D hyper-J2 notched specimen, u_y(ymax) = 0.001 t over `load_steps` steps
    (20 in the config: 2 % stretch, the notch yields from about step 8 on), symmetry planes,
    average-displacement objective, adjoint gradient w.r.t. every model parameter.
    part is None on one GPU; otherwise this rank's part of the SAME mesh (strong scaling): NCCL halo
    copies of Krylov vectors / Newton iterates and fp64 allreduces, issued by the library."""
    import torch
    import torch.distributed as dist
    from calibr8_b200.capi import HostProblem
    ns = mesh.node_sets if part is None else part.node_sets
    hp = HostProblem(ctx)
    hp.add_dbc(0, 0, ns["xmin"], "0.0")
    hp.add_dbc(0, 1, ns["ymin"], "0.0")
    hp.add_dbc(0, 2, ns["zmin"], "0.0")
    hp.add_dbc(0, 1, ns["ymax"], "0.001 * t")
    hp.finalize_dbcs()
    hp.set_solver(15, 1e-8, 1e-8, gmres_restart=100, gmres_max_iters=20000, linear_tol=LINEAR_TOL)
    hp.set_qoi_avg_disp()

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    # untimed: the same forward + adjoint evaluation once (what the first objective evaluation of a calibration
    # loop pays and every later one does not): builds the multigrid hierarchy, loads every kernel and
    # allocates the primal / adjoint histories of all load steps (~3 GB of cudaMalloc at 20 steps, 0.8-1.9 s
    # in the first process on a fresh box -- measured: 201 / 259 ms per load step with a one-step warm-up
    # against 164 in a second process on the same box)
    hp.set_time(load_steps, 1.0)
    hp.primal_solve(); hp.adjoint_gradient()
    hp.profile(True)                      # phase timers (stream-synchronised around each phase)
    p0 = hp.profile(True)
    s0 = hp.stats()
    sync(); t0 = time.perf_counter()
    J = hp.primal_solve()
    sync(); t1 = time.perf_counter()
    g = hp.adjoint_gradient()
    sync(); t2 = time.perf_counter()
    s1 = hp.stats()
    p1 = hp.profile(True)
    tt = torch.tensor([t2 - t0, t1 - t0, t2 - t1], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    tot, fwd, adj = [float(v) for v in tt.tolist()]
    xi_last = hp.get_step(load_steps)[1]
    n_own = ctx.n_elems if part is None else part.n_owned_elems
    yielded = torch.tensor([float((xi_last[:n_own, -1] > 0).sum()), float(n_own)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(yielded)
    its = s1["linear_iters"] - s0["linear_iters"]
    asm = s1["assemblies"] - s0["assemblies"]
    out = {"metric": "forward+adjoint gradient wall-time/load step", "unit": "ms",
           "value": tot / load_steps * 1e3, "forward_ms_per_load_step": fwd / load_steps * 1e3,
           "adjoint_ms_per_load_step": adj / load_steps * 1e3, "load_steps": load_steps,
           "assemblies": asm, "krylov_iterations": its,
           "krylov_iterations_per_assembly": its / max(asm, 1),
           "objective": J, "gradient": [float(v) for v in g],
           "gradient_parameters": list(ctx.param_names),
           "yielded_fraction_last_step": float(yielded[0] / yielded[1]),
           "phase_seconds": {k: p1[k] - p0[k] for k in p1},
           "preconditioner": ctx.preconditioner_info(),
           "note": "BASELINE configs[1] at spec: Newton tol 1e-8 and GMRES rel tol 1e-6 as in the reference deck "
                   "(test/primal/notch_hyper_J2.yaml.in), GMRES(100), aggregation-AMG right preconditioner (fine-level "
                   "operator kept in fp32 inside the preconditioner only; the Krylov iteration, its residuals and the "
                   "converged solution are fp64); one untimed evaluation of the same objective + gradient first "
                   "(hierarchy, kernel load, history allocation: the timed one is a steady-state evaluation of a "
                   "calibration loop); phase timers on (they add stream synchronisations)"}
    if part is not None:
        out["scaling"] = "strong"
        out["partition"] = {"parts": world, "owned_elems_rank0": part.n_owned_elems,
                            "halo_elems_rank0": part.n_elems - part.n_owned_elems,
                            "ghost_nodes_rank0": part.n_nodes - part.n_owned_nodes,
                            "neighbours_rank0": int(part.nbr_rank.size)}
        out["comm_rank0"] = ctx.comm_stats()
        out["p2p_rank0"] = ctx.p2p_active()
    hp.close()
    return out


def kernel_rows(ctx, st, dfma_peak, hbm_peak):
    """Every other kernel of the path on the bench state (1 GPU): K2..K6 with their own roofline.
    K3/K4/K6 are fp64-bound AD sweeps (reference flop count from the op-counting oracle, golden fixture);
    K2/K5 evaluate in plain doubles and are gather-bound (algorithmic bytes: conn + nodal + state)."""
    import torch
    from calibr8_b200.capi import make_qoi
    fl = json.load(open(os.path.join(ROOT, "tests", "golden", "flop_counts.json")))
    x, xp, xi, xip, A, b = st["x"], st["xp"], st["xi"], st["xip"], st["A"], st["b"]
    n, nn = ctx.n_elems, ctx.n_nodes

    def timeit(fn, pre=lambda: None, reps=5):
        ts = []
        for k in range(reps + 2):
            pre()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            if k >= 2:
                ts.append(e0.elapsed_time(e1))
        return float(np.median(ts))

    dev = x.device
    g = ctx.alloc("xi"); f = torch.zeros(ctx.xi_ld * ctx.nx, dtype=torch.float64, device=dev)
    rhs = ctx.alloc("b"); z = ctx.alloc("x"); z.copy_(torch.randn(z.shape, dtype=z.dtype, device=dev, generator=torch.Generator(device=dev).manual_seed(1)))
    phi = ctx.alloc("xi")
    sc = torch.zeros(8, dtype=torch.float64, device=dev)
    grad = torch.zeros(64, dtype=torch.float64, device=dev)
    q = make_qoi("avg_disp")
    t = {}
    t["K2"] = timeit(lambda: ctx.global_residual(x, xp, xi, xip, b), lambda: b.zero_())
    t["K3"] = timeit(lambda: ctx.adjoint_jacobian(q, x, xp, xi, xip, g, f, A, rhs), lambda: (rhs.zero_(), g.zero_()))
    t["K4"] = timeit(lambda: ctx.adjoint_local(x, xp, xi, xip, z, phi, g, f))
    t["K5"] = timeit(lambda: ctx.qoi_value(q, x, xp, xi, xip, 0, sc), lambda: sc.zero_())
    t["K6"] = timeit(lambda: ctx.qoi_gradient(q, x, xp, xi, xip, z, phi, grad), lambda: grad.zero_())
    names = {"K2": "eval_global_residual", "K3": "eval_adjoint_jacobian (element kernel + transposed BSR gather)",
             "K4": "solve_adjoint_local", "K5": "eval_qoi", "K6": "eval_qoi_gradient"}
    nxi, nx = ctx.nxi, ctx.nx
    # algorithmic bytes per launch: connectivity + xi streams + nodal fields once (+ outputs)
    nodal = nn * (24 + 32 + 32)
    bytes_ = {"K2": n * (16 + 2 * nxi * 8) + nodal + nn * 32,
              "K5": n * 16 + nn * (24 + 32)}
    rows = {}
    for k in ("K2", "K3", "K4", "K5", "K6"):
        r = {"reference_fn": names[k], "ms": t[k], "M_qp_per_s": n / t[k] / 1e3}
        if k in bytes_:
            gbs = bytes_[k] / (t[k] * 1e-3) * 1e-9
            r["roofline"] = {"bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak,
                             "algorithmic_bytes_per_launch": bytes_[k],
                             "note": "plain-double evaluation (no AD in the reference either), bound by the nodal gather"}
        else:
            tf = fl[k] * n / (t[k] * 1e-3) * 1e-12
            r["roofline"] = {"bound": "fp64", "achieved": tf, "peak": dfma_peak, "unit": "TFLOP/s", "frac": tf / dfma_peak,
                             "algorithmic_flops_per_qp": fl[k], "flops_definition": "reference algorithm's count"}
        rows[k] = r
    return rows


def general_path_rows(mesh, fields, dev_index, stream, dfma_peak):
    """K1 on the states that take the GENERAL path of the local solve (iterated return map, exp / pow
    in the yield law): hyper-J2 with Voce + power-law hardening, and the Hill model of BASELINE configs[2]."""
    import torch
    from calibr8_b200.capi import Context
    fl = json.load(open(os.path.join(ROOT, "tests", "golden", "flop_counts.json")))
    (u1, p1), (u2, p2) = fields
    rows = {}
    for key, (ltype, params, amp_scale) in GENERAL_PATHS.items():
        c = Context(dev_index)
        c.set_mesh(mesh.dim, mesh.conn, mesh.coords)
        c.set_model("mechanics", ltype, params, **LOCAL)
        c.set_stream(stream.cuda_stream)
        x, xp, x0 = c.alloc("x"), c.alloc("x"), c.alloc("x")
        xi0, xip, xi = c.alloc("xi"), c.alloc("xi"), c.alloc("xi")
        A, b, path = c.alloc("A"), c.alloc("b"), c.alloc("path")
        c.pack_x(u2 * amp_scale, p2, x); c.pack_x(u1 * amp_scale, p1, xp)
        c.init_xi(xi0); c.init_xi(xip)
        assert c.forward_jacobian(xp, x0, xi0, xip, None, b) == 0
        ts = []
        for k in range(8):
            b.zero_(); xi.copy_(xip)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); c.forward_jacobian(x, xp, xip, xi, A, b, path, check=False); e1.record()
            torch.cuda.synchronize()
            if k >= 3:
                ts.append(e0.elapsed_time(e1))
        assert c.forward_jacobian(x, xp, xip, xi, None, None, path) == 0
        ms = float(np.median(ts))
        tf = fl["K1_" + key] * c.n_elems / (ms * 1e-3) * 1e-12
        rows[key] = {"local_residual": ltype, "params": params, "ms": ms, "M_qp_per_s": c.n_elems / ms / 1e3,
                     "plastic_fraction": float(path.to(torch.float32).mean().item()),
                     "reference_newton_iterations_per_plastic_point": fl["newton_iters_plastic_" + key],
                     "roofline": {"bound": "fp64", "achieved": tf, "peak": dfma_peak, "unit": "TFLOP/s",
                                  "frac": tf / dfma_peak, "algorithmic_flops_per_qp": fl["K1_" + key],
                                  "flops_definition": "reference algorithm's count"}}
        c.close()
        del x, xp, x0, xi0, xip, xi, A, b, path
        torch.cuda.empty_cache()
    return rows


def measured_traffic():
    """DRAM bytes per launch of the K1 kernels from the committed ncu --set full summary of this round
    (profiles/r02d_ncu_summary.json -- the final persistent element kernel --, written by tools/ncu_extract.py;
    earlier captures as the fallback); None when no file is there."""
    for name in ("r02d_ncu_summary.json", "r02b_ncu_summary.json", "r02_ncu_summary.json"):
        try:
            d = json.load(open(os.path.join(ROOT, "profiles", name)))
            ks = d["kernels"]
            el = next(v for k, v in ks.items() if k.startswith("k_forward_jacobian"))
            ga = next(v for k, v in ks.items() if k.startswith("k_bsr_gather"))
            return {"traffic": el["dram_bytes"] + ga["dram_bytes"], "element_kernel": el, "gather": ga,
                    "source": "profiles/" + name + " (" + d.get("command", "") + ")"}
        except Exception:
            continue
    return None


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n_cells_sample = 12
    for _ in range(min(args.warmup, 1)):
        cpu_sample(cores, n_cells_sample)
    t0 = time.perf_counter()
    vals = [cpu_sample(cores, n_cells_sample) for _ in range(args.steps)]
    wall = time.perf_counter() - t0
    v = float(np.mean([r["evals_per_s"] for r in vals]))
    sample = (f"{cores} threads x one {n_cells_sample}^3-cell notched box "
              f"({vals[0]['n_elems']} tets each) of the same hyper-J2 state per step, "
              f"full eval_forward_jacobian incl. CSR scatter")
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * float(np.mean([r["seconds"] for r in vals])),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": "3D finite-strain hyper_J2 mixed u-p, notched box tets (bounded sample)",
                   "note": "CPU restatement of the reference algorithm (oracle/), not the Trilinos binary: "
                           "the reference cannot be built offline (DESIGN.md)"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "wall_s": wall,
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from calibr8_b200 import partition
    from calibr8_b200.capi import Context

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    # the contract is ONE JSON line on stdout: native libraries (NCCL's version banner) write to fd 1
    # directly, so everything but the final line is sent to stderr
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)

    # ---- the SAME ~1M-tet mesh at every N (strong scaling): N > 1 splits it into N parts (recursive
    # coordinate bisection); a part assembles the complete rows of its owned nodes, so it evaluates its
    # owned elements plus the halo elements touching an owned node.  Only OWNED quadrature points count.
    mesh = workload_mesh()
    fields = workload_fields(mesh, seed=0)
    (u1, p1), (u2, p2) = fields
    part = None
    ctx = Context(local_rank)
    if world > 1:
        _, part = partition.partition_mesh(mesh, world, rank=rank)
        ctx.set_mesh(mesh.dim, part.conn, part.coords)
    else:
        ctx.set_mesh(mesh.dim, mesh.conn, mesh.coords)
    ctx.set_model("mechanics", "hyper_J2", PARAMS, **LOCAL)
    # one explicit (non-default) stream for torch's fills/copies, the CUDA-event timers and every
    # kernel of the library (CUDA-graph capture in the Krylov solver needs a real stream)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    if part is not None:
        ctx.set_partition(part)

        def bcast(raw):
            t = torch.zeros(128, dtype=torch.uint8, device=dev)
            if raw is not None:
                t.copy_(torch.tensor(list(raw), dtype=torch.uint8))
            dist.broadcast(t, 0)
            return bytes(t.cpu().tolist())
        ctx.nccl_init(rank, world, bcast)
        loc = lambda a, nc: part.localize_nodal(a, nc)
        u1, p1, u2, p2 = loc(u1, 3), loc(p1, 1), loc(u2, 3), loc(p2, 1)
    n_local = ctx.n_elems                                  # owned + halo elements of this rank
    n_owned = part.n_owned_elems if part is not None else ctx.n_elems
    n_total = mesh.n_elems                                 # = sum of the owned counts

    # --- untimed set-up: history state xi_prev from one assembly at the previous synthetic state
    x, xp, x0 = ctx.alloc("x"), ctx.alloc("x"), ctx.alloc("x")
    xi0, xip, xi = ctx.alloc("xi"), ctx.alloc("xi"), ctx.alloc("xi")
    A, b, path = ctx.alloc("A"), ctx.alloc("b"), ctx.alloc("path")
    ctx.pack_x(u2, p2, x); ctx.pack_x(u1, p1, xp)
    ctx.init_xi(xi0); ctx.init_xi(xip)
    nf = ctx.forward_jacobian(xp, x0, xi0, xip, None, b)
    assert nf == 0, "set-up assembly failed"
    b.zero_()

    def step():
        b.zero_()                          # A is overwritten by the two-phase assembly
        xi.copy_(xip)                      # Disc::create_primal: the step starts from step-1's state
        ctx.forward_jacobian(x, xp, xip, xi, A, b, path, check=False)

    # roofline denominators measured on this box with the same timer
    dfma_peak = ctx.bench_dfma(4096)
    copy_peak = ctx.bench_copy()

    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()
    nf = ctx.forward_jacobian(x, xp, xip, xi, None, None, path)   # status check outside the timing
    assert nf == 0
    pl = torch.tensor([float(path[:n_owned].to(torch.float32).sum().item()), float(n_owned)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(pl)
    plastic = float(pl[0] / pl[1])

    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    # The clock sampler is started and left to reach its polling loop BEFORE anything is timed (the start-up
    # of an nvidia-smi process attaches every GPU of the box and stalls launches for a moment: with one
    # process per rank that covered the whole 9 ms timed region of the 8-GPU run, 0.46 instead of 0.34 ms
    # per step), then the same step spins untimed for 0.3 s so that the samples show the clocks UNDER THIS
    # LOAD even when the timed K steps are shorter than one polling period; three more untimed steps after
    # the barrier, then the timed region.
    sampler = ClockSampler(local_rank)
    sampler.start()
    sampler.wait_first_sample(3.0)
    if world > 1:
        dist.barrier()
    t_spin = time.perf_counter()
    while time.perf_counter() - t_spin < 0.3:
        for _ in range(8):
            step()
        torch.cuda.synchronize()   # keeps the launch queue short: the loop then really lasts 0.3 s
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
           for _ in range(args.steps)]
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    for _ in range(3):
        step()
    ev[0].record()
    for k in range(args.steps):
        b.zero_(); xi.copy_(xip)
        kev[k][0].record()
        ctx.forward_jacobian(x, xp, xip, xi, A, b, path, check=False)
        kev[k][1].record()
        ev[k + 1].record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    clocks = sampler.stop()
    clocks["window"] = "0.3 s untimed spin of the same step + the timed steps (nvidia-smi -lms 100)"
    total_ms = ev[0].elapsed_time(ev[-1])
    k_ms = float(np.mean([a.elapsed_time(c) for a, c in kev]))

    # --- checksums of the assembled system over the OWNED rows (N > 1: compared with the one-GPU values
    # of tests/golden/bench_n1.json -> parity of the partitioned assembly, measured by the driver's own run)
    y = ctx.alloc("x")
    ctx.spmv(A, x, y)
    torch.cuda.synchronize()
    nrow = (part.n_owned_nodes if part is not None else ctx.n_nodes) * ctx.nb
    chk = torch.stack([(b[:nrow] * b[:nrow]).sum(), (y[:nrow] * y[:nrow]).sum()])
    if world > 1:
        dist.all_reduce(chk)
    checks = {"b_dot_b": float(chk[0]), "Ax_dot_Ax": float(chk[1])}

    # --- e2e: host nodal buffers through the resident-state entry point
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    hu, hp_ = pin(u2), pin(p2)
    hbu, hbp = torch.zeros(ctx.n_nodes * 3).double().pin_memory(), torch.zeros(ctx.n_nodes).double().pin_memory()
    xi1_host = ctx.unpack_xi(xip)
    ctx.state_set_prev(u1, p1, xi1_host)

    hun, hpn, hbun, hbpn = hu.numpy(), hp_.numpy(), hbu.numpy(), hbp.numpy()

    def e2e_step():
        # hyper_J2's local Newton starts from the trial state of xi_prev, so the resident
        # current-step xi needs no reset between repeated assemblies of the same iterate
        return ctx.state_forward_jacobian(hun, hpn, hbun, hbpn)

    for _ in range(2):
        e2e_step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0) / args.steps
    h2d = (ctx.n_nodes * 4) * 8
    d2h = (ctx.n_nodes * 4) * 8 + 4

    # --- K9 BSR SpMV on the matrix just assembled (HBM-bound): y = A x, 20 launches
    for _ in range(3):
        ctx.spmv(A, x, y)
    sev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    sev[0].record()
    for _ in range(20):
        ctx.spmv(A, x, y)
    sev[1].record()
    torch.cuda.synchronize()
    spmv_ms = sev[0].elapsed_time(sev[1]) / 20
    spmv_bytes = ctx.nnzb * (ctx.nb * ctx.nb * 8 + 4) + ctx.n_nodes * (4 + 2 * ctx.nb * 8)

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)

    # --- the other kernels of the path and the general-path rows (one GPU only; secondary legs must
    # not lose the headline line)
    kernels = general = None
    if world == 1 and not args.no_kernels:
        try:
            kernels = kernel_rows(ctx, dict(x=x, xp=xp, xi=xi, xip=xip, A=A, b=b), dfma_peak, hbm_peak)
        except Exception as ex:   # noqa: BLE001
            kernels = {"error": repr(ex)[:300]}

    # --- the second half of BASELINE.json's metric: forward load-step solves + adjoint objective
    # gradient through the C++ host solvers (Newton + line search, AMG-GMRES, reverse sweep) on the
    # same mesh and model, BASELINE configs[1]'s 20 load steps
    cal = None
    if not args.no_solve:
        try:
            cal = calibration_step(ctx, mesh, part, args.load_steps, world, dev)
        except Exception as ex:   # noqa: BLE001
            # (every rank sees the same allreduced norms, so a solver failure is raised on all ranks at
            # the same point of the collective sequence)
            cal = {"metric": "forward+adjoint gradient wall-time/load step", "error": repr(ex)[:300]}

    # max over ranks
    t = torch.tensor([total_ms, k_ms, e2e_s], dtype=torch.float64, device=dev)
    halo = torch.tensor([float(n_local - n_owned)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(halo)
    total_ms, k_ms, e2e_s = [float(v) for v in t.tolist()]
    ms_per_step = total_ms / args.steps
    value = n_total / (ms_per_step * 1e-3)
    e2e_value = n_total / e2e_s
    n_dev = ctx.n_nodes
    nnzb = ctx.nnzb
    del x, xp, x0, xi0, xip, xi, A, b, path, y
    ctx.close()
    torch.cuda.empty_cache()
    if world == 1 and not args.no_kernels:
        try:
            general = general_path_rows(mesh, fields, local_rank, stream, dfma_peak)
        except Exception as ex:   # noqa: BLE001
            general = {"error": repr(ex)[:300]}

    if rank == 0:
        flops_qp = flops_per_qp_reference()
        # per-launch figures are those of the slowest rank's kernels: its local elements, owned + halo
        achieved_tf = flops_qp * n_local / (k_ms * 1e-3) * 1e-12
        bytes_per_launch = (n_local * (16 + 64 + 64 + 64 + 64 + 1) + n_dev * (24 + 32 + 32 + 32) + nnzb * 16 * 8)
        traffic = measured_traffic() if world == 1 else None
        golden_path = os.path.join(ROOT, "tests", "golden", "bench_n1.json")
        parity = None
        if args.write_golden and world == 1 and cal and "objective" in cal:
            json.dump({"command": "python bench.py --write-golden (one B200)", "n_elems": n_total, "checks": checks,
                       "load_steps": cal["load_steps"], "objective": cal["objective"], "gradient": cal["gradient"]},
                      open(golden_path, "w"), indent=1)
        elif os.path.exists(golden_path):
            gold = json.load(open(golden_path))
            rel = lambda a, c: abs(a - c) / abs(c) if c != 0 else abs(a)
            parity = {"reference": "tests/golden/bench_n1.json (one-GPU run of this bench)",
                      "rel_b": rel(checks["b_dot_b"], gold["checks"]["b_dot_b"]) / 2,
                      "rel_Ax": rel(checks["Ax_dot_Ax"], gold["checks"]["Ax_dot_Ax"]) / 2}
            ok = parity["rel_b"] < 1e-10 and parity["rel_Ax"] < 1e-10
            if cal and "objective" in cal and cal["load_steps"] == gold["load_steps"]:
                g0, g1 = np.array(gold["gradient"]), np.array(cal["gradient"])
                parity["rel_J"] = rel(cal["objective"], gold["objective"])
                parity["rel_grad"] = float(np.abs(g1 - g0).max() / np.abs(g0).max())
                ok = ok and parity["rel_J"] < 1e-8 and parity["rel_grad"] < 1e-8
            parity["ok"] = bool(ok)
            parity["tolerances"] = "assembled residual / matrix action 1e-10, objective and adjoint gradient 1e-8 (north_star)"
        cpu = None
        if world == 1 and not args.no_cpu:
            cores = os.cpu_count() or 1
            ca = min([cpu_sample(cores, 12) for _ in range(2)], key=lambda r: r["seconds"])
            c1 = cpu_sample(1, 20)
            cpu = {"value": ca["evals_per_s"], "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": f"{cores} threads x one {ca['n_elems']}-tet notched box (12 cells/side) of the same hyper-J2 "
                             f"state, full eval_forward_jacobian incl. CSR scatter, {ca['seconds']:.2f} s",
                   "single_core": {"value": c1["evals_per_s"], "sample": f"one {c1['n_elems']}-tet box, {c1['seconds']:.1f} s"}}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {
                "workload": f"3D finite-strain hyper_J2 + mixed u-p mechanics, notched unit box, "
                            f"{N_CELLS} cells/side Kuhn tets (BASELINE configs[1]); the same mesh at every N",
                "n_elems": n_total, "n_elems_slowest_rank_incl_halo": n_local,
                "halo_elems_total": int(halo.item()), "n_nodes_rank0": n_dev, "nnz_blocks_4x4_rank0": nnzb,
                "plastic_fraction": plastic, "local_newton": LOCAL,
                "timed_region": "memset(b) + xi<-xi_prev copy + K1 (eval_forward_jacobian: element kernel + BSR gather) per step",
                "l2_policy": "inputs larger than L2 at N<=2 (A 4x4-BSR values %.0f MB + state + element scratch %.0f MB per rank)"
                             % (nnzb * 128 / 1e6, n_local * 2048 / 1e6),
                "parallelism": ("%d ranks, RCB element partition, owned QPs counted, halo elements evaluated redundantly, "
                                "no data-path collective in K1" % world) if world > 1 else "1 GPU",
            },
            "kernel_ms": k_ms,
            "roofline": {"bound": "fp64", "achieved": achieved_tf, "peak": dfma_peak, "unit": "TFLOP/s",
                         "frac": achieved_tf / dfma_peak,
                         "peak_source": "DFMA micro-benchmark measured in this run (c8_bench_dfma); "
                                        "MEASURED_PEAKS.json has no fp64 entry",
                         "algorithmic_flops_per_qp": flops_qp,
                         "flops_definition": "reference algorithm's count (16-wide AD on every op, 4-5 local Newton "
                                             "iterations per plastic point; op-counting oracle); the kernel executes fewer "
                                             "-- executed DFMA TFLOP/s: profiles/README.md",
                         "kernels": "k_forward_jacobian_persistent<Cfg<3,0,HyperJ2<3>,4>,true> + k_bsr_gather<4,4,false>",
                         "traffic": traffic["traffic"] if traffic else None,
                         "traffic_source": traffic["source"] if traffic else None},
            "roofline_hbm": {"bound": "hbm", "achieved": bytes_per_launch / (k_ms * 1e-3) * 1e-9,
                             "peak": hbm_peak, "unit": "GB/s",
                             "frac": bytes_per_launch / (k_ms * 1e-3) * 1e-9 / hbm_peak,
                             "algorithmic_bytes_per_launch": bytes_per_launch,
                             "copy_gbs_measured_this_run": copy_peak,
                             "traffic": traffic["traffic"] if traffic else None},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": e2e_s * 1e3,
                    "api": "c8_state_forward_jacobian (host nodal iterate in, host residual + status out)"},
            "gpu_launches": 2 * args.steps * world,   # element kernel + BSR gather per step and rank
            "clocks": clocks,
            "checksums": checks,
        }
        line["roofline_spmv"] = {"bound": "hbm", "achieved": spmv_bytes / (spmv_ms * 1e-3) * 1e-9,
                                 "peak": hbm_peak, "unit": "GB/s",
                                 "frac": spmv_bytes / (spmv_ms * 1e-3) * 1e-9 / hbm_peak,
                                 "kernel": "k_bsr_spmv<4>", "ms_per_launch": spmv_ms,
                                 "algorithmic_bytes_per_launch": spmv_bytes,
                                 "bytes_definition": "8 B/value + 4 B/block index + rowptr + x read + y write"}
        if parity:
            line["parity_vs_n1"] = parity
        if cal:
            line["forward_adjoint"] = cal
        if kernels:
            line["kernels"] = kernels
        if general:
            line["general_path"] = general
        if cpu:
            line["cpu_baseline"] = cpu
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-solve", action="store_true", help="skip the forward+adjoint load-step leg")
    ap.add_argument("--no-kernels", action="store_true", help="skip the K2..K6 and general-path rows")
    ap.add_argument("--load-steps", type=int, default=20, help="BASELINE configs[1]: 20 load steps")
    ap.add_argument("--write-golden", action="store_true",
                    help="one GPU: store the checksums / objective / gradient as tests/golden/bench_n1.json")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
