#!/usr/bin/env python
"""bench.py — the femgl Newton step (assembly + GMRES(30)/block-Jacobi + line search) on N B200s of one box.

    python bench.py --gpus N --steps K --warmup W            # this framework (CUDA, through the C ABI)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's own CPU code (oracle/_ref) on the host cores

Workload (default, every N): BASELINE.json configs[4] = the north-star target — ONE cube, Q1, global refinement 7,
38 640 402 DoFs, Morton-partitioned over the N GPUs (strong scaling: same problem, same iterates at every N).  It fits
one B200 because the operator is applied matrix-free (no 148 GB matrix).  At N=1 the line also carries "c2": the same
measurement on configs[1] (Q1 r5, 646 866 DoFs).  `--workload c2|c3|c5` or --refine/--global-refine/--degree select others.

A "step" is one Newton step of run.cc:214-218 (assemble_system, solve(tol), newton_iteration).  The run follows the
reference's stop rule (run.cc:234-250, converge accuracy 5e-6): when the residual drops below it the run is over and the
next step starts a new run from the initial condition, so all K timed steps are steps a real run executes.  Steps are timed
one by one with CUDA events on the library's stream (the reset to the initial condition is outside the timed region);
`ms_per_step` = (max over ranks of the summed step times) / K.

Prints ONE JSON line.  `value` = DoFs advanced one Newton step per second, whole job; `roofline` is the dominant kernel
(the operator apply inside GMRES, HBM-bound) timed live; `e2e` repeats the steps through the C ABI with pinned HOST buffers
in and out every step; `cpu_baseline` is the reference's literal term code on a bounded cell sample, `cpu_port` the
optimised CPU restatement (fair CPU number); for N > 1 `multi_gpu_parity` compares the first Newton steps with a single-GPU
context of the same global mesh built on rank 0.
"""
import argparse
import json
import multiprocessing as mp
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

K123 = 0.42072
# Matep(p = 25 bar, t = 0.5, SCC on): tests/golden/matep.json, SURVEY.md App. B KAT-1
COEF = dict(alpha=-0.5, beta=(-0.010850915879921348, 0.020598836429398658, 0.02117724292551364, 0.019780381922869742,
                              -0.023091499302424053), gapB=3.9900156313155422, bt=2.0)
LIN_TOL = 1e-1       # "Cycle 0 linear solver tol" default, declare.cc:203
LS_STEP = 0.83       # "primary step length of dampped newton iteration" default, declare.cc:287
MAX_LIN_IT = 10000   # declare.cc:277
RESTART = 30         # SolverFGMRES default max_basis_size
CONVERGE_ACC = 5e-6  # "converge accuracy" default, declare.cc:249 (run.cc:234-250)
MAX_NEWTON = 41      # "Number of interations" default 40, inclusive upper bound (run.cc:207)

# vh_mg_params of the multigrid runs (--mg-pre / --mg-post).  V(0,1): the right-hand side is restricted directly, so an inner
# GMRES step costs two fine operator applies instead of the three of the library default V(1,1).  Measured at C5 on one B200
# (profiles/r02g_mg_cycle_variants.txt): 107 ms per Newton step and 9 steps to converge against 126 ms and 8 steps.
MG_PARAMS = {"pre": 0, "post": 1}
WORKLOADS = {  # name -> (degree, refine per GPU (weak) or None, global refine (strong) or None)
    "c2": (1, 5, None),   # BASELINE configs[1]; for N > 1: N root cubes stacked along z (weak)
    "c3": (2, None, 5),   # BASELINE configs[2]
    "c5": (1, None, 7),   # BASELINE configs[4], the north-star target
}


def coef_vector():
    return np.array([K123, K123, K123, COEF["alpha"], *COEF["beta"], COEF["bt"]])


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """Polls nvidia-smi (one-shot queries, ~5 Hz) from a thread while the timed region runs."""

    def __init__(self, gpu_index):
        self.rows = []
        self.gpu = gpu_index
        self.family = "clocks_event_reasons"
        self._stop = threading.Event()
        self._thr = None

    def _query(self):
        f = self.family
        return ("clocks.sm,clocks.max.sm,power.draw,%s.hw_slowdown,%s.hw_thermal_slowdown,%s.sw_thermal_slowdown,"
                "%s.sw_power_cap" % (f, f, f, f))

    def _once(self):
        p = subprocess.run(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self._query(), "--format=csv,noheader,nounits"],
                           capture_output=True, text=True, timeout=20)
        return p.returncode, p.stdout.strip()

    def _loop(self):
        while not self._stop.is_set():
            try:
                rc, out = self._once()
                if rc == 0 and out:
                    self.rows.append([s.strip() for s in out.splitlines()[0].split(",")])
            except Exception:
                pass
            self._stop.wait(0.15)

    def start(self):
        try:
            rc, out = self._once()
            if rc != 0 or "not a valid" in out.lower():
                self.family = "clocks_throttle_reasons"
            self._thr = threading.Thread(target=self._loop, daemon=True)
            self._thr.start()
        except Exception:
            self._thr = None

    def stop(self):
        if self._thr is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self._stop.set()
        self._thr.join(timeout=30)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def build_mesh(n_gpus, refine, degree, global_refine=None):
    import verkko_hem_repo_b200 as vh
    L = 20.0  # "cube half side length" default (declare.cc:159)
    if global_refine is not None:
        # strong scaling (BASELINE configs[4]): ONE cube refined `global_refine` times, Morton-partitioned over the ranks
        m = vh.Mesh(degree, [-L, -L, -L], [L, L, L], base=(1, 1, 1), face_bid=(1, 1, 1, 1, 4, 4), n_global_refine=global_refine)
        m.finalize(n_gpus)
        return m
    # weak scaling: N root cubes stacked along z, each refined `refine` times; Morton order keeps each root
    # contiguous, so rank r owns root r (the p4est partition of a 1 x 1 x N brick).
    m = vh.Mesh(degree, [-L, -L, -L], [L, L, -L + 2 * L * n_gpus], base=(1, 1, n_gpus), face_bid=(1, 1, 1, 1, 4, 4),
                n_global_refine=refine)
    m.finalize(n_gpus)
    return m


def problem_size(degree, refine, global_refine, n_gpus):
    """(n_dofs, n_cells) of the workload without building it (the reference arm needs no mesh)."""
    if global_refine is None:
        n_side, n_z = 2 ** refine, 2 ** refine * n_gpus
    else:
        n_side = n_z = 2 ** global_refine
    return 18 * (degree * n_side + 1) ** 2 * (degree * n_z + 1), n_side * n_side * n_z


def workload_string(degree, refine, global_refine, n_dofs, n_cells):
    """config.workload — identical in both arms (the driver compares the two lines)."""
    return ("femgl 3D cube Q%d, global refinement %s (%d DoFs total, %d cells), B-phase IC, "
            "z walls AdGR diffuse (bt=2), p=25 bar t=0.5 SCC on, linear tol 1e-1, damped Newton 0.83, converge accuracy 5e-6"
            % (degree, ("%d per GPU" % refine) if global_refine is None else ("%d, one cube split over the GPUs" % global_refine),
               n_dofs, n_cells))


def config_dict(degree, refine, global_refine, n_dofs, n_cells, n_gpus):
    """The `config` object of the JSON line - identical in both arms (ours and --impl reference) for the same command line."""
    return {"workload": workload_string(degree, refine, global_refine, n_dofs, n_cells),
            "l2": "the working set of one operator apply (the H_q tables: 11.5 KB per Q1 cell, 38.9 KB per Q2 cell, on every rank) exceeds "
                  "the 126 MB L2 in the Newton steps; the stand-alone kernel timings flush L2 between launches",
            "parallelism": "subdomain x%d (Morton partition, NCCL halo + peer-memory all-reduce)" % n_gpus,
            "stop_rule": "run.cc:234-250: a run ends at residual <= 5e-6, the next timed step starts a new run from the IC"}


def initial_state(T):
    from helpers import b_phase_state, MATEP_SCC_ON
    return b_phase_state(T, MATEP_SCC_ON, noise=0.0)


def newton_step(ctx):
    """run.cc:214-218 + iteration.cc:128-210 through the C ABI.  Returns (lin_its, n_trials, res_norm, rhs_norm)."""
    bn = ctx.assemble()
    its, _ = ctx.solve(LIN_TOL, MAX_LIN_IT, RESTART)
    n_trials = 0
    cur = bn
    for i in range(100):
        ctx.line_search_trial(LS_STEP ** i)
        cur = ctx.residual()
        n_trials += 1
        if cur < bn:
            break
    ctx.accept_trial()
    return its, n_trials, cur, bn


def run_steps(ctx, x0, steps, host_buffer=None):
    """`steps` Newton steps under the reference's stop rule (run.cc:234-250): a run ends when the residual is <= the converge
    accuracy (or after MAX_NEWTON steps) and the next step starts a new run from the initial condition x0.  Each step is
    timed on its own with CUDA events on the library's stream; resets are not timed.  host_buffer (pinned): the e2e variant
    — the step's input goes host->device and its result device->host inside the timed region.
    Returns (summed ms, list of runs, each a list of (its, trials, residual, rhs_norm))."""
    runs, cur_run, total_ms = [], [], 0.0
    if host_buffer is None:
        ctx.set_solution(x0)
    else:
        host_buffer[:] = x0
    for _ in range(steps):
        ctx.timer_start()
        if host_buffer is not None:
            ctx.set_solution(host_buffer)        # H2D of the step's input
        h = newton_step(ctx)
        if host_buffer is not None:
            ctx.get_solution(out=host_buffer)    # D2H of the step's result, straight into the pinned buffer
        total_ms += ctx.timer_stop()
        cur_run.append(h)
        if h[2] <= CONVERGE_ACC or len(cur_run) >= MAX_NEWTON:
            runs.append(cur_run)
            cur_run = []
            if host_buffer is None:
                ctx.set_solution(x0)
            else:
                host_buffer[:] = x0
    if cur_run:
        runs.append(cur_run)
    return total_ms, runs


def summarize_runs(runs):
    """Compact keys that survive a truncated log: totals over the timed steps and the first run's history."""
    steps = [h for r in runs for h in r]
    first = runs[0] if runs else []
    return {"newton_steps": len(steps), "gmres_its_total": int(sum(h[0] for h in steps)),
            "line_search_trials_total": int(sum(h[1] for h in steps)),
            "runs_completed": int(sum(1 for r in runs if r and r[-1][2] <= CONVERGE_ACC)),
            "steps_to_converge": (len(first) if first and first[-1][2] <= CONVERGE_ACC else None),
            "first_run_gmres_its": [int(h[0]) for h in first], "first_run_trials": [int(h[1]) for h in first],
            "first_run_residuals": [float("%.6e" % h[2]) for h in first]}


# ------------------------------------------------------------------------------------------
# CPU legs (the only places that touch oracle/): reference arm, cpu_baseline, cpu_port
# ------------------------------------------------------------------------------------------
def _ref_worker(args):
    seed, n_cells_each, h, coef = args
    import femgl_oracle as O
    rng = np.random.default_rng(seed)
    amp = COEF["gapB"] * float(np.float32(0.577350269))
    t0 = time.time()
    t_mat = t_res = 0.0
    for _ in range(n_cells_each):
        U = np.zeros((8, 18))
        U[:, [0, 4, 8]] = amp
        U += 0.05 * COEF["gapB"] * rng.uniform(-1, 1, U.shape)
        t = time.time()
        O.ref_cell(1, [0, 0, 0], h, U.ravel(), coef, [(4, 4)], want_matrix=True)   # assemble_system cell
        t_mat += time.time() - t
        t = time.time()
        O.ref_cell(1, [0, 0, 0], h, U.ravel(), coef, [(4, 4)], want_matrix=False)  # one compute_residual cell
        t_res += time.time() - t
    return time.time() - t0, t_mat, t_res


def reference_sample(h_cell, cells_per_core=1):
    """Time the reference's literal (q,i,j) loops + term functions on `cores * cells_per_core` Q1 cells, one worker
    process per host core (the reference is 1 thread per MPI rank, sol/src/main.cc:108)."""
    import femgl_oracle as O
    if not O.have_ref():
        return None
    cores = os.cpu_count() or 1
    coef = coef_vector()
    t0 = time.time()
    with mp.get_context("fork").Pool(cores) as pool:
        parts = pool.map(_ref_worker, [(1000 + k, cells_per_core, [h_cell] * 3, coef) for k in range(cores)])
    wall = time.time() - t0
    n_cells = cores * cells_per_core
    t_mat = sum(p[1] for p in parts) / cores   # mean per-worker seconds spent in the Jacobian / residual cells
    t_res = sum(p[2] for p in parts) / cores
    return dict(cores=cores, cells=n_cells, wall_s=wall, cells_per_s=n_cells / wall,
                assembly_cells_per_s=n_cells / max(t_mat, 1e-9), residual_cells_per_s=n_cells / max(t_res, 1e-9))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    p = args.degree
    n_dofs, n_cells = problem_size(p, args.refine, args.global_refine, args.gpus)
    n_side = 2 ** (args.refine if args.global_refine is None else args.global_refine)
    h_cell = 40.0 / n_side
    for _ in range(args.warmup if args.warmup < 2 else 1):
        reference_sample(h_cell, 1)
    t_list, last, asm_cps, res_cps = [], None, [], []
    for _ in range(args.steps):
        last = reference_sample(h_cell, 2)
        if last is None:
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libvhref.so not built"}))
            return
        t_list.append(last["wall_s"])
        asm_cps.append(last["assembly_cells_per_s"])
        res_cps.append(last["residual_cells_per_s"])
    cps = last["cells"] * len(t_list) / sum(t_list)
    scale = 1.0
    q2_note = ""
    if p == 2:
        # the sample is timed on Q1 cells (one Q2 cell takes ~100 s in the reference's loops); a Q2 cell visits
        # (486^2 * 27) / (144^2 * 8) = 38.4x as many (q,i,j) triples, each the same work (assemble.cc:188-252)
        scale = (486.0 ** 2 * 27.0) / (144.0 ** 2 * 8.0)
        q2_note = "; Q2 rate = Q1 rate / 38.4 (ratio of (q,i,j) triples per cell)"
    cps /= scale
    # one Newton step of the reference = one assembly + >= 1 residual evaluation over all cells; its ML-AMG solve is
    # not reproducible here and is left out, so this is an UPPER bound on the reference's throughput.
    step_s = n_cells / cps
    val = n_dofs / step_s
    dofs_per_cell = n_dofs / n_cells
    out = {"impl": "reference", "metric": "femgl Newton-step throughput",
           "value": val, "unit": "DoF/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": step_s * 1e3, "higher_is_better": True, "scaling": "weak" if args.global_refine is None else "strong",
           "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": config_dict(args.degree, args.refine, args.global_refine, n_dofs, n_cells, args.gpus),
           "reference_arm": "the reference's own CPU code for this path (oracle/_ref: its verbatim cell_mat_vec term files "
                            "driven by its literal (q,i,j) loops) on all host cores; %d-cell sample per step, extrapolated "
                            "linearly to the workload; its ML-AMG solve is left out (upper bound on its throughput)" % last["cells"],
           # NOT like-for-like with the GPU arm's Newton step (which includes the GMRES solve): compare phase by phase
           "comparable": False,
           "phases": {"assembly_dofs_per_s": float(np.mean(asm_cps)) / scale * dofs_per_cell,
                      "residual_dofs_per_s": float(np.mean(res_cps)) / scale * dofs_per_cell,
                      "note": "per-phase rates of the reference's literal loops on all cores (sampled cells, extrapolated): compare "
                              "with the GPU line's assembly.dofs_per_s and kernels.residual_ms; the solve is not in this arm"},
           "cpu_baseline": {"value": val, "unit": "DoF/s", "cores": last["cores"], "kind": "reference",
                            "sample": "%d Q1 cells/step (Jacobian + 1 residual) by the reference's verbatim cell_mat_vec "
                                      "term files driven by its literal (q,i,j) loops; %.2f cells/s on %d cores, "
                                      "extrapolated linearly to %d cells%s" % (last["cells"], cps, last["cores"], n_cells, q2_note)},
           "e2e": {"value": val, "unit": "DoF/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out))


def cpu_port_numbers(its_per_step, trials_per_step, n_cells_workload, n_dofs_workload):
    """The optimised CPU restatement (oracle O2: closed-form H_q in C with one pthread per core, scipy BSR SpMV, numpy
    vectors) on a BOUNDED sample — a Q1 r4 cube (4 096 cells, 88 434 DoFs) with the workload's coefficients — scaled
    linearly in cells/DoFs to the workload.  The "fair CPU" numbers SURVEY.md section 8(d) / BASELINE.md section 3 ask for:
    assembly DoF/s, SpMV GB/s and a full Newton step with the same GMRES(30)+block-Jacobi iteration count as the GPU run."""
    import femgl_oracle as O
    import scipy.sparse as sp
    import verkko_hem_repo_b200 as vh
    from helpers import MATEP_SCC_ON, b_phase_state
    T = vh.Mesh(1, [-20.0] * 3, [20.0] * 3, base=(1, 1, 1), face_bid=(1, 1, 1, 1, 4, 4), n_global_refine=4).finalize(1).tables(0)
    coef = coef_vector()
    x = b_phase_state(T, MATEP_SCC_ON, noise=0.0)
    fptr, fno, fbid = T.face_csr()
    dummy = np.zeros(1, np.int32)
    t0 = time.perf_counter()
    O.cells(1, T.cell_nodes, T.cell_origin, T.cell_h, x, coef, fptr, fno if fno.size else dummy, fbid if fbid.size else dummy,
            want_matrix=True)
    t_cells = time.perf_counter() - t0
    t0 = time.perf_counter()
    O.cells(1, T.cell_nodes, T.cell_origin, T.cell_h, x, coef, fptr, fno if fno.size else dummy, fbid if fbid.size else dummy,
            want_matrix=False)
    t_res = time.perf_counter() - t0
    t0 = time.perf_counter()
    A, rhs = O.assemble_global(T, x, coef, True)       # cells + constrained scatter (scipy)
    t_asm = time.perf_counter() - t0
    B = sp.bsr_matrix(A, blocksize=(18, 18))
    z = np.random.default_rng(3).uniform(-1, 1, A.shape[1])
    B @ z
    t0 = time.perf_counter()
    for _ in range(5):
        B @ z
    t_spmv = (time.perf_counter() - t0) / 5
    spmv_bytes = 8 * 324 * B.indices.size + 4 * B.indices.size + 4 * B.indptr.size + 16 * A.shape[0]
    Minv = O.block_jacobi_inverse(A, T.n_owned_nodes)
    t0 = time.perf_counter()
    _, its_s, _, _ = O.gmres_block_jacobi(A, rhs, Minv, LIN_TOL * np.linalg.norm(rhs))
    t_solve = time.perf_counter() - t0
    per_it = t_solve / max(its_s, 1)
    sc = n_cells_workload / T.n_cells
    step_s = sc * (t_cells + max(t_asm - t_cells, 0.0) + its_per_step * per_it + trials_per_step * t_res)
    return {"kind": "port", "cores": os.cpu_count(),
            "sample": "Q1 r4 cube (%d cells, %d DoFs): cell matrices by oracle/femgl_oracle.c (closed-form H_q, one pthread per "
                      "core) %.2f s, constrained scatter (scipy, 1 core) %.2f s, residual cells %.3f s, BSR18 SpMV (scipy, 1 core) "
                      "%.3f s, GMRES(30)+block-Jacobi %.3f s per iteration (numpy); scaled x%.0f in cells to the workload with the "
                      "GPU run's %.1f linear iterations and %.1f residual evaluations per Newton step"
                      % (T.n_cells, 18 * T.n_owned_nodes, t_cells, max(t_asm - t_cells, 0.0), t_res, t_spmv, per_it, sc, its_per_step,
                         trials_per_step),
            "assembly_dofs_per_s": 18 * T.n_owned_nodes / t_cells, "assembly_with_scatter_dofs_per_s": 18 * T.n_owned_nodes / t_asm,
            "residual_dofs_per_s": 18 * T.n_owned_nodes / t_res,
            "spmv_gbs": spmv_bytes / t_spmv / 1e9, "newton_step_ms": step_s * 1e3, "newton_step_dofs_per_s": n_dofs_workload / step_s}


# ------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------
def kernel_numbers(ctx, T, info, degree, hbm_peak, peak_src, traffic_key):
    """Live kernel timings (CUDA events, L2 flushed between launches) of rank 0's kernels and the roofline of the dominant one."""
    nnzb, nb = info["nnzb"], T.n_owned_nodes
    n = 8 if degree == 1 else 27
    n_fast_blocks = info["n_packed_blocks"]
    mf_mode = int(info.get("spmv_matrix_free", 0))
    # algorithmic bytes of one operator apply (BSR18, every 18x18 block stored in full)
    spmv_alg = 8 * 324 * nnzb + 4 * nnzb + 4 * (nb + 1) + 16 * 18 * nb
    # bytes this implementation really moves per apply:
    #   matrix-free: the packed H_q tables (8*180*n_q B per cell) + the cell gather (8*dpc) + the cell products written and
    #   re-read once (2 * 8*dpc) + y; constrained rows (full blocks) as stored
    slow = 8 * 324 * (nnzb - n_fast_blocks)
    mf_moved = 8 * 180 * n * T.n_cells + 3 * 8 * 18 * n * T.n_cells + slow + 8 * 18 * nb
    packed_moved = 8 * 180 * n_fast_blocks + slow + 4 * nnzb + 4 * (nb + 1) + 16 * 18 * nb
    moved = mf_moved if mf_mode else packed_moved
    t_apply = ctx.time_kernel(0, reps=10, flush_l2=True)
    t_asm = ctx.time_kernel(1, reps=3, flush_l2=True)
    t_pw = ctx.time_kernel(5, reps=3, flush_l2=True)
    t_res = ctx.time_kernel(2, reps=3, flush_l2=True)
    t_bj = ctx.time_kernel(3, reps=5, flush_l2=True)
    fp64_peak = ctx.measure_fp64_peak()
    traffic = None
    try:  # DRAM bytes per launch measured once with `ncu --set full` for this exact workload (profiles/traffic.json)
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            tj = json.load(f)
        kname = "k_points_apply" if mf_mode else "k_spmv_sym18"
        if traffic_key in tj and kname in tj[traffic_key]:
            traffic = tj[traffic_key][kname]["read_bytes"] + tj[traffic_key][kname]["write_bytes"]
    except Exception:
        traffic = None
    gbs = moved / (t_apply * 1e-3) / 1e9
    # frac: DRAM bytes ncu counted for this exact workload (traffic) when a capture exists, else the model of moved bytes
    real = traffic if traffic is not None else moved
    roof = {"bound": "hbm", "kernel": ("k_points<APPLY>+k_gather_apply" if mf_mode else "k_spmv_sym18") if n_fast_blocks else "k_spmv_bsr18",
            "achieved": real / (t_apply * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s", "frac": real / (t_apply * 1e-3) / 1e9 / hbm_peak,
            "traffic": traffic, "peak_source": peak_src, "ms_per_launch": t_apply, "moved_bytes_per_launch": moved,
            "achieved_moved_model": gbs, "frac_moved_model": gbs / hbm_peak,
            "achieved_algorithmic": spmv_alg / (t_apply * 1e-3) / 1e9, "algorithmic_bytes_per_launch": spmv_alg,
            "note": "achieved/frac = DRAM bytes of one operator apply (ncu capture of this workload: traffic; without a capture the "
                    "model of the bytes this implementation moves: moved_bytes_per_launch) / live CUDA-event time (/ measured HBM "
                    "peak); achieved_algorithmic uses SURVEY 8(d)'s BSR18 bytes (every 18x18 block streamed in full; the matrix "
                    "is never formed, so that figure exceeds the peak)"}
    asm_flops = 2.0 * n * n * n * 336 * T.n_cells
    asm = {"ms": t_asm, "dofs_per_s": 18 * nb / (t_asm * 1e-3), "pointwise_ms": t_pw,
           "note": "Jacobian phase of vh_assemble in the matrix-free default: pointwise kernel (H_q tables, cell rhs) + diagonal "
                   "blocks for block-Jacobi + rhs gather; the lattice rows are never assembled",
           "algorithmic_flops": asm_flops, "tflops_fp64_algorithmic": asm_flops / (t_asm * 1e-3) / 1e12,
           "fp64_peak_tflops_measured": fp64_peak,
           "frac_fp64_algorithmic": asm_flops / (t_asm * 1e-3) / 1e12 / fp64_peak if fp64_peak else None}
    kern = {"apply_ms": t_apply, "residual_ms": t_res,
            "residual_gbs": (8 * 18 * n * T.n_cells + 8 * 18 * nb) / (t_res * 1e-3) / 1e9,
            "block_jacobi_apply_ms": t_bj, "block_jacobi_apply_gbs": (2592 * nb + 16 * 18 * nb) / (t_bj * 1e-3) / 1e9,
            "operator_apply_mode": ("packed-spmv", "matrix-free", "table-free", "matrix-free-v2")[mf_mode]}
    return roof, asm, kern


def build_context(vh, dist, rank, world, local_rank, degree, refine, global_refine, precond, mg_coarsest):
    """Context of the workload on this rank; with precond == "mg" also the coarser levels of the multigrid hierarchy (the same
    box refined fewer times, same partition) attached to it.  Returns (mesh, tables, ctx, levels, create seconds)."""
    mesh = build_mesh(world, refine, degree, global_refine)
    T = mesh.tables(rank)
    t0 = time.perf_counter()
    ctx = vh.Context(T, device=local_rank)

    def init(c):
        if world > 1:
            uid = [vh.Context.nccl_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(uid, src=0)
            c.comm_init(rank, world, uid[0])
        c.set_coef_vector(coef_vector())

    init(ctx)
    levels = [ctx]
    if precond == "mg":
        top = global_refine if global_refine is not None else refine
        mf, Tf, cf = mesh, T, ctx
        for lv in range(top - 1, mg_coarsest - 1, -1):
            mc = build_mesh(world, lv if global_refine is None else None, degree, lv if global_refine is not None else None)
            Tc = mc.tables(rank)
            cc = vh.Context(Tc, device=local_rank)
            init(cc)
            err = None
            try:
                cf.mg_attach(cc, *vh.mg_prolongation(mf, Tf, mc, Tc))
            except (vh.VhError, RuntimeError) as exc:   # vh_mg_attach validates locally: agree before the next collective call
                err = str(exc)
            if world > 1:
                import torch
                flag = torch.tensor([0.0 if err is None else 1.0], device="cuda")
                dist.all_reduce(flag)
                if float(flag.item()) > 0 and err is None:
                    err = "vh_mg_attach failed on another rank"
            if err is not None:
                raise vh.VhError(-2, "multigrid hierarchy rejected: " + err)
            levels.append(cc)
            mf, Tf, cf = mc, Tc, cc
        ctx.set_preconditioner("multigrid", **MG_PARAMS)
    return mesh, T, ctx, levels, time.perf_counter() - t0


def measure(vh, torch, dist, rank, world, local_rank, degree, refine, global_refine, steps, warmup, want_kernels, want_parity,
            precond="bj", mg_coarsest=5):
    """Build the workload on `world` ranks and measure it.  Returns the dict of results (rank 0) or None."""
    mesh, T, ctx, levels, create_s = build_context(vh, dist, rank, world, local_rank, degree, refine, global_refine, precond, mg_coarsest)
    x0 = initial_state(T)[:18 * T.n_owned_nodes]
    n_dofs = 18 * mesh.n_nodes
    info = ctx.info()

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if dist is None:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- warm-up ----
    run_steps(ctx, x0, warmup)
    # ---- timed: K Newton steps, state resident in HBM ----
    ctx.timers(reset=True)
    barrier()
    ms, runs = run_steps(ctx, x0, steps)
    barrier()
    tm = ctx.timers()
    ms = max_over_ranks(ms)
    # GL free energy at the end of the first run (outside the timed region)
    try:
        final_energy = ctx.energy(0)
    except Exception:
        final_energy = None
    # ---- e2e: same steps, host buffers in and out of the C ABI every step (pinned host memory) ----
    xh = torch.from_numpy(x0.copy()).pin_memory().numpy()
    barrier()
    t0 = time.perf_counter()
    ms_e2e, _ = run_steps(ctx, x0, steps, host_buffer=xh)
    wall_e2e = time.perf_counter() - t0
    ms_e2e = max_over_ranks(ms_e2e)
    barrier()

    # ---- the north star's preconditioner next to the multigrid headline: the same context with nodal block-Jacobi ----
    bj = None
    if precond == "mg":
        ctx.set_preconditioner("block-jacobi")
        n_bj = min(steps, 5)
        run_steps(ctx, x0, 1)
        barrier()
        ms_bj, runs_bj = run_steps(ctx, x0, n_bj)
        barrier()
        ms_bj = max_over_ranks(ms_bj)
        s_bj = summarize_runs(runs_bj)
        bj = {"preconditioner": "nodal 18x18 block-Jacobi (north star)", "steps": n_bj, "ms_per_step": ms_bj / n_bj,
              "gmres_its_per_step": s_bj["gmres_its_total"] / n_bj, "first_run_gmres_its": s_bj["first_run_gmres_its"],
              "first_run_residuals": s_bj["first_run_residuals"]}
        ctx.set_preconditioner("multigrid", **MG_PARAMS)

    summ = summarize_runs(runs)
    its_total = max(summ["gmres_its_total"], 1)
    res = {"n_dofs": n_dofs, "n_cells": mesh.n_cells, "ms_per_step": ms / steps, "value": n_dofs / (ms / steps * 1e-3),
           "newton": summ, "final_energy": final_energy,
           "phase_ms_per_step": {k: tm[k] / steps for k in ("assemble", "residual", "solve", "vector", "precond_setup")},
           "ms_per_gmres_it": tm["solve"] / its_total, "gmres_its_per_step": summ["gmres_its_total"] / steps,
           # deal.II counts from the second inner step of a cycle on (SolverFGMRES): a solve reporting k iterations ran k + 1
           # inner steps (preconditioner apply + operator apply + Gram-Schmidt) when it needs no restart
           "ms_per_inner_step": (tm["solve"] - tm["precond_setup"]) / (summ["gmres_its_total"] + summ["newton_steps"]),
           "e2e": {"value": n_dofs / (ms_e2e / steps * 1e-3), "unit": "DoF/s", "h2d_bytes_per_step": int(8 * 18 * T.n_owned_nodes),
                   "d2h_bytes_per_step": int(8 * 18 * T.n_owned_nodes), "ms_per_step": ms_e2e / steps,
                   "wall_ms_per_step": wall_e2e / steps * 1e3},
           "gpu_launches": tm["launches"], "context_create_s": create_s, "n_owned_dofs_rank0": int(18 * T.n_owned_nodes),
           "memory": {"device_bytes_rank0": info["device_bytes"], "nnzb": info["nnzb"], "fast_rows": info["n_fast_rows"],
                      "slow_cells": info["n_slow_cells"]}}
    if bj is not None:
        res["block_jacobi"] = bj
    if world > 1:  # per-iteration cost of the two exchange steps (collective calls: 20 back-to-back each)
        try:
            res["halo_ms_per_exchange"] = max_over_ranks(ctx.time_kernel(9, reps=3, flush_l2=False)) / 20.0
            res["allreduce_ms_per_dot"] = max_over_ranks(ctx.time_kernel(10, reps=3, flush_l2=False)) / 20.0
        except Exception as exc:
            res["exchange_timing_error"] = str(exc)
    if want_kernels:
        ctx.set_solution(x0)
        ctx.assemble()
        hbm_peak, peak_src = peaks()
        key = "q%d_%s_%dgpu" % (degree, ("r%d" % refine) if global_refine is None else ("g%d" % global_refine), world)
        roof, asm, kern = kernel_numbers(ctx, T, info, degree, hbm_peak, peak_src, key)
        res.update(roofline=roof, assembly=asm, kernels=kern)

    # ---- multi-GPU parity: rank 0 repeats the first Newton steps on ONE GPU with the same global mesh ----
    if want_parity and world > 1:
        par = None
        if rank == 0:
            try:
                if global_refine is not None:  # strong scaling: the same cube on one rank, same preconditioner
                    _, T1, c1, lv1, _ = build_context(vh, None, 0, 1, local_rank, degree, None, global_refine, precond, mg_coarsest)
                else:                          # weak scaling: the N stacked root cubes on one rank (block-Jacobi runs only)
                    if precond == "mg":
                        raise RuntimeError("parity run with the multigrid preconditioner needs a strong-scaling workload")
                    m1 = vh.Mesh(degree, [-20.0] * 3, [20.0, 20.0, -20.0 + 40.0 * world], base=(1, 1, world),
                                 face_bid=(1, 1, 1, 1, 4, 4), n_global_refine=refine).finalize(1)
                    T1 = m1.tables(0)
                    c1 = vh.Context(T1, device=local_rank)
                    c1.set_coef_vector(coef_vector())
                    lv1 = [c1]
                x1 = initial_state(T1)[:18 * T1.n_owned_nodes]
                n_par = min(3, len(runs[0]))
                ms1, runs1 = run_steps(c1, x1, n_par)
                a, b = runs[0][:n_par], runs1[0][:n_par]
                par = {"steps_compared": n_par,
                       "max_rel_residual_diff": float(max(abs(p[2] - q[2]) / abs(q[2]) for p, q in zip(a, b))),
                       "max_rel_rhs_norm_diff": float(max(abs(p[3] - q[3]) / abs(q[3]) for p, q in zip(a, b))),
                       "its_equal": bool(all(p[0] == q[0] and p[1] == q[1] for p, q in zip(a, b))),
                       "gmres_its_n_gpus": [int(p[0]) for p in a], "gmres_its_1_gpu": [int(q[0]) for q in b],
                       "single_gpu_ms_per_step_first_steps": ms1 / n_par,
                       "n_gpu_ms_first_steps_note": "the 1-GPU time covers only the first %d steps of a run (fewer linear "
                                                    "iterations than the average step)" % n_par}
                for c in lv1:
                    c.close()
            except Exception as exc:
                par = {"error": str(exc)}
        barrier()
        res["multi_gpu_parity"] = par
    res["preconditioner"] = ("multigrid V(%d,%d) cycle, %d levels (coarsest: refinement %d), Chebyshev/block-Jacobi smoothing"
                             % (MG_PARAMS.get("pre", 1), MG_PARAMS.get("post", 1), len(levels), mg_coarsest)) if precond == "mg" \
        else "nodal 18x18 block-Jacobi"
    for c in levels:
        c.close()
    return res


def run_ours(args):
    # NCCL / torch print banners on stdout; the contract is ONE JSON line there, so everything else goes to stderr.
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    import torch
    import verkko_hem_repo_b200 as vh

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node %d for --gpus %d" % (args.gpus, args.gpus))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    # clocks and throttle reasons are polled (nvidia-smi, ~5 Hz) over warm-up + timed + e2e steps of the headline workload
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    fallback = None
    try:
        main = measure(vh, torch, dist, rank, world, local_rank, args.degree, args.refine, args.global_refine, args.steps, args.warmup,
                       want_kernels=True, want_parity=not args.no_parity, precond=args.precond, mg_coarsest=args.mg_coarsest)
    except vh.VhError as exc:
        # the scalars that decide convergence are replicated, so every rank fails in the same call; block-Jacobi (the north
        # star's preconditioner) needs no hierarchy and is the documented fallback of the multigrid mode
        if args.precond != "mg":
            raise
        fallback = "multigrid run failed (%s); measured with block-Jacobi instead" % exc
        print("bench: " + fallback, file=sys.stderr)
        main = measure(vh, torch, dist, rank, world, local_rank, args.degree, args.refine, args.global_refine, args.steps, args.warmup,
                       want_kernels=True, want_parity=not args.no_parity, precond="bj", mg_coarsest=args.mg_coarsest)
    clocks = sampler.stop() if rank == 0 else None
    if clocks is not None:
        clocks["window"] = "warm-up + timed + e2e Newton steps of the headline workload"

    # configs[1] (C2) next to the headline on one GPU: the same measurement, kernels included
    c2 = None
    if world == 1 and args.workload == "c5" and not args.no_c2:
        try:
            c2 = measure(vh, torch, dist, rank, world, local_rank, 1, 5, None, args.steps, args.warmup, want_kernels=True,
                         want_parity=False)
            c2["workload"] = workload_string(1, 5, None, c2["n_dofs"], c2["n_cells"])
        except Exception as exc:
            c2 = {"error": str(exc)}

    cpu = port = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        n_side = 2 ** (args.refine if args.global_refine is None else args.global_refine)
        s = reference_sample(40.0 / n_side, 6)
        if s is not None:
            step_s = main["n_cells"] / s["cells_per_s"]
            cpu = {"value": main["n_dofs"] / step_s, "unit": "DoF/s", "cores": s["cores"], "kind": "reference",
                   "sample": "%d Q1 cells (Jacobian + 1 residual each) through the reference's verbatim term files and literal "
                             "(q,i,j) loops in %.1f s on %d cores; extrapolated linearly to %d cells; solve excluded (upper bound "
                             "on the reference's throughput)" % (s["cells"], s["wall_s"], s["cores"], main["n_cells"]),
                   "assembly_dofs_per_s": s["assembly_cells_per_s"] * main["n_dofs"] / main["n_cells"],
                   "residual_dofs_per_s": s["residual_cells_per_s"] * main["n_dofs"] / main["n_cells"]}
        if args.degree == 1:
            try:
                nst = max(main["newton"]["newton_steps"], 1)
                port = cpu_port_numbers(main["newton"]["gmres_its_total"] / nst, main["newton"]["line_search_trials_total"] / nst,
                                        main["n_cells"], main["n_dofs"])
            except Exception as exc:
                port = {"error": str(exc)}

    if rank == 0:
        strong = args.global_refine is not None
        out = {"metric": "femgl Newton-step throughput", "value": main["value"], "unit": "DoF/s", "n_gpus": world,
               "steps": args.steps, "warmup": args.warmup, "ms_per_step": main["ms_per_step"], "higher_is_better": True,
               "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
               "config": config_dict(args.degree, args.refine, args.global_refine, main["n_dofs"], main["n_cells"], world)}
        for k in ("preconditioner", "newton", "final_energy", "phase_ms_per_step", "gmres_its_per_step", "ms_per_gmres_it", "ms_per_inner_step", "block_jacobi",
                  "halo_ms_per_exchange",
                  "allreduce_ms_per_dot", "roofline", "assembly", "kernels", "e2e", "gpu_launches", "multi_gpu_parity", "memory",
                  "context_create_s"):
            if k in main:
                out[k] = main[k]
        if fallback is not None:
            out["preconditioner_fallback"] = fallback
        out["clocks"] = clocks
        out["cpu_baseline"] = cpu
        out["cpu_port"] = port
        if cpu is not None or port is not None:
            vs = {"note": "phase-by-phase GPU/CPU ratios (reported baseline, not a target); the reference's literal loops have no "
                          "solve leg, so a whole-step ratio is only formed against the port"}
            if cpu is not None:
                vs["assembly_vs_reference_literal"] = main["assembly"]["dofs_per_s"] / cpu["assembly_dofs_per_s"]
                vs["residual_vs_reference_literal"] = (main["n_owned_dofs_rank0"] / (main["kernels"]["residual_ms"] * 1e-3)) / cpu["residual_dofs_per_s"]
            if port is not None and "error" not in port:
                vs["assembly_vs_port"] = main["assembly"]["dofs_per_s"] / port["assembly_with_scatter_dofs_per_s"]
                vs["newton_step_vs_port"] = port["newton_step_ms"] / main["ms_per_step"]
            out["vs_cpu"] = vs
        if c2 is not None:
            out["c2"] = c2
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(out), flush=True)
        os.dup2(2, 1)
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c5", choices=sorted(WORKLOADS),
                    help="c5 (default): BASELINE configs[4], Q1 global refinement 7 split over the GPUs (strong scaling); "
                         "c2: configs[1], Q1 r5 per GPU (weak); c3: configs[2], Q2 r5 split over the GPUs")
    ap.add_argument("--refine", type=int, default=None, help="override: refinements per GPU (weak scaling)")
    ap.add_argument("--degree", type=int, default=None)
    ap.add_argument("--global-refine", type=int, default=None,
                    help="override: one cube with this many global refinements split over the GPUs (strong scaling)")
    ap.add_argument("--precond", default="auto", choices=["auto", "bj", "mg"],
                    help="GMRES preconditioner: bj = nodal block-Jacobi (north star), mg = geometric multigrid V-cycle over the "
                         "coarser global refinements of the same box (Q1, strong-scaling workloads; the line then also carries the "
                         "block-Jacobi step time of the same context under block_jacobi); auto = mg where a hierarchy exists "
                         "(Q1 with --global-refine above --mg-coarsest, i.e. the c5 default), bj elsewhere")
    ap.add_argument("--mg-coarsest", type=int, default=5, help="coarsest refinement level of the multigrid hierarchy")
    ap.add_argument("--mg-pre", type=int, default=None, help="Chebyshev degree of the pre-smoother (bench default 0; library default 1)")
    ap.add_argument("--mg-post", type=int, default=None, help="Chebyshev degree of the post-smoother (default 1)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-c2", action="store_true", help="skip the configs[1] measurement next to the headline (N = 1)")
    ap.add_argument("--no-parity", action="store_true", help="skip the single-GPU parity run of rank 0 (N > 1)")
    args = ap.parse_args()
    d, r, g = WORKLOADS[args.workload]
    if args.refine is not None or args.global_refine is not None:
        r, g = args.refine, args.global_refine
        args.workload = "custom"
    if args.degree is not None:
        d = args.degree
        if d != WORKLOADS.get(args.workload, (d,))[0]:
            args.workload = "custom"
    args.degree, args.refine, args.global_refine = d, r, g
    if args.mg_pre is not None:
        MG_PARAMS["pre"] = args.mg_pre
    if args.mg_post is not None:
        MG_PARAMS["post"] = args.mg_post
    if args.precond == "auto":
        args.precond = "mg" if (d == 1 and g is not None and g > args.mg_coarsest) else "bj"
    if args.refine is None and args.global_refine is None:
        raise SystemExit("--refine or --global-refine needed")
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
