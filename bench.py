#!/usr/bin/env python
"""bench.py — the femgl Newton step (assembly + GMRES(30)/block-Jacobi + line search) on N B200s of one box.

    python bench.py --gpus N --steps K --warmup W            # this framework (CUDA, through the C ABI)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's own CPU code (oracle/_ref) on the host cores

A "step" is one Newton step of run.cc:214-218 (assemble_system, solve(tol), newton_iteration) on the synthetic
cube of BASELINE.json configs[1]: Q1, hyper_cube(-20,20), global refinement 5, 646 866 DoFs per GPU
(weak scaling: N GPUs solve an N-times taller box, partitioned along the Morton curve).
The W warm-up steps and the K timed steps both start from the uniform B-phase initial condition
(setup_uniform_B-phase.cc:245-259), so the timed region is Newton iterations 1..K of a real run.

Prints ONE JSON line.  `value` = DoFs advanced one Newton step per second, whole job (ms_per_step is the
Newton-step wall time the BASELINE metric names); `roofline` is the SpMV kernel (HBM-bound) timed live with CUDA
events; `assembly` and `kernels` break the step down; `e2e` repeats the step through the C ABI with host
buffers in and out every step; `cpu_baseline` is the reference's literal term code on a bounded cell sample.
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
    if global_refine is not None:
        # strong scaling (BASELINE configs[4]): ONE cube refined `global_refine` times, Morton-partitioned over the ranks
        L = 20.0
        m = vh.Mesh(degree, [-L, -L, -L], [L, L, L], base=(1, 1, 1), face_bid=(1, 1, 1, 1, 4, 4), n_global_refine=global_refine)
        m.finalize(n_gpus)
        return m
    # weak scaling: N root cubes stacked along z, each refined `refine` times; Morton order keeps each root
    # contiguous, so rank r owns root r (the p4est partition of a 1 x 1 x N brick).
    L = 20.0  # "cube half side length" default (declare.cc:159)
    m = vh.Mesh(degree, [-L, -L, -L], [L, L, -L + 2 * L * n_gpus], base=(1, 1, n_gpus), face_bid=(1, 1, 1, 1, 4, 4),
                n_global_refine=refine)
    m.finalize(n_gpus)
    return m


def workload_string(degree, refine, global_refine, n_dofs, n_cells):
    """config.workload — identical in both arms (the driver compares the two lines)."""
    return ("femgl 3D cube Q%d, global refinement %s (%d DoFs total, %d cells), B-phase IC, "
            "z walls AdGR diffuse (bt=2), p=25 bar t=0.5 SCC on, linear tol 1e-1, damped Newton 0.83"
            % (degree, ("%d per GPU" % refine) if global_refine is None else ("%d, one cube split over the GPUs" % global_refine),
               n_dofs, n_cells))


def initial_state(T):
    from helpers import b_phase_state, MATEP_SCC_ON
    return b_phase_state(T, MATEP_SCC_ON, noise=0.0)


def newton_step(ctx):
    """run.cc:214-218 + iteration.cc:128-210 through the C ABI.  Returns (lin_its, n_trials, res_norm)."""
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
    return its, n_trials, cur


# ------------------------------------------------------------------------------------------
# reference arm: the reference's own term files (oracle/_ref) on the host cores
# ------------------------------------------------------------------------------------------
def _ref_worker(args):
    seed, n_cells_each, h, coef = args
    import femgl_oracle as O
    rng = np.random.default_rng(seed)
    amp = COEF["gapB"] * float(np.float32(0.577350269))
    t0 = time.time()
    for _ in range(n_cells_each):
        U = np.zeros((8, 18))
        U[:, [0, 4, 8]] = amp
        U += 0.05 * COEF["gapB"] * rng.uniform(-1, 1, U.shape)
        O.ref_cell(1, [0, 0, 0], h, U.ravel(), coef, [(4, 4)], want_matrix=True)   # assemble_system cell
        O.ref_cell(1, [0, 0, 0], h, U.ravel(), coef, [(4, 4)], want_matrix=False)  # one compute_residual cell
    return time.time() - t0


def reference_sample(refine, cells_per_core=1):
    """Time the reference's literal (q,i,j) loops + term functions on `cores * cells_per_core` Q1 cells, one worker
    process per host core (the reference is 1 thread per MPI rank, sol/src/main.cc:108)."""
    import femgl_oracle as O
    if not O.have_ref():
        return None
    cores = os.cpu_count() or 1
    n_side = 2 ** refine
    h = [40.0 / n_side] * 3
    coef = coef_vector()
    t0 = time.time()
    with mp.get_context("fork").Pool(cores) as pool:
        pool.map(_ref_worker, [(1000 + k, cells_per_core, h, coef) for k in range(cores)])
    wall = time.time() - t0
    n_cells = cores * cells_per_core
    return dict(cores=cores, cells=n_cells, wall_s=wall, cells_per_s=n_cells / wall)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    p = args.degree
    if args.global_refine is None:  # weak scaling: `gpus` root cubes stacked along z (build_mesh)
        n_side, n_z = 2 ** args.refine, 2 ** args.refine * args.gpus
    else:                           # strong scaling: one cube
        n_side = n_z = 2 ** args.global_refine
    n_nodes = (p * n_side + 1) ** 2 * (p * n_z + 1)
    n_cells = n_side * n_side * n_z
    n_dofs = 18 * n_nodes
    for _ in range(args.warmup if args.warmup < 2 else 1):
        reference_sample(args.refine, 1)
    t_list, last = [], None
    for _ in range(args.steps):
        last = reference_sample(args.refine, 2)
        if last is None:
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libvhref.so not built"}))
            return
        t_list.append(last["wall_s"])
    cps = last["cells"] * len(t_list) / sum(t_list)
    q2_note = ""
    if p == 2:
        # the sample is timed on Q1 cells (one Q2 cell takes ~100 s in the reference's loops); a Q2 cell visits
        # (486^2 * 27) / (144^2 * 8) = 38.4x as many (q,i,j) triples, each the same work (assemble.cc:188-252)
        cps /= (486.0 ** 2 * 27.0) / (144.0 ** 2 * 8.0)
        q2_note = "; Q2 rate = Q1 rate / 38.4 (ratio of (q,i,j) triples per cell)"
    # one Newton step of the reference = one assembly + >= 1 residual evaluation over all cells; its ML-AMG solve is
    # not reproducible here and is left out, so this is an UPPER bound on the reference's throughput.
    step_s = n_cells / cps
    val = n_dofs / step_s
    out = {"impl": "reference", "metric": "femgl Newton-step throughput",
           "value": val, "unit": "DoF/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": step_s * 1e3, "higher_is_better": True, "scaling": "weak" if args.global_refine is None else "strong",
           "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": workload_string(args.degree, args.refine, args.global_refine, n_dofs, n_cells),
                      "reference_arm": "the reference's own CPU code for this path (oracle/_ref: its verbatim cell_mat_vec term files "
                                       "driven by its literal (q,i,j) loops) on all host cores; %d-cell sample per step, extrapolated "
                                       "linearly to the workload; its ML-AMG solve is left out (upper bound on its throughput)"
                                       % last["cells"]},
           "cpu_baseline": {"value": val, "unit": "DoF/s", "cores": last["cores"], "kind": "reference",
                            "sample": "%d Q1 cells/step (Jacobian + 1 residual) by the reference's verbatim cell_mat_vec "
                                      "term files driven by its literal (q,i,j) loops; %.2f cells/s on %d cores, "
                                      "extrapolated linearly to %d cells%s" % (last["cells"], cps, last["cores"], n_cells, q2_note)},
           "e2e": {"value": val, "unit": "DoF/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out))


# ------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------
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

    mesh = build_mesh(world, args.refine, args.degree, args.global_refine)
    T = mesh.tables(rank)
    ctx = vh.Context(T, device=local_rank)
    if world > 1:
        uid = [vh.Context.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        ctx.comm_init(rank, world, uid[0])
    ctx.set_coef_vector(coef_vector())
    if args.spmv_mf:  # whole run with the matrix-free operator apply inside GMRES (opt-in this round)
        ctx.set_spmv_matrix_free(True)
    # the mode this run really uses (the library's default may come from VH_SPMV_MF): 0 packed SpMV, 1 matrix-free, 2 table-free
    mf_mode = int(ctx.info().get("spmv_matrix_free", 0))
    use_mf = mf_mode != 0
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

    # clocks and throttle reasons are polled (nvidia-smi, ~5 Hz) from the first warm-up step to the end of the e2e loop: the
    # same workload runs throughout, and the device-timed K steps alone (tens of milliseconds) are shorter than one poll
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()

    # ---- warm-up ----
    ctx.set_solution(x0)
    for _ in range(args.warmup):
        newton_step(ctx)

    # ---- timed: K Newton steps from the initial condition, state resident in HBM ----
    ctx.set_solution(x0)
    ctx.timers(reset=True)
    hist = []
    barrier()
    ctx.timer_start()
    for _ in range(args.steps):
        hist.append(newton_step(ctx))
    ms = ctx.timer_stop()
    barrier()
    tm = ctx.timers()
    # GL free energy of the state the timed steps ended in (outside the timed region; the reference never evaluates the
    # functional, SURVEY.md section 0.6 — vh_energy integrates the functional whose half-gradient is the residual)
    try:
        final_energy = ctx.energy(0)
    except Exception:
        final_energy = None
    ms = max_over_ranks(ms)
    ms_per_step = ms / args.steps
    value = n_dofs / (ms_per_step * 1e-3)

    # ---- e2e: same steps, host buffers in and out of the C ABI every step (pinned host memory) ----
    xh = torch.from_numpy(x0.copy()).pin_memory().numpy()
    barrier()
    t0 = time.perf_counter()
    ctx.timer_start()
    for _ in range(args.steps):
        ctx.set_solution(xh)          # H2D of the step's input
        newton_step(ctx)
        ctx.get_solution(out=xh)      # D2H of the step's result, straight into the pinned buffer
    ms_e2e = max_over_ranks(ctx.timer_stop())
    wall_e2e = time.perf_counter() - t0
    barrier()
    e2e_val = n_dofs / (ms_e2e / args.steps * 1e-3)
    clocks = sampler.stop() if rank == 0 else None
    if clocks is not None:
        clocks["window"] = "warm-up + timed + e2e Newton steps (same workload)"

    # ---- live kernel timings for the roofline (rank 0's kernels; inputs larger than L2 or L2 flushed) ----
    ctx.set_solution(x0)
    ctx.assemble()
    nnzb, nb = info["nnzb"], T.n_owned_nodes
    # algorithmic bytes of one SpMV by SURVEY.md §8(d) (BSR18: every 18x18 block stored in full) ...
    spmv_bytes = 8 * 324 * nnzb + 4 * nnzb + 4 * (nb + 1) + 16 * 18 * nb
    # ... and the bytes this implementation really has to move: lattice rows are stored as packed symmetric blocks
    # (180 doubles instead of 324), general-scatter rows in full
    n_fast_blocks = info["n_packed_blocks"]
    packed = n_fast_blocks > 0
    spmv_moved = 8 * (180 * n_fast_blocks + 324 * (nnzb - n_fast_blocks)) + 4 * nnzb + 4 * (nb + 1) + 16 * 18 * nb
    spmv_moved_packed = spmv_moved
    t_spmv = ctx.time_kernel(0, reps=20, flush_l2=True)
    t_asm = ctx.time_kernel(1, reps=5, flush_l2=True)
    t_pw = ctx.time_kernel(5, reps=5, flush_l2=True)
    t_rows = ctx.time_kernel(6, reps=5, flush_l2=True)
    t_res = ctx.time_kernel(2, reps=5, flush_l2=True)
    t_bj = ctx.time_kernel(3, reps=10, flush_l2=True)
    fp64_peak = ctx.measure_fp64_peak()
    hbm_peak, peak_src = peaks()
    n = 8 if args.degree == 1 else 27
    asm_flops = 2.0 * n * n * n * 336 * T.n_cells
    asm_bytes = 8 * 324 * nnzb + 8 * 18 * n * T.n_cells  # SURVEY's write-once figure (full blocks)
    asm_moved = 8 * (180 * n_fast_blocks + 324 * (nnzb - n_fast_blocks)) + 8 * 18 * n * T.n_cells
    spmv_gbs = spmv_bytes / (t_spmv * 1e-3) / 1e9
    traffic = None
    try:  # DRAM bytes per launch measured once with `ncu --set full` for this exact workload (profiles/traffic.json)
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            tj = json.load(f)
        key = "q%d_r%d_%dgpu" % (args.degree, args.refine, world)
        if args.global_refine is None and key in tj:
            kname = "k_spmv_sym18" if packed else "k_spmv_bsr18"
            traffic = tj[key][kname]["read_bytes"] + tj[key][kname]["write_bytes"]
    except Exception:
        traffic = None
    # matrix-free apply: the H_q tables (8*180*n_q bytes per cell) + the per-cell products written and gathered once
    mf_moved = 8 * 180 * n * T.n_cells + 2 * 8 * 18 * n * T.n_cells + 8 * 324 * (nnzb - n_fast_blocks) + 16 * 18 * nb
    if use_mf:
        spmv_moved, traffic = mf_moved, None
    moved_gbs = spmv_moved / (t_spmv * 1e-3) / 1e9
    kernel_name = ("k_points<APPLY>+k_gather_apply" if use_mf else "k_spmv_sym18") if n_fast_blocks else "k_spmv_bsr18"
    roof = {"bound": "hbm", "kernel": kernel_name, "achieved": spmv_gbs, "peak": hbm_peak,
            "unit": "GB/s", "frac": spmv_gbs / hbm_peak, "traffic": traffic, "peak_source": peak_src, "ms_per_launch": t_spmv,
            "algorithmic_bytes_per_launch": spmv_bytes, "moved_bytes_per_launch": spmv_moved, "moved_gbs": moved_gbs,
            "hbm_utilization": moved_gbs / hbm_peak,
            "note": "achieved/frac use SURVEY's algorithmic bytes (full 18x18 blocks); the packed symmetric storage moves "
                    "moved_bytes_per_launch (0.56x), so frac can exceed 1 while hbm_utilization (= real DRAM bytes / time / peak) "
                    "stays below 1"}
    asm = {"ms": t_asm, "dofs_per_s": 18 * nb / (t_asm * 1e-3), "tflops_fp64": asm_flops / (t_asm * 1e-3) / 1e12,
           "fp64_peak_tflops_measured": fp64_peak, "frac_fp64": asm_flops / (t_asm * 1e-3) / 1e12 / fp64_peak if fp64_peak else None,
           "store_gbs": asm_bytes / (t_asm * 1e-3) / 1e9, "frac_hbm": asm_bytes / (t_asm * 1e-3) / 1e9 / hbm_peak,
           "pointwise_ms": t_pw, "rows_ms": t_rows, "algorithmic_flops": asm_flops, "algorithmic_bytes": asm_bytes,
           "moved_bytes": asm_moved,
           # flops the row-owner kernels really execute: packed symmetric entries (180 of 324) and, at Q2, the
           # sum-factorised contraction (3 x 81 FMAs per entry and row node instead of 27 x 27)
           "executed_flops": 2.0 * 180 * T.n_cells * (8 * 64 if args.degree == 1 else 27 * 243)}

    # the other operator-apply mode, kernel-only (the timed Newton steps above used the mode of this run)
    other = None
    if packed:
        try:
            ctx.set_spmv_matrix_free(0 if use_mf else 1)
            t_other = ctx.time_kernel(0, reps=20, flush_l2=True)
            om = spmv_moved_packed if use_mf else mf_moved
            other = {"mode": "packed-spmv" if use_mf else "matrix-free", "ms": t_other, "moved_bytes": om,
                     "moved_gbs": om / (t_other * 1e-3) / 1e9, "hbm_utilization": om / (t_other * 1e-3) / 1e9 / hbm_peak}
        except Exception as exc:  # never let the side measurement take the bench line down
            other = {"error": str(exc)}
        finally:
            try:
                ctx.set_spmv_matrix_free(mf_mode)
            except Exception:
                pass

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        s = reference_sample(args.refine, 6)
        if s is not None:
            n_cells = mesh.n_cells
            step_s = n_cells / s["cells_per_s"]
            cpu = {"value": n_dofs / step_s, "unit": "DoF/s", "cores": s["cores"], "kind": "reference",
                   "sample": "%d Q1 cells (Jacobian + 1 residual each) through the reference's verbatim term files and literal "
                             "(q,i,j) loops in %.1f s on %d cores; extrapolated linearly to %d cells; solve excluded"
                             % (s["cells"], s["wall_s"], s["cores"], n_cells)}

    # the optimised CPU restatement (oracle O2, C, one pthread per host core) on the same cells: the "fair CPU" assembly
    # rate SURVEY.md section 8(d) asks for next to the reference's literal loops (reported baseline, not a target)
    if cpu is not None and args.degree == 1:
        try:
            import femgl_oracle as O
            nsmp = min(T.n_cells, 2048)
            xs = np.zeros(18 * T.n_local_nodes)
            xs[:x0.size] = x0
            fptr, fno, fbid = T.face_csr()
            dummy = np.zeros(1, np.int32)
            t0 = time.perf_counter()
            O.cells(1, T.cell_nodes[:nsmp], T.cell_origin[:nsmp], T.cell_h[:nsmp], xs, coef_vector(), fptr[:nsmp + 1],
                    fno if fno.size else dummy, fbid if fbid.size else dummy, want_matrix=True)
            dt = time.perf_counter() - t0
            asm["cpu_port"] = {"dofs_per_s": 18 * nb / (T.n_cells / (nsmp / dt)), "unit": "DoF/s", "cores": os.cpu_count(), "kind": "port",
                               "sample": "%d Q1 cell matrices + rhs by oracle/femgl_oracle.c (closed-form H_q, pthreads) in %.2f s; "
                                         "scatter excluded; extrapolated linearly to %d cells" % (nsmp, dt, T.n_cells)}
        except Exception as exc:
            asm["cpu_port"] = {"error": str(exc)}

    if rank == 0:
        out = {"metric": "femgl Newton-step throughput",
               "value": value, "unit": "DoF/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
               "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak" if args.global_refine is None else "strong", "vs_baseline": None, "dtype": "f64",
               "data": "synthetic",
               "config": {"workload": workload_string(args.degree, args.refine, args.global_refine, n_dofs, mesh.n_cells),
                          "l2": "matrix (%.2f GB/GPU) exceeds the 126 MB L2; kernel timings flush L2 between launches"
                                % (8 * 324 * nnzb / 1e9),
                          "parallelism": "subdomain x%d (Morton partition, NCCL halo + all-reduce)" % world},
               "newton_history": [{"gmres_its": h[0], "line_search_trials": h[1], "residual": h[2]} for h in hist],
               "final_energy": final_energy,
               "phase_ms_per_step": {k: tm[k] / args.steps for k in ("assemble", "residual", "solve", "vector")},
               "roofline": roof, "assembly": asm,
               # the other HBM-bound kernels with SURVEY.md section 8(d)'s algorithmic bytes: residual assembly 8*dpc*cells (gather)
               # + 8*N (write); block-Jacobi apply 2592*nb + 16*N
               "kernels": {"spmv_ms": t_spmv, "spmv_gbs": spmv_gbs, "residual_ms": t_res,
                           "residual_gbs": (8 * 18 * n * T.n_cells + 8 * 18 * nb) / (t_res * 1e-3) / 1e9,
                           "block_jacobi_apply_ms": t_bj, "block_jacobi_apply_gbs": (2592 * nb + 16 * 18 * nb) / (t_bj * 1e-3) / 1e9,

                           "operator_apply_mode": ("packed-spmv", "matrix-free", "table-free", "matrix-free-v2")[mf_mode], "other_apply_mode": other},
               "e2e": {"value": e2e_val, "unit": "DoF/s", "h2d_bytes_per_step": int(8 * 18 * T.n_owned_nodes),
                       "d2h_bytes_per_step": int(8 * 18 * T.n_owned_nodes), "ms_per_step": ms_e2e / args.steps,
                       "wall_ms_per_step": wall_e2e / args.steps * 1e3},
               "gpu_launches": tm["launches"], "clocks": clocks, "cpu_baseline": cpu,
               "matrix": {"nnzb": nnzb, "fast_rows": info["n_fast_rows"], "slow_cells": info["n_slow_cells"],
                          "device_bytes": info["device_bytes"]}}
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(out), flush=True)
        os.dup2(2, 1)
    ctx.close()
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--refine", type=int, default=5)
    ap.add_argument("--degree", type=int, default=1)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--spmv-mf", action="store_true",
                    help="run GMRES with the matrix-free operator apply (vh_set_spmv_matrix_free) instead of the packed SpMV")
    ap.add_argument("--global-refine", type=int, default=None,
                    help="strong scaling: one cube with this many global refinements split over the GPUs (7 = BASELINE C5, 38.6M DoFs)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
