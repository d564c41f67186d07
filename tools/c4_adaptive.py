"""BASELINE configs[3] (C4) on N GPUs: the slab with a flat A/B interface (BnA.h:130-163), AdGR z walls, adaptive refinement
cycles with hanging nodes, under the reference's control flow (run.cc:182-256): per cycle setup_system -> Newton steps until
the stuck/converged rule fires -> refine_grid (refine.cc:133-179: indicator, fixed-number selection, SolutionTransfer,
constraints_solution.distribute).

The mesh side (indicator, flags, refinement, repartitioning) is the host's job in the reference too (deal.II + p4est) and
runs replicated on every rank here; the Newton steps go through the C ABI with one context per rank and cycle.  Because the
Morton partition moves when cells are refined, the state is transferred the way deal.II's parallel SolutionTransfer does
after repartitioning: gather the owned values, interpolate on the host mesh, hand every rank its new owned part.

    python -m torch.distributed.run --nproc-per-node N ... tools/c4_adaptive.py [--cycles 4] [--check-single] [--json out.json]
    python tools/c4_adaptive.py                      (one GPU)
    ... --dry                                        (no GPU: the Newton step is replaced by a fixed smooth change of the state;
                                                      exercises the partition / gather / transfer logic under gloo on CPU)

--check-single: rank 0 repeats the whole run on ONE GPU and the summary carries the P-independence verdict (identical
refinement flags, linear-iteration and line-search counts; residual norms within 1e-10 relative, or within 1e-10 of the run's
first right-hand side for the nearly converged steps; final state within 1e-8)."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import verkko_hem_repo_b200 as vh  # noqa: E402
from helpers import coef_vector, distribute_constraints  # noqa: E402

LIN_TOL, LS_STEP, CONVERGE_ACC = 1e-1, 0.83, 5e-6


def bna_state(xyz, mat, ratio, half_z):
    """BnA.h:135-162: A phase below the plane z = ratio * Lz, B phase above (values as the reference's float constants)."""
    x = np.zeros((xyz.shape[0], 18))
    is_b = xyz[:, 2] >= ratio * half_z
    e_a = mat["gapA"] * float(np.float32(0.7071067811865475))
    e_b = mat["gapB"] * float(np.float32(0.5773502691896258))
    x[is_b, 0] = x[is_b, 4] = x[is_b, 8] = e_b
    x[~is_b, 0] = x[~is_b, 10] = e_a
    return x.ravel()


def refinement_flags(eta, ratio):
    """refine_and_coarsen_fixed_number(refine_ratio, 0): exactly floor(ratio * n) cells, largest indicators first.  The indicator
    is compared after rounding to 9 digits of the largest one and ties go to the lower cell index, so that partitions whose
    states differ in the last bits select the same cells."""
    n = eta.size
    q = np.round(eta / max(float(eta.max()), 1e-300), 9)
    order = np.lexsort((np.arange(n), -q))
    flags = np.zeros(n, dtype=np.uint8)
    flags[order[: int(np.floor(ratio * n))]] = 1
    return flags


def dry_step(x_owned, node_global):
    """stand-in for a Newton step without a GPU: a fixed smooth change keyed on the GLOBAL node id"""
    g = node_global.astype(np.float64)[:, None] * 18.0 + np.arange(18)[None, :]
    return (x_owned.reshape(-1, 18) * (1.0 + 1e-3 * np.sin(0.37 * g))).ravel()


def run(args, rank, world, device, dist, flags_in=None):
    """One complete adaptive run on `world` ranks.  Returns (summary dict, list of flag arrays, final global state)."""
    mat = vh.matep(25.0, 0.5, True)
    coef = coef_vector(bt=2.0)
    hx, hy, hz = args.half
    mesh = vh.Mesh(1, [-hx, -hy, -hz], [hx, hy, hz], face_bid=(1, 1, 1, 1, 4, 4), n_global_refine=args.initial_refine).finalize(world)
    x_global = bna_state(mesh.node_xyz(), mat, args.ratio, hz)
    cycles, all_flags, history = [], [], []
    keeper = None
    for cycle in range(args.cycles + 1):
        r_last = 0.0                                               # run.cc:206: residual_last_iter of this cycle
        t0 = time.perf_counter()
        T = mesh.tables(rank)
        t_tables = time.perf_counter() - t0
        own = T.node_global[: T.n_owned_nodes]
        # local_solution on this partition: owned + ghost values of the global state, constraints_solution.distribute
        x_local = distribute_constraints(T, x_global.reshape(-1, 18)[T.node_global].ravel())
        ctx = None
        t_create = 0.0
        if not args.dry:
            t0 = time.perf_counter()
            ctx = vh.Context(T, device=device)
            if world > 1:
                if keeper is None or args.new_comm_per_cycle:
                    uid = [vh.Context.nccl_unique_id() if rank == 0 else None]
                    dist.broadcast_object_list(uid, src=0)
                    ctx.comm_init(rank, world, uid[0])
                else:   # setup_system() of a later cycle: same ranks, same communicator (vh_comm_share)
                    ctx.comm_share(keeper)
            ctx.set_coef_vector(coef)
            if args.cheb_degree > 0:   # polynomial preconditioner: Chebyshev(degree) of block-Jacobi (no mesh hierarchy needed)
                ctx.set_preconditioner("chebyshev", coarse_degree=args.cheb_degree, coarse_range=args.cheb_range)
            ctx.set_solution(x_local[: 18 * T.n_owned_nodes])
            t_create = time.perf_counter() - t0
        x_owned = x_local[: 18 * T.n_owned_nodes].copy()
        n_steps, ms_steps, res = 0, 0.0, 0.0
        for it in range(args.max_newton + 1):                      # run.cc:207: iteration_loop <= n_iteration
            if args.dry:
                x_owned = dry_step(x_owned, own)
                bn, its, trials, res = 0.0, 0, 1, 1.0 / (1 + cycle + it)
            else:
                ctx.timer_start()
                bn = ctx.assemble()
                its, _ = ctx.solve(args.lin_tol, args.max_lin_it, args.restart)
                trials = 0
                for i in range(100):
                    ctx.line_search_trial(LS_STEP ** i)
                    res = ctx.residual()
                    trials += 1
                    if res < bn:
                        break
                ctx.accept_trial()
                ms_steps += ctx.timer_stop()
            n_steps += 1
            history.append(dict(cycle=cycle, iteration=it, rhs_norm=bn, linear_its=its, trials=trials, residual=res))
            if rank == 0:
                print("[%d rank(s)] cycle %d (%d DoFs) iteration %d: rhs %.6e, %d linear its, %d trials, residual %.6e"
                      % (world, cycle, 18 * mesh.n_nodes, it, bn, its, trials, res), file=sys.stderr, flush=True)
            if abs(res - r_last) < args.threshold and res > CONVERGE_ACC and cycle < args.cycles:   # run.cc:235-240
                break
            if res <= CONVERGE_ACC:
                break
            r_last = res
        info = ctx.info() if ctx is not None else {}
        if ctx is not None:
            x_owned = ctx.get_solution()
            if world > 1 and keeper is None and not args.new_comm_per_cycle:
                keeper = ctx        # cycle 0's (small) context stays alive as the owner of the communicator
            else:
                ctx.close()
        # gather the owned parts into the global state (every rank keeps a copy: the mesh side runs replicated)
        t0 = time.perf_counter()
        if world > 1:
            parts = [None] * world
            dist.all_gather_object(parts, (own, x_owned))
        else:
            parts = [(own, x_owned)]
        x_global = np.zeros(18 * mesh.n_nodes)
        for g, v in parts:
            x_global.reshape(-1, 18)[g] = v.reshape(-1, 18)
        t_gather = time.perf_counter() - t0
        rec = dict(cycle=cycle, n_cells=int(mesh.n_cells), n_dofs=int(18 * mesh.n_nodes), n_hanging_nodes=int(mesh.n_hanging_nodes),
                   newton_steps=n_steps, ms_per_newton_step=(ms_steps / n_steps if n_steps else None), residual=res,
                   t_tables_s=t_tables, t_context_s=t_create, t_gather_s=t_gather,
                   fast_rows=info.get("n_fast_rows"), slow_cells=info.get("n_slow_cells"), owned_nodes_this_rank=int(T.n_owned_nodes))
        if res <= CONVERGE_ACC or cycle == args.cycles:
            cycles.append(rec)
            break
        # refine_grid (refine.cc:133-179) on the replicated host mesh
        t0 = time.perf_counter()
        flags = refinement_flags(mesh.kelly_indicator(x_global), args.refine_ratio) if flags_in is None else flags_in[cycle]
        all_flags.append(flags)
        new = mesh.clone()
        new.refine(flags)
        new.finalize(world)
        x_global = new.interpolate_from(mesh, x_global)           # SolutionTransfer::interpolate (distribute: next cycle's setup)
        rec["t_refine_transfer_s"] = time.perf_counter() - t0
        cycles.append(rec)
        mesh = new
    if keeper is not None:
        keeper.close()
    return dict(world=world, cycles=cycles, history=history), all_flags, x_global


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cycles", type=int, default=4)
    # Default configuration (measured, profiles/r02g_c4_probes.txt): slab 6 x 4 x 8 refined 4 times (h = 0.375 x 0.25 x 0.5: the
    # coherence length, 1 in these units, is resolved), the reference's default cycle threshold 1.0 (declare.cc:203-240) and
    # GMRES(100).  What does NOT work, in the oracle as on the GPU: a 20 x 20 x 40 slab refined 3 times (h = 2.5 .. 5) - the
    # discontinuous BnA state sits near a saddle of the under-resolved functional and the line search stalls; and GMRES(30) +
    # block-Jacobi on the final mesh (cells of 5 sizes, 7.5 M DoFs): the Jacobian is indefinite around the A/B interface and
    # the restarted iteration stagnates (10 000 iterations without reaching 1e-1), while GMRES(100) needs 400-2 000 per step.
    ap.add_argument("--initial-refine", type=int, default=4)
    ap.add_argument("--half", type=float, nargs=3, default=[3.0, 2.0, 4.0])
    ap.add_argument("--ratio", type=float, default=0.1, help="A-phase block range ratio")
    ap.add_argument("--refine-ratio", type=float, default=0.3)
    ap.add_argument("--threshold", type=float, default=1.0, help="Cycle x refinement threshold (all cycles; reference default 1.0)")
    ap.add_argument("--max-newton", type=int, default=10, help="Number of interations")
    ap.add_argument("--restart", type=int, default=100, help="GMRES restart length (.prm key of the mirror: GMRES restart length)")
    ap.add_argument("--lin-tol", type=float, default=LIN_TOL, help="Cycle x linear solver tol (all cycles)")
    ap.add_argument("--max-lin-it", type=int, default=10000, help="maximum linear iteration number")
    ap.add_argument("--cheb-degree", type=int, default=0, help="> 0: Chebyshev polynomial of this degree around block-Jacobi as the preconditioner")
    ap.add_argument("--cheb-range", type=float, default=30.0)
    ap.add_argument("--new-comm-per-cycle", action="store_true", help="ncclCommInitRank for every cycle's context instead of vh_comm_share")
    ap.add_argument("--dry", action="store_true")
    ap.add_argument("--check-single", action="store_true")
    ap.add_argument("--json", default=None)
    args = ap.parse_args()
    rank, world, lr = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        if args.dry:
            dist.init_process_group("gloo")
        else:
            torch.cuda.set_device(lr)
            dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    out, flags, xg = run(args, rank, world, lr, dist)
    ok = True
    if rank == 0:
        if args.check_single and world > 1:
            one, flags1, xg1 = run(args, 0, 1, lr, None)
            same_flags = len(flags) == len(flags1) and all(np.array_equal(a, b) for a, b in zip(flags, flags1))
            ha, hb = out["history"], one["history"]
            same_counts = len(ha) == len(hb) and all(a["linear_its"] == b["linear_its"] and a["trials"] == b["trials"] for a, b in zip(ha, hb))
            rel = max((abs(a["residual"] - b["residual"]) / abs(b["residual"]) for a, b in zip(ha, hb)), default=0.0) if len(ha) == len(hb) else None
            # near convergence the residual norm is a difference of O(1) terms: its rounding noise scales with the first
            # right-hand side of the run, not with the (1e-6 times smaller) residual itself
            r0 = (hb[0]["rhs_norm"] if hb else 0.0) or 1.0
            rel0 = max((abs(a["residual"] - b["residual"]) / r0 for a, b in zip(ha, hb)), default=0.0) if len(ha) == len(hb) else None
            sol = float(np.abs(xg - xg1).max() / np.abs(xg1).max()) if xg.shape == xg1.shape else None
            # measured (profiles/r02g_c4_adaptive_8gpu.json): after ~8 000 GMRES(100) iterations on the final 7.5 M-DoF mesh the
            # 8-rank and the 1-rank runs still take identical iteration counts; their residual norms differ by 8e-12 of the
            # first right-hand side and the final states by 3e-9 (different summation orders of the inner products)
            ok = bool(same_flags and same_counts and rel is not None and (rel <= 1e-10 or rel0 <= 1e-10) and sol is not None and sol <= 1e-8)
            out["p_independence"] = dict(same_refinement_flags=bool(same_flags), same_iteration_counts=bool(same_counts),
                                         max_rel_residual_diff=rel, max_residual_diff_over_initial_rhs=rel0, max_rel_solution_diff=sol, ok=ok,
                                         single_gpu_ms_per_newton_step=[c["ms_per_newton_step"] for c in one["cycles"]])
        line = json.dumps(out)
        if args.json:
            with open(args.json, "w") as f:
                f.write(line + "\n")
        for c in out["cycles"]:
            print("cycle %d: %d cells, %d DoFs, %d hanging nodes, %d Newton steps, %s ms/step, residual %.3e"
                  % (c["cycle"], c["n_cells"], c["n_dofs"], c["n_hanging_nodes"], c["newton_steps"],
                     ("%.2f" % c["ms_per_newton_step"]) if c["ms_per_newton_step"] else "-", c["residual"]))
        if "p_independence" in out:
            print("P-INDEPENDENCE", "OK" if ok else "FAILED", json.dumps(out["p_independence"]))
        print("C4 ADAPTIVE DONE")
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
