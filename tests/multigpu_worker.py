"""torchrun worker for tests/test_gpu_driver.py::test_two_gpu_run_equals_one_gpu_run: every rank runs two Newton steps
on its subdomain; rank 0 also runs the same problem alone on its GPU and compares histories and the solution."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import verkko_hem_repo_b200 as vh  # noqa: E402
from helpers import b_phase_state, coef_vector  # noqa: E402


def newton(ctx, n):
    hist = []
    for _ in range(n):
        bn = ctx.assemble()
        its, _ = ctx.solve(1e-1)
        k = 0
        for i in range(100):
            ctx.line_search_trial(0.83 ** i)
            cur = ctx.residual()
            k += 1
            if cur < bn:
                break
        ctx.accept_trial()
        hist.append((bn, its, k, cur, ctx.energy(0)))
    return hist


def main():
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    coef = coef_vector(bt=2.0)
    periodic = len(sys.argv) > 1 and sys.argv[1] == "periodic"  # the reference's active grid: x/y-periodic slab

    hanging = len(sys.argv) > 1 and sys.argv[1] == "hanging"    # C4-shaped: locally refined slab, several levels
    mg = len(sys.argv) > 1 and sys.argv[1] == "mg"              # multigrid V-cycle preconditioner, 3 levels (r4, r3, r2)

    def make(n_ranks):
        if periodic:
            return vh.periodic_slab(1, 3, half=(2.0, 2.0, 1.0), n_ranks=n_ranks)
        if hanging:
            m = vh.Mesh(1, [-3, -2, -4], [3, 2, 4], n_global_refine=2)
            for _ in range(2):  # two "adaptive cycles": the 30 % of the cells nearest the plane z = 0.4
                d = np.abs(m.cell_centers()[:, 2] - 0.4)
                m.refine(d <= np.sort(d)[int(0.3 * m.n_cells)])
            return m.finalize(n_ranks)
        return vh.unit_cube(1, 4 if mg else 3, half=2.0, n_ranks=n_ranks)

    def hierarchy(n_ranks, r, fine_mesh, fine_tables, fine_ctx, collective):
        """coarser levels of the multigrid hierarchy on the same partition, attached to fine_ctx"""
        keep = []
        mf, Tf, cf = fine_mesh, fine_tables, fine_ctx
        for lv in (3, 2):
            mc = vh.unit_cube(1, lv, half=2.0, n_ranks=n_ranks)
            Tc = mc.tables(r)
            cc = vh.Context(Tc, device=lr)
            if collective:   # the levels of a hierarchy share the fine level's communicator (vh_comm_share)
                cc.comm_share(fine_ctx)
            cc.set_coef_vector(coef)
            cf.mg_attach(cc, *vh.mg_prolongation(mf, Tf, mc, Tc))
            keep.append(cc)
            mf, Tf, cf = mc, Tc, cc
        fine_ctx.set_preconditioner("multigrid")
        return keep

    mesh = make(world)
    T = mesh.tables(rank)
    ctx = vh.Context(T, device=lr)
    uid = [vh.Context.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    ctx.comm_init(rank, world, uid[0])
    ctx.set_coef_vector(coef)
    levels = hierarchy(world, rank, mesh, T, ctx, True) if mg else []
    x = b_phase_state(T, seed=9)
    ctx.set_solution(x[:18 * T.n_owned_nodes])
    hist = newton(ctx, 2)
    sol = ctx.get_solution()
    gathered = [None] * world
    dist.all_gather_object(gathered, (T.node_xyz[:T.n_owned_nodes].copy(), sol))
    ok = True
    if rank == 0:
        m1 = make(1)
        T1 = m1.tables(0)
        c1 = vh.Context(T1, device=lr)
        c1.set_coef_vector(coef)
        lv1 = hierarchy(1, 0, m1, T1, c1, False) if mg else []
        c1.set_solution(b_phase_state(T1, seed=9))
        h1 = newton(c1, 2)
        s1 = c1.get_solution().reshape(-1, 18)
        key = {tuple(np.round(p, 9)): i for i, p in enumerate(T1.node_xyz)}
        for a, b in zip(hist, h1):
            ok &= abs(a[0] - b[0]) <= 1e-10 * b[0] and a[1] == b[1] and a[2] == b[2] and abs(a[3] - b[3]) <= 1e-10 * b[3]
            ok &= abs(a[4] - b[4]) <= 1e-10 * abs(b[4])
        for xyz, s in gathered:
            idx = np.array([key[tuple(np.round(p, 9))] for p in xyz])
            ok &= np.abs(s.reshape(-1, 18) - s1[idx]).max() <= 1e-9 * np.abs(s1).max()
        print("hist multi", hist)
        print("hist single", h1)
        print("MULTIGPU PARITY OK" if ok else "MULTIGPU PARITY FAILED")
    ctx.close()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
