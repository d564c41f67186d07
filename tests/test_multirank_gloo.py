"""world_size-2 test of the N>1 host logic on CPU (gloo): per-rank tables + halo plan + all-reduced inner products.

Each rank assembles its owned rows with the oracle from its own (owned + ghost-layer) cells, exchanges ghost values
with torch.distributed exactly as the halo plan prescribes (what vh_halo.cu does with NCCL send/recv), and runs the
oracle's GMRES with all-reduced dots.  The result must equal the 1-rank run: same iteration count, same update."""
import os
import socket

import numpy as np
import pytest

import femgl_oracle as O
import verkko_hem_repo_b200 as vh
from helpers import b_phase_state, coef_vector


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, refine, out_dir):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        coef = coef_vector(bt=2.0)
        m = vh.unit_cube(1, refine, half=2.0, n_ranks=world)
        T = m.tables(rank)
        NO = 18 * T.n_owned_nodes

        def halo(x_owned):
            x = np.zeros(18 * T.n_local_nodes)
            x[:NO] = x_owned
            xv = x.reshape(-1, 18)
            reqs, bufs = [], []
            for k, p in enumerate(T.peer_rank):
                send = torch.from_numpy(np.ascontiguousarray(xv[T.send_nodes[T.send_ptr[k]:T.send_ptr[k + 1]]]))
                recv = torch.zeros(int(T.recv_ptr[k + 1] - T.recv_ptr[k]), 18, dtype=torch.float64)
                if send.numel():
                    reqs.append(dist.isend(send, int(p)))
                if recv.numel():
                    reqs.append(dist.irecv(recv, int(p)))
                bufs.append((k, recv))
            for r in reqs:
                r.wait()
            for k, recv in bufs:
                xv[T.recv_nodes[T.recv_ptr[k]:T.recv_ptr[k + 1]]] = recv.numpy()
            return x

        def dot(a, b):
            t = torch.tensor([float(a @ b)], dtype=torch.float64)
            dist.all_reduce(t)
            return float(t.item())

        x_local = b_phase_state(T, seed=5)          # seeded on GLOBAL node ids -> same field on every partition
        x_local = halo(x_local[:NO])                # ghosts through the plan (must reproduce the seeded values)
        assert np.abs(x_local - b_phase_state(T, seed=5)).max() == 0.0
        A, rhs = O.assemble_global(T, x_local, coef, True)
        Minv = O.block_jacobi_inverse(A, T.n_owned_nodes)
        bn = np.sqrt(dot(rhs, rhs))
        d, its, res, ok = O.gmres_block_jacobi(A, rhs, Minv, 1e-6 * bn, matvec_halo=halo, dot=dot)
        np.savez(os.path.join(out_dir, "rank%d.npz" % rank), d=d, its=its, res=res, ok=ok, bn=bn,
                 xyz=T.node_xyz[:T.n_owned_nodes])
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2])
def test_two_rank_gmres_equals_one_rank(tmp_path, world):
    import torch.multiprocessing as mp
    refine = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, refine, str(tmp_path)), nprocs=world, join=True)
    coef = coef_vector(bt=2.0)
    T1 = vh.unit_cube(1, refine, half=2.0).tables(0)
    x1 = b_phase_state(T1, seed=5)
    A1, rhs1 = O.assemble_global(T1, x1, coef, True)
    Minv1 = O.block_jacobi_inverse(A1, T1.n_owned_nodes)
    bn1 = np.linalg.norm(rhs1)
    d1, its1, res1, ok1 = O.gmres_block_jacobi(A1, rhs1, Minv1, 1e-6 * bn1)
    key1 = {tuple(np.round(p, 9)): i for i, p in enumerate(T1.node_xyz)}
    n_seen = 0
    for r in range(world):
        R = np.load(os.path.join(str(tmp_path), "rank%d.npz" % r))
        assert bool(R["ok"]) and int(R["its"]) == its1
        assert abs(float(R["bn"]) - bn1) <= 1e-12 * bn1
        idx = np.array([key1[tuple(np.round(p, 9))] for p in R["xyz"]])
        want = d1.reshape(-1, 18)[idx].ravel()
        assert np.abs(R["d"] - want).max() <= 1e-9 * np.abs(d1).max()
        n_seen += idx.size
    assert n_seen == T1.n_owned_nodes


def test_c4_adaptive_harness_partition_logic_two_ranks():
    """tools/c4_adaptive.py --dry under gloo: across three refinement cycles with a moving Morton partition, the gathered state,
    the refinement flags and the transferred state on 2 ranks equal the 1-rank run (the GPU Newton step is replaced by a fixed
    change of the state keyed on the global node id)."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(root, "tools", "c4_adaptive.py"), "--dry", "--cycles", "3", "--initial-refine", "3",
           "--check-single"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "P-INDEPENDENCE OK" in r.stdout and "C4 ADAPTIVE DONE" in r.stdout
