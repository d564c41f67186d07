"""Shared fixtures for the parity tests: coefficient sets, initial states, BSR -> scipy conversion."""
import numpy as np

K123 = 0.42072
# Matep at p = 25 bar, t = 0.5 (reproduced bit-exact by oracle/_ref in tests/golden/matep.json)
MATEP_SCC_ON = dict(alpha=-0.5, beta=(-0.010850915879921348, 0.020598836429398658, 0.02117724292551364,
                                      0.019780381922869742, -0.023091499302424053), gapB=3.9900156313155422)
MATEP_SCC_OFF = dict(alpha=-0.5, beta=(-0.010656959214250182, 0.021313918428500365, 0.021313918428500365,
                                       0.021313918428500365, -0.021313918428500365), gapB=3.7517075313463422)


def gpu_count():
    """Visible CUDA devices, without importing torch (its first import on a fresh box takes about a minute)."""
    import os
    import subprocess
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis is not None:
        return len([v for v in vis.split(",") if v.strip()])
    try:
        out = subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True, timeout=60).stdout
        return sum(1 for line in out.splitlines() if line.startswith("GPU "))
    except Exception:
        return 0


def coef_vector(mat=MATEP_SCC_ON, bt=2.0):
    """[K1,K2,K3,alpha,beta1..5,bt] — K1=K2=K3=0.42072 (femgl.h:320-322)."""
    return np.array([K123, K123, K123, mat["alpha"], *mat["beta"], bt], dtype=np.float64)


def b_phase_state(T, mat=MATEP_SCC_ON, noise=0.05, seed=20250101):
    """Uniform B-phase u11=u22=u33=gapB*0.577350269f (setup_uniform_B-phase.cc:245-259) plus a seeded
    perturbation keyed on the GLOBAL node id (so every partition sees the same field), constraints distributed."""
    amp = mat["gapB"] * float(np.float32(0.577350269))
    x = np.zeros((T.n_local_nodes, 18))
    x[:, [0, 4, 8]] = amp
    if noise:
        gmax = int(T.node_global.max()) + 1 if T.n_local_nodes else 1
        rng = np.random.default_rng(seed)
        full = rng.uniform(-1.0, 1.0, size=(gmax, 18)) if gmax < 4_000_000 else None
        if full is not None:
            x += noise * mat["gapB"] * full[T.node_global]
        else:  # hash-based for very large meshes
            g = T.node_global.astype(np.uint64)[:, None] * np.uint64(18) + np.arange(18, dtype=np.uint64)[None, :]
            h = (g * np.uint64(0x9E3779B97F4A7C15) + np.uint64(seed)) & np.uint64(0xFFFFFFFFFFFF)
            x += noise * mat["gapB"] * (h.astype(np.float64) / float(0xFFFFFFFFFFFF) * 2.0 - 1.0)
    return distribute_constraints(T, x.ravel())


def distribute_constraints(T, x_local):
    """constraints_solution.distribute (setup_uniform_B-phase.cc:262): constrained entries from their masters, Dirichlet
    entries zero.  Plain numpy on the host tables — input preparation for tests, tools and bench.py, independent of oracle/."""
    x = np.array(x_local, dtype=np.float64, copy=True)
    if T.c_dof.size:
        cnt = np.diff(T.c_ptr)
        vals = np.zeros(T.c_dof.size)
        if T.c_master.size:
            np.add.at(vals, np.repeat(np.arange(T.c_dof.size), cnt), T.c_weight * x_local[T.c_master])
        x[T.c_dof] = vals
    return x


def bsr_to_csr(row_ptr, col, vals, n_local_nodes):
    import scipy.sparse as sp
    nb = row_ptr.size - 1
    B = sp.bsr_matrix((vals, col, row_ptr), shape=(18 * nb, 18 * n_local_nodes))
    return B.tocsr()


def blockrow_rel_error(A_gpu, A_ora):
    """max over entries of |a-b| / (max |entry| of that node's 18 rows in the oracle matrix)."""
    D = abs(A_gpu - A_ora).tocsr()
    Aabs = abs(A_ora).tocsr()
    n = A_ora.shape[0]
    rowmax = np.zeros(n)
    rmax_ora = Aabs.max(axis=1).toarray().ravel()
    rowmax[:] = rmax_ora
    scale = np.repeat(rowmax.reshape(-1, 18).max(axis=1), 18)
    dmax = D.max(axis=1).toarray().ravel()
    scale = np.where(scale > 0, scale, 1.0)
    return float((dmax / scale).max())
