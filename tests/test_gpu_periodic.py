"""GPU parity on the reference's ACTIVE grid (femgl/CMakeLists.txt:51): x/y-periodic slab with AdGR z walls
(makegrid_retangle-z-AdGR_xy-periodic.cc:167-219, setup_weak-coupling-PDW-configuration.cc:128-206).  The periodic
identity constraints reach the CUDA path as ordinary constraint lines of the C ABI: rows next to the periodic seam and
the rows of the image nodes' masters go through the row-owner general scatter (k_rows_slow), everything else through
the lattice-row kernels.  Same tolerances as tests/test_gpu_parity.py."""
import numpy as np
import pytest

import femgl_oracle as O
import verkko_hem_repo_b200 as vh
from helpers import MATEP_SCC_ON, b_phase_state, blockrow_rel_error, bsr_to_csr, coef_vector, gpu_count

pytestmark = pytest.mark.gpu


def _ctx(T, coef):
    ctx = vh.Context(T)
    ctx.set_coef_vector(coef)
    return ctx


@pytest.mark.parametrize("degree,refine", [(1, 2), (1, 3), (2, 1), (2, 2)])
def test_periodic_assembly_matches_oracle(degree, refine):
    T = vh.periodic_slab(degree, refine, half=(1.0, 1.5, 0.75)).tables(0)
    coef = coef_vector(MATEP_SCC_ON, 2.0)
    x = b_phase_state(T, seed=7)
    ctx = _ctx(T, coef)
    ctx.set_solution(x)
    rhs_norm = ctx.assemble()
    info = ctx.info()
    assert info["n_slow_cells"] > 0
    if refine >= 2:
        assert info["n_fast_rows"] > 0
    A_ora, rhs_ora = O.assemble_global(T, x, coef, True)
    A_gpu = bsr_to_csr(*ctx.export_matrix_bsr(), T.n_local_nodes)
    err = blockrow_rel_error(A_gpu, A_ora)
    assert err <= 1e-12, "matrix blockrow-relative error %.3e" % err
    rhs = ctx.get_rhs()
    assert np.abs(rhs - rhs_ora).max() <= 1e-12 * np.abs(rhs_ora).max()
    assert abs(rhs_norm - np.linalg.norm(rhs_ora)) <= 1e-12 * np.linalg.norm(rhs_ora)
    e_ora = O.energy_global(T, x, coef)
    assert abs(ctx.energy(0) - e_ora) <= 1e-12 * abs(e_ora)
    # SpMV over the mixed (packed lattice + full constrained) storage
    z = np.random.default_rng(3).uniform(-1, 1, A_ora.shape[1])
    y, y_ora = ctx.spmv(z), A_ora @ z
    assert np.abs(y - y_ora).max() <= 1e-13 * np.abs(y_ora).max()
    ctx.close()


@pytest.mark.parametrize("degree,refine", [(1, 3), (2, 1)])
def test_periodic_newton_history_matches_oracle(degree, refine):
    """Three damped Newton steps on the periodic slab: GMRES iteration counts and line-search trials identical, residual
    norms / energy to 1e-10, and the accepted state stays periodic (image DoFs equal their masters bit for bit)."""
    T = vh.periodic_slab(degree, refine, half=(2.0, 2.0, 1.0)).tables(0)
    coef = coef_vector(MATEP_SCC_ON, 2.0)
    x_ora = b_phase_state(T, noise=0.02, seed=11)
    ctx = _ctx(T, coef)
    ctx.set_solution(x_ora)
    for step in range(3):
        o = O.newton_step(T, x_ora, coef, 1e-1)
        bn = ctx.assemble()
        its, _ = ctx.solve(1e-1)
        n_trials = 0
        for i in range(100):
            ctx.line_search_trial(0.83 ** i)
            cur = ctx.residual()
            n_trials += 1
            if cur < bn:
                break
        ctx.accept_trial()
        assert abs(bn - o["rhs_norm"]) <= 1e-10 * o["rhs_norm"]
        assert its == o["lin_its"] and n_trials == o["n_trials"]
        assert abs(cur - o["res_norm"]) <= 1e-10 * o["res_norm"]
        e_ora = O.energy_global(T, o["x"], coef)
        assert abs(ctx.energy(0) - e_ora) <= 1e-10 * abs(e_ora)
        x_ora = o["x"]
    sol = ctx.get_solution()
    assert np.abs(sol - x_ora).max() <= 1e-9 * np.abs(x_ora).max()
    assert np.abs(O.distribute(T, sol) - sol).max() == 0.0
    ctx.close()


PERIODIC_PRM = """
subsection physical parameters
  set pressure in bar = 25.0
  set t_reduced = 0.5
  set AdGR diffuse length = 2.0
end
subsection control parameters
  set geometry = retangle-xy-periodic
  set half x length of retangle = 3.0
  set half y length of retangle = 2.0
  set half z length of retangle = 1.5
  set Number of initial global refinments = 3
  set Number of refinements = 0
  set Number of interations = 3
  set Cycle 0 refinement threshold = 1e-12
end
"""


def test_femgl_run_on_the_active_periodic_grid_matches_oracle():
    """FemGL::run() (C++ driver mirror) with the reference's active grid variant selected: per-step record equals the oracle's."""
    out = vh.run_prm(PERIODIC_PRM)
    T = vh.periodic_slab(1, 3, half=(3.0, 2.0, 1.5)).tables(0)
    mat = vh.matep(25.0, 0.5, True)
    coef = np.array([0.42072] * 3 + [mat["alpha"], mat["beta1"], mat["beta2"], mat["beta3"], mat["beta4"], mat["beta5"], 2.0])
    x = b_phase_state(T, noise=0.0)
    assert len(out["history"]) >= 2
    for rec in out["history"]:
        o = O.newton_step(T, x, coef, 1e-1)
        assert abs(rec["rhs_norm"] - o["rhs_norm"]) <= 1e-10 * o["rhs_norm"]
        assert rec["linear_its"] == o["lin_its"] and rec["trials"] == o["n_trials"]
        assert abs(rec["residual"] - o["res_norm"]) <= 1e-10 * o["res_norm"]
        x = o["x"]
    assert np.abs(out["solution"] - x).max() <= 1e-9 * np.abs(x).max()


def test_two_gpu_periodic_run_equals_one_gpu_run():
    import os
    import subprocess
    import sys
    if gpu_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29537", os.path.join(root, "tests", "multigpu_worker.py"), "periodic"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "MULTIGPU PARITY OK" in r.stdout
