"""GPU tests of the host driver mirror (FemGL::run through the C ABI), multi-GPU parity, and size-independent
properties at the BASELINE size C2 (Q1 r5, 646 866 DoFs)."""
import os
import subprocess
import sys

import numpy as np
import pytest

import femgl_oracle as O
import verkko_hem_repo_b200 as vh
from helpers import MATEP_SCC_ON, b_phase_state, coef_vector, gpu_count

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

PRM = """
subsection physical parameters
  set pressure in bar = 25.0
  set t_reduced = 0.5
  set AdGR diffuse length = 2.0
  set trun on Strong Coupling Correction = true
end
subsection control parameters
  set cube half side length = 2.0
  set Number of initial global refinments = %d
  set Number of refinements = %d
  set Number of interations = %d
  set Cycle 0 refinement threshold = %g
  set converge accuracy = 1e-7
end
"""


def test_femgl_run_matches_oracle_newton_loop():
    """C1-shaped run through FemGL::run() (C++ driver mirror): the per-step record (||rhs||, GMRES iterations, line-search
    trials, ||R||, energy) equals the oracle's Newton loop with the stop logic of run.cc:234-250."""
    out = vh.run_prm(PRM % (3, 0, 5, 1e-12))
    hist = out["history"]
    T = vh.unit_cube(1, 3, half=2.0).tables(0)
    coef = coef_vector(MATEP_SCC_ON, 2.0)
    x = b_phase_state(T, noise=0.0)
    assert len(hist) >= 3
    for rec in hist:
        o = O.newton_step(T, x, coef, 1e-1)
        assert abs(rec["rhs_norm"] - o["rhs_norm"]) <= 1e-10 * o["rhs_norm"]
        assert rec["linear_its"] == o["lin_its"] and rec["trials"] == o["n_trials"]
        assert abs(rec["residual"] - o["res_norm"]) <= 1e-10 * o["res_norm"]
        e = O.energy_global(T, o["x"], coef)
        assert abs(rec["energy"] - e) <= 1e-10 * abs(e)
        x = o["x"]
    assert np.abs(out["solution"] - x).max() <= 1e-9 * np.abs(x).max()
    # the reference's observable lines are preserved
    assert "Refinement Cycle is 0" in out["log"] and "Solved in" in out["log"] and "step length alpha is:" in out["log"]


def test_femgl_run_with_adaptive_cycles_converges_and_refines():
    """Two adaptive cycles (hanging nodes -> general scatter path + solution transfer): the run completes, the mesh
    grows, and the residual of the transferred solution keeps decreasing inside each cycle."""
    out = vh.run_prm(PRM % (2, 2, 3, 1e3))
    hist = out["history"]
    cycles = sorted(set(h["cycle"] for h in hist))
    assert cycles == [0, 1, 2]
    for c in cycles:
        rs = [h["residual"] for h in hist if h["cycle"] == c]
        assert all(np.isfinite(rs)) and rs[-1] <= rs[0] * 1.0000001
    assert out["solution"].size > 18 * 125  # refined beyond the 5^3-node start mesh
    assert "adaptive_refine_grid() call is done !" in out["log"]


def test_error_propagates_like_solvercontrol_noconvergence():
    with pytest.raises(RuntimeError) as e:
        vh.run_prm((PRM % (2, 0, 2, 1e-12)).replace("set converge accuracy = 1e-7",
                                                    "set converge accuracy = 1e-7\n  set maximum linear iteration number = 1\n"
                                                    "  set Cycle 0 linear solver tol = 1e-12"))
    assert "no convergence" in str(e.value)


def test_c2_size_properties():
    """BASELINE config C2 (Q1 r5): properties that hold at any size — symmetry of the Jacobian, Jacobian = derivative of
    the residual, residual = -1/2 gradient of the energy, constrained rows decoupled."""
    m = vh.unit_cube(1, 5, half=20.0)
    T = m.tables(0)
    assert 18 * m.n_nodes == 646866
    coef = coef_vector(MATEP_SCC_ON, 2.0)
    x = b_phase_state(T)
    ctx = vh.Context(T)
    ctx.set_coef_vector(coef)
    ctx.set_solution(x)
    ctx.assemble()
    info = ctx.info()
    assert info["nnzb"] == 912673 and info["n_fast_rows"] == m.n_nodes and info["n_slow_cells"] == 0
    rng = np.random.default_rng(3)
    con = np.zeros(x.size, dtype=bool)
    con[T.c_dof] = True
    u = rng.uniform(-1, 1, x.size)
    v = rng.uniform(-1, 1, x.size)
    Au, Av = ctx.spmv(u), ctx.spmv(v)
    assert abs(v @ Au - u @ Av) <= 1e-11 * abs(v @ Au)             # symmetric operator
    z = np.where(con, 1.0, 0.0)
    Az = ctx.spmv(z)
    assert np.abs(Az[~con]).max() == 0.0 and (Az[con] > 0).all()    # constrained DoFs: positive diagonal only
    rhs0 = ctx.get_rhs()
    assert np.abs(rhs0[con]).max() == 0.0
    # directional derivative of the residual along a constraint-compatible direction d
    d = np.where(con, 0.0, rng.uniform(-1, 1, x.size))
    eps = 1e-5
    ctx2 = vh.Context(T)
    ctx2.set_coef_vector(coef)
    rp = []
    en = []
    for s in (+1, -1):
        ctx2.set_solution(x + s * eps * d)
        ctx2.assemble()
        rp.append(ctx2.get_rhs())
        en.append(ctx2.energy(0))
    Ad = ctx.spmv(d)
    fd = -(rp[0] - rp[1]) / (2 * eps)
    assert np.abs(fd[~con] - Ad[~con]).max() <= 1e-6 * np.abs(Ad).max()
    # rhs = -R = -1/2 dF/dx
    assert abs((en[0] - en[1]) / (2 * eps) - (-2.0 * (rhs0 @ d))) <= 1e-6 * abs(2.0 * (rhs0 @ d))
    ctx.close()
    ctx2.close()


@pytest.mark.parametrize("mode", ["cube", "mg"])
def test_two_gpu_run_equals_one_gpu_run(mode):
    """mode mg: the same with the multigrid V-cycle preconditioner (three levels, every level partitioned over both ranks)."""
    if gpu_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533" if mode == "cube" else "29534", os.path.join(ROOT, "tests", "multigpu_worker.py")] + \
          ([] if mode == "cube" else [mode])
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "MULTIGPU PARITY OK" in r.stdout


SLAB_PRM = """
subsection physical parameters
  set pressure in bar = 25.0
  set t_reduced = 0.5
  set AdGR diffuse length = 2.0
end
subsection control parameters
  set geometry = retangle
  set initial condition = BnA
  set half x length of retangle = 3.0
  set half y length of retangle = 2.0
  set half z length of retangle = 4.0
  set A-phase block range ratio = 0.1
  set Number of initial global refinments = 2
  set Number of refinements = %d
  set Number of interations = 2
  set Cycle 0 refinement threshold = %g
end
"""


def test_slab_with_ab_interface_matches_oracle_first_steps():
    """C4-shaped configuration (hyper_rectangle slab, flat A/B wall initial condition BnA.h:130-163, AdGR z walls):
    the first Newton steps through FemGL::run() equal the oracle's on the same anisotropic mesh."""
    out = vh.run_prm(SLAB_PRM % (0, 1e-12))
    m = vh.Mesh(1, [-3, -2, -4], [3, 2, 4], n_global_refine=2).finalize(1)
    T = m.tables(0)
    mat = vh.matep(25.0, 0.5, True)
    coef = np.array([0.42072] * 3 + [mat["alpha"], mat["beta1"], mat["beta2"], mat["beta3"], mat["beta4"], mat["beta5"], 2.0])
    x = np.zeros((T.n_local_nodes, 18))
    isB = T.node_xyz[:, 2] >= 0.1 * 4.0
    eA = mat["gapA"] * float(np.float32(0.7071067811865475))
    eB = mat["gapB"] * float(np.float32(0.5773502691896258))
    x[isB, 0] = x[isB, 4] = x[isB, 8] = eB
    x[~isB, 0] = x[~isB, 10] = eA
    x = O.distribute(T, x.ravel())
    for rec in out["history"]:
        o = O.newton_step(T, x, coef, 1e-1)
        assert abs(rec["rhs_norm"] - o["rhs_norm"]) <= 1e-10 * o["rhs_norm"]
        assert rec["linear_its"] == o["lin_its"] and rec["trials"] == o["n_trials"]
        assert abs(rec["residual"] - o["res_norm"]) <= 1e-10 * o["res_norm"]
        x = o["x"]


def test_slab_adaptive_cycles_run_to_completion():
    out = vh.run_prm(SLAB_PRM % (2, 1e3))
    hist = out["history"]
    assert sorted(set(h["cycle"] for h in hist)) == [0, 1, 2]
    assert all(np.isfinite(h["residual"]) and np.isfinite(h["energy"]) for h in hist)
    # the refinement concentrates on the A/B interface: the mesh grows but stays far below uniform refinement
    n0 = 18 * 5 * 5 * 5
    assert n0 < out["solution"].size < 18 * 17 ** 3


def test_device_side_solution_transfer_matches_host_interpolation():
    """refine.cc:128-130/171-175 on the device (vh_transfer_solution): the state transferred to the refined mesh equals the
    host's FE interpolation followed by constraints_solution.distribute, hanging nodes included; the driver reports the
    rebuild time of the cycle."""
    old = vh.Mesh(1, [-3, -2, -4], [3, 2, 4], n_global_refine=2).finalize(1)
    new = old.clone()
    d = np.abs(new.cell_centers()[:, 2] - 0.4)
    new.refine(d <= np.sort(d)[int(0.3 * new.n_cells)])
    new.finalize(1)
    assert new.n_hanging_nodes > 0
    To, Tn = old.tables(0), new.tables(0)
    coef = coef_vector(MATEP_SCC_ON, 2.0)
    xo = b_phase_state(To, seed=21)
    co, cn = vh.Context(To), vh.Context(Tn)
    for c in (co, cn):
        c.set_coef_vector(coef)
    co.set_solution(xo)
    cn.transfer_solution_from(co, *new.transfer_table(old))
    from helpers import distribute_constraints
    want = distribute_constraints(Tn, new.interpolate_from(old, xo))
    got = cn.get_solution()
    assert np.abs(got - want).max() <= 1e-14 * np.abs(want).max()
    # the transferred state is a usable Newton state: one step on the new mesh equals the oracle's from the same state
    bn = cn.assemble()
    o = O.newton_step(Tn, want, coef, 1e-1)
    assert abs(bn - o["rhs_norm"]) <= 1e-10 * o["rhs_norm"]
    co.close()
    cn.close()
    out = vh.run_prm(SLAB_PRM % (1, 1e3))
    first = [h for h in out["history"] if h["iteration"] == 0]
    assert len(first) == 2 and all(h["t_setup_ms"] > 0 for h in first)
    assert "solution transfer" in out["log"]


@pytest.mark.parametrize("world", [1, 2])
def test_c4_adaptive_cycles_multi_gpu_equal_single_gpu(world):
    """BASELINE configs[3] through tools/c4_adaptive.py: slab with the A/B interface, adaptive cycles with hanging nodes under
    the reference's cycle rule (run.cc:182-256).  world = 2: the partition moves with every cycle; rank 0 repeats the run on one
    GPU and requires identical refinement flags and iteration counts and residual norms within 1e-10."""
    if gpu_count() < world:
        pytest.skip("needs %d GPUs" % world)
    tool = os.path.join(ROOT, "tools", "c4_adaptive.py")
    if world == 1:
        cmd = [sys.executable, tool, "--cycles", "2", "--initial-refine", "3", "--half", "3", "2", "4"]
    else:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
               "--master-port", "29541", tool, "--cycles", "2", "--initial-refine", "3", "--half", "3", "2", "4", "--check-single"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "C4 ADAPTIVE DONE" in r.stdout
    if world > 1:
        assert "P-INDEPENDENCE OK" in r.stdout


def test_snapshot_is_taken_in_stream_order_and_survives_the_next_step():
    """vh_snapshot_begin / vh_snapshot_wait (output path, io.cc:106-170): the snapshot holds local_solution and the Newton
    update as they were when it was requested, although the next Newton step runs before it is collected."""
    T = vh.unit_cube(1, 3, half=2.0).tables(0)
    ctx = vh.Context(T)
    ctx.set_coef_vector(coef_vector(MATEP_SCC_ON, 2.0))
    ctx.set_solution(b_phase_state(T, seed=4))
    ctx.snapshot_begin()                       # before any solve: the update block is zero
    s0, u0 = ctx.snapshot_wait()
    assert np.array_equal(s0[:ctx.n_owned], ctx.get_solution()) and not u0.any()

    def step():
        bn = ctx.assemble()
        ctx.solve(1e-1)
        for i in range(100):
            ctx.line_search_trial(0.83 ** i)
            if ctx.residual() < bn:
                break
        ctx.accept_trial()

    step()
    sol1, upd1 = ctx.get_solution(), ctx.get_newton_update()
    ctx.snapshot_begin()
    step()                                     # the state moves on while the snapshot travels
    s, u = ctx.snapshot_wait()
    assert np.array_equal(s[:ctx.n_owned], sol1) and np.array_equal(u[:ctx.n_owned], upd1)
    assert not np.array_equal(ctx.get_solution(), sol1)
    ctx.close()


def test_femgl_run_writes_vtu_after_every_newton_step(tmp_path, monkeypatch):
    """FemGL::output_results of the mirror (additive key "write vtu output"): one .vtu + .pvtu per Newton step in the
    reference's directories (run.cc:221-227), written by a thread next to the Newton loop; the last file holds the final state."""
    from test_vtu_output import NAMES, read_vtu
    monkeypatch.chdir(tmp_path)
    prm = SLAB_PRM % (1, 1e3)
    out = vh.run_prm(prm.replace("set geometry = retangle", "set geometry = retangle\n  set write vtu output = true"))
    hist = out["history"]
    assert os.path.exists(tmp_path / "setup_config" / "solution_00.0.vtu")
    for h in hist:
        d = tmp_path / ("refine-cycle_%d" % h["cycle"])
        assert os.path.exists(d / ("solution_%02d.0.vtu" % h["iteration"])) and os.path.exists(d / ("solution_%02d.pvtu" % h["iteration"]))
    last = hist[-1]
    v = read_vtu(str(tmp_path / ("refine-cycle_%d" % last["cycle"]) / ("solution_%02d.0.vtu" % last["iteration"])))
    sol = out["solution"].reshape(-1, 18)
    assert v["n_points"] == sol.shape[0]
    for c in range(18):
        assert np.array_equal(v[NAMES[18 + c]], sol[:, c])
    assert np.abs(np.stack([v[n] for n in NAMES[:18]])).max() > 0.0       # the Newton update of the last step
    assert "output: the Newton loop waited" in out["log"]
