"""GPU parity: the CUDA hot path (through the C ABI of include/vh_femgl.h) against the CPU oracle on the same
seeded inputs.  Tolerances are the north star's: matrix / residual entries 1e-12 relative (blockrow-scaled, see
SURVEY.md §7 hard part 3), Newton residual norms and energy 1e-10 relative, identical iteration counts."""
import numpy as np
import pytest

import femgl_oracle as O
import verkko_hem_repo_b200 as vh
from helpers import MATEP_SCC_OFF, MATEP_SCC_ON, b_phase_state, blockrow_rel_error, bsr_to_csr, coef_vector

pytestmark = pytest.mark.gpu


def _ctx(T, coef):
    ctx = vh.Context(T)
    ctx.set_coef_vector(coef)
    return ctx


def _check_assembly(T, coef, x, expect_fast=None):
    ctx = _ctx(T, coef)
    NO = 18 * T.n_owned_nodes
    ctx.set_solution(x[:NO])
    rhs_norm = ctx.assemble()
    info = ctx.info()
    if expect_fast is not None:
        assert (info["n_fast_rows"] == T.n_owned_nodes) == expect_fast, info
    A_ora, rhs_ora = O.assemble_global(T, x, coef, True)
    A_gpu = bsr_to_csr(*ctx.export_matrix_bsr(), T.n_local_nodes)
    err = blockrow_rel_error(A_gpu, A_ora)
    assert err <= 1e-12, "matrix blockrow-relative error %.3e" % err
    rhs = ctx.get_rhs()
    assert np.abs(rhs - rhs_ora).max() <= 1e-12 * np.abs(rhs_ora).max()
    assert abs(rhs_norm - np.linalg.norm(rhs_ora)) <= 1e-12 * np.linalg.norm(rhs_ora)
    ctx.close()
    return info


@pytest.mark.parametrize("refine,bt,mat", [(2, 2.0, MATEP_SCC_ON), (3, 2.0, MATEP_SCC_ON), (3, 1e10, MATEP_SCC_OFF)])
def test_q1_uniform_assembly_matches_oracle(refine, bt, mat):
    T = vh.unit_cube(1, refine, half=2.0).tables(0)
    coef = coef_vector(mat, bt)
    _check_assembly(T, coef, b_phase_state(T, mat), expect_fast=True)


def test_q1_all_walls_and_anisotropic_box():
    m = vh.Mesh(1, [-1.0, -2.0, -0.5], [1.5, 1.0, 0.75], base=(2, 3, 1), face_bid=(2, 2, 3, 1, 4, 4), n_global_refine=1).finalize(1)
    T = m.tables(0)
    _check_assembly(T, coef_vector(MATEP_SCC_ON, 0.7), b_phase_state(T), expect_fast=True)


def test_q1_hanging_nodes_general_scatter():
    """Rows next to hanging nodes: the row-owner kernel k_rows_slow."""
    m = vh.Mesh(1, [-2, -2, -2], [2, 2, 2], n_global_refine=2)
    c = m.cell_centers()
    m.refine((np.abs(c[:, 2]) < 1.1) & (c[:, 0] < 0.1))
    m.finalize(1)
    assert m.n_hanging_nodes > 0
    T = m.tables(0)
    info = _check_assembly(T, coef_vector(MATEP_SCC_ON, 2.0), b_phase_state(T))
    assert info["n_slow_cells"] > 0 and info["n_fast_rows"] > 0


@pytest.mark.parametrize("refine", [1, 2])
def test_q2_assembly_matches_oracle(refine):
    """Q2 lattice rows (vertex / edge / face / interior nodes): sum-factorised row-owner kernel, packed storage."""
    T = vh.unit_cube(2, refine, half=2.0).tables(0)
    _check_assembly(T, coef_vector(MATEP_SCC_ON, 2.0), b_phase_state(T), expect_fast=True)


def test_q2_all_walls_and_anisotropic_box():
    m = vh.Mesh(2, [-1.0, -2.0, -0.5], [1.5, 1.0, 0.75], base=(2, 1, 1), face_bid=(2, 2, 3, 1, 4, 4), n_global_refine=1).finalize(1)
    T = m.tables(0)
    _check_assembly(T, coef_vector(MATEP_SCC_ON, 0.7), b_phase_state(T), expect_fast=True)


def test_q2_hanging_nodes():
    m = vh.Mesh(2, [-2, -2, -2], [2, 2, 2], n_global_refine=1)
    fl = np.zeros(m.n_cells, dtype=np.uint8)
    fl[0] = 1
    m.refine(fl)
    m.finalize(1)
    assert m.n_hanging_nodes > 0
    T = m.tables(0)
    info = _check_assembly(T, coef_vector(MATEP_SCC_ON, 2.0), b_phase_state(T))
    assert info["n_slow_cells"] > 0 and info["n_fast_rows"] > 0


def test_residual_and_energy_match_oracle():
    T = vh.unit_cube(1, 3, half=2.0).tables(0)
    coef = coef_vector(MATEP_SCC_ON, 2.0)
    x = b_phase_state(T)
    ctx = _ctx(T, coef)
    ctx.set_solution(x)
    ctx.assemble()
    e_gpu = ctx.energy(0)
    e_ora = O.energy_global(T, x, coef)
    assert abs(e_gpu - e_ora) <= 1e-12 * abs(e_ora)
    its, _ = ctx.solve(1e-1)
    ctx.line_search_trial(0.83)
    r_norm = ctx.residual()
    d = ctx.get_newton_update()
    xt = O.distribute(T, x + 0.83 * d)
    _, r_ora = O.assemble_global(T, xt, coef, False)
    r = ctx.get_residual()
    assert np.abs(r - r_ora).max() <= 1e-12 * np.abs(r_ora).max()
    assert abs(r_norm - np.linalg.norm(r_ora)) <= 1e-12 * np.linalg.norm(r_ora)
    assert abs(ctx.energy(1) - O.energy_global(T, xt, coef)) <= 1e-12 * abs(e_ora)
    ctx.close()


@pytest.mark.parametrize("degree,refine", [(1, 3), (2, 2)])
def test_spmv_and_block_jacobi_match_oracle(degree, refine):
    T = vh.unit_cube(degree, refine, half=2.0).tables(0)
    coef = coef_vector(MATEP_SCC_ON, 2.0)
    x = b_phase_state(T)
    ctx = _ctx(T, coef)
    ctx.set_solution(x)
    ctx.assemble()
    A, _ = O.assemble_global(T, x, coef, True)
    rng = np.random.default_rng(11)
    z = rng.uniform(-1, 1, A.shape[1])
    y = ctx.spmv(z)
    y_ora = A @ z
    assert np.abs(y - y_ora).max() <= 1e-13 * np.abs(y_ora).max()
    Minv = O.block_jacobi_inverse(A, T.n_owned_nodes)
    p = ctx.precondition(z)
    p_ora = np.einsum("ijk,ik->ij", Minv, z.reshape(-1, 18)).ravel()
    assert np.abs(p - p_ora).max() <= 1e-11 * np.abs(p_ora).max()
    ctx.close()


@pytest.mark.parametrize("mgs", ["auto", "64", "1000"])
@pytest.mark.parametrize("degree,refine,tol", [(1, 3, 1e-1), (1, 3, 1e-8), (2, 2, 1e-6)])
def test_gmres_history_matches_oracle(degree, refine, tol, mgs, monkeypatch):
    """mgs: variant of the modified Gram-Schmidt step — register-resident fused kernel (auto at this size), the streaming
    fused kernel that long vectors (C5) take, and the kernel chain (no cooperative launch)."""
    if mgs != "auto":
        monkeypatch.setenv("VH_MGS_MODE", mgs)
    T = vh.unit_cube(degree, refine, half=2.0).tables(0)
    coef = coef_vector(MATEP_SCC_ON, 2.0)
    x = b_phase_state(T)
    ctx = _ctx(T, coef)
    ctx.set_solution(x)
    bn = ctx.assemble()
    its, res = ctx.solve(tol)
    A, rhs = O.assemble_global(T, x, coef, True)
    Minv = O.block_jacobi_inverse(A, T.n_owned_nodes)
    d_ora, its_ora, res_ora, ok = O.gmres_block_jacobi(A, rhs, Minv, tol * np.linalg.norm(rhs))
    assert ok and its == its_ora
    assert abs(res - res_ora) <= 1e-8 * bn
    d = ctx.get_newton_update()
    d_ora = O.distribute(T, d_ora)
    assert np.abs(d - d_ora).max() <= 1e-9 * np.abs(d_ora).max()
    # the solve really solved: ||b - A d|| <= tol ||b|| (true residual, unconstrained rows)
    assert np.linalg.norm(rhs - A @ d) <= 1.01 * tol * np.linalg.norm(rhs)
    ctx.close()


def test_newton_history_matches_oracle():
    """C1-like run (Q1 cube, B-phase IC, z walls Dirichlet-masked + Robin): residual norms and energy to 1e-10,
    identical linear-iteration and line-search counts per Newton step."""
    T = vh.unit_cube(1, 3, half=2.0).tables(0)
    coef = coef_vector(MATEP_SCC_ON, 2.0)
    x = b_phase_state(T, noise=0.0)
    ctx = _ctx(T, coef)
    ctx.set_solution(x)
    x_ora = x.copy()
    for step in range(4):
        o = O.newton_step(T, x_ora, coef, 1e-1)
        bn = ctx.assemble()
        its, _ = ctx.solve(1e-1)
        n_trials = 0
        for i in range(100):
            alpha = 0.83 ** i
            ctx.line_search_trial(alpha)
            cur = ctx.residual()
            n_trials += 1
            if cur < bn:
                break
        ctx.accept_trial()
        assert abs(bn - o["rhs_norm"]) <= 1e-10 * o["rhs_norm"]
        assert its == o["lin_its"] and n_trials == o["n_trials"]
        assert abs(cur - o["res_norm"]) <= 1e-10 * o["res_norm"]
        e_gpu = ctx.energy(0)
        e_ora = O.energy_global(T, o["x"], coef)
        assert abs(e_gpu - e_ora) <= 1e-10 * abs(e_ora)
        x_ora = o["x"]
    assert np.abs(ctx.get_solution() - x_ora).max() <= 1e-9 * np.abs(x_ora).max()
    ctx.close()


def test_errors_are_reported_not_swallowed():
    T = vh.unit_cube(1, 1).tables(0)
    ctx = vh.Context(T)
    with pytest.raises(vh.VhError):
        ctx.assemble()  # coefficients not set
    ctx.set_coef_vector(coef_vector())
    with pytest.raises(vh.VhError):
        ctx.solve(1e-1)  # no matrix yet
    ctx.close()


def test_call_order_guards_of_the_newton_step():
    """VH_ERR_STATE for calls out of the order run.cc:214-218 / iteration.cc:128-210 prescribe; in particular nothing of the
    replaced state survives vh_accept_trial: a second line search, residual, acceptance, trial energy or solve must fail
    instead of applying the stale Newton update once more."""
    T = vh.unit_cube(1, 2, half=2.0).tables(0)
    ctx = vh.Context(T)
    with pytest.raises(vh.VhError):
        ctx.assemble()                         # before vh_set_coefficients
    ctx.set_coef_vector(coef_vector(MATEP_SCC_ON, 2.0))
    ctx.set_solution(b_phase_state(T, seed=3))
    for call in (lambda: ctx.solve(1e-1), lambda: ctx.line_search_trial(1.0), ctx.residual, ctx.accept_trial, lambda: ctx.energy(1)):
        with pytest.raises(vh.VhError):
            call()                             # nothing assembled / solved / tried yet
    ctx.assemble()
    with pytest.raises(vh.VhError):
        ctx.line_search_trial(1.0)             # before vh_solve
    ctx.solve(1e-1)
    with pytest.raises(vh.VhError):
        ctx.accept_trial()                     # before vh_line_search_trial
    ctx.line_search_trial(1.0)
    ctx.residual()
    ctx.energy(1)
    ctx.accept_trial()
    x1 = ctx.get_solution()
    for call in (lambda: ctx.line_search_trial(1.0), ctx.residual, ctx.accept_trial, lambda: ctx.energy(1), lambda: ctx.solve(1e-1)):
        with pytest.raises(vh.VhError):
            call()                             # the update, the trial vector and the matrix belonged to the replaced state
    assert np.array_equal(ctx.get_solution(), x1)
    ctx.energy(0)
    ctx.assemble()                             # the next Newton step starts normally
    ctx.solve(1e-1)
    ctx.close()
