"""GPU parity of the matrix-free operator apply (the default; VH_SPMV_MF=0 selects the assembled packed SpMV): inside vh_solve / vh_spmv the lattice rows are applied as
sum_cells K_cell z_cell, with the bulk part H_q z_q read back from the packed H_q tables of the assembly and the gradient /
Robin forms evaluated from z (k_points<APPLY> + k_gather_apply), instead of streaming the assembled blocks.  The result
must equal the assembled operator: against the oracle's matrix (1e-13) and against the default SpMV of a second context,
with identical GMRES / Newton histories."""
import os

import numpy as np
import pytest

import femgl_oracle as O
import verkko_hem_repo_b200 as vh
from helpers import MATEP_SCC_ON, b_phase_state, coef_vector

pytestmark = pytest.mark.gpu

NAMES = ["q1-cube", "q1-walls-aniso", "q1-specular", "q2-cube", "q2-walls-aniso", "q1-hanging", "q1-periodic"]


def _mesh(name):
    """-> (per-rank tables, AdGR diffuse length bt)"""
    if name == "q1-cube":
        return vh.unit_cube(1, 3, half=2.0).tables(0), 2.0
    if name == "q1-walls-aniso":
        m = vh.Mesh(1, [-1.0, -2.0, -0.5], [1.5, 1.0, 0.75], base=(2, 3, 1), face_bid=(2, 2, 3, 1, 4, 4), n_global_refine=1)
        return m.finalize(1).tables(0), 0.7
    if name == "q1-specular":
        return vh.unit_cube(1, 2, half=2.0).tables(0), 1e10
    if name == "q2-cube":
        return vh.unit_cube(2, 2, half=2.0).tables(0), 2.0
    if name == "q2-walls-aniso":
        m = vh.Mesh(2, [-1.0, -2.0, -0.5], [1.5, 1.0, 0.75], base=(2, 1, 1), face_bid=(2, 2, 3, 1, 4, 4), n_global_refine=1)
        return m.finalize(1).tables(0), 0.7
    if name == "q1-hanging":
        m = vh.Mesh(1, [-2, -2, -2], [2, 2, 2], n_global_refine=2)
        c = m.cell_centers()
        m.refine((np.abs(c[:, 2]) < 1.1) & (c[:, 0] < 0.1))
        return m.finalize(1).tables(0), 2.0
    if name == "q1-periodic":
        return vh.periodic_slab(1, 3, half=(1.0, 1.5, 0.75)).tables(0), 2.0
    raise KeyError(name)


@pytest.mark.parametrize("name", NAMES)
def test_matrix_free_apply_equals_assembled_operator(name, monkeypatch):
    T, bt = _mesh(name)
    coef = coef_vector(MATEP_SCC_ON, bt)
    x = b_phase_state(T, seed=13)
    A, _ = O.assemble_global(T, x, coef, True)
    rng = np.random.default_rng(17)
    zs = [rng.uniform(-1, 1, A.shape[1]) for _ in range(2)]
    mf = vh.Context(T)                       # default: matrix-free apply, lattice rows not assembled
    monkeypatch.setenv("VH_SPMV_MF", "0")
    ref = vh.Context(T)                      # assembled packed SpMV
    monkeypatch.delenv("VH_SPMV_MF")
    assert mf.info()["spmv_matrix_free"] == 1 and ref.info()["spmv_matrix_free"] == 0
    for ctx in (mf, ref):
        ctx.set_coef_vector(coef)
        ctx.set_solution(x)
        ctx.assemble()
    for z in zs:
        y_ora = A @ z
        y_mf, y_ref = mf.spmv(z), ref.spmv(z)
        assert np.abs(y_mf - y_ora).max() <= 1e-13 * np.abs(y_ora).max(), name
        assert np.abs(y_mf - y_ref).max() <= 1e-13 * np.abs(y_ora).max(), name
    # a second apply after a solve and a residual evaluation (the cell scratch is shared with the residual path)
    mf.solve(1e-1)
    mf.line_search_trial(1.0)
    mf.residual()
    y_ora = A @ zs[0]
    assert np.abs(mf.spmv(zs[0]) - y_ora).max() <= 1e-13 * np.abs(y_ora).max()
    # the mode is also a run-time switch of the ABI (vh_set_spmv_matrix_free)
    ref.set_spmv_matrix_free(True)
    assert ref.info()["spmv_matrix_free"] == 1
    assert np.abs(ref.spmv(zs[1]) - A @ zs[1]).max() <= 1e-13 * np.abs(A @ zs[1]).max()
    ref.set_spmv_matrix_free(False)
    assert np.abs(ref.spmv(zs[1]) - y_ref).max() == 0.0
    mf.close()
    ref.close()


@pytest.mark.parametrize("degree,refine,tol", [(1, 3, 1e-1), (1, 3, 1e-8), (2, 2, 1e-6)])
def test_matrix_free_gmres_history_matches_oracle(degree, refine, tol, monkeypatch):
    monkeypatch.setenv("VH_SPMV_MF", "1")
    T = vh.unit_cube(degree, refine, half=2.0).tables(0)
    coef = coef_vector(MATEP_SCC_ON, 2.0)
    x = b_phase_state(T)
    ctx = vh.Context(T)
    assert ctx.info()["spmv_matrix_free"] == 1
    ctx.set_coef_vector(coef)
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
    assert np.linalg.norm(rhs - A @ d) <= 1.01 * tol * np.linalg.norm(rhs)
    ctx.close()


def test_matrix_free_newton_history_matches_oracle(monkeypatch):
    monkeypatch.setenv("VH_SPMV_MF", "1")
    T = vh.unit_cube(1, 3, half=2.0).tables(0)
    coef = coef_vector(MATEP_SCC_ON, 2.0)
    x_ora = b_phase_state(T, noise=0.0)
    ctx = vh.Context(T)
    ctx.set_coef_vector(coef)
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
        x_ora = o["x"]
    assert np.abs(ctx.get_solution() - x_ora).max() <= 1e-9 * np.abs(x_ora).max()
    ctx.close()


@pytest.mark.parametrize("mode", [2, 3])
@pytest.mark.parametrize("name", NAMES)
def test_table_free_apply_equals_assembled_operator(name, mode):
    """Mode 2 of the operator apply: H(A_q) z_q evaluated from the Newton state (vh_hessian_apply), no H_q table read.
    Mode 3: second formulation of the table apply at Q1 (csrc/vh_apply_v2.cuh; Q2 contexts fall back to mode 1)."""
    T, bt = _mesh(name)
    coef = coef_vector(MATEP_SCC_ON, bt)
    x = b_phase_state(T, seed=13)
    A, _ = O.assemble_global(T, x, coef, True)
    ctx = vh.Context(T)
    ctx.set_coef_vector(coef)
    ctx.set_solution(x)
    ctx.assemble()
    ctx.set_spmv_matrix_free(mode)
    assert ctx.info()["spmv_matrix_free"] == mode
    rng = np.random.default_rng(19)
    for _ in range(2):
        z = rng.uniform(-1, 1, A.shape[1])
        y_ora = A @ z
        assert np.abs(ctx.spmv(z) - y_ora).max() <= 1e-13 * np.abs(y_ora).max(), name
    its, _ = ctx.solve(1e-1)
    ctx.set_spmv_matrix_free(0)
    its0, _ = ctx.solve(1e-1)
    assert its == its0
    ctx.close()


@pytest.mark.parametrize("name", ["q1-cube", "q1-walls-aniso", "q2-cube", "q1-hanging", "q1-periodic"])
def test_lazy_rows_block_jacobi_and_solve_match_oracle(name, monkeypatch):
    """The default mode (matrix-free apply, lazy rows): vh_assemble forms only the diagonal blocks of the lattice rows
    (k_diag_cells + k_diag_gather).  The preconditioner, the GMRES history and the update must equal the oracle's, and the
    rows must appear on demand (export) identical to an ordinary assembly."""
    monkeypatch.setenv("VH_SPMV_MF", "1")
    monkeypatch.setenv("VH_MF_LAZY_ROWS", "1")
    T, bt = _mesh(name)
    coef = coef_vector(MATEP_SCC_ON, bt)
    x = b_phase_state(T, seed=13)
    A, rhs = O.assemble_global(T, x, coef, True)
    ctx = vh.Context(T)
    ctx.set_coef_vector(coef)
    ctx.set_solution(x)
    ctx.assemble()
    Minv = O.block_jacobi_inverse(A, T.n_owned_nodes)
    z = np.random.default_rng(29).uniform(-1, 1, A.shape[1])
    p_ora = np.einsum("ijk,ik->ij", Minv, z.reshape(-1, 18)[:T.n_owned_nodes]).ravel()
    assert np.abs(ctx.precondition(z) - p_ora).max() <= 1e-11 * np.abs(p_ora).max()
    its, _ = ctx.solve(1e-1)
    d_ora, its_ora, _, ok = O.gmres_block_jacobi(A, rhs, Minv, 1e-1 * np.linalg.norm(rhs))
    assert ok and its == its_ora
    d_ora = O.distribute(T, d_ora)
    assert np.abs(ctx.get_newton_update() - d_ora).max() <= 1e-9 * np.abs(d_ora).max()
    from helpers import blockrow_rel_error, bsr_to_csr
    A_gpu = bsr_to_csr(*ctx.export_matrix_bsr(), T.n_local_nodes)   # assembles the stale rows on demand
    assert blockrow_rel_error(A_gpu, A) <= 1e-12
    ctx.set_spmv_matrix_free(0)                                      # and the packed SpMV works afterwards
    y_ora = A @ z
    assert np.abs(ctx.spmv(z) - y_ora).max() <= 1e-13 * np.abs(y_ora).max()
    ctx.close()
