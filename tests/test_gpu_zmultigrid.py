"""GPU parity of the multigrid V-cycle preconditioner (csrc/vh_multigrid.cu, vh_mg_attach / vh_set_preconditioner) against
the oracle running the same cycle: one V-cycle applied to a vector, the GMRES history, and the Newton history."""
import numpy as np
import pytest

import femgl_oracle as O
import verkko_hem_repo_b200 as vh
from helpers import MATEP_SCC_ON, b_phase_state, coef_vector

pytestmark = pytest.mark.gpu


def _hierarchy(refine, n_levels, half=20.0, bt=2.0):
    meshes = [vh.Mesh(1, [-half] * 3, [half] * 3, n_global_refine=refine - k).finalize(1) for k in range(n_levels)]
    tabs = [m.tables(0) for m in meshes]
    prol = [vh.mg_prolongation(meshes[k], tabs[k], meshes[k + 1], tabs[k + 1]) for k in range(n_levels - 1)]
    coef = coef_vector(MATEP_SCC_ON, bt)
    ctxs = [vh.Context(T) for T in tabs]
    for c in ctxs:
        c.set_coef_vector(coef)
    for k in range(n_levels - 1):
        ctxs[k].mg_attach(ctxs[k + 1], *prol[k])
    return tabs, prol, coef, ctxs


@pytest.mark.parametrize("refine,n_levels,half,params", [(3, 2, 2.0, {}), (4, 3, 2.0, {}), (4, 2, 20.0, {}),
                                                         (3, 2, 2.0, dict(pre=2, post=3, smoothing_range=8.0, coarse_degree=5)),
                                                         (4, 3, 2.0, dict(pre=0, post=1)), (4, 3, 2.0, dict(pre=1, post=0)),
                                                         (3, 1, 2.0, dict(coarse_degree=4, coarse_range=10.0))])  # no hierarchy: Chebyshev polynomial
def test_vcycle_and_gmres_history_match_oracle(refine, n_levels, half, params):
    """half = 2: cells much smaller than the coherence length (gradient-dominated, the regime of the fine BASELINE meshes);
    half = 20, r4 -> r3: the coarsest spacing at which the re-discretised coarse operator still helps (tools/mg_experiment.py)."""
    tabs, prol, coef, ctxs = _hierarchy(refine, n_levels, half=half)
    T, fine = tabs[0], ctxs[0]
    x = b_phase_state(T, seed=5)
    fine.set_preconditioner("multigrid", **params)
    fine.set_solution(x)
    bn = fine.assemble()
    levels, lam = O.mg_setup(tabs, prol, x, coef, **params)
    A = levels[0]["A"]
    # one V-cycle on a (Dirichlet-masked) vector
    v = np.random.default_rng(2).uniform(-1, 1, A.shape[0])
    v[levels[0]["mask"]] = 0.0
    z = fine.precondition(v)
    for k in range(n_levels):  # the eigenvalue bounds of the smoothers
        assert abs(fine.mg_lambda(k) - lam[k]) <= 1e-10 * lam[k], (k, fine.mg_lambda(k), lam[k])
    z_ora = O.mg_vcycle(levels, 0, v, **params)
    assert np.abs(z - z_ora).max() <= 1e-11 * np.abs(z_ora).max()
    # GMRES with the V-cycle as right preconditioner: identical iteration count, same update
    for tol in (1e-1, 1e-6):
        its, res = fine.solve(tol)
        rhs = fine.get_rhs()
        d_ora, its_ora, res_ora, ok = O.gmres_right(A, rhs, lambda u: O.mg_vcycle(levels, 0, u, **params), tol * np.linalg.norm(rhs))
        assert ok and its == its_ora, (its, its_ora)
        d = fine.get_newton_update()
        d_ora = O.distribute(T, d_ora)
        assert np.abs(d - d_ora).max() <= 1e-8 * np.abs(d_ora).max()
        assert np.linalg.norm(rhs - A @ d) <= 1.01 * tol * bn
    # block-Jacobi is still selectable on the same context
    fine.set_preconditioner("block-jacobi")
    its_bj, _ = fine.solve(1e-1)
    Minv = O.block_jacobi_inverse(A, T.n_owned_nodes)
    _, its_bj_ora, _, _ = O.gmres_block_jacobi(A, rhs, Minv, 1e-1 * np.linalg.norm(rhs))
    assert its_bj == its_bj_ora
    for c in ctxs:
        c.close()


def test_multigrid_newton_history_matches_oracle():
    tabs, prol, coef, ctxs = _hierarchy(3, 2, half=2.0)
    T, fine = tabs[0], ctxs[0]
    fine.set_preconditioner("multigrid")
    x_ora = b_phase_state(T, noise=0.0)
    fine.set_solution(x_ora)
    lam = None
    for _ in range(3):
        bn = fine.assemble()
        its, _ = fine.solve(1e-1)
        n_trials = 0
        for i in range(100):
            fine.line_search_trial(0.83 ** i)
            cur = fine.residual()
            n_trials += 1
            if cur < bn:
                break
        fine.accept_trial()
        o, lam = O.newton_step_mg(tabs, prol, x_ora, coef, 1e-1, lam=lam)
        assert abs(bn - o["rhs_norm"]) <= 1e-10 * o["rhs_norm"]
        assert its == o["lin_its"] and n_trials == o["n_trials"]
        assert abs(cur - o["res_norm"]) <= 1e-10 * o["res_norm"]
        x_ora = o["x"]
    assert np.abs(fine.get_solution() - x_ora).max() <= 1e-9 * np.abs(x_ora).max()
    for c in ctxs:
        c.close()


def test_multigrid_attach_errors():
    tabs, prol, coef, ctxs = _hierarchy(3, 2)
    other = vh.Context(tabs[1])
    with pytest.raises(vh.VhError):
        ctxs[0].mg_attach(other, *prol[0])                 # already has a coarse level
    other.set_preconditioner("multigrid")                  # nothing attached: the Chebyshev polynomial of block-Jacobi (allowed)
    other.set_preconditioner("block-jacobi")
    with pytest.raises(vh.VhError):
        other.set_preconditioner("multigrid", pre=0, post=0)   # a cycle without any smoothing
    ptr, cn, w = prol[0]
    lone = vh.Context(tabs[0])
    with pytest.raises(vh.VhError):
        lone.mg_attach(other, ptr, cn, 0.5 * w)            # incomplete interpolation rows
    # destroying the coarse level first falls back to block-Jacobi instead of leaving a dangling level
    ctxs[0].set_preconditioner("multigrid")
    ctxs[1].close()
    ctxs[0].set_coef_vector(coef)
    ctxs[0].set_solution(b_phase_state(tabs[0]))
    ctxs[0].assemble()
    its, _ = ctxs[0].solve(1e-1)
    assert its > 0
    for c in (ctxs[0], other, lone):
        c.close()
