"""Host side of refine_grid (refine.cc:109-181) in the mini host: the SolutionTransfer table and the Kelly-type indicator."""
import numpy as np
import pytest

import verkko_hem_repo_b200 as vh


def _refined_pair(degree, seed=5):
    old = vh.Mesh(degree, [-1.0, -0.5, 0.0], [1.0, 1.0, 1.5], base=(2, 1, 1), face_bid=(1, 1, 2, 1, 4, 4), n_global_refine=1).finalize(1)
    new = old.clone()
    rng = np.random.default_rng(seed)
    new.refine(rng.uniform(size=new.n_cells) < 0.4)
    new.finalize(1)
    return old, new


@pytest.mark.parametrize("degree", [1, 2])
def test_transfer_table_is_the_fe_interpolation(degree):
    old, new = _refined_pair(degree)
    ptr, src, w = new.transfer_table(old)
    assert ptr.size == new.n_nodes + 1 and ptr[-1] == src.size == w.size
    rowsum = np.add.reduceat(w, ptr[:-1])
    assert np.abs(rowsum - 1.0).max() <= 1e-14                      # partition of unity
    # a polynomial of the element's degree is reproduced exactly at the new nodes
    xo, xn = old.node_xyz(), new.node_xyz()
    f = (lambda X: 1.0 + 2.0 * X[:, 0] - 0.5 * X[:, 1] + 0.25 * X[:, 2]) if degree == 1 else \
        (lambda X: 1.0 + X[:, 0] * X[:, 1] - 0.5 * X[:, 2] ** 2 + X[:, 0] ** 2)
    vals = np.add.reduceat(w * f(xo)[src], ptr[:-1])
    assert np.abs(vals - f(xn)).max() <= 1e-13
    # and it is what interpolate_from computes
    field = np.random.default_rng(1).uniform(-1, 1, (old.n_nodes, 18))
    got = np.zeros((new.n_nodes, 18))
    np.add.at(got, np.repeat(np.arange(new.n_nodes), np.diff(ptr)), w[:, None] * field[src])
    assert np.abs(got.ravel() - new.interpolate_from(old, field.ravel())).max() <= 1e-14


def test_kelly_indicator_sees_gradient_jumps_only():
    m = vh.Mesh(1, [-1.0] * 3, [1.0] * 3, n_global_refine=3).finalize(1)
    X = m.node_xyz()
    lin = np.zeros((m.n_nodes, 18))
    lin[:, 0] = 1.0 + 2.0 * X[:, 0] - X[:, 2]
    lin[:, 11] = 0.3 * X[:, 1]
    assert np.abs(m.kelly_indicator(lin.ravel())).max() <= 1e-12     # a globally linear field has no jumps
    kink = np.zeros((m.n_nodes, 18))
    kink[:, 4] = np.abs(X[:, 2] - 0.25)                              # kink on the cell faces at z = 0.25
    eta = m.kelly_indicator(kink.ravel())
    c = m.cell_centers()
    near = np.abs(c[:, 2] - 0.25) < 0.13
    assert eta[near].min() > 0 and np.abs(eta[~near]).max() <= 1e-12
    # jump 2 across a face of area h^2, h_K = sqrt(3) h:  eta^2 = h_K/24 * h^2 * 4
    h = 0.25
    assert np.allclose(eta[near], np.sqrt(np.sqrt(3) * h / 24 * h * h * 4.0), rtol=1e-12)


def test_kelly_indicator_across_hanging_faces_and_periodic_pairs():
    m = vh.Mesh(1, [-1.0] * 3, [1.0] * 3, n_global_refine=2)
    c = m.cell_centers()
    m.refine(c[:, 0] < 0)                                            # a refinement interface at x = 0
    m.finalize(1)
    X = m.node_xyz()
    lin = np.zeros((m.n_nodes, 18))
    lin[:, 2] = X[:, 0] + 0.5 * X[:, 1]
    assert np.abs(m.kelly_indicator(lin.ravel())).max() <= 1e-12     # also across the hanging faces
    per = vh.periodic_slab(1, 2, half=(1.0, 1.0, 1.0))
    Xp = per.node_xyz()
    f = np.zeros((per.n_nodes, 18))
    f[:, 0] = np.abs(Xp[:, 0])                                       # kinks at x = 0 and, through the periodic pair, at x = +-1
    eta = per.kelly_indicator(f.ravel())
    cp = per.cell_centers()                                          # 4 cells per direction: every cell touches a kink ...
    assert eta.min() > 0 and np.allclose(eta, eta[0], rtol=1e-12)    # ... and all see the same jump (2) on one face
    g = np.zeros((per.n_nodes, 18))
    g[:, 0] = np.abs(Xp[:, 2])                                       # z is a wall direction: only the kink at z = 0 counts
    eta = per.kelly_indicator(g.ravel())
    assert eta[np.abs(cp[:, 2]) < 0.5].min() > 0 and np.abs(eta[np.abs(cp[:, 2]) > 0.5]).max() <= 1e-12


@pytest.mark.parametrize("n_ranks", [1, 2, 4, 8])
def test_multigrid_prolongation_tables_on_nested_partitions(n_ranks):
    """The table vh_mg_attach receives: rows of owned fine nodes are complete interpolations whose parents are local on the
    same rank's coarse level, and every coarse owned node coincides with a fine node owned by the same rank."""
    mf = vh.Mesh(1, [-20.0] * 3, [20.0] * 3, n_global_refine=3).finalize(n_ranks)
    mc = vh.Mesh(1, [-20.0] * 3, [20.0] * 3, n_global_refine=2).finalize(n_ranks)
    for r in range(n_ranks):
        Tf, Tc = mf.tables(r), mc.tables(r)
        ptr, idx, w = vh.mg_prolongation(mf, Tf, mc, Tc)
        assert ptr.size == Tf.n_local_nodes + 1 and idx.max() < Tc.n_local_nodes
        rows = np.repeat(np.arange(Tf.n_local_nodes), np.diff(ptr))
        rowsum = np.bincount(rows, weights=w, minlength=Tf.n_local_nodes)
        assert np.abs(rowsum[:Tf.n_owned_nodes] - 1.0).max() <= 1e-14
        one = (np.abs(w - 1.0) <= 1e-14) & (rows < Tf.n_owned_nodes)
        assert np.isin(np.arange(Tc.n_owned_nodes), idx[one]).all()
        # trilinear: a linear function of the coordinates is reproduced at the owned fine nodes
        f = lambda X: 1.0 + 0.5 * X[:, 0] - 0.25 * X[:, 1] + 2.0 * X[:, 2]  # noqa: E731
        got = np.bincount(rows, weights=w * f(Tc.node_xyz)[idx], minlength=Tf.n_local_nodes)
        assert np.abs(got[:Tf.n_owned_nodes] - f(Tf.node_xyz)[:Tf.n_owned_nodes]).max() <= 1e-12
