"""Randomised multi-level refinement of the mini host's box mesh (what four adaptive cycles of run.cc:182-195 produce):
2:1 balance, hanging-node constraint lines (chains resolved), partition tables.  Properties checked on every mesh:
 * polynomial reproduction: the nodal interpolant of a global tri-linear (Q1) / tri-quadratic (Q2) polynomial satisfies
   every constraint line exactly (AffineConstraints::distribute leaves it unchanged);
 * conformity: after distribute() of a RANDOM nodal field the FE function is single-valued across random points of
   random cell faces (coarse/fine interfaces included);
 * every rank's tables pass vh_validate_mesh_desc and the ranks' owned nodes tile the mesh."""
import numpy as np
import pytest

import femgl_oracle as O
import verkko_hem_repo_b200 as vh
from helpers import distribute_constraints


def _random_mesh(degree, seed, rounds, face_bid=(1, 1, 1, 1, 1, 1), frac=0.15):
    rng = np.random.default_rng(seed)
    m = vh.Mesh(degree, [-1.0, -0.5, 0.0], [1.0, 1.0, 1.5], base=(2, 1, 1), face_bid=face_bid, n_global_refine=1)
    for _ in range(rounds):
        m.refine(rng.uniform(size=m.n_cells) < frac)
    return m


def _lagrange(degree, t, xi):
    if degree == 1:
        return xi if t else 1.0 - xi
    return [(2 * xi - 1) * (xi - 1), 4 * xi * (1 - xi), xi * (2 * xi - 1)][t]


def _fe_value(T, support_t, x, cell, point):
    xi = (point - T.cell_origin[cell]) / T.cell_h[cell]
    val = np.zeros(18)
    for a, node in enumerate(T.cell_nodes[cell]):
        w = np.prod([_lagrange(T.degree, support_t[a][d], xi[d]) for d in range(3)])
        if w != 0.0:
            val += w * x.reshape(-1, 18)[node]
    return val


@pytest.mark.parametrize("degree,seed,rounds", [(1, 1, 3), (1, 2, 4), (1, 3, 3), (2, 4, 2), (2, 5, 3)])
def test_random_refinement_constraints_are_exact_and_conforming(degree, seed, rounds):
    m = _random_mesh(degree, seed, rounds).finalize(1)
    T = m.tables(0)
    assert m.n_hanging_nodes > 0
    levels = np.unique(np.round(T.cell_h[:, 0], 12))
    assert levels.size >= 2
    # closed lines: no master is itself constrained
    con = np.zeros(18 * T.n_local_nodes, dtype=bool)
    con[T.c_dof] = True
    assert not con[T.c_master].any()
    # weights of every hanging line sum to one (constants are reproduced)
    cnt = np.diff(T.c_ptr)
    sums = np.add.reduceat(np.append(T.c_weight, 0.0), T.c_ptr[:-1])[cnt > 0] if T.c_master.size else np.zeros(0)
    assert np.abs(sums - 1.0).max() < 1e-13
    # polynomial reproduction
    X = T.node_xyz
    p = degree
    poly = sum(((i + 1.0) - 0.3 * j + 0.7 * k) * X[:, 0] ** i * X[:, 1] ** j * X[:, 2] ** k
               for i in range(p + 1) for j in range(p + 1) for k in range(p + 1))
    xp = np.repeat(poly[:, None], 18, axis=1) * (1.0 + 0.1 * np.arange(18))[None, :]
    assert np.abs(distribute_constraints(T, xp.ravel()) - xp.ravel()).max() <= 1e-12 * np.abs(xp).max()
    # conformity of a random field
    rng = np.random.default_rng(100 + seed)
    x = distribute_constraints(T, rng.uniform(-1, 1, 18 * T.n_local_nodes))
    _, _, _, support = O.fe_tables(degree)
    support_t = np.rint(support * degree).astype(int)
    lo, hi = T.cell_origin, T.cell_origin + T.cell_h
    n_checked = n_interfaces = 0
    for _ in range(300):
        k = int(rng.integers(T.n_cells))
        d = int(rng.integers(3))
        side = int(rng.integers(2))
        pt = lo[k] + rng.uniform(0.02, 0.98, 3) * T.cell_h[k]
        pt[d] = hi[k, d] if side else lo[k, d]
        inside = ((lo <= pt + 1e-12) & (hi >= pt - 1e-12)).all(axis=1)
        inside[k] = False
        others = np.nonzero(inside)[0]
        if others.size == 0:
            continue  # boundary face
        v0 = _fe_value(T, support_t, x, k, pt)
        for o in others:
            assert np.abs(_fe_value(T, support_t, x, int(o), pt) - v0).max() <= 1e-12, (k, o, pt)
            n_interfaces += abs(T.cell_h[o, 0] - T.cell_h[k, 0]) > 1e-12
        n_checked += 1
    assert n_checked > 100 and n_interfaces > 5


@pytest.mark.parametrize("degree,seed,rounds,n_ranks", [(1, 7, 3, 3), (1, 8, 4, 8), (2, 9, 2, 4)])
def test_random_refinement_partition_tables(degree, seed, rounds, n_ranks):
    m = _random_mesh(degree, seed, rounds, face_bid=(1, 1, 2, 1, 4, 4)).finalize(n_ranks)
    owned = np.zeros(m.n_nodes, dtype=int)
    tabs = [m.tables(r) for r in range(n_ranks)]
    for r, T in enumerate(tabs):
        assert vh.validate_tables(T) == "", r
        owned[T.node_global[:T.n_owned_nodes]] += 1
        # every constrained DoF of a visited cell finds its masters locally: distribute is closed on the rank
        x = np.random.default_rng(r).uniform(-1, 1, 18 * T.n_local_nodes)
        y = distribute_constraints(T, x)
        assert np.abs(distribute_constraints(T, y) - y).max() == 0.0
    assert (owned == 1).all()
    for r, T in enumerate(tabs):  # halo plans agree pairwise
        for k, p in enumerate(T.peer_rank):
            send = T.node_global[T.send_nodes[T.send_ptr[k]:T.send_ptr[k + 1]]]
            Tp = tabs[p]
            kk = list(Tp.peer_rank).index(r)
            assert np.array_equal(send, Tp.node_global[Tp.recv_nodes[Tp.recv_ptr[kk]:Tp.recv_ptr[kk + 1]]])


def test_random_refinement_rows_are_partition_independent():
    """Oracle rows of every rank (owned + ghost-layer cells, multi-level hanging nodes, masked walls) equal the 1-rank rows."""
    from helpers import b_phase_state, coef_vector
    coef = coef_vector(bt=2.0)
    bid = (1, 1, 2, 1, 4, 4)
    T1 = _random_mesh(1, 11, 3, face_bid=bid).finalize(1).tables(0)
    mP = _random_mesh(1, 11, 3, face_bid=bid).finalize(4)
    key1 = {tuple(np.round(p, 9)): i for i, p in enumerate(T1.node_xyz)}
    x1 = b_phase_state(T1, seed=5)
    A1, r1 = O.assemble_global(T1, x1, coef, True)
    A1 = A1.tocsr()
    for r in range(4):
        T = mP.tables(r)
        perm = np.array([key1[tuple(np.round(p, 9))] for p in T.node_xyz])
        x = x1.reshape(-1, 18)[perm].ravel()
        A, rhs = O.assemble_global(T, x, coef, True)
        dof_perm = (18 * perm[:, None] + np.arange(18)[None, :]).ravel()
        want = A1[dof_perm[:18 * T.n_owned_nodes]][:, dof_perm]
        assert abs(A - want).max() <= 1e-13 * abs(A1).max()
        assert np.abs(rhs - r1[dof_perm[:18 * T.n_owned_nodes]]).max() <= 1e-13 * np.abs(r1).max()


@pytest.mark.parametrize("world", [2, 3, 4, 8])
def test_multigrid_levels_are_nested_on_every_partition(world):
    """What vh_mg_attach requires of two consecutive levels (csrc/vh_multigrid.cu), checked on the host tables for the partitions
    bench.py runs (1/2/4/8 ranks) and an odd one: the prolongation rows of the owned fine nodes are complete (weights sum to 1,
    every parent is local on the coarse level) and every coarse owned node coincides with a fine node owned by the same rank."""
    import verkko_hem_repo_b200 as vh
    meshes = [vh.unit_cube(1, lv, half=20.0, n_ranks=world) for lv in (4, 3, 2)]
    for rank in range(world):
        tabs = [m.tables(rank) for m in meshes]
        for k in range(2):
            Tf, Tc = tabs[k], tabs[k + 1]
            ptr, cn, w = vh.mg_prolongation(meshes[k], Tf, meshes[k + 1], Tc)
            assert ptr.size - 1 == Tf.n_local_nodes and (cn.size == 0 or (cn.min() >= 0 and cn.max() < Tc.n_local_nodes))
            rows = np.repeat(np.arange(Tf.n_local_nodes), np.diff(ptr))
            total = np.bincount(rows, weights=w, minlength=Tf.n_local_nodes)[:Tf.n_owned_nodes]
            assert total.size == 0 or np.abs(total - 1.0).max() <= 1e-12
            one = (np.abs(w - 1.0) <= 1e-12) & (rows < Tf.n_owned_nodes) & (cn < Tc.n_owned_nodes)
            has = np.zeros(Tc.n_owned_nodes, dtype=bool)
            has[cn[one]] = True
            assert has.all()
