"""CPU tests of the periodic box (the reference's ACTIVE grid/setup pair, femgl/CMakeLists.txt:51,61):
makegrid_retangle-z-AdGR_xy-periodic.cc:167-219 tags x faces 5/6 and y faces 7/8 periodic, z faces id 4;
setup_weak-coupling-PDW-configuration.cc:128-206 turns the pairs into identity constraints
(DoFTools::make_periodicity_constraints) next to the hanging-node and masked-Dirichlet lines.

The tables are checked structurally, and the discrete problem built from them (oracle constrained scatter) is checked
through a property only a correct periodic identification has: invariance under a cyclic shift of the field by one
cell along a periodic direction."""
import numpy as np
import pytest

import femgl_oracle as O
import verkko_hem_repo_b200 as vh
from helpers import b_phase_state, coef_vector


@pytest.mark.parametrize("degree,refine", [(1, 2), (2, 1)])
def test_periodic_tables_structure(degree, refine):
    m = vh.periodic_slab(degree, refine, half=(1.0, 1.5, 0.5))
    T = m.tables(0)
    n1 = degree * 2 ** refine + 1
    # upper x face + upper y face - their common edge
    assert m.n_periodic_nodes == (2 * n1 - 1) * n1 and m.n_hanging_nodes == 0
    xyz = T.node_xyz
    con = np.zeros(18 * T.n_local_nodes, dtype=bool)
    con[T.c_dof] = True
    on_upper = (np.abs(xyz[:, 0] - 1.0) < 1e-13) | (np.abs(xyz[:, 1] - 1.5) < 1e-13)
    on_wall = np.abs(np.abs(xyz[:, 2]) - 0.5) < 1e-13
    # every DoF of an upper-face node is constrained; elsewhere only the 6 masked wall components (femgl.h:300-301)
    assert con.reshape(-1, 18)[on_upper].all()
    rest = con.reshape(-1, 18)[~on_upper]
    assert not rest[:, [0, 1, 3, 4, 6, 7, 9, 10, 12, 13, 15, 16]].any()
    assert (rest[:, [2, 5, 8, 11, 14, 17]].all(axis=1) == on_wall[~on_upper]).all()
    # identity lines: one master, weight 1, same component, image = same point with the periodic coordinates folded back;
    # masked wall components of periodic nodes close to "no master" (value 0)
    cnt = np.diff(T.c_ptr)
    for k, dof in enumerate(T.c_dof):
        nd, c = divmod(int(dof), 18)
        if not on_upper[nd]:
            assert cnt[k] == 0
            continue
        if on_wall[nd] and c % 3 == 2:
            assert cnt[k] == 0
            continue
        assert cnt[k] == 1 and T.c_weight[T.c_ptr[k]] == 1.0
        mdof = int(T.c_master[T.c_ptr[k]])
        assert mdof % 18 == c and not con[mdof]
        want = xyz[nd].copy()
        if abs(want[0] - 1.0) < 1e-13:
            want[0] = -1.0
        if abs(want[1] - 1.5) < 1e-13:
            want[1] = -1.5
        assert np.abs(xyz[mdof // 18] - want).max() < 1e-13
    # periodic faces are not walls: only the z faces feed the Robin term
    assert set(T.wall_face_bid.tolist()) == {4}


def test_bad_periodic_ids_are_rejected():
    with pytest.raises(RuntimeError):
        vh.Mesh(1, [-1] * 3, [1] * 3, face_bid=(5, 1, 1, 1, 4, 4), n_global_refine=2).finalize(1)
    with pytest.raises(RuntimeError):
        vh.Mesh(1, [-1] * 3, [1] * 3, face_bid=(7, 8, 1, 1, 4, 4), n_global_refine=2).finalize(1)
    with pytest.raises(RuntimeError):  # one cell across: a cell would hold a node and its own image
        vh.Mesh(1, [-1] * 3, [1] * 3, face_bid=(5, 6, 1, 1, 4, 4), n_global_refine=0).finalize(1)


def _lagrange(degree, t, xi):
    if degree == 1:
        return xi if t else 1.0 - xi
    return [(2 * xi - 1) * (xi - 1), 4 * xi * (1 - xi), xi * (2 * xi - 1)][t]


def _fe_value(T, x, cell, point):
    """FE function of the local vector x in cell `cell` at the physical point `point` (18 components)."""
    _, _, _, support = O.fe_tables(T.degree)
    xi = (np.asarray(point) - T.cell_origin[cell]) / T.cell_h[cell]
    assert (xi > -1e-12).all() and (xi < 1 + 1e-12).all()
    val = np.zeros(18)
    for a, node in enumerate(T.cell_nodes[cell]):
        t = np.rint(support[a] * T.degree).astype(int)
        w = np.prod([_lagrange(T.degree, t[d], xi[d]) for d in range(3)])
        val += w * x.reshape(-1, 18)[node]
    return val


@pytest.mark.parametrize("degree,refine", [(1, 2), (2, 1)])
def test_differently_refined_periodic_faces_stay_conforming(degree, refine):
    """Local refinement that reaches one periodic face only: Mesh::refine balances 2:1 across the seam and the nodes of
    the finer face without a counterpart hang on the coarse cell across the seam (what make_periodicity_constraints does
    for refined faces).  After distribute() the FE function must be single-valued across both periodic pairs, at matching
    and non-matching places alike, and the assembled system must still be the gradient of the energy."""
    half = np.array([1.0, 1.5, 0.75])
    m = vh.Mesh(degree, -half, half, face_bid=(5, 6, 7, 8, 4, 4), n_global_refine=refine)
    c = m.cell_centers()
    fl = (c[:, 0] < -half[0] + 0.6) & (c[:, 1] < -half[1] + 0.8) & (c[:, 2] < 0)  # touches the lower x and lower y faces only
    m.refine(fl)
    m.finalize(1)
    assert m.n_hanging_nodes > 0 and m.n_periodic_nodes > 0
    T = m.tables(0)
    rng = np.random.default_rng(2)
    x = O.distribute(T, rng.uniform(-1, 1, 18 * T.n_local_nodes))
    assert np.abs(O.distribute(T, x) - x).max() == 0.0
    lo, hi = T.cell_origin, T.cell_origin + T.cell_h
    n_nonmatching = 0
    for d in (0, 1):
        lower = np.nonzero(np.abs(lo[:, d] + half[d]) < 1e-12)[0]
        upper = np.nonzero(np.abs(hi[:, d] - half[d]) < 1e-12)[0]
        for _ in range(200):
            p = rng.uniform(-half, half)
            p_lo, p_hi = p.copy(), p.copy()
            p_lo[d], p_hi[d] = -half[d], half[d]
            k_lo = next(k for k in lower if (lo[k] <= p_lo + 1e-12).all() and (hi[k] >= p_lo - 1e-12).all())
            k_hi = next(k for k in upper if (lo[k] <= p_hi + 1e-12).all() and (hi[k] >= p_hi - 1e-12).all())
            n_nonmatching += abs(T.cell_h[k_lo, 0] - T.cell_h[k_hi, 0]) > 1e-12
            v_lo, v_hi = _fe_value(T, x, k_lo, p_lo), _fe_value(T, x, k_hi, p_hi)
            assert np.abs(v_lo - v_hi).max() <= 1e-12, (d, p)
    assert n_nonmatching > 10  # the samples really crossed differently refined places
    # the constrained system is still the gradient of the energy: -2 rhs = dE/dx along free directions
    coef = coef_vector(bt=2.0)
    xs = b_phase_state(T, seed=4)
    _, rhs = O.assemble_global(T, xs, coef, False)
    con = np.zeros(18 * T.n_local_nodes, dtype=bool)
    con[T.c_dof] = True
    free = np.nonzero(~con)[0]
    for i in rng.choice(free, size=6, replace=False):
        eps = 1e-5
        e = []
        for sgn in (1, -1):
            xp = xs.copy()
            xp[i] += sgn * eps
            e.append(O.energy_global(T, O.distribute(T, xp), coef))
        fd = (e[0] - e[1]) / (2 * eps)
        assert abs(fd + 2.0 * rhs[i]) <= 1e-6 * max(1.0, abs(fd)), (i, fd, rhs[i])


def _shift_x(T, x, n_cells_x, hx, x_lo, x_hi):
    """Field y(p) = x(p - hx e_x) with periodic wrap, evaluated on the nodes of T (unconstrained nodes first, then distribute)."""
    key = {tuple(np.round(p, 9)): i for i, p in enumerate(T.node_xyz)}
    L = x_hi - x_lo
    src = np.empty(T.n_local_nodes, dtype=np.int64)
    for i, p in enumerate(T.node_xyz):
        q = p.copy()
        q[0] = x_lo + ((q[0] - hx - x_lo) % L)
        if abs(q[0] - x_hi) < 1e-9:
            q[0] = x_lo
        src[i] = key[tuple(np.round(q, 9))]
    return x.reshape(-1, 18)[src].ravel(), src


@pytest.mark.parametrize("degree,refine", [(1, 2), (2, 1)])
def test_periodic_problem_is_translation_invariant(degree, refine):
    """Energy, |rhs| and the assembled operator are invariant under a one-cell cyclic shift along x — true only if the
    upper face is identified with the lower one in the scatter (matrix rows/columns AND rhs)."""
    half = (1.0, 1.5, 0.5)
    m = vh.periodic_slab(degree, refine, half=half)
    T = m.tables(0)
    coef = coef_vector(bt=2.0)
    x = b_phase_state(T, seed=3)  # distribute() makes it periodic
    hx = 2.0 * half[0] / 2 ** refine
    y, src = _shift_x(T, x, 2 ** refine, hx, -half[0], half[0])
    y = O.distribute(T, y)
    assert np.abs(y - O.distribute(T, y)).max() == 0.0
    e0, e1 = O.energy_global(T, x, coef), O.energy_global(T, y, coef)
    assert abs(e0 - e1) <= 1e-12 * abs(e0)
    A0, r0 = O.assemble_global(T, x, coef, True)
    A1, r1 = O.assemble_global(T, y, coef, True)
    assert abs(np.linalg.norm(r0) - np.linalg.norm(r1)) <= 1e-12 * np.linalg.norm(r0)
    # rhs and operator are the shifted ones on the unconstrained DoFs
    dsrc = (18 * src[:, None] + np.arange(18)[None, :]).ravel()
    con = np.zeros(18 * T.n_local_nodes, dtype=bool)
    con[T.c_dof] = True
    free = ~con & ~con[dsrc]
    assert np.abs(r1[free] - r0[dsrc][free]).max() <= 1e-12 * np.abs(r0).max()
    A0s = A0.tocsr()[dsrc][:, dsrc]
    D = abs(A1 - A0s).tocsr()[free][:, free]
    assert D.max() <= 1e-12 * abs(A0).max()
    # the masters' rows really collect the far side: a lower-face node couples to nodes next to the upper face
    xyz = T.node_xyz
    lower = np.nonzero((np.abs(xyz[:, 0] + half[0]) < 1e-13) & (np.abs(xyz[:, 2]) < 0.2) & (np.abs(xyz[:, 1]) < 0.2))[0]
    assert lower.size
    cols = A0.tocsr()[18 * lower[0]].indices // 18
    assert (xyz[cols, 0] > half[0] - hx - 1e-9).any() and (xyz[cols, 0] < -half[0] + hx + 1e-9).any()


def test_periodic_partition_independence():
    """Rows assembled per rank (owned + ghost-layer cells, constraint masters across the periodic seam included) equal the
    1-rank rows: the multi-GPU path needs no compress(add) exchange on the periodic grid either."""
    coef = coef_vector(bt=2.0)
    half = (1.0, 1.0, 1.0)
    T1 = vh.periodic_slab(1, 2, half=half).tables(0)
    mP = vh.periodic_slab(1, 2, half=half, n_ranks=3)
    key1 = {tuple(np.round(p, 9)): i for i, p in enumerate(T1.node_xyz)}
    x1 = b_phase_state(T1, seed=5)
    A1, r1 = O.assemble_global(T1, x1, coef, True)
    A1 = A1.tocsr()
    n_owned = 0
    for r in range(3):
        T = mP.tables(r)
        perm = np.array([key1[tuple(np.round(p, 9))] for p in T.node_xyz])
        x = x1.reshape(-1, 18)[perm].ravel()
        assert np.abs(O.distribute(T, x) - x).max() == 0.0  # every visible constrained DoF finds its masters locally
        A, rhs = O.assemble_global(T, x, coef, True)
        dof_perm = (18 * perm[:, None] + np.arange(18)[None, :]).ravel()
        want = A1[dof_perm[:18 * T.n_owned_nodes]][:, dof_perm]
        assert abs(A - want).max() <= 1e-13 * abs(A1).max()
        assert np.abs(rhs - r1[dof_perm[:18 * T.n_owned_nodes]]).max() <= 1e-13 * np.abs(r1).max()
        n_owned += T.n_owned_nodes
    assert n_owned == T1.n_owned_nodes


@pytest.mark.parametrize("degree", [1, 2])
def test_hanging_node_partition_independence(degree):
    """C4-shaped tables (locally refined slab, hanging nodes) split over 3 ranks: every rank's owned rows, assembled from its
    owned + ghost-layer cells with the constraint lines it sees, equal the 1-rank rows — also for the rows of masters whose
    hanging nodes live in another rank's cells."""
    coef = coef_vector(bt=2.0)

    def make(n_ranks):
        m = vh.Mesh(degree, [-2, -2, -2], [2, 2, 2], n_global_refine=2 if degree == 1 else 1)
        c = m.cell_centers()
        m.refine((np.abs(c[:, 2]) < 1.1) & (c[:, 0] < 0.1))
        return m.finalize(n_ranks)

    m1 = make(1)
    assert m1.n_hanging_nodes > 0
    T1 = m1.tables(0)
    mP = make(3)
    key1 = {tuple(np.round(p, 9)): i for i, p in enumerate(T1.node_xyz)}
    x1 = b_phase_state(T1, seed=5)
    A1, r1 = O.assemble_global(T1, x1, coef, True)
    A1 = A1.tocsr()
    n_owned = 0
    for r in range(3):
        T = mP.tables(r)
        perm = np.array([key1[tuple(np.round(p, 9))] for p in T.node_xyz])
        x = x1.reshape(-1, 18)[perm].ravel()
        A, rhs = O.assemble_global(T, x, coef, True)
        dof_perm = (18 * perm[:, None] + np.arange(18)[None, :]).ravel()
        want = A1[dof_perm[:18 * T.n_owned_nodes]][:, dof_perm]
        assert abs(A - want).max() <= 1e-13 * abs(A1).max()
        assert np.abs(rhs - r1[dof_perm[:18 * T.n_owned_nodes]]).max() <= 1e-13 * np.abs(r1).max()
        n_owned += T.n_owned_nodes
    assert n_owned == T1.n_owned_nodes


def test_differently_refined_periodic_faces_partition_independence():
    """The same mesh type split over 3 ranks: owned rows from owned + ghost-layer cells equal the 1-rank rows."""
    coef = coef_vector(bt=2.0)
    half = np.array([1.0, 1.5, 0.75])

    def make(n_ranks):
        m = vh.Mesh(1, -half, half, face_bid=(5, 6, 7, 8, 4, 4), n_global_refine=2)
        c = m.cell_centers()
        m.refine((c[:, 0] < -half[0] + 0.6) & (c[:, 1] < -half[1] + 0.8) & (c[:, 2] < 0))
        return m.finalize(n_ranks)

    T1 = make(1).tables(0)
    mP = make(3)
    key1 = {tuple(np.round(p, 9)): i for i, p in enumerate(T1.node_xyz)}
    x1 = b_phase_state(T1, seed=5)
    A1, r1 = O.assemble_global(T1, x1, coef, True)
    A1 = A1.tocsr()
    for r in range(3):
        T = mP.tables(r)
        assert vh.validate_tables(T) == ""
        perm = np.array([key1[tuple(np.round(p, 9))] for p in T.node_xyz])
        x = x1.reshape(-1, 18)[perm].ravel()
        A, rhs = O.assemble_global(T, x, coef, True)
        dof_perm = (18 * perm[:, None] + np.arange(18)[None, :]).ravel()
        want = A1[dof_perm[:18 * T.n_owned_nodes]][:, dof_perm]
        assert abs(A - want).max() <= 1e-13 * abs(A1).max()
        assert np.abs(rhs - r1[dof_perm[:18 * T.n_owned_nodes]]).max() <= 1e-13 * np.abs(r1).max()
