"""GPU parity AT SIZE: sampled block rows of the assembled Jacobian, the rhs, and the matrix-free operator apply of
BASELINE configs C2 (Q1 r5, 646 866 DoFs) and of a Q2 mesh of the same size (Q2 r4) against the oracle — int32 offsets, row
schedules and the larger-than-L2 regime are exercised with a checker, not only with self-consistency properties.

The oracle cannot assemble 3e8 entries in Python, but a block row only needs the cells incident to its node: the checker
assembles a SUB-MESH made of those cells (oracle cell matrices in C + the constrained scatter of femgl_oracle.py) and the
sampled rows of that sub-assembly are complete rows of the global matrix (uniform meshes: no hanging-node masters)."""
import types

import numpy as np
import pytest

import femgl_oracle as O
import verkko_hem_repo_b200 as vh
from helpers import MATEP_SCC_ON, b_phase_state, blockrow_rel_error, bsr_to_csr, coef_vector

pytestmark = pytest.mark.gpu


def _submesh(T, nodes):
    """Table view holding only the cells that touch the sampled nodes (same node numbering, same constraints)."""
    sel = np.isin(T.cell_nodes, nodes).any(axis=1)
    keep = np.nonzero(sel)[0]
    S = types.SimpleNamespace()
    for k in ("degree", "n_owned_nodes", "n_ghost_nodes", "n_local_nodes", "c_dof", "c_ptr", "c_master", "c_weight", "node_global"):
        setattr(S, k, getattr(T, k))
    S.cell_nodes, S.cell_origin, S.cell_h = T.cell_nodes[keep], T.cell_origin[keep], T.cell_h[keep]
    S.cell_owned = T.cell_owned[keep]
    S.n_cells = keep.size
    remap = -np.ones(T.n_cells, dtype=np.int64)
    remap[keep] = np.arange(keep.size)
    fsel = sel[T.wall_face_cell]
    S.wall_face_cell = remap[T.wall_face_cell[fsel]].astype(np.int32)
    S.wall_face_no, S.wall_face_bid = T.wall_face_no[fsel], T.wall_face_bid[fsel]
    S.face_csr = types.MethodType(vh.RankTables.face_csr, S)
    return S


@pytest.mark.parametrize("degree,refine,n_dofs", [(1, 5, 646866), (2, 4, 646866)])
def test_sampled_rows_rhs_and_apply_match_oracle_at_size(degree, refine, n_dofs):
    m = vh.unit_cube(degree, refine, half=20.0)
    T = m.tables(0)
    assert 18 * m.n_nodes == n_dofs
    coef = coef_vector(MATEP_SCC_ON, 2.0)
    x = b_phase_state(T)
    rng = np.random.default_rng(41)
    # 256 random rows plus the corners / wall nodes most likely to go wrong (first, last, a wall node, an interior node)
    nodes = np.unique(np.concatenate([rng.choice(T.n_owned_nodes, 256, replace=False), [0, T.n_owned_nodes - 1]]))
    S = _submesh(T, nodes)
    A_sub, rhs_sub = O.assemble_global(S, x, coef, True)
    rows = (18 * nodes[:, None] + np.arange(18)[None, :]).ravel()

    ctx = vh.Context(T)
    ctx.set_coef_vector(coef)
    ctx.set_solution(x)
    bn = ctx.assemble()
    assert ctx.info()["n_fast_rows"] == m.n_nodes
    # rhs entries of the sampled rows
    rhs = ctx.get_rhs()
    assert np.abs(rhs[rows] - rhs_sub[rows]).max() <= 1e-12 * np.abs(rhs_sub[rows]).max()
    assert np.isfinite(bn) and abs(bn - np.linalg.norm(rhs)) <= 1e-12 * bn
    # operator apply (the default matrix-free path) on the sampled rows: rows of A_sub are complete, so A_sub[rows] z = (A z)[rows]
    z = rng.uniform(-1, 1, 18 * T.n_local_nodes)
    y = ctx.spmv(z)
    y_ora = A_sub[rows] @ z
    assert np.abs(y[rows] - y_ora).max() <= 1e-13 * np.abs(y_ora).max() * 10
    # the assembled block rows themselves (assembled on demand by the export) — entrywise, block-row relative 1e-12
    A_gpu = bsr_to_csr(*ctx.export_matrix_bsr(), T.n_local_nodes)
    err = blockrow_rel_error(A_gpu[rows], A_sub[rows])
    assert err <= 1e-12, "matrix block-row relative error %.3e" % err
    # and the assembled packed SpMV agrees with the matrix-free apply everywhere
    ctx.set_spmv_matrix_free(0)
    y2 = ctx.spmv(z)
    assert np.abs(y2 - y).max() <= 1e-13 * np.abs(y).max() * 10
    ctx.close()
