"""The product's pointwise kernel family k_points (csrc/vh_points_kernel.cuh: assembly tables + cell rhs, residual,
energy, matrix-free operator apply from the H_q tables, table-free operator apply) compiled by g++ and executed lane by
lane under a small CUDA emulation (tests/native/cuda_emu.h: one fibre per CUDA thread, __syncthreads / __syncwarp /
__shfl_*_sync as cooperative yields, dynamic shared memory per block), checked against the oracle's cell matrices.
This is how kernel LOGIC written without GPU access gets checked in the CPU-only container; parity on hardware is what
tests/test_gpu_*.py assert."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

import femgl_oracle as O
import verkko_hem_repo_b200 as vh
from helpers import MATEP_SCC_ON, b_phase_state, coef_vector

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _lib():
    src = os.path.join(ROOT, "tests", "native", "points_emu_host.cc")
    csrc = os.path.join(ROOT, "verkko-hem-repo_b200", "csrc")
    deps = [src, os.path.join(ROOT, "tests", "native", "cuda_emu.h")] + [os.path.join(csrc, f) for f in
                                                                          ("vh_points_kernel.cuh", "vh_diag_kernel.cuh", "vh_apply_v2.cuh", "vh_gather_kernels.cuh", "vh_block_invert.cuh", "vh_pointwise.cuh", "vh_internal.h")]
    out = os.path.join(ROOT, "tests", "native", "_build", "libvhpoints_emu.so")
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(p) for p in deps):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        subprocess.check_call(["/usr/bin/g++", "-O1", "-std=c++17", "-w", "-shared", "-fPIC", "-I", "/usr/local/cuda/include",
                               "-I", os.path.join(ROOT, "include"), "-o", out, src])
    return ctypes.CDLL(out)


def _p(a, t=ctypes.c_double):
    return a.ctypes.data_as(ctypes.POINTER(t))


def _run(T, mode, x, coef, Hq=None, x_state=None):
    L = _lib()
    n = T.cell_nodes.shape[1]
    nq = n
    N, dN, w, _ = O.fe_tables(T.degree)
    Gref = np.einsum("q,aqx,bqy->abxy", w, dN, dN).copy()
    Mf = np.zeros((6, n, n))
    for f in range(6):
        Nf, wf = O.face_tables(T.degree, f)
        Mf[f] = np.einsum("q,aq,bq->ab", wf, Nf, Nf)
    h4 = np.concatenate([T.cell_h, T.cell_h.prod(axis=1, keepdims=True)], axis=1).copy()
    faces = np.zeros(T.n_cells, dtype=np.uint32)
    for c, f, b in zip(T.wall_face_cell, T.wall_face_no, T.wall_face_bid):
        faces[c] |= np.uint32(int(b) << (4 * int(f)))
    owned = np.ascontiguousarray(T.cell_owned, dtype=np.uint8)
    nodes = np.ascontiguousarray(T.cell_nodes, dtype=np.int32)
    if Hq is None:
        Hq = np.zeros(T.n_cells * nq * 180)
    Rc = np.full((T.n_cells, 18 * n), np.nan)
    Dc = np.zeros((T.n_cells, 18 * n))
    avgD = np.zeros(T.n_cells)
    Ec = np.zeros(T.n_cells)
    xs = x if x_state is None else x_state
    rc = L.vht_points_emulated(T.degree, mode, T.n_cells, _p(nodes, ctypes.c_int32), _p(h4), _p(faces, ctypes.c_uint32),
                               _p(owned, ctypes.c_uint8), _p(x), _p(xs), _p(N), _p(dN), _p(w), _p(Gref), _p(Mf), _p(coef), _p(Hq),
                               _p(Rc), _p(Dc), _p(avgD), _p(Ec))
    assert rc == 0
    return Hq, Rc, Dc, avgD, Ec


def _mesh(kind):
    if kind == "q1-walls":
        return vh.Mesh(1, [-1.0, -2.0, -0.5], [1.5, 1.0, 0.75], base=(2, 3, 1), face_bid=(2, 2, 3, 1, 4, 4), n_global_refine=0).finalize(1).tables(0)
    if kind == "q1-ragged":  # 18 cells: the last warp of the last block is only half full
        return vh.Mesh(1, [-1.0, -1.0, -1.0], [2.0, 2.0, 1.0], base=(3, 3, 2), face_bid=(1, 1, 1, 1, 4, 4), n_global_refine=0).finalize(1).tables(0)
    if kind == "q2-walls":
        return vh.Mesh(2, [-1.0, -2.0, -0.5], [1.5, 1.0, 0.75], base=(2, 1, 1), face_bid=(2, 2, 3, 1, 4, 4), n_global_refine=0).finalize(1).tables(0)
    raise KeyError(kind)


@pytest.mark.parametrize("kind,bt", [("q1-walls", 0.7), ("q1-ragged", 2.0), ("q1-ragged", 1e10), ("q2-walls", 0.7)])
def test_emulated_pointwise_kernels_match_oracle_cells(kind, bt):
    T = _mesh(kind)
    coef = coef_vector(MATEP_SCC_ON, bt)
    x = b_phase_state(T, seed=31)
    fptr, fno, fbid = T.face_csr()
    dummy = np.zeros(1, np.int32)
    K, r, e = O.cells(T.degree, T.cell_nodes, T.cell_origin, T.cell_h, x, coef, fptr, fno if fno.size else dummy,
                      fbid if fbid.size else dummy, want_matrix=True, want_energy=True)
    dofs = (18 * T.cell_nodes.astype(np.int64)[:, :, None] + np.arange(18)[None, None, :]).reshape(T.n_cells, -1)
    scale = np.abs(K).max()
    # assembly mode: cell rhs, cell-matrix diagonal, mean |diag|, and the packed H_q tables (used below)
    Hq, Rc, Dc, avgD, _ = _run(T, 0, x, coef)
    assert np.abs(Rc - r).max() <= 1e-12 * np.abs(r).max()
    diag = np.einsum("eii->ei", K)
    assert np.abs(Dc - diag).max() <= 1e-12 * scale
    assert np.abs(avgD - np.abs(diag).mean(axis=1)).max() <= 1e-12 * scale
    # residual-only mode and the energy
    _, Rc1, _, _, _ = _run(T, 1, x, coef)
    assert np.abs(Rc1 - r).max() <= 1e-12 * np.abs(r).max()
    _, _, _, _, Ec = _run(T, 4, x, coef)
    assert np.abs(Ec - e).max() <= 1e-12 * np.abs(e).max()
    # operator apply: from the tables (mode 2) and table-free (mode 3), for two random directions
    rng = np.random.default_rng(5)
    for _ in range(2):
        z = rng.uniform(-1, 1, x.size)
        want = np.einsum("eij,ej->ei", K, z[dofs])
        _, Y2, _, _, _ = _run(T, 2, z, coef, Hq=Hq)
        assert np.abs(Y2 - want).max() <= 1e-12 * np.abs(want).max(), "apply from the H_q tables"
        _, Y3, _, _, _ = _run(T, 3, z, coef, x_state=x)
        assert np.abs(Y3 - want).max() <= 1e-12 * np.abs(want).max(), "table-free apply"


@pytest.mark.parametrize("kind", ["q1-ragged", "q2-walls"])
def test_emulated_diagonal_block_kernels_match_oracle(kind):
    """k_diag_cells + k_diag_gather (block-Jacobi's diagonal blocks without assembling the lattice rows): the packed bulk
    part P_II of every node's diagonal block equals the oracle's, summed over the node's incident cells."""
    L = _lib()
    T = _mesh(kind)
    n = T.cell_nodes.shape[1]
    coef = coef_vector(MATEP_SCC_ON, 2.0)
    x = b_phase_state(T, seed=33)
    Hq, _, _, _, _ = _run(T, 0, x, coef)          # the tables the assembly kernel wrote (emulated)
    # bulk-only cell matrices from the oracle: gradient coefficients and Robin faces switched off
    bulk = coef.copy()
    bulk[0:3] = 0.0
    bulk[9] = 1e10
    K, _, _ = O.cells(T.degree, T.cell_nodes, T.cell_origin, T.cell_h, x, bulk, want_matrix=True)
    want = np.zeros((T.n_local_nodes, 18, 18))
    for e in range(T.n_cells):
        for a, node in enumerate(T.cell_nodes[e]):
            want[node] += K[e, 18 * a:18 * a + 18, 18 * a:18 * a + 18]
    # row tables as vh_create builds them: incident cells (<= 8) and the local index of the node in each
    fast_rows = np.arange(T.n_local_nodes, dtype=np.int32)
    fast_cells = np.full((T.n_local_nodes, 8), -1, dtype=np.int32)
    fast_a = np.zeros((T.n_local_nodes, 8), dtype=np.int8)
    fill = np.zeros(T.n_local_nodes, dtype=int)
    for e in range(T.n_cells):
        for a, node in enumerate(T.cell_nodes[e]):
            fast_cells[node, fill[node]] = e
            fast_a[node, fill[node]] = a
            fill[node] += 1
    diag_pos = (3 * np.arange(T.n_local_nodes) + 1).astype(np.int32)  # any scattered block positions
    N, _, w, _ = O.fe_tables(T.degree)
    Dblk = np.zeros(T.n_cells * n * 180)
    pvals = np.full((3 * T.n_local_nodes + 2) * 180, np.nan)
    rc = L.vht_diag_emulated(T.degree, T.n_cells, _p(N), _p(w), _p(Hq), T.n_local_nodes, _p(fast_rows, ctypes.c_int32),
                             _p(fast_cells, ctypes.c_int32), _p(fast_a, ctypes.c_int8), _p(diag_pos, ctypes.c_int32), _p(Dblk), _p(pvals))
    assert rc == 0
    native = ctypes.CDLL(os.path.join(ROOT, "tests", "native", "_build", "libvhpw.so")) if os.path.exists(
        os.path.join(ROOT, "tests", "native", "_build", "libvhpw.so")) else None
    if native is None:
        import test_host_logic
        native = test_host_logic._native_pointwise_lib()
    P = pvals.reshape(-1, 180)
    scale = np.abs(want).max()
    for node in range(T.n_local_nodes):
        blk = P[diag_pos[node]]
        for c in range(18):
            for d in range(c, 18):
                v = blk[native.vht_sym_index(c, d)]
                assert abs(v - want[node, c, d]) <= 1e-12 * scale, (node, c, d)
    # untouched blocks stay untouched
    assert np.isnan(P[0]).all() and np.isnan(P[2]).all()


@pytest.mark.parametrize("kind,bt", [("q1-walls", 0.7), ("q1-ragged", 2.0), ("q1-ragged", 1e10)])
def test_emulated_apply_v2_matches_oracle_cells(kind, bt):
    """Second formulation of the Q1 matrix-free apply (bulk part through the H_q tables with lane = point, gradient and
    Robin forms as a per-cell 8x8 table with lane = node): K_cell z_cell of the oracle for every cell."""
    L = _lib()
    T = _mesh(kind)
    coef = coef_vector(MATEP_SCC_ON, bt)
    x = b_phase_state(T, seed=31)
    fptr, fno, fbid = T.face_csr()
    dummy = np.zeros(1, np.int32)
    K, _, _ = O.cells(1, T.cell_nodes, T.cell_origin, T.cell_h, x, coef, fptr, fno if fno.size else dummy,
                      fbid if fbid.size else dummy, want_matrix=True)
    dofs = (18 * T.cell_nodes.astype(np.int64)[:, :, None] + np.arange(18)[None, None, :]).reshape(T.n_cells, -1)
    Hq, _, _, _, _ = _run(T, 0, x, coef)
    N, dN, w, _ = O.fe_tables(1)
    Gref = np.einsum("q,aqx,bqy->abxy", w, dN, dN).copy()
    Mf = np.zeros((6, 8, 8))
    for f in range(6):
        Nf, wf = O.face_tables(1, f)
        Mf[f] = np.einsum("q,aq,bq->ab", wf, Nf, Nf)
    h4 = np.concatenate([T.cell_h, T.cell_h.prod(axis=1, keepdims=True)], axis=1).copy()
    faces = np.zeros(T.n_cells, dtype=np.uint32)
    for c, f, b in zip(T.wall_face_cell, T.wall_face_no, T.wall_face_bid):
        faces[c] |= np.uint32(int(b) << (4 * int(f)))
    nodes = np.ascontiguousarray(T.cell_nodes, dtype=np.int32)
    rng = np.random.default_rng(6)
    for _ in range(2):
        z = rng.uniform(-1, 1, x.size)
        want = np.einsum("eij,ej->ei", K, z[dofs])
        Yc = np.full((T.n_cells, 144), np.nan)
        rc = L.vht_apply_v2_emulated(T.n_cells, _p(nodes, ctypes.c_int32), _p(h4), _p(faces, ctypes.c_uint32), _p(z), _p(N), _p(dN), _p(w),
                                     _p(Gref), _p(Mf), _p(coef), _p(Hq), _p(Yc))
        assert rc == 0
        assert np.abs(Yc - want).max() <= 1e-12 * np.abs(want).max()


@pytest.mark.parametrize("kind,bt,apply_mode", [("q1-walls", 0.7, 2), ("q1-walls", 0.7, 3), ("q1-ragged", 2.0, 2), ("q2-walls", 0.7, 2),
                                                 ("q2-walls", 0.7, 3)])
def test_emulated_lattice_row_pipeline_matches_oracle_global(kind, bt, apply_mode):
    """End to end on the CPU, for meshes whose rows are all lattice rows: pointwise assembly kernel -> k_rhs_fast (system_rhs
    and the constrained-diagonal values of distribute_local_to_global) -> matrix-free operator apply (k_points<APPLY> from the
    tables or table-free, then k_gather_apply with the Dirichlet rule) against the oracle's GLOBAL rhs and matrix."""
    L = _lib()
    T = _mesh(kind)
    n = T.cell_nodes.shape[1]
    dpc = 18 * n
    coef = coef_vector(MATEP_SCC_ON, bt)
    x = b_phase_state(T, seed=35)
    A, rhs_ora = O.assemble_global(T, x, coef, True)
    Hq, Rc, Dc, avgD, _ = _run(T, 0, x, coef)
    nn = T.n_local_nodes
    fast_rows = np.arange(nn, dtype=np.int32)
    fast_cells = np.full((nn, 8), -1, dtype=np.int32)
    fast_a = np.zeros((nn, 8), dtype=np.int8)
    fill = np.zeros(nn, dtype=int)
    for e in range(T.n_cells):
        for a, node in enumerate(T.cell_nodes[e]):
            fast_cells[node, fill[node]] = e
            fast_a[node, fill[node]] = a
            fill[node] += 1
    dirmask = np.zeros(nn, dtype=np.uint32)
    assert T.c_master.size == 0                      # only masked Dirichlet lines on these meshes
    for dof in T.c_dof:
        dirmask[dof // 18] |= np.uint32(1 << int(dof % 18))
    rhs = np.full(18 * nn, np.nan)
    cdiag = np.full(18 * nn, np.nan)
    dummy = np.zeros(1)
    I32, I8, U32 = ctypes.c_int32, ctypes.c_int8, ctypes.c_uint32
    rc = L.vht_gather_emulated(0, nn, dpc, _p(fast_rows, I32), _p(fast_cells, I32), _p(fast_a, I8), _p(dirmask, U32), _p(Rc), _p(Dc),
                               _p(avgD), _p(cdiag), _p(dummy), _p(rhs))
    assert rc == 0
    assert np.abs(rhs - rhs_ora).max() <= 1e-12 * np.abs(rhs_ora).max()
    rng = np.random.default_rng(8)
    z = rng.uniform(-1, 1, 18 * nn)
    con = np.zeros(18 * nn, dtype=bool)
    con[T.c_dof] = True
    zm = np.where(con, 0.0, z)                         # what k_mask_dirichlet hands to the cell kernel
    if apply_mode == 2:
        _, Yc, _, _, _ = _run(T, 2, zm, coef, Hq=Hq)
    else:
        _, Yc, _, _, _ = _run(T, 3, zm, coef, x_state=x)
    y = np.full(18 * nn, np.nan)
    rc = L.vht_gather_emulated(1, nn, dpc, _p(fast_rows, I32), _p(fast_cells, I32), _p(fast_a, I8), _p(dirmask, U32), _p(Yc), _p(Dc),
                               _p(avgD), _p(cdiag), _p(z), _p(y))
    assert rc == 0
    y_ora = A @ z
    assert np.abs(y - y_ora).max() <= 1e-12 * np.abs(y_ora).max()


def _sym_index(c, d):
    """vh_sym_index of csrc/vh_pointwise.cuh: position of (c, d), c <= d, in the packed P180 layout"""
    m = c >> 1
    start = (36 * m - 2 * m * m + 18) if (c & 1) else (38 * m - 2 * m * m)
    return start + d - 2 * m


@pytest.mark.parametrize("packed", [0, 1])
def test_emulated_block_invert(packed):
    """k_block_invert (Gauss-Jordan with deferred scaling, pivoting by redux/ballot, staged through shared memory): A * minv = I
    for full blocks and for packed lattice blocks Sym(P) + kron(I_6, M) with Dirichlet-masked rows; a singular block is counted."""
    L = _lib()
    rng = np.random.default_rng(11 + packed)
    n = 13                                     # not a multiple of the 4 blocks per CTA
    M9 = rng.uniform(-1, 1, (3, 3))
    M9 = M9 + M9.T
    dirmask = np.zeros(n, dtype=np.uint32)
    cdiag = rng.uniform(1.0, 2.0, (n, 18))
    A = np.zeros((n, 18, 18))
    if packed:
        blocks = np.zeros((n, 180))
        dirmask[3] = (1 << 2) | (1 << 5) | (1 << 17)
        dirmask[7] = 0x3FFFF
        for i in range(n):
            S = rng.uniform(-1, 1, (18, 18))
            S = S + S.T + (0.0 if i % 3 else 6.0) * np.eye(18)   # indefinite and definite blocks
            for c in range(18):
                for d in range(c, 18):
                    blocks[i, _sym_index(c, d)] = S[c, d]
            A[i] = S + np.kron(np.eye(6), M9)
            for c in range(18):
                if (dirmask[i] >> c) & 1:
                    A[i, c, :] = 0.0
                    A[i, :, c] = 0.0
                    A[i, c, c] = cdiag[i, c]
    else:
        blocks = rng.uniform(-1, 1, (n, 18, 18))
        blocks[5] = np.eye(18)[rng.permutation(18)] * rng.uniform(0.5, 2.0, 18)   # needs pivoting in every step
        blocks[9][:, 4] = 0.0                                                     # singular
        A[:] = blocks
    minv = np.full((n, 18, 18), np.nan)
    nsing = ctypes.c_int(0)
    rc = L.vht_block_invert_emulated(n, packed, _p(np.ascontiguousarray(blocks)), _p(np.ascontiguousarray(M9)), _p(dirmask, ctypes.c_uint32),
                                     _p(cdiag), _p(minv), ctypes.byref(nsing))
    assert rc == 0
    assert nsing.value == (0 if packed else 1)
    for i in range(n):
        if not packed and i == 9:
            assert np.array_equal(minv[i], np.eye(18))
            continue
        assert np.abs(A[i] @ minv[i] - np.eye(18)).max() <= 1e-11 * np.linalg.cond(A[i]), i
