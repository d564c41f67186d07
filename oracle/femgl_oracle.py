"""TEST INFRASTRUCTURE — the parity checker, not the product.

Python face of the CPU oracle for VerHem's femgl Newton hot path.  Only tests/,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs may
import this module.  The CUDA library never does.

* ``ref``    : oracle/_ref/libvhref.so — the reference's OWN ``cell_mat_vec/*.cc`` + ``matep.cc``
               compiled verbatim (oracle/Makefile ``ref``), driven by literal transcriptions of
               the loops of assemble.cc:177-361 / residual.cc:166-289 ("O1").
* ``lib``    : oracle/_build/libvhoracle.so — the C restatement femgl_oracle.c ("O2", cell level).
* the rest of this file restates, in numpy/scipy, what deal.II does around those cells:
  constrained scatter (``AffineConstraints::distribute_local_to_global``, call sites
  assemble.cc:356-361, residual.cc:287-289), the FGMRES driver (solve.cc:156-183) with the
  north star's nodal 18x18 block-Jacobi preconditioner, the line search of iteration.cc:128-210
  and the stop logic of run.cc:207-256.  [deal.II-internal] semantics are those of SURVEY.md
  Appendix A.4/A.5; they are unpinned by any reference test (the reference has none).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_dp = ctypes.POINTER(ctypes.c_double)
_ip = ctypes.POINTER(ctypes.c_int)


def _P(a):
    return None if a is None else a.ctypes.data_as(_dp)


def _I(a):
    return None if a is None else a.ctypes.data_as(_ip)


def build(ref=True):
    """Compile the checkers (oracle C restatement; reference objects when /root/reference exists)."""
    subprocess.check_call(["make", "-s", "-C", _HERE, "oracle"])
    if ref:
        subprocess.check_call(["make", "-s", "-C", _HERE, "ref"])


_lib = None
_ref = None


def lib():
    global _lib
    if _lib is None:
        p = os.path.join(_HERE, "_build", "libvhoracle.so")
        if not os.path.exists(p):
            build(ref=False)
        L = ctypes.CDLL(p)
        L.vho_pointwise.argtypes = [_dp, _dp, _dp, _dp, _dp]
        L.vho_fe_tables.argtypes = [ctypes.c_int, _dp, _dp, _dp, _dp]
        L.vho_face_tables.argtypes = [ctypes.c_int, ctypes.c_int, _dp, _dp]
        L.vho_cell.argtypes = [ctypes.c_int, _dp, _dp, _dp, _dp, ctypes.c_int, _ip, _ip, _dp, _dp, _dp]
        L.vho_cells.argtypes = [ctypes.c_int, ctypes.c_int, _ip, _dp, _dp, _dp, _dp, _ip, _ip, _ip, _dp, _dp, _dp]
        _lib = L
    return _lib


def have_ref():
    return os.path.exists(os.path.join(_HERE, "_ref", "libvhref.so"))


def ref():
    """The reference's own objects (O1).  Present when built here or shipped prebuilt in oracle/_ref."""
    global _ref
    if _ref is None:
        p = os.path.join(_HERE, "_ref", "libvhref.so")
        if not os.path.exists(p):
            build(ref=True)
        L = ctypes.CDLL(p)
        L.vhref_matep.argtypes = [ctypes.c_double, ctypes.c_double, ctypes.c_int, _dp]
        L.vhref_rhs_terms.argtypes = [_dp] * 5
        L.vhref_lhs_terms.argtypes = [_dp] * 7
        L.vhref_lhs_grad_terms.argtypes = [_dp] * 5
        L.vhref_cell.argtypes = [ctypes.c_int, ctypes.c_int, _dp, _dp, _dp, _dp, _dp, ctypes.c_int, _ip, ctypes.c_int,
                                 _dp, _dp, ctypes.c_int, _dp, _dp]
        _ref = L
    return _ref


# ------------------------------------------------------------------------------------------
# cell level
# ------------------------------------------------------------------------------------------
def nodes_per_cell(degree):
    return 8 if degree == 1 else 27


def fe_tables(degree):
    n = nodes_per_cell(degree)
    nq = n
    N = np.zeros((n, nq))
    dN = np.zeros((n, nq, 3))
    w = np.zeros(nq)
    xi = np.zeros((n, 3))
    lib().vho_fe_tables(degree, _P(N), _P(dN), _P(w), _P(xi))
    return N, dN, w, xi


def face_tables(degree, face_no):
    n = nodes_per_cell(degree)
    nqf = (degree + 1) ** 2
    Nf = np.zeros((n, nqf))
    wf = np.zeros(nqf)
    lib().vho_face_tables(degree, face_no, _P(Nf), _P(wf))
    return Nf, wf


def pointwise(A18, coef):
    A18 = np.ascontiguousarray(A18, dtype=np.float64)
    coef = np.ascontiguousarray(coef, dtype=np.float64)
    g = np.zeros(18)
    H = np.zeros((18, 18))
    f = np.zeros(1)
    lib().vho_pointwise(_P(A18), _P(coef), _P(g), _P(H), _P(f))
    return g, H, f[0]


def cell(degree, origin, h, U, coef, faces=(), want_matrix=True):
    """O2 cell matrix / rhs / energy.  faces = [(face_no, boundary_id), ...]."""
    n = nodes_per_cell(degree)
    dpc = 18 * n
    origin = np.ascontiguousarray(origin, dtype=np.float64)
    h = np.ascontiguousarray(h, dtype=np.float64)
    U = np.ascontiguousarray(U, dtype=np.float64)
    coef = np.ascontiguousarray(coef, dtype=np.float64)
    fno = np.array([f[0] for f in faces], dtype=np.int32)
    fbid = np.array([f[1] for f in faces], dtype=np.int32)
    K = np.zeros((dpc, dpc)) if want_matrix else None
    r = np.zeros(dpc)
    e = np.zeros(1)
    lib().vho_cell(degree, _P(origin), _P(h), _P(U), _P(coef), len(faces), _I(fno), _I(fbid), _P(K), _P(r), _P(e))
    return K, r, e[0]


def ref_cell(degree, origin, h, U, coef, faces=(), want_matrix=True):
    """O1: the reference's literal (q,i,j) loops over its own term functions, on one box cell.

    FEValues / FEFaceValues are replaced by tables built here (real-space gradients, JxW)."""
    n = nodes_per_cell(degree)
    dpc = 18 * n
    nq = n
    nqf = (degree + 1) ** 2
    h = np.asarray(h, dtype=np.float64)
    N, dN, w, _ = fe_tables(degree)
    dNr = np.ascontiguousarray(dN / h[None, None, :])
    JxW = np.ascontiguousarray(w * h.prod())
    U = np.ascontiguousarray(U, dtype=np.float64)
    coef = np.ascontiguousarray(coef, dtype=np.float64)
    nf = len(faces)
    Nf = np.zeros((max(nf, 1), n, nqf))
    JxWf = np.zeros((max(nf, 1), nqf))
    bid = np.zeros(max(nf, 1), dtype=np.int32)
    for k, (fno, b) in enumerate(faces):
        t, wf = face_tables(degree, fno)
        nd = fno // 2
        area = np.prod([h[d] for d in range(3) if d != nd])
        Nf[k] = t
        JxWf[k] = wf * area
        bid[k] = b
    K = np.zeros((dpc, dpc)) if want_matrix else None
    r = np.zeros(dpc)
    ref().vhref_cell(n, nq, _P(N), _P(dNr), _P(JxW), _P(U), _P(coef), nf, _I(bid), nqf, _P(Nf), _P(JxWf),
                     1 if want_matrix else 0, _P(K), _P(r))
    return K, r


def ref_matep(p, t, scc):
    out = np.zeros(12)
    ref().vhref_matep(float(p), float(t), int(bool(scc)), _P(out))
    keys = ["alpha", "beta1", "beta2", "beta3", "beta4", "beta5", "gapA", "gapB", "fA", "fB", "Tcp_mK", "tAB_RWS"]
    return dict(zip(keys, out.tolist()))


def cells(degree, cell_nodes, cell_origin, cell_h, x, coef, face_ptr=None, face_no=None, face_bid=None, want_matrix=True,
          want_energy=False):
    """O2 over many cells (OpenMP).  x: [n_local_nodes*18].  Returns K[cells,dpc,dpc] | None, r[cells,dpc], e[cells]|None."""
    n = nodes_per_cell(degree)
    dpc = 18 * n
    cell_nodes = np.ascontiguousarray(cell_nodes, dtype=np.int32)
    nc = cell_nodes.shape[0]
    cell_origin = np.ascontiguousarray(cell_origin, dtype=np.float64)
    cell_h = np.ascontiguousarray(cell_h, dtype=np.float64)
    x = np.ascontiguousarray(x, dtype=np.float64)
    coef = np.ascontiguousarray(coef, dtype=np.float64)
    if face_ptr is None:
        face_ptr = np.zeros(nc + 1, dtype=np.int32)
        face_no = np.zeros(1, dtype=np.int32)
        face_bid = np.zeros(1, dtype=np.int32)
    face_ptr = np.ascontiguousarray(face_ptr, dtype=np.int32)
    face_no = np.ascontiguousarray(face_no, dtype=np.int32)
    face_bid = np.ascontiguousarray(face_bid, dtype=np.int32)
    K = np.zeros((nc, dpc, dpc)) if want_matrix else None
    r = np.zeros((nc, dpc))
    e = np.zeros(nc) if want_energy else None
    lib().vho_cells(degree, nc, _I(cell_nodes), _P(cell_origin), _P(cell_h), _P(x), _P(coef), _I(face_ptr), _I(face_no),
                    _I(face_bid), _P(K), _P(r), _P(e))
    return K, r, e
