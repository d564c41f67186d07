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


# ------------------------------------------------------------------------------------------
# global level: constrained scatter, GMRES + block-Jacobi, Newton with line search
# ------------------------------------------------------------------------------------------
def _constraint_maps(T):
    """C (n_local_dofs x n_local_dofs, scipy CSR) with C[i,i]=1 for unconstrained i, C[i,m]=w for constrained i;
    and the boolean mask of constrained DoFs.  Closed, homogeneous constraints."""
    import scipy.sparse as sp
    NL = 18 * T.n_local_nodes
    con = np.zeros(NL, dtype=bool)
    con[T.c_dof] = True
    rows = [np.nonzero(~con)[0]]
    cols = [np.nonzero(~con)[0]]
    vals = [np.ones(rows[0].size)]
    if T.c_master.size:
        cnt = np.diff(T.c_ptr)
        rows.append(np.repeat(T.c_dof, cnt))
        cols.append(T.c_master.astype(np.int64))
        vals.append(T.c_weight)
    C = sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(NL, NL))
    return C, con


def assemble_global(T, x_local, coef, want_matrix=True, chunk=256):
    """What assemble_system() / compute_residual() leave in system_matrix / system_rhs on this rank's owned rows.

    Restates AffineConstraints::distribute_local_to_global (call sites assemble.cc:356-361, residual.cc:287-289,
    rules in SURVEY.md A.4): unconstrained rows/columns receive C^T K C, constrained rows/columns receive nothing
    but sum_cells |a_ii| (mean |diag| of the cell if a_ii == 0) on the diagonal; rhs = C^T r, 0 on constrained DoFs.
    Returns (A: scipy CSR [n_owned_dofs x n_local_dofs] or None, rhs[n_owned_dofs])."""
    import scipy.sparse as sp
    n = nodes_per_cell(T.degree)
    dpc = 18 * n
    NL = 18 * T.n_local_nodes
    NO = 18 * T.n_owned_nodes
    fptr, fno, fbid = T.face_csr()
    C, con = _constraint_maps(T)
    dofs = (18 * T.cell_nodes.astype(np.int64)[:, :, None] + np.arange(18)[None, None, :]).reshape(T.n_cells, dpc)
    rvec = np.zeros(NL)
    diag = np.zeros(NL)
    Ku = sp.csr_matrix((NL, NL)) if want_matrix else None
    dummy = np.zeros(1, np.int32)
    for c0 in range(0, T.n_cells, chunk):
        c1 = min(T.n_cells, c0 + chunk)
        sl = slice(c0, c1)
        fp = fptr[c0:c1 + 1] - fptr[c0]
        has = fptr[c1] > fptr[c0]
        K, r, _ = cells(T.degree, T.cell_nodes[sl], T.cell_origin[sl], T.cell_h[sl], x_local, coef, fp,
                        fno[fptr[c0]:fptr[c1]] if has else dummy, fbid[fptr[c0]:fptr[c1]] if has else dummy,
                        want_matrix=want_matrix)
        np.add.at(rvec, dofs[sl].ravel(), r.ravel())
        if want_matrix:
            rr = np.repeat(dofs[sl], dpc, axis=1).ravel()
            cc = np.tile(dofs[sl], (1, dpc)).ravel()
            Ku = Ku + sp.csr_matrix((K.ravel(), (rr, cc)), shape=(NL, NL))
            d = np.abs(np.einsum("eii->ei", K))
            avg = d.mean(axis=1, keepdims=True)
            d = np.where(d != 0.0, d, avg)
            cm = con[dofs[sl]]
            np.add.at(diag, dofs[sl][cm], d[cm])
    rhs = (C.T @ rvec)
    rhs[con] = 0.0
    A = None
    if want_matrix:
        A = C.T @ Ku @ C + sp.diags(diag)
        A = sp.csr_matrix(A)[:NO, :]
    return A, rhs[:NO]


def block_jacobi_inverse(A, n_owned_nodes):
    """Inverse of the nodal 18x18 diagonal blocks of A (owned rows); returns [n_owned_nodes,18,18]."""
    import scipy.sparse as sp
    B = sp.bsr_matrix(A[:, :18 * n_owned_nodes], blocksize=(18, 18))
    B.sort_indices()
    D = np.zeros((n_owned_nodes, 18, 18))
    for I in range(n_owned_nodes):
        s, e = B.indptr[I], B.indptr[I + 1]
        k = np.searchsorted(B.indices[s:e], I)
        D[I] = B.data[s + k]
    return np.linalg.inv(D)


def gmres_block_jacobi(A, b, Minv, tol_abs, max_it=10000, restart=30, matvec_halo=None, dot=None):
    """Right-preconditioned restarted GMRES with deal.II SolverFGMRES iteration semantics:
    modified Gram-Schmidt in deal.II's add_and_dot order; from the second inner step on the (j+1) x j
    Hessenberg block is solved by least squares and its residual checked with ++accumulated_iterations.
    A: [n_owned x n_local]; matvec_halo(z_owned) -> z_local fills ghosts (identity for one rank);
    dot(a,b) all-reduced inner product.  Returns (x_owned, iterations, residual, converged)."""
    NO = b.size
    nb = Minv.shape[0]
    if matvec_halo is None:
        matvec_halo = lambda z: z  # noqa: E731
    if dot is None:
        dot = lambda a, c: float(a @ c)  # noqa: E731

    def prec(v):
        return np.einsum("ijk,ik->ij", Minv, v.reshape(nb, 18)).ravel()

    def check(step, val):
        if val <= tol_abs:
            return "success"
        if step >= max_it or np.isnan(val):
            return "failure"
        return "iterate"

    x = np.zeros(NO)
    acc = 0
    res = 0.0
    state = "iterate"
    m = restart
    while state == "iterate":
        aux = b - A @ matvec_halo(x) if np.any(x) else b.copy()
        beta = np.sqrt(dot(aux, aux))
        res = beta
        state = check(acc, res)
        if state == "success":
            break
        H = np.zeros((m + 1, m))
        V = np.zeros((m, NO))
        y = np.zeros(0)
        a = beta
        for j in range(m):
            V[j] = aux / a if a != 0 else 0.0
            z = prec(V[j])
            aux = A @ matvec_halo(z)
            H[0, j] = dot(aux, V[0])
            for i in range(1, j + 1):
                aux = aux - H[i - 1, j] * V[i - 1]
                H[i, j] = dot(aux, V[i])
            aux = aux - H[j, j] * V[j]
            a = np.sqrt(dot(aux, aux))
            H[j + 1, j] = a
            if j > 0:
                H1 = H[:j + 1, :j]
                prhs = np.zeros(j + 1)
                prhs[0] = beta
                y, _, _, _ = np.linalg.lstsq(H1, prhs, rcond=None)
                res = float(np.linalg.norm(prhs - H1 @ y))
                acc += 1
                state = check(acc, res)
                if state != "iterate":
                    break
        if y.size:
            x = x + prec(y @ V[:y.size])
    return x, acc, res, state == "success"


def distribute(T, x_local):
    """AffineConstraints::distribute: constrained entries overwritten from their masters (solve.cc:181, iteration.cc:180)."""
    x = x_local.copy()
    if T.c_dof.size:
        cnt = np.diff(T.c_ptr)
        vals = np.zeros(T.c_dof.size)
        if T.c_master.size:
            np.add.at(vals, np.repeat(np.arange(T.c_dof.size), cnt), T.c_weight * x_local[T.c_master])
        x[T.c_dof] = vals
    return x


def energy_global(T, x_local, coef):
    fptr, fno, fbid = T.face_csr()
    dummy = np.zeros(1, np.int32)
    _, _, e = cells(T.degree, T.cell_nodes, T.cell_origin, T.cell_h, x_local, coef, fptr,
                    fno if fno.size else dummy, fbid if fbid.size else dummy, want_matrix=False, want_energy=True)
    return float(e[T.cell_owned.astype(bool)].sum())


def newton_step(T, x, coef, lin_tol, max_lin_it=10000, restart=30, damped=True, step=0.83):
    """One pass of run.cc:214-218 on a single rank: assemble_system, solve(tol), newton_iteration.
    x: owned == local DoF vector.  Returns dict(x, rhs_norm, lin_its, lin_res, res_norm, alpha, n_trials, ...)."""
    assert T.n_ghost_nodes == 0
    A, rhs = assemble_global(T, x, coef, True)
    Minv = block_jacobi_inverse(A, T.n_owned_nodes)
    bn = float(np.linalg.norm(rhs))
    d, its, lres, ok = gmres_block_jacobi(A, rhs, Minv, lin_tol * bn, max_lin_it, restart)
    if not ok:
        raise RuntimeError("oracle GMRES did not converge")
    d = distribute(T, d)
    prev = bn
    alpha = 1.0
    n_trials = 0
    for i in range(100 if damped else 1):
        alpha = step ** i if damped else 1.0
        xt = distribute(T, x + alpha * d)
        _, r = assemble_global(T, xt, coef, False)
        cur = float(np.linalg.norm(r))
        n_trials += 1
        if damped and cur < prev:
            break
    return dict(x=xt, delta=d, rhs=rhs, rhs_norm=bn, lin_its=its, lin_res=lres, res_norm=cur, alpha=alpha,
                n_trials=n_trials, A=A, residual=r)


# ------------------------------------------------------------------------------------------
# geometric multigrid V-cycle as the GMRES preconditioner (csrc/vh_multigrid.cu restated; the GPU's optional replacement of
# the reference's ML-AMG, solve.cc:130-154).  Levels: finest first, single rank.  Same arithmetic as the device code:
# re-discretised coarse Jacobians at the injected state, Chebyshev around nodal block-Jacobi, lambda_max by power iteration
# from a fixed start vector ONCE per level (kept in `lam`), zero initial guess, Dirichlet DoFs masked after the transfers.
# ------------------------------------------------------------------------------------------
MG_DEFAULTS = dict(pre=1, post=1, smoothing_range=4.0, coarse_degree=8, coarse_range=30.0, n_power=8, safety=1.1)


def _dirichlet_mask(T):
    m = np.zeros(18 * T.n_local_nodes, dtype=bool)
    if T.c_dof.size:
        m[T.c_dof[np.diff(T.c_ptr) == 0]] = True
    return m


def mg_setup(tables, prolongations, x_fine, coef, lam=None, **params):
    """tables: per-level RankTables (finest first); prolongations[k] = (ptr, coarse_node, weight) between level k and k+1 (the
    table vh_mg_attach receives).  Returns (levels, lam): lam is the list of eigenvalue bounds, to be passed back in for the
    following Newton steps (the device estimates them once per context)."""
    import scipy.sparse as sp
    P = {**MG_DEFAULTS, **params}
    levels = []
    x = np.asarray(x_fine, dtype=np.float64)
    for k, T in enumerate(tables):
        assert T.n_ghost_nodes == 0
        A, _ = assemble_global(T, x, coef, True)
        A = sp.csr_matrix(A)
        Dinv = block_jacobi_inverse(A, T.n_owned_nodes)
        nb = T.n_owned_nodes
        prec = (lambda D, n: (lambda v: np.einsum("ijk,ik->ij", D, v.reshape(n, 18)).ravel()))(Dinv, nb)
        mask = _dirichlet_mask(T)
        lev = dict(T=T, A=A, prec=prec, mask=mask)
        if lam is not None:
            lev["lam"] = lam[k]
        else:
            g = (18 * T.node_global[:nb, None] + np.arange(18)[None, :]).ravel()
            v = np.where(g % 2 == 1, -1.0, 1.0) * (1.0 + (g % 7) / 7.0)
            v[mask] = 0.0
            for _ in range(P["n_power"]):
                v = v / np.sqrt(v @ v)
                v = prec(A @ v)
            lev["lam"] = P["safety"] * float(np.sqrt(v @ v))
        if k + 1 < len(tables):
            ptr, cn, w = prolongations[k]
            n_c = tables[k + 1].n_local_nodes
            Pn = sp.csr_matrix((w, cn, ptr), shape=(T.n_local_nodes, n_c))
            lev["P"] = sp.kron(Pn, sp.identity(18), format="csr")
            # injection of the Newton state: the fine node that coincides with the coarse node (its weight-1 entry)
            rows = np.repeat(np.arange(T.n_local_nodes), np.diff(ptr))
            one = np.abs(w - 1.0) <= 1e-12
            inj = np.zeros(n_c, dtype=np.int64)
            inj[cn[one]] = rows[one]
            x = x.reshape(-1, 18)[inj].ravel()
        levels.append(lev)
    return levels, [lev["lam"] for lev in levels]


def _chebyshev(lev, b, x, degree, ratio):
    if degree == 0:  # V(0,k) / V(k,0): no smoothing on this side of the cycle
        return np.zeros_like(b) if x is None else x
    A, prec = lev["A"], lev["prec"]
    lmax, lmin = lev["lam"], lev["lam"] / ratio
    theta, delta = 0.5 * (lmax + lmin), 0.5 * (lmax - lmin)
    sigma = theta / delta
    rho = 1.0 / sigma
    r = b if x is None else b - A @ x
    d = prec(r) * (1.0 / theta)
    x = d.copy() if x is None else x + d
    for _ in range(1, degree):
        rho_new = 1.0 / (2.0 * sigma - rho)
        r = b - A @ x
        d = (rho_new * rho) * d + (2.0 * rho_new / delta) * prec(r)
        x = x + d
        rho = rho_new
    return x


def mg_vcycle(levels, k, b, **params):
    """z = V(b) on level k with zero initial guess."""
    P = {**MG_DEFAULTS, **params}
    lev = levels[k]
    if k == len(levels) - 1:
        return _chebyshev(lev, b, None, P["coarse_degree"], P["coarse_range"])
    x = _chebyshev(lev, b, None, P["pre"], P["smoothing_range"])
    r = b - lev["A"] @ x if P["pre"] > 0 else b
    rc = lev["P"].T @ r
    rc[levels[k + 1]["mask"]] = 0.0
    xc = mg_vcycle(levels, k + 1, rc, **params)
    x = x + np.where(lev["mask"], 0.0, lev["P"] @ xc)
    x[lev["mask"]] = 0.0
    return _chebyshev(lev, b, x, P["post"], P["smoothing_range"])


def gmres_right(A, b, prec, tol_abs, max_it=10000, restart=30):
    """SolverFGMRES iteration semantics (A.5) with an arbitrary right preconditioner (callable): z_j = M^-1 v_j is stored and the
    update is x += sum_j y_j z_j, as in deal.II and in vh_gmres.cu with the multigrid preconditioner."""
    NO = b.size

    def check(step, val):
        if val <= tol_abs:
            return "success"
        if step >= max_it or np.isnan(val):
            return "failure"
        return "iterate"

    x = np.zeros(NO)
    acc, res, state, m = 0, 0.0, "iterate", restart
    while state == "iterate":
        aux = b - A @ x if np.any(x) else b.copy()
        beta = float(np.sqrt(aux @ aux))
        res = beta
        state = check(acc, res)
        if state == "success":
            break
        H = np.zeros((m + 1, m))
        V = np.zeros((m, NO))
        Z = np.zeros((m, NO))
        y = np.zeros(0)
        a = beta
        for j in range(m):
            V[j] = aux / a if a != 0 else 0.0
            Z[j] = prec(V[j])
            aux = A @ Z[j]
            H[0, j] = aux @ V[0]
            for i in range(1, j + 1):
                aux = aux - H[i - 1, j] * V[i - 1]
                H[i, j] = aux @ V[i]
            aux = aux - H[j, j] * V[j]
            a = float(np.sqrt(aux @ aux))
            H[j + 1, j] = a
            if j > 0:
                H1 = H[:j + 1, :j]
                prhs = np.zeros(j + 1)
                prhs[0] = beta
                y, _, _, _ = np.linalg.lstsq(H1, prhs, rcond=None)
                res = float(np.linalg.norm(prhs - H1 @ y))
                acc += 1
                state = check(acc, res)
                if state != "iterate":
                    break
        if y.size:
            x = x + y @ Z[:y.size]
    return x, acc, res, state == "success"


def newton_step_mg(tables, prolongations, x, coef, lin_tol, lam=None, damped=True, step=0.83, **params):
    """newton_step with the multigrid V-cycle as the GMRES preconditioner.  Returns (dict as newton_step, lam)."""
    T = tables[0]
    levels, lam = mg_setup(tables, prolongations, x, coef, lam=lam, **params)
    A = levels[0]["A"]
    _, rhs = assemble_global(T, x, coef, False)
    bn = float(np.linalg.norm(rhs))
    d, its, lres, ok = gmres_right(A, rhs, lambda v: mg_vcycle(levels, 0, v, **params), lin_tol * bn)
    if not ok:
        raise RuntimeError("oracle GMRES (multigrid) did not converge")
    d = distribute(T, d)
    alpha, n_trials, cur, xt = 1.0, 0, bn, x
    for i in range(100 if damped else 1):
        alpha = step ** i if damped else 1.0
        xt = distribute(T, x + alpha * d)
        _, r = assemble_global(T, xt, coef, False)
        cur = float(np.linalg.norm(r))
        n_trials += 1
        if damped and cur < bn:
            break
    return dict(x=xt, delta=d, rhs=rhs, rhs_norm=bn, lin_its=its, lin_res=lres, res_norm=cur, alpha=alpha, n_trials=n_trials,
                A=A, levels=levels), lam
