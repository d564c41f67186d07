"""Experiment (not product): would a geometric multigrid V-cycle beat block-Jacobi as the GMRES preconditioner at r6/r7?
Every level is an ordinary GPU context; operator and block-Jacobi applies go through the C ABI with host buffers (slow, but
the question is the ITERATION COUNT).  Usage: python tools/mg_experiment.py <refine> <coarsest> <newton steps>"""
import os
import sys
import time

import numpy as np
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import verkko_hem_repo_b200 as vh  # noqa: E402
from helpers import MATEP_SCC_ON, b_phase_state, coef_vector  # noqa: E402


def transfer_nodes(Tf, Tc):
    xf, xc = Tf.node_xyz, Tc.node_xyz
    lo = xc.min(axis=0)
    axes = [np.unique(np.round(xc[:, d] - lo[d], 9)) for d in range(3)]
    H = [a[1] - a[0] for a in axes]
    nper = [a.size for a in axes]
    key_c = np.zeros(xc.shape[0], dtype=np.int64)
    for d in range(3):
        key_c = key_c * (nper[d] + 1) + np.rint((xc[:, d] - lo[d]) / H[d]).astype(np.int64)
    order = np.argsort(key_c)
    skeys = key_c[order]
    t = [(xf[:, d] - lo[d]) / H[d] for d in range(3)]
    i0 = [np.minimum(np.floor(t[d] + 1e-9).astype(np.int64), nper[d] - 1) for d in range(3)]
    fr = [t[d] - i0[d] for d in range(3)]
    nf = xf.shape[0]
    rows, cols, vals = [], [], []
    for corner in range(8):
        w = np.ones(nf)
        key = np.zeros(nf, dtype=np.int64)
        for d in range(3):
            b = (corner >> d) & 1
            w = w * (fr[d] if b else 1.0 - fr[d])
            key = key * (nper[d] + 1) + i0[d] + b
        sel = np.abs(w) > 1e-12
        pos = np.minimum(np.searchsorted(skeys, key[sel]), skeys.size - 1)
        hit = skeys[pos] == key[sel]
        rows.append(np.nonzero(sel)[0][hit])
        cols.append(order[pos[hit]])
        vals.append(w[sel][hit])
    return sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(nf, xc.shape[0]))


def dmask(T):
    m = np.zeros(18 * T.n_local_nodes, dtype=bool)
    if T.c_dof.size:
        m[T.c_dof[np.diff(T.c_ptr) == 0]] = True
    return m


def cheb(lev, b, x, degree, ratio):
    lmax, lmin = lev["lam"], lev["lam"] / ratio
    theta, delta = 0.5 * (lmax + lmin), 0.5 * (lmax - lmin)
    sigma = theta / delta
    rho = 1.0 / sigma
    r = b.copy() if x is None else b - lev["A"](x)
    d = lev["M"](r) / theta
    x = d.copy() if x is None else x + d
    for _ in range(1, degree):
        rho_new = 1.0 / (2.0 * sigma - rho)
        r = b - lev["A"](x)
        d = rho_new * rho * d + (2.0 * rho_new / delta) * lev["M"](r)
        x = x + d
        rho = rho_new
    return x


def vcycle(levels, k, b, pre, post, ratio, cdeg):
    lev = levels[k]
    if k == len(levels) - 1:
        return cheb(lev, b, None, cdeg, 30.0)
    x = cheb(lev, b, None, pre, ratio)
    r = b - lev["A"](x)
    rc = (lev["P"].T @ r.reshape(-1, 18)).ravel()
    rc[levels[k + 1]["mask"]] = 0.0
    xc = vcycle(levels, k + 1, rc, pre, post, ratio, cdeg)
    x = x + (lev["P"] @ xc.reshape(-1, 18)).ravel()
    x[lev["mask"]] = 0.0
    return cheb(lev, b, x, post, ratio)


def gmres_right(A, b, prec, tol_abs, max_it=300, m=30):
    x = np.zeros(b.size)
    acc, state = 0, "iterate"
    while state == "iterate":
        aux = b - A(x) if np.any(x) else b.copy()
        beta = float(np.sqrt(aux @ aux))
        if beta <= tol_abs:
            break
        if acc >= max_it:
            return x, acc, False
        H = np.zeros((m + 1, m))
        V = []
        y = np.zeros(0)
        a = beta
        for j in range(m):
            V.append(aux / a)
            aux = A(prec(V[j]))
            for i in range(j + 1):
                H[i, j] = aux @ V[i]
                aux = aux - H[i, j] * V[i]
            a = float(np.sqrt(aux @ aux))
            H[j + 1, j] = a
            if j > 0:
                H1 = H[:j + 1, :j]
                prhs = np.zeros(j + 1)
                prhs[0] = beta
                y, _, _, _ = np.linalg.lstsq(H1, prhs, rcond=None)
                res = float(np.linalg.norm(prhs - H1 @ y))
                acc += 1
                if res <= tol_abs:
                    state = "success"
                    break
                if acc >= max_it:
                    state = "failure"
                    break
        if y.size:
            x = x + prec(sum(yi * Vi for yi, Vi in zip(y, V)))
    return x, acc, state == "success"


def main():
    r, rmin, nsteps = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
    L = 20.0
    coef = coef_vector(MATEP_SCC_ON, 2.0)
    tabs, ctxs = [], []
    for lv in range(r, rmin - 1, -1):
        T = vh.Mesh(1, [-L] * 3, [L] * 3, base=(1, 1, 1), face_bid=(1, 1, 1, 1, 4, 4), n_global_refine=lv).finalize(1).tables(0)
        c = vh.Context(T)
        c.set_coef_vector(coef)
        tabs.append(T)
        ctxs.append(c)
    Ps, injs = [], []
    for k in range(len(tabs) - 1):
        Pn = transfer_nodes(tabs[k], tabs[k + 1])
        Ps.append(Pn)
        Pc = Pn.tocsc()
        big = Pc.data > 1.0 - 1e-9
        inj = np.zeros(tabs[k + 1].n_local_nodes, dtype=np.int64)
        cols = np.repeat(np.arange(Pc.shape[1]), np.diff(Pc.indptr))
        inj[cols[big]] = Pc.indices[big]
        injs.append(inj)
    x = b_phase_state(tabs[0], MATEP_SCC_ON, noise=0.0)
    fine = ctxs[0]
    for step in range(nsteps):
        t0 = time.time()
        fine.set_solution(x)
        bn = fine.assemble()
        rhs = fine.get_rhs()
        its_bj, _ = fine.solve(1e-1)
        # coarse levels: re-discretised at the injected state
        levels = []
        xs = x
        for k, (T, c) in enumerate(zip(tabs, ctxs)):
            if k > 0:
                xs = xs.reshape(-1, 18)[injs[k - 1]].ravel()
                c.set_solution(xs)
                c.assemble()
            lev = {"A": c.spmv, "M": c.precondition, "mask": dmask(T), "P": Ps[k] if k < len(Ps) else None}
            n = 18 * T.n_owned_nodes
            v = np.where((np.arange(n) % 2) == 0, 1.0, -1.0) * (1.0 + (np.arange(n) % 7) / 7.0)
            v[lev["mask"]] = 0.0
            lam = 1.0
            for _ in range(8):
                v = v / np.linalg.norm(v)
                v = c.precondition(c.spmv(v))
                lam = float(np.linalg.norm(v))
            lev["lam"] = 1.1 * lam
            levels.append(lev)
        out = []
        for pre, post, ratio, cdeg in ((2, 2, 8.0, 12), (1, 1, 4.0, 8)):
            _, its, ok = gmres_right(fine.spmv, rhs, lambda v: vcycle(levels, 0, v, pre, post, ratio, cdeg), 1e-1 * bn)
            out.append(("V(%d,%d)" % (pre, post), its, ok))
        # Newton path of the block-Jacobi run (what bench.py executes)
        prev = bn
        for i in range(100):
            fine.line_search_trial(0.83 ** i)
            cur = fine.residual()
            if cur < prev:
                break
        fine.accept_trial()
        x = fine.get_solution()
        print("step", step, "||rhs|| %.4e" % bn, "BJ its", its_bj, "MG", out, "lam", [round(l["lam"], 3) for l in levels],
              "t %.0fs" % (time.time() - t0), flush=True)


if __name__ == "__main__":
    main()
