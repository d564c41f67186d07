"""Generate the golden fixtures of tests/golden/ from the REFERENCE'S OWN CODE (O1).

Run in the build container, where /root/reference exists:

    python tests/golden/make_golden.py

It compiles oracle/_ref/libvhref.so (the reference's cell_mat_vec/*.cc + matep.cc, verbatim, in
place) and records what that code returns on seeded inputs.  The fixtures are what pins the
C restatement (oracle/femgl_oracle.c) and, through it, the CUDA path; the GPU box has no
/root/reference, so tests there read only these files.

Fixtures:
  matep.json        Matep coefficients on a (p, t, SCC) grid              (matep.cc:46-406)
  pointwise.npz     the 6 rhs / 6 lhs bulk forms + 2 gradient forms on random A, all (i,j)
  cell_q1.npz       one Q1 box cell with two Robin wall faces: full 144x144 matrix + rhs
  cell_q1_res.npz   residual.cc variant (rhs only) of another Q1 cell, bt = 1e10 (faces off)
  cell_q2.npz       one Q2 box cell (486 DoFs, one wall face): rhs, K@z, z^T K, sampled entries
"""
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
import femgl_oracle as O  # noqa: E402


def coef_for(p, t, scc, bt):
    m = O.ref_matep(p, t, scc)
    return np.array([0.42072, 0.42072, 0.42072, m["alpha"], m["beta1"], m["beta2"], m["beta3"], m["beta4"], m["beta5"], bt])


def main():
    O.build(ref=True)
    R = O.ref()
    P = O._P
    rng = np.random.default_rng(20250101)

    # ---- matep ----
    grid = []
    for scc in (0, 1):
        for p in (0.0, 3.3, 12.0, 21.22, 25.0, 33.9):
            for t in (0.1, 0.5, 0.9):
                grid.append({"p": p, "t": t, "scc": scc, **O.ref_matep(p, t, scc)})
    with open(os.path.join(HERE, "matep.json"), "w") as f:
        json.dump(grid, f, indent=0)

    # ---- pointwise forms ----
    nA = 4
    A = rng.uniform(-1, 1, size=(nA, 18))
    rhs = np.zeros((nA, 18, 6))
    lhs = np.zeros((nA, 18, 18, 6))

    def phi(c):
        pu = np.zeros(9)
        pv = np.zeros(9)
        (pu if c < 9 else pv)[c % 9] = 1.0
        return pu, pv

    o6 = np.zeros(6)
    for k in range(nA):
        u = np.ascontiguousarray(A[k, :9])
        v = np.ascontiguousarray(A[k, 9:])
        for i in range(18):
            pui, pvi = phi(i)
            R.vhref_rhs_terms(P(u), P(v), P(pui), P(pvi), P(o6))
            rhs[k, i] = o6
            for j in range(18):
                puj, pvj = phi(j)
                R.vhref_lhs_terms(P(u), P(v), P(pui), P(pvi), P(puj), P(pvj), P(o6))
                lhs[k, i, j] = o6
    gi = np.array([0.11, -0.23, 0.37])
    gj = np.array([-0.41, 0.13, 0.29])
    grad = np.zeros((18, 18, 2))
    o2 = np.zeros(2)

    def gphi(c, g):
        gu = np.zeros(27)
        gv = np.zeros(27)
        for k in range(3):
            (gu if c < 9 else gv)[9 * k + c % 9] = g[k]
        return gu, gv

    for i in range(18):
        for j in range(18):
            gui, gvi = gphi(i, gi)
            guj, gvj = gphi(j, gj)
            R.vhref_lhs_grad_terms(P(gui), P(gvi), P(guj), P(gvj), P(o2))
            grad[i, j] = o2
    np.savez_compressed(os.path.join(HERE, "pointwise.npz"), A=A, rhs=rhs, lhs=lhs, grad_i=gi, grad_j=gj, grad=grad)

    # ---- Q1 cell, two wall faces, diffuse walls ----
    coef = coef_for(25.0, 0.5, 1, 2.0)
    U = rng.uniform(-1, 1, size=144) * 2.0
    h = np.array([0.7, 1.1, 0.9])
    faces = [(4, 4), (1, 2)]
    t0 = time.time()
    K, r = O.ref_cell(1, [0, 0, 0], h, U, coef, faces)
    print("Q1 cell (literal reference loops): %.1f s" % (time.time() - t0))
    np.savez_compressed(os.path.join(HERE, "cell_q1.npz"), U=U, h=h, coef=coef, faces=np.array(faces), K=K, r=r)

    # ---- Q1 residual-only, specular walls (faces must be ignored), weak-coupling betas ----
    coef2 = coef_for(12.0, 0.9, 0, 1e10)
    U2 = rng.uniform(-1, 1, size=144) * 3.0
    h2 = np.array([1.25, 1.25, 1.25])
    _, r2 = O.ref_cell(1, [0, 0, 0], h2, U2, coef2, [(5, 4)], want_matrix=False)
    np.savez_compressed(os.path.join(HERE, "cell_q1_res.npz"), U=U2, h=h2, coef=coef2, faces=np.array([(5, 4)]), r=r2)

    # ---- Q2 cell ----
    U3 = rng.uniform(-1, 1, size=486) * 2.0
    h3 = np.array([0.8, 0.6, 1.3])
    faces3 = [(2, 3)]
    t0 = time.time()
    K3, r3 = O.ref_cell(2, [0, 0, 0], h3, U3, coef, faces3)
    print("Q2 cell (literal reference loops): %.1f s" % (time.time() - t0))
    z = rng.uniform(-1, 1, size=(486, 2))
    idx = rng.integers(0, 486, size=(4000, 2))
    np.savez_compressed(os.path.join(HERE, "cell_q2.npz"), U=U3, h=h3, coef=coef, faces=np.array(faces3), r=r3, z=z, Kz=K3 @ z,
                        zK=z.T @ K3, idx=idx, Kidx=K3[idx[:, 0], idx[:, 1]], Kdiag=np.diag(K3).copy(),
                        Kabsmax=np.abs(K3).max())


if __name__ == "__main__":
    main()
