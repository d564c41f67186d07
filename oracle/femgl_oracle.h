/* TEST INFRASTRUCTURE — the parity checker, not the product.
 *
 * "O2": a plain-C CPU restatement of the arithmetic of VerHem's femgl Newton hot path
 * (cell level).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may
 * call it.  The CUDA library never links or loads it.
 *
 * Pinned against the reference's own code ("O1" = oracle/_ref/libvhref.so, the reference's
 * cell_mat_vec term files compiled verbatim) by tests/test_oracle_vs_reference.py, and against the
 * golden vectors of tests/golden/ (generated from O1 by tests/golden/make_golden.py).
 *
 * Conventions (/root/reference/femgl/src/femgl.cc:133-137):
 *   component c in [0,18): c<9 -> u[c/3][c%3] (real part), c>=9 -> v[(c-9)/3][(c-9)%3] (imag)
 *   local DoF i of a cell = 18*a + c   (a = local node in deal.II FE_Q order)
 *   coef[10] = {K1, K2, K3, alpha, beta1..beta5, bt}
 */
#ifndef VH_FEMGL_ORACLE_H
#define VH_FEMGL_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

/* nodes per cell / quadrature points for degree 1|2 */
int vho_nodes_per_cell(int degree);
int vho_n_q(int degree);
int vho_n_qf(int degree);

/* Reference-cell tables.  N[a*nq+q], dN[(a*nq+q)*3+k] (d/dxi_k on the unit cell), w[q];
 * node_xi[a*3+k] = support point of local node a; quadrature = QGauss<3>(degree+1), x fastest. */
void vho_fe_tables(int degree, double *N, double *dN, double *w, double *node_xi);
/* Face tables for deal.II face_no 0..5 (x=0,x=1,y=0,y=1,z=0,z=1): Nf[a*nqf+q], wf[q] (unit face). */
void vho_face_tables(int degree, int face_no, double *Nf, double *wf);

/* Pointwise bulk terms at one quadrature point: g[18] = alpha*A + 2*sum beta_k G_k,
 * H[18*18] = dg/dA (row-major, symmetric), f = bulk energy density alpha*I0 + sum beta_k I_k. */
void vho_pointwise(const double *A18, const double *coef, double *g18, double *H324, double *f);

/* One axis-aligned box cell [origin, origin+h].  U[18n] local DoF values.
 * faces: n_faces wall faces, face_no[f] in 0..5, face_bid[f] in {2,3,4}; skipped if bt>=1e10.
 * Outputs (may be NULL): K[dpc*dpc] row-major cell matrix, r[dpc] cell rhs (= -residual),
 * energy[1] cell contribution to the GL functional F. */
void vho_cell(int degree, const double *origin, const double *h, const double *U, const double *coef, int n_faces,
              const int *face_no, const int *face_bid, double *K, double *r, double *energy);

/* Batch over cells (OpenMP).  cell_nodes[n_cells*n] index into x (18 values per node).
 * wall faces given as CSR per cell: face_ptr[n_cells+1], face_no[], face_bid[].
 * K may be NULL (residual only).  K layout [cell][dpc][dpc], r [cell][dpc], energy [cell]. */
void vho_cells(int degree, int n_cells, const int *cell_nodes, const double *cell_origin, const double *cell_h,
               const double *x, const double *coef, const int *face_ptr, const int *face_no, const int *face_bid,
               double *K, double *r, double *energy);

#ifdef __cplusplus
}
#endif
#endif
