/* TEST INFRASTRUCTURE — the parity checker, not the product.  See femgl_oracle.h.
 *
 * CPU restatement ("O2") of the cell-level arithmetic of the reference's
 *   FemGL::assemble_system   /root/reference/femgl/src/assemble.cc:177-361
 *   FemGL::compute_residual  /root/reference/femgl/src/residual.cc:166-289
 * using the closed forms of SURVEY.md Appendix A.2/A.3 instead of the (q,i,j) loops over
 * 3x3 FullMatrix products.  Deliberately written with straightforward complex 3x3 algebra
 * (directional derivatives of the cubic matrix polynomials), independent of the CUDA kernels.
 */
#include "femgl_oracle.h"

#include <complex.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef double complex cx;

/* ---------- 3x3 complex helpers (row-major, index 3*mu+j) ---------- */
static void mm(cx *C, const cx *A, const cx *B)
{
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j)
      {
        cx s = 0;
        for (int k = 0; k < 3; ++k)
          s += A[3 * i + k] * B[3 * k + j];
        C[3 * i + j] = s;
      }
}
static void tp(cx *C, const cx *A)
{
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j)
      C[3 * i + j] = A[3 * j + i];
}
static void cj(cx *C, const cx *A)
{
  for (int i = 0; i < 9; ++i)
    C[i] = conj(A[i]);
}
static cx trace3(const cx *A) { return A[0] + A[4] + A[8]; }
static void mm3(cx *C, const cx *X, const cx *Y, const cx *Z)
{
  cx T[9];
  mm(T, X, Y);
  mm(C, T, Z);
}

/* g(A) = alpha A + 2 sum_k beta_k G_k(A):
 *  G1 = tr(A A^T) A*    (cell_vec_rhs_beta1.cc:108-165)
 *  G2 = tr(A A^+) A     (cell_vec_rhs_beta2.cc:108-146)
 *  G3 = A A^T A*        (cell_vec_rhs_beta3.cc:108-204)
 *  G4 = A A^+ A         (cell_vec_rhs_beta4.cc:108-210)
 *  G5 = A* A^T A        (cell_vec_rhs_beta5.cc:108-216)            */
static void g_of_A(const cx *A, const double *coef, cx *g)
{
  const double alpha = coef[3], *b = coef + 4;
  cx           At[9], Ac[9], Ah[9], AAt[9], AAh[9], G[9];
  tp(At, A);
  cj(Ac, A);
  tp(Ah, Ac);
  mm(AAt, A, At);
  mm(AAh, A, Ah);
  const cx T = trace3(AAt), S = trace3(AAh);
  for (int i = 0; i < 9; ++i)
    g[i] = alpha * A[i] + 2.0 * (b[0] * T * Ac[i] + b[1] * S * A[i]);
  mm3(G, A, At, Ac);
  for (int i = 0; i < 9; ++i)
    g[i] += 2.0 * b[2] * G[i];
  mm3(G, A, Ah, A);
  for (int i = 0; i < 9; ++i)
    g[i] += 2.0 * b[3] * G[i];
  mm3(G, Ac, At, A);
  for (int i = 0; i < 9; ++i)
    g[i] += 2.0 * b[4] * G[i];
}

/* Directional derivative dg(A)[E]  (exact derivative of the forms above; the reference's
 * mat_lhs_beta_k(phi_i,phi_j) = d/d eps vec_rhs_beta_k(A + eps phi_j; phi_i)). */
static void dg_of_A(const cx *A, const cx *E, const double *coef, cx *dg)
{
  const double alpha = coef[3], *b = coef + 4;
  cx           At[9], Ac[9], Ah[9], Et[9], Ec[9], Eh[9], AAt[9], AAh[9], P[9], Q[9];
  tp(At, A);
  cj(Ac, A);
  tp(Ah, Ac);
  tp(Et, E);
  cj(Ec, E);
  tp(Eh, Ec);
  mm(AAt, A, At);
  mm(AAh, A, Ah);
  const cx T = trace3(AAt), S = trace3(AAh);
  mm(P, E, At);
  const cx dT = 2.0 * trace3(P); /* d tr(A A^T) = 2 tr(E A^T) */
  mm(Q, E, Ah);
  const double dS = 2.0 * creal(trace3(Q)); /* d tr(A A^+) = 2 Re tr(E A^+) */
  for (int i = 0; i < 9; ++i)
    dg[i] = alpha * E[i] + 2.0 * (b[0] * (dT * Ac[i] + T * Ec[i]) + b[1] * (dS * A[i] + S * E[i]));
  cx X[9];
  /* G3 */
  mm3(X, E, At, Ac);
  for (int i = 0; i < 9; ++i)
    dg[i] += 2.0 * b[2] * X[i];
  mm3(X, A, Et, Ac);
  for (int i = 0; i < 9; ++i)
    dg[i] += 2.0 * b[2] * X[i];
  mm3(X, A, At, Ec);
  for (int i = 0; i < 9; ++i)
    dg[i] += 2.0 * b[2] * X[i];
  /* G4 */
  mm3(X, E, Ah, A);
  for (int i = 0; i < 9; ++i)
    dg[i] += 2.0 * b[3] * X[i];
  mm3(X, A, Eh, A);
  for (int i = 0; i < 9; ++i)
    dg[i] += 2.0 * b[3] * X[i];
  mm3(X, A, Ah, E);
  for (int i = 0; i < 9; ++i)
    dg[i] += 2.0 * b[3] * X[i];
  /* G5 */
  mm3(X, Ec, At, A);
  for (int i = 0; i < 9; ++i)
    dg[i] += 2.0 * b[4] * X[i];
  mm3(X, Ac, Et, A);
  for (int i = 0; i < 9; ++i)
    dg[i] += 2.0 * b[4] * X[i];
  mm3(X, Ac, At, E);
  for (int i = 0; i < 9; ++i)
    dg[i] += 2.0 * b[4] * X[i];
}

void vho_pointwise(const double *A18, const double *coef, double *g18, double *H324, double *f)
{
  cx A[9], g[9];
  for (int i = 0; i < 9; ++i)
    A[i] = A18[i] + I * A18[9 + i];
  if (g18)
    {
      g_of_A(A, coef, g);
      for (int i = 0; i < 9; ++i)
        {
          g18[i]     = creal(g[i]);
          g18[9 + i] = cimag(g[i]);
        }
    }
  if (H324)
    for (int d = 0; d < 18; ++d)
      {
        cx E[9], dg[9];
        for (int i = 0; i < 9; ++i)
          E[i] = 0;
        E[d % 9] = (d < 9) ? 1.0 : I;
        dg_of_A(A, E, coef, dg);
        for (int i = 0; i < 9; ++i)
          {
            H324[18 * i + d]       = creal(dg[i]);
            H324[18 * (9 + i) + d] = cimag(dg[i]);
          }
      }
  if (f)
    { /* bulk free-energy density, SURVEY.md A.1 */
      const double alpha = coef[3], *b = coef + 4;
      cx           At[9], Ac[9], Ah[9], AAt[9], AAh[9], C1[9], C2[9];
      tp(At, A);
      cj(Ac, A);
      tp(Ah, Ac);
      mm(AAt, A, At);
      mm(AAh, A, Ah);
      const cx     T  = trace3(AAt);
      const double S  = creal(trace3(AAh));
      const double I1 = creal(T * conj(T)), I2 = S * S;
      cj(C1, AAt);
      mm(C2, AAt, C1);
      const double I3 = creal(trace3(C2));
      mm(C2, AAh, AAh);
      const double I4 = creal(trace3(C2));
      cj(C1, AAh);
      mm(C2, AAh, C1);
      const double I5 = creal(trace3(C2));
      *f              = alpha * S + b[0] * I1 + b[1] * I2 + b[2] * I3 + b[3] * I4 + b[4] * I5;
    }
}

/* ---------- finite element tables [deal.II-internal semantics, SURVEY.md A.3] ---------- */
int vho_nodes_per_cell(int degree) { return degree == 1 ? 8 : 27; }
int vho_n_q(int degree) { return degree == 1 ? 8 : 27; }
int vho_n_qf(int degree) { return degree == 1 ? 4 : 9; }

/* Gauss-Legendre on [0,1], QGauss<1>(n) for n = 2,3 */
static void gauss01(int n, double *x, double *w)
{
  if (n == 2)
    {
      const double d = 0.5 / sqrt(3.0);
      x[0] = 0.5 - d, x[1] = 0.5 + d;
      w[0] = w[1] = 0.5;
    }
  else
    {
      const double d = 0.5 * sqrt(0.6);
      x[0] = 0.5 - d, x[1] = 0.5, x[2] = 0.5 + d;
      w[0] = w[2] = 5.0 / 18.0, w[1] = 8.0 / 18.0;
    }
}

/* 1-D Lagrange basis on equidistant support points t/degree, t = 0..degree; value and derivative */
static void lagrange1d(int degree, int t, double xi, double *val, double *der)
{
  if (degree == 1)
    {
      *val = t ? xi : 1.0 - xi;
      *der = t ? 1.0 : -1.0;
    }
  else
    {
      if (t == 0)
        *val = (2 * xi - 1) * (xi - 1), *der = 4 * xi - 3;
      else if (t == 1)
        *val = 4 * xi * (1 - xi), *der = 4 - 8 * xi;
      else
        *val = xi * (2 * xi - 1), *der = 4 * xi - 1;
    }
}

/* local node a -> tensor index (tx,ty,tz), each in 0..degree.
 * Q1: vertices lexicographic (x fastest).  Q2: deal.II hierarchical order
 * vertices(8), lines(12), quads(6), hex(1); line/quad numbering of GeometryInfo<3>. */
static void node_tensor_index(int degree, int a, int *t)
{
  if (degree == 1)
    {
      t[0] = a & 1, t[1] = (a >> 1) & 1, t[2] = (a >> 2) & 1;
      return;
    }
  static const int q2[27][3] = {
    {0, 0, 0}, {2, 0, 0}, {0, 2, 0}, {2, 2, 0}, {0, 0, 2}, {2, 0, 2}, {0, 2, 2}, {2, 2, 2}, /* vertices */
    {0, 1, 0}, {2, 1, 0}, {1, 0, 0}, {1, 2, 0},                                              /* lines 0-3 (z=0) */
    {0, 1, 2}, {2, 1, 2}, {1, 0, 2}, {1, 2, 2},                                              /* lines 4-7 (z=1) */
    {0, 0, 1}, {2, 0, 1}, {0, 2, 1}, {2, 2, 1},                                              /* lines 8-11 (along z) */
    {0, 1, 1}, {2, 1, 1}, {1, 0, 1}, {1, 2, 1}, {1, 1, 0}, {1, 1, 2},                        /* quads x0,x1,y0,y1,z0,z1 */
    {1, 1, 1}};
  t[0] = q2[a][0], t[1] = q2[a][1], t[2] = q2[a][2];
}

static void shape_at(int degree, int a, const double *xi, double *val, double *grad)
{
  int t[3];
  node_tensor_index(degree, a, t);
  double v[3], d[3];
  for (int k = 0; k < 3; ++k)
    lagrange1d(degree, t[k], xi[k], &v[k], &d[k]);
  *val = v[0] * v[1] * v[2];
  if (grad)
    {
      grad[0] = d[0] * v[1] * v[2];
      grad[1] = v[0] * d[1] * v[2];
      grad[2] = v[0] * v[1] * d[2];
    }
}

void vho_fe_tables(int degree, double *N, double *dN, double *w, double *node_xi)
{
  const int n = vho_nodes_per_cell(degree), n1 = degree + 1, nq = n1 * n1 * n1;
  double    gx[3], gw[3];
  gauss01(n1, gx, gw);
  for (int q = 0; q < nq; ++q)
    {
      const int    qx = q % n1, qy = (q / n1) % n1, qz = q / (n1 * n1);
      const double xi[3] = {gx[qx], gx[qy], gx[qz]};
      if (w)
        w[q] = gw[qx] * gw[qy] * gw[qz];
      for (int a = 0; a < n; ++a)
        {
          double val, grad[3];
          shape_at(degree, a, xi, &val, grad);
          if (N)
            N[a * nq + q] = val;
          if (dN)
            for (int k = 0; k < 3; ++k)
              dN[(a * nq + q) * 3 + k] = grad[k];
        }
    }
  if (node_xi)
    for (int a = 0; a < n; ++a)
      {
        int t[3];
        node_tensor_index(degree, a, t);
        for (int k = 0; k < 3; ++k)
          node_xi[a * 3 + k] = (double)t[k] / degree;
      }
}

void vho_face_tables(int degree, int face_no, double *Nf, double *wf)
{
  const int n = vho_nodes_per_cell(degree), n1 = degree + 1, nqf = n1 * n1;
  double    gx[3], gw[3];
  gauss01(n1, gx, gw);
  const int nd = face_no / 2, side = face_no % 2;
  const int d0 = (nd == 0) ? 1 : 0, d1 = (nd == 2) ? 1 : 2; /* tangential dims, lower index fastest */
  for (int q = 0; q < nqf; ++q)
    {
      const int q0 = q % n1, q1 = q / n1;
      double    xi[3];
      xi[nd] = side;
      xi[d0] = gx[q0];
      xi[d1] = gx[q1];
      if (wf)
        wf[q] = gw[q0] * gw[q1];
      for (int a = 0; a < n; ++a)
        {
          double val;
          shape_at(degree, a, xi, &val, 0);
          Nf[a * nqf + q] = val;
        }
    }
}

/* ---------- cell level: SURVEY.md A.3 ---------- */
void vho_cell(int degree, const double *origin, const double *h, const double *U, const double *coef, int n_faces,
              const int *face_no, const int *face_bid, double *K, double *r, double *energy)
{
  (void)origin;
  const int    n = vho_nodes_per_cell(degree), nq = vho_n_q(degree), nqf = vho_n_qf(degree), dpc = 18 * n;
  const double K1 = coef[0], K23 = coef[1] + coef[2], bt = coef[9];
  double      *N = malloc(sizeof(double) * n * nq), *dN = malloc(sizeof(double) * n * nq * 3), *w = malloc(sizeof(double) * nq);
  vho_fe_tables(degree, N, dN, w, 0);
  const double vol = h[0] * h[1] * h[2];
  if (K)
    memset(K, 0, sizeof(double) * dpc * dpc);
  if (r)
    memset(r, 0, sizeof(double) * dpc);
  double E = 0;
  for (int q = 0; q < nq; ++q)
    {
      const double JxW = w[q] * vol;
      double       A[18], dA[18][3], g[18], H[324], f;
      for (int c = 0; c < 18; ++c)
        {
          A[c] = 0;
          dA[c][0] = dA[c][1] = dA[c][2] = 0;
        }
      for (int a = 0; a < n; ++a)
        for (int c = 0; c < 18; ++c)
          {
            A[c] += U[18 * a + c] * N[a * nq + q];
            for (int k = 0; k < 3; ++k)
              dA[c][k] += U[18 * a + c] * dN[(a * nq + q) * 3 + k] / h[k];
          }
      vho_pointwise(A, coef, g, K ? H : 0, &f);
      /* d_mu(A) = sum_z d_z A_{mu z}, per part (u,v) and spin row mu */
      double div[6];
      for (int pm = 0; pm < 6; ++pm)
        div[pm] = dA[3 * pm + 0][0] + dA[3 * pm + 1][1] + dA[3 * pm + 2][2];
      if (energy)
        {
          double e = f;
          for (int c = 0; c < 18; ++c)
            e += K1 * (dA[c][0] * dA[c][0] + dA[c][1] * dA[c][1] + dA[c][2] * dA[c][2]);
          for (int pm = 0; pm < 6; ++pm)
            e += K23 * div[pm] * div[pm];
          E += e * JxW;
        }
      for (int a = 0; a < n; ++a)
        {
          const double  Na = N[a * nq + q];
          const double *ga = &dN[(a * nq + q) * 3];
          const double  gra[3] = {ga[0] / h[0], ga[1] / h[1], ga[2] / h[2]};
          if (r)
            for (int c = 0; c < 18; ++c)
              {
                const double t = Na * g[c] + K1 * (gra[0] * dA[c][0] + gra[1] * dA[c][1] + gra[2] * dA[c][2]) +
                                 K23 * gra[c % 3] * div[c / 3];
                r[18 * a + c] -= t * JxW;
              }
          if (K)
            for (int b = 0; b < n; ++b)
              {
                const double  Nb = N[b * nq + q];
                const double *gb = &dN[(b * nq + q) * 3];
                const double  grb[3] = {gb[0] / h[0], gb[1] / h[1], gb[2] / h[2]};
                const double  gg = gra[0] * grb[0] + gra[1] * grb[1] + gra[2] * grb[2];
                for (int c = 0; c < 18; ++c)
                  for (int d = 0; d < 18; ++d)
                    {
                      double t = Na * Nb * H[18 * c + d];
                      if (c == d)
                        t += K1 * gg;
                      if (c / 3 == d / 3) /* same part and same spin row mu */
                        t += K23 * gra[c % 3] * grb[d % 3];
                      K[(size_t)(18 * a + c) * dpc + 18 * b + d] += t * JxW;
                    }
              }
        }
    }
  /* Robin wall faces (assemble.cc:286-348): +K1/bt * face mass on components whose orbital index != normal */
  if (bt < 1e10)
    {
      double *Nf = malloc(sizeof(double) * n * nqf), *wf = malloc(sizeof(double) * nqf);
      for (int f = 0; f < n_faces; ++f)
        {
          const int bid = face_bid[f];
          if (bid < 2 || bid > 4)
            continue;
          const int normal = bid - 2; /* id 2/3/4 <-> orbital column x/y/z zeroed (phi_vector2matrix.cc:160-194) */
          const int nd     = face_no[f] / 2;
          double    area   = 1;
          for (int k = 0; k < 3; ++k)
            if (k != nd)
              area *= h[k];
          vho_face_tables(degree, face_no[f], Nf, wf);
          for (int q = 0; q < nqf; ++q)
            {
              const double JxW = wf[q] * area, s = K1 / bt * JxW;
              double       A[18];
              for (int c = 0; c < 18; ++c)
                {
                  A[c] = 0;
                  for (int a = 0; a < n; ++a)
                    A[c] += U[18 * a + c] * Nf[a * nqf + q];
                }
              for (int c = 0; c < 18; ++c)
                {
                  if (c % 3 == normal)
                    continue;
                  if (energy)
                    E += s * A[c] * A[c];
                  for (int a = 0; a < n; ++a)
                    {
                      if (r)
                        r[18 * a + c] -= s * Nf[a * nqf + q] * A[c];
                      if (K)
                        for (int b = 0; b < n; ++b)
                          K[(size_t)(18 * a + c) * dpc + 18 * b + c] += s * Nf[a * nqf + q] * Nf[b * nqf + q];
                    }
                }
            }
        }
      free(Nf);
      free(wf);
    }
  if (energy)
    *energy = E;
  free(N);
  free(dN);
  free(w);
}

/* Batch over cells with a plain pthread fan-out (this image has no libgomp). */
#include <pthread.h>
#include <unistd.h>

typedef struct
{
  int           degree, e0, e1;
  const int    *cell_nodes;
  const double *cell_origin, *cell_h, *x, *coef;
  const int    *face_ptr, *face_no, *face_bid;
  double       *K, *r, *energy;
} cells_job;

static void *cells_worker(void *arg)
{
  const cells_job *j = (const cells_job *)arg;
  const int        n = vho_nodes_per_cell(j->degree), dpc = 18 * n;
  for (int e = j->e0; e < j->e1; ++e)
    {
      double U[18 * 27];
      for (int a = 0; a < n; ++a)
        memcpy(U + 18 * a, j->x + 18 * (size_t)j->cell_nodes[(size_t)e * n + a], 18 * sizeof(double));
      const int f0 = j->face_ptr ? j->face_ptr[e] : 0, nf = j->face_ptr ? j->face_ptr[e + 1] - f0 : 0;
      vho_cell(j->degree, j->cell_origin + 3 * (size_t)e, j->cell_h + 3 * (size_t)e, U, j->coef, nf, j->face_no + f0,
               j->face_bid + f0, j->K ? j->K + (size_t)e * dpc * dpc : 0, j->r ? j->r + (size_t)e * dpc : 0,
               j->energy ? j->energy + e : 0);
    }
  return 0;
}

void vho_cells(int degree, int n_cells, const int *cell_nodes, const double *cell_origin, const double *cell_h,
               const double *x, const double *coef, const int *face_ptr, const int *face_no, const int *face_bid,
               double *K, double *r, double *energy)
{
  long nt = sysconf(_SC_NPROCESSORS_ONLN);
  if (nt < 1)
    nt = 1;
  if (nt > 64)
    nt = 64;
  if (nt > n_cells)
    nt = n_cells > 0 ? n_cells : 1;
  pthread_t th[64];
  cells_job jobs[64];
  for (long t = 0; t < nt; ++t)
    {
      cells_job j = {degree, (int)((long long)n_cells * t / nt), (int)((long long)n_cells * (t + 1) / nt),
                     cell_nodes, cell_origin, cell_h, x, coef, face_ptr, face_no, face_bid, K, r, energy};
      jobs[t]     = j;
      pthread_create(&th[t], 0, cells_worker, &jobs[t]);
    }
  for (long t = 0; t < nt; ++t)
    pthread_join(th[t], 0);
}
