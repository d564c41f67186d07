// Pointwise bulk terms of the GL functional for the 3x3 complex order parameter A = u + i v.
//
// Replaces, in closed form, the reference's per-(q,i,j) evaluation through 3x3 FullMatrix products:
//   vec_rhs_alpha / vec_rhs_beta1..5   /root/reference/femgl/src/cell_mat_vec/cell_vec_rhs_{alpha,beta1..5}.cc:108-216
//   mat_lhs_alpha / mat_lhs_beta1..5   /root/reference/femgl/src/cell_mat_vec/cell_mat_lhs_{alpha,beta1..5}.cc:108-342
// with  g = alpha A + 2 sum_k beta_k G_k   (the explicit 2.0 is at assemble.cc:237,265),
//   G1 = tr(AA^T) A*, G2 = tr(AA^+) A, G3 = A A^T A*, G4 = A A^+ A, G5 = A* A^T A,
// and H = dg/dA (18x18 real, symmetric).  Because every test direction is a unit matrix s*e_nu e_k^T
// (s = 1 or i), each column of H needs only the four 3x3 products R = AA^T, Q = AA^+, P = A^+A, S = A^T A.
//
// Layout: A[18] = {u row-major (9), v row-major (9)}; prod[72] = {R,Q,P,S} x 9 entries x (re,im).
// All functions are __host__ __device__ so the same source is unit-tested on the CPU (tests/native).
#ifndef VH_POINTWISE_CUH
#define VH_POINTWISE_CUH

#ifndef __CUDACC__
#include <cmath>
#define VH_HD inline
#else
#define VH_HD __host__ __device__ __forceinline__
#endif

struct vh_cx
{
  double re, im;
};
VH_HD vh_cx vh_cmul(vh_cx a, vh_cx b) { return vh_cx{a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re}; }
VH_HD vh_cx vh_conj(vh_cx a) { return vh_cx{a.re, -a.im}; }
VH_HD vh_cx vh_ld(const double *A, int mu, int j) { return vh_cx{A[3 * mu + j], A[9 + 3 * mu + j]}; }
VH_HD vh_cx vh_ldp(const double *prod, int m, int i, int j)
{
  const double *p = prod + 18 * m + 2 * (3 * i + j);
  return vh_cx{p[0], p[1]};
}

// One entry e in [0,36) of the product table: matrix m = e/9, (i,j) = ((e%9)/3, e%3).
VH_HD void vh_product_entry(const double *A, int e, double *out2)
{
  const int m = e / 9, i = (e % 9) / 3, j = e % 3;
  double    re = 0, im = 0;
#pragma unroll
  for (int k = 0; k < 3; ++k)
    {
      vh_cx x, y;
      if (m == 0) // R = A A^T
        x = vh_ld(A, i, k), y = vh_ld(A, j, k);
      else if (m == 1) // Q = A A^+
        x = vh_ld(A, i, k), y = vh_conj(vh_ld(A, j, k));
      else if (m == 2) // P = A^+ A
        x = vh_conj(vh_ld(A, k, i)), y = vh_ld(A, k, j);
      else // S = A^T A
        x = vh_ld(A, k, i), y = vh_ld(A, k, j);
      const vh_cx z = vh_cmul(x, y);
      re += z.re;
      im += z.im;
    }
  out2[0] = re;
  out2[1] = im;
}

// Component c of the bulk residual density g (c<9: Re g_{mu j}, c>=9: Im g_{mu j}).
VH_HD double vh_g_component(const double *A, const double *prod, int c, double alpha, const double *beta)
{
  const int   mu = (c % 9) / 3, j = c % 3;
  const vh_cx a = vh_ld(A, mu, j);
  const vh_cx T{prod[0] + prod[8] + prod[16], prod[1] + prod[9] + prod[17]}; // tr R
  const double Sr = prod[18 + 0] + prod[18 + 8] + prod[18 + 16];             // tr Q (real)
  vh_cx        g  = vh_cmul(T, vh_conj(a));
  g.re *= beta[0];
  g.im *= beta[0];
  g.re += beta[1] * Sr * a.re;
  g.im += beta[1] * Sr * a.im;
#pragma unroll
  for (int k = 0; k < 3; ++k)
    {
      const vh_cx akj = vh_ld(A, k, j);
      const vh_cx t3  = vh_cmul(vh_ldp(prod, 0, mu, k), vh_conj(akj));              // R A*
      const vh_cx t4  = vh_cmul(vh_ldp(prod, 1, mu, k), akj);                       // Q A
      const vh_cx t5  = vh_cmul(vh_conj(vh_ld(A, mu, k)), vh_ldp(prod, 3, k, j));   // A* S
      g.re += beta[2] * t3.re + beta[3] * t4.re + beta[4] * t5.re;
      g.im += beta[2] * t3.im + beta[3] * t4.im + beta[4] * t5.im;
    }
  const double re = alpha * a.re + 2.0 * g.re, im = alpha * a.im + 2.0 * g.im;
  return c < 9 ? re : im;
}

// Column d of H = dg/dA: out[c], c = 0..17.  Direction E = s e_nu e_k^T, s = 1 (d<9) or i (d>=9).
VH_HD void vh_hessian_column(const double *A, const double *prod, int d, double alpha, const double *beta, double *out)
{
  const int   nu = (d % 9) / 3, k = d % 3;
  const bool  imag = d >= 9;
  const vh_cx s  = imag ? vh_cx{0.0, 1.0} : vh_cx{1.0, 0.0};
  const vh_cx sc = vh_conj(s);
  const vh_cx T{prod[0] + prod[8] + prod[16], prod[1] + prod[9] + prod[17]};
  const double Sr = prod[18 + 0] + prod[18 + 8] + prod[18 + 16];
  const vh_cx  ank = vh_ld(A, nu, k);
  vh_cx        dT  = vh_cmul(s, ank); // d tr(AA^T) = 2 s A_{nu k}
  dT.re *= 2.0;
  dT.im *= 2.0;
  const double dS = 2.0 * A[d]; // d tr(AA^+) = 2 Re(conj(A_{nu k}) s)
  const double b1 = 2.0 * beta[0], b2 = 2.0 * beta[1], b3 = 2.0 * beta[2], b4 = 2.0 * beta[3], b5 = 2.0 * beta[4];
#pragma unroll
  for (int mu = 0; mu < 3; ++mu)
    {
      const vh_cx amk = vh_ld(A, mu, k);
#pragma unroll
      for (int j = 0; j < 3; ++j)
        {
          const vh_cx amj = vh_ld(A, mu, j), anj = vh_ld(A, nu, j);
          // dense part
          vh_cx       t  = vh_cmul(dT, vh_conj(amj));
          double      re = b1 * t.re + b2 * dS * amj.re, im = b1 * t.im + b2 * dS * amj.im;
          t  = vh_cmul(s, vh_cmul(amk, vh_conj(anj))); // A E^T A*
          re += b3 * t.re;
          im += b3 * t.im;
          t = vh_cmul(sc, vh_cmul(amk, anj)); // A E^+ A
          re += b4 * t.re;
          im += b4 * t.im;
          t = vh_cmul(s, vh_cmul(vh_conj(amk), anj)); // A* E^T A
          re += b5 * t.re;
          im += b5 * t.im;
          if (mu == nu)
            { // E A^T A*, E A^+ A, E* A^T A : only row nu
              const vh_cx p = vh_ldp(prod, 2, k, j);
              t  = vh_cmul(s, vh_conj(p));
              re += b3 * t.re;
              im += b3 * t.im;
              t = vh_cmul(s, p);
              re += b4 * t.re;
              im += b4 * t.im;
              t = vh_cmul(sc, vh_ldp(prod, 3, k, j));
              re += b5 * t.re;
              im += b5 * t.im;
            }
          if (j == k)
            { // A A^T E*, A A^+ E, A* A^T E : only column k
              const vh_cx q = vh_ldp(prod, 1, mu, nu);
              t = vh_cmul(sc, vh_ldp(prod, 0, mu, nu));
              re += b3 * t.re;
              im += b3 * t.im;
              t = vh_cmul(s, q);
              re += b4 * t.re;
              im += b4 * t.im;
              t = vh_cmul(s, vh_conj(q));
              re += b5 * t.re;
              im += b5 * t.im;
              if (mu == nu)
                { // alpha E + 2 beta1 T E* + 2 beta2 tr(AA^+) E
                  t = vh_cmul(T, sc);
                  re += alpha * s.re + b1 * t.re + b2 * Sr * s.re;
                  im += alpha * s.im + b1 * t.im + b2 * Sr * s.im;
                }
            }
          out[3 * mu + j]     = re;
          out[9 + 3 * mu + j] = im;
        }
    }
}

// Bulk free-energy density  alpha I0 + sum beta_k I_k   from the product table.
VH_HD double vh_bulk_energy(const double *prod, double alpha, const double *beta)
{
  const vh_cx  T{prod[0] + prod[8] + prod[16], prod[1] + prod[9] + prod[17]};
  const double Sr = prod[18 + 0] + prod[18 + 8] + prod[18 + 16];
  double       I3 = 0, I4 = 0, I5 = 0;
#pragma unroll
  for (int e = 0; e < 9; ++e)
    {
      const double rr = prod[2 * e], ri = prod[2 * e + 1], qr = prod[18 + 2 * e], qi = prod[18 + 2 * e + 1];
      I3 += rr * rr + ri * ri;
      I4 += qr * qr + qi * qi;
      I5 += qr * qr - qi * qi;
    }
  return alpha * Sr + beta[0] * (T.re * T.re + T.im * T.im) + beta[1] * Sr * Sr + beta[2] * I3 + beta[3] * I4 + beta[4] * I5;
}


// ---- table-driven evaluation of single Hessian entries (what the kernels use) -------------------------
// Every dense term of dg[E] is one entry of   Z1[x,y] = A_x conj(A_y)   or   Z2[x,y] = A_x A_y
// (x,y = 3*mu+j matrix positions), so H_q costs 162 complex products + a handful of adds per entry instead
// of ~7 complex products per entry.  ztab[324] = {Z1 (81 x (re,im)), Z2 (81 x (re,im))}.
VH_HD void vh_ztable_entry(const double *A, int e, double *ztab)
{
  const int   which = e / 81, xy = e - 81 * which, x = xy / 9, y = xy - 9 * x;
  const vh_cx ax{A[x], A[9 + x]}, ay{A[y], A[9 + y]};
  const vh_cx z = vh_cmul(ax, which ? ay : vh_conj(ay));
  ztab[2 * e]     = z.re;
  ztab[2 * e + 1] = z.im;
}

// H[c][d] from the tables; identical (to rounding) to vh_hessian_column(..., d, ...)[c].
VH_HD double vh_hessian_entry(const double *A, const double *prod, const double *ztab, int c, int d, double alpha,
                              const double *beta)
{
  const int  pc = c / 9, mu = (c % 9) / 3, j = c % 3;
  const bool pd = d >= 9;
  const int  nu = (d % 9) / 3, k = d % 3;
  const int  x_nk = 3 * nu + k, x_mj = 3 * mu + j, x_mk = 3 * mu + k, x_nj = 3 * nu + j;
  double     re = 0.0, im = 0.0;
  // G += w * s * z   and   G += w * conj(s) * z   with s = 1 (d<9) or i (d>=9)
#define VH_ADD_S(zr, zi, w)        \
  do                               \
    {                              \
      if (!pd)                     \
        {                          \
          re += (w) * (zr);        \
          im += (w) * (zi);        \
        }                          \
      else                         \
        {                          \
          re -= (w) * (zi);        \
          im += (w) * (zr);        \
        }                          \
    }                              \
  while (0)
#define VH_ADD_SB(zr, zi, w)       \
  do                               \
    {                              \
      if (!pd)                     \
        {                          \
          re += (w) * (zr);        \
          im += (w) * (zi);        \
        }                          \
      else                         \
        {                          \
          re += (w) * (zi);        \
          im -= (w) * (zr);        \
        }                          \
    }                              \
  while (0)
  const double  b1 = 2.0 * beta[0], b2 = 2.0 * beta[1], b3 = 2.0 * beta[2], b4 = 2.0 * beta[3], b5 = 2.0 * beta[4];
  const double *z1a = ztab + 2 * (9 * x_nk + x_mj);       // A_{nu k} conj(A_{mu j})
  const double *z1b = ztab + 2 * (9 * x_mk + x_nj);       // A_{mu k} conj(A_{nu j})
  const double *z2b = ztab + 162 + 2 * (9 * x_mk + x_nj); // A_{mu k} A_{nu j}
  VH_ADD_S(z1a[0], z1a[1], 2.0 * b1);
  re += 2.0 * b2 * A[d] * A[x_mj];
  im += 2.0 * b2 * A[d] * A[9 + x_mj];
  VH_ADD_S(z1b[0], z1b[1], b3);
  VH_ADD_SB(z2b[0], z2b[1], b4);
  VH_ADD_S(z1b[0], -z1b[1], b5);
  if (mu == nu)
    {
      const double *p = prod + 36 + 2 * (3 * k + j); // P = A^+ A
      const double *q = prod + 54 + 2 * (3 * k + j); // S = A^T A
      VH_ADD_S(p[0], -p[1], b3);
      VH_ADD_S(p[0], p[1], b4);
      VH_ADD_SB(q[0], q[1], b5);
    }
  if (j == k)
    {
      const double *r = prod + 0 + 2 * (3 * mu + nu);  // R = A A^T
      const double *q = prod + 18 + 2 * (3 * mu + nu); // Q = A A^+
      VH_ADD_SB(r[0], r[1], b3);
      VH_ADD_S(q[0], q[1], b4);
      VH_ADD_S(q[0], -q[1], b5);
      if (mu == nu)
        {
          const double Tr = prod[0] + prod[8] + prod[16], Ti = prod[1] + prod[9] + prod[17];
          const double Sr = prod[18 + 0] + prod[18 + 8] + prod[18 + 16];
          VH_ADD_S(alpha + b2 * Sr, 0.0, 1.0);
          VH_ADD_SB(Tr, Ti, b1);
        }
    }
#undef VH_ADD_S
#undef VH_ADD_SB
  return pc ? im : re;
}


// ---- per-thread term lists: H[c][d] = cst + wprod*tab[pa]*tab[pb] + sum_{k<16} w[k]*tab[off[k]] ---------------
// tab is the per-quadrature-point table {A[18] | prod[72] | Z1[162] | Z2[162]} (VH_TQ doubles).  For a fixed (c,d)
// every term of vh_hessian_entry reads ONE table entry with a fixed signed weight, so a thread that owns (c,d)
// sets the list up once and then spends 16 loads + 16 FMAs per quadrature point.
#define VH_TQ_A 0
#define VH_TQ_P 18
#define VH_TQ_Z 90
#define VH_TQ 414
#define VH_NTERMS 16

struct vh_terms
{
  int    off[VH_NTERMS];
  double w[VH_NTERMS];
  double cst, wprod;
  int    pa, pb;
};

// G += wt * (s or conj(s)) * (z or conj(z)), z = tab[o], tab[o+1]; only component pc of G is wanted.
VH_HD void vh_term(vh_terms &T, int &n, int o, double wt, bool sbar, bool cj, int pc, bool pd)
{
  // (sign, component) of z that lands in component pc of the result
  int    comp;
  double sg = 1.0;
  if (!pd)
    comp = pc; // s = 1: re -> re, im -> im
  else if (!sbar)
    { // s = i: re = -zi, im = +zr
      comp = 1 - pc;
      if (pc == 0)
        sg = -1.0;
    }
  else
    { // conj(s) = -i: re = +zi, im = -zr
      comp = 1 - pc;
      if (pc == 1)
        sg = -1.0;
    }
  if (cj && comp == 1)
    sg = -sg;
  T.off[n] = o + comp;
  T.w[n]   = wt * sg;
  ++n;
}

VH_HD void vh_entry_terms(int c, int d, double alpha, const double *beta, vh_terms &T)
{
  const int  pc = c / 9, mu = (c % 9) / 3, j = c % 3;
  const bool pd = d >= 9;
  const int  nu = (d % 9) / 3, k = d % 3;
  const int  x_nk = 3 * nu + k, x_mj = 3 * mu + j, x_mk = 3 * mu + k, x_nj = 3 * nu + j;
  const double b1 = 2.0 * beta[0], b2 = 2.0 * beta[1], b3 = 2.0 * beta[2], b4 = 2.0 * beta[3], b5 = 2.0 * beta[4];
  int        n = 0;
  const int  Z1 = VH_TQ_Z, Z2 = VH_TQ_Z + 162, P = VH_TQ_P;
  vh_term(T, n, Z1 + 2 * (9 * x_nk + x_mj), 2.0 * b1, false, false, pc, pd);
  vh_term(T, n, Z1 + 2 * (9 * x_mk + x_nj), b3, false, false, pc, pd);
  vh_term(T, n, Z2 + 2 * (9 * x_mk + x_nj), b4, true, false, pc, pd);
  vh_term(T, n, Z1 + 2 * (9 * x_mk + x_nj), b5, false, true, pc, pd);
  const bool rowm = mu == nu, colm = j == k;
  // row nu only:  E A^T A* (conj P), E A^+ A (P), E* A^T A (S)
  vh_term(T, n, P + 36 + 2 * (3 * k + j), rowm ? b3 : 0.0, false, true, pc, pd);
  vh_term(T, n, P + 36 + 2 * (3 * k + j), rowm ? b4 : 0.0, false, false, pc, pd);
  vh_term(T, n, P + 54 + 2 * (3 * k + j), rowm ? b5 : 0.0, true, false, pc, pd);
  // column k only:  A A^T E* (R), A A^+ E (Q), A* A^T E (conj Q)
  vh_term(T, n, P + 0 + 2 * (3 * mu + nu), colm ? b3 : 0.0, true, false, pc, pd);
  vh_term(T, n, P + 18 + 2 * (3 * mu + nu), colm ? b4 : 0.0, false, false, pc, pd);
  vh_term(T, n, P + 18 + 2 * (3 * mu + nu), colm ? b5 : 0.0, false, true, pc, pd);
  // both:  (alpha + 2 beta2 tr(AA^+)) E  and  2 beta1 tr(AA^T) E*
  const bool both = rowm && colm;
  for (int i = 0; i < 3; ++i)
    vh_term(T, n, P + 18 + 8 * i, both ? b2 : 0.0, false, false, pc, pd); // Q_ii (imaginary parts are 0 up to rounding)
  for (int i = 0; i < 3; ++i)
    vh_term(T, n, P + 0 + 8 * i, both ? b1 : 0.0, true, false, pc, pd); // R_ii
  // alpha * s lands in component pc only if pc == pd
  T.cst   = (both && (pc == (pd ? 1 : 0))) ? alpha : 0.0;
  T.wprod = 2.0 * b2; // 4 beta2 a_d * (component pc of A_{mu j}) = 2 b2 a_c a_d
  T.pa    = VH_TQ_A + c;
  T.pb    = VH_TQ_A + d;
}

VH_HD double vh_entry_eval(const double *tab, const vh_terms &T)
{
  double v = fma(T.wprod * tab[T.pa], tab[T.pb], T.cst);
#pragma unroll
  for (int k = 0; k < VH_NTERMS; ++k)
    v = fma(T.w[k], tab[T.off[k]], v);
  return v;
}

// ---- register-resident evaluation (what k_points_q1 uses): one thread owns one quadrature point ---------------------
// The thread keeps A (18 doubles) and the UNIQUE entries of the four products (R, S complex symmetric: 6 complex each;
// Q, P Hermitian: 6 real parts + 3 imaginary parts each; 42 doubles instead of 72) in registers and evaluates the packed
// entries (c,d) with c, d compile-time constants after unrolling, so only the terms that are not structurally zero are
// executed (7.7 on average instead of the 17 table terms of vh_entry_eval) and there is no index arithmetic at all.
struct vh_prods
{
  double Rr[6], Ri[6]; // R = A A^T
  double Qr[6], Qi[3]; // Q = A A^+   (Q_ji = conj Q_ij)
  double Pr[6], Pi[3]; // P = A^+ A   (P_ji = conj P_ij)
  double Sr[6], Si[6]; // S = A^T A
};
// (0,0)=0 (0,1)=1 (0,2)=2 (1,1)=3 (1,2)=4 (2,2)=5
VH_HD int vh_sym6(int i, int j) { return i <= j ? (i == 0 ? j : (i == 1 ? 2 + j : 5)) : (j == 0 ? i : (j == 1 ? 2 + i : 5)); }
VH_HD void vh_prods_compute(const double *a, vh_prods &p)
{
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = i; j < 3; ++j)
      {
        double rr = 0, ri = 0, qr = 0, qi = 0, pr = 0, pi = 0, sr = 0, si = 0;
#pragma unroll
        for (int k = 0; k < 3; ++k)
          {
            const double ux = a[3 * i + k], vx = a[9 + 3 * i + k], uy = a[3 * j + k], vy = a[9 + 3 * j + k]; // A_ik, A_jk
            rr += ux * uy - vx * vy;
            ri += ux * vy + vx * uy;
            qr += ux * uy + vx * vy;
            qi += vx * uy - ux * vy;
            const double us = a[3 * k + i], vs = a[9 + 3 * k + i], ut = a[3 * k + j], vt = a[9 + 3 * k + j]; // A_ki, A_kj
            pr += us * ut + vs * vt;
            pi += us * vt - vs * ut;
            sr += us * ut - vs * vt;
            si += us * vt + vs * ut;
          }
        const int e = vh_sym6(i, j);
        p.Rr[e] = rr, p.Ri[e] = ri, p.Qr[e] = qr, p.Pr[e] = pr, p.Sr[e] = sr, p.Si[e] = si;
        if (i < j)
          p.Qi[i + j - 1] = qi, p.Pi[i + j - 1] = pi;
      }
}
VH_HD vh_cx vh_getR(const vh_prods &p, int i, int j) { return vh_cx{p.Rr[vh_sym6(i, j)], p.Ri[vh_sym6(i, j)]}; }
VH_HD vh_cx vh_getS(const vh_prods &p, int i, int j) { return vh_cx{p.Sr[vh_sym6(i, j)], p.Si[vh_sym6(i, j)]}; }
VH_HD vh_cx vh_getQ(const vh_prods &p, int i, int j)
{
  return vh_cx{p.Qr[vh_sym6(i, j)], i == j ? 0.0 : (i < j ? p.Qi[i + j - 1] : -p.Qi[i + j - 1])};
}
VH_HD vh_cx vh_getP(const vh_prods &p, int i, int j)
{
  return vh_cx{p.Pr[vh_sym6(i, j)], i == j ? 0.0 : (i < j ? p.Pi[i + j - 1] : -p.Pi[i + j - 1])};
}

// weights of the Hessian terms; b_k = 2 beta_k
struct vh_hweights
{
  double alpha, b1x2, b2x2, b3, b4, b5, b35p, b35m, b34p, b43m, b45p, b45m, b1, b2;
};
VH_HD vh_hweights vh_make_hweights(double alpha, const double *beta)
{
  const double b1 = 2.0 * beta[0], b2 = 2.0 * beta[1], b3 = 2.0 * beta[2], b4 = 2.0 * beta[3], b5 = 2.0 * beta[4];
  return vh_hweights{alpha, 2.0 * b1, 2.0 * b2, b3, b4, b5, b3 + b5, b3 - b5, b3 + b4, b4 - b3, b4 + b5, b4 - b5, b1, b2};
}
// acc + component pc of  (s or conj s) * (wr zr + i wi zi),  s = 1 (pd = 0) or i (pd = 1)
VH_HD double vh_acc_s(double acc, bool pd, bool pc, bool sbar, double wr, double wi, double zr, double zi)
{
  if (!pd)
    return pc ? fma(wi, zi, acc) : fma(wr, zr, acc);
  if (!sbar)
    return pc ? fma(wr, zr, acc) : fma(-wi, zi, acc);
  return pc ? fma(-wr, zr, acc) : fma(wi, zi, acc);
}
// per-point scalars of the "row == column" terms: c0 = alpha + 2 beta2 tr Q, (tr, ti) = 2 beta1 tr R
struct vh_hdiag
{
  double c0, tr, ti;
};
VH_HD vh_hdiag vh_make_hdiag(const vh_prods &p, const vh_hweights &w)
{
  return vh_hdiag{fma(w.b2, p.Qr[0] + p.Qr[3] + p.Qr[5], w.alpha), w.b1 * (p.Rr[0] + p.Rr[3] + p.Rr[5]),
                  w.b1 * (p.Ri[0] + p.Ri[3] + p.Ri[5])};
}
// H[c][d]; same terms as vh_hessian_entry.  Meant to be called with c, d known at compile time.
VH_HD double vh_h_entry(const double *a, const vh_prods &p, const vh_hweights &w, const vh_hdiag &hd, int c, int d)
{
  const bool pc = c >= 9, pd = d >= 9;
  const int  mu = (c % 9) / 3, j = c % 3, nu = (d % 9) / 3, k = d % 3;
  const int  X = 3 * nu + k, Y = 3 * mu + j, M = 3 * mu + k, N = 3 * nu + j;
  // z1a = A_X conj(A_Y), z1b = A_M conj(A_N), z2b = A_M A_N : only the component that lands in pc is live
  double v = vh_acc_s(0.0, pd, pc, false, w.b1x2, w.b1x2, fma(a[X], a[Y], a[9 + X] * a[9 + Y]), fma(a[9 + X], a[Y], -(a[X] * a[9 + Y])));
  v        = fma(w.b2x2 * a[c], a[d], v);
  v = vh_acc_s(v, pd, pc, false, w.b35p, w.b35m, fma(a[M], a[N], a[9 + M] * a[9 + N]), fma(a[9 + M], a[N], -(a[M] * a[9 + N])));
  v = vh_acc_s(v, pd, pc, true, w.b4, w.b4, fma(a[M], a[N], -(a[9 + M] * a[9 + N])), fma(a[M], a[9 + N], a[9 + M] * a[N]));
  if (mu == nu)
    { // row nu only: s (b3 conj P + b4 P)_kj + b5 conj(s) S_kj
      const vh_cx P = vh_getP(p, k, j), S = vh_getS(p, k, j);
      v = vh_acc_s(v, pd, pc, false, w.b34p, w.b43m, P.re, P.im);
      v = vh_acc_s(v, pd, pc, true, w.b5, w.b5, S.re, S.im);
    }
  if (j == k)
    { // column k only: b3 conj(s) R + s (b4 Q + b5 conj Q), entry (mu,nu)
      const vh_cx R = vh_getR(p, mu, nu), Q = vh_getQ(p, mu, nu);
      v = vh_acc_s(v, pd, pc, true, w.b3, w.b3, R.re, R.im);
      v = vh_acc_s(v, pd, pc, false, w.b45p, w.b45m, Q.re, Q.im);
      if (mu == nu)
        { // s (alpha + 2 beta2 tr Q) + 2 beta1 conj(s) tr R
          if (pc == pd)
            v += pd ? hd.c0 - hd.tr : hd.c0 + hd.tr;
          else
            v += hd.ti;
        }
    }
  return v;
}
// all 18 components of g = alpha A + 2 sum_k beta_k G_k from the unique products
VH_HD void vh_g_all(const double *a, const vh_prods &p, const vh_hweights &w, double *g)
{
  const double Tr = p.Rr[0] + p.Rr[3] + p.Rr[5], Ti = p.Ri[0] + p.Ri[3] + p.Ri[5], Sq = p.Qr[0] + p.Qr[3] + p.Qr[5];
  const double c0 = fma(w.b2, Sq, w.alpha);
#pragma unroll
  for (int mu = 0; mu < 3; ++mu)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      {
        const double u = a[3 * mu + j], v = a[9 + 3 * mu + j];
        // alpha a + b2 trQ a + b1 T conj(a)
        double re = fma(w.b1, Tr * u + Ti * v, c0 * u), im = fma(w.b1, Ti * u - Tr * v, c0 * v);
#pragma unroll
        for (int k = 0; k < 3; ++k)
          {
            const vh_cx akj = vh_ld(a, k, j), amk = vh_ld(a, mu, k);
            const vh_cx t3 = vh_cmul(vh_getR(p, mu, k), vh_conj(akj)), t4 = vh_cmul(vh_getQ(p, mu, k), akj),
                        t5 = vh_cmul(vh_conj(amk), vh_getS(p, k, j));
            re += w.b3 * t3.re + w.b4 * t4.re + w.b5 * t5.re;
            im += w.b3 * t3.im + w.b4 * t4.im + w.b5 * t5.im;
          }
        g[3 * mu + j]     = re;
        g[9 + 3 * mu + j] = im;
      }
}
VH_HD double vh_bulk_energy_u(const vh_prods &p, double alpha, const double *beta)
{
  const double Tr = p.Rr[0] + p.Rr[3] + p.Rr[5], Ti = p.Ri[0] + p.Ri[3] + p.Ri[5], Sq = p.Qr[0] + p.Qr[3] + p.Qr[5];
  double       I3 = 0, I4 = 0, I5 = 0;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = i; j < 3; ++j)
      {
        const int    e = vh_sym6(i, j);
        const double m = i == j ? 1.0 : 2.0, qi = i == j ? 0.0 : p.Qi[i + j - 1];
        I3 += m * (p.Rr[e] * p.Rr[e] + p.Ri[e] * p.Ri[e]);
        I4 += m * (p.Qr[e] * p.Qr[e] + qi * qi);
        I5 += m * (p.Qr[e] * p.Qr[e] - qi * qi);
      }
  return alpha * Sq + beta[0] * (Tr * Tr + Ti * Ti) + beta[1] * Sq * Sq + beta[2] * I3 + beta[3] * I4 + beta[4] * I5;
}

// Storage of a Q1 cell's eight packed H_q (k_points_q1 -> row-owner kernels): [pair p = e/2][slot] double2 with
// slot = q XOR (p & 7).  The eight quadrature points of one pair fill one 128-byte line (the producer's 8 lanes of a
// cell write it with one 16-byte store each); the XOR keeps the consumer's per-entry reads free of bank conflicts.
VH_HD int vh_hq8_index(int q, int e) { return (((e >> 1) << 3) + (q ^ ((e >> 1) & 7))) * 2 + (e & 1); }

// Packed layout of a symmetric 18x18 ("P180"): row c stores the entries (c,d) for d = 2*(c/2) .. 17, so every row
// starts at an EVEN column and every 16-byte pair (d, d+1) of the packed array belongs to one row and pairs with an
// aligned 16-byte pair of the vector in the SpMV.  For odd c the first stored entry (c, c-1) is a dummy that is always 0.
// 180 doubles per matrix (171 unique + 9 dummies).
#define VH_SYMP 180
VH_HD int vh_sym_rowstart(int c)
{
  const int m = c >> 1;
  return (c & 1) ? 36 * m - 2 * m * m + 18 : 38 * m - 2 * m * m;
}
// index of (c,d), c <= d
VH_HD int vh_sym_index(int c, int d) { return vh_sym_rowstart(c) + d - 2 * (c >> 1); }
// t += Sym(P) z for one packed block whose 90 16-byte pieces are fetched by ld(piece, v0, v1): piece p of row c holds
// the entries (c,d), (c,d+1) with d = 2*(c/2) + 2k.  Fully unrolled: every index is a compile-time constant, so z and t
// stay in registers (matrix-free operator apply, k_apply_cells).
template <class LD>
VH_HD void vh_sym_matvec(LD ld, const double *z, double *t)
{
#pragma unroll
  for (int c = 0; c < 18; ++c)
#pragma unroll
    for (int d = 2 * (c >> 1); d < 18; d += 2)
      {
        double v0, v1;
        ld(vh_sym_index(c, d) >> 1, v0, v1);
        if (d > c)
          { // off-diagonal entry (c,d): feeds rows c and d
            t[c] += v0 * z[d];
            t[d] += v0 * z[c];
          }
        else if (d == c)
          t[c] += v0 * z[c];
        // d == c - 1: the zero dummy of an odd row
        if (d + 1 > c)
          {
            t[c] += v1 * z[d + 1];
            t[d + 1] += v1 * z[c];
          }
        else
          t[c] += v1 * z[c]; // d + 1 == c: the diagonal entry of an odd row
      }
}

// ---------------------------------------------------------------------------------------------------------------
// Table-free H z (roadmap of the matrix-free operator apply): the directional derivative of
// g = alpha A + 2 sum_k beta_k G_k at A in the direction Z (z[18] in the layout of A),
//   dG1 = 2 tr(A Z^T) A* + tr(AA^T) Z*          dG2 = 2 Re tr(A Z^+) A + tr(AA^+) Z
//   dG3 = Z (A^T A*) + (A Z^T) A* + (A A^T) Z*   dG4 = Z (A^+ A) + (A Z^+) A + (A A^+) Z
//   dG5 = Z* (A^T A) + (A Z^+)* A + (A A^+)* Z
// from A and the four products R = AA^T, Q = AA^+, P = A^+A, S = A^T A (prod[72], vh_product_entry): eight 3x3 complex
// products, no 18x18 table.  Equals the symmetric H of vh_hessian_column / vh_h_entry applied to z (host-tested).
VH_HD void vh_hessian_apply(const double *A, const double *prod, const double *z, double alpha, const double *beta, double *out)
{
  vh_cx a[3][3], zz[3][3], M1[3][3], M2[3][3], WL[3][3], WR[3][3];
  vh_cx dT{0.0, 0.0};
  double dS = 0.0;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      {
        a[i][j]  = vh_ld(A, i, j);
        zz[i][j] = vh_ld(z, i, j);
        const vh_cx t = vh_cmul(a[i][j], zz[i][j]);
        dT.re += 2.0 * t.re;
        dT.im += 2.0 * t.im;
        dS += 2.0 * (a[i][j].re * zz[i][j].re + a[i][j].im * zz[i][j].im);
      }
  const vh_cx  T{prod[0] + prod[8] + prod[16], prod[1] + prod[9] + prod[17]}; // tr R
  const double Sr = prod[18 + 0] + prod[18 + 8] + prod[18 + 16];             // tr Q (real)
  // M1 = A Z^T, M2 = A Z^+
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int k = 0; k < 3; ++k)
      {
        vh_cx m1{0.0, 0.0}, m2{0.0, 0.0};
#pragma unroll
        for (int j = 0; j < 3; ++j)
          {
            const vh_cx t1 = vh_cmul(a[i][j], zz[k][j]), t2 = vh_cmul(a[i][j], vh_conj(zz[k][j]));
            m1.re += t1.re, m1.im += t1.im, m2.re += t2.re, m2.im += t2.im;
          }
        M1[i][k] = m1;
        M2[i][k] = m2;
      }
  // z-independent combinations: WL = beta4 Q + beta5 Q* (multiplies Z from the left), WR = beta3 P* + beta4 P (from the right)
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      {
        const vh_cx q = vh_ldp(prod, 1, i, j), p = vh_ldp(prod, 2, i, j);
        WL[i][j] = vh_cx{(beta[3] + beta[4]) * q.re, (beta[3] - beta[4]) * q.im};
        WR[i][j] = vh_cx{(beta[2] + beta[3]) * p.re, (beta[3] - beta[2]) * p.im};
      }
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      {
        // beta1 dG1 + beta2 dG2
        vh_cx t = vh_cmul(dT, vh_conj(a[i][j]));
        vh_cx u = vh_cmul(T, vh_conj(zz[i][j]));
        double re = beta[0] * (t.re + u.re) + beta[1] * (dS * a[i][j].re + Sr * zz[i][j].re);
        double im = beta[0] * (t.im + u.im) + beta[1] * (dS * a[i][j].im + Sr * zz[i][j].im);
#pragma unroll
        for (int k = 0; k < 3; ++k)
          {
            // left products: WL Z + beta3 R Z*
            t = vh_cmul(WL[i][k], zz[k][j]);
            u = vh_cmul(vh_ldp(prod, 0, i, k), vh_conj(zz[k][j]));
            re += t.re + beta[2] * u.re;
            im += t.im + beta[2] * u.im;
            // right products: Z WR + beta5 Z* S
            t = vh_cmul(zz[i][k], WR[k][j]);
            u = vh_cmul(vh_conj(zz[i][k]), vh_ldp(prod, 3, k, j));
            re += t.re + beta[4] * u.re;
            im += t.im + beta[4] * u.im;
            // middle products: beta3 M1 A* + (beta4 M2 + beta5 M2*) A
            t = vh_cmul(M1[i][k], vh_conj(a[k][j]));
            const vh_cx m{(beta[3] + beta[4]) * M2[i][k].re, (beta[3] - beta[4]) * M2[i][k].im};
            u = vh_cmul(m, a[k][j]);
            re += beta[2] * t.re + u.re;
            im += beta[2] * t.im + u.im;
          }
        out[3 * i + j]     = alpha * zz[i][j].re + 2.0 * re;
        out[9 + 3 * i + j] = alpha * zz[i][j].im + 2.0 * im;
      }
}

// fills c[e], d[e] for e in [0,180); dummies have d[e] = c[e] - 1 (i.e. d < c)
inline void vh_sym_tables(unsigned char *tc, unsigned char *td)
{
  int e = 0;
  for (int c = 0; c < 18; ++c)
    for (int d = 2 * (c >> 1); d < 18; ++d, ++e)
      {
        tc[e] = (unsigned char)c;
        td[e] = (unsigned char)d;
      }
}

#endif
