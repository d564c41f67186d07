// Test-only host build of the product's __host__ __device__ pointwise math (csrc/vh_pointwise.cuh),
// so the closed forms the CUDA kernels use are checked against the oracle on the CPU-only container.
// Never part of the product: the product evaluates these functions on the device only.
#include "../../verkko-hem-repo_b200/csrc/vh_pointwise.cuh"

extern "C" void vht_pointwise(const double *A, const double *coef, double *g18, double *H324, double *f)
{
  double prod[72];
  for (int e = 0; e < 36; ++e)
    vh_product_entry(A, e, prod + 2 * e);
  const double alpha = coef[3], *beta = coef + 4;
  for (int c = 0; c < 18; ++c)
    g18[c] = vh_g_component(A, prod, c, alpha, beta);
  for (int d = 0; d < 18; ++d)
    {
      double col[18];
      vh_hessian_column(A, prod, d, alpha, beta, col);
      for (int c = 0; c < 18; ++c)
        H324[18 * c + d] = col[c];
    }
  *f = vh_bulk_energy(prod, alpha, beta);
}
// same outputs through the table-driven single-entry evaluation the kernels use
extern "C" void vht_hessian_entries(const double *A, const double *coef, double *H324)
{
  double prod[72], ztab[324];
  for (int e = 0; e < 36; ++e)
    vh_product_entry(A, e, prod + 2 * e);
  for (int e = 0; e < 162; ++e)
    vh_ztable_entry(A, e, ztab);
  for (int c = 0; c < 18; ++c)
    for (int d = 0; d < 18; ++d)
      H324[18 * c + d] = vh_hessian_entry(A, prod, ztab, c, d, coef[3], coef + 4);
}
// ... and through the per-thread term lists (16 loads + 16 FMAs per entry and quadrature point)
extern "C" void vht_hessian_terms(const double *A, const double *coef, double *H324)
{
  double tab[VH_TQ];
  for (int i = 0; i < 18; ++i)
    tab[VH_TQ_A + i] = A[i];
  for (int e = 0; e < 36; ++e)
    vh_product_entry(A, e, tab + VH_TQ_P + 2 * e);
  for (int e = 0; e < 162; ++e)
    vh_ztable_entry(A, e, tab + VH_TQ_Z);
  for (int c = 0; c < 18; ++c)
    for (int d = 0; d < 18; ++d)
      {
        vh_terms T;
        vh_entry_terms(c, d, coef[3], coef + 4, T);
        H324[18 * c + d] = vh_entry_eval(tab, T);
      }
}
extern "C" int vht_sym_index(int c, int d) { return vh_sym_index(c, d); }
// ... and through the register-resident evaluation of k_points_q1 (unique products, compile-time entry formulas)
extern "C" void vht_points_q1(const double *A, const double *coef, double *g18, double *H324, double *f)
{
  vh_prods p;
  vh_prods_compute(A, p);
  const vh_hweights w  = vh_make_hweights(coef[3], coef + 4);
  const vh_hdiag    hd = vh_make_hdiag(p, w);
  vh_g_all(A, p, w, g18);
  for (int c = 0; c < 18; ++c)
    for (int d = 0; d < 18; ++d)
      H324[18 * c + d] = vh_h_entry(A, p, w, hd, c, d);
  *f = vh_bulk_energy_u(p, coef[3], coef + 4);
}
extern "C" int vht_hq8_index(int q, int e) { return vh_hq8_index(q, e); }
// the packed symmetric mat-vec of the matrix-free operator apply (k_apply_cells), on a block stored in the Q1 cell layout
// (hq: one cell's 1440 doubles, vh_hq8_index) or in the plain packed layout (q < 0: hq is 180 doubles)
extern "C" void vht_sym_matvec(const double *hq, int q, const double *z, double *t)
{
  for (int c = 0; c < 18; ++c)
    t[c] = 0.0;
  vh_sym_matvec(
    [&](int p, double &v0, double &v1) {
      if (q < 0)
        {
          v0 = hq[2 * p];
          v1 = hq[2 * p + 1];
        }
      else
        {
          v0 = hq[((p << 3) + (q ^ (p & 7))) * 2];
          v1 = hq[((p << 3) + (q ^ (p & 7))) * 2 + 1];
        }
    },
    z, t);
}
// table-free H z: the directional derivative of g at A in the direction z (roadmap of the matrix-free apply)
extern "C" void vht_hessian_apply(const double *A, const double *coef, const double *z, double *out18)
{
  double prod[72];
  for (int e = 0; e < 36; ++e)
    vh_product_entry(A, e, prod + 2 * e);
  vh_hessian_apply(A, prod, z, coef[3], coef + 4, out18);
}
