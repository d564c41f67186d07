// Jacobian / residual assembly of the femgl Newton step on sm_100a.
//
// Replaces FemGL::assemble_system (/root/reference/femgl/src/assemble.cc:108-372) and
// FemGL::compute_residual (/root/reference/femgl/src/residual.cc:109-297), including the constrained
// scatter AffineConstraints::distribute_local_to_global (call sites assemble.cc:356-361, residual.cc:287-289).
//
// Three kernels instead of the reference's (cell, q, i, j) loop over 3x3 FullMatrix products:
//   1. k_pointwise    one CTA per cell: gather the cell's DoFs, interpolate A and grad A at the quadrature
//                     points, evaluate g_q (18) and the symmetric bulk Hessian H_q (171 unique entries) in closed
//                     form (vh_pointwise.cuh), store H_q, and reduce the cell rhs / cell diagonal / cell energy.
//   2. k_rows_fast_q1 ROW-OWNER assembly: one CTA per matrix block row.  Each 18x18 block of the row is
//                     accumulated in registers over the <= 8 incident cells x 8 quadrature points and written
//                     exactly once with 16-byte stores: no atomics, no memset (write-once traffic = the matrix).
//                     Gradient (K1, K2+K3) and Robin-face terms are geometry-only and added at write time.
//                     Component-masked Dirichlet DoFs follow deal.II's rule (zero row/column, |a_ii| on the diagonal).
//   3. k_cells_slow   general constrained scatter (hanging nodes / any topology / Q2): one CTA per cell,
//                     constraint lines resolved per entry, atomics into the rows that kernel 2 does not own.
#include "vh_internal.h"
#include "vh_pointwise.cuh"

#include <cstdio>

__constant__ double  c_W1[512];   // Q1: w_q N_a(q) N_b(q), index (a*8+b)*8+q
__constant__ uint8_t c_symc[VH_SYMP], c_symd[VH_SYMP];

namespace
{
__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
    v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// sum over the whole block; result valid in thread 0.  s_red needs 32 doubles.
__device__ __forceinline__ double block_sum(double v, double *s_red)
{
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0)
    s_red[wid] = v;
  __syncthreads();
  double r = 0;
  if (wid == 0)
    {
      r = lane < nw ? s_red[lane] : 0.0;
      r = warp_sum(r);
    }
  return r;
}

// ------------------------------------------------------------------------------------------------
// 1. pointwise kernel
// ------------------------------------------------------------------------------------------------
template <int NN, int NQ>
__global__ void __launch_bounds__(NQ == 8 ? 160 : 512) k_pointwise(const int32_t *__restrict__ cell_nodes, const double *__restrict__ cell_h,
                            const uint32_t *__restrict__ cell_faces, const uint8_t *__restrict__ cell_owned,
                            const double *__restrict__ x, VhTables tab, VhCoef cf, int want_h, int want_e,
                            double *__restrict__ Hq, double *__restrict__ Rc, double *__restrict__ Dc,
                            double *__restrict__ avgD, double *__restrict__ Ec)
{
  extern __shared__ double sm[];
  double *sU  = sm;               // [NN*18]
  double *sA  = sU + NN * 18;     // [NQ*18]
  double *sdA = sA + NQ * 18;     // [NQ*18*3]
  double *sP  = sdA + NQ * 54;    // [NQ*72]
  double *sg  = sP + NQ * 72;     // [NQ*18]
  double *sN  = sg + NQ * 18;     // [NN*NQ]
  double *sdN = sN + NN * NQ;     // [NN*NQ*3]
  double *swq = sdN + NN * NQ * 3; // [NQ]
  double *sred = swq + NQ;        // [32]
  double *sH  = sred + 32;        // [NQ*324] (only touched when want_h)

  const int     t    = threadIdx.x;
  const int64_t cell = blockIdx.x;
  for (int i = t; i < NN * NQ; i += blockDim.x)
    sN[i] = tab.N[i];
  for (int i = t; i < NN * NQ * 3; i += blockDim.x)
    sdN[i] = tab.dN[i];
  if (t < NQ)
    swq[t] = tab.wq[t];
  for (int i = t; i < NN * 18; i += blockDim.x)
    {
      const int a = i / 18, c = i - 18 * a;
      sU[i]       = x[18 * (int64_t)cell_nodes[cell * NN + a] + c];
    }
  const double h0 = cell_h[4 * cell], h1 = cell_h[4 * cell + 1], h2 = cell_h[4 * cell + 2], vol = cell_h[4 * cell + 3];
  const double ih[3] = {1.0 / h0, 1.0 / h1, 1.0 / h2};
  __syncthreads();

  // FE interpolation of the state and its gradient (s_vector2matrix.cc:154-162, 203-213)
  if (t < 18 * NQ)
    {
      const int q = t / 18, c = t - 18 * q;
      double    A = 0, d0 = 0, d1 = 0, d2 = 0;
#pragma unroll 4
      for (int a = 0; a < NN; ++a)
        {
          const double u = sU[a * 18 + c];
          A += sN[a * NQ + q] * u;
          d0 += sdN[(a * NQ + q) * 3 + 0] * u;
          d1 += sdN[(a * NQ + q) * 3 + 1] * u;
          d2 += sdN[(a * NQ + q) * 3 + 2] * u;
        }
      sA[t]           = A;
      sdA[3 * t + 0]  = d0 * ih[0];
      sdA[3 * t + 1]  = d1 * ih[1];
      sdA[3 * t + 2]  = d2 * ih[2];
    }
  __syncthreads();
  for (int i = t; i < NQ * 36; i += blockDim.x)
    {
      const int q = i / 36, e = i - 36 * q;
      vh_product_entry(sA + q * 18, e, sP + q * 72 + 2 * e);
    }
  __syncthreads();
  if (t < 18 * NQ)
    {
      const int q = t / 18, c = t - 18 * q;
      sg[t]       = vh_g_component(sA + q * 18, sP + q * 72, c, cf.alpha, cf.beta);
      if (want_h)
        {
          double col[18];
          vh_hessian_column(sA + q * 18, sP + q * 72, c, cf.alpha, cf.beta, col);
#pragma unroll
          for (int cc = 0; cc < 18; ++cc)
            sH[q * VH_BLK + cc * 18 + c] = col[cc];
        }
    }
  __syncthreads();
  if (want_h)
    {
      double *dst = Hq + cell * (int64_t)(NQ * VH_SYMP);
      for (int i = t; i < NQ * VH_SYMP; i += blockDim.x)
        {
          const int q = i / VH_SYMP, s = i - VH_SYMP * q;
          double    v = 0.0;
          if (s < VH_SYM)
            v = sH[q * VH_BLK + c_symc[s] * 18 + c_symd[s]];
          dst[i] = v;
        }
    }

  // cell rhs (assemble.cc:257-276) and cell-matrix diagonal
  const uint32_t faces = cell_faces[cell];
  const bool     robin = (cf.bt < 1e10) && faces != 0u;
  double         absd  = 0.0;
  if (t < 18 * NN)
    {
      const int a = t / 18, c = t - 18 * a, pm = c / 3, xc = c - 3 * pm;
      double    r = 0.0, dg = 0.0;
      for (int q = 0; q < NQ; ++q)
        {
          const double JxW = swq[q] * vol, Na = sN[a * NQ + q];
          const double gx = sdN[(a * NQ + q) * 3 + 0] * ih[0], gy = sdN[(a * NQ + q) * 3 + 1] * ih[1],
                       gz = sdN[(a * NQ + q) * 3 + 2] * ih[2];
          const double *dA  = sdA + 3 * (q * 18 + c);
          const double *dAr = sdA + 3 * (q * 18 + 3 * pm);
          const double  div = dAr[0] + dAr[4] + dAr[8];
          const double  gxc = xc == 0 ? gx : (xc == 1 ? gy : gz);
          r += JxW * (Na * sg[q * 18 + c] + cf.K1 * (gx * dA[0] + gy * dA[1] + gz * dA[2]) + cf.K23 * gxc * div);
          if (want_h)
            dg += JxW * Na * Na * sH[q * VH_BLK + c * 18 + c];
        }
      if (want_h)
        {
          const double *G = tab.Gref + (size_t)(a * NN + a) * 9;
          dg += vol * (cf.K1 * (G[0] * ih[0] * ih[0] + G[4] * ih[1] * ih[1] + G[8] * ih[2] * ih[2]) +
                       cf.K23 * G[4 * xc] * ih[xc] * ih[xc]);
        }
      if (robin)
        for (int f = 0; f < 6; ++f)
          {
            const int bid = (faces >> (4 * f)) & 15u;
            if (bid < 2 || bid > 4 || xc == bid - 2)
              continue;
            const double  hn   = f / 2 == 0 ? h0 : (f / 2 == 1 ? h1 : h2);
            const double  s    = cf.K1 / cf.bt * (vol / hn);
            const double *M    = tab.Mf + (size_t)(f * NN + a) * NN;
            double        accf = 0.0;
            for (int b = 0; b < NN; ++b)
              accf += M[b] * sU[b * 18 + c];
            r += s * accf;
            dg += s * M[a];
          }
      Rc[cell * (int64_t)(18 * NN) + t] = -r;
      if (want_h)
        {
          Dc[cell * (int64_t)(18 * NN) + t] = dg;
          absd                              = fabs(dg);
        }
    }
  if (want_h)
    {
      const double s = block_sum(absd, sred);
      if (t == 0)
        avgD[cell] = s / (double)(18 * NN);
    }
  if (want_e)
    { // GL functional, SURVEY.md A.1 (only locally owned cells count; ghost cells are assembled redundantly)
      double e = 0.0;
      if (t < 18 * NQ && cell_owned[cell])
        {
          const int     q = t / 18, c = t - 18 * q;
          const double  JxW = swq[q] * vol;
          const double *dA  = sdA + 3 * t;
          e                 = cf.K1 * (dA[0] * dA[0] + dA[1] * dA[1] + dA[2] * dA[2]);
          if (c < 6)
            {
              const double *dAr = sdA + 3 * (q * 18 + 3 * c);
              const double  div = dAr[0] + dAr[4] + dAr[8];
              e += cf.K23 * div * div;
            }
          if (c == 6)
            e += vh_bulk_energy(sP + q * 72, cf.alpha, cf.beta);
          e *= JxW;
        }
      if (robin && t < 18 && cell_owned[cell])
        for (int f = 0; f < 6; ++f)
          {
            const int bid = (faces >> (4 * f)) & 15u;
            if (bid < 2 || bid > 4 || (t % 3) == bid - 2)
              continue;
            const double hn = f / 2 == 0 ? h0 : (f / 2 == 1 ? h1 : h2);
            const double s  = cf.K1 / cf.bt * (vol / hn);
            double       ef = 0.0;
            for (int a = 0; a < NN; ++a)
              {
                const double *M = tab.Mf + (size_t)(f * NN + a) * NN;
                double        m = 0.0;
                for (int b = 0; b < NN; ++b)
                  m += M[b] * sU[b * 18 + t];
                ef += sU[a * 18 + t] * m;
              }
            e += s * ef;
          }
      const double s = block_sum(e, sred);
      if (t == 0)
        Ec[cell] = s;
    }
}

// ------------------------------------------------------------------------------------------------
// 2. row-owner Jacobian kernel for Q1 rows whose neighbourhood is a piece of a structured lattice
// ------------------------------------------------------------------------------------------------
#define VH_FAST_THREADS 96
#define VH_FAST_ACTIVE 86 /* 86 threads x 2 packed entries = 172 */

__global__ void __launch_bounds__(VH_FAST_THREADS)
  k_rows_fast_q1(const int32_t *__restrict__ fast_rows, const int32_t *__restrict__ fast_cells,
                 const int8_t *__restrict__ fast_slot, const int32_t *__restrict__ row_ptr, const int32_t *__restrict__ col,
                 const uint32_t *__restrict__ dirmask, const double *__restrict__ cell_h,
                 const uint32_t *__restrict__ cell_faces, const double *__restrict__ Hq, const double *__restrict__ Dc,
                 const double *__restrict__ avgD, VhTables tab, VhCoef cf, double *__restrict__ vals)
{
  __shared__ __align__(16) double s_tile[2][VH_BLK];
  __shared__ double               s_GS[27 * 9];
  __shared__ double               s_FS[27 * 3];
  __shared__ int                  s_cells[8];
  __shared__ int                  s_pos[27];

  const int t = threadIdx.x;
  const int r = blockIdx.x;
  const int I = fast_rows[r];
  if (t < 8)
    s_cells[t] = fast_cells[(size_t)r * 8 + t];
  if (t >= 32 && t < 59)
    s_pos[t - 32] = fast_slot[(size_t)r * 32 + (t - 32)];
  __syncthreads();

  // geometry-only terms per stencil slot:  GS[s][x][y] = sum_(o,b)->s vol_o Gref[7-o][b][x][y] / (h_x h_y)
  for (int i = t; i < 27 * 9; i += VH_FAST_THREADS)
    {
      const int s = i / 9, xy = i - 9 * s, xx = xy / 3, yy = xy - 3 * xx;
      const int sx = s % 3, sy = (s / 3) % 3, sz = s / 9;
      double    acc = 0.0;
      for (int o = 0; o < 8; ++o)
        {
          const int e = s_cells[o];
          const int bx = sx - (o & 1), by = sy - ((o >> 1) & 1), bz = sz - (o >> 2);
          if (e < 0 || bx < 0 || bx > 1 || by < 0 || by > 1 || bz < 0 || bz > 1)
            continue;
          const int     b = bx + 2 * by + 4 * bz, a = 7 - o;
          const double *h = cell_h + 4 * (size_t)e;
          acc += h[3] / (h[xx] * h[yy]) * tab.Gref[(size_t)(a * 8 + b) * 9 + xy];
        }
      s_GS[i] = acc;
    }
  const bool robin = cf.bt < 1e10;
  for (int i = t; i < 27 * 3; i += VH_FAST_THREADS)
    {
      const int s = i / 3, xx = i - 3 * s;
      const int sx = s % 3, sy = (s / 3) % 3, sz = s / 9;
      double    acc = 0.0;
      if (robin)
        for (int o = 0; o < 8; ++o)
          {
            const int e = s_cells[o];
            const int bx = sx - (o & 1), by = sy - ((o >> 1) & 1), bz = sz - (o >> 2);
            if (e < 0 || bx < 0 || bx > 1 || by < 0 || by > 1 || bz < 0 || bz > 1)
              continue;
            const uint32_t faces = cell_faces[e];
            if (!faces)
              continue;
            const int     b = bx + 2 * by + 4 * bz, a = 7 - o;
            const double *h = cell_h + 4 * (size_t)e;
            for (int f = 0; f < 6; ++f)
              {
                const int bid = (faces >> (4 * f)) & 15u;
                if (bid < 2 || bid > 4 || xx == bid - 2)
                  continue;
                acc += cf.K1 / cf.bt * (h[3] / h[f / 2]) * tab.Mf[(size_t)(f * 8 + a) * 8 + b];
              }
          }
      s_FS[i] = acc;
    }

  // bulk part:  acc[s](c,d) = sum_o sum_q sum_b->s  vol_o w_q N_a(q) N_b(q) H_{o,q}(c,d)
  double acc0[27], acc1[27];
#pragma unroll
  for (int s = 0; s < 27; ++s)
    acc0[s] = acc1[s] = 0.0;
  if (t < VH_FAST_ACTIVE)
    {
#pragma unroll
      for (int o = 0; o < 8; ++o)
        {
          const int e = s_cells[o];
          if (e >= 0)
            {
              const double   vol = cell_h[4 * (size_t)e + 3];
              const double2 *Hp  = reinterpret_cast<const double2 *>(Hq + (size_t)e * (8 * VH_SYMP)) + t;
#pragma unroll
              for (int q = 0; q < 8; ++q)
                {
                  const double2 hv = __ldg(Hp + q * (VH_SYMP / 2));
                  const double  hx = hv.x * vol, hy = hv.y * vol;
#pragma unroll
                  for (int b = 0; b < 8; ++b)
                    {
                      const int    s = ((o & 1) + (b & 1)) + 3 * (((o >> 1) & 1) + ((b >> 1) & 1)) + 9 * ((o >> 2) + (b >> 2));
                      const double w = c_W1[((7 - o) * 8 + b) * 8 + q];
                      acc0[s]        = fma(w, hx, acc0[s]);
                      acc1[s]        = fma(w, hy, acc1[s]);
                    }
                }
            }
        }
    }
  __syncthreads();

  const uint32_t maskI = dirmask[I];
  const int      rp    = row_ptr[I];
  const int      p0    = 2 * t, p1 = 2 * t + 1;
  int            c0 = 0, d0 = 0, c1 = 0, d1 = 0;
  if (t < VH_FAST_ACTIVE)
    {
      c0 = c_symc[p0], d0 = c_symd[p0];
      c1 = c_symc[p1], d1 = c_symd[p1];
    }
  const bool have1 = t < VH_FAST_ACTIVE && p1 < VH_SYM;
  int        n_done = 0;
#pragma unroll
  for (int s = 0; s < 27; ++s)
    {
      const int pos = s_pos[s];
      if (pos < 0)
        continue; // block-uniform
      double        *tile  = s_tile[n_done & 1];
      const int      J     = col[rp + pos];
      const uint32_t maskJ = dirmask[J];
      const double   trG   = cf.K1 * (s_GS[s * 9 + 0] + s_GS[s * 9 + 4] + s_GS[s * 9 + 8]);
      if (t < VH_FAST_ACTIVE)
        {
#pragma unroll
          for (int k = 0; k < 2; ++k)
            {
              if (k == 1 && !have1)
                break;
              const int    c = k ? c1 : c0, d = k ? d1 : d0;
              const double a = k ? acc1[s] : acc0[s];
              const int    xc = c % 3, xd = d % 3;
              double       vcd = a, vdc = a;
              if (c == d)
                vcd += trG + s_FS[s * 3 + xc];
              if (c / 3 == d / 3)
                {
                  vcd += cf.K23 * s_GS[s * 9 + xc * 3 + xd];
                  vdc += cf.K23 * s_GS[s * 9 + xd * 3 + xc];
                }
              // component-masked Dirichlet DoFs: row and column dropped (distribute_local_to_global)
              if (((maskI >> c) & 1u) || ((maskJ >> d) & 1u))
                vcd = 0.0;
              if (((maskI >> d) & 1u) || ((maskJ >> c) & 1u))
                vdc = 0.0;
              if (s == 13 && c == d && ((maskI >> c) & 1u))
                { // constrained diagonal: sum over cells of |a_ii| (mean |diag| of the cell if a_ii == 0)
                  double dsum = 0.0;
                  for (int o = 0; o < 8; ++o)
                    {
                      const int e = s_cells[o];
                      if (e < 0)
                        continue;
                      double dv = fabs(Dc[(size_t)e * 144 + (7 - o) * 18 + c]);
                      if (dv == 0.0)
                        dv = avgD[e];
                      dsum += dv;
                    }
                  vcd = dsum;
                }
              tile[c * 18 + d] = vcd;
              if (c != d)
                tile[d * 18 + c] = vdc;
            }
        }
      __syncthreads();
      double2       *dst = reinterpret_cast<double2 *>(vals + (size_t)(rp + pos) * VH_BLK);
      const double2 *src = reinterpret_cast<const double2 *>(tile);
      for (int i = t; i < VH_BLK / 2; i += VH_FAST_THREADS)
        dst[i] = src[i];
      ++n_done;
    }
}

// rhs of fast rows: deterministic gather of the cell rhs over the incident cells (no atomics)
__global__ void k_rhs_fast_q1(int n_fast, const int32_t *__restrict__ fast_rows, const int32_t *__restrict__ fast_cells,
                              const uint32_t *__restrict__ dirmask, const double *__restrict__ Rc, double *__restrict__ rhs)
{
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= (int64_t)n_fast * 18)
    return;
  const int r = (int)(gid / 18), c = (int)(gid - 18 * (int64_t)r);
  const int I = fast_rows[r];
  double    s = 0.0;
#pragma unroll
  for (int o = 0; o < 8; ++o)
    {
      const int e = fast_cells[(size_t)r * 8 + o];
      if (e >= 0)
        s += Rc[(size_t)e * 144 + (7 - o) * 18 + c];
    }
  if ((dirmask[I] >> c) & 1u)
    s = 0.0;
  rhs[(size_t)I * 18 + c] = s;
}

// ------------------------------------------------------------------------------------------------
// 3. general constrained scatter
// ------------------------------------------------------------------------------------------------
__global__ void k_zero_slow_rows(int n_slow_rows, const int32_t *__restrict__ slow_rows, const int32_t *__restrict__ row_ptr,
                                 double *__restrict__ vals, double *__restrict__ rhs, int want_matrix)
{
  const int I = slow_rows[blockIdx.x];
  if (threadIdx.x < 18)
    rhs[(size_t)I * 18 + threadIdx.x] = 0.0;
  if (!want_matrix)
    return;
  const size_t b0 = (size_t)row_ptr[I] * VH_BLK, b1 = (size_t)row_ptr[I + 1] * VH_BLK;
  for (size_t i = b0 + threadIdx.x; i < b1; i += blockDim.x)
    vals[i] = 0.0;
}

struct SlowArgs
{
  const int32_t *slow_cells, *cell_nodes, *row_ptr, *col;
  const double  *cell_h;
  const uint32_t *cell_faces;
  const uint8_t *row_slow;
  const int32_t *line_of, *cptr, *cmaster;
  const double  *cweight;
  const double  *Hq, *Rc, *avgD;
  int            n_owned;
};

__device__ __forceinline__ void slow_add(const SlowArgs &A, int I, int c, int J, int d, double v, double *vals)
{
  if (I >= A.n_owned || !A.row_slow[I])
    return;
  int lo = A.row_ptr[I], hi = A.row_ptr[I + 1] - 1;
  while (lo < hi)
    {
      const int mid = (lo + hi) >> 1;
      if (A.col[mid] < J)
        lo = mid + 1;
      else
        hi = mid;
    }
  if (A.col[lo] == J)
    atomicAdd(vals + (size_t)lo * VH_BLK + c * 18 + d, v);
}

template <int NN, int NQ>
__global__ void k_cells_slow(SlowArgs A, VhTables tab, VhCoef cf, int want_matrix, double *__restrict__ vals,
                             double *__restrict__ rhs)
{
  extern __shared__ double sm[];
  double *sH  = sm;                  // [NQ*172]
  double *sN  = sH + NQ * VH_SYMP;   // [NN*NQ]
  double *swq = sN + NN * NQ;        // [NQ]
  __shared__ int s_nodes[NN];

  const int     t    = threadIdx.x;
  const int64_t cell = A.slow_cells[blockIdx.x];
  if (want_matrix)
    for (int i = t; i < NQ * VH_SYMP; i += blockDim.x)
      sH[i] = A.Hq[cell * (int64_t)(NQ * VH_SYMP) + i];
  for (int i = t; i < NN * NQ; i += blockDim.x)
    sN[i] = tab.N[i];
  if (t < NQ)
    swq[t] = tab.wq[t];
  if (t < NN)
    s_nodes[t] = A.cell_nodes[cell * NN + t];
  __syncthreads();
  const double *h   = A.cell_h + 4 * cell;
  const double  vol = h[3];
  const double  ih[3] = {1.0 / h[0], 1.0 / h[1], 1.0 / h[2]};
  const uint32_t faces = A.cell_faces[cell];
  const bool     robin = (cf.bt < 1e10) && faces != 0u;

  if (want_matrix)
    for (int pair = 0; pair < NN * NN; ++pair)
      {
        const int     a = pair / NN, b = pair - NN * a;
        const double *G = tab.Gref + (size_t)pair * 9;
        for (int e = t; e < VH_BLK; e += blockDim.x)
          {
            const int c = e / 18, d = e - 18 * c;
            const int sidx = c <= d ? vh_sym_index(c, d) : vh_sym_index(d, c);
            double    v = 0.0;
            for (int q = 0; q < NQ; ++q)
              v += swq[q] * sN[a * NQ + q] * sN[b * NQ + q] * sH[q * VH_SYMP + sidx];
            v *= vol;
            if (c == d)
              v += vol * cf.K1 * (G[0] * ih[0] * ih[0] + G[4] * ih[1] * ih[1] + G[8] * ih[2] * ih[2]);
            if (c / 3 == d / 3)
              v += vol * cf.K23 * G[(c % 3) * 3 + d % 3] * ih[c % 3] * ih[d % 3];
            if (robin && c == d)
              for (int f = 0; f < 6; ++f)
                {
                  const int bid = (faces >> (4 * f)) & 15u;
                  if (bid < 2 || bid > 4 || (c % 3) == bid - 2)
                    continue;
                  v += cf.K1 / cf.bt * (vol / h[f / 2]) * tab.Mf[(size_t)(f * NN + a) * NN + b];
                }
            // AffineConstraints::distribute_local_to_global
            const int gi = 18 * s_nodes[a] + c, gj = 18 * s_nodes[b] + d;
            const int li = A.line_of[gi], lj = A.line_of[gj];
            if (li < 0 && lj < 0)
              slow_add(A, s_nodes[a], c, s_nodes[b], d, v, vals);
            else
              {
                const int r0 = li < 0 ? 0 : A.cptr[li], r1 = li < 0 ? 1 : A.cptr[li + 1];
                const int q0 = lj < 0 ? 0 : A.cptr[lj], q1 = lj < 0 ? 1 : A.cptr[lj + 1];
                for (int rr = r0; rr < r1; ++rr)
                  {
                    const int    rd = li < 0 ? gi : A.cmaster[rr];
                    const double rw = li < 0 ? 1.0 : A.cweight[rr];
                    for (int qq = q0; qq < q1; ++qq)
                      {
                        const int    cd = lj < 0 ? gj : A.cmaster[qq];
                        const double cw = lj < 0 ? 1.0 : A.cweight[qq];
                        slow_add(A, rd / 18, rd % 18, cd / 18, cd % 18, rw * cw * v, vals);
                      }
                  }
                if (gi == gj && li >= 0)
                  {
                    double dv = fabs(v);
                    if (dv == 0.0)
                      dv = A.avgD[cell];
                    slow_add(A, s_nodes[a], c, s_nodes[a], c, dv, vals);
                  }
              }
          }
      }
  // vector part (rhs[m_k] += w_k r_i; constrained entries stay 0)
  for (int i = t; i < 18 * NN; i += blockDim.x)
    {
      const int    a = i / 18, c = i - 18 * a;
      const double r = A.Rc[cell * (int64_t)(18 * NN) + i];
      const int    gi = 18 * s_nodes[a] + c, li = A.line_of[gi];
      if (li < 0)
        {
          const int I = s_nodes[a];
          if (I < A.n_owned && A.row_slow[I])
            atomicAdd(rhs + gi, r);
        }
      else
        for (int rr = A.cptr[li]; rr < A.cptr[li + 1]; ++rr)
          {
            const int rd = A.cmaster[rr], I = rd / 18;
            if (I < A.n_owned && A.row_slow[I])
              atomicAdd(rhs + rd, A.cweight[rr] * r);
          }
    }
}

template <int NN, int NQ>
size_t pointwise_smem(bool want_h)
{
  size_t n = (size_t)NN * 18 + NQ * 18 + NQ * 54 + NQ * 72 + NQ * 18 + NN * NQ + NN * NQ * 3 + NQ + 32;
  if (want_h)
    n += (size_t)NQ * VH_BLK;
  return n * sizeof(double);
}
} // namespace

int vhk_upload_constants(vh_ctx *ctx)
{
  // symmetric packing tables
  uint8_t sc[VH_SYMP], sd[VH_SYMP];
  int     k = 0;
  for (int c = 0; c < 18; ++c)
    for (int d = c; d < 18; ++d)
      {
        sc[k] = (uint8_t)c;
        sd[k] = (uint8_t)d;
        ++k;
      }
  sc[VH_SYM] = 17;
  sd[VH_SYM] = 17; // padding slot (never written back)
  VH_CUDA(cudaMemcpyToSymbol(c_symc, sc, sizeof(sc)));
  VH_CUDA(cudaMemcpyToSymbol(c_symd, sd, sizeof(sd)));
  return VH_OK;
}

int vhk_upload_w1(vh_ctx *ctx, const double *W1)
{
  VH_CUDA(cudaMemcpyToSymbol(c_W1, W1, 512 * sizeof(double)));
  return VH_OK;
}

int vhk_pointwise(vh_ctx *ctx, const double *x_local, bool want_h, bool want_e)
{
  if (ctx->n_cells == 0)
    return VH_OK;
  if (ctx->degree == 1)
    {
      const size_t smem = pointwise_smem<8, 8>(want_h);
      k_pointwise<8, 8><<<ctx->n_cells, 160, smem, ctx->stream>>>(ctx->cell_nodes, ctx->cell_h, ctx->cell_faces, ctx->cell_owned,
                                                                 x_local, ctx->tab, ctx->coef, want_h, want_e, ctx->Hq, ctx->Rc,
                                                                 ctx->Dc, ctx->avgD, ctx->Ec);
    }
  else
    {
      const size_t smem = pointwise_smem<27, 27>(want_h);
      static bool  attr_set = false;
      if (!attr_set)
        {
          VH_CUDA(cudaFuncSetAttribute(k_pointwise<27, 27>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)pointwise_smem<27, 27>(true)));
          attr_set = true;
        }
      k_pointwise<27, 27><<<ctx->n_cells, 512, smem, ctx->stream>>>(ctx->cell_nodes, ctx->cell_h, ctx->cell_faces,
                                                                   ctx->cell_owned, x_local, ctx->tab, ctx->coef, want_h, want_e,
                                                                   ctx->Hq, ctx->Rc, ctx->Dc, ctx->avgD, ctx->Ec);
    }
  VH_LAUNCH_CHECK();
  return VH_OK;
}

int vhk_rows_fast(vh_ctx *ctx)
{
  if (ctx->n_fast == 0)
    return VH_OK;
  k_rows_fast_q1<<<ctx->n_fast, VH_FAST_THREADS, 0, ctx->stream>>>(ctx->fast_rows, ctx->fast_cells, ctx->fast_slot, ctx->row_ptr,
                                                                  ctx->col, ctx->dirmask, ctx->cell_h, ctx->cell_faces, ctx->Hq,
                                                                  ctx->Dc, ctx->avgD, ctx->tab, ctx->coef, ctx->vals);
  VH_LAUNCH_CHECK();
  return VH_OK;
}

int vhk_rhs_fast(vh_ctx *ctx, double *rhs_out)
{
  if (ctx->n_fast == 0)
    return VH_OK;
  const int64_t n = (int64_t)ctx->n_fast * 18;
  k_rhs_fast_q1<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(ctx->n_fast, ctx->fast_rows, ctx->fast_cells, ctx->dirmask,
                                                                     ctx->Rc, rhs_out);
  VH_LAUNCH_CHECK();
  return VH_OK;
}

int vhk_rows_slow(vh_ctx *ctx, bool want_matrix, double *rhs_out)
{
  if (ctx->n_slow_rows == 0)
    return VH_OK;
  k_zero_slow_rows<<<ctx->n_slow_rows, 128, 0, ctx->stream>>>(ctx->n_slow_rows, ctx->slow_rows, ctx->row_ptr, ctx->vals, rhs_out,
                                                             want_matrix ? 1 : 0);
  VH_LAUNCH_CHECK();
  if (ctx->n_slow_cells == 0)
    return VH_OK;
  SlowArgs A;
  A.slow_cells = ctx->slow_cells;
  A.cell_nodes = ctx->cell_nodes;
  A.row_ptr    = ctx->row_ptr;
  A.col        = ctx->col;
  A.cell_h     = ctx->cell_h;
  A.cell_faces = ctx->cell_faces;
  A.row_slow   = ctx->row_slow;
  A.line_of    = ctx->cons[0].line_of;
  A.cptr       = ctx->cons[0].ptr;
  A.cmaster    = ctx->cons[0].master;
  A.cweight    = ctx->cons[0].weight;
  A.Hq         = ctx->Hq;
  A.Rc         = ctx->Rc;
  A.avgD       = ctx->avgD;
  A.n_owned    = ctx->n_owned;
  if (ctx->degree == 1)
    {
      const size_t smem = (size_t)(8 * VH_SYMP + 64 + 8) * sizeof(double);
      k_cells_slow<8, 8><<<ctx->n_slow_cells, 352, smem, ctx->stream>>>(A, ctx->tab, ctx->coef, want_matrix ? 1 : 0, ctx->vals,
                                                                       rhs_out);
    }
  else
    {
      const size_t smem = (size_t)(27 * VH_SYMP + 729 + 27) * sizeof(double);
      k_cells_slow<27, 27><<<ctx->n_slow_cells, 352, smem, ctx->stream>>>(A, ctx->tab, ctx->coef, want_matrix ? 1 : 0, ctx->vals,
                                                                         rhs_out);
    }
  VH_LAUNCH_CHECK();
  return VH_OK;
}
