// Inverse of the nodal 18x18 diagonal blocks: setup of the block-Jacobi preconditioner the north star prescribes in place
// of the reference's ML-AMG (preconditioner.initialize, /root/reference/femgl/src/solve.cc:130-154).
//
// One warp per block, 4 blocks per CTA.  The block is staged through shared memory in both directions so that every
// global access is a coalesced 16-byte piece (1440 B packed / 2592 B full in, 2592 B out: the kernel's HBM traffic), and
// the elimination itself is register-resident: lane r keeps row r of the working matrix.
//
// In-place Gauss-Jordan with partial pivoting and DEFERRED row scaling: step k picks the not yet used row with the largest
// |a[r][k]| (one redux.sync on the float-truncated magnitudes + a ballot; ties go to the lowest lane).  The pivot row is
// broadcast UNSCALED through a double-buffered shared row (one __syncwarp per step) and every other row subtracts
// (a[r][k] / pivot) times it; the pivot lane runs the same instructions with a zero multiplier, so there is no divergent
// scaling pass and a step costs 18 DFMA per lane.  The freed column k stores the column of the inverse that became
// non-trivial in this step.  Rows are never swapped and never normalised on the way: at the end the lane that pivoted on
// column k holds pivot_k times row k of A^-1 with its columns in the order pl_0 .. pl_17, and scales it once.
// The reciprocals are rcp.approx + two Newton steps (full double precision up to the last bit, no slow path).
//
// Included by vh_linalg.cu; compiled for the host by tests/native/points_emu_host.cc under the CUDA emulation.
#ifndef VH_BLOCK_INVERT_CUH
#define VH_BLOCK_INVERT_CUH

#include "vh_internal.h"
#include "vh_pointwise.cuh"

#define VH_INV_WARPS 4

VH_HD double vh_fast_rcp(double x)
{
#ifdef __CUDA_ARCH__
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double e = fma(-x, y, 1.0);
  y        = fma(y, e, y);
  e        = fma(-x, y, 1.0);
  return fma(y, e, y);
#else
  return 1.0 / x;
#endif
}

__global__ void __launch_bounds__(VH_INV_WARPS * 32)
  k_block_invert(int n_rows, const int32_t *__restrict__ diag_pos, const double *__restrict__ vals, double *__restrict__ minv,
                 int *__restrict__ n_singular, const double *__restrict__ pvals, int cm_stride, int diag_slot, const int32_t *__restrict__ fast_index,
                 const int32_t *__restrict__ fast_class, const double *__restrict__ class_M, const uint32_t *__restrict__ dirmask,
                 const double *__restrict__ cdiag, const double *__restrict__ dpack, int packed)
{
  __shared__ __align__(16) double s_blk[VH_INV_WARPS][VH_BLK];   // the block on its way in (packed or full), then A^-1 on its way out
  __shared__ __align__(16) double s_row[VH_INV_WARPS][2][18];    // pivot row of the current step (double-buffered)
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int row  = blockIdx.x * VH_INV_WARPS + wid;
  if (row >= n_rows)
    return;
  const int rr = lane < 18 ? lane : 17;
  double    a[18];
  double   *sb = s_blk[wid];
  const int fi = packed ? fast_index[row] : -1;
  if (fi >= 0)
    { // packed row: diagonal block = Sym(P) + kron(I_6, M_13), Dirichlet rows/columns -> sum_cells |a_ii| on the diagonal;
      // P comes from dpack[fast row] while the lattice rows are not assembled (matrix-free default), else from the row itself
      const double2 *P = reinterpret_cast<const double2 *>(dpack ? dpack + (size_t)fi * VH_SYMP : pvals + (size_t)diag_pos[row] * VH_SYMP);
#pragma unroll
      for (int r = 0; r < 3; ++r)
        if (lane + 32 * r < VH_SYMP / 2)
          reinterpret_cast<double2 *>(sb)[lane + 32 * r] = __ldg(P + lane + 32 * r);
      const double  *M  = class_M + (size_t)fast_class[fi] * cm_stride + diag_slot * 10 + 3 * (rr % 3);
      const double   m0 = __ldg(M), m1 = __ldg(M + 1), m2 = __ldg(M + 2);
      const uint32_t mI = dirmask[row];
      const double   cd = cdiag[(size_t)row * 18 + rr];
      const bool     rmask = (mI >> rr) & 1u;
      const int      rs = vh_sym_rowstart(rr) - 2 * (rr >> 1); // vh_sym_index(rr, c) = rs + c
      __syncwarp();
#pragma unroll
      for (int c = 0; c < 18; ++c)
        {
          double v = sb[rr <= c ? rs + c : vh_sym_index(c, rr)];
          if (rr / 3 == c / 3)
            v += (c % 3 == 0) ? m0 : ((c % 3 == 1) ? m1 : m2);
          if (rmask || ((mI >> c) & 1u))
            v = (rr == c) ? cd : 0.0;
          a[c] = v;
        }
    }
  else
    {
      const double2 *B = reinterpret_cast<const double2 *>(vals + (size_t)diag_pos[row] * VH_BLK);
#pragma unroll
      for (int r = 0; r < 6; ++r)
        if (lane + 32 * r < VH_BLK / 2)
          reinterpret_cast<double2 *>(sb)[lane + 32 * r] = __ldg(B + lane + 32 * r);
      __syncwarp();
#pragma unroll
      for (int c = 0; c < 9; ++c)
        {
          const double2 v = *reinterpret_cast<const double2 *>(sb + rr * 18 + 2 * c);
          a[2 * c]        = v.x;
          a[2 * c + 1]    = v.y;
        }
    }
  bool     used  = lane >= 18;
  int      mycol = 0;
  double   dpiv  = 1.0;          // the pivot of this lane's row
  bool     singular = false;
  unsigned pk[3] = {0u, 0u, 0u}; // pivot lanes pl_0 .. pl_17, 5 bits each (warp-uniform)
#pragma unroll
  for (int k = 0; k < 18; ++k)
    {
      const unsigned key  = used ? 0u : __float_as_uint(fabsf(__double2float_ru(fabs(a[k]))));
      const unsigned best = __reduce_max_sync(0xffffffffu, key);
      if (best == 0u)
        {
          singular = true;
          break;
        }
      const int    pl  = __ffs(__ballot_sync(0xffffffffu, key == best)) - 1;
      const bool   me  = lane == pl;
      const double f   = a[k];
      const double piv = __shfl_sync(0xffffffffu, f, pl);
      pk[k / 6] |= (unsigned)pl << (5 * (k % 6));
      double *sr = s_row[wid][k & 1];
      if (me)
        {
          dpiv  = piv;
          used  = true;
          mycol = k;
          a[k]  = 1.0;
#pragma unroll
          for (int c = 0; c < 9; ++c)
            *reinterpret_cast<double2 *>(sr + 2 * c) = make_double2(a[2 * c], a[2 * c + 1]);
        }
      else
        a[k] = 0.0;
      const double g = me ? 0.0 : -f * vh_fast_rcp(piv);
      __syncwarp();
#pragma unroll
      for (int c = 0; c < 9; ++c)
        {
          const double2 p = *reinterpret_cast<const double2 *>(sr + 2 * c);
          a[2 * c]        = fma(g, p.x, a[2 * c]);
          a[2 * c + 1]    = fma(g, p.y, a[2 * c + 1]);
        }
    }
  if (singular)
    {
      if (lane == 0)
        atomicAdd(n_singular, 1);
      for (int i = lane; i < VH_BLK; i += 32)
        minv[(size_t)row * VH_BLK + i] = (i / 18 == i % 18) ? 1.0 : 0.0;
      return;
    }
  __syncwarp(); // every lane has read its row of the staged input (and the last pivot row)
  if (lane < 18)
    {
      const double s   = vh_fast_rcp(dpiv);
      double      *out = sb + mycol * 18;
#pragma unroll
      for (int m = 0; m < 18; ++m)
        out[(pk[m / 6] >> (5 * (m % 6))) & 31u] = a[m] * s;
    }
  __syncwarp();
  double2 *dst = reinterpret_cast<double2 *>(minv + (size_t)row * VH_BLK);
#pragma unroll
  for (int r = 0; r < 6; ++r)
    if (lane + 32 * r < VH_BLK / 2)
      dst[lane + 32 * r] = reinterpret_cast<const double2 *>(sb)[lane + 32 * r];
}

#endif
