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
#include <cstdlib>

__constant__ double  c_W1[512];   // Q1: w_q N_a(q) N_b(q), index (a*8+b)*8+q
__constant__ uint8_t c_symc[VH_SYMP], c_symd[VH_SYMP];
__constant__ double  c_T2[27];    // Q2, one direction: w_q l_a(q) l_b(q), index (t_a*3 + t_b)*3 + q  (t = 0 low, 1 mid, 2 high)
__constant__ uint8_t c_q2t[27 * 3]; // Q2: tensor index (t_x, t_y, t_z) of local node a (deal.II hierarchical order)

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
__global__ void __launch_bounds__(NQ == 8 ? 192 : 512)
  k_pointwise(const int32_t *__restrict__ cell_nodes, const double *__restrict__ cell_h, const uint32_t *__restrict__ cell_faces,
              const uint8_t *__restrict__ cell_owned, const double *__restrict__ x, VhTables tab, VhCoef cf, int want_h, int want_e,
              double *__restrict__ Hq, double *__restrict__ Rc, double *__restrict__ Dc, double *__restrict__ avgD,
              double *__restrict__ Ec)
{
  extern __shared__ double sm[];
  double *sU   = sm;                // [NN*18]
  double *sT   = sU + NN * 18;      // [NQ][VH_TQ]  per quadrature point: A(18) | products R,Q,P,S (72) | Z1,Z2 (324)
  double *sdA  = sT + NQ * VH_TQ;   // [NQ*18*3]
  double *sg   = sdA + NQ * 54;     // [NQ*18]
  double *sHd  = sg + NQ * 18;      // [NQ*18]   diagonal of H_q (unscaled)
  double *sN   = sHd + NQ * 18;     // [NN*NQ]
  double *sdN  = sN + NN * NQ;      // [NN*NQ*3]
  double *swq  = sdN + NN * NQ * 3; // [NQ]
  double *sred = swq + NQ;          // [32]

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
      sT[q * VH_TQ + VH_TQ_A + c] = A;
      sdA[3 * t + 0]              = d0 * ih[0];
      sdA[3 * t + 1]              = d1 * ih[1];
      sdA[3 * t + 2]              = d2 * ih[2];
    }
  __syncthreads();
  // per quadrature point: 36 product entries (+ 162 Z-table entries for the Hessian)
  {
    const int per_q = want_h ? 198 : 36;
    for (int i = t; i < NQ * per_q; i += blockDim.x)
      {
        const int q = i / per_q, e = i - per_q * q;
        double   *T = sT + q * VH_TQ;
        if (e < 36)
          vh_product_entry(T + VH_TQ_A, e, T + VH_TQ_P + 2 * e);
        else
          vh_ztable_entry(T + VH_TQ_A, e - 36, T + VH_TQ_Z);
      }
  }
  __syncthreads();
  if (t < 18 * NQ)
    {
      const int q = t / 18, c = t - 18 * q;
      sg[t]       = vh_g_component(sT + q * VH_TQ + VH_TQ_A, sT + q * VH_TQ + VH_TQ_P, c, cf.alpha, cf.beta);
    }
  if (want_h)
    { // the 171 unique entries of every H_q, stored pre-multiplied by the cell volume (JxW = w_q * vol).
      // A thread owns ONE packed entry (c,d): its 16-term list is set up once and reused for every quadrature point.
      const int n_groups = blockDim.x / VH_SYMP, grp = t / VH_SYMP, sidx = t - VH_SYMP * grp;
      if (grp < n_groups)
        {
          double *dst = Hq + cell * (int64_t)(NQ * VH_SYMP) + sidx;
          if (c_symd[sidx] >= c_symc[sidx])
            {
              const int c = c_symc[sidx], d = c_symd[sidx];
              vh_terms  T;
              vh_entry_terms(c, d, cf.alpha, cf.beta, T);
              for (int q = grp; q < NQ; q += n_groups)
                {
                  const double v = vh_entry_eval(sT + q * VH_TQ, T);
                  if (c == d)
                    sHd[q * 18 + c] = v;
                  dst[q * VH_SYMP] = v * vol;
                }
            }
          else
            for (int q = grp; q < NQ; q += n_groups)
              dst[q * VH_SYMP] = 0.0;
        }
    }
  __syncthreads();

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
            dg += JxW * Na * Na * sHd[q * 18 + c];
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
            e += vh_bulk_energy(sT + q * VH_TQ + VH_TQ_P, cf.alpha, cf.beta);
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
// 1b. pointwise kernel, one THREAD per quadrature point: csrc/vh_points_kernel.cuh (k_points, VhPt, VH_PT_WARPS)
// ------------------------------------------------------------------------------------------------
#include "vh_points_kernel.cuh"
#include "vh_diag_kernel.cuh"
#include "vh_apply_v2.cuh"

// ------------------------------------------------------------------------------------------------
// 2. row-owner Jacobian kernel for Q1 rows whose neighbourhood is a piece of a structured lattice
// ------------------------------------------------------------------------------------------------
#define VH_FAST_STAGES 4
#define VH_CELL_H_BYTES (8 * VH_SYMP * 8) /* one cell's 8 x 172 doubles: 11008 B, a multiple of 16 */
#define VH_FAST_SMEM (VH_FAST_STAGES * VH_CELL_H_BYTES + VH_BLK * 8 + 32 * 8 + VH_FAST_STAGES * 8 + (36 + 28) * 4)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
  asm volatile("{\n"
               ".reg .pred p;\n"
               "VH_WAIT_%=:\n"
               "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
               "@p bra VH_DONE_%=;\n"
               "bra VH_WAIT_%=;\n"
               "VH_DONE_%=:\n"
               "}" ::"r"(smem_u32(bar)),
               "r"(parity)
               : "memory");
}
// TMA 1-D bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// EPT = packed Hessian entries per thread: 2 -> 96 threads/row (86 active), 108 accumulator registers, 12 warps/SM;
//                                          1 -> 192 threads/row (172 active), 54 accumulator registers, 24 warps/SM.
// PACK = true: the row is stored as packed symmetric blocks (172 doubles each): the accumulators go straight from
//               registers to global memory with coalesced stores; geometry terms and Dirichlet masks are applied by the
//               consumers (SpMV, block-Jacobi setup, export), so the expansion stage disappears.
template <int EPT, bool PACK>
__global__ void __launch_bounds__(192 / EPT, 4)
  k_rows_fast_q1(const int32_t *__restrict__ fast_rows, const int32_t *__restrict__ fast_cells,
                 const int8_t *__restrict__ fast_slot, const int32_t *__restrict__ fast_class,
                 const double *__restrict__ class_M, const int32_t *__restrict__ row_ptr, const int32_t *__restrict__ col,
                 const uint32_t *__restrict__ dirmask, const double *__restrict__ Hq, const double *__restrict__ Dc,
                 const double *__restrict__ avgD, VhCoef cf, double *__restrict__ vals)
{
  extern __shared__ __align__(128) unsigned char smraw[];
  double   *s_H    = reinterpret_cast<double *>(smraw);                                     // [4][8*172], later [27][172]
  double   *s_cls  = reinterpret_cast<double *>(smraw + VH_FAST_STAGES * VH_CELL_H_BYTES);  // [27][12]: GS(9), FS(3)
  double   *s_tr   = s_cls + VH_BLK;                                                        // [27] tr(GS), padded to 32
  uint64_t *s_bar  = reinterpret_cast<uint64_t *>(s_tr + 32);                               // [4]
  int      *s_cells = reinterpret_cast<int *>(s_bar + VH_FAST_STAGES);                      // [8]
  int      *s_pos   = s_cells + 8;                                                          // [27] (+1 pad)
  uint32_t *s_maskJ = reinterpret_cast<uint32_t *>(s_pos + 28);                             // [27]

  constexpr int NT = 192 / EPT, ACT = VH_SYMP / EPT;
  const int     t = threadIdx.x;
  const int     r = blockIdx.x;
  const int     I = fast_rows[r];
  if (t == 0)
    { // TMA producer, first thing in the CTA: the DRAM/L2 latency of the first four cells' tables overlaps the prologue
#pragma unroll
      for (int k = 0; k < VH_FAST_STAGES; ++k)
        mbar_init(s_bar + k, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
#pragma unroll
      for (int o = 0; o < VH_FAST_STAGES; ++o)
        {
          const int e = fast_cells[(size_t)r * 8 + o];
          if (e >= 0)
            {
              mbar_expect_tx(s_bar + o, VH_CELL_H_BYTES);
              bulk_g2s(s_H + (size_t)o * (8 * VH_SYMP), Hq + (size_t)e * (8 * VH_SYMP), VH_CELL_H_BYTES, s_bar + o);
            }
        }
    }
  if (t >= 64 && t < 72)
    s_cells[t - 64] = fast_cells[(size_t)r * 8 + (t - 64)];
  if (t >= 32 && t < 59)
    s_pos[t - 32] = fast_slot[(size_t)r * 32 + (t - 32)];
  if constexpr (!PACK)
    { // geometry-only part of every block of this row's stencil: block += kron(I_6, M_s), M_s 3x3 (stride 10, [9] = 0)
      const double *cls = class_M + (size_t)fast_class[r] * 270;
      for (int i = t; i < 270; i += NT)
        s_cls[i] = cls[i];
    }
  __syncthreads();

  // TMA producer: thread 0 streams the incident cells' pre-scaled H_q tables (11 KB each, contiguous) into the ring
  auto issue = [&](int o) {
    const int e = s_cells[o];
    if (e >= 0)
      {
        uint64_t *bar = s_bar + (o & (VH_FAST_STAGES - 1));
        mbar_expect_tx(bar, VH_CELL_H_BYTES);
        bulk_g2s(s_H + (size_t)(o & (VH_FAST_STAGES - 1)) * (8 * VH_SYMP), Hq + (size_t)e * (8 * VH_SYMP), VH_CELL_H_BYTES, bar);
      }
  };
  // bulk part:  acc[s](c,d) = sum_o sum_q sum_b->s  w_q N_a(q) N_b(q) (vol_o H_{o,q})(c,d)
  double acc[EPT][27];
#pragma unroll
  for (int s = 0; s < 27; ++s)
#pragma unroll
    for (int k = 0; k < EPT; ++k)
      acc[k][s] = 0.0;
  uint32_t uses = 0; // bit k: parity of the next completed phase of stage k
#pragma unroll
  for (int o = 0; o < 8; ++o)
    {
      if (o == 2 || o == 4)
        { // stages {0,1} (resp. {2,3}) have been consumed by every thread: refill them with cells o+2, o+3
          __syncthreads();
          if (t == 0)
            {
              issue(o + 2);
              issue(o + 3);
            }
        }
      const int e = s_cells[o];
      if (e >= 0)
        {
          const int st = o & (VH_FAST_STAGES - 1);
          mbar_wait(s_bar + st, (uses >> st) & 1u);
          uses ^= 1u << st;
          if (t < ACT)
            {
              // cell table layout [pair][q XOR (pair & 7)] (vh_hq8_index): the stage base is 128-byte aligned, so the
              // address of point q is the address of point 0 with bits 4..6 flipped by q
              const int      pr0 = (EPT * t) >> 1;
              const uint32_t Hs0 = smem_u32(s_H + (size_t)st * (8 * VH_SYMP)) + (uint32_t)((pr0 << 7) + ((pr0 & 7) << 4) + ((EPT * t) & 1) * 8);
#pragma unroll
              for (int q = 0; q < 8; ++q)
                {
                  double hv[EPT];
                  if constexpr (EPT == 2)
                    asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(hv[0]), "=d"(hv[1]) : "r"(Hs0 ^ (uint32_t)(q << 4)));
                  else
                    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(hv[0]) : "r"(Hs0 ^ (uint32_t)(q << 4)));
#pragma unroll
                  for (int b = 0; b < 8; ++b)
                    {
                      const int    s = ((o & 1) + (b & 1)) + 3 * (((o >> 1) & 1) + ((b >> 1) & 1)) + 9 * ((o >> 2) + (b >> 2));
                      const double w = c_W1[((7 - o) * 8 + b) * 8 + q];
#pragma unroll
                      for (int k = 0; k < EPT; ++k)
                        acc[k][s] = fma(w, hv[k], acc[k][s]);
                    }
                }
            }
        }
    }

  if constexpr (PACK)
    { // packed storage: block (rp+pos) holds the 172 packed entries contiguously; thread t owns entries EPT*t..
      const int rp = row_ptr[I];
      if (t < ACT)
        {
#pragma unroll
          for (int s = 0; s < 27; ++s)
            {
              const int pos = s_pos[s];
              if (pos < 0)
                continue;
              double *dst = vals + (size_t)(rp + pos) * VH_SYMP + EPT * t;
              if constexpr (EPT == 2)
                __stcs(reinterpret_cast<double2 *>(dst), make_double2(acc[0][s], acc[1][s]));
              else
                __stcs(dst, acc[0][s]);
            }
        }
      return;
    }

  // ---- write the block row ----
  // 1. dump the packed symmetric accumulators of all 27 slots into the (now idle) TMA ring: [27][172] doubles
  __syncthreads();
  double *s_sym = s_H;
  if (t < ACT)
    {
#pragma unroll
      for (int s = 0; s < 27; ++s)
        {
          if constexpr (EPT == 2)
            reinterpret_cast<double2 *>(s_sym + s * VH_SYMP)[t] = make_double2(acc[0][s], acc[1][s]);
          else
            s_sym[s * VH_SYMP + t] = acc[0][s];
        }
    }
  const uint32_t maskI = dirmask[I];
  const int      rp    = row_ptr[I];
  if (t < 27)
    { // per-slot column Dirichlet masks, so the store loop has no dependent global loads
      const int pos = s_pos[t];
      s_maskJ[t]    = pos >= 0 ? dirmask[col[rp + pos]] : 0u;
    }
  __syncthreads();
  // 2. every thread owns EPT fixed 16-byte pieces of the 18x18 block (double2 #t [and #t+96]): the packed offsets and
  //    geometry selectors of its entries are loop invariants; the slot loop is rolled and barrier-free.
  constexpr int NE = 2 * EPT;
  int           soff[NE], gsel[NE];
  uint32_t      rbit[NE], cbit[NE];
#pragma unroll
  for (int k = 0; k < NE; ++k)
    {
      const int i = t + NT * (k >> 1); // double2 index inside the block
      const int c = min((2 * i) / 18, 17), d = (2 * i) % 18 + (k & 1);
      soff[k] = c <= d ? vh_sym_index(c, d) : vh_sym_index(d, c);
      gsel[k] = (c / 3 == d / 3) ? (c % 3) * 3 + d % 3 : 9; // entry 9 of every M_s is 0
      rbit[k] = 1u << c;
      cbit[k] = 1u << d;
    }
  const bool first = t < VH_BLK / 2, second = EPT == 2 && t + NT < VH_BLK / 2;
#pragma unroll 3
  for (int s = 0; s < 27; ++s)
    {
      const int pos = s_pos[s];
      if (pos < 0)
        continue; // block-uniform
      const double  *sy    = s_sym + s * VH_SYMP;
      const double  *M     = s_cls + s * 10;
      const uint32_t maskJ = s_maskJ[s];
      double         v[NE];
#pragma unroll
      for (int k = 0; k < NE; ++k)
        v[k] = sy[soff[k]] + M[gsel[k]];
      if ((maskI | maskJ) != 0u)
        { // component-masked Dirichlet DoFs (block-uniform branch): row and column dropped (distribute_local_to_global)
#pragma unroll
          for (int k = 0; k < NE; ++k)
            if ((maskI & rbit[k]) || (maskJ & cbit[k]))
              v[k] = 0.0;
          if (s == 13)
            { // constrained diagonal: sum over cells of |a_ii| (mean |diag| of the cell if a_ii == 0)
#pragma unroll
              for (int k = 0; k < NE; ++k)
                if (rbit[k] == cbit[k] && (maskI & rbit[k]))
                  {
                    const int c    = 31 - __clz(rbit[k]);
                    double    dsum = 0.0;
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
                    v[k] = dsum;
                  }
            }
        }
      double2 *dst = reinterpret_cast<double2 *>(vals + (size_t)(rp + pos) * VH_BLK);
      if (first)
        __stcs(dst + t, make_double2(v[0], v[1])); // streaming stores: the block is not re-read by this kernel
      if constexpr (EPT == 2)
        if (second)
          __stcs(dst + t + NT, make_double2(v[2], v[3]));
    }
}

// ------------------------------------------------------------------------------------------------
// 2q. row-owner Jacobian kernel for Q2 lattice rows (packed storage), sum-factorised
// ------------------------------------------------------------------------------------------------
// Row node I is local node a_k (tensor index (ax,ay,az)) of its 1/2/4/8 incident cells; its stencil has up to 5x5x5
// slots.  For one incident cell and one packed Hessian entry e the 27 contributions
//     out[bx][by][bz] = sum_q w_q N_a(q) N_b(q) (vol H_q)[e],   N_a(q) = l_ax(qx) l_ay(qy) l_az(qz)
// factorise over the directions:  W[a][b][q] = T[ax][bx][qx] T[ay][by][qy] T[az][bz][qz]  with the 27-entry 1-D table T.
// Thread e keeps its 27 H values in registers and contracts qx, qy, qz in turn: 3 x 81 = 243 FMAs instead of 27 x 27 =
// 729 (the full Q2 cell matrix costs 2.36 MFLOP instead of SURVEY's 13.2 MFLOP).  Slots fed by several cells are
// accumulated by the SAME thread with plain load-add-store on its own entry (first-writer mask from the host): no
// atomics, no memset, and the row (<= 180 KB) stays in L2 between the visits.
template <int MINB>
__global__ void __launch_bounds__(192, MINB)
  k_rows_fast_q2(const int32_t *__restrict__ fast_rows, const int32_t *__restrict__ fast_cells, const int8_t *__restrict__ fast_a,
                 const int8_t *__restrict__ fast_slot, const uint32_t *__restrict__ fast_first, const int32_t *__restrict__ row_ptr,
                 const double *__restrict__ Hq, double *__restrict__ vals)
{
  __shared__ int8_t   s_pos[128];
  __shared__ int      s_cells[8], s_a[8];
  __shared__ uint32_t s_first[8];
  const int t = threadIdx.x, r = blockIdx.x;
  if (t < 128)
    s_pos[t] = fast_slot[(size_t)r * 128 + t];
  if (t >= 128 && t < 136)
    {
      s_cells[t - 128] = fast_cells[(size_t)r * 8 + (t - 128)];
      s_a[t - 128]     = fast_a[(size_t)r * 8 + (t - 128)];
      s_first[t - 128] = fast_first[(size_t)r * 8 + (t - 128)];
    }
  __syncthreads();
  if (t >= VH_SYMP)
    return;
  double *row = vals + (size_t)row_ptr[fast_rows[r]] * VH_SYMP + t;
  for (int k = 0; k < 8; ++k)
    {
      const int e = s_cells[k];
      if (e < 0)
        break;
      const int      a = s_a[k], ax = c_q2t[3 * a], ay = c_q2t[3 * a + 1], az = c_q2t[3 * a + 2];
      const uint32_t first = s_first[k];
      const double  *H = Hq + (size_t)e * (27 * VH_SYMP) + t;
      double         h[27];
#pragma unroll
      for (int q = 0; q < 27; ++q)
        h[q] = __ldg(H + q * VH_SYMP);
      double T[9];
      // contract qx:  t1[bx][qy,qz]
#pragma unroll
      for (int i = 0; i < 9; ++i)
        T[i] = c_T2[ax * 9 + i];
      double t1[3][9];
#pragma unroll
      for (int bx = 0; bx < 3; ++bx)
#pragma unroll
        for (int j = 0; j < 9; ++j)
          t1[bx][j] = fma(T[3 * bx + 2], h[3 * j + 2], fma(T[3 * bx + 1], h[3 * j + 1], T[3 * bx] * h[3 * j]));
      // contract qy:  t2[bx][by][qz]
#pragma unroll
      for (int i = 0; i < 9; ++i)
        T[i] = c_T2[ay * 9 + i];
      double t2[3][3][3];
#pragma unroll
      for (int bx = 0; bx < 3; ++bx)
#pragma unroll
        for (int by = 0; by < 3; ++by)
#pragma unroll
          for (int qz = 0; qz < 3; ++qz)
            t2[bx][by][qz] = fma(T[3 * by + 2], t1[bx][3 * qz + 2], fma(T[3 * by + 1], t1[bx][3 * qz + 1], T[3 * by] * t1[bx][3 * qz]));
      // contract qz and add into the row
#pragma unroll
      for (int i = 0; i < 9; ++i)
        T[i] = c_T2[az * 9 + i];
      const int base = (2 - ax) + 5 * (2 - ay) + 25 * (2 - az);
#pragma unroll
      for (int bz = 0; bz < 3; ++bz)
#pragma unroll
        for (int by = 0; by < 3; ++by)
#pragma unroll
          for (int bx = 0; bx < 3; ++bx)
            {
              double  v   = fma(T[3 * bz + 2], t2[bx][by][2], fma(T[3 * bz + 1], t2[bx][by][1], T[3 * bz] * t2[bx][by][0]));
              double *dst = row + (size_t)s_pos[base + bx + 5 * by + 25 * bz] * VH_SYMP;
              if (!((first >> (bx + 3 * by + 9 * bz)) & 1u))
                v += *dst;
              *dst = v;
            }
    }
}

// ------------------------------------------------------------------------------------------------
// 2c. row-owner kernel with the contraction on the FP64 tensor cores (mma.sync m8n8k4 -> SASS DMMA.8x8x4).
//     For one incident cell o the row's contribution is the GEMM  out[b][entry] = sum_q W_o[b][q] * H_o[q][entry]
//     with M = 8 column nodes b, K = 8 quadrature points (two k-steps of 4), N = 172 packed entries (22 tiles of 8).
//     On B200 DMMA has the same FLOP/s as DFMA (measured 37.1 vs 36.7 TFLOP/s, tools/dmma_peak.cu) but needs 8x fewer
//     issue slots, which is what the scalar kernel is short of.
//     Row permutation: MMA row m of octant o holds column node b = m XOR o, so that row m always accumulates slots of
//     parity m (slot coordinate s_d = 1 if m_d else 2*o_d): contributions of different cells to one slot stay in the
//     same lane and are reduced in registers; no cross-lane traffic, no atomics.  C fragments: one set per octant.
//     6 warps x 4 tiles; A fragments (weights) come from a 4 KB shared table, B fragments from the TMA ring.
// ------------------------------------------------------------------------------------------------
#define VH_MMA_THREADS 192
#define VH_MMA_STAGES 8 /* all incident cells' tables in flight at once: 88 KB ring, two CTAs per SM */
#define VH_MMA_SMEM (VH_MMA_STAGES * VH_CELL_H_BYTES + 272 * 8 + 512 * 8 + VH_MMA_STAGES * 8 + 64 * 4)

__global__ void __launch_bounds__(VH_MMA_THREADS, 2)
  k_rows_mma_q1(const int32_t *__restrict__ fast_rows, const int32_t *__restrict__ fast_cells, const int8_t *__restrict__ fast_slot,
                const int32_t *__restrict__ fast_class, const double *__restrict__ class_M, const double *__restrict__ afrag,
                const int32_t *__restrict__ row_ptr, const int32_t *__restrict__ col, const uint32_t *__restrict__ dirmask,
                const double *__restrict__ Hq, const double *__restrict__ Dc, const double *__restrict__ avgD,
                double *__restrict__ vals)
{
  extern __shared__ __align__(128) unsigned char smraw[];
  double   *s_H     = reinterpret_cast<double *>(smraw);                                   // [8][8*172], later [27][172]
  double   *s_cls   = reinterpret_cast<double *>(smraw + VH_MMA_STAGES * VH_CELL_H_BYTES); // [27][10] (+2)
  double   *s_A     = s_cls + 272;                                                         // [8 o][2 kstep][32 lanes]
  uint64_t *s_bar   = reinterpret_cast<uint64_t *>(s_A + 512);                             // [8]
  int      *s_cells = reinterpret_cast<int *>(s_bar + VH_MMA_STAGES);                      // [8]
  int      *s_pos   = s_cells + 8;                                                         // [27] (+1)
  uint32_t *s_maskJ = reinterpret_cast<uint32_t *>(s_pos + 28);                            // [27]

  constexpr int NT = VH_MMA_THREADS;
  const int     t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int     r = blockIdx.x;
  const int     I = fast_rows[r];
  if (t == 0)
    { // TMA producer, first thing in the CTA: every incident cell's table (11 KB each) is requested at once
#pragma unroll
      for (int k = 0; k < VH_MMA_STAGES; ++k)
        mbar_init(s_bar + k, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
#pragma unroll
      for (int o = 0; o < 8; ++o)
        {
          const int e = fast_cells[(size_t)r * 8 + o];
          if (e >= 0)
            {
              mbar_expect_tx(s_bar + o, VH_CELL_H_BYTES);
              bulk_g2s(s_H + (size_t)o * (8 * VH_SYMP), Hq + (size_t)e * (8 * VH_SYMP), VH_CELL_H_BYTES, s_bar + o);
            }
        }
    }
  // all row metadata is fetched now, behind the TMA latency: nothing global is touched again until the stores
  const uint32_t maskI = dirmask[I];
  const int      rp    = row_ptr[I];
  if (t >= 64 && t < 72)
    s_cells[t - 64] = fast_cells[(size_t)r * 8 + (t - 64)];
  if (t >= 32 && t < 59)
    {
      const int pos     = fast_slot[(size_t)r * 32 + (t - 32)];
      s_pos[t - 32]     = pos;
      s_maskJ[t - 32]   = pos >= 0 ? dirmask[col[rp + pos]] : 0u;
    }
  {
    const double *cls = class_M + (size_t)fast_class[r] * 270;
    for (int i = t; i < 270; i += NT)
      s_cls[i] = cls[i];
    for (int i = t; i < 512; i += NT)
      s_A[i] = afrag[i];
  }
  __syncthreads();

  // B-fragment offsets of this lane inside a cell table: row q = 4*kstep + lane%4, column = packed entry of tile + lane/4
  int boff[2][4];
#pragma unroll
  for (int ks = 0; ks < 2; ++ks)
#pragma unroll
    for (int tl = 0; tl < 4; ++tl)
      {
        const int e  = 8 * (4 * warp + tl) + (lane >> 2);
        boff[ks][tl] = vh_hq8_index(4 * ks + (lane & 3), e < VH_SYMP ? e : 18); // out-of-range columns read the zero dummy (1,0)
      }
  double acc[8][4][2];
#pragma unroll
  for (int o = 0; o < 8; ++o)
#pragma unroll
    for (int tl = 0; tl < 4; ++tl)
      acc[o][tl][0] = acc[o][tl][1] = 0.0;

#pragma unroll
  for (int o = 0; o < 8; ++o)
    {
      const int e = s_cells[o];
      if (e >= 0)
        {
          mbar_wait(s_bar + o, 0u);
          const double *Hs = s_H + (size_t)o * (8 * VH_SYMP);
#pragma unroll
          for (int ks = 0; ks < 2; ++ks)
            {
              const double a = s_A[(o * 2 + ks) * 32 + lane];
#pragma unroll
              for (int tl = 0; tl < 4; ++tl)
                {
                  const double b = Hs[boff[ks][tl]];
                  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                               : "+d"(acc[o][tl][0]), "+d"(acc[o][tl][1])
                               : "d"(a), "d"(b));
                }
            }
        }
    }

  // ---- reduce the octant sets inside each lane and dump the packed accumulators: [27][172] in the idle TMA ring ----
  __syncthreads();
  double   *s_sym = s_H;
  const int m = lane >> 2; // MMA row = slot parity class (m_x, m_y, m_z)
#pragma unroll
  for (int d = 0; d < 3; ++d)
    if ((m >> d) & 1)
      { // coordinate d of the slot is 1 whatever the octant: octants o and o|(1<<d) feed the same slot
#pragma unroll
        for (int o = 0; o < 8; ++o)
          if (!((o >> d) & 1))
#pragma unroll
            for (int tl = 0; tl < 4; ++tl)
              {
                acc[o][tl][0] += acc[o | (1 << d)][tl][0];
                acc[o][tl][1] += acc[o | (1 << d)][tl][1];
              }
      }
#pragma unroll
  for (int o = 0; o < 8; ++o)
    if ((o & m) == 0)
      {
        const int sx = (m & 1) ? 1 : 2 * (o & 1), sy = (m & 2) ? 1 : 2 * ((o >> 1) & 1), sz = (m & 4) ? 1 : 2 * (o >> 2);
        double   *dst = s_sym + (sx + 3 * sy + 9 * sz) * VH_SYMP;
#pragma unroll
        for (int tl = 0; tl < 4; ++tl)
          {
            const int e = 8 * (4 * warp + tl) + 2 * (lane & 3);
            if (e < VH_SYMP)
              *reinterpret_cast<double2 *>(dst + e) = make_double2(acc[o][tl][0], acc[o][tl][1]);
          }
      }
  __syncthreads();

  // ---- store: thread t < 162 owns double2 #t of every 18x18 block ----
  if (t < VH_BLK / 2)
    {
      int      soff[2], gsel[2];
      uint32_t rbit[2], cbit[2];
#pragma unroll
      for (int k = 0; k < 2; ++k)
        {
          const int c = (2 * t) / 18, d = (2 * t) % 18 + k;
          soff[k] = c <= d ? vh_sym_index(c, d) : vh_sym_index(d, c);
          gsel[k] = (c / 3 == d / 3) ? (c % 3) * 3 + d % 3 : 9;
          rbit[k] = 1u << c;
          cbit[k] = 1u << d;
        }
      // fully unrolled: the 108 shared-memory loads of the 27 slots are batched ahead of the adds and the stores
      double2 v[27];
#pragma unroll
      for (int s = 0; s < 27; ++s)
        {
          const double *sy = s_sym + s * VH_SYMP;
          const double *G  = s_cls + s * 10;
          v[s]             = make_double2(sy[soff[0]] + G[gsel[0]], sy[soff[1]] + G[gsel[1]]);
        }
#pragma unroll
      for (int s = 0; s < 27; ++s)
        {
          const int pos = s_pos[s];
          if (pos < 0)
            continue; // block-uniform
          const uint32_t maskJ = s_maskJ[s];
          double         v0 = v[s].x, v1 = v[s].y;
          if ((maskI | maskJ) != 0u)
            {
              if ((maskI & rbit[0]) || (maskJ & cbit[0]))
                v0 = 0.0;
              if ((maskI & rbit[1]) || (maskJ & cbit[1]))
                v1 = 0.0;
              if (s == 13)
                {
#pragma unroll
                  for (int k = 0; k < 2; ++k)
                    if (rbit[k] == cbit[k] && (maskI & rbit[k]))
                      { // constrained diagonal: sum over cells of |a_ii| (mean |diag| of the cell if a_ii == 0)
                        const int c    = 31 - __clz(rbit[k]);
                        double    dsum = 0.0;
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
                        if (k == 0)
                          v0 = dsum;
                        else
                          v1 = dsum;
                      }
                }
            }
          __stcs(reinterpret_cast<double2 *>(vals + (size_t)(rp + pos) * VH_BLK) + t, make_double2(v0, v1));
        }
    }
}

// Store-bandwidth probe with the row kernel's write pattern (one CTA per block row, 16-byte stores, no other work):
// what the write-once matrix store costs on its own.  Used by vh_time_kernel(what=7) only.
__global__ void __launch_bounds__(192) k_store_probe(int n_rows, const int32_t *__restrict__ row_ptr, double *__restrict__ vals, int mode)
{
  const int r = blockIdx.x;
  if (r >= n_rows)
    return;
  double2      *dst = reinterpret_cast<double2 *>(vals + (size_t)row_ptr[r] * VH_BLK);
  const int     n2  = (row_ptr[r + 1] - row_ptr[r]) * (VH_BLK / 2);
  const double2 v   = make_double2(1.0, 2.0);
  for (int i = threadIdx.x; i < n2; i += blockDim.x)
    {
      if (mode == 0)
        __stcs(dst + i, v);
      else
        dst[i] = v;
    }
}

// rhs gather of the lattice rows and the gather of the matrix-free apply: csrc/vh_gather_kernels.cuh
#include "vh_gather_kernels.cuh"

// Expansion of the packed rows into full 18x18 blocks (export / diagnostics): block = Sym(P) + kron(I_6, M_slot), Dirichlet
// rows and columns zeroed, constrained diagonal from cdiag.
__global__ void k_expand_packed(int n_fast, int ps_stride, int cm_stride, const int32_t *__restrict__ fast_rows,
                                const uint8_t *__restrict__ fast_posslot,
                                const int32_t *__restrict__ fast_class, const double *__restrict__ class_M,
                                const int32_t *__restrict__ row_ptr, const int32_t *__restrict__ col,
                                const uint32_t *__restrict__ dirmask, const double *__restrict__ pvals,
                                const double *__restrict__ cdiag, double *__restrict__ full)
{
  const int r = blockIdx.x;
  if (r >= n_fast)
    return;
  const int      I = fast_rows[r], rp = row_ptr[I], nb = row_ptr[I + 1] - rp;
  const uint32_t maskI = dirmask[I];
  const double  *M0 = class_M + (size_t)fast_class[r] * cm_stride;
  for (int i = threadIdx.x; i < nb * VH_BLK; i += blockDim.x)
    {
      const int      pos = i / VH_BLK, e = i - VH_BLK * pos, c = e / 18, d = e - 18 * c;
      const int      J = col[rp + pos];
      const uint32_t maskJ = dirmask[J];
      const double  *P = pvals + (size_t)(rp + pos) * VH_SYMP;
      double         v = P[c <= d ? vh_sym_index(c, d) : vh_sym_index(d, c)];
      if (c / 3 == d / 3)
        v += M0[fast_posslot[(size_t)r * ps_stride + pos] * 10 + (c % 3) * 3 + d % 3];
      if (((maskI >> c) & 1u) || ((maskJ >> d) & 1u))
        v = 0.0;
      if (J == I && c == d && ((maskI >> c) & 1u))
        v = cdiag[(size_t)I * 18 + c];
      full[(size_t)(rp + pos) * VH_BLK + e] = v;
    }
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
              v += swq[q] * sN[a * NQ + q] * sN[b * NQ + q] * sH[NQ == 8 ? vh_hq8_index(q, sidx) : q * VH_SYMP + sidx]; // pre-scaled by vol
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

// ------------------------------------------------------------------------------------------------
// 3b. row-owner assembly of the constrained rows (hanging-node neighbourhoods, constraint masters)
// ------------------------------------------------------------------------------------------------
// One CTA per owned row node I that the lattice kernels do not take.  The host lists the (cell, local node a) pairs that
// feed the row: a is I itself or a hanging node whose constraint line names I as a master (weight w).  Thread (c,d) owns
// entry (c,d) of every block of the row: for each pair it contracts the cell's packed H_q table (staged in shared memory)
// with the weights of all column nodes b, adds the gradient / Robin terms and applies deal.II's
// distribute_local_to_global rule: row weight of (a,c) -> (I,c), column b either direct or spread over
// the masters of its line, constrained rows reduced to  sum_cells |a_ii|  on the diagonal.  The row belongs to this CTA
// alone and entry (c,d) to one thread, so the accumulation is a plain read-modify-write on pre-zeroed blocks: no atomics.
// (Constraint lines that couple different components - none are generated by the hosts in this repository - would let
// two threads meet in one entry; such contexts keep the atomic cell scatter k_cells_slow.)
struct SlowRowArgs
{
  const int32_t *slow_rows, *srow_ptr, *srow_cell;
  const int8_t  *srow_a;       // local node of the pair | 64 if that node is the row node itself
  const int16_t *srow_posb;    // [pairs][nn] position of the cell's node b in the row (-1: not a column)
  const double  *srow_wr;      // [pairs][18] weight of local row (a,c) in global row (I,c)
  const uint32_t *srow_bcons;  // [pairs]     bit b: node b of the cell has constrained DoFs
  const int32_t *srow_mnode;   // [pairs][nn][MAXM] master nodes of the constrained column node b (-1 padded), MAXM = 4 (Q1) / 9 (Q2)
  const int16_t *srow_mpos;    // [pairs][nn][MAXM] their positions in the row
  const int32_t *srow_posI;    // [rows]      position of the diagonal block
  const uint32_t *srow_cons;   // [rows]      constrained components of the row node
  const int32_t *cell_nodes, *row_ptr, *col;
  const double  *cell_h;
  const uint32_t *cell_faces;
  const int32_t *line_of, *cptr, *cmaster;
  const double  *cweight;
  const double  *Hq, *Rc, *Dc, *avgD;
};

// rhs / residual of the constrained rows: one thread per (row, component), deterministic gather (constrained DoFs stay 0)
__global__ void k_rhs_slow(int n_slow_rows, int nn, SlowRowArgs A, double *__restrict__ rhs)
{
  const int gid = blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= n_slow_rows * 18)
    return;
  const int r = gid / 18, c = gid - 18 * r, I = A.slow_rows[r];
  double    s = 0.0;
  for (int k = A.srow_ptr[r]; k < A.srow_ptr[r + 1]; ++k)
    {
      const int64_t e = A.srow_cell[k];
      const int     a = A.srow_a[k] & 63;
      const double  w = A.srow_wr[(size_t)k * 18 + c];
      if (w != 0.0)
        s += w * A.Rc[e * (18 * nn) + a * 18 + c];
    }
  rhs[(size_t)I * 18 + c] = s;
}

// Per (row, pair) metadata is resolved on the host (positions of the cell's nodes in the row, row weights, which column
// nodes carry constrained DoFs), so the device code has no dependent index loads; the next pair's H_q table and
// metadata are fetched into registers while the current pair is being contracted.
template <int NN, int NQ>
__global__ void __launch_bounds__(352, NN == 8 ? 2 : 1)
  k_rows_slow(SlowRowArgs A, VhTables tab, VhCoef cf, double *__restrict__ vals)
{
  constexpr int NT = 352, HPT = (NQ * VH_SYMP + NT - 1) / NT; // table doubles per thread
  extern __shared__ double sm[];
  double *sH = sm;                 // [NQ * 180] packed H_q table of the current cell
  double *sN = sH + NQ * VH_SYMP;  // [NN][NQ]   shape values
  double *sW = sN + NN * NQ;       // [NQ]       quadrature weights
  __shared__ int    s_posb[NN], s_nodes[NN];
  __shared__ double s_wr[18];
  constexpr int MAXM = NN == 8 ? 4 : 9;
  __shared__ int   s_mnode[NN * MAXM];
  __shared__ short s_mpos[NN * MAXM];
  __shared__ double s_geo[NN * 12]; // per column node b: M[3][3] = vol K23 G_xy/(h_x h_y), then D[3] = K1 trace part + Robin
  const int      t = threadIdx.x, r = blockIdx.x;
  const int      I = A.slow_rows[r], rp = A.row_ptr[I];
  const int      c = t / 18, d = t - 18 * c;
  const bool     ent = t < VH_BLK;
  const int      sidx = ent ? (c <= d ? vh_sym_index(c, d) : vh_sym_index(d, c)) : 0;
  const int      gsel = (ent && c / 3 == d / 3) ? (c % 3) * 3 + d % 3 : -1;
  const int      posI = A.srow_posI[r];
  const uint32_t consI = A.srow_cons[r]; // components of node I that are constrained
  for (int i = t; i < NN * NQ; i += NT)
    sN[i] = tab.N[i];
  if (t < NQ)
    sW[t] = tab.wq[t];
  const int k0 = A.srow_ptr[r], k1 = A.srow_ptr[r + 1];
  // prefetch registers: table slice, per-node metadata, and the pair's scalars (every thread keeps its own copy)
  double   pH[HPT];
  int      p_posb = -1, p_node = 0, p_a = 0;
  int64_t  p_e = 0;
  double   p_wr = 0.0, p_h0 = 1.0, p_h1 = 1.0, p_h2 = 1.0, p_vol = 0.0;
  uint32_t p_bcons = 0u, p_faces = 0u;
  auto fetch = [&](int k) {
    p_e = A.srow_cell[k];
#pragma unroll
    for (int i = 0; i < HPT; ++i)
      pH[i] = (t + i * NT < NQ * VH_SYMP) ? __ldg(A.Hq + p_e * (NQ * VH_SYMP) + t + i * NT) : 0.0;
    if (t < NN)
      {
        p_posb = A.srow_posb[(size_t)k * NN + t];
        p_node = A.cell_nodes[p_e * NN + t];
      }
    if (t >= 32 && t < 50)
      p_wr = A.srow_wr[(size_t)k * 18 + (t - 32)];
    p_a     = A.srow_a[k];
    p_bcons = A.srow_bcons[k];
    p_faces = A.cell_faces[p_e];
    p_h0 = A.cell_h[4 * p_e], p_h1 = A.cell_h[4 * p_e + 1], p_h2 = A.cell_h[4 * p_e + 2], p_vol = A.cell_h[4 * p_e + 3];
  };
  if (k0 < k1)
    fetch(k0);
  for (int k = k0; k < k1; ++k)
    {
      const int64_t  e = p_e;
      const int      a = p_a & 63;
      const bool     self = p_a & 64;   // the pair's node is I itself
      const uint32_t bcons = p_bcons;   // column nodes with constrained DoFs
      const uint32_t faces = p_faces;
      const double   h[3] = {p_h0, p_h1, p_h2}, vol = p_vol;
      __syncthreads(); // everybody is done with the previous pair's shared data
#pragma unroll
      for (int i = 0; i < HPT; ++i)
        if (t + i * NT < NQ * VH_SYMP)
          sH[t + i * NT] = pH[i];
      if (t < NN)
        {
          s_posb[t]  = p_posb;
          s_nodes[t] = p_node;
        }
      if (t >= 32 && t < 50)
        s_wr[t - 32] = p_wr;
      if (bcons)
        for (int i = t; i < NN * MAXM; i += NT)
          {
            s_mnode[i] = A.srow_mnode[(size_t)k * (NN * MAXM) + i];
            s_mpos[i]  = A.srow_mpos[(size_t)k * (NN * MAXM) + i];
          }
      for (int i = t; i < NN * 12; i += NT)
        { // geometry-only part of the cell matrix for row node a and every column node b (gradient forms + Robin faces)
          const int     b = i / 12, j = i - 12 * b;
          const double *G = tab.Gref + (size_t)(a * NN + b) * 9;
          double        v;
          if (j < 9)
            v = vol * cf.K23 * G[j] / (h[j / 3] * h[j % 3]);
          else
            {
              const int x = j - 9;
              v = vol * cf.K1 * (G[0] / (h[0] * h[0]) + G[4] / (h[1] * h[1]) + G[8] / (h[2] * h[2]));
              if ((cf.bt < 1e10) && faces != 0u)
                for (int f = 0; f < 6; ++f)
                  {
                    const int bid = (faces >> (4 * f)) & 15u;
                    if (bid < 2 || bid > 4 || x == bid - 2)
                      continue;
                    v += cf.K1 / cf.bt * (vol / h[f / 2]) * tab.Mf[(size_t)(f * NN + a) * NN + b];
                  }
            }
          s_geo[i] = v;
        }
      __syncthreads();
      if (k + 1 < k1)
        fetch(k + 1);
      if (!ent)
        continue;
      if (c == d && self && ((consI >> c) & 1u) && posI >= 0)
        { // constrained DoF: its row keeps only  sum_cells |a_ii|  (the cell mean |diag| if a_ii == 0)
          double dv = fabs(A.Dc[e * (18 * NN) + a * 18 + c]);
          if (dv == 0.0)
            dv = A.avgD[e];
          vals[(size_t)(rp + posI) * VH_BLK + t] += dv;
        }
      const double wr = s_wr[c];
      if (wr == 0.0)
        continue;
      double       hq[NQ]; // w_q N_a(q) (vol H_q)[c][d]
#pragma unroll
      for (int q = 0; q < NQ; ++q)
        hq[q] = sW[q] * sN[a * NQ + q] * sH[NQ == 8 ? vh_hq8_index(q, sidx) : q * VH_SYMP + sidx];
      // the old values of this thread's entry in the blocks of the unconstrained column nodes: all loads in flight at once
      // (the row belongs to this CTA, the entry to this thread: plain read-modify-write)
      double old[NN == 8 ? 8 : 1];
      if constexpr (NN == 8)
        {
#pragma unroll
          for (int b = 0; b < 8; ++b)
            {
              const int pos = s_posb[b];
              old[b] = (pos >= 0 && !((bcons >> b) & 1u)) ? vals[(size_t)(rp + pos) * VH_BLK + t] : 0.0;
            }
        }
#pragma unroll(NN == 8 ? 8 : 1)
      for (int b = 0; b < NN; ++b)
        {
          double v = 0.0;
#pragma unroll
          for (int q = 0; q < NQ; ++q)
            v = fma(sN[b * NQ + q], hq[q], v);
          v += gsel >= 0 ? s_geo[b * 12 + gsel] : 0.0;
          if (c == d)
            v += s_geo[b * 12 + 9 + c % 3];
          v *= wr;
          if constexpr (NN == 8)
            {
              if (!((bcons >> b) & 1u))
                { // all 18 DoFs of node b unconstrained: its block takes the preloaded value (first pass)
                  const int pos = s_posb[b];
                  if (pos >= 0)
                    vals[(size_t)(rp + pos) * VH_BLK + t] = old[b] + v;
                }
              else
                old[b] = v; // second pass below: masters may share a block with a first-pass node
            }
          else
            {
              const int lj = ((bcons >> b) & 1u) ? A.line_of[18 * s_nodes[b] + d] : -1;
              if (lj < 0)
                {
                  const int pos = s_posb[b];
                  if (pos >= 0)
                    vals[(size_t)(rp + pos) * VH_BLK + t] += v;
                }
              else
                for (int p = A.cptr[lj]; p < A.cptr[lj + 1]; ++p)
                  { // constrained column: spread over its masters (same component: checked on the host)
                    const int md = A.cmaster[p], J = md / 18;
                    int       pos = -1;
#pragma unroll
                    for (int m = 0; m < MAXM; ++m)
                      if (s_mnode[b * MAXM + m] == J)
                        pos = s_mpos[b * MAXM + m];
                    if (pos >= 0)
                      vals[(size_t)(rp + pos) * VH_BLK + c * 18 + md % 18] += A.cweight[p] * v;
                  }
            }
        }
      if constexpr (NN == 8)
        if (bcons)
#pragma unroll
          for (int b = 0; b < 8; ++b)
            if ((bcons >> b) & 1u)
              {
                const double v  = old[b];
                const int    lj = A.line_of[18 * s_nodes[b] + d];
                if (lj < 0)
                  {
                    const int pos = s_posb[b];
                    if (pos >= 0)
                      vals[(size_t)(rp + pos) * VH_BLK + t] += v;
                  }
                else
                  for (int p = A.cptr[lj]; p < A.cptr[lj + 1]; ++p)
                    {
                      const int md = A.cmaster[p], J = md / 18;
                      int       pos = -1;
#pragma unroll
                      for (int m = 0; m < MAXM; ++m)
                        if (s_mnode[b * MAXM + m] == J)
                          pos = s_mpos[b * MAXM + m];
                      if (pos >= 0)
                        vals[(size_t)(rp + pos) * VH_BLK + c * 18 + md % 18] += A.cweight[p] * v;
                    }
              }
    }
}

template <int NN, int NQ>
size_t pointwise_smem(bool want_h)
{
  (void)want_h;
  size_t n = (size_t)NN * 18 + (size_t)NQ * VH_TQ + NQ * 54 + NQ * 18 + NQ * 18 + NN * NQ + NN * NQ * 3 + NQ + 32;
  return n * sizeof(double);
}
} // namespace

int vhk_upload_constants(vh_ctx *ctx)
{
  // symmetric packing tables
  uint8_t sc[VH_SYMP], sd[VH_SYMP];
  vh_sym_tables(sc, sd); // dummies (d < c) are written as 0 by the pointwise kernel and never read back as entries
  VH_CUDA(cudaMemcpyToSymbol(c_symc, sc, sizeof(sc)));
  VH_CUDA(cudaMemcpyToSymbol(c_symd, sd, sizeof(sd)));
  return VH_OK;
}

int vhk_upload_q2(vh_ctx *ctx, const double *T2, const uint8_t *q2t)
{
  VH_CUDA(cudaMemcpyToSymbol(c_T2, T2, 27 * sizeof(double)));
  VH_CUDA(cudaMemcpyToSymbol(c_q2t, q2t, 81));
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
  const bool legacy_q2 = getenv("VH_Q2_POINTWISE_LEGACY") && getenv("VH_Q2_POINTWISE_LEGACY")[0] == '1';
  const vh_hweights hw = vh_make_hweights(ctx->coef.alpha, ctx->coef.beta);
#define VH_LAUNCH_POINTS(NN, H, E)                                                                                                \
  k_points<NN, H, E><<<grid, VH_PT_WARPS * 32, VhPt<NN>::SMEM, ctx->stream>>>(ctx->n_cells, ctx->cell_nodes, ctx->cell_h,         \
                                                                             ctx->cell_faces, ctx->cell_owned, x_local, ctx->tab, \
                                                                             ctx->coef, hw, ctx->Hq, ctx->Rc, ctx->Dc, ctx->avgD, \
                                                                             ctx->Ec)
#define VH_LAUNCH_POINTS_MODE(NN)      \
  if (want_h && want_e)                \
    VH_LAUNCH_POINTS(NN, true, true);  \
  else if (want_h)                     \
    VH_LAUNCH_POINTS(NN, true, false); \
  else if (want_e)                     \
    VH_LAUNCH_POINTS(NN, false, true); \
  else                                 \
    VH_LAUNCH_POINTS(NN, false, false)
#define VH_POINTS_ATTR(NN)                                                                                                        \
  VH_CUDA(cudaFuncSetAttribute(k_points<NN, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)VhPt<NN>::SMEM));     \
  VH_CUDA(cudaFuncSetAttribute(k_points<NN, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)VhPt<NN>::SMEM));      \
  VH_CUDA(cudaFuncSetAttribute(k_points<NN, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)VhPt<NN>::SMEM));    \
  VH_CUDA(cudaFuncSetAttribute(k_points<NN, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)VhPt<NN>::SMEM))
  static unsigned long long attr_mask = 0;
  if (vh_first_time_on_device(attr_mask, ctx->device))
    {
      VH_POINTS_ATTR(8);
      VH_POINTS_ATTR(27);
    }
  if (ctx->degree == 1)
    {
      const int grid = (ctx->n_cells + 4 * VH_PT_WARPS - 1) / (4 * VH_PT_WARPS);
      VH_LAUNCH_POINTS_MODE(8);
    }
  else if (!legacy_q2)
    {
      const int grid = (ctx->n_cells + VH_PT_WARPS - 1) / VH_PT_WARPS;
      VH_LAUNCH_POINTS_MODE(27);
    }
#undef VH_LAUNCH_POINTS
#undef VH_LAUNCH_POINTS_MODE
#undef VH_POINTS_ATTR
  else
    {
      const size_t smem = pointwise_smem<27, 27>(want_h);
      static unsigned long long attr_mask2 = 0;
      if (vh_first_time_on_device(attr_mask2, ctx->device))
        VH_CUDA(cudaFuncSetAttribute(k_pointwise<27, 27>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)pointwise_smem<27, 27>(true)));
      k_pointwise<27, 27><<<ctx->n_cells, 512, smem, ctx->stream>>>(ctx->cell_nodes, ctx->cell_h, ctx->cell_faces,
                                                                   ctx->cell_owned, x_local, ctx->tab, ctx->coef, want_h, want_e,
                                                                   ctx->Hq, ctx->Rc, ctx->Dc, ctx->avgD, ctx->Ec);
    }
  VH_LAUNCH_CHECK();
  return VH_OK;
}

// Matrix-free apply of the lattice rows: y_fast = A z without touching the assembled blocks.  z_masked: local vector
// (owned + ghosts) with zeros at the Dirichlet DoFs; x_orig: the unmasked vector (constrained-diagonal term only).
// Uses ctx->Rc as the per-cell scratch: the cell rhs it holds after vh_assemble / vh_residual has been gathered by then.
int vhk_apply_fast(vh_ctx *ctx, const double *z_masked, const double *x_orig, double *y_owned)
{
  if (ctx->n_fast == 0 || ctx->n_cells == 0)
    return VH_OK;
  if (ctx->spmv_mf_v2 && ctx->degree == 1)
    { // second formulation (Q1): lane = point for the bulk part, lane = node with a per-cell geometry table for the rest
      k_apply_q1_v2<<<(ctx->n_cells + 4 * VH_V2_WARPS - 1) / (4 * VH_V2_WARPS), VH_V2_WARPS * 32, 0, ctx->stream>>>(
        ctx->n_cells, ctx->cell_nodes, ctx->cell_h, ctx->cell_faces, z_masked, ctx->tab, ctx->coef, ctx->Hq, ctx->Rc);
      VH_LAUNCH_CHECK();
      const int64_t n0 = (int64_t)ctx->n_fast * 18;
      k_gather_apply<<<(unsigned)((n0 + 255) / 256), 256, 0, ctx->stream>>>(ctx->n_fast, ctx->dpc, ctx->fast_rows, ctx->fast_cells, ctx->fast_a,
                                                                           ctx->dirmask, ctx->Rc, ctx->cdiag, x_orig, y_owned);
      VH_LAUNCH_CHECK();
      return VH_OK;
    }
  if (ctx->spmv_mf_table_free)
    { // table-free variant: the Jacobian is the one of the state the last vh_assemble saw (ctx->x_sol until vh_accept_trial,
      // which invalidates the matrix anyway)
      const vh_hweights hw0 = vh_make_hweights(ctx->coef.alpha, ctx->coef.beta);
      static unsigned long long tf_mask = 0;
      if (vh_first_time_on_device(tf_mask, ctx->device))
        {
          VH_CUDA(cudaFuncSetAttribute(k_points<8, false, false, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)VhPt<8>::SMEM_TFREE));
          VH_CUDA(cudaFuncSetAttribute(k_points<27, false, false, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)VhPt<27>::SMEM_TFREE));
        }
      if (ctx->degree == 1)
        k_points<8, false, false, true, true><<<(ctx->n_cells + 4 * VH_PT_WARPS - 1) / (4 * VH_PT_WARPS), VH_PT_WARPS * 32, VhPt<8>::SMEM_TFREE, ctx->stream>>>(
          ctx->n_cells, ctx->cell_nodes, ctx->cell_h, ctx->cell_faces, ctx->cell_owned, z_masked, ctx->tab, ctx->coef, hw0, ctx->Hq, ctx->Rc,
          ctx->Dc, ctx->avgD, ctx->Ec, ctx->x_sol);
      else
        k_points<27, false, false, true, true><<<(ctx->n_cells + VH_PT_WARPS - 1) / VH_PT_WARPS, VH_PT_WARPS * 32, VhPt<27>::SMEM_TFREE, ctx->stream>>>(
          ctx->n_cells, ctx->cell_nodes, ctx->cell_h, ctx->cell_faces, ctx->cell_owned, z_masked, ctx->tab, ctx->coef, hw0, ctx->Hq, ctx->Rc,
          ctx->Dc, ctx->avgD, ctx->Ec, ctx->x_sol);
      VH_LAUNCH_CHECK();
      const int64_t n0 = (int64_t)ctx->n_fast * 18;
      k_gather_apply<<<(unsigned)((n0 + 255) / 256), 256, 0, ctx->stream>>>(ctx->n_fast, ctx->dpc, ctx->fast_rows, ctx->fast_cells, ctx->fast_a,
                                                                           ctx->dirmask, ctx->Rc, ctx->cdiag, x_orig, y_owned);
      VH_LAUNCH_CHECK();
      return VH_OK;
    }
  const vh_hweights hw = vh_make_hweights(ctx->coef.alpha, ctx->coef.beta); // unused by the apply
  static unsigned long long attr_mask = 0;
  if (vh_first_time_on_device(attr_mask, ctx->device))
    {
      VH_CUDA(cudaFuncSetAttribute(k_points<8, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)VhPt<8>::SMEM));
      VH_CUDA(cudaFuncSetAttribute(k_points<27, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)VhPt<27>::SMEM));
    }
  if (ctx->degree == 1)
    {
      const int grid = (ctx->n_cells + 4 * VH_PT_WARPS - 1) / (4 * VH_PT_WARPS);
      k_points<8, false, false, true><<<grid, VH_PT_WARPS * 32, VhPt<8>::SMEM, ctx->stream>>>(
        ctx->n_cells, ctx->cell_nodes, ctx->cell_h, ctx->cell_faces, ctx->cell_owned, z_masked, ctx->tab, ctx->coef, hw, ctx->Hq, ctx->Rc,
        ctx->Dc, ctx->avgD, ctx->Ec);
    }
  else
    {
      const int grid = (ctx->n_cells + VH_PT_WARPS - 1) / VH_PT_WARPS;
      k_points<27, false, false, true><<<grid, VH_PT_WARPS * 32, VhPt<27>::SMEM, ctx->stream>>>(
        ctx->n_cells, ctx->cell_nodes, ctx->cell_h, ctx->cell_faces, ctx->cell_owned, z_masked, ctx->tab, ctx->coef, hw, ctx->Hq, ctx->Rc,
        ctx->Dc, ctx->avgD, ctx->Ec);
    }
  VH_LAUNCH_CHECK();
  const int64_t n = (int64_t)ctx->n_fast * 18;
  k_gather_apply<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(ctx->n_fast, ctx->dpc, ctx->fast_rows, ctx->fast_cells, ctx->fast_a,
                                                                      ctx->dirmask, ctx->Rc, ctx->cdiag, x_orig, y_owned);
  VH_LAUNCH_CHECK();
  return VH_OK;
}

int vhk_diag_fast(vh_ctx *ctx)
{
  if (ctx->n_fast == 0 || ctx->n_cells == 0)
    return VH_OK;
  if (!ctx->Dblk)
    VH_TRY(vh_dev_alloc(ctx, &ctx->Dblk, (size_t)ctx->n_cells * ctx->nn * VH_SYMP));
  if (ctx->degree == 1)
    k_diag_cells<8><<<ctx->n_cells, 192, 0, ctx->stream>>>(ctx->n_cells, ctx->tab.N, ctx->tab.wq, ctx->Hq, ctx->Dblk);
  else
    k_diag_cells<27><<<ctx->n_cells, 192, 0, ctx->stream>>>(ctx->n_cells, ctx->tab.N, ctx->tab.wq, ctx->Hq, ctx->Dblk);
  VH_LAUNCH_CHECK();
  k_diag_gather<<<ctx->n_fast, 192, 0, ctx->stream>>>(ctx->n_fast, ctx->nn, ctx->fast_rows, ctx->fast_cells, ctx->fast_a, ctx->diag_pos,
                                                    ctx->Dblk, ctx->pvals);
  VH_LAUNCH_CHECK();
  return VH_OK;
}

int vhk_ensure_rows(vh_ctx *ctx)
{
  if (!ctx->rows_stale)
    return VH_OK;
  VH_TRY(vhk_rows_fast(ctx)); // from the H_q tables of the last vh_assemble (the pointwise kernel is not rerun)
  ctx->rows_stale = false;
  return VH_OK;
}

int vhk_rows_fast(vh_ctx *ctx)
{
  if (ctx->n_fast == 0)
    return VH_OK;
  if (ctx->degree == 2)
    {
      static int minb = 0;
      if (!minb)
        {
          const char *e = getenv("VH_Q2_ROWS_MINB"); // tuning knob: resident CTAs per SM the kernel is compiled for (2, 3 or 4)
          minb          = e ? atoi(e) : 4;
        }
      if (minb == 2)
        k_rows_fast_q2<2><<<ctx->n_fast, 192, 0, ctx->stream>>>(ctx->fast_rows, ctx->fast_cells, ctx->fast_a, ctx->fast_slot,
                                                               ctx->fast_first, ctx->row_ptr, ctx->Hq, ctx->pvals);
      else if (minb == 4)
        k_rows_fast_q2<4><<<ctx->n_fast, 192, 0, ctx->stream>>>(ctx->fast_rows, ctx->fast_cells, ctx->fast_a, ctx->fast_slot,
                                                               ctx->fast_first, ctx->row_ptr, ctx->Hq, ctx->pvals);
      else
        k_rows_fast_q2<3><<<ctx->n_fast, 192, 0, ctx->stream>>>(ctx->fast_rows, ctx->fast_cells, ctx->fast_a, ctx->fast_slot,
                                                               ctx->fast_first, ctx->row_ptr, ctx->Hq, ctx->pvals);
      VH_LAUNCH_CHECK();
      return VH_OK;
    }
  static int                ept = 0;
  static unsigned long long rows_attr_mask = 0;
  if (!ept)
    {
      const char *e = getenv("VH_ROWS_EPT"); // tuning knob: packed entries per thread (1 or 2)
      ept           = (e && e[0] == '2') ? 2 : 1;
    }
  if (vh_first_time_on_device(rows_attr_mask, ctx->device))
    {
      VH_CUDA(cudaFuncSetAttribute(k_rows_fast_q1<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, VH_FAST_SMEM));
      VH_CUDA(cudaFuncSetAttribute(k_rows_fast_q1<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, VH_FAST_SMEM));
      VH_CUDA(cudaFuncSetAttribute(k_rows_fast_q1<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, VH_FAST_SMEM));
      VH_CUDA(cudaFuncSetAttribute(k_rows_fast_q1<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, VH_FAST_SMEM));
    }
  static int use_mma = -1;
  if (use_mma < 0)
    {
      const char *e = getenv("VH_ROWS_MMA"); // tuning knob (full-format storage only): 1 = FP64 tensor-core kernel, 0 = scalar DFMA kernel (default, faster)
      use_mma       = (e && e[0] == '1') ? 1 : 0;
    }
  static unsigned long long mma_attr_mask = 0;
  if (use_mma && vh_first_time_on_device(mma_attr_mask, ctx->device))
    VH_CUDA(cudaFuncSetAttribute(k_rows_mma_q1, cudaFuncAttributeMaxDynamicSharedMemorySize, VH_MMA_SMEM));
  if (ctx->packed)
    { // packed symmetric storage: no expansion stage, half the bytes
      if (ept == 2)
        k_rows_fast_q1<2, true><<<ctx->n_fast, 96, VH_FAST_SMEM, ctx->stream>>>(ctx->fast_rows, ctx->fast_cells, ctx->fast_slot,
                                                                               ctx->fast_class, ctx->class_M, ctx->row_ptr, ctx->col,
                                                                               ctx->dirmask, ctx->Hq, ctx->Dc, ctx->avgD, ctx->coef,
                                                                               ctx->pvals);
      else
        k_rows_fast_q1<1, true><<<ctx->n_fast, 192, VH_FAST_SMEM, ctx->stream>>>(ctx->fast_rows, ctx->fast_cells, ctx->fast_slot,
                                                                                ctx->fast_class, ctx->class_M, ctx->row_ptr, ctx->col,
                                                                                ctx->dirmask, ctx->Hq, ctx->Dc, ctx->avgD, ctx->coef,
                                                                                ctx->pvals);
    }
  else if (use_mma)
    k_rows_mma_q1<<<ctx->n_fast, VH_MMA_THREADS, VH_MMA_SMEM, ctx->stream>>>(ctx->fast_rows, ctx->fast_cells, ctx->fast_slot,
                                                                           ctx->fast_class, ctx->class_M, ctx->afrag, ctx->row_ptr,
                                                                           ctx->col, ctx->dirmask, ctx->Hq, ctx->Dc, ctx->avgD,
                                                                           ctx->vals);
  else if (ept == 2)
    k_rows_fast_q1<2, false><<<ctx->n_fast, 96, VH_FAST_SMEM, ctx->stream>>>(ctx->fast_rows, ctx->fast_cells, ctx->fast_slot, ctx->fast_class,
                                                                     ctx->class_M, ctx->row_ptr, ctx->col, ctx->dirmask, ctx->Hq,
                                                                     ctx->Dc, ctx->avgD, ctx->coef, ctx->vals);
  else
    k_rows_fast_q1<1, false><<<ctx->n_fast, 192, VH_FAST_SMEM, ctx->stream>>>(ctx->fast_rows, ctx->fast_cells, ctx->fast_slot, ctx->fast_class,
                                                                      ctx->class_M, ctx->row_ptr, ctx->col, ctx->dirmask, ctx->Hq,
                                                                      ctx->Dc, ctx->avgD, ctx->coef, ctx->vals);
  VH_LAUNCH_CHECK();
  return VH_OK;
}

int vhk_expand_packed(vh_ctx *ctx, double *full_vals)
{
  if (!ctx->packed || ctx->n_fast == 0)
    return VH_OK;
  k_expand_packed<<<ctx->n_fast, 256, 0, ctx->stream>>>(ctx->n_fast, ctx->slot_stride, ctx->n_slots * 10, ctx->fast_rows, ctx->fast_posslot, ctx->fast_class, ctx->class_M,
                                                       ctx->row_ptr, ctx->col, ctx->dirmask, ctx->pvals, ctx->cdiag, full_vals);
  VH_LAUNCH_CHECK();
  return VH_OK;
}

int vhk_store_probe(vh_ctx *ctx, int mode)
{
  k_store_probe<<<ctx->n_owned, 192, 0, ctx->stream>>>(ctx->n_owned, ctx->row_ptr, ctx->vals, mode);
  VH_LAUNCH_CHECK();
  return VH_OK;
}

int vhk_rhs_fast(vh_ctx *ctx, double *rhs_out, bool with_cdiag)
{
  if (ctx->n_fast == 0)
    return VH_OK;
  const int64_t n = (int64_t)ctx->n_fast * 18;
  // the constrained-diagonal values are (re)computed whenever the Jacobian was (the cell diagonals Dc are fresh then)
  k_rhs_fast<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(ctx->n_fast, ctx->dpc, ctx->fast_rows, ctx->fast_cells, ctx->fast_a,
                                                                  ctx->dirmask, ctx->Rc, rhs_out, ctx->Dc, ctx->avgD,
                                                                  (ctx->packed && with_cdiag) ? ctx->cdiag : nullptr);
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
  if (ctx->slow_row_owner)
    { // row-owner kernel: no atomics (constraint lines stay within one component)
      SlowRowArgs R;
      R.slow_rows  = ctx->slow_rows;
      R.srow_ptr   = ctx->srow_ptr;
      R.srow_cell  = ctx->srow_cell;
      R.srow_a     = ctx->srow_a;
      R.srow_posb  = ctx->srow_posb;
      R.srow_wr    = ctx->srow_wr;
      R.srow_bcons = ctx->srow_bcons;
      R.srow_mnode = ctx->srow_mnode;
      R.srow_mpos  = ctx->srow_mpos;
      R.srow_posI  = ctx->srow_posI;
      R.srow_cons  = ctx->srow_cons;
      R.cell_nodes = ctx->cell_nodes;
      R.row_ptr    = ctx->row_ptr;
      R.col        = ctx->col;
      R.cell_h     = ctx->cell_h;
      R.cell_faces = ctx->cell_faces;
      R.line_of    = ctx->cons[0].line_of;
      R.cptr       = ctx->cons[0].ptr;
      R.cmaster    = ctx->cons[0].master;
      R.cweight    = ctx->cons[0].weight;
      R.Hq         = ctx->Hq;
      R.Rc         = ctx->Rc;
      R.Dc         = ctx->Dc;
      R.avgD       = ctx->avgD;
      k_rhs_slow<<<(ctx->n_slow_rows * 18 + 127) / 128, 128, 0, ctx->stream>>>(ctx->n_slow_rows, ctx->nn, R, rhs_out);
      VH_LAUNCH_CHECK();
      if (!want_matrix)
        return VH_OK;
      if (ctx->degree == 1)
        k_rows_slow<8, 8><<<ctx->n_slow_rows, 352, (8 * VH_SYMP + 64 + 8) * sizeof(double), ctx->stream>>>(R, ctx->tab, ctx->coef, ctx->vals);
      else
        {
          static unsigned long long attr_mask = 0;
          if (vh_first_time_on_device(attr_mask, ctx->device))
            VH_CUDA(cudaFuncSetAttribute(k_rows_slow<27, 27>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)((27 * VH_SYMP + 729 + 27) * sizeof(double))));
          k_rows_slow<27, 27><<<ctx->n_slow_rows, 352, (27 * VH_SYMP + 729 + 27) * sizeof(double), ctx->stream>>>(R, ctx->tab, ctx->coef,
                                                                                                                 ctx->vals);
        }
      VH_LAUNCH_CHECK();
      return VH_OK;
    }
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
