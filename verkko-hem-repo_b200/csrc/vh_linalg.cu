// FP64 Krylov building blocks on sm_100a: BSR(18) SpMV, nodal 18x18 block-Jacobi, fused vector kernels.
//
// These replace what deal.II/Trilinos do inside SolverFGMRES::solve (/root/reference/femgl/src/solve.cc:171-174):
// Epetra CRS SpMV, the preconditioner apply (ML-AMG in the reference; the north star prescribes nodal
// block-Jacobi), Vector::add_and_dot / l2_norm of the modified Gram-Schmidt loop, and
// AffineConstraints::distribute (solve.cc:181, iteration.cc:141,180).  All are HBM-bound; the design rules are
// 16-byte coalesced streaming of the 2592-byte blocks, x gathered through L1/L2, and deterministic
// two-stage reductions whose results stay on the device (no host round trip inside the Gram-Schmidt loop).
#include "vh_internal.h"
#include "vh_p2p.cuh"

#include <cstdlib>

#include <cooperative_groups.h>
namespace cg = cooperative_groups;

                  // SpMV row-end reduction: partial-sum slots per component

namespace
{
__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
    v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ------------------------------------------------------------------------------------------------
// SpMV: one warp per block row.  A block is 162 double2; lane l streams double2 l, l+32, ..., l+160.
// Element k of a block belongs to matrix row (2k)/18 = k/9 and columns 2(k%9), 2(k%9)+1, so for a fixed
// (lane, round) the matrix row is the same in every block of the block row: six private accumulators per lane,
// one segmented reduction per block row at the end.
// ------------------------------------------------------------------------------------------------
#define VH_SPMV_WARPS 8
__global__ void __launch_bounds__(VH_SPMV_WARPS * 32)
  k_spmv_bsr18(int n_rows, const int32_t *__restrict__ rows, const int32_t *__restrict__ row_ptr, const int32_t *__restrict__ col,
               const double *__restrict__ vals, const double *__restrict__ x, double *__restrict__ y)
{
  __shared__ double s_part[VH_SPMV_WARPS][6 * 32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int ridx = blockIdx.x * VH_SPMV_WARPS + wid;
  if (ridx >= n_rows)
    return;
  const int row = rows ? rows[ridx] : ridx; // rows != nullptr: only the general-scatter rows (packed storage elsewhere)
  const int b0 = row_ptr[row], b1 = row_ptr[row + 1];
  double    acc[6] = {0, 0, 0, 0, 0, 0};
  int       xoff[6];
#pragma unroll
  for (int r = 0; r < 6; ++r)
    xoff[r] = 2 * ((lane + 32 * r) % 9);
  const bool last = lane < 2; // round 5 covers double2 160,161 only
  for (int base = b0; base < b1; base += 32)
    {
      // the row's block-column indices, one per lane, handed out by shuffle (no dependent global load per block)
      const int nchunk = min(32, b1 - base);
      const int mycol  = lane < nchunk ? __ldg(col + base + lane) : 0;
      for (int j = 0; j < nchunk; j += 2)
        {
          const bool     two = j + 1 < nchunk;
          const double2 *B0  = reinterpret_cast<const double2 *>(vals + (size_t)(base + j) * VH_BLK);
          const double2 *B1  = B0 + (two ? VH_BLK / 2 : 0);
          double2        v0[6], v1[6];
          // 12 independent 16-byte streaming loads per lane in flight (every matrix byte is used once per SpMV)
#pragma unroll
          for (int r = 0; r < 5; ++r)
            v0[r] = __ldcs(B0 + lane + 32 * r);
          v0[5] = last ? __ldcs(B0 + lane + 160) : make_double2(0.0, 0.0);
#pragma unroll
          for (int r = 0; r < 5; ++r)
            v1[r] = two ? __ldcs(B1 + lane + 32 * r) : make_double2(0.0, 0.0);
          v1[5] = (two && last) ? __ldcs(B1 + lane + 160) : make_double2(0.0, 0.0);
          const double *x0 = x + 18 * (size_t)__shfl_sync(0xffffffffu, mycol, j);
          const double *x1 = x + 18 * (size_t)__shfl_sync(0xffffffffu, mycol, two ? j + 1 : j);
#pragma unroll
          for (int r = 0; r < 6; ++r)
            {
              const double2 xa = *reinterpret_cast<const double2 *>(x0 + xoff[r]);
              const double2 xb = *reinterpret_cast<const double2 *>(x1 + xoff[r]);
              acc[r]           = fma(v0[r].x, xa.x, fma(v0[r].y, xa.y, acc[r]));
              acc[r]           = fma(v1[r].x, xb.x, fma(v1[r].y, xb.y, acc[r]));
            }
        }
    }
#pragma unroll
  for (int r = 0; r < 6; ++r)
    s_part[wid][32 * r + lane] = acc[r];
  __syncwarp();
  if (lane < 18)
    {
      double s = 0.0;
#pragma unroll
      for (int k = 0; k < 9; ++k)
        s += s_part[wid][9 * lane + k];
      y[(size_t)row * 18 + lane] = s;
    }
}

// ------------------------------------------------------------------------------------------------
// SpMV over the PACKED rows: block = Sym(P) + kron(I_6, M_slot), Dirichlet columns/rows masked on the fly.
// One warp per row; a block is 90 double2 (1440 B, layout P180 of vh_pointwise.cuh): lane l streams double2 l, l+32,
// l+64.  A double2 holds (c,d),(c,d+1) with d even, so it needs ONE aligned 16-byte load of x_J[d..d+1] plus x_J[c]:
//     y_c += S0 x_d + S1 x_{d+1};   y_d += S0 x_c;   y_{d+1} += S1 x_c        (mirror terms, off the diagonal)
// (c,d) are lane constants, identical in every block: nine private partial sums per lane, gathered once per row through a
// constant index table.  The geometry part is three FMAs per block for lanes 0..17.  Four blocks are in flight.
// ------------------------------------------------------------------------------------------------
#define VH_PSPMV_WARPS 8
#define VH_PSPMV_NPART 9
// xg: vector the blocks are applied to.  Its Dirichlet DoFs must already be zero (true for every Krylov vector; the
// public vh_spmv masks a copy first), so no mask logic sits in the streaming loop.  xo: the unmasked vector, used only
// for the constrained diagonal  y_c = (sum_cells |a_ii|) x_c.
// lane_tab[3][32]: per (slot k, lane) packed  c | d_even<<8 | m0<<16 | m1<<17;  gather_tab[26][18]: partial-sum indices.
template <int NB, int MINB> // NB blocks in flight per warp, MINB resident CTAs per SM
__global__ void __launch_bounds__(VH_PSPMV_WARPS * 32, MINB)
  k_spmv_sym18(int n_fast, int ps_stride, int cm_stride, const int32_t *__restrict__ order, const int32_t *__restrict__ fast_rows,
               const uint8_t *__restrict__ fast_posslot,
               const int32_t *__restrict__ fast_class, const double *__restrict__ class_M, const int32_t *__restrict__ row_ptr,
               const int32_t *__restrict__ col, const uint32_t *__restrict__ dirmask, const double *__restrict__ pvals,
               const double *__restrict__ cdiag, const uint32_t *__restrict__ lane_tab, const uint16_t *__restrict__ gather_tab,
               const double *__restrict__ xg, const double *__restrict__ xo, double *__restrict__ y)
{
  __shared__ double s_part[VH_PSPMV_WARPS][VH_PSPMV_NPART * 32 + 1];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int r0   = blockIdx.x * VH_PSPMV_WARPS + wid;
  if (r0 >= n_fast)
    return;
  const int     r = order[r0];
  const int     I = fast_rows[r], b0 = row_ptr[I], b1 = row_ptr[I + 1];
  const double *M0 = class_M + (size_t)fast_class[r] * cm_stride;
  const bool    third = lane < (VH_SYMP / 2 - 64); // double2 #(lane+64) exists for lanes 0..25
  int           pc[3], pd[3];
  double        m0[3], m1[3];
#pragma unroll
  for (int k = 0; k < 3; ++k)
    {
      const uint32_t w = __ldg(lane_tab + 32 * k + lane);
      pc[k]            = w & 0xff;
      pd[k]            = (w >> 8) & 0xff;
      m0[k]            = (w >> 16) & 1u ? 1.0 : 0.0;
      m1[k]            = (w >> 17) & 1u ? 1.0 : 0.0;
    }
  double    ac[3] = {0, 0, 0}, a0[3] = {0, 0, 0}, a1[3] = {0, 0, 0};
  double    geo   = 0.0;
  const int gc = lane < 18 ? lane : 0, gg = 3 * (gc / 3), gx = gc % 3; // geometry: lane c, group base, orbital index

  // one block: 3 x 16 B of matrix, 3 x (16 B + 8 B) of x through L1, 12 FMAs (+3 for the geometry part in lanes 0..17)
#define VH_PSPMV_BLOCK(V, JJ, SL)                                                        \
  {                                                                                      \
    const double *xj = xg + 18 * (size_t)(JJ);                                           \
    _Pragma("unroll") for (int k = 0; k < 3; ++k)                                        \
    {                                                                                    \
      const double2 xd = *reinterpret_cast<const double2 *>(xj + pd[k]);                 \
      const double  xc = xj[pc[k]];                                                      \
      ac[k]            = fma((V)[k].x, xd.x, fma((V)[k].y, xd.y, ac[k]));                \
      a0[k]            = fma((V)[k].x, xc, a0[k]);                                       \
      a1[k]            = fma((V)[k].y, xc, a1[k]);                                       \
    }                                                                                    \
    if (lane < 18)                                                                       \
      {                                                                                  \
        const double *M = M0 + (SL)*10 + 3 * gx;                                         \
        geo             = fma(M[0], xj[gg], fma(M[1], xj[gg + 1], fma(M[2], xj[gg + 2], geo))); \
      }                                                                                  \
  }

  for (int base = b0; base < b1; base += 32)
    {
      const int nchunk = min(32, b1 - base);
      const int mycol  = lane < nchunk ? __ldg(col + base + lane) : 0;
      const int myslot = lane < nchunk ? (int)fast_posslot[(size_t)r * ps_stride + (base - b0) + lane] : 0;
      int       j      = 0;
      for (; j + NB <= nchunk; j += NB)
        { // NB blocks (NB x 3 x 16 B per lane) in flight, no tail logic here
          double2 v[NB][3];
#pragma unroll
          for (int u = 0; u < NB; ++u)
            {
              const double2 *B = reinterpret_cast<const double2 *>(pvals + (size_t)(base + j + u) * VH_SYMP);
              v[u][0]          = __ldcs(B + lane);
              v[u][1]          = __ldcs(B + lane + 32);
              v[u][2]          = third ? __ldcs(B + lane + 64) : make_double2(0.0, 0.0);
            }
#pragma unroll
          for (int u = 0; u < NB; ++u)
            {
              const int J  = __shfl_sync(0xffffffffu, mycol, j + u);
              const int sl = __shfl_sync(0xffffffffu, myslot, j + u);
              VH_PSPMV_BLOCK(v[u], J, sl)
            }
        }
      if (j < nchunk)
        { // tail (< NB blocks): one masked batch, so its loads are in flight together like in the main loop
          double2 v[NB][3];
#pragma unroll
          for (int u = 0; u < NB; ++u)
            {
              const bool     on = j + u < nchunk;
              const double2 *B  = reinterpret_cast<const double2 *>(pvals + (size_t)(base + (on ? j + u : j)) * VH_SYMP);
              v[u][0]           = on ? __ldcs(B + lane) : make_double2(0.0, 0.0);
              v[u][1]           = on ? __ldcs(B + lane + 32) : make_double2(0.0, 0.0);
              v[u][2]           = (on && third) ? __ldcs(B + lane + 64) : make_double2(0.0, 0.0);
            }
#pragma unroll
          for (int u = 0; u < NB; ++u)
            if (j + u < nchunk)
              {
                const int J  = __shfl_sync(0xffffffffu, mycol, j + u);
                const int sl = __shfl_sync(0xffffffffu, myslot, j + u);
                VH_PSPMV_BLOCK(v[u], J, sl)
              }
        }
    }
#undef VH_PSPMV_BLOCK
  // row end: nine partial sums per lane -> 18 components (fixed gather order: deterministic)
#pragma unroll
  for (int k = 0; k < 3; ++k)
    {
      s_part[wid][(3 * k + 0) * 32 + lane] = ac[k];
      s_part[wid][(3 * k + 1) * 32 + lane] = a0[k] * m0[k];
      s_part[wid][(3 * k + 2) * 32 + lane] = a1[k] * m1[k];
    }
  if (lane == 0)
    s_part[wid][VH_PSPMV_NPART * 32] = 0.0; // padding target of the gather lists
  __syncwarp();
  if (lane < 18)
    {
      double sum = geo;
#pragma unroll
      for (int i = 0; i < 26; ++i)
        sum += s_part[wid][__ldg(gather_tab + i * 18 + lane)];
      const uint32_t mI = dirmask[I];
      if ((mI >> lane) & 1u) // constrained row: only the diagonal entry sum_cells |a_ii|
        sum = cdiag[(size_t)I * 18 + lane] * xo[18 * (size_t)I + lane];
      y[(size_t)I * 18 + lane] = sum;
    }
}

// x~ = x with the homogeneous-Dirichlet DoFs zeroed (their matrix columns are empty)
__global__ void k_mask_dirichlet(int64_t n, const uint32_t *__restrict__ dirmask, const double *__restrict__ x, double *__restrict__ xm)
{
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n)
    xm[i] = ((dirmask[i / 18] >> (i % 18)) & 1u) ? 0.0 : x[i];
}

// ------------------------------------------------------------------------------------------------
// block-Jacobi: invert the 18x18 diagonal blocks (Gauss-Jordan, partial pivoting), one warp per block
// ------------------------------------------------------------------------------------------------
#include "vh_block_invert.cuh"

// y = blockdiag(minv) x : same streaming scheme as the SpMV with exactly one block per row
__global__ void __launch_bounds__(256)
  k_block_apply(int n_rows, const double *__restrict__ minv, const double *__restrict__ x, double *__restrict__ y)
{
  __shared__ double s_part[8][6 * 32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int row  = blockIdx.x * 8 + wid;
  if (row >= n_rows)
    return;
  const double2 *B  = reinterpret_cast<const double2 *>(minv + (size_t)row * VH_BLK);
  const double  *xj = x + 18 * (size_t)row;
#pragma unroll
  for (int r = 0; r < 6; ++r)
    {
      const int k = lane + 32 * r;
      double    a = 0.0;
      if (k < 162)
        {
          const double2 v  = __ldg(B + k);
          const double2 xv = *reinterpret_cast<const double2 *>(xj + 2 * (k % 9));
          a                = fma(v.x, xv.x, v.y * xv.y);
        }
      s_part[wid][k] = a;
    }
  __syncwarp();
  if (lane < 18)
    {
      double s = 0.0;
#pragma unroll
      for (int k = 0; k < 9; ++k)
        s += s_part[wid][9 * lane + k];
      y[(size_t)row * 18 + lane] = s;
    }
}

// v = x / sqrt(*nsq) (written out: it is the next Krylov basis vector) and y = blockdiag(minv) v, in one pass
__global__ void __launch_bounds__(256)
  k_block_apply_scaled(int n_rows, const double *__restrict__ minv, const double *__restrict__ x, const double *__restrict__ nsq,
                       double *__restrict__ v_out, double *__restrict__ y, VhPush H, unsigned long long push_seq)
{
  // H.n_peers > 0: fused ghost push.  The 18 values of an interface node are also stored into the ghost slot of every
  // neighbour's vector over NVLink; the last block to finish (ticket) posts this rank's completion flag to the neighbours.
  __shared__ double s_part[8][6 * 32];
  __shared__ int    s_pushed;
  if (H.n_peers > 0)
    {
      if (threadIdx.x == 0)
        s_pushed = 0;
      __syncthreads();
    }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int row  = blockIdx.x * 8 + wid;
  if (row < n_rows)
  {
  const double   nrm = sqrt(*nsq);
  const double   inv = nrm != 0.0 ? 1.0 / nrm : 0.0;
  const double2 *B   = reinterpret_cast<const double2 *>(minv + (size_t)row * VH_BLK);
  const double  *xj  = x + 18 * (size_t)row;
  if (lane < 9)
    {
      double2 xv = *reinterpret_cast<const double2 *>(xj + 2 * lane);
      xv.x *= inv;
      xv.y *= inv;
      *reinterpret_cast<double2 *>(v_out + 18 * (size_t)row + 2 * lane) = xv;
    }
#pragma unroll
  for (int r = 0; r < 6; ++r)
    {
      const int k = lane + 32 * r;
      double    a = 0.0;
      if (k < 162)
        {
          const double2 m  = __ldg(B + k);
          const double2 xv = *reinterpret_cast<const double2 *>(xj + 2 * (k % 9));
          a                = fma(m.x, xv.x * inv, m.y * (xv.y * inv));
        }
      s_part[wid][k] = a;
    }
  __syncwarp();
  if (lane < 18)
    {
      double s = 0.0;
#pragma unroll
      for (int k = 0; k < 9; ++k)
        s += s_part[wid][9 * lane + k];
      y[(size_t)row * 18 + lane] = s;
      if (H.n_peers > 0)
        for (int e = H.push_ptr[row]; e < H.push_ptr[row + 1]; ++e)
          {
            H.zpeer[H.push_peer[e]][18 * (size_t)H.push_dst[e] + lane] = s;
            s_pushed = 1;
          }
    }
  }
  if (H.n_peers > 0)
    {
      __syncthreads();
      if (threadIdx.x < 32)
        {
          bool last = false;
          if (threadIdx.x == 0)
            {
              if (s_pushed)
                __threadfence_system(); // the remote stores of this block before the ticket
              else
                __threadfence();
              last = atomicAdd(H.ticket, 1u) == gridDim.x - 1;
            }
          last = __shfl_sync(0xffffffffu, (int)last, 0);
          __syncwarp(); // memory ordering between lane 0's ticket and the other lanes' flag stores
          if (last)
            {
              __threadfence_system();
              if (lane < H.n_peers)
                vh_p2p_post(H.flag_dst[lane], push_seq, 0.0);
              if (lane == 0)
                *H.ticket = 0u;
            }
        }
    }
}

// waits until every neighbour has posted its completion flag for push number `seq`
__global__ void k_halo_wait(VhPush H, unsigned long long seq)
{
  if ((int)threadIdx.x < H.n_peers)
    (void)vh_p2p_wait(H.flag_src[threadIdx.x], seq, H.err);
}

// ------------------------------------------------------------------------------------------------
// Whole modified Gram-Schmidt step of GMRES in ONE cooperative kernel (single-rank contexts):
//   h_0 = w.v_0;  for i = 1..j: w -= h_{i-1} v_{i-1}, h_i = w.v_i;  w -= h_j v_j, h_{j+1} = w.w
// (deal.II's add_and_dot chain).  j+2 grid-wide barriers replace j+2 kernel launches; each thread keeps
// its slice of w in registers, so w is read and written exactly once.  Reductions are deterministic: per-block partials
// are summed in the same fixed order by every block.
// ------------------------------------------------------------------------------------------------
#define VH_MGS_THREADS 256
// EPT = elements of w per thread held in registers (as double2): 8 covers 1.2 M DoFs per GPU, 32 up to 4.8 M.
// EPT = 0 is the STREAMING variant for vectors of any length: w stays in global memory and is updated in place by the
// thread that owns the element (grid-stride slices), so there is still no kernel boundary and no host round trip between
// the j+2 dependent reductions; with 4.8 M DoFs per rank (C5 on 8 GPUs) w and the last basis vector stay L2-resident and
// only the next basis vector streams from HBM.
template <int EPT>
__global__ void __launch_bounds__(VH_MGS_THREADS, EPT == 8 ? 3 : (EPT == 0 ? 2 : 1))
  k_mgs_fused(int64_t n, double *__restrict__ w, const double *__restrict__ V, int64_t ld, int j, double *__restrict__ hcol,
              double *__restrict__ partials, unsigned int *__restrict__ tickets, VhP2P P, unsigned long long seq0,
              double *__restrict__ nrm2_out, volatile double *host_out, unsigned long long host_seq)
{
  // Launched cooperatively (all blocks co-resident): one GPU uses cg grid barriers, several GPUs wait through the mailboxes.
  __shared__ double s_w[VH_MGS_THREADS / 32];
  __shared__ double s_tot;
  __shared__ bool   s_last;
  const int         lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int64_t     base = ((int64_t)blockIdx.x * VH_MGS_THREADS + threadIdx.x) * 2;
  const int64_t     stride = (int64_t)gridDim.x * VH_MGS_THREADS * 2;
  double2           wr[EPT > 0 ? EPT / 2 : 1];
#pragma unroll
  for (int k = 0; k < EPT / 2; ++k)
    {
      const int64_t i = base + k * stride;
      wr[k]           = make_double2(0.0, 0.0);
      if (i + 1 < n)
        wr[k] = *reinterpret_cast<const double2 *>(w + i);
      else if (i < n)
        wr[k].x = w[i];
    }
  // Every basis vector is read ONCE: V[step] is the dot partner of step `step` and the subtracted vector of step + 1, and
  // (EPT = 8) the next one is requested before this step's reduction so that its latency hides behind the barrier.
  constexpr bool PF = EPT == 8;
#define VH_MGS_LOAD(vptr, dst)                                              \
  _Pragma("unroll") for (int k = 0; k < EPT / 2; ++k)                       \
  {                                                                         \
    const int64_t i = base + k * stride;                                    \
    (dst)[k]        = make_double2(0.0, 0.0);                               \
    if (i + 1 < n)                                                          \
      (dst)[k] = *reinterpret_cast<const double2 *>((vptr) + i);            \
    else if (i < n)                                                         \
      (dst)[k].x = (vptr)[i];                                               \
  }
  double2 cu[PF ? EPT / 2 : 1], nu[PF ? EPT / 2 : 1];
  if constexpr (PF)
    {
      VH_MGS_LOAD(V, nu)
    }
  double hprev = 0.0;
  for (int step = 0; step <= j + 1; ++step)
    {
      double s = 0.0;
      if constexpr (EPT == 0)
        { // streaming: this thread's slice of w is updated in global memory (nobody else touches these elements)
          const double *vp = step > 0 ? V + (size_t)(step - 1) * ld : nullptr; // subtract hprev * v_{step-1}
          const double *vu = step <= j ? V + (size_t)step * ld : nullptr;      // dot with v_step (or with w itself)
          double        s1 = 0.0;
          // main part: four 16-byte pieces of every stream requested before the first is used (12 loads in flight per thread:
          // with only 2 co-resident blocks per SM the memory-level parallelism has to come from the thread itself)
          int64_t i = base;
          for (; i + 3 * stride + 1 < n; i += 4 * stride)
            {
              double2 wv[4], pv[4], uv[4];
#pragma unroll
              for (int u = 0; u < 4; ++u)
                wv[u] = *reinterpret_cast<const double2 *>(w + i + u * stride);
              if (vp)
#pragma unroll
                for (int u = 0; u < 4; ++u)
                  pv[u] = __ldcs(reinterpret_cast<const double2 *>(vp + i + u * stride));
              if (vu)
#pragma unroll
                for (int u = 0; u < 4; ++u)
                  uv[u] = *reinterpret_cast<const double2 *>(vu + i + u * stride);
#pragma unroll
              for (int u = 0; u < 4; ++u)
                {
                  if (vp)
                    {
                      wv[u].x = fma(-hprev, pv[u].x, wv[u].x);
                      wv[u].y = fma(-hprev, pv[u].y, wv[u].y);
                      *reinterpret_cast<double2 *>(w + i + u * stride) = wv[u];
                    }
                  if (!vu)
                    uv[u] = wv[u];
                  s  = fma(wv[u].x, uv[u].x, s);
                  s1 = fma(wv[u].y, uv[u].y, s1);
                }
            }
          for (; i < n; i += stride)
            {
              const bool two = i + 1 < n;
              double2    wv  = two ? *reinterpret_cast<const double2 *>(w + i) : make_double2(w[i], 0.0);
              if (vp)
                {
                  const double2 pv = two ? __ldcs(reinterpret_cast<const double2 *>(vp + i)) : make_double2(vp[i], 0.0);
                  wv.x             = fma(-hprev, pv.x, wv.x);
                  wv.y             = fma(-hprev, pv.y, wv.y);
                  if (two)
                    *reinterpret_cast<double2 *>(w + i) = wv;
                  else
                    w[i] = wv.x;
                }
              double2 uv = wv;
              if (vu)
                uv = two ? *reinterpret_cast<const double2 *>(vu + i) : make_double2(vu[i], 0.0);
              s  = fma(wv.x, uv.x, s);
              s1 = fma(wv.y, uv.y, s1);
            }
          s += s1;
        }
      else if constexpr (PF)
        {
          // w -= hprev * V[step-1]  (cu still holds V[step-1] from the previous step)
          if (step > 0)
#pragma unroll
            for (int k = 0; k < EPT / 2; ++k)
              {
                wr[k].x = fma(-hprev, cu[k].x, wr[k].x);
                wr[k].y = fma(-hprev, cu[k].y, wr[k].y);
              }
          if (step <= j)
            {
#pragma unroll
              for (int k = 0; k < EPT / 2; ++k)
                cu[k] = nu[k];
              if (step + 1 <= j)
                {
                  const double *vn = V + (size_t)(step + 1) * ld;
                  VH_MGS_LOAD(vn, nu)
                }
#pragma unroll
              for (int k = 0; k < EPT / 2; ++k)
                s = fma(wr[k].x, cu[k].x, fma(wr[k].y, cu[k].y, s));
            }
          else
#pragma unroll
            for (int k = 0; k < EPT / 2; ++k)
              s = fma(wr[k].x, wr[k].x, fma(wr[k].y, wr[k].y, s));
        }
      else
        { // long vectors: w fills the registers, the basis vectors are read where they are used
          const double *vp = step > 0 ? V + (size_t)(step - 1) * ld : nullptr; // subtract hprev * v_{step-1}
          const double *vu = step <= j ? V + (size_t)step * ld : nullptr;      // dot with v_step (or with w itself)
#pragma unroll
          for (int k = 0; k < EPT / 2; ++k)
            {
              const int64_t i = base + k * stride;
              if (i >= n)
                continue;
              const bool two = i + 1 < n;
              if (vp)
                {
                  const double2 pv = two ? *reinterpret_cast<const double2 *>(vp + i) : make_double2(vp[i], 0.0);
                  wr[k].x          = fma(-hprev, pv.x, wr[k].x);
                  wr[k].y          = fma(-hprev, pv.y, wr[k].y);
                }
              double2 uv = wr[k];
              if (vu)
                uv = two ? *reinterpret_cast<const double2 *>(vu + i) : make_double2(vu[i], 0.0);
              s = fma(wr[k].x, uv.x, fma(wr[k].y, uv.y, s));
            }
        }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1)
        s += __shfl_xor_sync(0xffffffffu, s, o);
      if (lane == 0)
        s_w[wid] = s;
      __syncthreads();
      if (threadIdx.x == 0)
        {
          double b = 0.0;
          for (int k = 0; k < VH_MGS_THREADS / 32; ++k)
            b += s_w[k];
          partials[(size_t)step * gridDim.x + blockIdx.x] = b;
          if (P.n > 1)
            {
              __threadfence();
              s_last = atomicAdd(tickets + step, 1u) == gridDim.x - 1;
            }
        }
      if (P.n == 1)
        { // one GPU: a grid-wide barrier, then every block sums all partials of this step in the same order
          cg::this_grid().sync();
          if (wid == 0)
            {
              double t = 0.0;
              for (int b = lane; b < (int)gridDim.x; b += 32)
                t += partials[(size_t)step * gridDim.x + b];
#pragma unroll
              for (int o = 16; o > 0; o >>= 1)
                t += __shfl_xor_sync(0xffffffffu, t, o);
              if (lane == 0)
                s_tot = t;
            }
          __syncthreads();
          hprev = s_tot;
          if (blockIdx.x == 0 && threadIdx.x == 0)
            {
              hcol[step]     = hprev;
              host_out[step] = hprev; // mapped pinned host memory: the host polls instead of copying (see below)
            }
          __syncthreads();
          continue;
        }
      // several GPUs: no grid barrier at all - the blocks of all ranks meet in the peer-memory mailboxes
      __syncthreads();
      const unsigned long long seq  = seq0 + step;
      const int                slot = (int)(seq % VH_P2P_SLOTS);
      if (wid == 0)
        {
          if (s_last)
            { // the last block of this rank adds the partials in index order and posts the rank's total into the mailbox of
              // every rank (its own included): one NVLink store per peer, released by the sequence number
              __threadfence();
              double t = 0.0;
              for (int b = lane; b < (int)gridDim.x; b += 32)
                t += __ldcg(partials + (size_t)step * gridDim.x + b);
#pragma unroll
              for (int o = 16; o > 0; o >>= 1)
                t += __shfl_xor_sync(0xffffffffu, t, o);
              if (lane < P.n)
                vh_p2p_post(P.peer[lane] + slot * P.n + P.me, seq, t);
              if (lane == 0)
                tickets[step] = 0u; // ready for the next launch (nobody touches it again in this one)
            }
          // every block polls its own rank's mailbox and adds the ranks' totals in rank order (bit-identical everywhere)
          double got = 0.0;
          if (lane < P.n)
            got = vh_p2p_wait(P.peer[P.me] + slot * P.n + lane, seq, P.err);
          double sum = 0.0;
          for (int r = 0; r < P.n; ++r)
            sum += __shfl_sync(0xffffffffu, got, r);
          if (lane == 0)
            s_tot = sum;
        }
      __syncthreads();
      hprev = s_tot;
      if (blockIdx.x == 0 && threadIdx.x == 0)
        {
          hcol[step]     = hprev;
          host_out[step] = hprev;
        }
    }
  if (blockIdx.x == 0 && threadIdx.x == 0)
    { // hand the column to the host without a copy operation in the stream: values first, then the sequence flag
      *nrm2_out = hprev; // a^2 for the scaling kernel of the next inner step
      __threadfence_system();
      *reinterpret_cast<volatile unsigned long long *>(host_out + VH_SCAL_COUNT - 1) = host_seq;
    }
#pragma unroll
  for (int k = 0; k < EPT / 2; ++k)
    {
      const int64_t i = base + k * stride;
      if (i + 1 < n)
        *reinterpret_cast<double2 *>(w + i) = wr[k];
      else if (i < n)
        w[i] = wr[k].x;
    }
}

// ------------------------------------------------------------------------------------------------
// fused vector kernels with deterministic device-side reductions
// ------------------------------------------------------------------------------------------------
#define VH_RED_THREADS 256

// Final stage: the last block to finish sums the per-block partials in index order.
__device__ __forceinline__ void red_finish(double block_val, double *partials, unsigned int *ticket, double *out)
{
  __shared__ double s_w[VH_RED_THREADS / 32];
  __shared__ bool   s_last;
  const int         lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  double            v = warp_sum(block_val);
  if (lane == 0)
    s_w[wid] = v;
  __syncthreads();
  if (threadIdx.x == 0)
    {
      double s = 0.0;
      for (int i = 0; i < VH_RED_THREADS / 32; ++i)
        s += s_w[i];
      partials[blockIdx.x] = s;
      __threadfence();
      const unsigned int done = atomicAdd(ticket, 1u);
      s_last                  = (done == gridDim.x - 1);
    }
  __syncthreads();
  if (s_last)
    {
      __threadfence();
      double s = 0.0;
      for (int i = threadIdx.x; i < (int)gridDim.x; i += VH_RED_THREADS)
        s += ((volatile double *)partials)[i];
      s = warp_sum(s);
      __syncthreads();
      if (lane == 0)
        s_w[wid] = s;
      __syncthreads();
      if (threadIdx.x == 0)
        {
          double tot = 0.0;
          for (int i = 0; i < VH_RED_THREADS / 32; ++i)
            tot += s_w[i];
          *out    = tot;
          *ticket = 0u;
        }
    }
}

__global__ void __launch_bounds__(VH_RED_THREADS)
  k_dot(int64_t n, const double *__restrict__ a, const double *__restrict__ b, double *partials, unsigned int *ticket, double *out)
{
  double        s      = 0.0;
  const int64_t stride = (int64_t)gridDim.x * VH_RED_THREADS * 2;
  for (int64_t i = ((int64_t)blockIdx.x * VH_RED_THREADS + threadIdx.x) * 2; i < n; i += stride)
    {
      if (i + 1 < n)
        {
          const double2 av = *reinterpret_cast<const double2 *>(a + i), bv = *reinterpret_cast<const double2 *>(b + i);
          s                = fma(av.x, bv.x, fma(av.y, bv.y, s));
        }
      else
        s = fma(a[i], b[i], s);
    }
  red_finish(s, partials, ticket, out);
}

// w += (-*coef) v ; out = sum w*u   (u may alias w: the norm^2 of the updated w)
__global__ void __launch_bounds__(VH_RED_THREADS)
  k_add_and_dot(int64_t n, double *w, const double *__restrict__ coef, const double *__restrict__ v, const double *u,
                double *partials, unsigned int *ticket, double *out)
{
  const double  a      = -(*coef);
  const bool    self   = (u == w);
  double        s      = 0.0;
  const int64_t stride = (int64_t)gridDim.x * VH_RED_THREADS * 2;
  for (int64_t i = ((int64_t)blockIdx.x * VH_RED_THREADS + threadIdx.x) * 2; i < n; i += stride)
    {
      if (i + 1 < n)
        {
          double2       wv = *reinterpret_cast<double2 *>(w + i);
          const double2 vv = *reinterpret_cast<const double2 *>(v + i);
          wv.x             = fma(a, vv.x, wv.x);
          wv.y             = fma(a, vv.y, wv.y);
          *reinterpret_cast<double2 *>(w + i) = wv;
          const double2 uv = self ? wv : *reinterpret_cast<const double2 *>(u + i);
          s                = fma(wv.x, uv.x, fma(wv.y, uv.y, s));
        }
      else
        {
          const double wv = fma(a, v[i], w[i]);
          w[i]            = wv;
          s               = fma(wv, self ? wv : u[i], s);
        }
    }
  red_finish(s, partials, ticket, out);
}

__global__ void k_scale_to(int64_t n, double *__restrict__ dst, const double *__restrict__ src, const double *__restrict__ nsq)
{
  const double  nrm = sqrt(*nsq);
  const double  inv = nrm != 0.0 ? 1.0 / nrm : 0.0;
  const int64_t i   = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n)
    dst[i] = src[i] * inv;
}

// y += sum_{j<k} c_j V_j   (GMRES solution update)
__global__ void k_axpy_multi(int64_t n, double *__restrict__ y, const double *__restrict__ coefs, int k, const double *__restrict__ V,
                             int64_t ld)
{
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n)
    return;
  double s = y[i];
  for (int j = 0; j < k; ++j)
    s = fma(coefs[j], V[(size_t)j * ld + i], s);
  y[i] = s;
}

__global__ void k_axpby(int64_t n, double *__restrict__ z, double a, const double *__restrict__ x, double b, const double *__restrict__ y)
{
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n)
    z[i] = a * x[i] + (b != 0.0 ? b * y[i] : 0.0);
}

// AffineConstraints::distribute on the owned constrained DoFs: x[dof] = sum w x[master]
__global__ void k_distribute(int n_lines, int64_t n_owned_dofs, const int32_t *__restrict__ dof, const int32_t *__restrict__ ptr,
                             const int32_t *__restrict__ master, const double *__restrict__ weight, double *x)
{
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= n_lines)
    return;
  const int d = dof[l];
  if (d >= n_owned_dofs)
    return;
  double s = 0.0;
  for (int p = ptr[l]; p < ptr[l + 1]; ++p)
    s = fma(weight[p], x[master[p]], s);
  x[d] = s;
}

__global__ void __launch_bounds__(VH_RED_THREADS)
  k_masked_sum(int64_t n, const double *__restrict__ a, const uint8_t *__restrict__ mask, double *partials, unsigned int *ticket,
               double *out)
{
  double s = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * VH_RED_THREADS + threadIdx.x; i < n; i += (int64_t)gridDim.x * VH_RED_THREADS)
    if (!mask || mask[i])
      s += a[i];
  red_finish(s, partials, ticket, out);
}

inline unsigned red_grid(int64_t n)
{
  int64_t g = (n + (int64_t)VH_RED_THREADS * 8 - 1) / ((int64_t)VH_RED_THREADS * 8);
  if (g < 1)
    g = 1;
  if (g > 148 * 8)
    g = 148 * 8;
  return (unsigned)g;
}
} // namespace

int vhk_upload_linalg_constants(vh_ctx *ctx)
{
  uint8_t sc[VH_SYMP], sd[VH_SYMP];
  vh_sym_tables(sc, sd);
  // Lane table and gather lists of k_spmv_sym18.  double2 #p (entries e = 2p, 2p+1) lives in lane p%32, slot k = p/32 and
  // produces three partials: [3k+0] -> y_c, [3k+1] -> y_d (mirror of the first entry), [3k+2] -> y_{d+1} (mirror of the
  // second).  Component comp collects at most 9 + 17 = 26 of them; unused list positions point at a zero slot.
  std::vector<uint32_t> lt(96, 0u);
  std::vector<uint16_t> g(26 * 18, (uint16_t)(9 * 32));
  std::vector<int>      cnt(18, 0);
  for (int p = 0; p < 96; ++p)
    {
      const int pp = p < VH_SYMP / 2 ? p : VH_SYMP / 2 - 1; // lanes without a third double2 mirror the last one (their v is 0)
      const int c = sc[2 * pp], d = 2 * (sd[2 * pp + 1] >> 1);
      const int mir0 = (d != c && sd[2 * pp] >= sc[2 * pp]) ? 1 : 0, mir1 = (d + 1 != c) ? 1 : 0;
      lt[(p / 32) * 32 + p % 32] = (uint32_t)c | ((uint32_t)d << 8) | ((uint32_t)mir0 << 16) | ((uint32_t)mir1 << 17);
      if (p >= VH_SYMP / 2)
        continue;
      const int lane = p % 32, k = p / 32;
      auto      add = [&](int comp, int slot) {
        if (cnt[comp] < 26)
          g[(size_t)cnt[comp] * 18 + comp] = (uint16_t)(slot * 32 + lane);
        cnt[comp]++;
      };
      add(c, 3 * k + 0);
      if (mir0)
        add(d, 3 * k + 1);
      if (mir1)
        add(d + 1, 3 * k + 2);
    }
  for (int comp = 0; comp < 18; ++comp)
    if (cnt[comp] > 26)
      return vh_fail(ctx, VH_ERR_ARG, "internal: SpMV gather list overflow");
  VH_TRY(vh_dev_upload(ctx, &ctx->spmv_lane_tab, lt.data(), lt.size()));
  VH_TRY(vh_dev_upload(ctx, &ctx->spmv_gather_tab, g.data(), g.size()));
  return VH_OK;
}

int vhk_spmv(vh_ctx *ctx, const double *x_local, double *y_owned, bool x_is_masked)
{
  if (ctx->n_owned == 0)
    return VH_OK;
  if (ctx->packed)
    {
      const unsigned grid = (ctx->n_fast + VH_PSPMV_WARPS - 1) / VH_PSPMV_WARPS;
      const double *xg = x_local;
      if (!x_is_masked)
        { // arbitrary input: the blocks are applied to a copy whose Dirichlet DoFs are zeroed
          if (!ctx->xmask)
            VH_TRY(vh_dev_alloc(ctx, &ctx->xmask, (size_t)ctx->NL));
          k_mask_dirichlet<<<(unsigned)((ctx->NL + 255) / 256), 256, 0, ctx->stream>>>(ctx->NL, ctx->dirmask, x_local, ctx->xmask);
          VH_LAUNCH_CHECK();
          xg = ctx->xmask;
        }
      if (ctx->spmv_mf)
        { // matrix-free: K_cell z_cell from the H_q tables, gathered per lattice row; constrained rows below as usual
          VH_TRY(vhk_apply_fast(ctx, xg, x_local, y_owned));
          if (ctx->n_slow_rows > 0)
            {
              const unsigned gs = (ctx->n_slow_rows + VH_SPMV_WARPS - 1) / VH_SPMV_WARPS;
              k_spmv_bsr18<<<gs, VH_SPMV_WARPS * 32, 0, ctx->stream>>>(ctx->n_slow_rows, ctx->slow_rows, ctx->row_ptr, ctx->col, ctx->vals,
                                                                      x_local, y_owned);
              VH_LAUNCH_CHECK();
            }
          return VH_OK;
        }
      VH_TRY(vhk_ensure_rows(ctx)); // lattice rows that were left unassembled for the matrix-free mode (VH_MF_LAZY_ROWS=1)
#define VH_LAUNCH_PSPMV(NB, MINB)                                                                                                  \
  k_spmv_sym18<NB, MINB><<<grid, VH_PSPMV_WARPS * 32, 0, ctx->stream>>>(ctx->n_fast, ctx->slot_stride, ctx->n_slots * 10, ctx->spmv_order, ctx->fast_rows, ctx->fast_posslot, \
                                                                       ctx->fast_class, ctx->class_M, ctx->row_ptr, ctx->col,      \
                                                                       ctx->dirmask, ctx->pvals, ctx->cdiag, ctx->spmv_lane_tab,   \
                                                                       ctx->spmv_gather_tab, xg, x_local, y_owned)
      VH_LAUNCH_PSPMV(4, 2); // four blocks in flight per warp, two CTAs per SM (2/3/6 in flight measured 0.26-0.30 vs 0.256 ms at C2)
#undef VH_LAUNCH_PSPMV
      VH_LAUNCH_CHECK();
      if (ctx->n_slow_rows > 0)
        {
          const unsigned gs = (ctx->n_slow_rows + VH_SPMV_WARPS - 1) / VH_SPMV_WARPS;
          k_spmv_bsr18<<<gs, VH_SPMV_WARPS * 32, 0, ctx->stream>>>(ctx->n_slow_rows, ctx->slow_rows, ctx->row_ptr, ctx->col, ctx->vals,
                                                                  x_local, y_owned);
          VH_LAUNCH_CHECK();
        }
      return VH_OK;
    }
  const unsigned grid = (ctx->n_owned + VH_SPMV_WARPS - 1) / VH_SPMV_WARPS;
  k_spmv_bsr18<<<grid, VH_SPMV_WARPS * 32, 0, ctx->stream>>>(ctx->n_owned, nullptr, ctx->row_ptr, ctx->col, ctx->vals, x_local,
                                                            y_owned);
  VH_LAUNCH_CHECK();
  return VH_OK;
}

int vhk_block_jacobi_setup(vh_ctx *ctx)
{
  if (ctx->n_owned == 0)
    return VH_OK;
  int *d_sing = reinterpret_cast<int *>(ctx->scal + VH_SCAL_MISC);
  VH_CUDA(cudaMemsetAsync(d_sing, 0, sizeof(int), ctx->stream));
  k_block_invert<<<(ctx->n_owned + VH_INV_WARPS - 1) / VH_INV_WARPS, VH_INV_WARPS * 32, 0, ctx->stream>>>(ctx->n_owned, ctx->diag_pos, ctx->vals, ctx->minv, d_sing,
                                                                  ctx->packed ? ctx->pvals : nullptr, ctx->n_slots * 10, ctx->diag_slot, ctx->fast_index, ctx->fast_class,
                                                                  ctx->class_M, ctx->dirmask, ctx->cdiag,
                                                                  ctx->rows_stale ? ctx->dpack : nullptr, ctx->packed ? 1 : 0);
  VH_LAUNCH_CHECK();
  int h_sing = 0;
  VH_CUDA(cudaMemcpyAsync(&h_sing, d_sing, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  VH_CUDA(cudaStreamSynchronize(ctx->stream));
  if (h_sing)
    return vh_fail(ctx, VH_ERR_ARG, std::to_string(h_sing) + " singular 18x18 diagonal blocks in block-Jacobi setup");
  return VH_OK;
}

int vhk_block_jacobi_apply(vh_ctx *ctx, const double *x_owned, double *y_owned)
{
  if (ctx->n_owned == 0)
    return VH_OK;
  k_block_apply<<<(ctx->n_owned + 7) / 8, 256, 0, ctx->stream>>>(ctx->n_owned, ctx->minv, x_owned, y_owned);
  VH_LAUNCH_CHECK();
  return VH_OK;
}

int vhk_halo_wait(vh_ctx *ctx)
{
  k_halo_wait<<<1, 32, 0, ctx->stream>>>(ctx->zpush_dev, ctx->zpush_seq);
  VH_LAUNCH_CHECK();
  return VH_OK;
}

int vhk_block_jacobi_apply_scaled(vh_ctx *ctx, const double *x_owned, const double *nsq_dev, double *v_out, double *y_owned, bool push)
{
  if (ctx->n_owned == 0)
    return VH_OK;
  VhPush H = ctx->zpush_dev;
  if (!push)
    H.n_peers = 0;
  else
    ++ctx->zpush_seq;
  k_block_apply_scaled<<<(ctx->n_owned + 7) / 8, 256, 0, ctx->stream>>>(ctx->n_owned, ctx->minv, x_owned, nsq_dev, v_out, y_owned, H,
                                                                        ctx->zpush_seq);
  VH_LAUNCH_CHECK();
  return VH_OK;
}

// Returns VH_OK and sets *used = true when the fused cooperative kernel ran; *used = false means "not applicable here"
// (multi-rank context, vector too long for the register-resident slice, or no cooperative launch): use the kernel chain.
// Largest vector the fused kernel can hold in registers with EPT elements per thread (0: no cooperative launch).
static int64_t mgs_capacity(vh_ctx *ctx, int ept)
{
  int coop = 0, sms = 0, per_sm = 0;
  cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, ctx->device);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device);
  if (ept == 8)
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_mgs_fused<8>, VH_MGS_THREADS, 0);
  else if (ept == 32)
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_mgs_fused<32>, VH_MGS_THREADS, 0);
  else
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_mgs_fused<0>, VH_MGS_THREADS, 0);
  if (!coop)
    return 0;
  if (ept == 0) // streaming variant: the number of co-resident blocks (any vector length)
    return (int64_t)sms * (per_sm < 2 ? per_sm : 2);
  return (int64_t)sms * (per_sm < 4 ? per_sm : 4) * VH_MGS_THREADS * ept;
}
// 8 or 32 = elements per thread this rank needs for its owned vector, 64 = streaming variant (any length), 1000 = cannot fuse
// (no cooperative launch).  Ordered so that the maximum over the ranks is the mode every rank can run.
int vhk_mgs_mode_local(vh_ctx *ctx)
{
  if (const char *e = getenv("VH_MGS_MODE")) // test hook: force a variant (8, 32, 64 = streaming, 1000 = kernel chain)
    {
      const int m = atoi(e);
      if (m == 8 || m == 32 || m == 64 || m == 1000)
        return (m == 64 && mgs_capacity(ctx, 0) == 0) ? 1000 : m;
    }
  if (ctx->NO == 0)
    return 8;
  for (int ept : {8, 32})
    if (ctx->NO <= mgs_capacity(ctx, ept) && (int64_t)(VH_MAX_RESTART + 2) * ((ctx->NO + VH_MGS_THREADS * ept - 1) / (VH_MGS_THREADS * ept)) <= (int64_t)VH_MAX_RED_BLOCKS * 32)
      return ept;
  return mgs_capacity(ctx, 0) > 0 ? 64 : 1000;
}

int vhk_mgs_fused(vh_ctx *ctx, double *w, const double *V, int64_t ld, int j, double *hcol_dev, bool *used)
{
  *used = false;
  if (ctx->mgs_mode < 0) // single rank: decided here; multi-rank: agreed over all ranks in vh_comm_init
    ctx->mgs_mode = ctx->n_ranks == 1 ? vhk_mgs_mode_local(ctx) : 1000;
  if (ctx->mgs_mode > 64 || (ctx->n_ranks > 1 && !ctx->p2p) || (ctx->n_ranks == 1 && ctx->NO == 0))
    return VH_OK;
  const int     ept       = ctx->mgs_mode == 64 ? 0 : ctx->mgs_mode;
  const int64_t per_block = (int64_t)VH_MGS_THREADS * (ept ? ept : 2);
  int64_t       grid      = std::max<int64_t>(1, (ctx->NO + per_block - 1) / per_block);
  if (ept == 0)
    grid = std::min<int64_t>(grid, std::max<int64_t>(1, mgs_capacity(ctx, 0)));
  int64_t       n         = ctx->NO;
  VhP2P         P         = ctx->p2p_dev; // single rank: the mailbox is this context's own buffer (set up in vh_create)
  unsigned long long seq0 = ctx->p2p_seq + 1;
  double            *nrm2_out = ctx->scal + VH_SCAL_NRM2;
  double            *host_out = ctx->h_mgs;
  unsigned long long host_seq = ++ctx->h_mgs_seq;
  void *args[] = {&n, &w, (void *)&V, &ld, &j, &hcol_dev, &ctx->partials, &ctx->mgs_tickets, &P, &seq0, &nrm2_out, &host_out, &host_seq};
  cudaError_t e = cudaLaunchCooperativeKernel(ept == 8 ? (void *)k_mgs_fused<8> : (ept == 32 ? (void *)k_mgs_fused<32> : (void *)k_mgs_fused<0>), dim3((unsigned)grid),
                                              dim3(VH_MGS_THREADS), args, 0, ctx->stream);
  if (e != cudaSuccess)
    return vh_fail(ctx, VH_ERR_CUDA, std::string("cooperative launch: ") + cudaGetErrorString(e));
  if (ctx->n_ranks > 1)
    ctx->p2p_seq += j + 2;
  ctx->n_launches++;
  *used = true;
  return VH_OK;
}

int vhk_dot(vh_ctx *ctx, const double *a, const double *b, double *out)
{
  k_dot<<<red_grid(ctx->NO), VH_RED_THREADS, 0, ctx->stream>>>(ctx->NO, a, b, ctx->partials, ctx->ticket, out);
  VH_LAUNCH_CHECK();
  return vhk_allreduce_sum(ctx, out, 1);
}

int vhk_add_and_dot(vh_ctx *ctx, double *w, const double *coef_dev, const double *v, const double *u, double *out)
{
  k_add_and_dot<<<red_grid(ctx->NO), VH_RED_THREADS, 0, ctx->stream>>>(ctx->NO, w, coef_dev, v, u, ctx->partials, ctx->ticket, out);
  VH_LAUNCH_CHECK();
  return vhk_allreduce_sum(ctx, out, 1);
}

int vhk_scale_to(vh_ctx *ctx, double *dst, const double *src, const double *norm_sq_dev)
{
  if (ctx->NO == 0)
    return VH_OK;
  k_scale_to<<<(unsigned)((ctx->NO + 255) / 256), 256, 0, ctx->stream>>>(ctx->NO, dst, src, norm_sq_dev);
  VH_LAUNCH_CHECK();
  return VH_OK;
}

int vhk_axpy_dev(vh_ctx *ctx, double *y, const double *coefs_dev, int k, const double *V, int64_t ld)
{
  if (ctx->NO == 0 || k == 0)
    return VH_OK;
  k_axpy_multi<<<(unsigned)((ctx->NO + 255) / 256), 256, 0, ctx->stream>>>(ctx->NO, y, coefs_dev, k, V, ld);
  VH_LAUNCH_CHECK();
  return VH_OK;
}

int vhk_axpby(vh_ctx *ctx, double *z, double a, const double *x, double b, const double *y, int64_t n)
{
  if (n == 0)
    return VH_OK;
  k_axpby<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(n, z, a, x, b, y ? y : x);
  VH_LAUNCH_CHECK();
  return VH_OK;
}

int vhk_distribute(vh_ctx *ctx, int which, double *x_local)
{
  const VhConstraintsDev &C = ctx->cons[which];
  if (C.n_lines == 0)
    return VH_OK;
  k_distribute<<<(C.n_lines + 255) / 256, 256, 0, ctx->stream>>>(C.n_lines, ctx->NO, C.dof, C.ptr, C.master, C.weight, x_local);
  VH_LAUNCH_CHECK();
  return VH_OK;
}

int vhk_sum(vh_ctx *ctx, const double *a, const uint8_t *mask, int64_t n, double *out)
{
  k_masked_sum<<<red_grid(n), VH_RED_THREADS, 0, ctx->stream>>>(n, a, mask, ctx->partials, ctx->ticket, out);
  VH_LAUNCH_CHECK();
  return vhk_allreduce_sum(ctx, out, 1);
}

int vh_read_scalars(vh_ctx *ctx, const double *dev, int n, double *host)
{
  VH_CUDA(cudaMemcpyAsync(ctx->h_pinned, dev, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  VH_CUDA(cudaStreamSynchronize(ctx->stream));
  for (int i = 0; i < n; ++i)
    host[i] = ctx->h_pinned[i];
  return VH_OK;
}
