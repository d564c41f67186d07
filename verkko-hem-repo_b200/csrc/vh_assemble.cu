// Jacobian / residual assembly of the femgl Newton step on sm_100a.
//
// Replaces FemGL::assemble_system (/root/reference/femgl/src/assemble.cc:108-372) and
// FemGL::compute_residual (/root/reference/femgl/src/residual.cc:109-297), including the constrained
// scatter AffineConstraints::distribute_local_to_global (call sites assemble.cc:356-361, residual.cc:287-289).
//
// Kernels instead of the reference's (cell, q, i, j) loop over 3x3 FullMatrix products:
//   1. k_points       (vh_points_kernel.cuh) one THREAD per quadrature point: gather the cell's DoFs, interpolate A and
//                     grad A, evaluate g_q (18) and the symmetric bulk Hessian H_q (171 unique entries) in closed form
//                     (vh_pointwise.cuh), store the packed H_q table, reduce the cell rhs / cell diagonal / cell energy.
//                     The same kernel in APPLY mode is the matrix-free operator of the lattice rows.
//   2. k_diag_cells / k_diag_gather (vh_diag_kernel.cuh) diagonal 18x18 blocks of the lattice rows for block-Jacobi.
//   3. k_rows_fast_q1 / k_rows_fast_q2  ROW-OWNER assembly of the lattice rows, on demand only (export, assembled SpMV
//                     mode): one CTA per block row, every packed block accumulated in registers over the <= 8 incident
//                     cells and written exactly once: no atomics, no memset.  Gradient (K1, K2+K3) and Robin-face terms
//                     are geometry-only (class_M) and Dirichlet masks are applied by the consumers.
//   4. k_rows_slow    row-owner assembly of the constrained rows (hanging-node neighbourhoods, periodic seams,
//                     constraint masters) with deal.II's distribute_local_to_global rule resolved per entry.
#include "vh_internal.h"
#include "vh_pointwise.cuh"

#include <cstdio>
#include <cstdlib>

__constant__ double  c_W1[512];   // Q1: w_q N_a(q) N_b(q), index (a*8+b)*8+q
__constant__ double  c_T2[27];    // Q2, one direction: w_q l_a(q) l_b(q), index (t_a*3 + t_b)*3 + q  (t = 0 low, 1 mid, 2 high)
__constant__ uint8_t c_q2t[27 * 3]; // Q2: tensor index (t_x, t_y, t_z) of local node a (deal.II hierarchical order)

namespace
{
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
#define VH_CELL_H_BYTES (8 * VH_SYMP * 8) /* one cell's 8 x 180 doubles: 11520 B, a multiple of 16 */
#define VH_FAST_SMEM (VH_FAST_STAGES * VH_CELL_H_BYTES + VH_FAST_STAGES * 8 + (8 + 28) * 4)

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

// One thread per packed Hessian entry (192 threads per row, 180 active, 27 accumulator registers, 24 warps/SM).  The row is
// stored as packed symmetric blocks (180 doubles each): the accumulators go straight from registers to global memory with
// coalesced stores; geometry terms and Dirichlet masks are applied by the consumers (SpMV, block-Jacobi setup, export).
__global__ void __launch_bounds__(192, 4)
  k_rows_fast_q1(const int32_t *__restrict__ fast_rows, const int32_t *__restrict__ fast_cells,
                 const int8_t *__restrict__ fast_slot, const int32_t *__restrict__ row_ptr, const double *__restrict__ Hq,
                 double *__restrict__ vals)
{
  extern __shared__ __align__(128) unsigned char smraw[];
  double   *s_H     = reinterpret_cast<double *>(smraw);                                      // [4][8*180]
  uint64_t *s_bar   = reinterpret_cast<uint64_t *>(smraw + VH_FAST_STAGES * VH_CELL_H_BYTES); // [4]
  int      *s_cells = reinterpret_cast<int *>(s_bar + VH_FAST_STAGES);                        // [8]
  int      *s_pos   = s_cells + 8;                                                            // [27] (+1 pad)

  const int t = threadIdx.x;
  const int r = blockIdx.x;
  const int I = fast_rows[r];
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
  __syncthreads();

  // TMA producer: thread 0 streams the incident cells' pre-scaled H_q tables (11.5 KB each, contiguous) into the ring
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
  double acc[27];
#pragma unroll
  for (int s = 0; s < 27; ++s)
    acc[s] = 0.0;
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
          if (t < VH_SYMP)
            {
              // cell table layout [pair][q XOR (pair & 7)] (vh_hq8_index): the stage base is 128-byte aligned, so the
              // address of point q is the address of point 0 with bits 4..6 flipped by q
              const int      pr0 = t >> 1;
              const uint32_t Hs0 = smem_u32(s_H + (size_t)st * (8 * VH_SYMP)) + (uint32_t)((pr0 << 7) + ((pr0 & 7) << 4) + (t & 1) * 8);
#pragma unroll
              for (int q = 0; q < 8; ++q)
                {
                  double hv;
                  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(hv) : "r"(Hs0 ^ (uint32_t)(q << 4)));
#pragma unroll
                  for (int b = 0; b < 8; ++b)
                    {
                      const int    s = ((o & 1) + (b & 1)) + 3 * (((o >> 1) & 1) + ((b >> 1) & 1)) + 9 * ((o >> 2) + (b >> 2));
                      const double w = c_W1[((7 - o) * 8 + b) * 8 + q];
                      acc[s]         = fma(w, hv, acc[s]);
                    }
                }
            }
        }
    }
  // block (rp + pos) holds the 180 packed entries contiguously; thread t owns entry t of every block of the row
  const int rp = row_ptr[I];
  if (t < VH_SYMP)
    {
#pragma unroll
      for (int s = 0; s < 27; ++s)
        {
          const int pos = s_pos[s];
          if (pos >= 0)
            __stcs(vals + (size_t)(rp + pos) * VH_SYMP + t, acc[s]);
        }
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
__global__ void __launch_bounds__(192, 4)
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
// (Constraint lines that couple different components - none exist in the reference: Dirichlet, hanging-node and periodic
// lines all stay within one component - would let two threads meet in one entry; vh_create rejects such tables.)
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

} // namespace

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
  else
    {
      const int grid = (ctx->n_cells + VH_PT_WARPS - 1) / VH_PT_WARPS;
      VH_LAUNCH_POINTS_MODE(27);
    }
#undef VH_LAUNCH_POINTS
#undef VH_LAUNCH_POINTS_MODE
#undef VH_POINTS_ATTR
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
  if (!ctx->dpack)
    VH_TRY(vh_dev_alloc(ctx, &ctx->dpack, (size_t)ctx->n_fast * VH_SYMP));
  k_diag_gather<<<ctx->n_fast, 192, 0, ctx->stream>>>(ctx->n_fast, ctx->nn, ctx->fast_rows, ctx->fast_cells, ctx->fast_a, ctx->Dblk,
                                                    ctx->dpack);
  VH_LAUNCH_CHECK();
  return VH_OK;
}

int vhk_alloc_rows(vh_ctx *ctx)
{
  if (ctx->packed && !ctx->pvals)
    VH_TRY(vh_dev_alloc(ctx, &ctx->pvals, (size_t)ctx->nnzb * VH_SYMP));
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
  VH_TRY(vhk_alloc_rows(ctx));
  if (ctx->degree == 2)
    k_rows_fast_q2<<<ctx->n_fast, 192, 0, ctx->stream>>>(ctx->fast_rows, ctx->fast_cells, ctx->fast_a, ctx->fast_slot, ctx->fast_first,
                                                        ctx->row_ptr, ctx->Hq, ctx->pvals);
  else
    {
      static unsigned long long rows_attr_mask = 0;
      if (vh_first_time_on_device(rows_attr_mask, ctx->device))
        VH_CUDA(cudaFuncSetAttribute(k_rows_fast_q1, cudaFuncAttributeMaxDynamicSharedMemorySize, VH_FAST_SMEM));
      k_rows_fast_q1<<<ctx->n_fast, 192, VH_FAST_SMEM, ctx->stream>>>(ctx->fast_rows, ctx->fast_cells, ctx->fast_slot, ctx->row_ptr,
                                                                     ctx->Hq, ctx->pvals);
    }
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
  // row-owner kernel: no atomics (constraint lines stay within one component: checked in vh_create)
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
