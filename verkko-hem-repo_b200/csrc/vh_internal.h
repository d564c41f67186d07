// Internal declarations of the CUDA hot-path library (not installed; the public surface is include/vh_femgl.h).
#ifndef VH_INTERNAL_H
#define VH_INTERNAL_H

#include <cuda_runtime.h>
#include <stdint.h>

#include <memory>
#include <string>
#include <vector>

#include "vh_femgl.h"
#include "vh_pointwise.cuh"

#define VH_NCOMP 18
#define VH_BLK 324   /* 18*18 doubles per matrix block */
/* VH_SYMP = 180: packed symmetric 18x18, see vh_pointwise.cuh (171 unique entries + 9 zero dummies, even row starts) */

// ---- reference-cell tables (unit cell [0,1]^3), built on the host at vh_create and uploaded once ----
// Q1: nn = nq = 8, nqf = 4.  Q2: nn = nq = 27, nqf = 9.
struct VhTables
{
  int     degree, nn, nq;
  double *N;    // [nn][nq]           shape values at QGauss<3>(degree+1) points
  double *dN;   // [nn][nq][3]        unit-cell gradients
  double *wq;   // [nq]
  double *Gref; // [nn][nn][3][3]     sum_q wq d_x N_a d_y N_b   (gradient forms are geometry-only)
  double *Mf;   // [6][nn][nn]        unit-face mass  sum_qf wf Nf_a Nf_b  for face_no 0..5
};

// Mailbox all-reduce over peer memory.  Rank r owns mbox[VH_P2P_SLOTS][n_ranks] cells of {value, sequence}; an
// all-reduce number `seq` writes this rank's partial into cell [seq % SLOTS][me] of EVERY rank, then waits until its own
// cells [seq % SLOTS][0..n) carry `seq` and adds the values in rank order, so every rank gets the bit-identical sum.
#define VH_P2P_MAX_RANKS 8
#define VH_P2P_SLOTS 8
struct VhP2PCell // two 8-byte words {half of the value, 32-bit sequence}: each word is written and read atomically, so no
{                // fences are needed (the flag travels with the data, like NCCL's LL protocol)
  unsigned long long lo, hi;
};
struct VhP2P
{
  VhP2PCell *peer[VH_P2P_MAX_RANKS]; // peer[r] = mailbox of rank r (peer[me] is local memory)
  int        me, n;
  int       *err;
};

struct VhPush
{
  double    *zpeer[VH_P2P_MAX_RANKS];   // [peer slot] neighbour's zbuf
  VhP2PCell *flag_dst[VH_P2P_MAX_RANKS]; // [peer slot] my flag cell in the neighbour's mailbox
  VhP2PCell *flag_src[VH_P2P_MAX_RANKS]; // [peer slot] the neighbour's flag cell in my mailbox
  const int32_t *push_ptr, *push_dst;
  const int8_t  *push_peer;
  unsigned int  *ticket;
  int            n_peers;
  int           *err;
};

// Once-per-device guard for cudaFuncSetAttribute: function attributes belong to the device's context, and a process may
// hold contexts on several devices (one bit per device ordinal in a function-local static mask).
inline bool vh_first_time_on_device(unsigned long long &mask, int device)
{
  const unsigned long long bit = 1ull << (device & 63);
  if (mask & bit)
    return false;
  mask |= bit;
  return true;
}

struct VhCoef
{
  double K1, K23, alpha, beta[5], bt;
};

struct VhConstraintsDev
{
  int32_t  n_lines  = 0;
  int32_t *line_of  = nullptr; // [n_local_dofs] line index or -1
  int32_t *dof      = nullptr; // [n_lines]
  int32_t *ptr      = nullptr; // [n_lines+1]
  int32_t *master   = nullptr;
  double  *weight   = nullptr;
  bool     has_masters = false;
};

// Multigrid V-cycle parameters (vh_mg_params of the public header; defaults = what tools/mg_experiment.py measured on C5)
struct VhMGParams
{
  int    pre = 1, post = 1;  // Chebyshev degree of the pre- / post-smoother
  double range = 4.0;        // smoother interval [lambda_max / range, lambda_max] of M^-1 A
  int    coarse_degree = 8;  // Chebyshev degree on the coarsest level
  double coarse_range  = 30.0;
  int    n_power = 8;        // power iterations for lambda_max (once per level and context)
  double safety  = 1.1;
};

struct vh_ctx
{
  int          device = 0;
  cudaStream_t stream = nullptr;     // the stream every kernel of this context is launched on (a coarse multigrid level: its parent's)
  cudaStream_t own_stream = nullptr; // the stream this context created
  std::string  err;

  // sizes
  int     degree = 1, nn = 8, nq = 8, dpc = 144;
  int32_t n_owned = 0, n_ghost = 0, n_local = 0, n_cells = 0;
  int64_t NO = 0, NL = 0; // owned / local DoFs
  int64_t nnzb = 0;

  // mesh tables (device)
  int32_t  *cell_nodes = nullptr; // [n_cells][nn]
  double   *cell_h     = nullptr; // [n_cells][4]  hx,hy,hz,vol
  uint32_t *cell_faces = nullptr; // [n_cells]     4 bits per face: boundary id (0 = none)
  uint8_t  *cell_owned = nullptr;
  uint32_t *dirmask    = nullptr; // [n_local] bit c: DoF (node,c) is homogeneous Dirichlet in constraints_newton_update
  VhConstraintsDev cons[2];       // 0: newton update, 1: solution
  VhTables tab{};
  VhCoef   coef{};
  bool     coef_set = false;

  // BSR(18) matrix over owned rows
  int32_t *row_ptr  = nullptr; // [n_owned+1]
  int32_t *col      = nullptr; // [nnzb] local node ids, ascending inside a row
  double  *vals     = nullptr; // [nnzb][18][18]  full blocks of the constrained (row-owner scatter) rows
  // Packed storage of the lattice ("fast") rows: block = Sym(P) + kron(I_6, M_slot) with Dirichlet masks applied on
  // the fly, P = 171 unique entries of the symmetric bulk part (+9 zero dummies): 1440 B per block instead of 2592 B.
  bool     packed   = false;   // the context has lattice rows
  double  *pvals    = nullptr; // [nnzb][180], allocated on first use (vhk_alloc_rows)
  uint32_t *spmv_lane_tab   = nullptr; // [3][32]  lane constants of k_spmv_sym18
  uint16_t *spmv_gather_tab = nullptr; // [26][18] row-end gather lists of k_spmv_sym18
  double   *xmask   = nullptr; // [NL] scratch: Dirichlet-masked copy of an SpMV input (public vh_spmv only)
  double  *cdiag    = nullptr; // [n_owned][18] constrained-diagonal values sum_cells |a_ii| (0 for unconstrained DoFs)
  int32_t *diag_pos = nullptr; // [n_owned] block index of (I,I)
  double  *minv     = nullptr; // [n_owned][18][18] inverse diagonal blocks (block-Jacobi)
  std::vector<int32_t> h_row_ptr, h_col, h_fast_index;

  // row classification
  int32_t  n_fast = 0, n_slow_rows = 0, n_slow_cells = 0;
  int32_t *fast_rows  = nullptr; // [n_fast]
  int32_t *fast_cells = nullptr; // [n_fast][8]   incident cells, -1 = absent (Q1: octant order, row node is local vertex 7-o)
  uint32_t *fast_first = nullptr; // [n_fast][8]  Q2: bit j = bx+3by+9bz set if cell k is the first one to write that slot
  int32_t *spmv_order = nullptr; // [n_fast]      permutation of the fast rows for the SpMV: longest rows first
  int8_t  *fast_a     = nullptr; // [n_fast][8]   local index of the row node in each incident cell
  int32_t  n_slots = 27, slot_stride = 32, diag_slot = 13; // stencil slots (2p+1)^3, row stride of fast_slot/fast_posslot
  int8_t  *fast_slot  = nullptr; // [n_fast][slot_stride]  stencil slot -> position in the row, -1 = absent
  uint8_t *fast_posslot = nullptr; // [n_fast][slot_stride] position in the row -> stencil slot
  int32_t *fast_index = nullptr; // [n_owned]     index into fast_rows or -1
  int32_t *fast_class = nullptr; // [n_fast]      geometry class of the row's stencil
  double  *class_tab  = nullptr; // [n_classes][27][12] per slot: GS[3][3] = sum vol/(h_x h_y) Gref, FS[3] = sum area Mf (x != normal)
  int32_t  n_classes  = 0;
  double  *class_M    = nullptr; // [n_classes][27][10] coefficient-dependent 3x3 geometry block per slot (entry 9 = 0)
  std::vector<double> h_class_tab;
  int32_t *slow_rows  = nullptr; // [n_slow_rows]
  uint8_t *row_slow   = nullptr; // [n_owned]
  // row-owner lists of the constrained rows: (cell, local node) pairs feeding slow row r (the node itself or a hanging
  // node that names it as a master)
  int32_t *srow_ptr = nullptr, *srow_cell = nullptr; // [n_slow_rows + 1], [srow_ptr[n_slow_rows]]
  int8_t  *srow_a = nullptr;
  int16_t *srow_posb = nullptr;
  double  *srow_wr = nullptr;
  uint32_t *srow_bcons = nullptr, *srow_cons = nullptr;
  int32_t *srow_posI = nullptr, *srow_mnode = nullptr;
  int16_t *srow_mpos = nullptr;
  // node -> incident (cell, local node) lists for the deterministic rhs gather of fast rows
  // (fast rows use fast_cells; kept for Q2 later)

  // per-cell scratch written by the pointwise kernel
  double *Hq   = nullptr; // [n_cells][nq][180]  packed symmetric bulk Hessian (x cell volume) at every quadrature point
  double *Rc   = nullptr; // [n_cells][dpc]      cell rhs (= -cell residual)
  double *Dc   = nullptr; // [n_cells][dpc]      cell matrix diagonal (for the constrained-diagonal rule)
  double *avgD = nullptr; // [n_cells]           mean |diag| of the cell matrix
  double *Ec   = nullptr; // [n_cells]           cell energy

  // vectors
  double *x_sol = nullptr, *x_trial = nullptr, *delta = nullptr, *zbuf = nullptr; // [NL]
  double *rhs = nullptr, *resid = nullptr, *w = nullptr, *tmpo = nullptr;          // [NO]
  double *V   = nullptr;                                                           // [(restart+1)][NO]
  int     V_cap = 0;
  double *Zb    = nullptr; // [restart][NO] z_j = M^-1 v_j, kept only with the multigrid preconditioner (vh_gmres.cu)
  int     Zb_cap = 0;
  // reductions
  double *partials = nullptr; // [VH_MAX_RED_BLOCKS]
  double *scal     = nullptr; // device scalars [VH_SCAL_COUNT]
  unsigned int *ticket = nullptr;
  double *h_pinned = nullptr; // pinned host scratch [VH_SCAL_COUNT]
  double *h_mgs = nullptr;    // pinned + mapped: the fused Gram-Schmidt kernel writes H(:,j) here, flag in the last slot
  unsigned long long h_mgs_seq = 0;

  // halo
  int                  rank = 0, n_ranks = 1;
  void                *nccl_comm = nullptr;
  std::shared_ptr<void> nccl_holder; // owns the communicator; shared between the contexts of one rank (vh_comm_share)
  // peer-memory mailboxes (NVLink P2P through CUDA IPC) for latency-bound scalar all-reduces; see vh_halo.cu
  int                  mgs_mode = -1;      // fused Gram-Schmidt: elements per thread (8 / 32), 64 = streaming variant, 1000 = kernel chain, -1 = undecided
  bool                 p2p = false;
  VhP2P                p2p_dev;            // by-value kernel argument
  void                *p2p_mbox = nullptr; // this rank's mailbox (device)
  void                *p2p_open[VH_P2P_MAX_RANKS] = {}; // peers' mailboxes opened with cudaIpcOpenMemHandle
  unsigned long long   p2p_seq = 0;        // number of all-reduces issued so far (identical on every rank)
  int                 *p2p_err = nullptr;  // device flag: a wait timed out
  unsigned int        *mgs_tickets = nullptr; // [VH_MAX_RESTART + 2] arrival counters of the fused Gram-Schmidt kernel
  std::vector<int32_t> peer_rank, send_ptr, recv_ptr;
  int32_t             *send_nodes = nullptr, *recv_nodes = nullptr; // device lists
  double              *send_buf = nullptr, *recv_buf = nullptr;
  int64_t              n_send = 0, n_recv = 0;
  std::vector<int32_t> h_send_nodes, h_recv_nodes;
  // fused ghost push of the GMRES preconditioner kernel: z = M^-1 v of an interface node goes straight into the ghost slots of
  // the neighbours' zbuf over NVLink (peer memory through CUDA IPC), completion flags in the mailboxes; see vh_halo.cu
  bool      zpush = false;
  VhPush    zpush_dev;                         // by-value kernel argument
  void     *zpush_open[VH_P2P_MAX_RANKS] = {}; // peers' zbuf mapped with cudaIpcOpenMemHandle (indexed by peer slot)
  int32_t  *push_ptr = nullptr, *push_dst = nullptr; // CSR over owned nodes: destination ghost node index in the peer ...
  int8_t   *push_peer = nullptr;                      // ... and the peer slot
  unsigned int *push_ticket = nullptr;
  unsigned long long zpush_seq = 0;

  // operator apply of the lattice rows inside vhk_spmv: false = packed SpMV over the assembled blocks (default),
  // true = matrix-free from the H_q tables (VH_SPMV_MF=1; opt-in until measured on hardware)
  bool spmv_mf = false;
  // with spmv_mf: evaluate H(A_q) z_q from the Newton state instead of reading the H_q tables (VH_SPMV_MF=2 or
  // vh_set_spmv_matrix_free(ctx, 2); written after the round-1 GPU budget was spent: unverified on hardware)
  bool spmv_mf_table_free = false;
  bool spmv_mf_v2         = false; // mode 3 (Q1): second formulation of the table apply, csrc/vh_apply_v2.cuh (unverified on hardware)
  // VH_MF_LAZY_ROWS=1 (with spmv_mf): vh_assemble does not assemble the lattice rows — only their diagonal blocks, which
  // block-Jacobi needs (k_diag_cells + k_diag_gather); whoever needs the rows later (packed SpMV after a mode switch,
  // vh_export_matrix_bsr) assembles them on demand from the H_q tables.  Unverified on hardware in round 1.
  bool    rows_lazy = false, rows_stale = false;
  double *Dblk      = nullptr; // [n_cells][nn][180] per-(cell, node) diagonal contributions (allocated on first use)
  double *dpack     = nullptr; // [n_fast][180] packed diagonal blocks of the lattice rows (what block-Jacobi reads while rows_stale)

  // preconditioner of vh_solve: 0 = nodal block-Jacobi, 1 = multigrid V-cycle over the attached coarse levels (vh_multigrid.cu)
  int        precond = 0;
  vh_ctx    *mg_coarse = nullptr, *mg_parent = nullptr;
  VhMGParams mg_params;
  double     mg_lam = 0.0;                                  // safety * lambda_max(M^-1 A) of this level; 0 = not estimated yet
  double    *mg_x = nullptr, *mg_r = nullptr;               // [NL] level solution / residual (ghosted)
  double    *mg_b = nullptr, *mg_d = nullptr, *mg_t = nullptr; // [NO] level rhs, Chebyshev direction, A x
  int32_t   *mg_p_ptr = nullptr, *mg_p_coarse = nullptr;    // P: rows = owned nodes of this level, entries = LOCAL nodes of the coarse level
  double    *mg_p_w = nullptr;
  int32_t   *mg_rt_ptr = nullptr, *mg_rt_fine = nullptr;    // P^T: rows = owned nodes of the coarse level, entries = LOCAL nodes of this level
  double    *mg_rt_w = nullptr;
  int32_t   *mg_inj = nullptr;                              // [coarse n_owned] coincident node of this level
  int64_t   *node_global_dev = nullptr;                     // [n_owned] global node ids (start vector of the power iteration)

  // state flags
  bool have_matrix = false, have_update = false, have_trial = false;

  // accounting
  int64_t     n_launches = 0;
  double      t_ms[5]    = {0, 0, 0, 0, 0};
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr, ev3 = nullptr, ev4 = nullptr, ev_scal = nullptr;
  int64_t     device_bytes = 0;
  void       *flush_buf   = nullptr;
  size_t      flush_bytes = 0;
  // output path (vh_snapshot_begin / vh_snapshot_wait): device staging + pinned host copy of [local_solution | Newton update]
  double      *snap_dev = nullptr, *snap_host = nullptr;
  cudaStream_t snap_stream = nullptr;
  cudaEvent_t  ev_snap_ready = nullptr, ev_snap_done = nullptr;
  bool         snap_pending = false;
  bool         delta_holds_update = false; // delta still holds the last Newton update (also after vh_accept_trial)
};

#define VH_MAX_RED_BLOCKS 2048 /* ctx->partials holds 32x this many doubles (fused Gram-Schmidt: one set per step) */
// layout of the device scalar scratch ctx->scal
#define VH_MAX_RESTART 100
#define VH_SCAL_HCOL 0    /* Hessenberg column of the current GMRES inner step (<= VH_MAX_RESTART+1) */
#define VH_SCAL_NRM2 120  /* squared norm feeding the next basis-vector scaling */
#define VH_SCAL_Y 128     /* least-squares solution for the GMRES update (<= VH_MAX_RESTART) */
#define VH_SCAL_MISC 250
#define VH_SCAL_COUNT 256

// ---- error plumbing ----
int vh_fail(vh_ctx *ctx, int code, const std::string &msg);
#define VH_CUDA(call)                                                                                        \
  do                                                                                                         \
    {                                                                                                        \
      cudaError_t e__ = (call);                                                                              \
      if (e__ != cudaSuccess)                                                                                \
        return vh_fail(ctx, VH_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__));               \
    }                                                                                                        \
  while (0)
#define VH_TRY(call)           \
  do                           \
    {                          \
      int rc__ = (call);       \
      if (rc__ != VH_OK)       \
        return rc__;           \
    }                          \
  while (0)
#define VH_LAUNCH_CHECK()                                                                                    \
  do                                                                                                         \
    {                                                                                                        \
      ctx->n_launches++;                                                                                     \
      cudaError_t e__ = cudaGetLastError();                                                                  \
      if (e__ != cudaSuccess)                                                                                \
        return vh_fail(ctx, VH_ERR_CUDA, std::string("kernel launch: ") + cudaGetErrorString(e__));          \
    }                                                                                                        \
  while (0)

// ---- assembly (vh_assemble.cu) ----
int vhk_pointwise(vh_ctx *ctx, const double *x_local, bool want_hessian, bool want_energy);
int vhk_rows_fast(vh_ctx *ctx);
int vhk_rows_slow(vh_ctx *ctx, bool want_matrix, double *rhs_out);
int vhk_rhs_fast(vh_ctx *ctx, double *rhs_out, bool with_cdiag);
int vhk_upload_w1(vh_ctx *ctx, const double *W1);
int vhk_upload_q2(vh_ctx *ctx, const double *T2, const uint8_t *q2t);

// ---- linear algebra (vh_linalg.cu) ----
// x_is_masked: the caller guarantees zeros at the homogeneous-Dirichlet DoFs of x (all Krylov vectors satisfy this)
int vhk_spmv(vh_ctx *ctx, const double *x_local, double *y_owned, bool x_is_masked);
int vhk_upload_linalg_constants(vh_ctx *ctx);
int vhk_expand_packed(vh_ctx *ctx, double *full_vals);
// matrix-free apply of the lattice rows from the stored H_q tables (ctx->spmv_mf, VH_SPMV_MF=1); see k_points<APPLY>
int vhk_apply_fast(vh_ctx *ctx, const double *z_masked, const double *x_orig, double *y_owned);
// diagonal packed blocks of the lattice rows from the H_q tables (rows_lazy); vhk_ensure_rows assembles stale rows on demand
int vhk_diag_fast(vh_ctx *ctx);
int vhk_ensure_rows(vh_ctx *ctx);
int vhk_alloc_rows(vh_ctx *ctx); // packed row storage (pvals) is allocated on first use: the matrix-free default never needs it
int vhk_block_jacobi_setup(vh_ctx *ctx);
int vhk_block_jacobi_apply(vh_ctx *ctx, const double *x_owned, double *y_owned);
// push = true (only with ctx->zpush and y_owned == ctx->zbuf): also store the interface values into the neighbours' ghost slots
int vhk_block_jacobi_apply_scaled(vh_ctx *ctx, const double *x_owned, const double *nsq_dev, double *v_out, double *y_owned,
                                  bool push = false);
int vhk_mgs_fused(vh_ctx *ctx, double *w, const double *V, int64_t ld, int j, double *hcol_dev, bool *used);
// out_scalar[0] = sum_i a[i]*b[i] over owned DoFs (all-reduced over ranks); stream-ordered, result on device
int vhk_dot(vh_ctx *ctx, const double *a, const double *b, double *out_scalar);
// w += (-*coef) * v ; out = dot(w, u)   (deal.II Vector::add_and_dot with a = -coef)
int vhk_add_and_dot(vh_ctx *ctx, double *w, const double *coef_dev, const double *v, const double *u, double *out_scalar);
int vhk_scale_to(vh_ctx *ctx, double *dst, const double *src, const double *norm_sq_dev); // dst = src / sqrt(*norm_sq)
int vhk_axpy_dev(vh_ctx *ctx, double *y, const double *coefs_dev, int k, const double *V, int64_t ld); // y += sum_j c_j V_j
int vhk_axpby(vh_ctx *ctx, double *z, double a, const double *x, double b, const double *y, int64_t n); // z = a x + b y
int vhk_distribute(vh_ctx *ctx, int which, double *x_local); // AffineConstraints::distribute on owned constrained DoFs
int vhk_sum(vh_ctx *ctx, const double *a, const uint8_t *mask, int64_t n, double *out_scalar); // masked sum, all-reduced
int vh_read_scalars(vh_ctx *ctx, const double *dev, int n, double *host);                     // D2H + sync

// ---- halo (vh_halo.cu) ----
int vhk_halo_exchange(vh_ctx *ctx, double *x_local);
int vhk_allreduce_sum(vh_ctx *ctx, double *dev, int n);
int vhk_halo_wait(vh_ctx *ctx); // after a pushing vhk_block_jacobi_apply_scaled: wait for the neighbours' pushes
int vhk_mgs_mode_local(vh_ctx *ctx);

void vh_comm_destroy(vh_ctx *ctx);
int  vh_p2p_alloc_local(vh_ctx *ctx, int n_ranks);

// ---- multigrid preconditioner (vh_multigrid.cu) ----
int  vhk_assemble_device(vh_ctx *ctx); // Jacobian phase of vh_assemble without timers / norms (vh_context.cu)
int  vhk_mg_setup(vh_ctx *fine);       // per Newton step: coarse states, coarse Jacobians, block-Jacobi inverses, eigenvalue bounds
int  vhk_mg_apply(vh_ctx *fine, const double *v_owned, double *z_local); // z = V-cycle(v), zero initial guess
void vhk_mg_detach(vh_ctx *ctx);

// ---- GMRES (vh_gmres.cu) ----
int vh_gmres(vh_ctx *ctx, double tol_abs, int max_it, int restart, int *iterations, double *final_res);

// ---- helpers ----
template <typename T>
int vh_dev_alloc(vh_ctx *ctx, T **p, size_t count)
{
  *p = nullptr;
  if (count == 0)
    count = 1;
  cudaError_t e = cudaMalloc((void **)p, count * sizeof(T));
  if (e != cudaSuccess)
    return vh_fail(ctx, VH_ERR_CUDA, std::string("cudaMalloc(") + std::to_string(count * sizeof(T)) + " B): " + cudaGetErrorString(e));
  ctx->device_bytes += (int64_t)(count * sizeof(T));
  return VH_OK;
}
template <typename T>
int vh_dev_upload(vh_ctx *ctx, T **p, const T *host, size_t count)
{
  VH_TRY(vh_dev_alloc(ctx, p, count));
  if (count)
    VH_CUDA(cudaMemcpy(*p, host, count * sizeof(T), cudaMemcpyHostToDevice));
  return VH_OK;
}

#endif
