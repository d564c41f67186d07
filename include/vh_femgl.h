/* vh_femgl.h — C ABI of the B200-native femgl Newton hot path.
 *
 * Drop-in boundary for the three hot private members of VerHem's FemGL<dim>
 *   void assemble_system();            /root/reference/femgl/inc/femgl.h:132, femgl/src/assemble.cc:108-372
 *   void compute_residual();           /root/reference/femgl/inc/femgl.h:133, femgl/src/residual.cc:109-297
 *   void solve(const double &tol);     /root/reference/femgl/inc/femgl.h:134, femgl/src/solve.cc:108-187
 * and the vector work of
 *   void newton_iteration();           /root/reference/femgl/inc/femgl.h:135, femgl/src/iteration.cc:109-215
 *
 * The host (deal.II in production; verkko-hem-repo_b200/host in this repository) keeps mesh, DoF
 * numbering, constraints and refinement, and hands FLAT TABLES to this library once per mesh.
 * Plain pointers and sizes only; every array passed in is copied, the caller keeps ownership.
 * One context <-> one MPI rank <-> one GPU.  Calls are blocking, collective across ranks, and
 * single-caller-thread per context.  All functions return VH_OK (0) or a negative error code;
 * vh_last_error() gives the text (the deal.II adapter rethrows it as std::runtime_error so the
 * handler of /root/reference/sol/src/main.cc:120-145 still works).
 *
 * Numbering conventions
 *   node     : FE support point carrying 18 consecutive DoFs; local DoF = 18*local_node + c
 *   c        : 0..8 = u[c/3][c%3] (Re A), 9..17 = v (Im A)        (femgl.cc:133-137)
 *   local nodes [0,n_owned) are owned by this rank (their matrix rows live here),
 *   [n_owned, n_owned+n_ghost) are ghosts (locally_relevant \ locally_owned, femgl.h:304-305).
 */
#ifndef VH_FEMGL_H
#define VH_FEMGL_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VH_OK 0
#define VH_ERR_ARG (-1)           /* bad argument / inconsistent tables            */
#define VH_ERR_CUDA (-2)          /* CUDA runtime error (no device, OOM, launch)    */
#define VH_ERR_NCCL (-3)          /* NCCL error or NCCL library not loadable        */
#define VH_ERR_NOT_CONVERGED (-4) /* SolverControl::NoConvergence (solve.cc:171)    */
#define VH_ERR_UNSUPPORTED (-5)
#define VH_ERR_STATE (-6)         /* call order violated (e.g. solve before assemble) */

#define VH_NCCL_UNIQUE_ID_BYTES 128

typedef struct vh_ctx vh_ctx;

/* Constraint table in CSR form over LOCAL DoFs: the closed AffineConstraints object
 * (femgl.h:307-308; built in setup_*.cc:162-258).  Line k constrains DoF dof[k] to
 * sum_{p in [ptr[k],ptr[k+1])} weight[p] * x[master[p]]  (+ 0: all inhomogeneities are zero,
 * dirichlet.h:118-124,151-157).  A line with no entries is a homogeneous Dirichlet DoF.
 * Masters must be unconstrained (the object is closed) and local (owned or ghost). */
typedef struct vh_constraints
{
  int32_t        n_lines;
  const int32_t *dof;    /* [n_lines]   sorted ascending */
  const int32_t *ptr;    /* [n_lines+1] */
  const int32_t *master; /* [ptr[n_lines]] */
  const double  *weight; /* [ptr[n_lines]] */
} vh_constraints;

typedef struct vh_mesh_desc
{
  int32_t degree;        /* FE_Q degree: 1 or 2 (main.cc:116 hard-codes 1) */
  int32_t n_owned_nodes; /* locally_owned_dofs / 18                         */
  int32_t n_ghost_nodes;
  const int64_t *node_global; /* [n_owned+n_ghost] global node id (diagnostics, export) */

  /* Cells this rank visits: every active cell that can contribute to an owned row, i.e. the
   * locally owned cells plus one ghost layer (this replaces compress(add), assemble.cc:369-370).
   * cell_nodes: local node ids in deal.II FESystem(FE_Q(p),18) local order (vertices, lines, quads, hex). */
  int32_t        n_cells;
  const int32_t *cell_nodes;  /* [n_cells][(degree+1)^3]                       */
  const double  *cell_origin; /* [n_cells][3]  axis-aligned box cells          */
  const double  *cell_h;      /* [n_cells][3]                                  */
  const uint8_t *cell_owned;  /* [n_cells] cell->is_locally_owned() (energy is integrated over these) */

  /* Robin ("AdGR diffuse") wall faces, assemble.cc:286-292: boundary id 2|3|4 <-> normal x|y|z */
  int32_t        n_wall_faces;
  const int32_t *wall_face_cell; /* [n_wall_faces] index into cells        */
  const int8_t  *wall_face_no;   /* [n_wall_faces] deal.II face 0..5       */
  const int8_t  *wall_face_bid;  /* [n_wall_faces] 2|3|4                   */

  vh_constraints constraints_newton_update; /* femgl.h:307 */
  vh_constraints constraints_solution;      /* femgl.h:308 */

  /* Halo plan (ghosted assignment `locally_relevant = distributed`, solve.cc:183, iteration.cc:145,183,210) */
  int32_t        n_peers;
  const int32_t *peer_rank;  /* [n_peers]                                         */
  const int32_t *send_ptr;   /* [n_peers+1]                                       */
  const int32_t *send_nodes; /* owned local node ids to pack for each peer        */
  const int32_t *recv_ptr;   /* [n_peers+1]                                       */
  const int32_t *recv_nodes; /* ghost local node ids filled from each peer        */
} vh_mesh_desc;

/* Host-only consistency check of a descriptor (no CUDA call, works on a machine without a GPU): sizes, null arrays,
 * index ranges, sorted and closed constraint tables, masters that stay in their component's node range, wall-face and
 * halo-plan entries.  Returns VH_OK or VH_ERR_ARG with the first finding in msg (may be NULL).  vh_create runs it first. */
int vh_validate_mesh_desc(const vh_mesh_desc *desc, char *msg, int msg_len);

/* ---- lifetime (one context per mesh; destroy and re-create after refine_grid, refine.cc:109-181) ---- */
int         vh_create(const vh_mesh_desc *desc, int cuda_device, vh_ctx **out);
int         vh_destroy(vh_ctx *ctx);
const char *vh_last_error(const vh_ctx *ctx); /* ctx may be NULL: error of the last failed vh_create */

/* ---- multi-GPU: NCCL communicator over the ranks of one box (replaces MPI_COMM_WORLD, femgl.cc:111) ---- */
int vh_nccl_unique_id(void *id_out /* VH_NCCL_UNIQUE_ID_BYTES */); /* call on rank 0, broadcast by the host */
int vh_comm_init(vh_ctx *ctx, int rank, int n_ranks, const void *unique_id);
/* Give ctx the communicator of another context of the same rank instead of creating one (the mesh of the next adaptive cycle,
 * the levels of a multigrid hierarchy): setup_system() runs once per refinement cycle (refine.cc:126,169) and a new NCCL
 * communicator costs seconds on 8 ranks.  Collective; the communicator is released with its last context. */
int vh_comm_share(vh_ctx *ctx, vh_ctx *donor);

/* ---- coefficients: K1..K3 (femgl.h:320-322), alpha/beta1..5 from Matep (femgl.cc:161-168), bt (femgl.h:338) ---- */
int vh_set_coefficients(vh_ctx *ctx, double K1, double K2, double K3, double alpha, const double beta[5], double bt);

/* ---- state: local_solution (femgl.h:313).  Host arrays hold the OWNED DoFs, 18*n_owned doubles ---- */
int vh_set_solution(vh_ctx *ctx, const double *owned);     /* H2D + ghost refresh */
int vh_get_solution(vh_ctx *ctx, double *owned);           /* D2H */
int vh_get_newton_update(vh_ctx *ctx, double *owned);      /* locally_relevant_newton_solution (femgl.h:312) */
int vh_get_rhs(vh_ctx *ctx, double *owned);                /* system_rhs (femgl.h:315)      */
/* Output path (FemGL::output_results, io.cc:106-170, runs after every Newton step: run.cc:221-227).
 * vh_snapshot_begin enqueues, behind everything already enqueued for this context, a copy of the LOCAL (owned, then ghost)
 * parts of local_solution and of the Newton update (femgl.h:312-313: the two vectors DataOut reads) and their transfer to
 * pinned host memory on a second stream; it returns at once and the next Newton step may start.  vh_snapshot_wait blocks
 * until that snapshot has landed and hands out the two arrays (18 * n_local doubles each, library-owned, valid until the
 * next vh_snapshot_begin or vh_destroy); the Newton update is all zero before the first vh_solve. */
int vh_snapshot_begin(vh_ctx *ctx);
int vh_snapshot_wait(vh_ctx *ctx, const double **solution_local, const double **update_local);
int vh_get_residual(vh_ctx *ctx, double *owned);           /* residual_vector (femgl.h:316) */

/* ---- refine_grid(): SolutionTransfer::interpolate + constraints_solution.distribute + ghosted copy on the device
 *      (refine.cc:128-130, 171-175; run.cc:182-195).  `dst` is the context of the NEW mesh, `src` the one of the old mesh
 *      (same device, still alive); row i of the CSR table belongs to owned node i of dst and lists the nodes of src
 *      (LOCAL ids, owned or ghost) whose shape functions are non-zero at that node, with their values:
 *          dst.local_solution[i][c] = sum_{k in [ptr[i], ptr[i+1])} weight[k] * src.local_solution[src_node[k]][c].
 *      The host builds the table from the parent/child relation of the cells it refined (deal.II: the same data
 *      SolutionTransfer uses); nodes whose old cell lives on another rank must be filled by the host (vh_set_solution). ---- */
int vh_transfer_solution(vh_ctx *dst, vh_ctx *src, int32_t n_rows, const int32_t *ptr, const int32_t *src_node,
                         const double *weight);

/* ---- assemble_system(): system_matrix = R'(x), system_rhs = -R(x) at local_solution; returns ||rhs||_2 ---- */
int vh_assemble(vh_ctx *ctx, double *rhs_l2);

/* ---- solve(tol): GMRES(restart) with nodal 18x18 block-Jacobi, to ||r|| <= tol_rel*||rhs|| (solve.cc:159-160),
 *      then constraints_newton_update.distribute and ghost refresh (solve.cc:181-183).
 *      Iteration semantics follow deal.II SolverFGMRES.  VH_ERR_NOT_CONVERGED after max_it. ---- */
int vh_solve(vh_ctx *ctx, double tol_rel, int max_it, int restart, int *iterations, double *final_residual);

/* ---- preconditioner of vh_solve.  The reference builds a Trilinos-ML AMG hierarchy inside solve() (solve.cc:130-154);
 *      the north star prescribes nodal block-Jacobi, which is the default here.  For meshes that come with a hierarchy
 *      (global refinement: the host keeps the coarser meshes, deal.II: distribute_mg_dofs / MGTransfer) a geometric
 *      multigrid V-cycle can replace it: every coarser level is an ordinary context created from that level's tables on
 *      the same rank partition (vh_create, vh_comm_init, vh_set_coefficients), attached with its prolongation
 *          x_fine[i] = sum_{k in [ptr[i], ptr[i+1])} weight[k] * x_coarse[coarse_node[k]]
 *      (row i = LOCAL node i of the fine level, owned and ghost; coarse_node = LOCAL ids on the coarse level; rows of owned
 *      nodes must be complete, rows of ghost nodes list the parents that are local).  Levels chain: attach level l+1 to l.
 *      The coarse Jacobians are re-discretised at the injected Newton state in every vh_solve. ---- */
typedef struct vh_mg_params
{
  int32_t pre, post;        /* Chebyshev degree of the pre- / post-smoother around block-Jacobi (default 1, 1); either may be 0
                             * (V(0,k): the right-hand side is restricted directly, one operator apply fewer per cycle) */
  double  smoothing_range;  /* smoother interval [lambda_max/range, lambda_max] of M^-1 A          (default 4)    */
  int32_t coarse_degree;    /* Chebyshev degree on the coarsest level                              (default 8)    */
  double  coarse_range;     /*                                                                     (default 30)   */
  int32_t n_power;          /* power iterations for lambda_max, once per level and context         (default 8)    */
  double  safety;           /* lambda_max is multiplied by this                                    (default 1.1)  */
} vh_mg_params;
int vh_mg_attach(vh_ctx *fine, vh_ctx *coarse, int32_t n_rows, const int32_t *ptr, const int32_t *coarse_node, const double *weight);
/* kind: 0 = block-Jacobi, 1 = multigrid V-cycle over the levels attached with vh_mg_attach; params may be NULL (defaults).
 * With NO level attached kind 1 is the cycle's coarsest-level solver alone: a Chebyshev polynomial of degree coarse_degree in
 * the block-Jacobi-preconditioned operator on [lambda_max/coarse_range, lambda_max] - the polynomial preconditioner for
 * meshes without a hierarchy (adaptive cycles).  Collective. */
int vh_set_preconditioner(vh_ctx *ctx, int kind, const vh_mg_params *params);
/* diagnostics: safety * lambda_max(M^-1 A) the smoother of level `level` (0 = ctx) uses; 0 before the first vh_solve */
int vh_mg_get_lambda(vh_ctx *ctx, int level, double *lambda_max);

/* ---- newton_iteration() pieces (iteration.cc:128-210) ---- */
int vh_line_search_trial(vh_ctx *ctx, double alpha); /* trial = x + alpha*delta; constraints_solution.distribute; ghosts */
int vh_residual(vh_ctx *ctx, double *l2);            /* compute_residual() on the trial vector ('l', residual.cc:164) */
int vh_accept_trial(vh_ctx *ctx);                    /* local_solution = trial (iteration.cc:210) */

/* ---- GL free-energy functional F[A]; which = 0: local_solution, 1: trial vector ---- */
int vh_energy(vh_ctx *ctx, int which, double *energy);

/* ---- introspection for tests and benchmarks (not part of the reference's interface) ---- */
typedef struct vh_info
{
  int64_t n_owned_dofs, n_local_dofs;
  int64_t nnzb;          /* stored 18x18 blocks in owned rows            */
  int64_t n_fast_rows;   /* rows assembled by the write-once row kernel  */
  int64_t n_slow_cells;  /* cells routed through the constrained scatter */
  int64_t device_bytes;  /* device memory held by the context            */
  int64_t n_packed_blocks; /* blocks stored as packed symmetric 18x18 (180 doubles) instead of 324 doubles */
  int64_t spmv_matrix_free; /* 1: vh_solve / vh_spmv apply the lattice rows matrix-free from the H_q tables (VH_SPMV_MF=1); 2: table-free */
} vh_info;
int vh_get_info(vh_ctx *ctx, vh_info *info);
/* BSR(18) copy of the owned rows: row_ptr[n_owned+1], col[nnzb] (local node ids), vals[nnzb][18][18] row-major */
int vh_export_matrix_bsr(vh_ctx *ctx, int32_t *row_ptr, int32_t *col, double *vals);
/* y_owned = system_matrix * x_owned (ghosts refreshed inside); host buffers */
int vh_spmv(vh_ctx *ctx, const double *x_owned, double *y_owned);
/* y = M^-1 x with M = nodal 18x18 diagonal blocks of system_matrix; host buffers */
int vh_precondition(vh_ctx *ctx, const double *x_owned, double *y_owned);
/* Device-timed kernels (CUDA events on the launching stream; average ms per launch over `reps`):
 *   what: 0 SpMV, 1 Jacobian+rhs assembly, 2 residual assembly, 3 block-Jacobi apply, 4 fused add_and_dot,
 *         5 pointwise kernel only, 6 row-owner Jacobian kernel only;
 *         collective (every rank must call): 9 = 20 ghost refreshes, 10 = 20 inner products with their all-reduce */
int vh_time_kernel(vh_ctx *ctx, int what, int reps, int flush_l2, float *ms_avg);
/* Cumulative device time (ms) and launch counts since the last reset:
 *   [0] assemble [1] residual [2] solve [3] line search vector ops [4] the preconditioner setup inside [2] (block inverses,
 *   multigrid level re-discretisation; solve.cc:130-154); n_launches = kernels launched */
int vh_get_timers(vh_ctx *ctx, double ms[5], int64_t *n_launches, int reset);
/* CUDA-event stopwatch on the library's stream (the stream every kernel of this context is launched on). */
int vh_timer_start(vh_ctx *ctx);
int vh_timer_stop(vh_ctx *ctx, float *ms);
/* FP64 DFMA peak of the device this context lives on, measured by a register-resident FMA chain (TFLOP/s). */
int vh_measure_fp64_peak(vh_ctx *ctx, double *tflops);
/* Operator apply of the lattice rows inside vh_solve / vh_spmv (collective: same value on every rank):
 *   0 = packed SpMV over the assembled blocks (default), 1 = matrix-free from the H_q tables of the last vh_assemble,
 *   2 = matrix-free and table-free: H(A_q) z_q evaluated from the Newton state (unverified on hardware in round 1),
 *   3 = second formulation of mode 1 at Q1 (per-cell geometry table, higher occupancy; unverified on hardware in round 1).
 * The initial value comes from the environment variable VH_SPMV_MF.  VH_ERR_UNSUPPORTED if the context has no packed
 * lattice rows (nothing to apply matrix-free). */
int vh_set_spmv_matrix_free(vh_ctx *ctx, int on);

#ifdef __cplusplus
}
#endif
#endif
