// Geometric multigrid V-cycle as the right preconditioner of GMRES — the GPU replacement of the reference's Trilinos-ML AMG
//   preconditioner.initialize(system_matrix, additional_data)   (/root/reference/femgl/src/solve.cc:130-154)
// for meshes that come with a hierarchy (global refinement of a box: BASELINE configs C2/C3/C5).  Block-Jacobi
// (the north star's choice) stays the default and the fallback for meshes without a hierarchy (adaptive cycles).
//
// Every level is an ordinary context (vh_create on the coarser mesh, same partition): its operator is the RE-DISCRETISED
// Jacobian at the injected Newton state, applied matrix-free by the same k_points<APPLY> kernel; its smoother is the
// Chebyshev iteration around the nodal 18x18 block-Jacobi of that level.  The host supplies the prolongation between two
// consecutive levels as a CSR table over local nodes (deal.II: the entries of MGTransfer; mini host: trilinear weights).
//   V(b):  x = cheb_pre(b);  r = b - A x;  b_c = mask(P^T r);  x_c = V_c(b_c);  x += mask(P x_c);  x = cheb_post(b, x)
// with a zero initial guess, fixed degrees and fixed eigenvalue bounds: a FIXED LINEAR operator.  GMRES keeps z_j = V(v_j)
// with this preconditioner (deal.II's SolverFGMRES update x += sum y_j z_j; a cycle costs several operator applies and the
// iteration counts are single-digit, vh_gmres.cu).
// Exchange steps per level: ghost refresh of x before every operator apply, of r before the restriction and of x_c before the
// prolongation (NCCL halo of that level's plan); no reductions inside the cycle.
#include "vh_internal.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace
{
// d = c1 d + c2 (M^-1 r);  x = x + d.  flags bit 0: d has no previous value (d = c2 M^-1 r), bit 1: x has none (x = d).
// One warp per node, same streaming scheme as k_block_apply.
__global__ void __launch_bounds__(256)
  k_cheb_update(int n_rows, const double *__restrict__ minv, const double *__restrict__ r, double *__restrict__ d, double *__restrict__ x,
                double c1, double c2, int flags)
{
  __shared__ double s_part[8][6 * 32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int row  = blockIdx.x * 8 + wid;
  if (row >= n_rows)
    return;
  const double2 *B  = reinterpret_cast<const double2 *>(minv + (size_t)row * VH_BLK);
  const double  *rj = r + 18 * (size_t)row;
#pragma unroll
  for (int q = 0; q < 6; ++q)
    {
      const int k = lane + 32 * q;
      double    a = 0.0;
      if (k < 162)
        {
          const double2 m  = __ldg(B + k);
          const double2 rv = *reinterpret_cast<const double2 *>(rj + 2 * (k % 9));
          a                = fma(m.x, rv.x, m.y * rv.y);
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
      const size_t i  = (size_t)row * 18 + lane;
      const double dn = (flags & 1) ? c2 * s : fma(c1, d[i], c2 * s);
      d[i]            = dn;
      x[i]            = (flags & 2) ? dn : x[i] + dn;
    }
}

// b_c[I][c] = sum_k w_k r[f_k][c] over the fine LOCAL nodes feeding coarse owned node I (P^T), Dirichlet DoFs -> 0
__global__ void k_restrict(int n_rows, const int32_t *__restrict__ ptr, const int32_t *__restrict__ fine, const double *__restrict__ w,
                           const double *__restrict__ r_fine, const uint32_t *__restrict__ dirmask_c, double *__restrict__ b_c)
{
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= (int64_t)n_rows * 18)
    return;
  const int i = (int)(gid / 18), c = (int)(gid - 18 * (int64_t)i);
  double    s = 0.0;
  for (int k = ptr[i]; k < ptr[i + 1]; ++k)
    s = fma(w[k], r_fine[18 * (int64_t)fine[k] + c], s);
  b_c[gid] = ((dirmask_c[i] >> c) & 1u) ? 0.0 : s;
}

// x[f][c] += sum_k w_k x_c[c_k][c] for fine owned node f (P), Dirichlet DoFs stay 0
__global__ void k_prolong_add(int n_rows, const int32_t *__restrict__ ptr, const int32_t *__restrict__ coarse, const double *__restrict__ w,
                              const double *__restrict__ x_c, const uint32_t *__restrict__ dirmask_f, double *__restrict__ x)
{
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= (int64_t)n_rows * 18)
    return;
  const int i = (int)(gid / 18), c = (int)(gid - 18 * (int64_t)i);
  if ((dirmask_f[i] >> c) & 1u)
    {
      x[gid] = 0.0;
      return;
    }
  double s = 0.0;
  for (int k = ptr[i]; k < ptr[i + 1]; ++k)
    s = fma(w[k], x_c[18 * (int64_t)coarse[k] + c], s);
  x[gid] += s;
}

// coarse state <- the coincident fine node (injection of the Newton state for the re-discretised coarse Jacobian)
__global__ void k_inject(int n_rows, const int32_t *__restrict__ inj, const double *__restrict__ x_f, double *__restrict__ x_c)
{
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= (int64_t)n_rows * 18)
    return;
  const int i = (int)(gid / 18), c = (int)(gid - 18 * (int64_t)i);
  x_c[gid]    = x_f[18 * (int64_t)inj[i] + c];
}

// start vector of the power iteration: a fixed pattern of the GLOBAL DoF index (the same field on every partition)
__global__ void k_power_start(int n_owned, const int64_t *__restrict__ node_global, const uint32_t *__restrict__ dirmask, double *__restrict__ v)
{
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= (int64_t)n_owned * 18)
    return;
  const int     i = (int)(gid / 18), c = (int)(gid - 18 * (int64_t)i);
  const int64_t g = 18 * node_global[i] + c;
  double        s = ((g & 1) ? -1.0 : 1.0) * (1.0 + (double)(g % 7) / 7.0);
  if ((dirmask[i] >> c) & 1u)
    s = 0.0;
  v[gid] = s;
}

int ensure_work(vh_ctx *L)
{
  vh_ctx *ctx = L;
  if (L->mg_x)
    return VH_OK;
  VH_TRY(vh_dev_alloc(ctx, &L->mg_x, (size_t)L->NL));
  VH_TRY(vh_dev_alloc(ctx, &L->mg_r, (size_t)L->NL));
  VH_TRY(vh_dev_alloc(ctx, &L->mg_b, (size_t)L->NO));
  VH_TRY(vh_dev_alloc(ctx, &L->mg_d, (size_t)L->NO));
  VH_TRY(vh_dev_alloc(ctx, &L->mg_t, (size_t)L->NO));
  VH_CUDA(cudaMemsetAsync(L->mg_x, 0, sizeof(double) * (size_t)std::max<int64_t>(L->NL, 1), L->stream));
  VH_CUDA(cudaMemsetAsync(L->mg_r, 0, sizeof(double) * (size_t)std::max<int64_t>(L->NL, 1), L->stream));
  return VH_OK;
}

// r = b - A x on level L (x: local vector whose owned part is current; ghosts are refreshed here); r: owned part of `r`
int residual(vh_ctx *L, const double *b, double *x, double *r)
{
  VH_TRY(vhk_halo_exchange(L, x));
  VH_TRY(vhk_spmv(L, x, L->mg_t, true));
  VH_TRY(vhk_axpby(L, r, 1.0, b, -1.0, L->mg_t, L->NO));
  return VH_OK;
}

// `degree` Chebyshev steps for M^-1 A on [lam / ratio, lam]; has_x = false: zero initial guess (no operator apply in step 0)
int chebyshev(vh_ctx *L, const double *b, double *x, bool has_x, int degree, double ratio)
{
  vh_ctx *ctx = L;
  if (L->n_owned == 0)
    { // a rank without owned nodes on this level still takes part in the ghost refreshes of the operator applies
      for (int i = has_x ? 0 : 1; i < degree; ++i)
        VH_TRY(vhk_halo_exchange(L, x));
      return VH_OK;
    }
  if (degree == 0)
    { // no smoothing on this side of the cycle: a zero initial guess stays zero
      if (!has_x)
        VH_CUDA(cudaMemsetAsync(x, 0, sizeof(double) * (size_t)L->NO, L->stream));
      return VH_OK;
    }
  const double lmax = L->mg_lam, lmin = L->mg_lam / ratio;
  const double theta = 0.5 * (lmax + lmin), delta = 0.5 * (lmax - lmin), sigma = theta / delta;
  double       rho = 1.0 / sigma;
  const unsigned grid = (unsigned)((L->n_owned + 7) / 8);
  const double  *r = b;
  if (has_x)
    {
      VH_TRY(residual(L, b, x, L->mg_r));
      r = L->mg_r;
    }
  k_cheb_update<<<grid, 256, 0, L->stream>>>(L->n_owned, L->minv, r, L->mg_d, x, 0.0, 1.0 / theta, has_x ? 1 : 3);
  VH_LAUNCH_CHECK();
  for (int i = 1; i < degree; ++i)
    {
      const double rho_new = 1.0 / (2.0 * sigma - rho);
      VH_TRY(residual(L, b, x, L->mg_r));
      k_cheb_update<<<grid, 256, 0, L->stream>>>(L->n_owned, L->minv, L->mg_r, L->mg_d, x, rho_new * rho, 2.0 * rho_new / delta, 0);
      VH_LAUNCH_CHECK();
      rho = rho_new;
    }
  return VH_OK;
}

// largest eigenvalue of M^-1 A on level L by n_power power iterations from the fixed start vector (once per context)
int estimate_lambda(vh_ctx *L, const VhMGParams &P)
{
  vh_ctx *ctx = L;
  double *nrm = L->scal + VH_SCAL_MISC + 5; // slots MISC .. MISC+4 belong to vh_context.cu (VH_SCAL_COUNT = MISC + 6)
  if (L->n_owned > 0)
    {
      const int64_t n = (int64_t)L->n_owned * 18;
      k_power_start<<<(unsigned)((n + 255) / 256), 256, 0, L->stream>>>(L->n_owned, L->node_global_dev, L->dirmask, L->mg_x);
      VH_LAUNCH_CHECK();
    }
  double lam2 = 0.0;
  for (int it = 0; it < P.n_power; ++it)
    {
      VH_TRY(vhk_dot(L, L->mg_x, L->mg_x, nrm));
      VH_TRY(vhk_scale_to(L, L->mg_x, L->mg_x, nrm)); // v /= ||v||
      VH_TRY(vhk_halo_exchange(L, L->mg_x));
      VH_TRY(vhk_spmv(L, L->mg_x, L->mg_t, true));
      VH_TRY(vhk_block_jacobi_apply(L, L->mg_t, L->mg_x)); // v = M^-1 A v
    }
  VH_TRY(vhk_dot(L, L->mg_x, L->mg_x, nrm));
  VH_TRY(vh_read_scalars(L, nrm, 1, &lam2));
  if (!(lam2 > 0.0) || !std::isfinite(lam2))
    return vh_fail(ctx, VH_ERR_ARG, "multigrid: the power iteration for lambda_max(M^-1 A) broke down");
  L->mg_lam = P.safety * std::sqrt(lam2);
  return VH_OK;
}

// VH_MG_TRACE=1: CUDA events around the parts of every cycle, printed to stderr after the cycle (diagnostic, like VH_GMRES_TRACE)
struct MgTrace
{
  std::vector<cudaEvent_t>  ev;
  std::vector<const char *> name;
  std::vector<int>          level;
  bool                      on = getenv("VH_MG_TRACE") && getenv("VH_MG_TRACE")[0] == '1';
  void mark(vh_ctx *L, const char *what, int lv)
  {
    if (!on)
      return;
    cudaEvent_t e;
    cudaEventCreate(&e);
    cudaEventRecord(e, L->stream);
    ev.push_back(e);
    name.push_back(what);
    level.push_back(lv);
  }
  void flush(vh_ctx *L)
  {
    if (!on || ev.empty())
      return;
    cudaStreamSynchronize(L->stream);
    fprintf(stderr, "[mg trace]");
    for (size_t i = 1; i < ev.size(); ++i)
      {
        float ms = 0;
        cudaEventElapsedTime(&ms, ev[i - 1], ev[i]);
        fprintf(stderr, " L%d.%s %.0f", level[i], name[i], ms * 1e3f);
      }
    fprintf(stderr, " (us)\n");
    for (cudaEvent_t e : ev)
      cudaEventDestroy(e);
    ev.clear();
    name.clear();
    level.clear();
  }
};
MgTrace g_trace;

int vcycle(vh_ctx *L, const VhMGParams &P, const double *b, double *x, int lv = 0)
{
  vh_ctx *ctx = L;
  vh_ctx *C   = L->mg_coarse;
  if (!C)
    {
      VH_TRY(chebyshev(L, b, x, false, P.coarse_degree, P.coarse_range));
      g_trace.mark(L, "coarse", lv);
      return VH_OK;
    }
  VH_TRY(chebyshev(L, b, x, false, P.pre, P.range));
  g_trace.mark(L, "pre", lv);
  // r = b - A x, ghosts refreshed: the restriction of a coarse owned node reads fine nodes owned by the neighbours
  if (P.pre > 0)
    VH_TRY(residual(L, b, x, L->mg_r));
  else if (L->NO > 0) // x = 0: the residual is the right-hand side itself, no operator apply
    VH_CUDA(cudaMemcpyAsync(L->mg_r, b, sizeof(double) * (size_t)L->NO, cudaMemcpyDeviceToDevice, L->stream));
  VH_TRY(vhk_halo_exchange(L, L->mg_r));
  g_trace.mark(L, "residual", lv);
  if (C->n_owned > 0)
    {
      const int64_t n = (int64_t)C->n_owned * 18;
      k_restrict<<<(unsigned)((n + 255) / 256), 256, 0, L->stream>>>(C->n_owned, L->mg_rt_ptr, L->mg_rt_fine, L->mg_rt_w, L->mg_r, C->dirmask,
                                                                   C->mg_b);
      VH_LAUNCH_CHECK();
    }
  g_trace.mark(L, "restrict", lv);
  VH_TRY(vcycle(C, P, C->mg_b, C->mg_x, lv + 1));
  VH_TRY(vhk_halo_exchange(C, C->mg_x));
  if (L->n_owned > 0)
    {
      const int64_t n = (int64_t)L->n_owned * 18;
      k_prolong_add<<<(unsigned)((n + 255) / 256), 256, 0, L->stream>>>(L->n_owned, L->mg_p_ptr, L->mg_p_coarse, L->mg_p_w, C->mg_x, L->dirmask, x);
      VH_LAUNCH_CHECK();
    }
  g_trace.mark(L, "prolong", lv);
  VH_TRY(chebyshev(L, b, x, true, P.post, P.range));
  g_trace.mark(L, "post", lv);
  return VH_OK;
}
} // namespace

// Per Newton step ("Solve: setup preconditioner", solve.cc:133): coarse states by injection, re-discretised coarse
// Jacobians (H_q tables + diagonal blocks), their block-Jacobi inverses; eigenvalue bounds on first use.
int vhk_mg_setup(vh_ctx *fine)
{
  vh_ctx *ctx = fine;
  for (vh_ctx *L = fine; L; L = L->mg_coarse)
    {
      VH_TRY(ensure_work(L));
      vh_ctx *C = L->mg_coarse;
      if (!C)
        break;
      if (!C->coef_set)
        return vh_fail(ctx, VH_ERR_STATE, "multigrid: vh_set_coefficients was not called on a coarse level");
      if (C->n_owned > 0)
        {
          const int64_t n = (int64_t)C->n_owned * 18;
          k_inject<<<(unsigned)((n + 255) / 256), 256, 0, L->stream>>>(C->n_owned, L->mg_inj, L->x_sol, C->x_sol);
          VH_LAUNCH_CHECK();
        }
      VH_TRY(vhk_halo_exchange(C, C->x_sol));
      VH_TRY(vhk_assemble_device(C));
      VH_TRY(vhk_block_jacobi_setup(C));
      C->have_matrix = true;
      C->have_update = C->have_trial = false;
    }
  for (vh_ctx *L = fine; L; L = L->mg_coarse)
    if (L->mg_lam == 0.0)
      VH_TRY(estimate_lambda(L, fine->mg_params));
  return VH_OK;
}

// z = V(v): one V-cycle with zero initial guess; v: owned vector, z: LOCAL vector (owned part written, ghosts stale)
int vhk_mg_apply(vh_ctx *fine, const double *v_owned, double *z_local)
{
  g_trace.mark(fine, "start", 0);
  const int rc = vcycle(fine, fine->mg_params, v_owned, z_local);
  g_trace.flush(fine);
  return rc;
}

void vhk_mg_detach(vh_ctx *ctx)
{ // called by vh_destroy: unlink both directions, give the coarse level its own stream back
  if (ctx->mg_parent)
    {
      vh_ctx *p = ctx->mg_parent;
      p->mg_coarse = nullptr;
      if (p->precond == 1)
        p->precond = 0;
      ctx->stream    = ctx->own_stream;
      ctx->mg_parent = nullptr;
    }
  if (ctx->mg_coarse)
    {
      vh_ctx *c = ctx->mg_coarse;
      for (vh_ctx *k = c; k; k = k->mg_coarse)
        k->stream = k->own_stream; // the levels below ran on this context's stream
      c->mg_parent   = nullptr;
      ctx->mg_coarse = nullptr;
    }
  void *ptrs[] = {ctx->mg_x, ctx->mg_r, ctx->mg_b, ctx->mg_d, ctx->mg_t, ctx->mg_p_ptr, ctx->mg_p_coarse, ctx->mg_p_w, ctx->mg_rt_ptr,
                  ctx->mg_rt_fine, ctx->mg_rt_w, ctx->mg_inj};
  for (void *p : ptrs)
    if (p)
      cudaFree(p);
  ctx->mg_x = ctx->mg_r = ctx->mg_b = ctx->mg_d = ctx->mg_t = ctx->mg_p_w = ctx->mg_rt_w = nullptr;
  ctx->mg_p_ptr = ctx->mg_p_coarse = ctx->mg_rt_ptr = ctx->mg_rt_fine = ctx->mg_inj = nullptr;
}

extern "C" int vh_mg_attach(vh_ctx *fine, vh_ctx *coarse, int32_t n_rows, const int32_t *ptr, const int32_t *coarse_node, const double *weight)
{
  vh_ctx *ctx = fine;
  if (!fine)
    return vh_fail(nullptr, VH_ERR_ARG, "null context");
  if (!coarse || coarse == fine || coarse->device != fine->device || coarse->mg_parent || fine->mg_coarse)
    return vh_fail(ctx, VH_ERR_ARG, "vh_mg_attach: the coarse level must be another, unattached context on the same device");
  if (coarse->n_ranks != fine->n_ranks || coarse->rank != fine->rank)
    return vh_fail(ctx, VH_ERR_ARG, "vh_mg_attach: both levels must live on the same rank of communicators of the same size");
  if (n_rows != fine->n_local || !ptr || ptr[0] != 0 || (ptr[n_rows] > 0 && (!coarse_node || !weight)))
    return vh_fail(ctx, VH_ERR_ARG, "vh_mg_attach: the prolongation table must have one row per LOCAL node of the fine level");
  for (int i = 0; i < n_rows; ++i)
    if (ptr[i + 1] < ptr[i])
      return vh_fail(ctx, VH_ERR_ARG, "vh_mg_attach: ptr not monotone");
  const int nent = ptr[n_rows];
  for (int k = 0; k < nent; ++k)
    if (coarse_node[k] < 0 || coarse_node[k] >= coarse->n_local || !(weight[k] == weight[k]))
      return vh_fail(ctx, VH_ERR_ARG, "vh_mg_attach: coarse node out of range or weight not finite");
  // rows of the owned fine nodes must be complete interpolations (weights sum to 1): their parents are all local
  for (int i = 0; i < fine->n_owned; ++i)
    {
      double s = 0.0;
      for (int k = ptr[i]; k < ptr[i + 1]; ++k)
        s += weight[k];
      if (std::fabs(s - 1.0) > 1e-12)
        return vh_fail(ctx, VH_ERR_ARG, "vh_mg_attach: a fine owned node has parents that are not local on the coarse level "
                                        "(the two levels must share one nested partition)");
    }
  VH_CUDA(cudaSetDevice(fine->device));
  // P over the fine owned rows
  const int np = fine->n_owned ? ptr[fine->n_owned] : 0;
  VH_TRY(vh_dev_upload(ctx, &fine->mg_p_ptr, ptr, (size_t)fine->n_owned + 1));
  VH_TRY(vh_dev_upload(ctx, &fine->mg_p_coarse, coarse_node, (size_t)np));
  VH_TRY(vh_dev_upload(ctx, &fine->mg_p_w, weight, (size_t)np));
  // P^T restricted to the coarse OWNED nodes (rows) over the fine LOCAL nodes (entries), entries in ascending fine order;
  // injection: the fine node that coincides with the coarse node (weight 1)
  std::vector<int32_t> tptr((size_t)coarse->n_owned + 1, 0), inj((size_t)coarse->n_owned, -1);
  for (int k = 0; k < nent; ++k)
    if (coarse_node[k] < coarse->n_owned)
      tptr[coarse_node[k] + 1]++;
  for (int i = 0; i < coarse->n_owned; ++i)
    tptr[i + 1] += tptr[i];
  std::vector<int32_t> tf((size_t)tptr[coarse->n_owned]), fill(tptr.begin(), tptr.end() - 1);
  std::vector<double>  tw((size_t)tptr[coarse->n_owned]);
  for (int f = 0; f < n_rows; ++f)
    for (int k = ptr[f]; k < ptr[f + 1]; ++k)
      {
        const int c = coarse_node[k];
        if (c >= coarse->n_owned)
          continue;
        tf[fill[c]] = f;
        tw[fill[c]] = weight[k];
        fill[c]++;
        if (std::fabs(weight[k] - 1.0) <= 1e-12 && f < fine->n_owned)
          inj[c] = f;
      }
  for (int c = 0; c < coarse->n_owned; ++c)
    if (inj[c] < 0)
      return vh_fail(ctx, VH_ERR_ARG, "vh_mg_attach: a coarse owned node has no coincident fine node owned by this rank "
                                      "(nested meshes with the same ownership rule are required)");
  VH_TRY(vh_dev_upload(ctx, &fine->mg_rt_ptr, tptr.data(), tptr.size()));
  VH_TRY(vh_dev_upload(ctx, &fine->mg_rt_fine, tf.data(), tf.size()));
  VH_TRY(vh_dev_upload(ctx, &fine->mg_rt_w, tw.data(), tw.size()));
  VH_TRY(vh_dev_upload(ctx, &fine->mg_inj, inj.data(), inj.size()));
  // one stream for the whole hierarchy: every level below `fine` now launches on fine's stream
  VH_CUDA(cudaStreamSynchronize(coarse->stream));
  VH_CUDA(cudaStreamSynchronize(fine->stream));
  fine->mg_coarse   = coarse;
  coarse->mg_parent = fine;
  for (vh_ctx *k = coarse; k; k = k->mg_coarse)
    k->stream = fine->stream;
  return VH_OK;
}

extern "C" int vh_mg_get_lambda(vh_ctx *ctx, int level, double *lambda_max)
{
  if (!ctx || !lambda_max)
    return vh_fail(ctx, VH_ERR_ARG, "vh_mg_get_lambda: null argument");
  vh_ctx *L = ctx;
  for (int k = 0; k < level && L; ++k)
    L = L->mg_coarse;
  if (!L || level < 0)
    return vh_fail(ctx, VH_ERR_ARG, "vh_mg_get_lambda: no such level");
  *lambda_max = L->mg_lam;
  return VH_OK;
}

extern "C" int vh_set_preconditioner(vh_ctx *ctx, int kind, const vh_mg_params *p)
{
  if (!ctx)
    return vh_fail(nullptr, VH_ERR_ARG, "null context");
  if (kind != 0 && kind != 1)
    return vh_fail(ctx, VH_ERR_ARG, "vh_set_preconditioner: kind must be 0 (block-Jacobi) or 1 (multigrid)");
  // kind 1 without an attached coarse level: the cycle degenerates to its coarsest-level solver, i.e. a Chebyshev polynomial
  // of degree coarse_degree in the block-Jacobi-preconditioned operator on [lambda / coarse_range, lambda] - a polynomial
  // preconditioner for meshes without a hierarchy (adaptive cycles): GMRES(m) then spans m * coarse_degree operator powers
  // per restart cycle, which is what restarted GMRES with plain block-Jacobi lacks on strongly graded meshes.
  if (p)
    {
      if (p->pre < 0 || p->post < 0 || p->pre + p->post < 1 || p->coarse_degree < 1 || p->n_power < 1 || !(p->smoothing_range > 1.0) ||
          !(p->coarse_range > 1.0) || !(p->safety >= 1.0))
        return vh_fail(ctx, VH_ERR_ARG, "vh_set_preconditioner: pre, post >= 0 with pre + post >= 1, coarse_degree and n_power >= 1, "
                                        "ranges > 1, safety >= 1");
      ctx->mg_params.pre           = p->pre;
      ctx->mg_params.post          = p->post;
      ctx->mg_params.range         = p->smoothing_range;
      ctx->mg_params.coarse_degree = p->coarse_degree;
      ctx->mg_params.coarse_range  = p->coarse_range;
      ctx->mg_params.n_power       = p->n_power;
      ctx->mg_params.safety        = p->safety;
      for (vh_ctx *L = ctx; L; L = L->mg_coarse)
        L->mg_lam = 0.0; // re-estimate with the new power-iteration parameters
    }
  ctx->precond = kind;
  return VH_OK;
}
