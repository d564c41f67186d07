// Restarted GMRES with right nodal block-Jacobi preconditioning — the GPU replacement of
//   SolverFGMRES<LA::MPI::Vector>::solve(system_matrix, update, system_rhs, preconditioner)
//   (/root/reference/femgl/src/solve.cc:156-176; ML-AMG replaced by block-Jacobi as the north star prescribes).
//
// Iteration semantics are deal.II's SolverFGMRES ([deal.II-internal]):
//   * x0 = 0, r = b - A x, beta = ||r||; SolverControl is checked with (accumulated_iterations, beta) first;
//   * inner step j: v_j = r/a, z_j = M^-1 v_j, w = A z_j, modified Gram-Schmidt through
//     H(0,j) = w.v_0, H(i,j) = add_and_dot(-H(i-1,j), v_{i-1}, v_i), H(j+1,j) = a = sqrt(add_and_dot(-H(j,j), v_j, w));
//   * from j > 0 on, the (j+1) x j leading Hessenberg block is solved in the least-squares sense
//     (Householder QR) and its residual is checked with ++accumulated_iterations;
//   * at restart / exit x += sum_{i < y.size()} y_i z_i.
// Because M is a fixed linear operator here, z_i is not stored: sum y_i z_i = M^-1 (sum y_i v_i), which is
// what keeps 38M-DoF problems inside 180 GB (SURVEY.md §7 hard part 5).
//
// Device/host split: vectors and all O(N) work stay on the device; the Gram-Schmidt coefficients are produced
// by device-side reductions and consumed by the next fused kernel straight from device memory.  The host reads
// one Hessenberg column per inner step (a single small D2H) to run the tiny QR and the stopping test.
#include "vh_internal.h"

#include <atomic>
#include <cstdio>
#include <cstdlib>

#include <cmath>
#include <vector>

namespace
{
// min || rhs - H1 y ||_2 for the (rows x cols) column-major-free dense H1 (row-major, ld = cols), Householder QR.
// Returns the residual norm (deal.II Householder::least_squares).
double least_squares(std::vector<double> H1, int rows, int cols, std::vector<double> rhs, std::vector<double> &y)
{
  for (int k = 0; k < cols; ++k)
    {
      double sigma = 0.0;
      for (int i = k; i < rows; ++i)
        sigma += H1[i * cols + k] * H1[i * cols + k];
      const double nrm = std::sqrt(sigma);
      if (nrm == 0.0)
        continue;
      const double akk = H1[k * cols + k];
      const double s   = akk >= 0 ? -nrm : nrm;
      // v = a_k - s e_k, beta = 1/(s*(s - akk)) hmm: use standard form  Hh = I - v v^T / (v^T v / 2)
      std::vector<double> v(rows, 0.0);
      for (int i = k; i < rows; ++i)
        v[i] = H1[i * cols + k];
      v[k] -= s;
      double vtv = 0.0;
      for (int i = k; i < rows; ++i)
        vtv += v[i] * v[i];
      if (vtv == 0.0)
        continue;
      for (int c = k; c < cols; ++c)
        {
          double d = 0.0;
          for (int i = k; i < rows; ++i)
            d += v[i] * H1[i * cols + c];
          d *= 2.0 / vtv;
          for (int i = k; i < rows; ++i)
            H1[i * cols + c] -= d * v[i];
        }
      double d = 0.0;
      for (int i = k; i < rows; ++i)
        d += v[i] * rhs[i];
      d *= 2.0 / vtv;
      for (int i = k; i < rows; ++i)
        rhs[i] -= d * v[i];
    }
  y.assign(cols, 0.0);
  for (int k = cols - 1; k >= 0; --k)
    {
      double s = rhs[k];
      for (int c = k + 1; c < cols; ++c)
        s -= H1[k * cols + c] * y[c];
      y[k] = H1[k * cols + k] != 0.0 ? s / H1[k * cols + k] : 0.0;
    }
  double res = 0.0;
  for (int i = cols; i < rows; ++i)
    res += rhs[i] * rhs[i];
  return std::sqrt(res);
}
} // namespace

int vh_gmres(vh_ctx *ctx, double tol_abs, int max_it, int restart, int *iterations, double *final_res)
{
  const int64_t NO = ctx->NO;
  const int     m  = restart;
  double       *aux = ctx->w;
  double       *hcol  = ctx->scal + VH_SCAL_HCOL; // device: H(0..j+1, j) of the current inner step (last entry squared)
  double       *ycoef = ctx->scal + VH_SCAL_Y;    // device: y for the solution update
  double       *nrm2  = ctx->scal + VH_SCAL_NRM2;

  // x0 = 0 (solve.cc:131: freshly constructed distributed_newton_update)
  VH_CUDA(cudaMemsetAsync(ctx->delta, 0, sizeof(double) * ctx->NL, ctx->stream));

  enum
  {
    ITERATE,
    SUCCESS,
    FAILURE
  };
  auto check = [&](int step, double val) {
    if (val <= tol_abs)
      return (int)SUCCESS;
    if (step >= max_it || std::isnan(val))
      return (int)FAILURE;
    return (int)ITERATE;
  };

  // VH_GMRES_TRACE=1: CUDA events around every kernel group of the inner loop, printed to stderr after the solve (diagnostic)
  static const bool   trace = getenv("VH_GMRES_TRACE") && getenv("VH_GMRES_TRACE")[0] == '1';
  std::vector<cudaEvent_t> tev;
  std::vector<const char *> tname;
  auto mark = [&](const char *name) {
    if (!trace)
      return;
    cudaEvent_t e;
    cudaEventCreate(&e);
    cudaEventRecord(e, ctx->stream);
    tev.push_back(e);
    tname.push_back(name);
  };
  static const bool   speculate_env = !(getenv("VH_GMRES_SPECULATE") && getenv("VH_GMRES_SPECULATE")[0] == '0');
  const bool          speculate = speculate_env && ctx->precond != 1;
  // v_j = aux / a and z = M^-1 v_j (owned part of zbuf): block-Jacobi does both in one pass (and may push the interface values
  // into the neighbours' ghost slots); the multigrid V-cycle takes the scaled vector as its right-hand side
  const bool mg   = ctx->precond == 1;
  const bool push = ctx->zpush && !mg;
  auto       precondition_scaled = [&](const double *src, const double *a2, double *vj) -> int {
    if (!mg)
      return vhk_block_jacobi_apply_scaled(ctx, src, a2, vj, ctx->zbuf, push);
    VH_TRY(vhk_scale_to(ctx, vj, src, a2));
    return vhk_mg_apply(ctx, vj, ctx->zbuf);
  };
  // A V-cycle costs several operator applies, so with the multigrid preconditioner z_j = M^-1 v_j is kept (as deal.II's
  // SolverFGMRES does: x += sum y_j z_j) instead of spending one more cycle on M^-1 (sum y_j v_j) at the end, and no inner
  // step is enqueued speculatively (a wasted step would cost a whole cycle, the host round trip it hides only ~20 us).
  if (mg && (!ctx->Zb || ctx->Zb_cap < m))
    {
      if (ctx->Zb)
        {
          cudaFree(ctx->Zb);
          ctx->device_bytes -= (int64_t)ctx->Zb_cap * NO * (int64_t)sizeof(double);
          ctx->Zb = nullptr;
        }
      VH_TRY(vh_dev_alloc(ctx, &ctx->Zb, (size_t)m * (size_t)NO));
      ctx->Zb_cap = m;
    }
  int                 accumulated = 0;
  double              res = 0.0;
  int                 state = ITERATE;
  bool                x_is_zero = true;
  std::vector<double> H((size_t)(m + 1) * m, 0.0), y;
  do
    {
      // aux = b - A x
      if (x_is_zero)
        VH_CUDA(cudaMemcpyAsync(aux, ctx->rhs, sizeof(double) * NO, cudaMemcpyDeviceToDevice, ctx->stream));
      else
        {
          VH_TRY(vhk_halo_exchange(ctx, ctx->delta));
          VH_TRY(vhk_spmv(ctx, ctx->delta, ctx->tmpo, true)); // delta = M^-1 (V y): zero at Dirichlet DoFs
          VH_TRY(vhk_axpby(ctx, aux, 1.0, ctx->rhs, -1.0, ctx->tmpo, NO));
        }
      VH_TRY(vhk_dot(ctx, aux, aux, nrm2));
      double beta2;
      VH_TRY(vh_read_scalars(ctx, nrm2, 1, &beta2));
      const double beta = std::sqrt(beta2);
      res               = beta;
      state             = check(accumulated, res);
      if (state == SUCCESS)
        break;
      std::fill(H.begin(), H.end(), 0.0);
      y.clear();
      double        a    = beta;
      const double *a2_d = nrm2; // device location of a^2
      // Inner loop.  The host needs H(:,j) only to decide whether to stop, so the next step's basis-vector scaling,
      // preconditioner application, ghost refresh and SpMV - which depend on device data only - are enqueued BEFORE the
      // host waits for the coefficients: the GPU never idles during the host round trip.  No speculation right before an
      // expected stop (residual within 4x of the tolerance) or at the end of a restart cycle; a wrongly speculated step
      // only overwrites scratch vectors.  The decision depends on replicated scalars, so all ranks take the same path.
      bool have_next = false; // v_j, z = M^-1 v_j and aux = A z of the coming step are already enqueued
      for (int j = 0; j < m; ++j)
        {
          double *vj = ctx->V + (size_t)j * NO;
          if (!have_next)
            {
              mark("step");
              // v_j = aux / a and z = M^-1 v_j (owned part of zbuf) in one pass; ghosts refreshed; aux = A z
              if (a != 0.0)
                VH_TRY(precondition_scaled(aux, a2_d, vj));
              else
                {
                  VH_CUDA(cudaMemsetAsync(vj, 0, sizeof(double) * NO, ctx->stream));
                  VH_CUDA(cudaMemsetAsync(ctx->zbuf, 0, sizeof(double) * NO, ctx->stream));
                }
              mark("apply");
              if (push && a != 0.0)
                VH_TRY(vhk_halo_wait(ctx)); // the neighbours pushed their interface values into our ghost slots
              else
                VH_TRY(vhk_halo_exchange(ctx, ctx->zbuf));
              mark("halo");
              if (mg) // keep z_j
                VH_CUDA(cudaMemcpyAsync(ctx->Zb + (size_t)j * NO, ctx->zbuf, sizeof(double) * NO, cudaMemcpyDeviceToDevice, ctx->stream));
              VH_TRY(vhk_spmv(ctx, ctx->zbuf, aux, true)); // z = M^-1 v_j: zero at Dirichlet DoFs because v_j is
              mark("spmv");
            }
          have_next = false;
          mark("pre-mgs");
          // modified Gram-Schmidt; the coefficients never leave the device between the fused steps
          bool fused = false;
          VH_TRY(vhk_mgs_fused(ctx, aux, ctx->V, NO, j, hcol, &fused));
          if (!fused)
            {
              VH_TRY(vhk_dot(ctx, aux, ctx->V, hcol + 0));
              for (int i = 1; i <= j; ++i)
                VH_TRY(vhk_add_and_dot(ctx, aux, hcol + (i - 1), ctx->V + (size_t)(i - 1) * NO, ctx->V + (size_t)i * NO, hcol + i));
              VH_TRY(vhk_add_and_dot(ctx, aux, hcol + j, vj, aux, hcol + j + 1));
            }
          mark("mgs");
          if (!fused)
            {
              VH_CUDA(cudaMemcpyAsync(ctx->h_pinned, hcol, (j + 2) * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
              VH_CUDA(cudaEventRecord(ctx->ev_scal, ctx->stream));
              // keep a^2 where the next inner step's scaling kernel reads it, before hcol is overwritten
              VH_CUDA(cudaMemcpyAsync(nrm2, hcol + j + 1, sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
            } // fused: the kernel itself wrote a^2 to nrm2 and the column to mapped host memory (no copy in the stream)
          a2_d = nrm2;
          if (speculate && j + 1 < m && accumulated + 2 <= max_it && !(j > 0 && res <= 4.0 * tol_abs))
            {
              mark("step");
              VH_TRY(precondition_scaled(aux, a2_d, ctx->V + (size_t)(j + 1) * NO));
              mark("apply");
              if (push)
                VH_TRY(vhk_halo_wait(ctx));
              else
                VH_TRY(vhk_halo_exchange(ctx, ctx->zbuf));
              mark("halo");
              VH_TRY(vhk_spmv(ctx, ctx->zbuf, aux, true));
              mark("spmv");
              have_next = true;
            }
          std::vector<double> hc(j + 2);
          if (fused)
            { // poll the sequence flag the kernel writes after the column (system-scope fence in between)
              volatile unsigned long long *flag = reinterpret_cast<volatile unsigned long long *>(ctx->h_mgs + VH_SCAL_COUNT - 1);
              long                         spins = 0;
              while (*flag != ctx->h_mgs_seq)
                if ((++spins & 0xfff) == 0)
                  { // a failed launch / sticky error would leave us spinning: look at the stream now and then
                    const cudaError_t q = cudaStreamQuery(ctx->stream);
                    if (q != cudaSuccess && q != cudaErrorNotReady)
                      return vh_fail(ctx, VH_ERR_CUDA, std::string("GMRES: ") + cudaGetErrorString(q));
                    if (q == cudaSuccess && *flag != ctx->h_mgs_seq)
                      return vh_fail(ctx, VH_ERR_CUDA, "GMRES: the Gram-Schmidt kernel finished without publishing its column");
                  }
              std::atomic_thread_fence(std::memory_order_acquire);
              for (int i = 0; i < j + 2; ++i)
                hc[i] = reinterpret_cast<volatile double *>(ctx->h_mgs)[i];
            }
          else
            {
              VH_CUDA(cudaEventSynchronize(ctx->ev_scal));
              for (int i = 0; i < j + 2; ++i)
                hc[i] = ctx->h_pinned[i];
            }
          for (int i = 0; i <= j; ++i)
            H[(size_t)i * m + j] = hc[i];
          a                          = std::sqrt(hc[j + 1]);
          H[(size_t)(j + 1) * m + j] = a;
          if (a == 0.0)
            have_next = false; // breakdown: the speculated step divided by zero; it is redone through the memset path
          if (j > 0)
            {
              const int           rows = j + 1, cols = j;
              std::vector<double> H1((size_t)rows * cols), prhs(rows, 0.0);
              for (int r = 0; r < rows; ++r)
                for (int c = 0; c < cols; ++c)
                  H1[(size_t)r * cols + c] = H[(size_t)r * m + c];
              prhs[0] = beta;
              res     = least_squares(H1, rows, cols, prhs, y);
              state   = check(++accumulated, res);
              if (state != ITERATE)
                break;
            }
        }
      // x += sum_i y_i z_i = M^-1 (sum_i y_i v_i)
      if (!y.empty())
        {
          for (size_t i = 0; i < y.size(); ++i)
            ctx->h_pinned[i] = y[i];
          VH_CUDA(cudaMemcpyAsync(ycoef, ctx->h_pinned, sizeof(double) * y.size(), cudaMemcpyHostToDevice, ctx->stream));
          if (mg) // x += sum_i y_i z_i with the stored z_i
            VH_TRY(vhk_axpy_dev(ctx, ctx->delta, ycoef, (int)y.size(), ctx->Zb, NO));
          else
            {
              VH_CUDA(cudaMemsetAsync(ctx->tmpo, 0, sizeof(double) * NO, ctx->stream));
              VH_TRY(vhk_axpy_dev(ctx, ctx->tmpo, ycoef, (int)y.size(), ctx->V, NO));
              VH_TRY(vhk_block_jacobi_apply(ctx, ctx->tmpo, ctx->zbuf));
              VH_TRY(vhk_axpby(ctx, ctx->delta, 1.0, ctx->delta, 1.0, ctx->zbuf, NO));
            }
          VH_CUDA(cudaStreamSynchronize(ctx->stream)); // h_pinned is reused by the next read
          x_is_zero = false;
        }
    }
  while (state == ITERATE);

  if (trace && !tev.empty())
    {
      cudaStreamSynchronize(ctx->stream);
      fprintf(stderr, "[gmres trace] %d iterations:", accumulated);
      for (size_t i = 1; i < tev.size(); ++i)
        {
          float ms = 0;
          cudaEventElapsedTime(&ms, tev[i - 1], tev[i]);
          fprintf(stderr, " %s %.1f", tname[i], ms * 1e3f);
        }
      fprintf(stderr, " (us)\n");
      for (cudaEvent_t e : tev)
        cudaEventDestroy(e);
    }
  *iterations = accumulated;
  *final_res  = res;
  if (ctx->p2p)
    { // a peer that never posted its partial sum makes the mailbox wait give up after ~10 s and raise this flag
      int h_err = 0;
      VH_CUDA(cudaMemcpyAsync(&h_err, ctx->p2p_err, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
      VH_CUDA(cudaStreamSynchronize(ctx->stream));
      if (h_err)
        return vh_fail(ctx, VH_ERR_NCCL, "peer-memory all-reduce: a rank did not arrive (mailbox wait timed out)");
    }
  if (state != SUCCESS)
    return vh_fail(ctx, VH_ERR_NOT_CONVERGED,
                   "GMRES: no convergence after " + std::to_string(accumulated) + " iterations, residual " + std::to_string(res));
  return VH_OK;
}
