// Row gathers of the lattice rows: cell rhs -> system_rhs (+ the constrained-diagonal values), cell products of the
// matrix-free apply -> y.  Deterministic sums over the <= 8 incident cells of a row, Dirichlet rule applied here.  Same
// header for nvcc (included by vh_assemble.cu inside its anonymous namespace) and for the CPU emulation
// (tests/native/cuda_emu.h, tests/test_kernel_emulation.py).
#ifndef VH_GATHER_KERNELS_CUH
#define VH_GATHER_KERNELS_CUH

// rhs of fast rows: deterministic gather of the cell rhs over the incident cells (no atomics)
__global__ void k_rhs_fast(int n_fast, int dpc, const int32_t *__restrict__ fast_rows, const int32_t *__restrict__ fast_cells,
                           const int8_t *__restrict__ fast_a, const uint32_t *__restrict__ dirmask, const double *__restrict__ Rc,
                           double *__restrict__ rhs, const double *__restrict__ Dc, const double *__restrict__ avgD,
                           double *__restrict__ cdiag)
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
        s += Rc[(size_t)e * dpc + fast_a[(size_t)r * 8 + o] * 18 + c];
    }
  const bool masked = (dirmask[I] >> c) & 1u;
  if (masked)
    s = 0.0;
  rhs[(size_t)I * 18 + c] = s;
  if (cdiag)
    { // packed storage keeps the constrained-diagonal values sum_cells |a_ii| (distribute_local_to_global) beside the blocks
      double dsum = 0.0;
      if (masked)
        for (int o = 0; o < 8; ++o)
          {
            const int e = fast_cells[(size_t)r * 8 + o];
            if (e < 0)
              continue;
            double dv = fabs(Dc[(size_t)e * dpc + fast_a[(size_t)r * 8 + o] * 18 + c]);
            if (dv == 0.0)
              dv = avgD[e];
            dsum += dv;
          }
      cdiag[(size_t)I * 18 + c] = dsum;
    }
}

// y of the lattice rows from the cell products K_cell z_cell (k_points<APPLY>): deterministic gather over the incident cells;
// Dirichlet rows keep only their constrained-diagonal value times the unmasked input (as k_spmv_sym18 does)
__global__ void k_gather_apply(int n_fast, int dpc, const int32_t *__restrict__ fast_rows, const int32_t *__restrict__ fast_cells,
                               const int8_t *__restrict__ fast_a, const uint32_t *__restrict__ dirmask, const double *__restrict__ Yc,
                               const double *__restrict__ cdiag, const double *__restrict__ xo, double *__restrict__ y)
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
        s += Yc[(size_t)e * dpc + fast_a[(size_t)r * 8 + o] * 18 + c];
    }
  if ((dirmask[I] >> c) & 1u)
    s = cdiag[(size_t)I * 18 + c] * xo[(size_t)I * 18 + c];
  y[(size_t)I * 18 + c] = s;
}

#endif
