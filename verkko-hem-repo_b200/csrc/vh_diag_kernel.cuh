// Diagonal 18x18 blocks of the lattice rows straight from the H_q tables — what block-Jacobi needs when the operator is
// applied matrix-free and the lattice rows are therefore not assembled (VH_MF_LAZY_ROWS=1):
//   P_II = sum_{cells e containing node I} sum_q w_q N_a(q)^2 (vol H_q)(e)        (a = local index of I in e)
// k_diag_cells writes the per-(cell, node) contributions in the packed layout (coalesced over the 180 packed entries),
// k_diag_gather sums the <= 8 contributions of a row in a fixed order into dpack[fast row] (packed layout), which
// k_block_invert reads instead of the assembled diagonal block; its class_M / Dirichlet handling stays as it is.  Same header for nvcc and for the CPU emulation
// (tests/native/cuda_emu.h, tests/test_kernel_emulation.py).
#ifndef VH_DIAG_KERNEL_CUH
#define VH_DIAG_KERNEL_CUH

template <int NN>
__global__ void __launch_bounds__(192)
  k_diag_cells(int n_cells, const double *__restrict__ N, const double *__restrict__ wq, const double *__restrict__ Hq,
               double *__restrict__ Dblk)
{
  constexpr int     NQ = NN;
  __shared__ double sW[NN * NQ]; // w_q N_a(q)^2
  for (int i = threadIdx.x; i < NN * NQ; i += blockDim.x)
    sW[i] = wq[i % NQ] * N[i] * N[i];
  __syncthreads();
  const int     e    = threadIdx.x;
  const int64_t cell = blockIdx.x;
  if (e >= VH_SYMP || cell >= n_cells)
    return;
  double        h[NQ];
  const double *H = Hq + cell * (int64_t)(NQ * VH_SYMP);
#pragma unroll
  for (int q = 0; q < NQ; ++q)
    h[q] = H[NN == 8 ? vh_hq8_index(q, e) : q * VH_SYMP + e];
  double *D = Dblk + cell * (int64_t)(NN * VH_SYMP) + e;
#pragma unroll(NN == 8 ? 8 : 1)
  for (int a = 0; a < NN; ++a)
    {
      double acc = 0.0;
#pragma unroll
      for (int q = 0; q < NQ; ++q)
        acc = fma(sW[a * NQ + q], h[q], acc);
      D[a * VH_SYMP] = acc;
    }
}

__global__ void __launch_bounds__(192)
  k_diag_gather(int n_fast, int nn, const int32_t *__restrict__ fast_rows, const int32_t *__restrict__ fast_cells,
                const int8_t *__restrict__ fast_a, const double *__restrict__ Dblk, double *__restrict__ dpack)
{
  const int r = blockIdx.x, e = threadIdx.x;
  if (r >= n_fast || e >= VH_SYMP)
    return;
  double s = 0.0;
#pragma unroll
  for (int o = 0; o < 8; ++o)
    {
      const int c = fast_cells[(size_t)r * 8 + o];
      if (c >= 0)
        s += Dblk[((int64_t)c * nn + fast_a[(size_t)r * 8 + o]) * VH_SYMP + e];
    }
  dpack[(size_t)r * VH_SYMP + e] = s;
}

#endif
