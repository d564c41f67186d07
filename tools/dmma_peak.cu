// Micro-benchmark: FP64 tensor-core (mma.sync m8n8k4 -> SASS DMMA.8x8x4) vs DFMA peak on the device.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/dmma_peak tools/dmma_peak.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(256) k_dmma(double *out, int iters)
{
  double a = 1.0 + 1e-9 * threadIdx.x, b = 1.0 - 1e-9 * threadIdx.x;
  double c[8][2];
  for (int k = 0; k < 8; ++k) c[k][0] = c[k][1] = 0.0;
  for (int i = 0; i < iters; ++i)
#pragma unroll
    for (int k = 0; k < 8; ++k)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[k][0]), "+d"(c[k][1]) : "d"(a), "d"(b));
  double s = 0;
  for (int k = 0; k < 8; ++k) s += c[k][0] + c[k][1];
  if (s == 12345.678) out[0] = s;
}
__global__ void __launch_bounds__(256) k_dfma(double *out, int iters)
{
  double a[8];
  for (int k = 0; k < 8; ++k) a[k] = 1.0 + 1e-9 * (threadIdx.x + k);
  const double m = 1.0 + 1e-12, b = 1e-12;
  for (int i = 0; i < iters; ++i)
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] = fma(a[k], m, b);
  double s = 0;
  for (int k = 0; k < 8; ++k) s += a[k];
  if (s == 12345.678) out[0] = s;
}
int main()
{
  double *d; cudaMalloc(&d, 8);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int grid = 148 * 8, iters = 1 << 14;
  for (int which = 0; which < 2; ++which)
    for (int rep = 0; rep < 3; ++rep)
      {
        cudaEventRecord(e0);
        if (which == 0) k_dmma<<<grid, 256>>>(d, iters); else k_dfma<<<grid, 256>>>(d, iters);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        // DMMA: 8x8x4 = 256 FMA per warp instruction; DFMA: 32 FMA per warp instruction
        const double fma_per_thread = which == 0 ? 8.0 * iters * 256 / 32 : 8.0 * iters;
        printf("%s rep %d: %.3f ms  %.2f TFLOP/s\n", which == 0 ? "DMMA" : "DFMA", rep, ms, 2.0 * fma_per_thread * grid * 256 / (ms * 1e-3) * 1e-12);
      }
  return 0;
}
