// TEST INFRASTRUCTURE: a lane-by-lane CUDA emulation for the CPU-only container.
//
// Runs a __global__ function body compiled by g++ with every CUDA thread of a block as a fibre (ucontext) on ONE OS
// thread: __syncthreads / __syncwarp / __shfl_*_sync are cooperative yields to a round-robin scheduler, dynamic shared
// memory is a per-block heap buffer.  It checks kernel LOGIC (indexing, staging through shared memory, warp-level
// exchange patterns, numerics); it says nothing about performance, memory-model races between unsynchronised threads
// (fibres only switch at barriers) or PTX-only features (TMA, mbarrier, cooperative groups are not emulated).
// Used by tests/native/points_emu_host.cc for csrc/vh_points_kernel.cuh.
#ifndef VH_CUDA_EMU_H
#define VH_CUDA_EMU_H

#include <ucontext.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <stdexcept>
#include <string>
#include <vector>

namespace emu
{
struct uint3_
{
  unsigned x = 0, y = 0, z = 0;
};
struct Fiber
{
  ucontext_t ctx;
  char      *stack = nullptr;
  uint3_     tid;
  int        state = 0; // 0 runnable, 1 waiting at the block barrier, 2 waiting at its warp barrier, 3 finished
};
struct Block
{
  std::vector<Fiber>    fibers;
  uint3_                bid, bdim, gdim;
  std::vector<char>     dyn_smem;
  double                shfl[32 * 64]; // [warp][lane] exchange slots (as 64-bit patterns)
  ucontext_t            sched;
  int                   cur = -1;
  std::function<void()> body;
};
constexpr size_t kGuard = 65536; // bytes, a multiple of 16: wide enough to absorb a whole mis-sized buffer
inline Block *&block()
{
  static Block *b = nullptr;
  return b;
}
inline Fiber &cur() { return block()->fibers[block()->cur]; }

inline void yield_with(int state)
{
  Block *b            = block();
  b->fibers[b->cur].state = state;
  swapcontext(&b->fibers[b->cur].ctx, &b->sched);
}
inline void trampoline()
{
  block()->body();
  cur().state = 3;
  swapcontext(&cur().ctx, &block()->sched);
}

// One thread block: run all fibres to completion, releasing barriers when every (live) participant has arrived.
inline void run_block(Block &B, unsigned n_threads, size_t stack_bytes)
{
  block() = &B;
  B.fibers.assign(n_threads, Fiber());
  for (unsigned t = 0; t < n_threads; ++t)
    {
      Fiber &f = B.fibers[t];
      f.stack  = static_cast<char *>(std::malloc(stack_bytes));
      f.tid.x  = t;
      getcontext(&f.ctx);
      f.ctx.uc_stack.ss_sp   = f.stack;
      f.ctx.uc_stack.ss_size = stack_bytes;
      f.ctx.uc_link          = &B.sched;
      makecontext(&f.ctx, reinterpret_cast<void (*)()>(trampoline), 0);
    }
  for (;;)
    {
      bool progressed = false, all_done = true;
      for (unsigned t = 0; t < n_threads; ++t)
        if (B.fibers[t].state == 0)
          {
            B.cur = (int)t;
            swapcontext(&B.sched, &B.fibers[t].ctx);
            progressed = true;
          }
      for (unsigned t = 0; t < n_threads; ++t)
        all_done = all_done && B.fibers[t].state == 3;
      if (all_done)
        break;
      // block barrier: everybody who is not finished waits at it
      bool at_block = true;
      for (unsigned t = 0; t < n_threads; ++t)
        if (B.fibers[t].state != 1 && B.fibers[t].state != 3)
          at_block = false;
      if (at_block)
        {
          for (unsigned t = 0; t < n_threads; ++t)
            if (B.fibers[t].state == 1)
              B.fibers[t].state = 0;
          continue;
        }
      // warp barriers: release every warp whose live lanes all wait at the warp barrier
      bool released = false;
      for (unsigned w = 0; w * 32 < n_threads; ++w)
        {
          bool ok = true, any = false;
          for (unsigned l = 0; l < 32 && w * 32 + l < n_threads; ++l)
            {
              const int s = B.fibers[w * 32 + l].state;
              if (s == 2)
                any = true;
              else if (s != 3)
                ok = false;
            }
          if (ok && any)
            {
              for (unsigned l = 0; l < 32 && w * 32 + l < n_threads; ++l)
                if (B.fibers[w * 32 + l].state == 2)
                  B.fibers[w * 32 + l].state = 0;
              released = true;
            }
        }
      if (!released && !progressed)
        throw std::runtime_error("cuda_emu: deadlock (threads wait at different barriers)");
    }
  for (Fiber &f : B.fibers)
    std::free(f.stack);
  block() = nullptr;
}

// launch: body is called once per thread with threadIdx/blockIdx set
inline void launch(unsigned grid, unsigned block_threads, size_t smem_bytes, const std::function<void()> &body,
                   size_t stack_bytes = 512 * 1024)
{
  for (unsigned b = 0; b < grid; ++b)
    {
      Block B;
      B.bid.x  = b;
      B.bdim.x = block_threads;
      B.gdim.x = grid;
      // dynamic shared memory with guard zones on both sides: a kernel that writes outside its launch size is caught
      B.dyn_smem.assign(smem_bytes + 2 * kGuard + 64, (char)0x5a);
      B.body = body;
      run_block(B, block_threads, stack_bytes);
      const char *base = reinterpret_cast<const char *>((reinterpret_cast<uintptr_t>(B.dyn_smem.data()) + 15) & ~uintptr_t(15));
      for (size_t i = 0; i < kGuard; ++i)
        if (base[i] != (char)0x5a || base[kGuard + smem_bytes + i] != (char)0x5a)
          throw std::runtime_error("cuda_emu: write outside the dynamic shared memory of the launch (block " + std::to_string(b) + ")");
    }
}
inline double *dyn_smem()
{
  char *p = block()->dyn_smem.data();
  return reinterpret_cast<double *>(((reinterpret_cast<uintptr_t>(p) + 15) & ~uintptr_t(15)) + kGuard);
}
template <class T>
inline T shfl_common(T v, int src_lane)
{
  static_assert(sizeof(T) <= 8, "64-bit shuffles at most");
  Block *b    = block();
  const int t = (int)cur().tid.x, w = t / 32, l = t % 32;
  std::memcpy(&b->shfl[w * 32 + l], &v, sizeof(T));
  yield_with(2);
  T out;
  std::memcpy(&out, &b->shfl[w * 32 + (src_lane & 31)], sizeof(T));
  yield_with(2); // nobody overwrites a slot before every lane has read
  return out;
}
} // namespace emu

// ---- the CUDA surface the emulated kernels use ----
#ifndef __CUDACC__
#undef __launch_bounds__
#define __launch_bounds__(...)
#ifndef __global__
#define __global__
#endif
#undef __shared__
#define __shared__ static /* one copy per kernel instantiation: blocks run one after the other, all fibres of a block share it */
#endif
#define threadIdx (emu::cur().tid)
#define blockIdx (emu::block()->bid)
#define blockDim (emu::block()->bdim)
#define gridDim (emu::block()->gdim)
#define VH_DYNAMIC_SMEM(name) double *name = emu::dyn_smem()
inline void __syncthreads() { emu::yield_with(1); }
inline void __syncwarp(unsigned = 0xffffffffu) { emu::yield_with(2); }
template <class T>
inline T __shfl_xor_sync(unsigned, T v, int lane_mask)
{
  return emu::shfl_common(v, (int)(emu::cur().tid.x % 32) ^ lane_mask);
}
template <class T>
inline T __shfl_sync(unsigned, T v, int src_lane)
{
  return emu::shfl_common(v, src_lane);
}
template <class T>
inline T __ldg(const T *p)
{
  return *p;
}
// warp votes / reductions over all 32 lanes (the emulated kernels use them with every lane of the warp alive)
inline unsigned __reduce_max_sync(unsigned, unsigned v)
{
  emu::Block *b = emu::block();
  const int   t = (int)emu::cur().tid.x, w = t / 32, l = t % 32;
  std::memcpy(&b->shfl[w * 32 + l], &v, sizeof(v));
  emu::yield_with(2);
  unsigned m = 0;
  for (int k = 0; k < 32; ++k)
    {
      unsigned o;
      std::memcpy(&o, &b->shfl[w * 32 + k], sizeof(o));
      m = o > m ? o : m;
    }
  emu::yield_with(2);
  return m;
}
inline unsigned __ballot_sync(unsigned, bool pred)
{
  emu::Block    *b = emu::block();
  const int      t = (int)emu::cur().tid.x, w = t / 32, l = t % 32;
  const unsigned v = pred ? 1u : 0u;
  std::memcpy(&b->shfl[w * 32 + l], &v, sizeof(v));
  emu::yield_with(2);
  unsigned m = 0;
  for (int k = 0; k < 32; ++k)
    {
      unsigned o;
      std::memcpy(&o, &b->shfl[w * 32 + k], sizeof(o));
      m |= o << k;
    }
  emu::yield_with(2);
  return m;
}
inline int      __ffs(unsigned v) { return __builtin_ffs((int)v); }
inline unsigned __float_as_uint(float f)
{
  unsigned u;
  std::memcpy(&u, &f, sizeof(u));
  return u;
}
inline float __double2float_ru(double x)
{
  float f = (float)x;
  if ((double)f < x)
    f = std::nextafterf(f, INFINITY);
  return f;
}
inline int atomicAdd(int *p, int v)
{ // fibres of one OS thread: no race
  const int old = *p;
  *p += v;
  return old;
}
#endif
