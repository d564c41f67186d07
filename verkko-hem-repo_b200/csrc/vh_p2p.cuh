// Device side of the mailbox all-reduce over peer memory (layout and protocol: VhP2P in vh_internal.h).
#ifndef VH_P2P_CUH
#define VH_P2P_CUH
#include "vh_internal.h"

// A cell carries the value as two 8-byte words {32 bits of the value | 32-bit sequence number}.  Naturally aligned 8-byte
// stores and loads are single-copy atomic (also over NVLink), so a reader that sees the expected sequence number in both
// words has the whole value: no release/acquire fences, one volatile load pair per poll.
__device__ __forceinline__ void vh_p2p_post(VhP2PCell *dst, unsigned long long seq, double v)
{
  const unsigned long long bits = (unsigned long long)__double_as_longlong(v), tag = (seq & 0xffffffffull) << 32;
  asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(dst), "l"(tag | (bits & 0xffffffffull)), "l"(tag | (bits >> 32)) : "memory");
}
// Spins until the cell carries `seq`; ~10 s without it raises *err and returns 0 so that the kernel terminates.
__device__ __forceinline__ double vh_p2p_wait(const VhP2PCell *src, unsigned long long seq, int *err)
{
  const unsigned long long want = seq & 0xffffffffull;
  unsigned long long       lo = 0, hi = 0;
  const long long          t0 = clock64();
  for (;;)
    {
      asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(lo), "=l"(hi) : "l"(src) : "memory");
      if ((lo >> 32) == want && (hi >> 32) == want)
        break;
      if (clock64() - t0 > 20000000000ll)
        {
          *err = 1;
          return 0.0;
        }
    }
  return __longlong_as_double((long long)((lo & 0xffffffffull) | (hi << 32)));
}

// All-reduce (sum) of one double; to be called by ONE full warp of one block per rank; returns the sum on every lane.
__device__ __forceinline__ double vh_p2p_allreduce_warp(const VhP2P &P, unsigned long long seq, double v)
{
  const int lane = threadIdx.x & 31;
  const int slot = (int)(seq % VH_P2P_SLOTS);
  double    got  = 0.0;
  if (lane < P.n)
    {
      vh_p2p_post(P.peer[lane] + slot * P.n + P.me, seq, v);
      got = vh_p2p_wait(P.peer[P.me] + slot * P.n + lane, seq, P.err);
    }
  double sum = 0.0;
  for (int r = 0; r < P.n; ++r) // rank order: the same rounding on every rank
    sum += __shfl_sync(0xffffffffu, got, r);
  return sum;
}

#endif
