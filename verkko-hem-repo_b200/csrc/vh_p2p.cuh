// Device side of the mailbox all-reduce over peer memory (layout and protocol: VhP2P in vh_internal.h).
// To be called by ONE full warp of one block per rank; returns the sum on every lane.
#ifndef VH_P2P_CUH
#define VH_P2P_CUH
#include "vh_internal.h"

__device__ __forceinline__ double vh_p2p_allreduce_warp(const VhP2P &P, unsigned long long seq, double v)
{
  const int lane = threadIdx.x & 31;
  const int slot = (int)(seq % VH_P2P_SLOTS);
  double    got  = 0.0;
  if (lane < P.n)
    {
      VhP2PCell *dst = P.peer[lane] + slot * P.n + P.me;
      asm volatile("st.relaxed.sys.global.f64 [%0], %1;" ::"l"(&dst->val), "d"(v) : "memory");
      asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(&dst->seq), "l"(seq) : "memory");
      const VhP2PCell   *src = P.peer[P.me] + slot * P.n + lane;
      unsigned long long s   = 0;
      const long long    t0  = clock64();
      do
        {
          asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(s) : "l"(&src->seq) : "memory");
          if (s != seq && clock64() - t0 > 20000000000ll)
            { // ~10 s: a peer never arrived; flag it and carry on so that the kernel terminates
              *P.err = 1;
              break;
            }
        }
      while (s != seq);
      asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(got) : "l"(&src->val) : "memory");
    }
  double sum = 0.0;
  for (int r = 0; r < P.n; ++r) // rank order: the same rounding on every rank
    sum += __shfl_sync(0xffffffffu, got, r);
  return sum;
}


#endif
