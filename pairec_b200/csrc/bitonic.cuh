// bitonic.cuh — 1024-element bitonic sort with ONE element per thread (used by the recall's refine select).  The shared-memory version used by sort.cu and the recall's refine select pays a CTA barrier and
// a shared-memory round trip for each of its 55 compare-exchange stages (≈ 0.3 µs per stage); here the 40 stages whose
// partner is in the same warp (stride < 32) are register shuffles and only the 15 with stride >= 32 go through shared
// memory.  Same network, same strict total order, hence the same output.
#pragma once
#include <stdint.h>

namespace prg {

struct KeyIdx {
  uint64_t k;
  int32_t i;
};
__device__ __forceinline__ uint64_t bitonic_shfl(uint64_t v, uint32_t m) {
  return (uint64_t)__shfl_xor_sync(0xffffffffu, (unsigned long long)v, (int)m);
}
__device__ __forceinline__ KeyIdx bitonic_shfl(KeyIdx v, uint32_t m) {
  KeyIdx o;
  o.k = (uint64_t)__shfl_xor_sync(0xffffffffu, (unsigned long long)v.k, (int)m);
  o.i = __shfl_xor_sync(0xffffffffu, v.i, (int)m);
  return o;
}

// Called by all 1024 threads of the CTA; thread t passes element t and gets back the element of output position t.
// before(a, b): a precedes b in the output (strict total order; equal elements may come back in either slot).
// xch: shared memory, 1024 elements, free to overwrite; on return it may be reused at once.
template <typename E, typename Before>
__device__ __forceinline__ E bitonic_sort_1024(E e, E* xch, Before before) {
  const uint32_t t = threadIdx.x;
#pragma unroll 1
  for (uint32_t size = 2; size <= 1024u; size <<= 1) {
    const bool fwd = (t & size) == 0;   // this block of `size` elements is sorted in output order (else reversed)
#pragma unroll 1
    for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
      E o;
      if (stride >= 32u) {
        xch[t] = e;
        __syncthreads();
        o = xch[t ^ stride];
        __syncthreads();
      } else {
        o = bitonic_shfl(e, stride);
      }
      const bool lower = (t & stride) == 0;   // this thread is the lower position of the pair
      const bool keep_first = (lower == fwd);
      e = (keep_first == before(e, o)) ? e : o;
    }
  }
  return e;
}

}  // namespace prg
