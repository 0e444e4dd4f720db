// Device helpers shared by the batched CG kernels (kernels_cg.cu, kernels_cg_stencil.cu).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#ifndef MSGPU_FAST_SCALARS
#define MSGPU_FAST_SCALARS 1
#endif

namespace msgpu {
namespace cgc {

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Dot-product partial of a strip: NACC independent accumulators (fixed assignment j % NACC, fixed
// combination order) so that a long strip is not one serial chain of K dependent DFMAs.
template <int K, int NACC, class F>
__device__ __forceinline__ double strip_dot(F&& term) {
  double acc[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) acc[i] = 0.0;
#pragma unroll
  for (int j = 0; j < K; ++j) acc[j % NACC] += term(j);
#pragma unroll
  for (int w2 = 1; w2 < NACC; w2 *= 2)
#pragma unroll
    for (int i = 0; i + w2 < NACC; i += 2 * w2) acc[i] += acc[i + w2];
  return acc[0];
}

// Bulk prefetch of `bytes` (a multiple of 16) starting at the 16-byte block that contains p into L2.
__device__ __forceinline__ void l2_prefetch_bulk(const void* p, uint32_t bytes) {
  const unsigned long long addr = reinterpret_cast<unsigned long long>(p) & ~15ull;
  if (bytes) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;\n" ::"l"(addr), "r"(bytes) : "memory");
}

}  // namespace cgc
}  // namespace msgpu
