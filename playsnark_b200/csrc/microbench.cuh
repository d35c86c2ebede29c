// Integer-multiply pipe microbenchmarks: the roofline denominator for the MSM / NTT kernels.
// MEASURED_PEAKS.json carries no integer figure, so bench.py measures it on the device it runs on.
//   variant 0: IMAD      (mad.lo.u32, 32x32+32 -> 32)
//   variant 1: IMAD.HI   (mad.hi.u32)
//   variant 2: IMAD.WIDE (mad.wide.u32, 32x32+64 -> 64), independent accumulators
//   variant 3: IMAD.WIDE.X carry chains exactly as the field multiplier issues them
// Each thread keeps 8 independent dependency chains so the pipe, not latency, is the limit.
#pragma once
#include "backend.cuh"
#include "curve.cuh"

#if PS_GPU
namespace ps {

template <int VARIANT>
__global__ void __launch_bounds__(256) k_intpipe(uint32_t* out, int iters, uint32_t seed) {
  uint32_t a = seed * (threadIdx.x + 1) | 1u, b = (seed ^ 0x9e3779b9u) + blockIdx.x;
  uint32_t x[16];
#pragma unroll
  for (int i = 0; i < 16; i++) x[i] = a + i * 0x01000193u;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int rep = 0; rep < 8; rep++) {
      if (VARIANT == 0) {
#pragma unroll
        for (int i = 0; i < 16; i++) asm volatile("mad.lo.u32 %0,%1,%2,%0;" : "+r"(x[i]) : "r"(a), "r"(b));
      } else if (VARIANT == 1) {
#pragma unroll
        for (int i = 0; i < 16; i++) asm volatile("mad.hi.u32 %0,%1,%2,%0;" : "+r"(x[i]) : "r"(a), "r"(b));
      } else if (VARIANT == 2) {
#pragma unroll
        for (int i = 0; i < 16; i += 2) {
          uint64_t acc = ((uint64_t)x[i + 1] << 32) | x[i];
          asm volatile("mad.wide.u32 %0,%1,%2,%0;" : "+l"(acc) : "r"(a), "r"(b));
          x[i] = (uint32_t)acc; x[i + 1] = (uint32_t)(acc >> 32);
        }
      } else {
        // one 16-limb carry chain = 8 IMAD.WIDE.U32.X
        mad_chain<16, false>(x, x, b);
      }
    }
  }
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < 16; i++) s ^= x[i];
  if (s == 0x12345678u) out[0] = s;  // never true in practice: keeps the work alive
}

// instructions issued per thread per iteration
inline double intpipe_inst_per_iter(int variant) { return variant <= 1 ? 8.0 * 16 : 8.0 * 8; }

template <class F>
__global__ void __launch_bounds__(256) k_fieldmul(F* io, int iters) {
  uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  F x = io[tid], y = io[tid + gridDim.x * blockDim.x];
  for (int it = 0; it < iters; it++) {
    x = x * y;
    y = y * x;
  }
  io[tid] = x + y;
}

}  // namespace ps
#endif
