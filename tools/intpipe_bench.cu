// Standalone integer-pipe microbenchmarks (development tool; results recorded in profiles/).
// Measures warp-instruction issue cost on the fmaheavy pipe for the multiply flavours a Montgomery
// multiplier can be built from.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_build/intpipe_bench tools/intpipe_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CHAINS 4
template <int V>
__global__ void __launch_bounds__(256) k(uint32_t* out, int iters, uint32_t seed) {
  uint32_t a[12], x[CHAINS][12];
#pragma unroll
  for (int i = 0; i < 12; i++) a[i] = seed * (threadIdx.x + 7 + i) | 1u;
#pragma unroll
  for (int c = 0; c < CHAINS; c++)
#pragma unroll
    for (int i = 0; i < 12; i++) x[c][i] = a[i] ^ (0x9e3779b9u * (c + 1));
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int c = 0; c < CHAINS; c++) {
      uint32_t b = x[(c + 1) % CHAINS][0] | 1u;
      if (V == 0) {  // IMAD lo: 12 independent 32-bit mads
#pragma unroll
        for (int i = 0; i < 12; i++) asm volatile("mad.lo.u32 %0,%1,%2,%0;" : "+r"(x[c][i]) : "r"(a[i]), "r"(b));
      } else if (V == 1) {  // IMAD.HI
#pragma unroll
        for (int i = 0; i < 12; i++) asm volatile("mad.hi.u32 %0,%1,%2,%0;" : "+r"(x[c][i]) : "r"(a[i]), "r"(b));
      } else if (V == 2) {  // IMAD.WIDE, no carries: 6 per chain
#pragma unroll
        for (int i = 0; i < 12; i += 2) {
          uint64_t acc = ((uint64_t)x[c][i + 1] << 32) | x[c][i];
          asm volatile("mad.wide.u32 %0,%1,%2,%0;" : "+l"(acc) : "r"(a[i]), "r"(b));
          x[c][i] = (uint32_t)acc; x[c][i + 1] = (uint32_t)(acc >> 32);
        }
      } else if (V == 3) {  // the multiplier's carry chain: 6 x IMAD.WIDE.U32(.X)
        asm volatile("mad.lo.cc.u32 %0,%2,%3,%0; madc.hi.cc.u32 %1,%2,%3,%1;" : "+r"(x[c][0]), "+r"(x[c][1]) : "r"(a[0]), "r"(b));
#pragma unroll
        for (int i = 2; i < 12; i += 2)
          asm volatile("madc.lo.cc.u32 %0,%2,%3,%0; madc.hi.cc.u32 %1,%2,%3,%1;" : "+r"(x[c][i]), "+r"(x[c][i + 1]) : "r"(a[i]), "r"(b));
      } else if (V == 4) {  // carry-out only (each wide mad starts a fresh chain, result carry unused but generated)
#pragma unroll
        for (int i = 0; i < 12; i += 2) {
          uint32_t cy;
          asm volatile("mad.lo.cc.u32 %0,%3,%4,%0; madc.hi.cc.u32 %1,%3,%4,%1; addc.u32 %2,0,0;" : "+r"(x[c][i]), "+r"(x[c][i + 1]), "=r"(cy) : "r"(a[i]), "r"(b));
          x[c][(i + 2) % 12] ^= cy;
        }
      } else if (V == 5) {  // narrow carry chain: lo.cc on limbs (12 x IMAD with carry, no hi)
        asm volatile("mad.lo.cc.u32 %0,%1,%2,%0;" : "+r"(x[c][0]) : "r"(a[0]), "r"(b));
#pragma unroll
        for (int i = 1; i < 12; i++) asm volatile("madc.lo.cc.u32 %0,%1,%2,%0;" : "+r"(x[c][i]) : "r"(a[i]), "r"(b));
      } else if (V == 6) {  // narrow hi carry chain
        asm volatile("mad.hi.cc.u32 %0,%1,%2,%0;" : "+r"(x[c][0]) : "r"(a[0]), "r"(b));
#pragma unroll
        for (int i = 1; i < 12; i++) asm volatile("madc.hi.cc.u32 %0,%1,%2,%0;" : "+r"(x[c][i]) : "r"(a[i]), "r"(b));
      } else if (V == 7) {  // IADD3 carry chain on the ALU pipe for comparison
        asm volatile("add.cc.u32 %0,%0,%1;" : "+r"(x[c][0]) : "r"(b));
#pragma unroll
        for (int i = 1; i < 12; i++) asm volatile("addc.cc.u32 %0,%0,%1;" : "+r"(x[c][i]) : "r"(a[i]));
      }
    }
  }
  uint32_t s = 0;
#pragma unroll
  for (int c = 0; c < CHAINS; c++)
#pragma unroll
    for (int i = 0; i < 12; i++) s ^= x[c][i];
  if (s == 0x12345678u) out[0] = s;
}

template <int V>
void run(const char* name, double inst_per_iter, int sms) {
  uint32_t* d; cudaMalloc(&d, 16);
  int blocks = sms * 8, threads = 256, iters = 4000;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 3; rep++) {
    cudaEventRecord(e0);
    k<V><<<blocks, threads>>>(d, iters, 12345u + rep);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  double warp_inst = (double)blocks * threads / 32 * iters * inst_per_iter;
  double per_s = warp_inst * 32 / (best * 1e-3);
  printf("%-28s %.3e thread-inst/s  %.2f ms  => %.2f SM-cycles per warp-inst per SMSP at 1.965 GHz\n", name, per_s, best,
         (best * 1e-3 * 1.965e9) / (warp_inst / (sms * 4)));
  cudaFree(d);
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int sms = p.multiProcessorCount;
  printf("device %s, %d SMs\n", p.name, sms);
  run<0>("imad.lo", CHAINS * 12, sms);
  run<1>("imad.hi", CHAINS * 12, sms);
  run<2>("imad.wide", CHAINS * 6, sms);
  run<3>("imad.wide.x carry chain", CHAINS * 6, sms);
  run<4>("imad.wide carry-out only", CHAINS * 6, sms);
  run<5>("imad.lo.x carry chain", CHAINS * 12, sms);
  run<6>("imad.hi.x carry chain", CHAINS * 12, sms);
  run<7>("iadd3.x carry chain", CHAINS * 12, sms);
  return 0;
}
