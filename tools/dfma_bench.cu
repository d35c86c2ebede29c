// FP64-pipe Montgomery product microbenchmark (development tool; results recorded in profiles/).
//
// Question: the prover's kernels keep the integer multiplier (fmaheavy) 82 % busy; the
// FP64 pipe is a separate unit.  Does an Fp product built from DFMA on 52-bit limbs held in doubles beat the IMAD one
// (3.0e10 products/s measured, tools/intpipe_bench.cu), alone or side by side with it?
//
// The limb product: for integers a, b < 2^52 held exactly in doubles,
//   hi = fma_rz(a, b, 2^104)            -> mantissa field of hi = floor(a b / 2^52)
//   lo = fma_rz(a, b, (2^104 + 2^52) - hi) -> mantissa field of lo = a b mod 2^52
// (2 DFMA + 1 DADD); the bit patterns are summed as 64-bit integers per column and the exponent fields are taken off
// once per column.  Fp = 8 limbs of 52 bits, R = 2^416 > 2^35 p, so products of values below 2p stay below 2p and no
// conditional subtraction is needed.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/_build/dfma_bench tools/dfma_bench.cu
// Host self-check of the arithmetic (no GPU): g++ -O2 -x c++ -DDFMA_HOST_ONLY -o /tmp/dfma_host tools/dfma_bench.cu && /tmp/dfma_host
#include <cfenv>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#ifndef DFMA_HOST_ONLY
#include <cuda_runtime.h>
#include "../playsnark_b200/csrc/field.cuh"
#define HD __host__ __device__ __forceinline__
#else
#define HD inline
#endif

HD double fma_rz(double a, double b, double c) {
#ifdef __CUDA_ARCH__
  return __fma_rz(a, b, c);
#else
  return std::fma(a, b, c);  // the host check runs under fesetround(FE_TOWARDZERO)
#endif
}
HD uint64_t d2u(double d) {
#ifdef __CUDA_ARCH__
  return (uint64_t)__double_as_longlong(d);
#else
  uint64_t u; memcpy(&u, &d, 8); return u;
#endif
}
HD double u2d(uint64_t u) {
#ifdef __CUDA_ARCH__
  return __longlong_as_double((long long)u);
#else
  double d; memcpy(&d, &u, 8); return d;
#endif
}

constexpr uint64_t MASK52 = (1ull << 52) - 1;
constexpr uint64_t OFF_LO = 0x4330000000000000ull;  // exponent field of 2^52  (lo = 2^52 + L)
constexpr uint64_t OFF_HI = 0x4670000000000000ull;  // exponent field of 2^104 (hi = 2^104 + H 2^52)
constexpr uint64_t PINV52 = 0x3fffcfffcfffdull;     // -p^-1 mod 2^52
HD double P52(int i) {
  constexpr double t[8] = {(double)0xeffffffffaaabull, (double)0xfeb153ffffb9full, (double)0x6b0f6241eabffull, (double)0x12bf6730d2a0full,
                           (double)0x764774b84f385ull, (double)0x1ba7b6434bacdull, (double)0x1ea397fe69a4bull, (double)0x1a011ull};
  return t[i];
}

struct F52 { double v[8]; };

// a b / 2^416 mod p, result below 2p when a b < 2^416 p (limbs < 2^52 in and out)
HD F52 mont_mul52(const F52& a, const F52& b) {
  const double C1 = 0x1p104, C2 = 0x1p104 + 0x1p52;
  uint64_t col[16];
#pragma unroll
  for (int k = 0; k < 16; k++) col[k] = 0;
#pragma unroll
  for (int i = 0; i < 8; i++)
#pragma unroll
    for (int j = 0; j < 8; j++) {
      double hi = fma_rz(a.v[i], b.v[j], C1);
      double lo = fma_rz(a.v[i], b.v[j], C2 - hi);
      col[i + j] += d2u(lo);
      col[i + j + 1] += d2u(hi);
    }
  // exponent fields of the multiplication phase: column k holds nlo(k) low halves and nhi(k) = nlo(k-1) high halves
#pragma unroll
  for (int k = 0; k < 16; k++) {
    const int nlo = k < 8 ? k + 1 : 15 - k, nhi = k == 0 ? 0 : (k - 1 < 8 ? k : 16 - k);
    col[k] -= (uint64_t)nlo * OFF_LO + (uint64_t)nhi * OFF_HI;
  }
#pragma unroll
  for (int k = 0; k < 8; k++) {
    // rows r < k left k low halves and k high halves here
    col[k] -= (uint64_t)k * (OFF_LO + OFF_HI);
    const uint64_t q = (col[k] * PINV52) & MASK52;
    const double qd = u2d(q | OFF_LO) - 0x1p52;
#pragma unroll
    for (int j = 0; j < 8; j++) {
      double hi = fma_rz(qd, P52(j), C1);
      double lo = fma_rz(qd, P52(j), C2 - hi);
      col[k + j] += d2u(lo);
      col[k + j + 1] += d2u(hi);
    }
    col[k + 1] += (col[k] - OFF_LO) >> 52;  // low 52 bits are zero now
  }
  F52 r;
  uint64_t carry = 0;
#pragma unroll
  for (int k = 8; k < 16; k++) {
    const int nlo = 15 - k, nhi = 16 - k;  // reduction rows that reached this column
    uint64_t t = col[k] - ((uint64_t)nlo * OFF_LO + (uint64_t)nhi * OFF_HI) + carry;
    carry = t >> 52;
    r.v[k - 8] = u2d((t & MASK52) | OFF_LO) - 0x1p52;
  }
  return r;
}

// ---- host check against plain big-integer arithmetic ------------------------------------------------------------
typedef unsigned __int128 u128;
struct Big { uint64_t w[16]; };  // little-endian 64-bit words, 1024 bits
static Big big_from52(const double* v, int n) {
  Big b{}; for (int i = 0; i < n; i++) { uint64_t x = (uint64_t)v[i]; int bit = 52 * i; b.w[bit / 64] |= x << (bit % 64); if (bit % 64 > 12) b.w[bit / 64 + 1] |= x >> (64 - bit % 64); }
  return b;
}
static Big big_mul(const Big& a, const Big& b) {  // low 1024 bits
  Big r{};
  for (int i = 0; i < 16; i++) { u128 c = 0; for (int j = 0; i + j < 16; j++) { c += (u128)a.w[i] * b.w[j] + r.w[i + j]; r.w[i + j] = (uint64_t)c; c >>= 64; } }
  return r;
}
static int big_cmp(const Big& a, const Big& b) { for (int i = 15; i >= 0; i--) if (a.w[i] != b.w[i]) return a.w[i] < b.w[i] ? -1 : 1; return 0; }
static Big big_sub(const Big& a, const Big& b) { Big r; u128 br = 0; for (int i = 0; i < 16; i++) { u128 t = (u128)a.w[i] - b.w[i] - br; r.w[i] = (uint64_t)t; br = (t >> 64) & 1; } return r; }
static Big big_shl1(const Big& a) { Big r; uint64_t c = 0; for (int i = 0; i < 16; i++) { r.w[i] = (a.w[i] << 1) | c; c = a.w[i] >> 63; } return r; }
static Big big_mod(Big a, const Big& p) {  // binary long division, a < 2^1023
  Big s = p; int sh = 0;
  while (big_cmp(big_shl1(s), a) <= 0 && sh < 700) { s = big_shl1(s); sh++; }
  for (; sh >= 0; sh--) { if (big_cmp(a, s) >= 0) a = big_sub(a, s); Big t{}; uint64_t c = 0; for (int i = 15; i >= 0; i--) { t.w[i] = (s.w[i] >> 1) | c; c = s.w[i] << 63; } s = t; }
  return a;
}
static bool host_check() {
  std::fesetround(FE_TOWARDZERO);
  double pv[8]; for (int i = 0; i < 8; i++) pv[i] = P52(i);
  const Big p = big_from52(pv, 8);
  Big R{}; R.w[6] = 1ull << 32;  // 2^416
  uint64_t s = 0x243f6a8885a308d3ull;
  auto rnd = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; };
  bool ok = true;
  for (int t = 0; t < 2000 && ok; t++) {
    F52 a, b;
    for (int i = 0; i < 8; i++) { a.v[i] = (double)(rnd() & MASK52); b.v[i] = (double)(rnd() & MASK52); }
    a.v[7] = (double)(rnd() % 0x34022ull); b.v[7] = (double)(rnd() % 0x34022ull);  // below 2p
    if (t == 0) for (int i = 0; i < 8; i++) { a.v[i] = (double)MASK52; b.v[i] = (double)MASK52; }  // limb extremes (a b < 2^416 p still holds)
    if (t == 0) { a.v[7] = b.v[7] = (double)0x3ffffull; }
    F52 r = mont_mul52(a, b);
    for (int i = 0; i < 8; i++) if (r.v[i] < 0 || r.v[i] > (double)MASK52) ok = false;
    // r R = a b (mod p) and r < 2p
    Big lhs = big_mod(big_mul(big_from52(r.v, 8), R), p);
    Big rhs = big_mod(big_mul(big_from52(a.v, 8), big_from52(b.v, 8)), p);
    if (big_cmp(lhs, rhs) != 0) ok = false;
    if (big_cmp(big_from52(r.v, 8), big_shl1(p)) >= 0) ok = false;
  }
  std::fesetround(FE_TONEAREST);
  return ok;
}

#ifdef DFMA_HOST_ONLY
int main() { bool ok = host_check(); printf("host check of mont_mul52 (2000 random products vs big-integer arithmetic): %s\n", ok ? "ok" : "FAILED"); return ok ? 0 : 1; }
#else
// ---- device kernels ----------------------------------------------------------------------------------------------
#define CH 4
// V = 0: DFMA, independent chains.  1: the limb-product primitive (DFMA, DADD, DFMA, two 64-bit adds).
// 2: DFMA chains and IMAD.WIDE carry chains interleaved in one thread.  3: DFMA chains + IADD3 carry chains.
template <int V>
__global__ void __launch_bounds__(256) k_rate(uint64_t* out, int iters, uint32_t seed) {
  double x[CH][6], a[6];
  uint64_t acc[CH][2];
  uint32_t y[CH][12], ia[12];
#pragma unroll
  for (int i = 0; i < 6; i++) a[i] = (double)((seed * (threadIdx.x + 7 + i)) & 0xfffff) + 3.0;
#pragma unroll
  for (int i = 0; i < 12; i++) ia[i] = seed * (threadIdx.x + 11 + i) | 1u;
#pragma unroll
  for (int c = 0; c < CH; c++) {
#pragma unroll
    for (int i = 0; i < 6; i++) x[c][i] = a[i] + c;
#pragma unroll
    for (int i = 0; i < 12; i++) y[c][i] = ia[i] ^ (0x9e3779b9u * (c + 1));
    acc[c][0] = acc[c][1] = c;
  }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int c = 0; c < CH; c++) {
      if (V == 0 || V == 2 || V == 3) {
#pragma unroll
        for (int i = 0; i < 6; i++) x[c][i] = __fma_rz(x[c][i], 0.999999, a[i]);
      }
      if (V == 1) {
#pragma unroll
        for (int i = 0; i < 6; i++) {
          double hi = __fma_rz(x[c][i], a[i], 0x1p104);
          double lo = __fma_rz(x[c][i], a[i], (0x1p104 + 0x1p52) - hi);
          acc[c][0] += d2u(lo); acc[c][1] += d2u(hi);
          x[c][i] = u2d((acc[c][0] & MASK52) | OFF_LO);  // keeps the operand < 2^53 and the chain dependent
        }
      }
      if (V == 2) {
        const uint32_t b = y[(c + 1) % CH][0] | 1u;
        asm volatile("mad.lo.cc.u32 %0,%2,%3,%0; madc.hi.cc.u32 %1,%2,%3,%1;" : "+r"(y[c][0]), "+r"(y[c][1]) : "r"(ia[0]), "r"(b));
#pragma unroll
        for (int i = 2; i < 12; i += 2)
          asm volatile("madc.lo.cc.u32 %0,%2,%3,%0; madc.hi.cc.u32 %1,%2,%3,%1;" : "+r"(y[c][i]), "+r"(y[c][i + 1]) : "r"(ia[i]), "r"(b));
      }
      if (V == 3) {
        asm volatile("add.cc.u32 %0,%0,%1;" : "+r"(y[c][0]) : "r"(ia[0]));
#pragma unroll
        for (int i = 1; i < 12; i++) asm volatile("addc.cc.u32 %0,%0,%1;" : "+r"(y[c][i]) : "r"(ia[i]));
      }
    }
  }
  uint64_t s = 0;
#pragma unroll
  for (int c = 0; c < CH; c++) {
#pragma unroll
    for (int i = 0; i < 6; i++) s ^= d2u(x[c][i]);
#pragma unroll
    for (int i = 0; i < 12; i++) s ^= y[c][i];
    s ^= acc[c][0] ^ acc[c][1];
  }
  if (s == 0x123456789abcdefull) out[0] = s;
}

// Fp products: MODE 0 = DFMA product in every warp, 1 = IMAD product (the library's) in every warp,
// 2 = even warps IMAD, odd warps DFMA (both pipes of an SM sub-partition busy at once).
// NCH independent dependent-chains per thread; out receives chain 0 of the first DFMA / IMAD thread for the check.
template <int MODE, int NCH>
__global__ void __launch_bounds__(256) k_prod(double* out52, uint32_t* out32, int iters, uint32_t seed) {
  const int warp = threadIdx.x >> 5;
  const bool use_dfma = MODE == 0 || (MODE == 2 && (warp & 1));
  if (use_dfma) {
    F52 x[NCH], a;
#pragma unroll
    for (int i = 0; i < 8; i++) a.v[i] = (double)(((uint64_t)seed * 0x9e3779b97f4a7c15ull >> (i + 3)) & MASK52);
    a.v[7] = (double)(seed & 0xffff);
#pragma unroll
    for (int c = 0; c < NCH; c++) {
      x[c] = a; x[c].v[0] = (double)(((uint64_t)(threadIdx.x + 1) * 7919u + c) & MASK52);
    }
    for (int it = 0; it < iters; it++) {
#pragma unroll
      for (int c = 0; c < NCH; c++) x[c] = mont_mul52(x[c], a);
    }
    double s = 0;
#pragma unroll
    for (int c = 1; c < NCH; c++) s += x[c].v[0];
    if (blockIdx.x == 0 && (threadIdx.x == 0 || (MODE == 2 && threadIdx.x == 32))) {
      for (int i = 0; i < 8; i++) out52[i] = x[0].v[i];
      out52[8] = s;
    }
    if (s == -1.0) out52[9] = s;
  } else {
    ps::Fp x[NCH], a;
#pragma unroll
    for (int i = 0; i < 12; i++) a.v[i] = seed * (i + 3) | 1u;
    a.v[11] &= 0x0fffffffu;
#pragma unroll
    for (int c = 0; c < NCH; c++) { x[c] = a; x[c].v[0] = (threadIdx.x + 1) * 7919u + c; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
      for (int c = 0; c < NCH; c++) x[c] = x[c] * a;
    }
    uint32_t s = 0;
#pragma unroll
    for (int c = 0; c < NCH; c++) s ^= x[c].v[0] ^ x[c].v[11];
    if (blockIdx.x == 0 && threadIdx.x == 0)
      for (int i = 0; i < 12; i++) out32[i] = x[0].v[i];
    if (s == 0x12345678u) out32[12] = s;
  }
}

template <int V>
void run_rate(const char* name, double inst_per_iter, int sms) {
  uint64_t* d; cudaMalloc(&d, 16);
  int blocks = sms * 8, threads = 256, iters = 4000;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 3; rep++) {
    cudaEventRecord(e0);
    k_rate<V><<<blocks, threads>>>(d, iters, 12345u + rep);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  double warp_inst = (double)blocks * threads / 32 * iters * inst_per_iter;
  printf("%-44s %.3e thread-inst/s  %.2f ms  => %.2f SM-cycles per warp-inst per SMSP at 1.965 GHz\n", name, warp_inst * 32 / (best * 1e-3), best,
         (best * 1e-3 * 1.965e9) / (warp_inst / (sms * 4)));
  cudaFree(d);
}

template <int MODE, int NCH>
double run_prod(const char* name, int sms, int blocks_per_sm, bool check) {
  double* d52; uint32_t* d32; cudaMalloc(&d52, 16 * 8); cudaMalloc(&d32, 16 * 4);
  int blocks = sms * blocks_per_sm, threads = 256, iters = 2000;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 3; rep++) {
    cudaEventRecord(e0);
    k_prod<MODE, NCH><<<blocks, threads>>>(d52, d32, iters, 12345u);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  cudaError_t err = cudaGetLastError();
  double prods = (double)blocks * threads * iters * NCH;
  double per_s = prods / (best * 1e-3);
  cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, k_prod<MODE, NCH>);
  printf("%-44s %.3e Fp products/s  %.2f ms  (%d regs, %zu B local, %d blocks/SM)%s\n", name, per_s, best, fa.numRegs, (size_t)fa.localSizeBytes, blocks_per_sm,
         err == cudaSuccess ? "" : cudaGetErrorString(err));
  if (check && MODE != 1) {
    // the same chain on the host, with the same code under round-toward-zero
    double got[16]; cudaMemcpy(got, d52, sizeof got, cudaMemcpyDeviceToHost);
    std::fesetround(FE_TOWARDZERO);
    F52 a, x;
    for (int i = 0; i < 8; i++) a.v[i] = (double)(((uint64_t)12345u * 0x9e3779b97f4a7c15ull >> (i + 3)) & MASK52);
    a.v[7] = (double)(12345u & 0xffff);
    x = a; x.v[0] = (double)(((uint64_t)((MODE == 2 ? 32 : 0) + 1) * 7919u + 0) & MASK52);
    for (int it = 0; it < iters; it++) x = mont_mul52(x, a);
    std::fesetround(FE_TONEAREST);
    bool same = true; for (int i = 0; i < 8; i++) same = same && got[i] == x.v[i];
    printf("    device chain of %d DFMA products equals the host's: %s\n", iters, same ? "yes" : "NO");
  }
  cudaFree(d52); cudaFree(d32);
  return per_s;
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int sms = p.multiProcessorCount;
  printf("device %s, %d SMs\n", p.name, sms);
  printf("host check of mont_mul52 vs big-integer arithmetic: %s\n", host_check() ? "ok" : "FAILED");
  run_rate<0>("dfma.rz, independent chains", CH * 6, sms);
  run_rate<1>("limb product: dfma, dadd, dfma (+2 add64)", CH * 6 * 3, sms);
  run_rate<2>("dfma + imad.wide.x chains, per dfma+wide pair", CH * 6, sms);
  run_rate<3>("dfma + 2 iadd3.x, per triple", CH * 6, sms);
  run_prod<1, 2>("Fp product, IMAD (library), 2 chains", sms, 4, false);
  run_prod<0, 1>("Fp product, DFMA 52-bit limbs, 1 chain", sms, 4, true);
  run_prod<0, 2>("Fp product, DFMA 52-bit limbs, 2 chains", sms, 4, true);
  run_prod<0, 2>("Fp product, DFMA 52-bit limbs, 2 chains", sms, 2, false);
  run_prod<2, 1>("even warps IMAD / odd warps DFMA, 1 chain", sms, 4, true);
  run_prod<2, 2>("even warps IMAD / odd warps DFMA, 2 chains", sms, 4, true);
  return 0;
}
#endif
