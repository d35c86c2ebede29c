// Montgomery prime-field arithmetic on 32-bit limbs for sm_100a (BLS12-381 Fp: 12 limbs, Fr: 8).
//
// Replaces, on the device, the third-party big-integer arithmetic the reference reaches through
// kyber.Scalar / kyber.Point (reference call sites: algebra.go:92-105 Scalar.Mul/Add,
// algebra.go:348-359 Point.Mul/Add; modules named in go.mod:6-8).  Values are kept in Montgomery
// form x*R mod p with R = 2^(32N), always fully reduced to [0, p).
//
// The multiplier is a row-interleaved (CIOS) Montgomery product written as carry chains of
// mad.lo.cc / madc.hi.cc on *alternating* limbs: the partial products of the even limbs of the
// multiplicand land in one accumulator ("aligned"), those of the odd limbs in a second one that is
// shifted by one limb.  Inside one chain consecutive 64-bit products do not overlap, so a single
// carry flag ripples through it and ptxas can pair every lo/hi couple into one IMAD.WIDE.U32(.X).
#pragma once
#include <cstdint>
#include "constants.cuh"

namespace ps {

#ifdef __CUDACC__
#define PS_NOINLINE __host__ __device__ __noinline__
#define PS_DEV __host__ __device__ __forceinline__
#else
#define PS_NOINLINE
#define PS_DEV inline
#endif

// ---- carry-chain primitives -------------------------------------------------------------------
// Device: PTX add.cc / madc.* (the carry lives in the hardware flag between asm statements).
// Host (unit tests of the same templates, tests/host_check.cpp): a thread-local emulated flag.
#ifdef __CUDA_ARCH__
PS_DEV uint32_t ptx_add_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("add.cc.u32 %0,%1,%2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
PS_DEV uint32_t ptx_addc_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("addc.cc.u32 %0,%1,%2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
PS_DEV uint32_t ptx_addc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("addc.u32 %0,%1,%2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
PS_DEV uint32_t ptx_sub_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("sub.cc.u32 %0,%1,%2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
PS_DEV uint32_t ptx_subc_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("subc.cc.u32 %0,%1,%2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
PS_DEV uint32_t ptx_subc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("subc.u32 %0,%1,%2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
// (hi:lo) = a*b + (hi:lo) [+ carry]; carry out in the flag.  Written as a lo/hi pair so that ptxas
// fuses it into one IMAD.WIDE.U32(.X).
PS_DEV void ptx_mad_wide_cc(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) {
  asm volatile("mad.lo.cc.u32 %0,%2,%3,%0; madc.hi.cc.u32 %1,%2,%3,%1;" : "+r"(lo), "+r"(hi) : "r"(a), "r"(b)); }
PS_DEV void ptx_madc_wide_cc(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) {
  asm volatile("madc.lo.cc.u32 %0,%2,%3,%0; madc.hi.cc.u32 %1,%2,%3,%1;" : "+r"(lo), "+r"(hi) : "r"(a), "r"(b)); }
PS_DEV void ptx_mul_wide(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) {
  asm volatile("mul.lo.u32 %0,%2,%3; mul.hi.u32 %1,%2,%3;" : "=r"(lo), "=r"(hi) : "r"(a), "r"(b)); }
#else
inline uint32_t& host_cc() { static thread_local uint32_t cc = 0; return cc; }
inline uint32_t ptx_add_cc(uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a + b; host_cc() = (uint32_t)(t >> 32); return (uint32_t)t; }
inline uint32_t ptx_addc_cc(uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a + b + host_cc(); host_cc() = (uint32_t)(t >> 32); return (uint32_t)t; }
inline uint32_t ptx_addc(uint32_t a, uint32_t b) { return a + b + host_cc(); }
inline uint32_t ptx_sub_cc(uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a - b; host_cc() = (uint32_t)(t >> 63); return (uint32_t)t; }
inline uint32_t ptx_subc_cc(uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a - b - host_cc(); host_cc() = (uint32_t)(t >> 63); return (uint32_t)t; }
inline uint32_t ptx_subc(uint32_t a, uint32_t b) { return a - b - host_cc(); }
inline void ptx_mad_wide_cc(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) {
  unsigned __int128 t = (unsigned __int128)a * b + (((uint64_t)hi << 32) | lo);
  lo = (uint32_t)t; hi = (uint32_t)(t >> 32); host_cc() = (uint32_t)(t >> 64); }
inline void ptx_madc_wide_cc(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) {
  unsigned __int128 t = (unsigned __int128)a * b + (((uint64_t)hi << 32) | lo) + host_cc();
  lo = (uint32_t)t; hi = (uint32_t)(t >> 32); host_cc() = (uint32_t)(t >> 64); }
inline void ptx_mul_wide(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) {
  uint64_t t = (uint64_t)a * b; lo = (uint32_t)t; hi = (uint32_t)(t >> 32); }
#endif

// acc[0..N) = sum over k of a[2k] * b * 2^(64k)   (no carries: the products are disjoint)
template <int N>
PS_DEV void mul_chain(uint32_t* acc, const uint32_t* a, uint32_t b) {
#pragma unroll
  for (int j = 0; j < N; j += 2) ptx_mul_wide(acc[j], acc[j + 1], a[j], b);
}

// acc[0..N) += sum over k of a[2k] * b * 2^(64k); CARRY_IN consumes the pending carry flag at
// limb 0; the carry out of limb N-1 is left in the flag.
template <int N, bool CARRY_IN>
PS_DEV void mad_chain(uint32_t* acc, const uint32_t* a, uint32_t b) {
  if (CARRY_IN) ptx_madc_wide_cc(acc[0], acc[1], a[0], b);
  else ptx_mad_wide_cc(acc[0], acc[1], a[0], b);
#pragma unroll
  for (int j = 2; j < N; j += 2) ptx_madc_wide_cc(acc[j], acc[j + 1], a[j], b);
}

// acc[0..N) += sum over k of MOD[OFF+2k] * b * 2^(64k): the modulus limbs become immediates.
template <class P, int OFF>
PS_DEV void mad_chain_mod(uint32_t* acc, uint32_t b) {
  constexpr int N = P::N;
  ptx_mad_wide_cc(acc[0], acc[1], P::MOD(OFF), b);
#pragma unroll
  for (int j = 2; j < N; j += 2) ptx_madc_wide_cc(acc[j], acc[j + 1], P::MOD(OFF + j), b);
}

// ---- field element ----------------------------------------------------------------------------
template <class P> struct Fe;
template <class P> PS_NOINLINE Fe<P> fe_mul_call(const Fe<P>& a, const Fe<P>& b);

template <class P>
struct alignas(16) Fe {
  static constexpr int N = P::N;
  using Params = P;
  uint32_t v[P::N];

  PS_DEV static Fe zero() { Fe r;
#pragma unroll
    for (int i = 0; i < N; i++) r.v[i] = 0;
    return r; }
  PS_DEV static Fe one() { Fe r;
#pragma unroll
    for (int i = 0; i < N; i++) r.v[i] = P::ONE(i);
    return r; }
  template <uint32_t (*LIMB)(int)>
  PS_DEV static Fe from_const() { Fe r;
#pragma unroll
    for (int i = 0; i < N; i++) r.v[i] = LIMB(i);
    return r; }

  PS_DEV bool is_zero() const { uint32_t o = 0;
#pragma unroll
    for (int i = 0; i < N; i++) o |= v[i];
    return o == 0; }
  PS_DEV bool operator==(const Fe& b) const { uint32_t o = 0;
#pragma unroll
    for (int i = 0; i < N; i++) o |= v[i] ^ b.v[i];
    return o == 0; }
  PS_DEV bool operator!=(const Fe& b) const { return !(*this == b); }

  // r = a - p if a >= p (a < 2p on entry; `top` is the carry limb above v[N-1])
  PS_DEV static void final_sub(Fe& a, uint32_t top) {
    uint32_t t[N];
    t[0] = ptx_sub_cc(a.v[0], P::MOD(0));
#pragma unroll
    for (int i = 1; i < N; i++) t[i] = ptx_subc_cc(a.v[i], P::MOD(i));
    uint32_t borrow = ptx_subc(top, 0);  // 0 if a >= p, 0xffffffff otherwise
#pragma unroll
    for (int i = 0; i < N; i++) a.v[i] = borrow ? a.v[i] : t[i];
  }

  PS_DEV friend Fe operator+(const Fe& a, const Fe& b) {
    Fe r;
    r.v[0] = ptx_add_cc(a.v[0], b.v[0]);
#pragma unroll
    for (int i = 1; i < N; i++) r.v[i] = ptx_addc_cc(a.v[i], b.v[i]);
    uint32_t top = ptx_addc(0, 0);
    final_sub(r, top);
    return r;
  }
  PS_DEV friend Fe operator-(const Fe& a, const Fe& b) {
    Fe r;
    r.v[0] = ptx_sub_cc(a.v[0], b.v[0]);
#pragma unroll
    for (int i = 1; i < N; i++) r.v[i] = ptx_subc_cc(a.v[i], b.v[i]);
    uint32_t borrow = ptx_subc(0, 0);  // 0xffffffff when a < b
    uint32_t t[N];
#pragma unroll
    for (int i = 0; i < N; i++) t[i] = P::MOD(i) & borrow;
    r.v[0] = ptx_add_cc(r.v[0], t[0]);
#pragma unroll
    for (int i = 1; i < N - 1; i++) r.v[i] = ptx_addc_cc(r.v[i], t[i]);
    r.v[N - 1] = ptx_addc(r.v[N - 1], t[N - 1]);
    return r;
  }
  PS_DEV Fe neg() const { return is_zero() ? *this : (Fe::zero() - *this); }
  PS_DEV Fe dbl() const { return *this + *this; }

  // Montgomery product a*b/R mod p.
  PS_DEV friend Fe operator*(const Fe& a, const Fe& b) {
    // T = X + 2^32 * Y.  X receives the products of even limbs, Y those of odd limbs.
    uint32_t X[N], Y[N];
    uint32_t pend;  // limb to be added at weight 1 (spilled out of X by the previous row's shift)
    {
      const uint32_t bi = b.v[0];
      mul_chain<N>(X, a.v, bi);
      mul_chain<N>(Y, a.v + 1, bi);
      const uint32_t m = X[0] * P::INV;
      mad_chain_mod<P, 0>(X, m);
      uint32_t cx = ptx_addc(0, 0);
      mad_chain_mod<P, 1>(Y, m);
      // shift right one limb: X[0] == 0 now.
      pend = X[1];
#pragma unroll
      for (int k = 0; k < N - 2; k++) X[k] = X[k + 2];
      X[N - 2] = cx; X[N - 1] = 0;
      // roles swap: new aligned accumulator is Y, new shifted one is X
    }
#pragma unroll
    for (int i = 1; i < N; i++) {
      const uint32_t bi = b.v[i];
      // even rows of the loop (i odd) have Y aligned / X shifted, and vice versa.
      uint32_t* A = (i & 1) ? Y : X;  // aligned
      uint32_t* S = (i & 1) ? X : Y;  // shifted
      A[0] = ptx_add_cc(A[0], pend);                 // carry -> weight 2^32 == S[0]
      mad_chain<N, true>(S, a.v + 1, bi);            // odd limbs, consumes that carry
      mad_chain<N, false>(A, a.v, bi);               // even limbs
      uint32_t cx = ptx_addc(0, 0);
      const uint32_t m = A[0] * P::INV;
      mad_chain_mod<P, 0>(A, m);
      cx = ptx_addc(cx, 0);
      mad_chain_mod<P, 1>(S, m);
      pend = A[1];
#pragma unroll
      for (int k = 0; k < N - 2; k++) A[k] = A[k + 2];
      A[N - 2] = cx; A[N - 1] = 0;
    }
    // after N rows (N even) the aligned accumulator is X again
    uint32_t* A = (N & 1) ? Y : X;
    uint32_t* S = (N & 1) ? X : Y;
    Fe r;
    r.v[0] = ptx_add_cc(A[0], pend);
#pragma unroll
    for (int k = 1; k < N; k++) r.v[k] = ptx_addc_cc(A[k], S[k - 1]);
    uint32_t top = ptx_addc(S[N - 1], 0);
    final_sub(r, top);
    return r;
  }
  PS_DEV Fe sqr() const { return (*this) * (*this); }

  // (A dedicated squaring with the symmetric products taken once was measured slower inside the bucket-accumulation
  // kernel twice in round 2 and removed: interleaved with the reduction, 234 instead of 300 MACs, 82.4 vs 78.7 ms at
  // 2^24 G1 points; as an unreduced square + redc_wide, 222 MACs, 75.3 vs 74.2 ms (profiles/r02_ab_lazy.md).)

  PS_DEV Fe to_mont() const { return fe_mul_call(*this, from_const<P::R2>()); }
  PS_DEV Fe from_mont() const { Fe o = zero(); o.v[0] = 1; return fe_mul_call(*this, o); }

  // this^e, e = nlimbs little-endian 32-bit limbs (readable at run time: constant or global memory)
  PS_NOINLINE Fe pow(const uint32_t* e, int nlimbs) const {
    Fe acc = one();
    bool started = false;
    for (int i = nlimbs - 1; i >= 0; i--) {
      uint32_t w = e[i];
#pragma unroll 1
      for (int b = 31; b >= 0; b--) {
        if (started) acc = acc * acc;
        if ((w >> b) & 1) {
          if (started) acc = acc * (*this); else { acc = *this; started = true; }
        }
      }
    }
    return acc;
  }
};

// ---- unreduced products (sums of two products with ONE reduction: mul_sub_pair in curve.cuh) -------------------
// acc[0..N) += sum over k of MOD[OFF+2k] * b * 2^(64k), consuming the pending carry flag at limb 0
template <class P, int OFF>
PS_DEV void madc_chain_mod(uint32_t* acc, uint32_t b) {
  constexpr int N = P::N;
#pragma unroll
  for (int j = 0; j < N; j += 2) ptx_madc_wide_cc(acc[j], acc[j + 1], P::MOD(OFF + j), b);
}

// Unreduced products as plain integers of 2N limbs.  Two accumulators as in the Montgomery product: E takes the
// 64-bit partial products that start at even limb positions, O those that start at odd positions; every row is two
// independent carry chains of N/2 wide multiply-adds per product.  The carry out of a chain is added into the limb just
// above it, which no chain of that accumulator has covered yet at that row (it only ever holds such carries, so it
// cannot overflow).  That invariant is why a SUM of two products (DUAL: a b + c d < 2^(64N)) is accumulated row by row
// for both products together: adding the second product after the first would find every limb already full, and a
// carry added into a full limb is lost about once per 2^32 rows -- one wrong point per 2^24-point MSM, invisible to
// random operands (tests/test_host_arith.py drives it with saturated limbs).
template <class P, bool DUAL>
PS_DEV void wide_rows(uint32_t* E, uint32_t* O, const Fe<P>& a, const Fe<P>& b, const Fe<P>& c, const Fe<P>& d) {
  constexpr int N = P::N;
#pragma unroll
  for (int k = 0; k < 2 * N + 2; k++) { E[k] = 0; O[k] = 0; }
  mul_chain<N>(E, a.v, b.v[0]);
  mul_chain<N>(O + 1, a.v + 1, b.v[0]);
  if (DUAL) {
    mad_chain<N, false>(E, c.v, d.v[0]);
    E[N] = ptx_addc(E[N], 0);
    mad_chain<N, false>(O + 1, c.v + 1, d.v[0]);
    O[N + 1] = ptx_addc(O[N + 1], 0);
  }
#pragma unroll
  for (int i = 1; i < N; i++) {
    uint32_t* ge = (i & 1) ? O : E;   // grid of the products of the even limbs of a / c: they start at position i + 2k
    uint32_t* go = (i & 1) ? E : O;   // grid of the products of the odd limbs: position i + 1 + 2k
    mad_chain<N, false>(ge + i, a.v, b.v[i]);
    ge[i + N] = ptx_addc(ge[i + N], 0);
    mad_chain<N, false>(go + i + 1, a.v + 1, b.v[i]);
    go[i + 1 + N] = ptx_addc(go[i + 1 + N], 0);
    if (DUAL) {
      mad_chain<N, false>(ge + i, c.v, d.v[i]);
      ge[i + N] = ptx_addc(ge[i + N], 0);
      mad_chain<N, false>(go + i + 1, c.v + 1, d.v[i]);
      go[i + 1 + N] = ptx_addc(go[i + 1 + N], 0);
    }
  }
}
template <int N>
PS_DEV void wide_merge(uint32_t* T, const uint32_t* E, const uint32_t* O) {
  T[0] = E[0];
  T[1] = ptx_add_cc(E[1], O[1]);
#pragma unroll
  for (int k = 2; k < 2 * N - 1; k++) T[k] = ptx_addc_cc(E[k], O[k]);
  T[2 * N - 1] = ptx_addc(E[2 * N - 1], O[2 * N - 1]);
}
template <class P>
PS_DEV void mul_wide(uint32_t* T, const Fe<P>& a, const Fe<P>& b) {
  constexpr int N = P::N;
  uint32_t E[2 * N + 2], O[2 * N + 2];
  wide_rows<P, false>(E, O, a, b, a, b);
  wide_merge<N>(T, E, O);
}
// (a b + c d) / R mod p with ONE reduction; a b + c d < p R (e.g. all four below p, or p itself among them)
template <class P>
PS_DEV Fe<P> redc_wide(const uint32_t* T);
// T[0..2N) = a b + c d as a plain integer (must stay below 2^(64N))
template <class P>
PS_DEV void mul2_wide(uint32_t* T, const Fe<P>& a, const Fe<P>& b, const Fe<P>& c, const Fe<P>& d) {
  constexpr int N = P::N;
  uint32_t E[2 * N + 2], O[2 * N + 2];
  wide_rows<P, true>(E, O, a, b, c, d);
  wide_merge<N>(T, E, O);
}
template <class P>
PS_DEV Fe<P> mul2_lazy(const Fe<P>& a, const Fe<P>& b, const Fe<P>& c, const Fe<P>& d) {
  uint32_t T[2 * P::N];
  mul2_wide(T, a, b, c, d);
  return redc_wide<P>(T);
}
// p - a as a plain integer (a <= p): the additive inverse for use as an operand of an unreduced product (0 -> p)
template <class P>
PS_DEV Fe<P> neg_lazy(const Fe<P>& a) {
  constexpr int N = P::N;
  Fe<P> r;
  r.v[0] = ptx_sub_cc(P::MOD(0), a.v[0]);
#pragma unroll
  for (int k = 1; k < N - 1; k++) r.v[k] = ptx_subc_cc(P::MOD(k), a.v[k]);
  r.v[N - 1] = ptx_subc(P::MOD(N - 1), a.v[N - 1]);
  return r;
}

// T / R mod p for T < p R (2N limbs): Montgomery-reduce the low half with the row loop of the product (its
// multiplication rows left out), add the high half, one conditional subtraction.
template <class P>
PS_DEV Fe<P> redc_wide(const uint32_t* T) {
  constexpr int N = P::N;
  uint32_t X[N], Y[N];
  uint32_t pend;
#pragma unroll
  for (int k = 0; k < N; k++) { X[k] = T[k]; Y[k] = 0; }
  {
    const uint32_t m = X[0] * P::INV;
    mad_chain_mod<P, 0>(X, m);
    uint32_t cx = ptx_addc(0, 0);
    mad_chain_mod<P, 1>(Y, m);
    pend = X[1];
#pragma unroll
    for (int k = 0; k < N - 2; k++) X[k] = X[k + 2];
    X[N - 2] = cx; X[N - 1] = 0;
  }
#pragma unroll
  for (int i = 1; i < N; i++) {
    uint32_t* A = (i & 1) ? Y : X;  // aligned
    uint32_t* S = (i & 1) ? X : Y;  // shifted by one limb
    const uint32_t m = (A[0] + pend) * P::INV;
    A[0] = ptx_add_cc(A[0], pend);   // carry -> weight 2^32 == S[0]
    madc_chain_mod<P, 1>(S, m);      // consumes it
    mad_chain_mod<P, 0>(A, m);
    uint32_t cx = ptx_addc(0, 0);
    pend = A[1];
#pragma unroll
    for (int k = 0; k < N - 2; k++) A[k] = A[k + 2];
    A[N - 2] = cx; A[N - 1] = 0;
  }
  uint32_t* A = (N & 1) ? Y : X;
  uint32_t* S = (N & 1) ? X : Y;
  Fe<P> r;
  r.v[0] = ptx_add_cc(A[0], pend);
#pragma unroll
  for (int k = 1; k < N; k++) r.v[k] = ptx_addc_cc(A[k], S[k - 1]);
  uint32_t top = ptx_addc(S[N - 1], 0);   // zero: the reduced low half is at most p
  r.v[0] = ptx_add_cc(r.v[0], T[N]);
#pragma unroll
  for (int k = 1; k < N; k++) r.v[k] = ptx_addc_cc(r.v[k], T[N + k]);
  top = ptx_addc(top, 0);
  Fe<P>::final_sub(r, top);
  return r;
}

using Fp = Fe<FpParams>;
using Fr = Fe<FrParams>;

// explicit 128-bit loads of a field element (the compiler otherwise tends to issue one 32-bit load per
// limb when the value feeds multiplier operands directly, e.g. NTT twiddles)
template <class P>
PS_DEV Fe<P> fe_ld(const Fe<P>* p) {
#ifdef __CUDA_ARCH__
  Fe<P> r;
  const uint4* q = reinterpret_cast<const uint4*>(p);
#pragma unroll
  for (int i = 0; i < P::N / 4; i++) {
    uint4 t = __ldg(q + i);
    r.v[4 * i] = t.x; r.v[4 * i + 1] = t.y; r.v[4 * i + 2] = t.z; r.v[4 * i + 3] = t.w;
  }
  return r;
#else
  return *p;
#endif
}

// Out-of-line product: one copy of the ~300-instruction multiplier per field instead of one per call
// site.  Used wherever code size / compile time matters more than the call overhead (Fp2 towers,
// cold paths); the hot G1 kernels use the inlined operator*.
template <class P>
PS_NOINLINE Fe<P> fe_mul_call(const Fe<P>& a, const Fe<P>& b) { return a * b; }

}  // namespace ps

namespace ps {
#ifdef __CUDA_ARCH__
#define PS_CEXP(name) name
#else
// host copies of the constant-memory exponents
#define PS_CEXP(name) host_##name()
inline const uint32_t* host_c_FP_MOD_M2() { static uint32_t t[12]; for (int i = 0; i < 12; i++) t[i] = FpParams::MOD_M2(i); return t; }
inline const uint32_t* host_c_FP_SQRT_EXP() { static uint32_t t[12]; for (int i = 0; i < 12; i++) t[i] = FpParams::SQRT_EXP(i); return t; }
inline const uint32_t* host_c_FP_P_M3_D4() { static uint32_t t[12]; for (int i = 0; i < 12; i++) t[i] = FpParams::P_M3_D4(i); return t; }
inline const uint32_t* host_c_FP_P_M1_D2() { static uint32_t t[12]; for (int i = 0; i < 12; i++) t[i] = FpParams::P_M1_D2(i); return t; }
inline const uint32_t* host_c_FR_MOD_M2() { static uint32_t t[8]; for (int i = 0; i < 8; i++) t[i] = FrParams::MOD_M2(i); return t; }
#endif
// a^-1 (Montgomery form in and out; 0 -> 0) by the binary extended Euclidean algorithm on the raw limbs.
// Inversions sit at the very end of every call (one per result point, to-affine before encoding) where a
// single thread runs them: Fermat's a^(p-2) is a chain of ~570 dependent field products (0.6 ms for Fp on
// B200), the binary algorithm at most 2*bits rounds of shifts / subtractions on N limbs (~0.1 ms).
// With x = a R the loop yields x^-1 = a^-1 R^-1 as a plain integer; one product with R^3 restores the form.
template <class P>
PS_NOINLINE Fe<P> fe_inv_gcd(const Fe<P>& a) {
  constexpr int N = P::N;
  if (a.is_zero()) return a;
  uint32_t u[N], v[N], x1[N], x2[N];
#pragma unroll
  for (int i = 0; i < N; i++) { u[i] = a.v[i]; v[i] = P::MOD(i); x1[i] = 0; x2[i] = 0; }
  x1[0] = 1;
  // halve (y, x): y even; x <- x/2 mod p
  auto halve = [](uint32_t* y, uint32_t* x) {
    constexpr int N = P::N;
#pragma unroll
    for (int i = 0; i < N - 1; i++) y[i] = (y[i] >> 1) | (y[i + 1] << 31);
    y[N - 1] >>= 1;
    uint32_t top = 0;
    if (x[0] & 1) {
      x[0] = ptx_add_cc(x[0], P::MOD(0));
#pragma unroll
      for (int i = 1; i < N; i++) x[i] = ptx_addc_cc(x[i], P::MOD(i));
      top = ptx_addc(0, 0);
    }
#pragma unroll
    for (int i = 0; i < N - 1; i++) x[i] = (x[i] >> 1) | (x[i + 1] << 31);
    x[N - 1] = (x[N - 1] >> 1) | (top << 31);
  };
  // y -= z (y >= z); x <- x - w mod p
  auto reduce = [](uint32_t* y, const uint32_t* z, uint32_t* x, const uint32_t* w) {
    constexpr int N = P::N;
    y[0] = ptx_sub_cc(y[0], z[0]);
#pragma unroll
    for (int i = 1; i < N - 1; i++) y[i] = ptx_subc_cc(y[i], z[i]);
    y[N - 1] = ptx_subc(y[N - 1], z[N - 1]);
    x[0] = ptx_sub_cc(x[0], w[0]);
#pragma unroll
    for (int i = 1; i < N; i++) x[i] = ptx_subc_cc(x[i], w[i]);
    uint32_t borrow = ptx_subc(0, 0);
    if (borrow) {
      x[0] = ptx_add_cc(x[0], P::MOD(0));
#pragma unroll
      for (int i = 1; i < N - 1; i++) x[i] = ptx_addc_cc(x[i], P::MOD(i));
      x[N - 1] = ptx_addc(x[N - 1], P::MOD(N - 1));
    }
  };
  auto is_one = [](const uint32_t* y) {
    constexpr int N = P::N;
    uint32_t o = y[0] ^ 1u;
#pragma unroll
    for (int i = 1; i < N; i++) o |= y[i];
    return o == 0;
  };
  auto geq = [](const uint32_t* y, const uint32_t* z) {  // y >= z
    constexpr int N = P::N;
    (void)ptx_sub_cc(y[0], z[0]);
#pragma unroll
    for (int i = 1; i < N; i++) (void)ptx_subc_cc(y[i], z[i]);
    return ptx_subc(0, 0) == 0;
  };
#pragma unroll 1
  for (int guard = 0; guard < 4 * 32 * N; guard++) {  // at most 2 * bits rounds; the guard only bounds a corrupted input
    if (is_one(u) || is_one(v)) break;
#pragma unroll 1
    while ((u[0] & 1) == 0) halve(u, x1);
#pragma unroll 1
    while ((v[0] & 1) == 0) halve(v, x2);
    if (geq(u, v)) reduce(u, v, x1, x2); else reduce(v, u, x2, x1);
  }
  Fe<P> y;
  const bool from_u = is_one(u);
#pragma unroll
  for (int i = 0; i < N; i++) y.v[i] = from_u ? x1[i] : x2[i];
  const Fe<P> r2 = Fe<P>::template from_const<P::R2>();
  return y * (r2 * r2);  // y * R^3 / R = a^-1 R
}
// Fermat's a^(p-2): branch-free, the route for kernels that invert in every thread of a warp
PS_DEV Fp fp_inv(const Fp& a) { return a.pow(PS_CEXP(c_FP_MOD_M2), 12); }   // 0 -> 0
PS_DEV Fr fr_inv(const Fr& a) { return a.pow(PS_CEXP(c_FR_MOD_M2), 8); }
// binary algorithm: the route for the one-thread inversions at the end of a call (data-dependent branches)
PS_DEV Fp fp_inv_serial(const Fp& a) { return fe_inv_gcd(a); }
PS_DEV Fr fr_inv_serial(const Fr& a) { return fe_inv_gcd(a); }
PS_DEV Fp fp_sqrt_candidate(const Fp& a) { return a.pow(PS_CEXP(c_FP_SQRT_EXP), 12); }  // p = 3 mod 4
}  // namespace ps
