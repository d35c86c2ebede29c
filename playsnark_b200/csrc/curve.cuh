// BLS12-381 G1 / G2 group arithmetic for the prover's multi-scalar multiplications.
//
// Replaces kyber.Point.Add / Mul as the reference calls them inside Poly.BlindEval
// (algebra.go:348-359), sumBlind (groth16.go:134-141) and computeSolCommit (pinochio.go:222-229).
// Points are stored affine (bases) or in extended-Jacobian XYZZ form (accumulators):
//   x = X/ZZ, y = Y/ZZZ, ZZ^3 = ZZZ^2;  ZZ == 0 encodes the point at infinity.
// Affine infinity is encoded (0, 0), which is not on either curve (b != 0).
// The same templates serve G1 (F = Fp) and G2 (F = Fp2 = Fp[u]/(u^2+1)).
#pragma once
#include "field.cuh"

namespace ps {

// ---- Fp2 ---------------------------------------------------------------------------------------
// INLINE = false: base-field products are out-of-line calls (small code; used everywhere except the
// G2 bucket-accumulation hot loop).  INLINE = true: products are inlined (no argument traffic through
// the stack); same memory layout, so arrays of one kind can be viewed as the other.
template <bool INLINE>
struct alignas(16) Fp2T {
  Fp c0, c1;
  PS_DEV static Fp mulp(const Fp& a, const Fp& b) { if (INLINE) return a * b; else return fe_mul_call(a, b); }
  PS_DEV static Fp2T zero() { return Fp2T{Fp::zero(), Fp::zero()}; }
  PS_DEV static Fp2T one() { return Fp2T{Fp::one(), Fp::zero()}; }
  PS_DEV bool is_zero() const { return c0.is_zero() && c1.is_zero(); }
  PS_DEV bool operator==(const Fp2T& b) const { return c0 == b.c0 && c1 == b.c1; }
  PS_DEV bool operator!=(const Fp2T& b) const { return !(*this == b); }
  PS_DEV friend Fp2T operator+(const Fp2T& a, const Fp2T& b) { return Fp2T{a.c0 + b.c0, a.c1 + b.c1}; }
  PS_DEV friend Fp2T operator-(const Fp2T& a, const Fp2T& b) { return Fp2T{a.c0 - b.c0, a.c1 - b.c1}; }
  PS_DEV Fp2T neg() const { return Fp2T{c0.neg(), c1.neg()}; }
  PS_DEV Fp2T dbl() const { return Fp2T{c0.dbl(), c1.dbl()}; }
  // Karatsuba: 3 base-field products.  (Products on unreduced terms -- Karatsuba with 3 wide multiplications and 2
  // reductions, 720 instead of 864 multiply-adds, or schoolbook with 4 + 2 -- were measured SLOWER inside the G2 bucket
  // accumulation, 21.9 / 21.3 vs 19.8 ms at 2^20 points, and removed: that kernel sits at 255 registers and 2 warps
  // per scheduler and is bound by latency, not by the number of multiply-adds; profiles/r02_ab_lazy.md.)
  PS_DEV friend Fp2T operator*(const Fp2T& a, const Fp2T& b) {
    Fp t0 = mulp(a.c0, b.c0);
    Fp t1 = mulp(a.c1, b.c1);
    Fp t2 = mulp(a.c0 + a.c1, b.c0 + b.c1);
    return Fp2T{t0 - t1, t2 - t0 - t1};
  }
  // (c0 + c1 u)^2 = (c0+c1)(c0-c1) + 2 c0 c1 u: 2 base-field products
  PS_DEV Fp2T sqr() const {
    Fp s = c0 + c1, d = c0 - c1;
    Fp m = mulp(c0, c1);
    return Fp2T{mulp(s, d), m.dbl()};
  }
};
using Fp2 = Fp2T<false>;
using Fp2I = Fp2T<true>;

PS_DEV Fp2 fp2_inv(const Fp2& a) {
  Fp d = fp_inv(fe_mul_call(a.c0, a.c0) + fe_mul_call(a.c1, a.c1));
  return Fp2{fe_mul_call(a.c0, d), fe_mul_call(a.c1, d).neg()};
}

// the same with the binary-algorithm base-field inversion (one-thread tails, field.cuh)
PS_DEV Fp2 fp2_inv_serial(const Fp2& a) {
  Fp d = fp_inv_serial(fe_mul_call(a.c0, a.c0) + fe_mul_call(a.c1, a.c1));
  return Fp2{fe_mul_call(a.c0, d), fe_mul_call(a.c1, d).neg()};
}

template <class F> struct FieldInv;
template <> struct FieldInv<Fp> {
  PS_DEV static Fp inv(const Fp& a) { return fp_inv(a); }
  PS_DEV static Fp inv_serial(const Fp& a) { return fp_inv_serial(a); }
};
template <> struct FieldInv<Fp2> {
  PS_DEV static Fp2 inv(const Fp2& a) { return fp2_inv(a); }
  PS_DEV static Fp2 inv_serial(const Fp2& a) { return fp2_inv_serial(a); }
};
template <> struct FieldInv<Fp2I> {
  PS_DEV static Fp2I inv(const Fp2I& a) { Fp2 r = fp2_inv(Fp2{a.c0, a.c1}); return Fp2I{r.c0, r.c1}; }
  PS_DEV static Fp2I inv_serial(const Fp2I& a) { Fp2 r = fp2_inv_serial(Fp2{a.c0, a.c1}); return Fp2I{r.c0, r.c1}; }
};

// ---- points --------------------------------------------------------------------------------------
template <class F>
struct alignas(16) Affine {
  F x, y;
  PS_DEV bool is_inf() const { return x.is_zero() && y.is_zero(); }
  PS_DEV static Affine inf() { return Affine{F::zero(), F::zero()}; }
  PS_DEV Affine neg() const { return Affine{x, y.neg()}; }
};

template <class F>
struct alignas(16) XYZZ {
  F x, y, zz, zzz;
  PS_DEV bool is_inf() const { return zz.is_zero(); }
  PS_DEV static XYZZ inf() { return XYZZ{F::zero(), F::zero(), F::zero(), F::zero()}; }
  PS_DEV static XYZZ from_affine(const Affine<F>& p) {
    if (p.is_inf()) return inf();
    return XYZZ{p.x, p.y, F::one(), F::one()};
  }
  PS_DEV XYZZ neg() const { return XYZZ{x, y.neg(), zz, zzz}; }
};

// a b - c d, the shape of Y3 in every addition / doubling formula below.  The base field computes it as
// (a b + (p - c) d) / R with ONE Montgomery reduction (432 instead of 576 multiply-adds; same value bit for bit):
// measured on B200 inside the G1 bucket accumulation, 2^24 points: 78.98 -> 75.86 ms (tools/ab_fp2.py, round 2).
// -DPS_NO_LAZY_Y3 restores the two separate products (A/B builds, tools/ab_lazy.py).
PS_DEV Fp mul_sub_pair(const Fp& a, const Fp& b, const Fp& c, const Fp& d) {
#ifndef PS_NO_LAZY_Y3
  return mul2_lazy(a, b, neg_lazy(c), d);
#else
  return a * b - c * d;
#endif
}
// Fp2 keeps two plain products (the same merge over Fp2, three differences of products with one reduction each, measured
// 20.1 vs 19.7 ms in the G2 kernel).
template <bool I>
PS_DEV Fp2T<I> mul_sub_pair(const Fp2T<I>& a, const Fp2T<I>& b, const Fp2T<I>& c, const Fp2T<I>& d) { return a * b - c * d; }

// 2*P for affine P (mdbl-2008-s-1, a = 0)
template <class F>
PS_DEV XYZZ<F> xyzz_dbl_affine(const Affine<F>& p) {
  if (p.is_inf() || p.y.is_zero()) return XYZZ<F>::inf();
  F U = p.y.dbl();
  F V = U.sqr();
  F W = U * V;
  F S = p.x * V;
  F X2 = p.x.sqr();
  F M = X2.dbl() + X2;
  F X3 = M.sqr() - S.dbl();
  F Y3 = mul_sub_pair(M, S - X3, W, p.y);
  return XYZZ<F>{X3, Y3, V, W};
}

// 2*P (dbl-2008-s-1, a = 0)
template <class F>
PS_DEV XYZZ<F> xyzz_dbl(const XYZZ<F>& p) {
  if (p.is_inf() || p.y.is_zero()) return XYZZ<F>::inf();
  F U = p.y.dbl();
  F V = U.sqr();
  F W = U * V;
  F S = p.x * V;
  F X2 = p.x.sqr();
  F M = X2.dbl() + X2;
  F X3 = M.sqr() - S.dbl();
  F Y3 = mul_sub_pair(M, S - X3, W, p.y);
  return XYZZ<F>{X3, Y3, V * p.zz, W * p.zzz};
}

// out-of-line ("cold") versions, defined below: used on rare paths and in every kernel that is not
// the bucket-accumulation hot loop, to keep code size and compile time down
template <class F> PS_NOINLINE XYZZ<F> xyzz_dbl_c(const XYZZ<F>& p);
template <class F> PS_NOINLINE XYZZ<F> xyzz_dbl_affine_c(const Affine<F>& p);

// acc += q, q affine (madd-2008-s: 8M + 2S), all special cases handled
template <class F>
PS_DEV void xyzz_madd(XYZZ<F>& acc, const Affine<F>& q) {
  if (q.is_inf()) return;
  if (acc.is_inf()) { acc = XYZZ<F>{q.x, q.y, F::one(), F::one()}; return; }
  F U2 = q.x * acc.zz;
  F S2 = q.y * acc.zzz;
  F Pd = U2 - acc.x;
  F Rd = S2 - acc.y;
  if (Pd.is_zero()) {
    if (Rd.is_zero()) acc = xyzz_dbl_affine_c(q); else acc = XYZZ<F>::inf();
    return;
  }
  F PP = Pd.sqr();
  F PPP = Pd * PP;
  F Q = acc.x * PP;
  F X3 = Rd.sqr() - PPP - Q.dbl();
  F Y3 = mul_sub_pair(Rd, Q - X3, acc.y, PPP);
  acc.x = X3;
  acc.y = Y3;
  acc.zz = acc.zz * PP;
  acc.zzz = acc.zzz * PPP;
}

// acc += q, both XYZZ (add-2008-s: 12M + 2S)
template <class F>
PS_DEV void xyzz_add(XYZZ<F>& acc, const XYZZ<F>& q) {
  if (q.is_inf()) return;
  if (acc.is_inf()) { acc = q; return; }
  F U1 = acc.x * q.zz;
  F U2 = q.x * acc.zz;
  F S1 = acc.y * q.zzz;
  F S2 = q.y * acc.zzz;
  F Pd = U2 - U1;
  F Rd = S2 - S1;
  if (Pd.is_zero()) {
    if (Rd.is_zero()) acc = xyzz_dbl_c(acc); else acc = XYZZ<F>::inf();
    return;
  }
  F PP = Pd.sqr();
  F PPP = Pd * PP;
  F Q = U1 * PP;
  F X3 = Rd.sqr() - PPP - Q.dbl();
  F Y3 = mul_sub_pair(Rd, Q - X3, S1, PPP);
  acc.x = X3;
  acc.y = Y3;
  acc.zz = acc.zz * q.zz * PP;
  acc.zzz = acc.zzz * q.zzz * PPP;
}

template <class F> PS_NOINLINE XYZZ<F> xyzz_dbl_c(const XYZZ<F>& p) { return xyzz_dbl(p); }
template <class F> PS_NOINLINE XYZZ<F> xyzz_dbl_affine_c(const Affine<F>& p) { return xyzz_dbl_affine(p); }
template <class F> PS_NOINLINE void xyzz_add_c(XYZZ<F>& acc, const XYZZ<F>& q) { xyzz_add(acc, q); }
template <class F> PS_NOINLINE void xyzz_madd_c(XYZZ<F>& acc, const Affine<F>& q) { xyzz_madd(acc, q); }

template <class F>
PS_DEV Affine<F> xyzz_to_affine(const XYZZ<F>& p) {
  if (p.is_inf()) return Affine<F>::inf();
  // one inversion gives both: ZZ^3 = ZZZ^2  =>  1/ZZ = (ZZ/ZZZ)^2
  F zzz_inv = FieldInv<F>::inv(p.zzz);
  F t = zzz_inv * p.zz;
  F zz_inv = t.sqr();
  return Affine<F>{p.x * zz_inv, p.y * zzz_inv};
}

// k * P by left-to-right double-and-add over `nbits` bits of a little-endian limb array.
// Used for the handful of per-proof scalar multiplications (groth16.go:149,159,189-199) and as the
// plain reference inside self-tests; k = 0 or P = inf give inf like the reference's bit-serial loop.
template <class F>
PS_DEV XYZZ<F> xyzz_scalar_mul(const XYZZ<F>& p, const uint32_t* k, int nlimbs) {
  XYZZ<F> acc = XYZZ<F>::inf();
  for (int i = nlimbs - 1; i >= 0; i--) {
    uint32_t w = k[i];
#pragma unroll 1
    for (int b = 31; b >= 0; b--) {
      acc = xyzz_dbl_c(acc);
      if ((w >> b) & 1) xyzz_add_c(acc, p);
    }
  }
  return acc;
}

template <class F> PS_NOINLINE Affine<F> xyzz_to_affine_c(const XYZZ<F>& p) { return xyzz_to_affine(p); }
// for the handful of result points a call ends with (one thread each): binary-algorithm inversion
template <class F>
PS_NOINLINE Affine<F> xyzz_to_affine_serial(const XYZZ<F>& p) {
  if (p.is_inf()) return Affine<F>::inf();
  F zzz_inv = FieldInv<F>::inv_serial(p.zzz);
  F t = zzz_inv * p.zz;
  F zz_inv = t.sqr();
  return Affine<F>{p.x * zz_inv, p.y * zzz_inv};
}

using G1Affine = Affine<Fp>;
using G2Affine = Affine<Fp2>;
using G1XYZZ = XYZZ<Fp>;
using G2XYZZ = XYZZ<Fp2>;

}  // namespace ps
