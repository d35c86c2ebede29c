// Optimal-ate pairing product checks on BLS12-381 for the verifiers (Groth16Verify groth16.go:214-233,
// PHGR13Verify pinochio.go:281-375; Suite.Pair curve.go:36-38).
//
// The reference only ever compares GT values for equality, i.e. it decides  prod_i e(P_i, Q_i) == 1  (with one
// side's G1 points negated), which any non-degenerate bilinear map on the same groups decides identically.  So
// this file computes  (prod_i f_{|x|,Q_i}(P_i))^(3 (p^12 - 1) / r)  -- one shared final exponentiation per check,
// its hard part through the curve parameter x (3 (p^4 - p^2 + 1) / r = (x-1)^2 (x+p) (x^2+p^2-1) + 3) -- and
// compares it with 1.  Nothing here has to match a byte of the reference's GT representation.
//
// Tower: Fp2 = Fp[u]/(u^2+1), Fp6 = Fp2[v]/(v^3 - xi), Fp12 = Fp6[w]/(w^2 - v), xi = 1 + u.  One THREAD per Miller
// loop and per final exponentiation (constant-size work per proof: the parallelism is across pairs, checks and
// proofs of a batch); the functions are out of line so that the kernels stay a few thousand
// instructions instead of one inlined multiplier per product.
#pragma once
#include "curve.cuh"

namespace ps {

PS_DEV Fp2 fp2_mul_xi(const Fp2& a) { return Fp2{a.c0 - a.c1, a.c0 + a.c1}; }   // (a0 + a1 u)(1 + u)
PS_DEV Fp2 fp2_conj(const Fp2& a) { return Fp2{a.c0, a.c1.neg()}; }
PS_DEV Fp2 fp2_mul_fp(const Fp2& a, const Fp& k) { return Fp2{fe_mul_call(a.c0, k), fe_mul_call(a.c1, k)}; }
PS_NOINLINE Fp2 fp2_mul_call(const Fp2& a, const Fp2& b) { return a * b; }
PS_NOINLINE Fp2 fp2_sqr_call(const Fp2& a) { return a.sqr(); }

struct Fp6 {
  Fp2 c0, c1, c2;
  PS_DEV static Fp6 zero() { return Fp6{Fp2::zero(), Fp2::zero(), Fp2::zero()}; }
  PS_DEV static Fp6 one() { return Fp6{Fp2::one(), Fp2::zero(), Fp2::zero()}; }
  PS_DEV bool operator==(const Fp6& b) const { return c0 == b.c0 && c1 == b.c1 && c2 == b.c2; }
};
PS_DEV Fp6 fp6_add(const Fp6& a, const Fp6& b) { return Fp6{a.c0 + b.c0, a.c1 + b.c1, a.c2 + b.c2}; }
PS_DEV Fp6 fp6_sub(const Fp6& a, const Fp6& b) { return Fp6{a.c0 - b.c0, a.c1 - b.c1, a.c2 - b.c2}; }
PS_DEV Fp6 fp6_neg(const Fp6& a) { return Fp6{a.c0.neg(), a.c1.neg(), a.c2.neg()}; }
PS_DEV Fp6 fp6_mul_v(const Fp6& a) { return Fp6{fp2_mul_xi(a.c2), a.c0, a.c1}; }   // times v
// Karatsuba over Fp2: 6 products
PS_NOINLINE Fp6 fp6_mul(const Fp6& a, const Fp6& b) {
  const Fp2 t0 = fp2_mul_call(a.c0, b.c0), t1 = fp2_mul_call(a.c1, b.c1), t2 = fp2_mul_call(a.c2, b.c2);
  Fp6 r;
  r.c0 = t0 + fp2_mul_xi(fp2_mul_call(a.c1 + a.c2, b.c1 + b.c2) - (t1 + t2));
  r.c1 = fp2_mul_call(a.c0 + a.c1, b.c0 + b.c1) - (t0 + t1) + fp2_mul_xi(t2);
  r.c2 = fp2_mul_call(a.c0 + a.c2, b.c0 + b.c2) - (t0 + t2) + t1;
  return r;
}
PS_NOINLINE Fp6 fp6_inv(const Fp6& a) {
  const Fp2 c0 = fp2_sqr_call(a.c0) - fp2_mul_xi(fp2_mul_call(a.c1, a.c2));
  const Fp2 c1 = fp2_mul_xi(fp2_sqr_call(a.c2)) - fp2_mul_call(a.c0, a.c1);
  const Fp2 c2 = fp2_sqr_call(a.c1) - fp2_mul_call(a.c0, a.c2);
  const Fp2 t = fp2_mul_call(a.c0, c0) + fp2_mul_xi(fp2_mul_call(a.c2, c1) + fp2_mul_call(a.c1, c2));
  const Fp2 ti = fp2_inv_serial(t);
  return Fp6{fp2_mul_call(c0, ti), fp2_mul_call(c1, ti), fp2_mul_call(c2, ti)};
}

struct Fp12 {
  Fp6 c0, c1;
  PS_DEV static Fp12 one() { return Fp12{Fp6::one(), Fp6::zero()}; }
  PS_DEV bool operator==(const Fp12& b) const { return c0 == b.c0 && c1 == b.c1; }
};
PS_NOINLINE Fp12 fp12_mul(const Fp12& a, const Fp12& b) {
  const Fp6 t0 = fp6_mul(a.c0, b.c0), t1 = fp6_mul(a.c1, b.c1);
  Fp12 r;
  r.c1 = fp6_sub(fp6_mul(fp6_add(a.c0, a.c1), fp6_add(b.c0, b.c1)), fp6_add(t0, t1));
  r.c0 = fp6_add(t0, fp6_mul_v(t1));
  return r;
}
// (a0 + a1 w)^2 = (a0 + a1)(a0 + v a1) - t - v t + 2 t w,  t = a0 a1
PS_NOINLINE Fp12 fp12_sqr(const Fp12& a) {
  const Fp6 t = fp6_mul(a.c0, a.c1);
  Fp12 r;
  r.c0 = fp6_sub(fp6_sub(fp6_mul(fp6_add(a.c0, a.c1), fp6_add(a.c0, fp6_mul_v(a.c1))), t), fp6_mul_v(t));
  r.c1 = fp6_add(t, t);
  return r;
}
PS_DEV Fp12 fp12_conj(const Fp12& a) { return Fp12{a.c0, fp6_neg(a.c1)}; }
PS_NOINLINE Fp12 fp12_inv(const Fp12& a) {
  const Fp6 t = fp6_inv(fp6_sub(fp6_mul(a.c0, a.c0), fp6_mul_v(fp6_mul(a.c1, a.c1))));
  return Fp12{fp6_mul(a.c0, t), fp6_neg(fp6_mul(a.c1, t))};
}
// Frobenius x -> x^p.  As a polynomial in w (w^2 = v, w^6 = xi) the element is
//   c0.c0 + c1.c0 w + c0.c1 w^2 + c1.c1 w^3 + c0.c2 w^4 + c1.c2 w^5,   and (a w^k)^p = conj(a) gamma_k w^k
// with gamma_k = xi^(k (p-1) / 6) (FROBk_* of constants.cuh, derived by tools/gen_constants.py).
PS_NOINLINE Fp12 fp12_frob(const Fp12& a) {
  const Fp2 g1{Fp::from_const<FpParams::FROB1_C0>(), Fp::from_const<FpParams::FROB1_C1>()};
  const Fp2 g2{Fp::from_const<FpParams::FROB2_C0>(), Fp::from_const<FpParams::FROB2_C1>()};
  const Fp2 g3{Fp::from_const<FpParams::FROB3_C0>(), Fp::from_const<FpParams::FROB3_C1>()};
  const Fp2 g4{Fp::from_const<FpParams::FROB4_C0>(), Fp::from_const<FpParams::FROB4_C1>()};
  const Fp2 g5{Fp::from_const<FpParams::FROB5_C0>(), Fp::from_const<FpParams::FROB5_C1>()};
  Fp12 r;
  r.c0.c0 = fp2_conj(a.c0.c0);
  r.c1.c0 = fp2_mul_call(fp2_conj(a.c1.c0), g1);
  r.c0.c1 = fp2_mul_call(fp2_conj(a.c0.c1), g2);
  r.c1.c1 = fp2_mul_call(fp2_conj(a.c1.c1), g3);
  r.c0.c2 = fp2_mul_call(fp2_conj(a.c0.c2), g4);
  r.c1.c2 = fp2_mul_call(fp2_conj(a.c1.c2), g5);
  return r;
}

// Line through the (untwisted) G2 point T with twisted slope lam, evaluated at P = (xp, yp) and scaled by w^3
// (a factor of a proper subfield, which the final exponentiation removes):
//   (lam xt - yt) + (-lam xp) v + (yp) v w
PS_DEV Fp12 pairing_line(const Fp2& lam, const Fp2& xt, const Fp2& yt, const Fp& xp, const Fp& yp) {
  Fp12 l;
  l.c0 = Fp6{fp2_mul_call(lam, xt) - yt, fp2_mul_fp(lam.neg(), xp), Fp2::zero()};
  l.c1 = Fp6{Fp2::zero(), Fp2{yp, Fp::zero()}, Fp2::zero()};
  return l;
}

// f_{|x|, Q}(P), conjugated because x < 0.  Affine steps (one Fp2 inversion each, binary algorithm): the loop is a
// few hundred field products per bit either way, and the affine formulas are the ones the oracle restates.
PS_NOINLINE Fp12 miller_loop(const Affine<Fp>& P, const Affine<Fp2>& Q) {
  if (P.is_inf() || Q.is_inf()) return Fp12::one();
  Fp2 xt = Q.x, yt = Q.y;
  Fp12 f = Fp12::one();
  const uint64_t X = FpParams::BLS_X_ABS;
#pragma unroll 1
  for (int b = 62; b >= 0; b--) {
    const Fp2 xx = fp2_sqr_call(xt);
    Fp2 lam = fp2_mul_call(xx.dbl() + xx, fp2_inv_serial(yt.dbl()));
    f = fp12_mul(fp12_sqr(f), pairing_line(lam, xt, yt, P.x, P.y));
    Fp2 x3 = fp2_sqr_call(lam) - xt.dbl();
    yt = fp2_mul_call(lam, xt - x3) - yt;
    xt = x3;
    if ((X >> b) & 1) {
      lam = fp2_mul_call(Q.y - yt, fp2_inv_serial(Q.x - xt));
      f = fp12_mul(f, pairing_line(lam, xt, yt, P.x, P.y));
      x3 = fp2_sqr_call(lam) - xt - Q.x;
      yt = fp2_mul_call(lam, xt - x3) - yt;
      xt = x3;
    }
  }
  return fp12_conj(f);
}

// g^x for g in the cyclotomic subgroup (where the inverse is the conjugate): x = -|x|
PS_NOINLINE Fp12 fp12_pow_x(const Fp12& g) {
  const uint64_t X = FpParams::BLS_X_ABS;
  Fp12 r = g;
#pragma unroll 1
  for (int b = 62; b >= 0; b--) {
    r = fp12_sqr(r);
    if ((X >> b) & 1) r = fp12_mul(r, g);
  }
  return fp12_conj(r);
}

// f^(3 (p^12 - 1) / r)
PS_NOINLINE Fp12 final_exponentiation(const Fp12& f) {
  const Fp12 f1 = fp12_mul(fp12_conj(f), fp12_inv(f));            // f^(p^6 - 1)
  const Fp12 y = fp12_mul(fp12_frob(fp12_frob(f1)), f1);          // ^(p^2 + 1): now in the cyclotomic subgroup
  const Fp12 a = fp12_mul(fp12_pow_x(y), fp12_conj(y));           // y^(x - 1)
  const Fp12 b = fp12_mul(fp12_pow_x(a), fp12_conj(a));           // ^(x - 1)
  const Fp12 c = fp12_mul(fp12_pow_x(b), fp12_frob(b));           // ^(x + p)
  Fp12 d = fp12_mul(fp12_pow_x(fp12_pow_x(c)), fp12_frob(fp12_frob(c)));
  d = fp12_mul(d, fp12_conj(c));                                  // ^(x^2 + p^2 - 1)
  return fp12_mul(d, fp12_mul(fp12_sqr(y), y));                   // times y^3
}

// out[i] = f_{|x|, Q_i}(P_i)                                                          (thread per pair)
struct PairingMillerK {
  static constexpr int BLOCK = 32;
  PS_DEV static void run(uint32_t i, const Affine<Fp>* P, const Affine<Fp2>* Q, Fp12* out) { out[i] = miller_loop(P[i], Q[i]); }
};
// ok[t] = (prod of the Miller values [first[t], first[t+1])) ^ (3 (p^12-1)/r) == 1   (thread per check)
struct PairingCheckK {
  static constexpr int BLOCK = 32;
  PS_DEV static void run(uint32_t t, const Fp12* f, const uint32_t* first, uint8_t* ok) {
    Fp12 acc = Fp12::one();
    for (uint32_t i = first[t]; i < first[t + 1]; i++) acc = fp12_mul(acc, f[i]);
    ok[t] = final_exponentiation(acc) == Fp12::one() ? 1 : 0;
  }
};

}  // namespace ps
