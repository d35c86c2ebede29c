// Host build of the device field/curve templates (carry flag emulated, see field.cuh) so that the
// arithmetic can be checked against the oracle on a machine without a GPU.  Test-only.
#include "../playsnark_b200/csrc/curve.cuh"
#include <cstring>
using namespace ps;

template <class F> static F load_std(const uint32_t* a) { F r; memcpy(r.v, a, sizeof(r.v)); return r.to_mont(); }
template <class F> static void store_std(uint32_t* o, const F& a) { F r = a.from_mont(); memcpy(o, r.v, sizeof(r.v)); }
static Fp2 load2(const uint32_t* a) { return Fp2{load_std<Fp>(a), load_std<Fp>(a + 12)}; }
static void store2(uint32_t* o, const Fp2& a) { store_std(o, a.c0); store_std(o + 12, a.c1); }
static G1Affine load_g1(const uint32_t* a) { return G1Affine{load_std<Fp>(a), load_std<Fp>(a + 12)}; }
static void store_g1(uint32_t* o, const G1Affine& p) { store_std(o, p.x); store_std(o + 12, p.y); }
static G2Affine load_g2(const uint32_t* a) { return G2Affine{load2(a), load2(a + 24)}; }
static void store_g2(uint32_t* o, const G2Affine& p) { store2(o, p.x); store2(o + 24, p.y); }

extern "C" {
void hc_fp_mul(const uint32_t* a, const uint32_t* b, uint32_t* o) { store_std(o, load_std<Fp>(a) * load_std<Fp>(b)); }
void hc_fp_add(const uint32_t* a, const uint32_t* b, uint32_t* o) { store_std(o, load_std<Fp>(a) + load_std<Fp>(b)); }
void hc_fp_sub(const uint32_t* a, const uint32_t* b, uint32_t* o) { store_std(o, load_std<Fp>(a) - load_std<Fp>(b)); }
void hc_fp_inv(const uint32_t* a, uint32_t* o) { store_std(o, fp_inv(load_std<Fp>(a))); }
void hc_fp_inv_serial(const uint32_t* a, uint32_t* o) { store_std(o, fp_inv_serial(load_std<Fp>(a))); }
void hc_fr_inv_serial(const uint32_t* a, uint32_t* o) { store_std(o, fr_inv_serial(load_std<Fr>(a))); }
void hc_g2_to_affine_serial(const uint32_t* p, int pre, uint32_t* o) {
  G2XYZZ a = G2XYZZ::from_affine(load_g2(p));
  for (int i = 0; i < pre; i++) a = xyzz_dbl(a);
  G2Affine s = xyzz_to_affine_serial(a), f = xyzz_to_affine(a);
  store_g2(o, s); store_g2(o + 48, f);
}
void hc_fp_sqrt(const uint32_t* a, uint32_t* o) { store_std(o, fp_sqrt_candidate(load_std<Fp>(a))); }
void hc_fr_mul(const uint32_t* a, const uint32_t* b, uint32_t* o) { store_std(o, load_std<Fr>(a) * load_std<Fr>(b)); }
void hc_fr_add(const uint32_t* a, const uint32_t* b, uint32_t* o) { store_std(o, load_std<Fr>(a) + load_std<Fr>(b)); }
void hc_fr_sub(const uint32_t* a, const uint32_t* b, uint32_t* o) { store_std(o, load_std<Fr>(a) - load_std<Fr>(b)); }
void hc_fr_inv(const uint32_t* a, uint32_t* o) { store_std(o, fr_inv(load_std<Fr>(a))); }
void hc_fp_sqr(const uint32_t* a, uint32_t* o) { store_std(o, load_std<Fp>(a).sqr()); }
void hc_fr_sqr(const uint32_t* a, uint32_t* o) { store_std(o, load_std<Fr>(a).sqr()); }
// raw Montgomery square on the given limbs (no domain conversion): a*a/R mod p, also for unreduced-looking inputs < p
void hc_fp_montsqr_raw(const uint32_t* a, uint32_t* o) { Fp x; memcpy(x.v, a, 48); Fp r = x.sqr(); memcpy(o, r.v, 48); }
void hc_fr_montsqr_raw(const uint32_t* a, uint32_t* o) { Fr x; memcpy(x.v, a, 32); Fr r = x.sqr(); memcpy(o, r.v, 32); }
// raw Montgomery product on the given limbs (no domain conversion): a*b/R mod p
void hc_fp_montmul_raw(const uint32_t* a, const uint32_t* b, uint32_t* o) { Fp x, y; memcpy(x.v, a, 48); memcpy(y.v, b, 48); Fp r = x * y; memcpy(o, r.v, 48); }
void hc_fr_montmul_raw(const uint32_t* a, const uint32_t* b, uint32_t* o) { Fr x, y; memcpy(x.v, a, 32); memcpy(y.v, b, 32); Fr r = x * y; memcpy(o, r.v, 32); }
void hc_fp2_mul(const uint32_t* a, const uint32_t* b, uint32_t* o) { store2(o, load2(a) * load2(b)); }
void hc_fp2_sqr(const uint32_t* a, uint32_t* o) { store2(o, load2(a).sqr()); }
// building blocks of the unreduced products on raw limbs
void hc_fp_mul2_lazy_raw(const uint32_t* a, const uint32_t* b, const uint32_t* c, const uint32_t* d, uint32_t* o) {
  Fp x, y, z, w; memcpy(x.v, a, 48); memcpy(y.v, b, 48); memcpy(z.v, c, 48); memcpy(w.v, d, 48);
  Fp r = mul2_lazy(x, y, neg_lazy(z), w); memcpy(o, r.v, 48); }
// a b + c d as a plain 768-bit integer (the unreduced sum inside mul2_lazy), operands = any 12-limb integers
void hc_fp_mul2_wide_raw(const uint32_t* a, const uint32_t* b, const uint32_t* c, const uint32_t* d, uint32_t* o) {
  Fp x, y, z, w; memcpy(x.v, a, 48); memcpy(y.v, b, 48); memcpy(z.v, c, 48); memcpy(w.v, d, 48);
  mul2_wide(o, x, y, z, w); }
void hc_fp_mul_wide_raw(const uint32_t* a, const uint32_t* b, uint32_t* o) { Fp x, y; memcpy(x.v, a, 48); memcpy(y.v, b, 48); mul_wide(o, x, y); }
void hc_fp_redc_wide_raw(const uint32_t* t, uint32_t* o) { Fp r = redc_wide<FpParams>(t); memcpy(o, r.v, 48); }
void hc_fp2_inv(const uint32_t* a, uint32_t* o) { store2(o, fp2_inv(load2(a))); }

// out = pre*P (pre in {1,2,3}: makes the accumulator non-trivially projective) + Q via madd
void hc_g1_madd(const uint32_t* p, const uint32_t* q, int pre, uint32_t* o) {
  G1Affine P = load_g1(p), Q = load_g1(q);
  G1XYZZ acc = G1XYZZ::from_affine(P);
  if (pre >= 2) acc = xyzz_dbl(acc);
  if (pre >= 3) xyzz_madd(acc, P);
  xyzz_madd(acc, Q);
  store_g1(o, xyzz_to_affine(acc));
}
void hc_g1_add(const uint32_t* p, const uint32_t* q, int pre, uint32_t* o) {
  G1Affine P = load_g1(p), Q = load_g1(q);
  G1XYZZ a = G1XYZZ::from_affine(P), b = G1XYZZ::from_affine(Q);
  if (pre >= 2) { a = xyzz_dbl(a); b = xyzz_dbl(b); }
  xyzz_add(a, b);
  store_g1(o, xyzz_to_affine(a));
}
void hc_g1_mul(const uint32_t* p, const uint32_t* k, uint32_t* o) {
  store_g1(o, xyzz_to_affine(xyzz_scalar_mul(G1XYZZ::from_affine(load_g1(p)), k, 8)));
}
void hc_g2_madd(const uint32_t* p, const uint32_t* q, int pre, uint32_t* o) {
  G2Affine P = load_g2(p), Q = load_g2(q);
  G2XYZZ acc = G2XYZZ::from_affine(P);
  if (pre >= 2) acc = xyzz_dbl(acc);
  if (pre >= 3) xyzz_madd(acc, P);
  xyzz_madd(acc, Q);
  store_g2(o, xyzz_to_affine(acc));
}
void hc_g2_add(const uint32_t* p, const uint32_t* q, int pre, uint32_t* o) {
  G2Affine P = load_g2(p), Q = load_g2(q);
  G2XYZZ a = G2XYZZ::from_affine(P), b = G2XYZZ::from_affine(Q);
  if (pre >= 2) { a = xyzz_dbl(a); b = xyzz_dbl(b); }
  xyzz_add(a, b);
  store_g2(o, xyzz_to_affine(a));
}
void hc_g2_mul(const uint32_t* p, const uint32_t* k, uint32_t* o) {
  store_g2(o, xyzz_to_affine(xyzz_scalar_mul(G2XYZZ::from_affine(load_g2(p)), k, 8)));
}
}
