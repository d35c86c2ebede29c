/* CPU oracle, C part -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C restatement of the arithmetic-heavy pieces of the reference's proving path, used (a) by
 * tests/ to check larger cases than the Python oracle finishes in seconds and (b) by bench.py's
 * cpu_baseline / --impl reference legs as the CPU "port" of the reference algorithm.  Nothing under
 * playsnark_b200/ links or loads it.
 *
 * PARITY UNPINNED (see oracle/ps_oracle.py): the reference's own arithmetic lives in un-vendored Go
 * modules (go.mod:6-8) and cannot be built here; this file is validated against the Python oracle
 * (tests/test_oracle_c.py), which in turn is anchored on public constants and the reference's
 * algebraic self-checks.
 *
 * What is restated, literally (same operation counts as the Go code):
 *   Poly.BlindEval   algebra.go:348-359   sum_i p[i]*P[i], ONE bit-serial double-and-add per term
 *   Poly.Mul         algebra.go:92-105    schoolbook, every product computed (also by zero)
 *   Poly.Sub / Add   algebra.go:161-197
 *   Poly.Div2        algebra.go:140-159   long division, one full tPoly.Mul(p2) per step
 *   computeAggregatePoly  qap.go:164-175
 * Field arithmetic: Fp as 6x64-bit Montgomery limbs (what kilic/bls12-381 uses), Fr as 4x64.
 * G1 in Jacobian coordinates, scalar multiplication LSB-first over the scalar's bits.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef unsigned __int128 u128;

/* ---------------- generic Montgomery arithmetic on N 64-bit limbs ---------------- */
#define DEFINE_FIELD(NAME, N)                                                                      \
  typedef struct { uint64_t v[N]; } NAME;                                                          \
  static const uint64_t NAME##_MOD[N];                                                             \
  static uint64_t NAME##_INV;                                                                      \
  static NAME NAME##_R2, NAME##_ONE;                                                               \
  static int NAME##_geq_mod(const uint64_t* a) {                                                   \
    for (int i = N - 1; i >= 0; i--) { if (a[i] > NAME##_MOD[i]) return 1; if (a[i] < NAME##_MOD[i]) return 0; } \
    return 1;                                                                                      \
  }                                                                                                \
  static void NAME##_add(NAME* r, const NAME* a, const NAME* b) {                                  \
    u128 c = 0; uint64_t t[N];                                                                     \
    for (int i = 0; i < N; i++) { c += (u128)a->v[i] + b->v[i]; t[i] = (uint64_t)c; c >>= 64; }    \
    if (c || NAME##_geq_mod(t)) { u128 br = 0; for (int i = 0; i < N; i++) { u128 d = (u128)t[i] - NAME##_MOD[i] - (uint64_t)br; t[i] = (uint64_t)d; br = (d >> 64) & 1; } } \
    memcpy(r->v, t, sizeof t);                                                                     \
  }                                                                                                \
  static void NAME##_sub(NAME* r, const NAME* a, const NAME* b) {                                  \
    u128 br = 0; uint64_t t[N];                                                                    \
    for (int i = 0; i < N; i++) { u128 d = (u128)a->v[i] - b->v[i] - (uint64_t)br; t[i] = (uint64_t)d; br = (d >> 64) & 1; } \
    if (br) { u128 c = 0; for (int i = 0; i < N; i++) { c += (u128)t[i] + NAME##_MOD[i]; t[i] = (uint64_t)c; c >>= 64; } } \
    memcpy(r->v, t, sizeof t);                                                                     \
  }                                                                                                \
  static void NAME##_mul(NAME* r, const NAME* a, const NAME* b) {                                  \
    uint64_t t[N + 2]; memset(t, 0, sizeof t);                                                     \
    for (int i = 0; i < N; i++) {                                                                  \
      u128 c = 0;                                                                                  \
      for (int j = 0; j < N; j++) { c += (u128)a->v[j] * b->v[i] + t[j]; t[j] = (uint64_t)c; c >>= 64; } \
      c += t[N]; t[N] = (uint64_t)c; t[N + 1] = (uint64_t)(c >> 64);                               \
      uint64_t m = t[0] * NAME##_INV;                                                              \
      c = (u128)m * NAME##_MOD[0] + t[0]; c >>= 64;                                                \
      for (int j = 1; j < N; j++) { c += (u128)m * NAME##_MOD[j] + t[j]; t[j - 1] = (uint64_t)c; c >>= 64; } \
      c += t[N]; t[N - 1] = (uint64_t)c; t[N] = t[N + 1] + (uint64_t)(c >> 64);                    \
    }                                                                                              \
    if (t[N] || NAME##_geq_mod(t)) { u128 br = 0; for (int i = 0; i < N; i++) { u128 d = (u128)t[i] - NAME##_MOD[i] - (uint64_t)br; t[i] = (uint64_t)d; br = (d >> 64) & 1; } } \
    memcpy(r->v, t, N * 8);                                                                        \
  }                                                                                                \
  static int NAME##_is_zero(const NAME* a) { uint64_t o = 0; for (int i = 0; i < N; i++) o |= a->v[i]; return o == 0; } \
  static int NAME##_eq(const NAME* a, const NAME* b) { return memcmp(a->v, b->v, N * 8) == 0; }    \
  static void NAME##_from_be(NAME* r, const uint8_t* b) { /* big-endian bytes -> Montgomery */      \
    NAME t;                                                                                        \
    for (int i = 0; i < N; i++) { uint64_t w = 0; for (int k = 0; k < 8; k++) w = (w << 8) | b[(N - 1 - i) * 8 + k]; t.v[i] = w; } \
    NAME##_mul(r, &t, &NAME##_R2);                                                                 \
  }                                                                                                \
  static void NAME##_to_be(uint8_t* b, const NAME* a) {                                            \
    NAME one_raw; memset(&one_raw, 0, sizeof one_raw); one_raw.v[0] = 1;                           \
    NAME t; NAME##_mul(&t, a, &one_raw);                                                           \
    for (int i = 0; i < N; i++) for (int k = 0; k < 8; k++) b[(N - 1 - i) * 8 + k] = (uint8_t)(t.v[i] >> (8 * (7 - k))); \
  }                                                                                                \
  static void NAME##_pow(NAME* r, const NAME* a, const uint64_t* e, int elimbs) {                  \
    NAME acc = NAME##_ONE;                                                                         \
    for (int i = elimbs - 1; i >= 0; i--) for (int b = 63; b >= 0; b--) {                          \
      NAME##_mul(&acc, &acc, &acc); if ((e[i] >> b) & 1) NAME##_mul(&acc, &acc, a); }              \
    *r = acc;                                                                                      \
  }                                                                                                \
  static void NAME##_inv(NAME* r, const NAME* a) {                                                 \
    uint64_t e[N]; memcpy(e, NAME##_MOD, sizeof e); e[0] -= 2; NAME##_pow(r, a, e, N);             \
  }                                                                                                \
  static void NAME##_init(void) {                                                                  \
    uint64_t inv = 1; for (int i = 0; i < 6; i++) inv *= 2 - NAME##_MOD[0] * inv;                  \
    NAME##_INV = (uint64_t)0 - inv;                                                                \
    /* R mod p by repeated doubling of 1, then R^2 by 64*N more doublings */                        \
    NAME x; memset(&x, 0, sizeof x); x.v[0] = 1;                                                   \
    for (int i = 0; i < 64 * N; i++) NAME##_add(&x, &x, &x);                                       \
    NAME##_ONE = x;                                                                                \
    for (int i = 0; i < 64 * N; i++) NAME##_add(&x, &x, &x);                                       \
    NAME##_R2 = x;                                                                                 \
  }

DEFINE_FIELD(fp, 6)
DEFINE_FIELD(fr, 4)

static const uint64_t fp_MOD[6] = {0xb9feffffffffaaabull, 0x1eabfffeb153ffffull, 0x6730d2a0f6b0f624ull,
                                   0x64774b84f38512bfull, 0x4b1ba7b6434bacd7ull, 0x1a0111ea397fe69aull};
static const uint64_t fr_MOD[4] = {0xffffffff00000001ull, 0x53bda402fffe5bfeull, 0x3339d80809a1d805ull, 0x73eda753299d7d48ull};

static int g_init = 0;
static void ensure_init(void) { if (!g_init) { fp_init(); fr_init(); g_init = 1; } }

/* ---------------- G1, Jacobian (X, Y, Z), Z == 0 <=> infinity ---------------- */
typedef struct { fp x, y, z; } g1j;

static void g1_set_inf(g1j* p) { memset(p, 0, sizeof *p); p->x = fp_ONE; p->y = fp_ONE; }
static int g1_is_inf(const g1j* p) { return fp_is_zero(&p->z); }

static void g1_dbl(g1j* r, const g1j* p) {
  if (g1_is_inf(p) || fp_is_zero(&p->y)) { g1_set_inf(r); return; }
  fp a, b, c, d, e, f, t, x3, y3, z3;
  fp_mul(&a, &p->x, &p->x); fp_mul(&b, &p->y, &p->y); fp_mul(&c, &b, &b);
  fp_add(&t, &p->x, &b); fp_mul(&t, &t, &t); fp_sub(&t, &t, &a); fp_sub(&t, &t, &c); fp_add(&d, &t, &t);
  fp_add(&e, &a, &a); fp_add(&e, &e, &a); fp_mul(&f, &e, &e);
  fp_sub(&x3, &f, &d); fp_sub(&x3, &x3, &d);
  fp_sub(&t, &d, &x3); fp_mul(&y3, &e, &t);
  fp_add(&c, &c, &c); fp_add(&c, &c, &c); fp_add(&c, &c, &c); fp_sub(&y3, &y3, &c);
  fp_mul(&z3, &p->y, &p->z); fp_add(&z3, &z3, &z3);
  r->x = x3; r->y = y3; r->z = z3;
}

static void g1_add(g1j* r, const g1j* p, const g1j* q) {
  if (g1_is_inf(p)) { *r = *q; return; }
  if (g1_is_inf(q)) { *r = *p; return; }
  fp z1z1, z2z2, u1, u2, s1, s2, h, rr, hh, hhh, v, t, x3, y3, z3;
  fp_mul(&z1z1, &p->z, &p->z); fp_mul(&z2z2, &q->z, &q->z);
  fp_mul(&u1, &p->x, &z2z2); fp_mul(&u2, &q->x, &z1z1);
  fp_mul(&t, &q->z, &z2z2); fp_mul(&s1, &p->y, &t);
  fp_mul(&t, &p->z, &z1z1); fp_mul(&s2, &q->y, &t);
  if (fp_eq(&u1, &u2)) { if (fp_eq(&s1, &s2)) { g1_dbl(r, p); return; } g1_set_inf(r); return; }
  fp_sub(&h, &u2, &u1); fp_sub(&rr, &s2, &s1);
  fp_mul(&hh, &h, &h); fp_mul(&hhh, &h, &hh); fp_mul(&v, &u1, &hh);
  fp_mul(&x3, &rr, &rr); fp_sub(&x3, &x3, &hhh); fp_sub(&x3, &x3, &v); fp_sub(&x3, &x3, &v);
  fp_sub(&t, &v, &x3); fp_mul(&y3, &rr, &t); fp_mul(&t, &s1, &hhh); fp_sub(&y3, &y3, &t);
  fp_mul(&z3, &p->z, &q->z); fp_mul(&z3, &z3, &h);
  r->x = x3; r->y = y3; r->z = z3;
}

/* kyber Point.Mul -> kilic MulScalar: walks the scalar's bits, doubling a running base */
static void g1_mul_scalar(g1j* r, const g1j* p, const uint8_t* k_be32) {
  g1j acc, base = *p;
  g1_set_inf(&acc);
  for (int byte = 31; byte >= 0; byte--)
    for (int bit = 0; bit < 8; bit++) {
      if ((k_be32[byte] >> bit) & 1) g1_add(&acc, &acc, &base);
      g1_dbl(&base, &base);
    }
  *r = acc;
}

static void g1_from_affine_be(g1j* p, const uint8_t* b96) {
  if (b96[0] & 0x40) { g1_set_inf(p); return; }
  fp_from_be(&p->x, b96); fp_from_be(&p->y, b96 + 48); p->z = fp_ONE;
}
static void g1_to_affine_be(uint8_t* b96, const g1j* p) {
  if (g1_is_inf(p)) { memset(b96, 0, 96); b96[0] = 0x40; return; }
  fp zi, zi2, zi3, x, y;
  fp_inv(&zi, &p->z); fp_mul(&zi2, &zi, &zi); fp_mul(&zi3, &zi2, &zi);
  fp_mul(&x, &p->x, &zi2); fp_mul(&y, &p->y, &zi3);
  fp_to_be(b96, &x); fp_to_be(b96 + 48, &y);
}

/* ---------------- exported entry points (ctypes) ---------------- */

/* Poly.BlindEval, algebra.go:348-359.  points: n x 96 B zcash-uncompressed; scalars: n x 32 B
 * big-endian; out: 96 B uncompressed.  Returns the number of group additions performed. */
long oc_blind_eval_g1(const uint8_t* points, const uint8_t* scalars, long n, uint8_t* out96) {
  ensure_init();
  g1j acc, tmp, base;
  long adds = 0;
  g1_set_inf(&acc);
  for (long i = 0; i < n; i++) {
    g1_from_affine_be(&base, points + 96 * i);
    g1_mul_scalar(&tmp, &base, scalars + 32 * i);
    g1_add(&acc, &acc, &tmp);
    adds++;
  }
  g1_to_affine_be(out96, &acc);
  return adds;
}

/* k * P for one point (used to cross-check the C group law against the Python oracle) */
void oc_g1_mul(const uint8_t* point96, const uint8_t* k_be32, uint8_t* out96) {
  ensure_init();
  g1j p, r;
  g1_from_affine_be(&p, point96);
  g1_mul_scalar(&r, &p, k_be32);
  g1_to_affine_be(out96, &r);
}

void oc_fp_mul(const uint8_t* a48, const uint8_t* b48, uint8_t* out48) {
  ensure_init();
  fp a, b, r; fp_from_be(&a, a48); fp_from_be(&b, b48); fp_mul(&r, &a, &b); fp_to_be(out48, &r);
}
void oc_fr_mul(const uint8_t* a32, const uint8_t* b32, uint8_t* out32) {
  ensure_init();
  fr a, b, r; fr_from_be(&a, a32); fr_from_be(&b, b32); fr_mul(&r, &a, &b); fr_to_be(out32, &r);
}
void oc_fr_inv(const uint8_t* a32, uint8_t* out32) {
  ensure_init();
  fr a, r; fr_from_be(&a, a32); fr_inv(&r, &a); fr_to_be(out32, &r);
}

/* ---- polynomials over Fr (coefficients low degree first, Montgomery form internally) ---- */
static fr* poly_load(const uint8_t* be, long n) {
  fr* p = (fr*)malloc((n ? n : 1) * sizeof(fr));
  for (long i = 0; i < n; i++) fr_from_be(&p[i], be + 32 * i);
  return p;
}
/* Poly.Mul, algebra.go:92-105: la*lb products, none skipped */
static fr* poly_mul(const fr* a, long la, const fr* b, long lb) {
  long l = la + lb - 1;
  fr* o = (fr*)calloc(l, sizeof(fr));
  fr t;
  for (long i = 0; i < la; i++)
    for (long j = 0; j < lb; j++) { fr_mul(&t, &a[i], &b[j]); fr_add(&o[i + j], &o[i + j], &t); }
  return o;
}

/* computeAggregatePoly, qap.go:164-175: out[k] = sum_i w[i] * M[i][k]; M is m x n row-major */
void oc_aggregate(const uint8_t* M_be, const uint8_t* w_be, long m, long n, uint8_t* out_be) {
  ensure_init();
  fr* acc = (fr*)calloc(n, sizeof(fr));
  fr c, w, t;
  for (long i = 0; i < m; i++) {
    fr_from_be(&w, w_be + 32 * i);
    for (long k = 0; k < n; k++) { fr_from_be(&c, M_be + 32 * (i * n + k)); fr_mul(&t, &c, &w); fr_add(&acc[k], &acc[k], &t); }
  }
  for (long k = 0; k < n; k++) fr_to_be(out_be + 32 * k, &acc[k]);
  free(acc);
}

/* QAP.Quotient's polynomial part, qap.go:155-160: h = (a*b - c) / z by Poly.Mul, Poly.Sub and
 * Poly.Div2.  a, b, c have n coefficients, z has n+1; h receives n-1.  `faithful` != 0 spends the
 * reference's full cost in Div2 (a complete schoolbook tPoly.Mul(p2) per step, algebra.go:156);
 * otherwise multiplications by the monomial's zero coefficients are skipped (same values).
 * Returns 0 when the remainder is zero, 1 otherwise ("apocalypse"). */
int oc_quotient(const uint8_t* a_be, const uint8_t* b_be, const uint8_t* c_be, const uint8_t* z_be, long n, int faithful,
                uint8_t* h_be) {
  ensure_init();
  fr *a = poly_load(a_be, n), *b = poly_load(b_be, n), *c = poly_load(c_be, n), *z = poly_load(z_be, n + 1);
  long lr = 2 * n - 1, lz = n + 1;
  fr* r = poly_mul(a, n, b, n);
  for (long i = 0; i < n; i++) fr_sub(&r[i], &r[i], &c[i]);
  long lq = lr - lz + 1;
  fr* q = (fr*)calloc(lq > 0 ? lq : 1, sizeof(fr));
  fr zlead_inv, t, prod;
  fr_inv(&zlead_inv, &z[lz - 1]);
  while (lr > 0 && lr >= lz) {
    fr_mul(&t, &r[lr - 1], &zlead_inv);
    long deg = lr - lz;
    fr_add(&q[deg], &q[deg], &t);
    if (faithful) {
      fr* tp = (fr*)calloc(deg + 1, sizeof(fr));
      tp[deg] = t;
      fr* pr = poly_mul(tp, deg + 1, z, lz);
      for (long i = 0; i < lr; i++) fr_sub(&r[i], &r[i], &pr[i]);
      free(tp); free(pr);
    } else {
      for (long j = 0; j < lz; j++) { fr_mul(&prod, &t, &z[j]); fr_sub(&r[deg + j], &r[deg + j], &prod); }
    }
    lr--;
  }
  int nonzero = 0;
  for (long i = 0; i < lr; i++) if (!fr_is_zero(&r[i])) nonzero = 1;
  for (long i = 0; i < lq; i++) fr_to_be(h_be + 32 * i, &q[i]);
  free(a); free(b); free(c); free(z); free(r); free(q);
  return nonzero;
}

/* sum_i a[i] * b[i] mod r over n pairs of 32-byte big-endian scalars (values may be >= r: they are
 * reduced).  Checker for the exponent-level MSM test at the benchmark's sizes
 * (groth16_test.go:32-107 style: the MSM over bases k_i*G must equal (sum k_i s_i)*G). */
void oc_fr_dot(const uint8_t* a_be, const uint8_t* b_be, long n, uint8_t* out32) {
  ensure_init();
  fr acc, x, y, t;
  memset(&acc, 0, sizeof acc);
  for (long i = 0; i < n; i++) {
    fr_from_be(&x, a_be + 32 * i);
    fr_from_be(&y, b_be + 32 * i);
    fr_mul(&t, &x, &y);
    fr_add(&acc, &acc, &t);
  }
  fr_to_be(out32, &acc);
}
