/* CPU oracle, C part 2 -- TEST INFRASTRUCTURE, NOT PRODUCT CODE (same rules as ps_oracle.c, which
 * this file includes: nothing under playsnark_b200/ links or loads it).
 *
 * Plain-C restatement of the reference's WHOLE proving flow, with the reference's own algorithms and
 * operation counts, so that bench.py can time "the reference's CPU prover" beside the GPU one at the
 * sizes where the naive algorithms finish, and so that the C
 * port is checked end to end against the Python oracle (tests/test_oracle_c.py):
 *
 *   Interpolate / lagrangeBasis   algebra.go:254-280, 313-338   (one inversion per (j, m) pair)
 *   ToQAP / qapInterpolate        qap.go:35-93
 *   Poly.Eval                     algebra.go:107-115
 *   GeneratePowersCommit          algebra.go:371-384
 *   NewGroth16TrustedSetup        groth16.go:64-101, linearPolyForVar / fullLinearPoly :238-264
 *   Groth16Prove                  groth16.go:122-211  (sumBlind: m*n bit-serial scalar-muls, three times)
 *   NewPHGR13TrustedSetup (EK)    pinochio.go:93-141, generateEvalCommit :381-388
 *   PHGR13Prove                   pinochio.go:207-254
 *
 * Scalar multiplication follows kyber.Point.Mul -> kilic MulScalar as recalled in SURVEY 8 c2: LSB-first
 * over the scalar's BitLen with a running doubled base.  G2 is the same Jacobian code over Fp2 =
 * Fp[u]/(u^2+1).  The reference is single-threaded; `threads` > 1 spreads the independent iterations
 * of its outer loops (variables in sumBlind / ToQAP / the setups) over OpenMP threads for the
 * "all host threads" leg of bench.py -- same group elements, since point addition is associative.
 * PARITY UNPINNED, like the rest of the oracle (header of ps_oracle.py).
 */
#include "ps_oracle.c"

#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static double now_s(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

/* ---------------- Fp2 ---------------- */
typedef struct { fp c0, c1; } fp2;
static fp2 fp2_ONE;
static void fp2_add(fp2* r, const fp2* a, const fp2* b) { fp_add(&r->c0, &a->c0, &b->c0); fp_add(&r->c1, &a->c1, &b->c1); }
static void fp2_sub(fp2* r, const fp2* a, const fp2* b) { fp_sub(&r->c0, &a->c0, &b->c0); fp_sub(&r->c1, &a->c1, &b->c1); }
static void fp2_mul(fp2* r, const fp2* a, const fp2* b) {
  fp t0, t1, t2, s0, s1;
  fp_mul(&t0, &a->c0, &b->c0); fp_mul(&t1, &a->c1, &b->c1);
  fp_add(&s0, &a->c0, &a->c1); fp_add(&s1, &b->c0, &b->c1); fp_mul(&t2, &s0, &s1);
  fp_sub(&r->c0, &t0, &t1); fp_sub(&t2, &t2, &t0); fp_sub(&r->c1, &t2, &t1);
}
static int fp2_is_zero(const fp2* a) { return fp_is_zero(&a->c0) && fp_is_zero(&a->c1); }
static int fp2_eq(const fp2* a, const fp2* b) { return fp_eq(&a->c0, &b->c0) && fp_eq(&a->c1, &b->c1); }
static void fp2_inv(fp2* r, const fp2* a) {
  fp t0, t1, d;
  fp_mul(&t0, &a->c0, &a->c0); fp_mul(&t1, &a->c1, &a->c1); fp_add(&t0, &t0, &t1); fp_inv(&d, &t0);
  fp_mul(&r->c0, &a->c0, &d); fp_mul(&t1, &a->c1, &d);
  fp z; memset(&z, 0, sizeof z); fp_sub(&r->c1, &z, &t1);
}

/* ---------------- Jacobian group law, generic over the coordinate field ---------------- */
#define DEFINE_JAC(G, F)                                                                            \
  typedef struct { F x, y, z; } G;                                                                  \
  static void G##_set_inf(G* p) { memset(p, 0, sizeof *p); p->x = F##_ONE; p->y = F##_ONE; }        \
  static int G##_is_inf(const G* p) { return F##_is_zero(&p->z); }                                  \
  static void G##_dbl(G* r, const G* p) {                                                           \
    if (G##_is_inf(p) || F##_is_zero(&p->y)) { G##_set_inf(r); return; }                            \
    F a, b, c, d, e, f, t, x3, y3, z3;                                                              \
    F##_mul(&a, &p->x, &p->x); F##_mul(&b, &p->y, &p->y); F##_mul(&c, &b, &b);                      \
    F##_add(&t, &p->x, &b); F##_mul(&t, &t, &t); F##_sub(&t, &t, &a); F##_sub(&t, &t, &c); F##_add(&d, &t, &t); \
    F##_add(&e, &a, &a); F##_add(&e, &e, &a); F##_mul(&f, &e, &e);                                  \
    F##_sub(&x3, &f, &d); F##_sub(&x3, &x3, &d);                                                    \
    F##_sub(&t, &d, &x3); F##_mul(&y3, &e, &t);                                                     \
    F##_add(&c, &c, &c); F##_add(&c, &c, &c); F##_add(&c, &c, &c); F##_sub(&y3, &y3, &c);           \
    F##_mul(&z3, &p->y, &p->z); F##_add(&z3, &z3, &z3);                                             \
    r->x = x3; r->y = y3; r->z = z3;                                                                \
  }                                                                                                 \
  static void G##_add(G* r, const G* p, const G* q) {                                               \
    if (G##_is_inf(p)) { *r = *q; return; }                                                         \
    if (G##_is_inf(q)) { *r = *p; return; }                                                         \
    F z1z1, z2z2, u1, u2, s1, s2, h, rr, hh, hhh, v, t, x3, y3, z3;                                 \
    F##_mul(&z1z1, &p->z, &p->z); F##_mul(&z2z2, &q->z, &q->z);                                     \
    F##_mul(&u1, &p->x, &z2z2); F##_mul(&u2, &q->x, &z1z1);                                         \
    F##_mul(&t, &q->z, &z2z2); F##_mul(&s1, &p->y, &t);                                             \
    F##_mul(&t, &p->z, &z1z1); F##_mul(&s2, &q->y, &t);                                             \
    if (F##_eq(&u1, &u2)) { if (F##_eq(&s1, &s2)) { G##_dbl(r, p); return; } G##_set_inf(r); return; } \
    F##_sub(&h, &u2, &u1); F##_sub(&rr, &s2, &s1);                                                  \
    F##_mul(&hh, &h, &h); F##_mul(&hhh, &h, &hh); F##_mul(&v, &u1, &hh);                            \
    F##_mul(&x3, &rr, &rr); F##_sub(&x3, &x3, &hhh); F##_sub(&x3, &x3, &v); F##_sub(&x3, &x3, &v);  \
    F##_sub(&t, &v, &x3); F##_mul(&y3, &rr, &t); F##_mul(&t, &s1, &hhh); F##_sub(&y3, &y3, &t);     \
    F##_mul(&z3, &p->z, &q->z); F##_mul(&z3, &z3, &h);                                              \
    r->x = x3; r->y = y3; r->z = z3;                                                                \
  }                                                                                                 \
  static void G##_neg(G* r, const G* p) { F zero; memset(&zero, 0, sizeof zero); *r = *p; F##_sub(&r->y, &zero, &p->y); } \
  /* k * P, LSB-first over k's bit length with a running doubled base (kilic MulScalar) */          \
  static void G##_mul(G* r, const G* p, const fr* k_mont) {                                         \
    fr one_raw, k; memset(&one_raw, 0, sizeof one_raw); one_raw.v[0] = 1;                           \
    fr_mul(&k, k_mont, &one_raw); /* leave Montgomery form */                                       \
    int top = -1;                                                                                   \
    for (int i = 255; i >= 0; i--) if ((k.v[i >> 6] >> (i & 63)) & 1) { top = i; break; }           \
    G acc, base = *p;                                                                               \
    G##_set_inf(&acc);                                                                              \
    for (int i = 0; i <= top; i++) {                                                                \
      if ((k.v[i >> 6] >> (i & 63)) & 1) G##_add(&acc, &acc, &base);                                \
      G##_dbl(&base, &base);                                                                        \
    }                                                                                               \
    *r = acc;                                                                                       \
  }                                                                                                 \
  static void G##_to_affine(F* x, F* y, const G* p) { /* caller checks infinity */                  \
    F zi, zi2, zi3;                                                                                 \
    F##_inv(&zi, &p->z); F##_mul(&zi2, &zi, &zi); F##_mul(&zi3, &zi2, &zi);                         \
    F##_mul(x, &p->x, &zi2); F##_mul(y, &p->y, &zi3);                                               \
  }

DEFINE_JAC(e1, fp)
DEFINE_JAC(e2, fp2)

static e1 GEN1;
static e2 GEN2;
static int g_gen = 0;

/* generators as zcash-uncompressed bytes (96 B; 192 B with x = c1 || c0, y = c1 || c0) */
void op_set_generators(const uint8_t* g1_96, const uint8_t* g2_192) {
  ensure_init();
  fp2_ONE.c0 = fp_ONE; memset(&fp2_ONE.c1, 0, sizeof(fp));
  fp_from_be(&GEN1.x, g1_96); fp_from_be(&GEN1.y, g1_96 + 48); GEN1.z = fp_ONE;
  fp_from_be(&GEN2.x.c1, g2_192); fp_from_be(&GEN2.x.c0, g2_192 + 48);
  fp_from_be(&GEN2.y.c1, g2_192 + 96); fp_from_be(&GEN2.y.c0, g2_192 + 144); GEN2.z = fp2_ONE;
  g_gen = 1;
}
static void e1_out(uint8_t* b96, const e1* p) {
  if (e1_is_inf(p)) { memset(b96, 0, 96); b96[0] = 0x40; return; }
  fp x, y; e1_to_affine(&x, &y, p); fp_to_be(b96, &x); fp_to_be(b96 + 48, &y);
}
static void e2_out(uint8_t* b192, const e2* p) {
  if (e2_is_inf(p)) { memset(b192, 0, 192); b192[0] = 0x40; return; }
  fp2 x, y; e2_to_affine(&x, &y, p);
  fp_to_be(b192, &x.c1); fp_to_be(b192 + 48, &x.c0); fp_to_be(b192 + 96, &y.c1); fp_to_be(b192 + 144, &y.c0);
}

/* ---------------- polynomials ---------------- */
static void fr_set_i64(fr* r, long v) { /* Value.ToFieldElement, curve.go:17-19: SetInt64 reduces mod r */
  fr t; memset(&t, 0, sizeof t);
  if (v >= 0) { t.v[0] = (uint64_t)v; fr_mul(r, &t, &fr_R2); }
  else { t.v[0] = (uint64_t)(-v); fr_mul(&t, &t, &fr_R2); fr z; memset(&z, 0, sizeof z); fr_sub(r, &z, &t); }
}
static void poly_eval(fr* r, const fr* p, long n, const fr* x) { /* Poly.Eval, algebra.go:107-115 (Horner) */
  fr v; memset(&v, 0, sizeof v);
  for (long j = n - 1; j >= 0; j--) { fr_mul(&v, &v, x); fr_add(&v, &v, &p[j]); }
  *r = v;
}
/* Interpolate, algebra.go:254-280: p(j) = ys[j-1], j = 1..n; every Lagrange basis rebuilt from scratch
 * (lagrangeBasis :313-338: n-1 schoolbook products by (x - x_m) and n-1 inversions). out: n coefficients */
static void interpolate(fr* out, const fr* ys, long n) {
  fr* basis = (fr*)malloc((n + 1) * sizeof(fr));
  fr* next = (fr*)malloc((n + 1) * sizeof(fr));
  memset(out, 0, n * sizeof(fr));
  for (long j = 1; j <= n; j++) {
    long len = 1;
    basis[0] = fr_ONE;
    fr acc = fr_ONE, xj, xm, den, nxm, t;
    fr_set_i64(&xj, j);
    for (long m = 1; m <= n; m++) {
      if (m == j) continue;
      fr_set_i64(&xm, m);
      fr z; memset(&z, 0, sizeof z); fr_sub(&nxm, &z, &xm);
      /* basis = basis.Mul([-xm, 1]): every product computed, as Poly.Mul does (algebra.go:92-105) */
      memset(next, 0, (len + 1) * sizeof(fr));
      for (long i = 0; i < len; i++) {
        fr_mul(&t, &basis[i], &nxm); fr_add(&next[i], &next[i], &t);
        fr_mul(&t, &basis[i], &fr_ONE); fr_add(&next[i + 1], &next[i + 1], &t);
      }
      len++;
      fr* sw = basis; basis = next; next = sw;
      fr_sub(&den, &xj, &xm); fr_inv(&den, &den); fr_mul(&acc, &acc, &den);
    }
    for (long i = 0; i < len; i++) { fr_mul(&basis[i], &basis[i], &acc); fr_mul(&basis[i], &basis[i], &ys[j - 1]); }
    for (long i = 0; i < len && i < n; i++) fr_add(&out[i], &out[i], &basis[i]);
  }
  free(basis); free(next);
}

typedef struct {
  long n, m, nio;
  fr *left, *right, *out; /* m x n coefficients each */
  fr* z;                  /* n + 1 */
} qap_t;

/* ToQAP, qap.go:35-65: per-variable interpolation of the transposed gate matrices, z = prod (x - i).
 * mats: three n x m row-major matrices of Go ints (gate rows). */
static void to_qap(qap_t* q, const long* L, const long* R, const long* O, long n, long m, long nio, int threads) {
  q->n = n; q->m = m; q->nio = nio;
  q->left = (fr*)calloc(m * n, sizeof(fr)); q->right = (fr*)calloc(m * n, sizeof(fr)); q->out = (fr*)calloc(m * n, sizeof(fr));
  const long* mats[3] = {L, R, O};
  fr* dst[3] = {q->left, q->right, q->out};
  (void)threads;
#pragma omp parallel for num_threads(threads) schedule(dynamic, 1) collapse(2)
  for (int k = 0; k < 3; k++)
    for (long i = 0; i < m; i++) {
      fr* ys = (fr*)malloc(n * sizeof(fr));
      for (long j = 0; j < n; j++) fr_set_i64(&ys[j], mats[k][j * m + i]);
      interpolate(dst[k] + i * n, ys, n);
      free(ys);
    }
  /* z = (x-1)(x-2)...(x-n) by repeated Poly.Mul */
  q->z = (fr*)calloc(n + 2, sizeof(fr));
  fr* next = (fr*)calloc(n + 2, sizeof(fr));
  long len = 0;
  for (long i = 1; i <= n; i++) {
    fr xi, nxi, zero, t; memset(&zero, 0, sizeof zero);
    fr_set_i64(&xi, i); fr_sub(&nxi, &zero, &xi);
    if (len == 0) { q->z[0] = nxi; q->z[1] = fr_ONE; len = 2; continue; }
    memset(next, 0, (len + 1) * sizeof(fr));
    for (long a = 0; a < len; a++) {
      fr_mul(&t, &q->z[a], &nxi); fr_add(&next[a], &next[a], &t);
      fr_mul(&t, &q->z[a], &fr_ONE); fr_add(&next[a + 1], &next[a + 1], &t);
    }
    len++;
    memcpy(q->z, next, len * sizeof(fr));
  }
  free(next);
}
/* The same polynomials as to_qap (the interpolant is unique) without the reference's O(m n^3) cost, for
 * the PROVE-ONLY timings of bench.py at sizes where the reference's ToQAP itself no longer finishes: the
 * Lagrange basis l_j = z / ((x - j) z'(j)) once (synthetic division, z'(j) = (-1)^(n-j) (j-1)! (n-j)!),
 * then each variable's polynomial as the combination of the basis over its non-zero gates.  Untimed setup. */
static void to_qap_fast(qap_t* q, const long* L, const long* R, const long* O, long n, long m, long nio, int threads) {
  q->n = n; q->m = m; q->nio = nio;
  q->left = (fr*)calloc(m * n, sizeof(fr)); q->right = (fr*)calloc(m * n, sizeof(fr)); q->out = (fr*)calloc(m * n, sizeof(fr));
  q->z = (fr*)calloc(n + 2, sizeof(fr));
  fr zero; memset(&zero, 0, sizeof zero);
  q->z[0] = fr_ONE;
  long len = 1;
  for (long i = 1; i <= n; i++) {       /* z *= (x - i), in place from the top */
    fr xi; fr_set_i64(&xi, i);
    q->z[len] = q->z[len - 1];
    for (long a = len - 1; a >= 1; a--) { fr t; fr_mul(&t, &q->z[a], &xi); fr_sub(&q->z[a], &q->z[a - 1], &t); }
    { fr t; fr_mul(&t, &q->z[0], &xi); fr_sub(&q->z[0], &zero, &t); }
    len++;
  }
  fr* fact = (fr*)malloc((n + 1) * sizeof(fr));
  fact[0] = fr_ONE;
  for (long i = 1; i <= n; i++) { fr fi; fr_set_i64(&fi, i); fr_mul(&fact[i], &fact[i - 1], &fi); }
  fr* basis = (fr*)malloc(n * n * sizeof(fr));
  (void)threads;
#pragma omp parallel for num_threads(threads) schedule(dynamic, 8)
  for (long j = 1; j <= n; j++) {
    fr xj, zp, zpi; fr_set_i64(&xj, j);
    fr* b = basis + (j - 1) * n;
    /* synthetic division of z by (x - j): b[n-1] = z[n], b[k-1] = z[k] + j b[k] */
    b[n - 1] = q->z[n];
    for (long k = n - 1; k >= 1; k--) { fr t; fr_mul(&t, &b[k], &xj); fr_add(&b[k - 1], &q->z[k], &t); }
    fr_mul(&zp, &fact[j - 1], &fact[n - j]);
    if ((n - j) & 1) fr_sub(&zp, &zero, &zp);
    fr_inv(&zpi, &zp);
    for (long k = 0; k < n; k++) fr_mul(&b[k], &b[k], &zpi);
  }
  const long* mats[3] = {L, R, O};
  fr* dst[3] = {q->left, q->right, q->out};
#pragma omp parallel for num_threads(threads) schedule(dynamic, 8) collapse(2)
  for (int k = 0; k < 3; k++)
    for (long i = 0; i < m; i++) {
      fr* p = dst[k] + i * n;
      for (long j = 0; j < n; j++) {
        long v = mats[k][j * m + i];
        if (!v) continue;
        fr y, t; fr_set_i64(&y, v);
        for (long c = 0; c < n; c++) { fr_mul(&t, &basis[j * n + c], &y); fr_add(&p[c], &p[c], &t); }
      }
    }
  free(fact); free(basis);
}
static void qap_free(qap_t* q) { free(q->left); free(q->right); free(q->out); free(q->z); }

/* QAP.Quotient, qap.go:151-162 (computeAggregatePoly :164-175, Mul, Sub, Div2); h: n-1 coefficients.
 * Returns 1 on a non-zero remainder ("apocalypse"). */
static int qap_quotient(const qap_t* q, const fr* sol, fr* h) {
  long n = q->n, m = q->m;
  fr* a = (fr*)calloc(n, sizeof(fr)); fr* b = (fr*)calloc(n, sizeof(fr)); fr* c = (fr*)calloc(n, sizeof(fr));
  fr t;
  for (long i = 0; i < m; i++)
    for (long k = 0; k < n; k++) {
      fr_mul(&t, &q->left[i * n + k], &sol[i]); fr_add(&a[k], &a[k], &t);
      fr_mul(&t, &q->right[i * n + k], &sol[i]); fr_add(&b[k], &b[k], &t);
      fr_mul(&t, &q->out[i * n + k], &sol[i]); fr_add(&c[k], &c[k], &t);
    }
  long lr = 2 * n - 1, lz = n + 1;
  fr* r = poly_mul(a, n, b, n);
  for (long i = 0; i < n; i++) fr_sub(&r[i], &r[i], &c[i]);
  memset(h, 0, (n - 1) * sizeof(fr));
  fr zlead_inv;
  fr_inv(&zlead_inv, &q->z[lz - 1]);
  while (lr > 0 && lr >= lz) {   /* Div2, algebra.go:140-159: one full tPoly.Mul(p2) per step */
    fr_mul(&t, &r[lr - 1], &zlead_inv);
    long deg = lr - lz;
    fr_add(&h[deg], &h[deg], &t);
    fr* tp = (fr*)calloc(deg + 1, sizeof(fr));
    tp[deg] = t;
    fr* pr = poly_mul(tp, deg + 1, q->z, lz);
    for (long i = 0; i < lr; i++) fr_sub(&r[i], &r[i], &pr[i]);
    free(tp); free(pr);
    lr--;
  }
  int nonzero = 0;
  for (long i = 0; i < lr; i++) if (!fr_is_zero(&r[i])) nonzero = 1;
  free(a); free(b); free(c); free(r);
  return nonzero;
}

/* ---------------- Groth16 ---------------- */
typedef struct {
  e1 Alpha, Beta, Delta, *Xi, *NioLP, *XiT;
  e2 Beta2, Delta2, *Xi2;
} g16_key;

static void linear_poly_for_var(fr* r, const qap_t* q, long i, const fr* x, const fr* alpha, const fr* beta) {
  fr ui, vi, wi, t;   /* groth16.go:238-249 */
  poly_eval(&ui, q->left + i * q->n, q->n, x); fr_mul(&ui, &ui, beta);
  poly_eval(&vi, q->right + i * q->n, q->n, x); fr_mul(&vi, &vi, alpha);
  poly_eval(&wi, q->out + i * q->n, q->n, x);
  fr_add(&t, &ui, &vi); fr_add(r, &wi, &t);
}

/* NewGroth16TrustedSetup, groth16.go:64-101, prover side (IoLP / Gamma belong to the verifier) */
static void g16_setup(g16_key* k, const qap_t* q, const fr* alpha, const fr* beta, const fr* delta, const fr* x, int threads) {
  long n = q->n, m = q->m, diff = q->m - q->nio;
  (void)threads;
  e1_mul(&k->Alpha, &GEN1, alpha); e1_mul(&k->Beta, &GEN1, beta); e2_mul(&k->Beta2, &GEN2, beta);
  e1_mul(&k->Delta, &GEN1, delta); e2_mul(&k->Delta2, &GEN2, delta);
  k->Xi = (e1*)malloc(n * sizeof(e1)); k->Xi2 = (e2*)malloc(n * sizeof(e2));
  k->XiT = (e1*)malloc((n - 1) * sizeof(e1)); k->NioLP = (e1*)malloc((q->nio ? q->nio : 1) * sizeof(e1));
  fr* pw = (fr*)malloc(n * sizeof(fr));
  pw[0] = fr_ONE;
  for (long i = 1; i < n; i++) fr_mul(&pw[i], &pw[i - 1], x);     /* GeneratePowersCommit, algebra.go:371-384 */
  fr dinv, tx, txd;
  fr_inv(&dinv, delta);
  poly_eval(&tx, q->z, n + 1, x); fr_mul(&txd, &tx, &dinv);
#pragma omp parallel for num_threads(threads) schedule(dynamic, 4)
  for (long i = 0; i < n; i++) {
    e1_mul(&k->Xi[i], &GEN1, &pw[i]);
    e2_mul(&k->Xi2[i], &GEN2, &pw[i]);
    if (i < n - 1) { fr t; fr_mul(&t, &pw[i], &txd); e1_mul(&k->XiT[i], &GEN1, &t); }
  }
#pragma omp parallel for num_threads(threads) schedule(dynamic, 4)
  for (long i = diff; i < m; i++) {   /* fullLinearPoly(qap, diff, nbVars, ..., delta), groth16.go:254-264 */
    fr lp; linear_poly_for_var(&lp, q, i, x, alpha, beta); fr_mul(&lp, &lp, &dinv);
    e1_mul(&k->NioLP[i - diff], &GEN1, &lp);
  }
  free(pw);
}
static void g16_key_free(g16_key* k) { free(k->Xi); free(k->Xi2); free(k->XiT); free(k->NioLP); }

/* sumBlind, groth16.go:134-141: sum_i sol[i] * BlindEval(polys[i], xi) -- m*n scalar-muls */
#define DEFINE_SUMBLIND(G)                                                                          \
  static void G##_sum_blind(G* out, const fr* polys, const fr* sol, long m, long n, const G* xi, int threads) { \
    G total; G##_set_inf(&total);                                                                   \
    (void)threads;                                                                                  \
    _Pragma("omp parallel num_threads(threads)")                                                    \
    {                                                                                               \
      G local; G##_set_inf(&local);                                                                 \
      _Pragma("omp for schedule(dynamic, 1)")                                                       \
      for (long i = 0; i < m; i++) {                                                                \
        G uix, tmp; G##_set_inf(&uix);                                                              \
        for (long k = 0; k < n; k++) { G##_mul(&tmp, &xi[k], &polys[i * n + k]); G##_add(&uix, &uix, &tmp); } /* BlindEval */ \
        G##_mul(&uix, &uix, &sol[i]);                                                               \
        G##_add(&local, &local, &uix);                                                              \
      }                                                                                             \
      _Pragma("omp critical")                                                                       \
      G##_add(&total, &total, &local);                                                              \
    }                                                                                               \
    *out = total;                                                                                   \
  }
DEFINE_SUMBLIND(e1)
DEFINE_SUMBLIND(e2)

/* Groth16Prove, groth16.go:122-211, with (r, s) injected.  Returns 1 on "apocalypse". */
static int g16_prove(const g16_key* k, const qap_t* q, const fr* sol, const fr* r, const fr* s, e1* A, e2* B, e1* C, fr* h_out,
                     int threads) {
  long n = q->n, m = q->m, diff = q->m - q->nio;
  e1 t1; e2 t2;
  e1_sum_blind(A, q->left, sol, m, n, k->Xi, threads);
  e1_mul(&t1, &k->Delta, r); e1_add(A, A, &t1); e1_add(A, &k->Alpha, A);
  e2_sum_blind(B, q->right, sol, m, n, k->Xi2, threads);
  e2_mul(&t2, &k->Delta2, s); e2_add(B, B, &t2); e2_add(B, &k->Beta2, B);
  e1 nio, htd, As, B1, Br, rsd;
  e1_set_inf(&nio);
  for (long i = 0; i < q->nio; i++) { e1_mul(&t1, &k->NioLP[i], &sol[i + diff]); e1_add(&nio, &nio, &t1); }
  fr* h = h_out ? h_out : (fr*)malloc((n - 1) * sizeof(fr));
  int bad = qap_quotient(q, sol, h);
  e1_set_inf(&htd);
  for (long i = 0; i < n - 1; i++) { e1_mul(&t1, &k->XiT[i], &h[i]); e1_add(&htd, &htd, &t1); }   /* BlindEval */
  if (!h_out) free(h);
  e1_mul(&As, A, s);
  e1_sum_blind(&B1, q->right, sol, m, n, k->Xi, threads);
  e1_mul(&t1, &k->Delta, s); e1_add(&B1, &B1, &t1); e1_add(&B1, &B1, &k->Beta);
  e1_mul(&Br, &B1, r);
  fr rs; fr_mul(&rs, r, s);
  e1_mul(&rsd, &k->Delta, &rs); e1_neg(&rsd, &rsd);
  e1_set_inf(C);
  e1_add(C, C, &nio); e1_add(C, C, &htd); e1_add(C, C, &As); e1_add(C, C, &Br); e1_add(C, C, &rsd);
  return bad;
}

/* One run of the reference's Groth16 flow on an R1CS given as three n x m matrices of Go ints:
 * ToQAP -> trusted setup (toxic = alpha, beta, delta, x as 32 B big-endian) -> prove with (r, s).
 * witness: m x 32 B big-endian Fr.  out: A 96 B | B 192 B | C 96 B zcash-uncompressed; h_be (optional)
 * n-1 coefficients; seconds[3] = ToQAP, setup, prove; fast_qap != 0 replaces the reference's ToQAP by
 * to_qap_fast (same polynomials; for prove-only timings).  Returns 0, or 1 on "apocalypse". */
int op_groth16_flow(const long* L, const long* R, const long* O, long n, long m, long nio, const uint8_t* witness_be,
                    const uint8_t* toxic_be, const uint8_t* r_be, const uint8_t* s_be, int threads, int fast_qap,
                    uint8_t* out384, uint8_t* h_be, double* seconds) {
  if (!g_gen) return -1;
  if (threads < 1) threads = 1;
  qap_t q; g16_key k;
  double t0 = now_s();
  if (fast_qap) to_qap_fast(&q, L, R, O, n, m, nio, threads); else to_qap(&q, L, R, O, n, m, nio, threads);
  double t1 = now_s();
  fr tox[4], r, s;
  for (int i = 0; i < 4; i++) fr_from_be(&tox[i], toxic_be + 32 * i);
  fr_from_be(&r, r_be); fr_from_be(&s, s_be);
  g16_setup(&k, &q, &tox[0], &tox[1], &tox[2], &tox[3], threads);
  double t2 = now_s();
  fr* sol = poly_load(witness_be, m);
  fr* h = (fr*)malloc((n > 1 ? n - 1 : 1) * sizeof(fr));
  e1 A, C; e2 B;
  int bad = g16_prove(&k, &q, sol, &r, &s, &A, &B, &C, h, threads);
  double t3 = now_s();
  e1_out(out384, &A); e2_out(out384 + 96, &B); e1_out(out384 + 288, &C);
  if (h_be) for (long i = 0; i < n - 1; i++) fr_to_be(h_be + 32 * i, &h[i]);
  if (seconds) { seconds[0] = t1 - t0; seconds[1] = t2 - t1; seconds[2] = t3 - t2; }
  free(sol); free(h); g16_key_free(&k); qap_free(&q);
  return bad;
}

/* ---------------- PHGR13 ---------------- */
/* generateEvalCommit, pinochio.go:381-388: { (shift * p_i(x)) * base } */
#define DEFINE_EVALCOMMIT(G)                                                                        \
  static void G##_eval_commit(G* out, const G* base, const fr* polys, long count, long n, const fr* x, const fr* shift, int threads) { \
    (void)threads;                                                                                  \
    _Pragma("omp parallel for num_threads(threads) schedule(dynamic, 4)")                           \
    for (long i = 0; i < count; i++) { fr t; poly_eval(&t, polys + i * n, n, x); fr_mul(&t, &t, shift); G##_mul(&out[i], base, &t); } \
  }
DEFINE_EVALCOMMIT(e1)
DEFINE_EVALCOMMIT(e2)

/* ToQAP -> NewPHGR13TrustedSetup's evaluation key (toxic = s, av, aw, ay, rv, rw, beta) -> PHGR13Prove.
 * out: hs, vss, yss, vass, wass, yass, gz (7 x 96 B) then wss (192 B), the order of ps_phgr13_prove.
 * seconds[3] = ToQAP, setup (EK only), prove. */
int op_phgr13_flow(const long* L, const long* R, const long* O, long n, long m, long nio, const uint8_t* witness_be,
                   const uint8_t* toxic_be, int threads, int fast_qap, uint8_t* out864, uint8_t* h_be, double* seconds) {
  if (!g_gen) return -1;
  if (threads < 1) threads = 1;
  qap_t q;
  double t0 = now_s();
  if (fast_qap) to_qap_fast(&q, L, R, O, n, m, nio, threads); else to_qap(&q, L, R, O, n, m, nio, threads);
  double t1 = now_s();
  fr tox[7];
  for (int i = 0; i < 7; i++) fr_from_be(&tox[i], toxic_be + 32 * i);
  const fr *s = &tox[0], *av = &tox[1], *aw = &tox[2], *ay = &tox[3], *rv = &tox[4], *rw = &tox[5], *beta = &tox[6];
  long diff = m - nio;
  e1 gv, g1w, gy; e2 gw; fr ry;
  e1_mul(&gv, &GEN1, rv); e2_mul(&gw, &GEN2, rw); e1_mul(&g1w, &GEN1, rw);
  fr_mul(&ry, rv, rw); e1_mul(&gy, &GEN1, &ry);
  e1* gsi = (e1*)malloc((n - 1) * sizeof(e1));
  {
    fr* pw = (fr*)malloc((n - 1) * sizeof(fr));
    pw[0] = fr_ONE;
    for (long i = 1; i < n - 1; i++) fr_mul(&pw[i], &pw[i - 1], s);
#pragma omp parallel for num_threads(threads) schedule(dynamic, 4)
    for (long i = 0; i < n - 1; i++) e1_mul(&gsi[i], &GEN1, &pw[i]);
    free(pw);
  }
  long cnt = nio ? nio : 1;
  e1* g1k[8]; for (int i = 0; i < 8; i++) g1k[i] = (e1*)malloc(cnt * sizeof(e1));
  e2* ws = (e2*)malloc(cnt * sizeof(e2));
  const fr *pl = q.left + diff * n, *pr = q.right + diff * n, *po = q.out + diff * n;
  e1_eval_commit(g1k[0], &gv, pl, nio, n, s, &fr_ONE, threads);   /* vs  */
  e2_eval_commit(ws, &gw, pr, nio, n, s, &fr_ONE, threads);       /* ws  */
  e1_eval_commit(g1k[1], &gy, po, nio, n, s, &fr_ONE, threads);   /* ys  */
  e1_eval_commit(g1k[2], &gv, pl, nio, n, s, av, threads);        /* vas */
  e1_eval_commit(g1k[3], &g1w, pr, nio, n, s, aw, threads);       /* was */
  e1_eval_commit(g1k[4], &gy, po, nio, n, s, ay, threads);        /* yas */
  e1_eval_commit(g1k[5], &gv, pl, nio, n, s, beta, threads);      /* vbs */
  e1_eval_commit(g1k[6], &g1w, pr, nio, n, s, beta, threads);     /* wbs */
  e1_eval_commit(g1k[7], &gy, po, nio, n, s, beta, threads);      /* ybs */
  double t2 = now_s();
  /* PHGR13Prove, pinochio.go:207-254 */
  fr* sol = poly_load(witness_be, m);
  fr* h = (fr*)malloc((n > 1 ? n - 1 : 1) * sizeof(fr));
  int bad = qap_quotient(&q, sol, h);
  e1 res[8], t; e2 wss, t2p;
  e1_set_inf(&res[0]);
  for (long i = 0; i < n - 1; i++) { e1_mul(&t, &gsi[i], &h[i]); e1_add(&res[0], &res[0], &t); }     /* hs */
  for (int kx = 0; kx < 8; kx++) {     /* computeSolCommit, pinochio.go:222-229 */
    e1 acc; e1_set_inf(&acc);
    for (long i = 0; i < nio; i++) { e1_mul(&t, &g1k[kx][i], &sol[diff + i]); e1_add(&acc, &acc, &t); }
    g1k[kx][0] = acc;  /* reuse slot 0 as the result */
  }
  e2_set_inf(&wss);
  for (long i = 0; i < nio; i++) { e2_mul(&t2p, &ws[i], &sol[diff + i]); e2_add(&wss, &wss, &t2p); }
  e1 gz; e1_add(&gz, &g1k[6][0], &g1k[7][0]); e1_add(&gz, &g1k[5][0], &gz);
  double t3 = now_s();
  e1_out(out864, &res[0]);
  e1_out(out864 + 96, &g1k[0][0]); e1_out(out864 + 192, &g1k[1][0]); e1_out(out864 + 288, &g1k[2][0]);
  e1_out(out864 + 384, &g1k[3][0]); e1_out(out864 + 480, &g1k[4][0]); e1_out(out864 + 576, &gz);
  e2_out(out864 + 672, &wss);
  if (h_be) for (long i = 0; i < n - 1; i++) fr_to_be(h_be + 32 * i, &h[i]);
  if (seconds) { seconds[0] = t1 - t0; seconds[1] = t2 - t1; seconds[2] = t3 - t2; }
  for (int i = 0; i < 8; i++) free(g1k[i]);
  free(ws); free(gsi); free(sol); free(h); qap_free(&q);
  return bad;
}

/* Poly.BlindEval over G1 (algebra.go:348-359) with the terms spread over `threads` OpenMP threads:
 * the "all host threads" form of oc_blind_eval_g1 for bench.py's --impl reference leg. */
long op_blind_eval_g1_mt(const uint8_t* points, const uint8_t* scalars, long n, int threads, uint8_t* out96) {
  ensure_init();
  if (threads < 1) threads = 1;
  e1 total; e1_set_inf(&total);
#pragma omp parallel num_threads(threads)
  {
    e1 local; e1_set_inf(&local);
#pragma omp for schedule(static)
    for (long i = 0; i < n; i++) {
      e1 base, tmp; fr k;
      if (points[96 * i] & 0x40) continue;
      fp_from_be(&base.x, points + 96 * i); fp_from_be(&base.y, points + 96 * i + 48); base.z = fp_ONE;
      fr_from_be(&k, scalars + 32 * i);
      e1_mul(&tmp, &base, &k);
      e1_add(&local, &local, &tmp);
    }
#pragma omp critical
    e1_add(&total, &total, &local);
  }
  e1_out(out96, &total);
  return n;
}

int op_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
