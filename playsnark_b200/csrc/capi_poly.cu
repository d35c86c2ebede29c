// C ABI, part 2: every entry point whose device work is Fr polynomial arithmetic -- NTT, the QAP loaders,
// QAP.Quotient (qap.go:151-162) in its dense and sparse forms, the scalar assembly of the Groth16 proof and
// the per-rank steps of the sharded prover.  This translation unit instantiates the kernels of ntt.cuh,
// poly.cuh, interp.cuh and the Fr half of codec.cuh; capi.cu (contexts, base sets, MSMs, proof assembly)
// reaches them through poly_api.cuh.
#include "codec.cuh"
#include "interp.cuh"
#include "multi_api.cuh"
#include "setup.cuh"

#include <algorithm>
#include <cstdlib>
#include <new>
#include <vector>

using namespace ps;

namespace ps {

int ctx_ntt_tables(ps_ctx* ctx, int log_n, const NttTables** out) {
  if (log_n < 0 || log_n > 30) return PS_ERR_ARG;
  NttTables& t = ctx->ntt_cache[log_n];
  if (t.log_n != log_n) PS_TRY(ntt_tables_build(ctx->stream, log_n, &t));
  *out = &t;
  return PS_OK;
}

// ---- Groth16 / PHGR13 scalar assembly ---------------------------------------------------------------
// dst[k] = src[k], k < n (device-to-device gather of Fr)
struct FrCopyK {
  static constexpr int BLOCK = 256;
  PS_DEV static void run(uint32_t k, const Fr* src, Fr* dst) { dst[k] = src[k]; }
};
// dst[k] = s*a[k] + r*b[k]
struct FrAxpbyK {
  static constexpr int BLOCK = 256;
  PS_DEV static void run(uint32_t k, Fr s, const Fr* a, Fr r, const Fr* b, Fr* dst) { dst[k] = s * a[k] + r * b[k]; }
};
// dst[0..3) = given constants
struct FrSet3K {
  static constexpr int BLOCK = 32;
  PS_DEV static void run(uint32_t k, Fr x0, Fr x1, Fr x2, int cnt, Fr* dst) {
    if ((int)k < cnt) dst[k] = k == 0 ? x0 : (k == 1 ? x1 : x2);
  }
};

// scalars (host, big-endian) -> device limbs; returns PS_ERR_ENCODING for values >= r
int stage_scalars(ps_ctx* ctx, const uint8_t* scalars_be, size_t n, int mont, uint32_t** d_out, uint32_t** d_err_out) {
  ps_stream_t st = ctx->stream;
  uint8_t* d_in = ctx->arena.take<uint8_t>(n * 32);
  uint32_t* d_sc = ctx->arena.take<uint32_t>(n * 8);
  uint32_t* d_err = ctx->arena.take<uint32_t>(1);
  if (!d_in || !d_sc || !d_err) return PS_ERR_ALLOC;
  PS_TRY(dev_memset(d_err, 0, 4, st));
  if (n) PS_TRY(dev_h2d(d_in, scalars_be, n * 32, st));
  PS_LAUNCH(FrFromBytesK, st, n, (const uint8_t*)d_in, d_sc, mont, d_err);
  *d_out = d_sc;
  *d_err_out = d_err;
  return PS_OK;
}

int run_quotient_sparse(ps_ctx* ctx, const ps_qap* q, const uint8_t* witness_be, QuotientBufs* o, bool want_c) {
  const SparseQap* sq = (const SparseQap*)q->sparse;
  const uint32_t n = (uint32_t)q->n, np = sq->np;   // real gates; tree leaves (rows above n are empty, their evaluations 0)
  ps_stream_t st = ctx->stream;
  uint32_t* d_w = nullptr;
  PS_TRY(stage_scalars(ctx, witness_be, q->m, 1, &d_w, &o->enc_err));
  o->w = (Fr*)d_w;
  Fr* ev = ctx->arena.take<Fr>((size_t)3 * np);
  Fr* coef = ctx->arena.take<Fr>((size_t)3 * np);
  o->h = ctx->arena.take<Fr>(np);
  o->flag = ctx->arena.take<uint32_t>(1);
  if (!ev || !coef || !o->h || !o->flag) return PS_ERR_ALLOC;
  PS_TRY(dev_memset(o->flag, 0, 4, st));
  PS_LAUNCH(SpmvK, st, (size_t)3 * np, np, 0u, (const uint32_t*)sq->mat[0].row_ptr, (const uint32_t*)sq->mat[0].col, (const Fr*)sq->mat[0].val,
            (const uint32_t*)sq->mat[1].row_ptr, (const uint32_t*)sq->mat[1].col, (const Fr*)sq->mat[1].val,
            (const uint32_t*)sq->mat[2].row_ptr, (const uint32_t*)sq->mat[2].col, (const Fr*)sq->mat[2].val, (const Fr*)o->w, ev);
  PS_LAUNCH(GateCheckK, st, np, np, (const Fr*)ev, o->flag);
  // only a and b are interpolated: c = a*b mod z never has to exist for the proof
  PS_TRY(interpolate_ap(ctx, sq, np, q->log_np, 2, ev, coef));
  o->a = coef; o->b = coef + np; o->c = coef + 2 * (size_t)np;   // np entries each, zero from n on
  PS_TRY(quotient_series(ctx, sq, n, q->log_np, o->a, o->b, o->h, want_c ? o->c : (Fr*)nullptr));
  return PS_OK;
}

int run_quotient(ps_ctx* ctx, const ps_qap* q, const uint8_t* witness_be, QuotientBufs* o, bool want_c) {
  if (!q->dense) return run_quotient_sparse(ctx, q, witness_be, o, want_c);
  const uint32_t np = 1u << q->log_np;
  uint32_t* d_w = nullptr;
  PS_TRY(stage_scalars(ctx, witness_be, q->m, 1, &d_w, &o->enc_err));
  o->w = (Fr*)d_w;
  o->a = ctx->arena.take<Fr>(np); o->b = ctx->arena.take<Fr>(np); o->c = ctx->arena.take<Fr>(np); o->h = ctx->arena.take<Fr>(np);
  o->flag = ctx->arena.take<uint32_t>(1);
  if (!o->a || !o->b || !o->c || !o->h || !o->flag) return PS_ERR_ALLOC;
  PS_TRY(qap_aggregate_dense(ctx, q, o->w, o->a, o->b, o->c));
  PS_TRY(quotient_from_abc(ctx, q, o->a, o->b, o->c, o->h, o->flag));
  return PS_OK;
}

int export_fr(ps_ctx* ctx, const Fr* d_src, size_t count, uint8_t* host_out) {
  uint8_t* d_bytes = ctx->arena.take<uint8_t>(count * 32);
  if (!d_bytes) return PS_ERR_ALLOC;
  PS_LAUNCH(FrToBytesK, ctx->stream, count, (const uint32_t*)d_src, d_bytes, 1);
  PS_TRY(dev_d2h(host_out, d_bytes, count * 32, ctx->stream));
  return PS_OK;
}

// Fr vectors in Montgomery form -> standard form, in place
struct FrFromMontK {
  static constexpr int BLOCK = 256;
  PS_DEV static void run(uint32_t i, Fr* a) { a[i] = a[i].from_mont(); }
};

int g16_build_scalars(ps_ctx* ctx, const ps_g16_key* key, const ps_qap* qap, const uint8_t* witness_be, const uint8_t* r_be,
                      const uint8_t* s_be, G16Scalars* o) {
  if (key->n != qap->n || key->n_nio != qap->n_io) return PS_ERR_LENGTH;
  ps_stream_t st = ctx->stream;
  const size_t n = qap->n, nio = qap->n_io, diff = qap->m - qap->n_io;
  PS_TRY(run_quotient(ctx, qap, witness_be, &o->qb));
  Fr hrs[2];
  PS_TRY(parse_fr(r_be, &hrs[0]));
  PS_TRY(parse_fr(s_be, &hrs[1]));
  const Fr r = hrs[0], s = hrs[1], rs = hrs[0] * hrs[1];
  o->nA = n + 2; o->nB = n + 2; o->nC = nio + (n - 1) + n + 3;
  o->scA = ctx->arena.take<Fr>(o->nA);
  o->scB = ctx->arena.take<Fr>(o->nB);
  o->scC = ctx->arena.take<Fr>(o->nC);
  if (!o->scA || !o->scB || !o->scC) return PS_ERR_ALLOC;
  const QuotientBufs& qb = o->qb;
  PS_LAUNCH(FrCopyK, st, n, (const Fr*)qb.a, o->scA);
  PS_LAUNCH(FrSet3K, st, 2, r, Fr::one(), Fr::zero(), 2, o->scA + n);
  PS_LAUNCH(FrCopyK, st, n, (const Fr*)qb.b, o->scB);
  PS_LAUNCH(FrSet3K, st, 2, s, Fr::one(), Fr::zero(), 2, o->scB + n);
  PS_LAUNCH(FrCopyK, st, nio, (const Fr*)(qb.w + diff), o->scC);
  PS_LAUNCH(FrCopyK, st, n - 1, (const Fr*)qb.h, o->scC + nio);
  PS_LAUNCH(FrAxpbyK, st, n, s, (const Fr*)qb.a, r, (const Fr*)qb.b, o->scC + nio + (n - 1));
  PS_LAUNCH(FrSet3K, st, 3, s, r, rs, 3, o->scC + nio + (n - 1) + n);
  return PS_OK;
}
// out[i] = in[i] in standard form
struct FrStdCopyK {
  static constexpr int BLOCK = 256;
  PS_DEV static void run(uint32_t i, const Fr* in, Fr* out) { out[i] = in[i].from_mont(); }
};
// *status |= (enc ? 1 : 0) | (rem ? 2 : 0)      (device-side status word: no host round trip per call)
struct StatusMergeK {
  static constexpr int BLOCK = 32;
  PS_DEV static void run(uint32_t i, const uint32_t* enc, const uint32_t* rem, uint32_t* status) {
    if (i) return;
    uint32_t v = ((enc && *enc) ? 1u : 0u) | ((rem && *rem) ? 2u : 0u);
    if (v) ps_atomic_or(status, v);
  }
};

// body of ps_qap_interp_part once the witness is on the device (Montgomery form); d_err may be null
int interp_part_run(ps_ctx* ctx, const ps_qap* qap, const Fr* d_w, const uint32_t* d_err, int which, size_t part, size_t parts,
                    int lp, void* d_out_evals, void* d_w_nio_out, void* d_status) {
  const SparseQap* sq = (const SparseQap*)qap->sparse;
  const uint32_t n = sq->np, ns = (uint32_t)(sq->np / parts), lo = (uint32_t)part * ns;   // n: leaves of the tree
  ps_stream_t st = ctx->stream;
  if (d_w_nio_out) PS_LAUNCH(FrStdCopyK, st, qap->n_io, d_w + (qap->m - qap->n_io), (Fr*)d_w_nio_out);
  Fr* ev = ctx->arena.take<Fr>((size_t)3 * ns);
  uint32_t* flag = ctx->arena.take<uint32_t>(1);
  if (!ev || !flag) return PS_ERR_ALLOC;
  PS_TRY(dev_memset(flag, 0, 4, st));
  PS_LAUNCH(SpmvK, st, (size_t)3 * ns, ns, lo, (const uint32_t*)sq->mat[0].row_ptr, (const uint32_t*)sq->mat[0].col, (const Fr*)sq->mat[0].val,
            (const uint32_t*)sq->mat[1].row_ptr, (const uint32_t*)sq->mat[1].col, (const Fr*)sq->mat[1].val,
            (const uint32_t*)sq->mat[2].row_ptr, (const uint32_t*)sq->mat[2].col, (const Fr*)sq->mat[2].val, d_w, ev);
  PS_LAUNCH(GateCheckK, st, ns, ns, (const Fr*)ev, flag);   // this rank's gates; every gate is checked by some rank
  // parts == 1: the whole tree, d_out_evals receives the qap->n coefficients (the tree produces np, zero from n on)
  Fr* coef = lp ? (Fr*)nullptr : ctx->arena.take<Fr>(n);
  if (!lp && !coef) return PS_ERR_ALLOC;
  PS_TRY(interpolate_from_leaves(ctx, sq, n, qap->log_np, 1, lo, ns, qap->log_np - lp, (const Fr*)(ev + (size_t)which * ns),
                                 lp ? (Fr*)d_out_evals : (Fr*)nullptr, coef));
  if (!lp) PS_TRY(dev_d2d(d_out_evals, coef, qap->n * sizeof(Fr), st));
  PS_LAUNCH(StatusMergeK, st, 1, d_err, (const uint32_t*)flag, (uint32_t*)d_status);
  return PS_OK;
}
// dst[k] = (s a[lo + k] + r b[lo + k]) in standard form
struct FrAxpbyStdK {
  static constexpr int BLOCK = 256;
  PS_DEV static void run(uint32_t k, Fr s, const Fr* a, Fr r, const Fr* b, Fr* dst) { dst[k] = (s * a[k] + r * b[k]).from_mont(); }
};

int g16_slice_scalars(ps_ctx* ctx, const KeySlice& sl, const uint8_t* r_be, const uint8_t* s_be, const Fr* d_a, const Fr* d_b,
                      const Fr* d_w, size_t diff, Fr* scA, Fr* scB, Fr* scC) {
  PS_TRY(begin_call(ctx));
  ps_stream_t st = ctx->stream;
  Fr r, s;
  PS_TRY(parse_fr(r_be, &r));
  PS_TRY(parse_fr(s_be, &s));
  const Fr rs = r * s;
  const size_t nx = sl.x_hi - sl.x_lo, nn = sl.n_hi - sl.n_lo, nt = sl.t_hi - sl.t_lo;
  PS_LAUNCH(FrStdCopyK, st, nx, d_a + sl.x_lo, scA);
  PS_LAUNCH(FrStdCopyK, st, nx, d_b + sl.x_lo, scB);
  PS_LAUNCH(FrStdCopyK, st, nn, d_w + diff + sl.n_lo, scC);
  PS_LAUNCH(FrAxpbyStdK, st, nx, s, d_a + sl.x_lo, r, d_b + sl.x_lo, scC + nn + nt);
  if (sl.consts) {
    PS_LAUNCH(FrSet3K, st, 2, r.from_mont(), Fr::one().from_mont(), Fr::zero(), 2, scA + nx);
    PS_LAUNCH(FrSet3K, st, 2, s.from_mont(), Fr::one().from_mont(), Fr::zero(), 2, scB + nx);
    PS_LAUNCH(FrSet3K, st, 3, s.from_mont(), r.from_mont(), rs.from_mont(), 3, scC + nn + nt + nx);
  }
  return PS_OK;
}

// dst[k] = src[idx[k]]
struct FrGatherK {
  static constexpr int BLOCK = 256;
  PS_DEV static void run(uint32_t k, const Fr* src, const uint32_t* idx, Fr* dst) { dst[k] = src[idx[k]]; }
};
// dst[k] = c[k] in standard form, k < cnt <= 8 (constants passed by value)
struct FrConstsK {
  static constexpr int BLOCK = 32;
  struct Pack { Fr v[8]; };
  PS_DEV static void run(uint32_t k, Pack c, Fr* dst) { dst[k] = c.v[k].from_mont(); }
};

int g16_setup_scalars(ps_ctx* ctx, const ps_qap* q, const uint8_t* toxic_be, G16SetupScalars* o) {
  ps_stream_t st = ctx->stream;
  const size_t n = q->n, m = q->m, diff = q->m - q->n_io;
  Fr tox[5];   // alpha, beta, delta, x, gamma
  for (int i = 0; i < 5; i++) {
    PS_TRY(parse_fr(toxic_be + 32 * i, &tox[i]));
    if (tox[i].is_zero()) return PS_ERR_ARG;
  }
  const Fr alpha = tox[0], beta = tox[1], delta = tox[2], x = tox[3], gamma = tox[4];
  Fr* ev = ctx->arena.take<Fr>(3 * m);
  Fr* zx = ctx->arena.take<Fr>(1);
  o->pw = ctx->arena.take<Fr>(n); o->pwt = ctx->arena.take<Fr>(n); o->lp = ctx->arena.take<Fr>(m); o->consts = ctx->arena.take<Fr>(8);
  if (!ev || !zx || !o->pw || !o->pwt || !o->lp || !o->consts) return PS_ERR_ALLOC;
  PS_TRY(qap_eval_all_at(ctx, q, x, ev, zx));
  const Fr dinv = fr_inv(delta), ginv = fr_inv(gamma);
  // lp[i] = (beta u_i + alpha v_i + w_i) / gamma (i < diff) or / delta   (linearPolyForVar, groth16.go:238-264)
  PS_LAUNCH(LinCombStdK, st, m, (uint32_t)m, (const Fr*)ev, beta, alpha, Fr::one(), (uint32_t)diff, ginv, dinv, o->lp);
  PS_LAUNCH(FrPowStdK, st, n, x, Fr::one(), o->pw);                    // GeneratePowersCommit(x, 1, n-1), groth16.go:79
  // XiT: x^k t(x) / delta, k <= n - 2 (groth16.go:93-96): t(x) lives on the device
  PS_LAUNCH(FrPowTableK, st, n - 1, x, dinv, o->pwt);
  PS_LAUNCH(FrScaleByK, st, n - 1, (const Fr*)zx, o->pwt);
  PS_LAUNCH(FrFromMontK, st, n - 1, o->pwt);
  FrConstsK::Pack pk;
  pk.v[0] = alpha; pk.v[1] = beta; pk.v[2] = delta; pk.v[3] = gamma;
  for (int i = 4; i < 8; i++) pk.v[i] = Fr::zero();
  PS_LAUNCH(FrConstsK, st, 4, pk, o->consts);
  return PS_OK;
}

int phgr13_setup_scalars(ps_ctx* ctx, const ps_qap* q, const uint8_t* toxic_be, Phgr13SetupScalars* o) {
  ps_stream_t st = ctx->stream;
  const size_t n = q->n, m = q->m;
  Fr tox[8];   // s, av, aw, ay, rv, rw, beta, gamma  (sampling order of pinochio.go:93-141)
  for (int i = 0; i < 8; i++) {
    PS_TRY(parse_fr(toxic_be + 32 * i, &tox[i]));
    if (tox[i].is_zero()) return PS_ERR_ARG;
  }
  const Fr s = tox[0], av = tox[1], aw = tox[2], ay = tox[3], rv = tox[4], rw = tox[5], beta = tox[6], gamma = tox[7];
  const Fr ry = rv * rw, zero = Fr::zero();
  Fr* ev = ctx->arena.take<Fr>(3 * m);
  Fr* zx = ctx->arena.take<Fr>(1);
  o->pw = ctx->arena.take<Fr>(n);
  for (int k = 0; k < 9; k++) o->ek[k] = ctx->arena.take<Fr>(m);
  o->consts = ctx->arena.take<Fr>(8);
  if (!ev || !zx || !o->pw || !o->consts) return PS_ERR_ALLOC;
  for (int k = 0; k < 9; k++) if (!o->ek[k]) return PS_ERR_ALLOC;
  PS_TRY(qap_eval_all_at(ctx, q, s, ev, zx));
  PS_LAUNCH(FrPowStdK, st, n - 1, s, Fr::one(), o->pw);                // gsi, pinochio.go:101
  // generateEvalCommit(base r*G, polys, s, shift) = (shift r p_i(s)) * G  (pinochio.go:381-388); all m variables are
  // computed (the verification key commits to all of them, the evaluation key to the last nbIO)
  const Fr cu[9] = {rv, zero, zero, rv * av, zero, zero, rv * beta, zero, zero};   // vs ws ys vas was yas vbs wbs ybs
  const Fr cv[9] = {zero, rw, zero, zero, rw * aw, zero, zero, rw * beta, zero};
  const Fr cw[9] = {zero, zero, ry, zero, zero, ry * ay, zero, zero, ry * beta};
  for (int k = 0; k < 9; k++)
    PS_LAUNCH(LinCombStdK, st, m, (uint32_t)m, (const Fr*)ev, cu[k], cv[k], cw[k], 0u, Fr::one(), Fr::one(), o->ek[k]);
  // verification-key constants: av, aw, ay, gamma, beta*gamma, t(s)*ry
  Fr* tsy = ctx->arena.take<Fr>(1);
  if (!tsy) return PS_ERR_ALLOC;
  FrConstsK::Pack pk;
  pk.v[0] = av; pk.v[1] = aw; pk.v[2] = ay; pk.v[3] = gamma; pk.v[4] = beta * gamma; pk.v[5] = ry; pk.v[6] = zero; pk.v[7] = zero;
  PS_LAUNCH(FrConstsK, st, 6, pk, o->consts);
  // consts[5] = t(s) * ry
  PS_TRY(dev_d2d(tsy, zx, sizeof(Fr), st));
  PS_LAUNCH(FrPowTableK, st, 1, Fr::one(), ry, tsy + 0);   // tsy[0] = ry (Montgomery)
  PS_LAUNCH(FrScaleByK, st, 1, (const Fr*)zx, tsy);        // * t(s)
  PS_LAUNCH(FrStdCopyK, st, 1, (const Fr*)tsy, o->consts + 5);
  return PS_OK;
}

}  // namespace ps

extern "C" {

// ---- NTT --------------------------------------------------------------------------------------------
int ps_ntt_fr(ps_ctx* ctx, uint8_t* data_be, unsigned log_n, int inverse, const uint8_t* coset_be) {
  if (!data_be || log_n > 28) return PS_ERR_ARG;
  PS_TRY(begin_call(ctx));
  ps_stream_t st = ctx->stream;
  const size_t n = (size_t)1 << log_n;
  const NttTables* tabs = nullptr;
  PS_TRY(ctx_ntt_tables(ctx, (int)log_n, &tabs));
  uint32_t *d_x = nullptr, *d_err = nullptr;
  PS_TRY(stage_scalars(ctx, data_be, n, 1, &d_x, &d_err));
  Fr* x = (Fr*)d_x;
  Fr* y = ctx->arena.take<Fr>(n);
  Fr* pw = ctx->arena.take<Fr>(n);
  if (!y || !pw) return PS_ERR_ALLOC;
  Fr g = Fr::one();
  if (coset_be) {
    Fr t;
    for (int j = 0; j < 8; j++)
      t.v[j] = ((uint32_t)coset_be[4 * (7 - j)] << 24) | ((uint32_t)coset_be[4 * (7 - j) + 1] << 16) |
               ((uint32_t)coset_be[4 * (7 - j) + 2] << 8) | (uint32_t)coset_be[4 * (7 - j) + 3];
    if (!limbs_lt_mod<FrParams>(t.v) || t.is_zero()) return PS_ERR_ENCODING;
    g = t.to_mont();
  }
  if (!inverse) {
    if (coset_be) {
      PS_LAUNCH(FrPowTableK, st, n, g, Fr::one(), pw);
      PS_LAUNCH(FrMulTableK, st, n, x, (const Fr*)pw);
    }
    PS_TRY(ntt_forward(st, x, (int)log_n, tabs->tw));
    PS_LAUNCH(BitRevK, st, n, (const Fr*)x, y, (int)log_n);
  } else {
    PS_LAUNCH(BitRevK, st, n, (const Fr*)x, y, (int)log_n);
    PS_TRY(ntt_inverse_unscaled(st, y, (int)log_n, tabs->tw_inv));
    Fr ninv = fr_inv(fr_host_from_u64(n));
    PS_LAUNCH(FrPowTableK, st, n, fr_inv(g), ninv, pw);
    PS_LAUNCH(FrMulTableK, st, n, y, (const Fr*)pw);
  }
  PS_TRY(export_fr(ctx, y, n, data_be));
  return check_err_flag(ctx, d_err, PS_ERR_ENCODING);
}

// ---- QAP --------------------------------------------------------------------------------------------
int ps_qap_load_dense(ps_ctx* ctx, size_t n_gates, size_t n_vars, size_t n_io, const uint8_t* left,
                      const uint8_t* right, const uint8_t* out, const uint8_t* z, ps_qap** qap) {
  if (!qap || !left || !right || !out || !z || n_gates < 2 || n_vars < 1 || n_io > n_vars) return PS_ERR_ARG;
  if (n_gates > (1u << 26) || n_vars * n_gates > ((size_t)1 << 34)) return PS_ERR_UNSUPPORTED;
  PS_TRY(begin_call(ctx));
  ps_stream_t st = ctx->stream;
  ps_qap* q = new (std::nothrow) ps_qap();
  if (!q) return PS_ERR_ALLOC;
  q->n = n_gates; q->m = n_vars; q->n_io = n_io; q->dense = true;
  const size_t mn = n_vars * n_gates;
  int rc = PS_OK;
  uint32_t* d_err = ctx->arena.take<uint32_t>(1);
  uint8_t* d_bytes = ctx->arena.take<uint8_t>(mn * 32);
  Fr* d_z = ctx->arena.take<Fr>(n_gates + 1);
  if (!d_err || !d_bytes || !d_z) rc = PS_ERR_ALLOC;
  if (rc == PS_OK) rc = dev_memset(d_err, 0, 4, st);
  const uint8_t* srcs[3] = {left, right, out};
  Fr** dsts[3] = {&q->left, &q->right, &q->out};
  for (int k = 0; k < 3 && rc == PS_OK; k++) {
    rc = dev_alloc((void**)dsts[k], mn * sizeof(Fr));
    if (rc == PS_OK) rc = dev_h2d(d_bytes, srcs[k], mn * 32, st);
    if (rc == PS_OK) rc = ps_launch<FrFromBytesK>(st, mn, (const uint8_t*)d_bytes, (uint32_t*)*dsts[k], 1, d_err);
  }
  if (rc == PS_OK) rc = dev_h2d(d_bytes, z, (n_gates + 1) * 32, st);
  if (rc == PS_OK) rc = ps_launch<FrFromBytesK>(st, n_gates + 1, (const uint8_t*)d_bytes, (uint32_t*)d_z, 1, d_err);
  if (rc == PS_OK) rc = qap_prepare_tables(ctx, q, d_z);
  if (rc == PS_OK) rc = check_err_flag(ctx, d_err, PS_ERR_ENCODING);
  if (rc != PS_OK) { qap_release(q); return rc; }
  *qap = q;
  return PS_OK;
}

int ps_qap_load_r1cs(ps_ctx* ctx, size_t n_gates, size_t n_vars, size_t n_io, const uint32_t* l_row_ptr, const uint32_t* l_col,
                     const uint8_t* l_val, const uint32_t* r_row_ptr, const uint32_t* r_col, const uint8_t* r_val,
                     const uint32_t* o_row_ptr, const uint32_t* o_col, const uint8_t* o_val, ps_qap** qap) {
  if (!qap || !l_row_ptr || !r_row_ptr || !o_row_ptr || n_gates < 2 || n_vars < 1 || n_io > n_vars) return PS_ERR_ARG;
  if (n_gates > (1u << 26) || n_vars > (1u << 28)) return PS_ERR_UNSUPPORTED;
  PS_TRY(begin_call(ctx));
  ps_stream_t st = ctx->stream;
  ps_qap* q = new (std::nothrow) ps_qap();
  SparseQap* sq = new (std::nothrow) SparseQap();
  if (!q || !sq) { delete q; delete sq; return PS_ERR_ALLOC; }
  q->n = n_gates; q->m = n_vars; q->n_io = n_io; q->dense = false; q->sparse = sq;
  int k = 1;
  while (((size_t)1 << k) < n_gates) k++;
  const size_t np = (size_t)1 << k;   // leaves of the interpolation tree; gates n..np-1 are empty rows
  sq->np = (uint32_t)np;
  if (const char* e = getenv("PLAYSNARK_B200_SPMVT_SEG")) { long v = atol(e); if (v >= 1) sq->seg_len = (uint32_t)v; }   // tests: short segments
  const uint32_t* rps[3] = {l_row_ptr, r_row_ptr, o_row_ptr};
  const uint32_t* cols[3] = {l_col, r_col, o_col};
  const uint8_t* vals[3] = {l_val, r_val, o_val};
  int rc = PS_OK;
  uint32_t* d_err = ctx->arena.take<uint32_t>(1);
  if (!d_err) rc = PS_ERR_ALLOC;
  if (rc == PS_OK) rc = dev_memset(d_err, 0, 4, st);
  for (int i = 0; i < 3 && rc == PS_OK; i++) {
    size_t nnz = rps[i][n_gates];
    if (rps[i][0] != 0 || (nnz && (!cols[i] || !vals[i]))) { rc = PS_ERR_ARG; break; }
    for (size_t j = 0; j < n_gates && rc == PS_OK; j++) if (rps[i][j] > rps[i][j + 1]) rc = PS_ERR_ARG;
    for (size_t t = 0; t < nnz && rc == PS_OK; t++) if (cols[i][t] >= n_vars) rc = PS_ERR_ARG;
    if (rc != PS_OK) break;
    CsrDev& m = sq->mat[i];
    m.nnz = nnz;
    std::vector<uint32_t> rp_pad(np + 1, (uint32_t)nnz);
    std::copy(rps[i], rps[i] + n_gates + 1, rp_pad.begin());
    rc = dev_alloc((void**)&m.row_ptr, (np + 1) * 4);
    if (rc == PS_OK) rc = dev_alloc((void**)&m.col, nnz * 4);
    if (rc == PS_OK) rc = dev_alloc((void**)&m.val, nnz * sizeof(Fr));
    if (rc == PS_OK) rc = dev_h2d(m.row_ptr, rp_pad.data(), (np + 1) * 4, st);
    if (rc == PS_OK) rc = dev_sync(st);   // rp_pad goes out of scope
    if (rc == PS_OK && nnz) rc = dev_h2d(m.col, cols[i], nnz * 4, st);
    uint8_t* d_bytes = ctx->arena.take<uint8_t>(nnz * 32);
    if (rc == PS_OK && !d_bytes) rc = PS_ERR_ALLOC;
    if (rc == PS_OK && nnz) rc = dev_h2d(d_bytes, vals[i], nnz * 32, st);
    if (rc == PS_OK) rc = ps_launch<FrFromBytesK>(st, nnz, (const uint8_t*)d_bytes, (uint32_t*)m.val, 1, d_err);
  }
  // transposes (per variable: the gates that use it) for the trusted setups: counting sort on the host, values gathered
  // on the device from the already converted CSR values
  for (int i = 0; i < 3 && rc == PS_OK; i++) {
    const size_t nnz = rps[i][n_gates];
    std::vector<uint32_t> cp(n_vars + 1, 0), gate(nnz ? nnz : 1), src(nnz ? nnz : 1);
    for (size_t t = 0; t < nnz; t++) cp[cols[i][t] + 1]++;
    for (size_t v = 0; v < n_vars; v++) cp[v + 1] += cp[v];
    std::vector<uint32_t> fill(cp.begin(), cp.end() - 1);
    for (size_t j = 0; j < n_gates; j++)
      for (uint32_t t = rps[i][j]; t < rps[i][j + 1]; t++) {
        const uint32_t pos = fill[cols[i][t]]++;
        gate[pos] = (uint32_t)j; src[pos] = t;
      }
    CsrDev& mt = sq->matT[i];
    mt.nnz = nnz;
    rc = dev_alloc((void**)&mt.row_ptr, (n_vars + 1) * 4);
    if (rc == PS_OK) rc = dev_alloc((void**)&mt.col, nnz * 4);
    if (rc == PS_OK) rc = dev_alloc((void**)&mt.val, nnz * sizeof(Fr));
    uint32_t* d_src = ctx->arena.take<uint32_t>(nnz);
    if (rc == PS_OK && !d_src) rc = PS_ERR_ALLOC;
    if (rc == PS_OK) rc = dev_h2d(mt.row_ptr, cp.data(), (n_vars + 1) * 4, st);
    if (rc == PS_OK && nnz) rc = dev_h2d(mt.col, gate.data(), nnz * 4, st);
    if (rc == PS_OK && nnz) rc = dev_h2d(d_src, src.data(), nnz * 4, st);
    if (rc == PS_OK) rc = ps_launch<FrGatherK>(st, nnz, (const Fr*)sq->mat[i].val, (const uint32_t*)d_src, mt.val);
    // long rows of the transpose, cut into segments (setup.cuh: SpmvTSegK / SpmvTLongSumK)
    std::vector<uint32_t> seg_lo, seg_hi, lrow, seg_ptr(1, 0);
    for (size_t v = 0; v < n_vars; v++) {
      if (cp[v + 1] - cp[v] <= sq->seg_len) continue;
      for (uint32_t lo = cp[v]; lo < cp[v + 1]; lo += sq->seg_len) { seg_lo.push_back(lo); seg_hi.push_back(std::min(cp[v + 1], lo + sq->seg_len)); }
      lrow.push_back((uint32_t)v);
      seg_ptr.push_back((uint32_t)seg_lo.size());
    }
    LongRows& lr = sq->longT[i];
    lr.nseg = (uint32_t)seg_lo.size(); lr.nrow = (uint32_t)lrow.size();
    if (lr.nrow) {
      if (rc == PS_OK) rc = dev_alloc((void**)&lr.seg_lo, lr.nseg * 4);
      if (rc == PS_OK) rc = dev_alloc((void**)&lr.seg_hi, lr.nseg * 4);
      if (rc == PS_OK) rc = dev_alloc((void**)&lr.row, lr.nrow * 4);
      if (rc == PS_OK) rc = dev_alloc((void**)&lr.seg_ptr, (lr.nrow + 1) * 4);
      if (rc == PS_OK) rc = dev_h2d(lr.seg_lo, seg_lo.data(), lr.nseg * 4, st);
      if (rc == PS_OK) rc = dev_h2d(lr.seg_hi, seg_hi.data(), lr.nseg * 4, st);
      if (rc == PS_OK) rc = dev_h2d(lr.row, lrow.data(), lr.nrow * 4, st);
      if (rc == PS_OK) rc = dev_h2d(lr.seg_ptr, seg_ptr.data(), (lr.nrow + 1) * 4, st);
    }
    if (rc == PS_OK) rc = dev_sync(st);   // the host vectors go out of scope
  }
  Fr* d_z = ctx->arena.take<Fr>(n_gates + 1);
  if (rc == PS_OK && !d_z) rc = PS_ERR_ALLOC;
  if (rc == PS_OK) rc = inv_zprime_build(ctx, sq, (uint32_t)n_gates, (uint32_t)np);
  if (rc == PS_OK) rc = ztree_build(ctx, sq, (uint32_t)n_gates, (uint32_t)np, k, d_z);
  if (rc == PS_OK) rc = twist_tables_build(ctx, sq, (uint32_t)np, k);
  if (rc == PS_OK) rc = series_tables_build(ctx, sq, (uint32_t)n_gates, (uint32_t)np, k, d_z);
  q->log_np = k;
  if (rc == PS_OK) rc = check_err_flag(ctx, d_err, PS_ERR_ENCODING);
  if (rc != PS_OK) { ps_qap_free(q); return rc; }
  *qap = q;
  return PS_OK;
}

void ps_qap_free(ps_qap* qap) {
  if (qap && qap->sparse) { ((SparseQap*)qap->sparse)->release(); delete (SparseQap*)qap->sparse; qap->sparse = nullptr; }
  qap_release(qap);
}

int ps_quotient(ps_ctx* ctx, const ps_qap* qap, const uint8_t* witness_be, uint8_t* out_h, uint8_t* out_abc) {
  if (!qap || !witness_be || !out_h) return PS_ERR_ARG;
  PS_TRY(begin_call(ctx));
  QuotientBufs qb;
  PS_TRY(run_quotient(ctx, qap, witness_be, &qb, out_abc != nullptr));
  PS_TRY(export_fr(ctx, qb.h, qap->n - 1, out_h));
  if (out_abc) {
    PS_TRY(export_fr(ctx, qb.a, qap->n, out_abc));
    PS_TRY(export_fr(ctx, qb.b, qap->n, out_abc + qap->n * 32));
    PS_TRY(export_fr(ctx, qb.c, qap->n, out_abc + 2 * qap->n * 32));
  }
  PS_TRY(check_err_flag(ctx, qb.enc_err, PS_ERR_ENCODING));
  return check_err_flag(ctx, qb.flag, PS_ERR_REMAINDER);
}

// One aggregate polynomial of the sparse QAP in coefficient form (which = 0: a, 1: b): the SpMV and
// the interpolation on {1..n} for that polynomial only, so that two GPUs can share the work.
int ps_qap_aggregate_one(ps_ctx* ctx, const ps_qap* qap, const uint8_t* witness_be, int which, void* d_out_coef) {
  if (!ctx || !qap || !witness_be || !d_out_coef || which < 0 || which > 1) return PS_ERR_ARG;
  if (qap->dense) return PS_ERR_UNSUPPORTED;
  PS_TRY(begin_call(ctx));
  const SparseQap* sq = (const SparseQap*)qap->sparse;
  const uint32_t n = sq->np;   // leaves of the tree (the caller's buffer takes the first qap->n coefficients)
  ps_stream_t st = ctx->stream;
  uint32_t *d_w = nullptr, *d_err = nullptr;
  PS_TRY(stage_scalars(ctx, witness_be, qap->m, 1, &d_w, &d_err));
  Fr* ev = ctx->arena.take<Fr>((size_t)3 * n);
  Fr* coef = ctx->arena.take<Fr>(n);
  if (!ev || !coef) return PS_ERR_ALLOC;
  PS_LAUNCH(SpmvK, st, (size_t)3 * n, n, 0u, (const uint32_t*)sq->mat[0].row_ptr, (const uint32_t*)sq->mat[0].col, (const Fr*)sq->mat[0].val,
            (const uint32_t*)sq->mat[1].row_ptr, (const uint32_t*)sq->mat[1].col, (const Fr*)sq->mat[1].val,
            (const uint32_t*)sq->mat[2].row_ptr, (const uint32_t*)sq->mat[2].col, (const Fr*)sq->mat[2].val, (const Fr*)d_w, ev);
  PS_TRY(interpolate_ap(ctx, sq, n, qap->log_np, 1, ev + (size_t)which * n, coef));
  PS_TRY(dev_d2d(d_out_coef, coef, qap->n * sizeof(Fr), st));
  return check_err_flag(ctx, d_err, PS_ERR_ENCODING);
}

// ps_g16_scalars with the aggregate polynomials a, b already interpolated (device, n Montgomery
// coefficients each, e.g. by ps_qap_aggregate_one on two GPUs): gate check, series division, assembly.
int ps_g16_scalars_from_ab(ps_ctx* ctx, const ps_g16_key* key, const ps_qap* qap, const uint8_t* witness_be, const uint8_t* r_be,
                           const uint8_t* s_be, const void* d_a, const void* d_b, void* d_scA, void* d_scC, void* d_scB) {
  if (!key || !qap || !witness_be || !r_be || !s_be || !d_a || !d_b || !d_scA || !d_scC || !d_scB) return PS_ERR_ARG;
  if (qap->dense) return PS_ERR_UNSUPPORTED;
  if (key->n != qap->n || key->n_nio != qap->n_io) return PS_ERR_LENGTH;
  PS_TRY(begin_call(ctx));
  ps_stream_t st = ctx->stream;
  const SparseQap* sq = (const SparseQap*)qap->sparse;
  const uint32_t n = (uint32_t)qap->n, np = sq->np;
  const size_t nio = qap->n_io, diff = qap->m - qap->n_io;
  uint32_t *d_w = nullptr, *d_err = nullptr;
  PS_TRY(stage_scalars(ctx, witness_be, qap->m, 1, &d_w, &d_err));
  Fr* w = (Fr*)d_w;
  Fr* ev = ctx->arena.take<Fr>((size_t)3 * np);
  Fr* h = ctx->arena.take<Fr>(n);
  uint32_t* flag = ctx->arena.take<uint32_t>(1);
  if (!ev || !h || !flag) return PS_ERR_ALLOC;
  PS_TRY(dev_memset(flag, 0, 4, st));
  PS_LAUNCH(SpmvK, st, (size_t)3 * np, np, 0u, (const uint32_t*)sq->mat[0].row_ptr, (const uint32_t*)sq->mat[0].col, (const Fr*)sq->mat[0].val,
            (const uint32_t*)sq->mat[1].row_ptr, (const uint32_t*)sq->mat[1].col, (const Fr*)sq->mat[1].val,
            (const uint32_t*)sq->mat[2].row_ptr, (const uint32_t*)sq->mat[2].col, (const Fr*)sq->mat[2].val, (const Fr*)w, ev);
  PS_LAUNCH(GateCheckK, st, np, np, (const Fr*)ev, flag);
  const Fr* a = (const Fr*)d_a;
  const Fr* b = (const Fr*)d_b;
  PS_TRY(quotient_series(ctx, sq, n, qap->log_np, a, b, h, (Fr*)nullptr));
  Fr hrs[2];
  PS_TRY(parse_fr(r_be, &hrs[0]));
  PS_TRY(parse_fr(s_be, &hrs[1]));
  const Fr r = hrs[0], s = hrs[1], rs = hrs[0] * hrs[1];
  Fr* scA = (Fr*)d_scA; Fr* scB = (Fr*)d_scB; Fr* scC = (Fr*)d_scC;
  PS_LAUNCH(FrCopyK, st, n, a, scA);
  PS_LAUNCH(FrSet3K, st, 2, r, Fr::one(), Fr::zero(), 2, scA + n);
  PS_LAUNCH(FrCopyK, st, n, b, scB);
  PS_LAUNCH(FrSet3K, st, 2, s, Fr::one(), Fr::zero(), 2, scB + n);
  PS_LAUNCH(FrCopyK, st, nio, (const Fr*)(w + diff), scC);
  PS_LAUNCH(FrCopyK, st, (size_t)n - 1, (const Fr*)h, scC + nio);
  PS_LAUNCH(FrAxpbyK, st, n, s, a, r, b, scC + nio + (n - 1));
  PS_LAUNCH(FrSet3K, st, 3, s, r, rs, 3, scC + nio + (n - 1) + n);
  PS_LAUNCH(FrFromMontK, st, (size_t)n + 2, scA);
  PS_LAUNCH(FrFromMontK, st, (size_t)n + 2, scB);
  PS_LAUNCH(FrFromMontK, st, nio + (n - 1) + n + 3, scC);
  PS_TRY(check_err_flag(ctx, d_err, PS_ERR_ENCODING));
  return check_err_flag(ctx, flag, PS_ERR_REMAINDER);
}

int ps_qap_interp_part(ps_ctx* ctx, const ps_qap* qap, const uint8_t* witness_be, int which, size_t part, size_t parts,
                       void* d_out_evals, void* d_w_nio_out, void* d_status) {
  if (!ctx || !qap || !witness_be || !d_out_evals || !d_status || which < 0 || which > 1) return PS_ERR_ARG;
  if (qap->dense) return PS_ERR_UNSUPPORTED;
  const int lp = log2_exact(parts);
  if (lp < 0 || parts > ((size_t)1 << qap->log_np) / 2 || part >= parts) return PS_ERR_ARG;
  PS_TRY(begin_call(ctx));
  uint32_t *d_w = nullptr, *d_err = nullptr;
  PS_TRY(stage_scalars(ctx, witness_be, qap->m, 1, &d_w, &d_err));
  return interp_part_run(ctx, qap, (const Fr*)d_w, d_err, which, part, parts, lp, d_out_evals, d_w_nio_out, d_status);
}

int ps_qap_interp_part_dev(ps_ctx* ctx, const ps_qap* qap, const void* d_witness_mont, int which, size_t part, size_t parts,
                           void* d_out_evals, void* d_w_nio_out, void* d_status) {
  if (!ctx || !qap || !d_witness_mont || !d_out_evals || !d_status || which < 0 || which > 1) return PS_ERR_ARG;
  if (qap->dense) return PS_ERR_UNSUPPORTED;
  const int lp = log2_exact(parts);
  if (lp < 0 || parts > ((size_t)1 << qap->log_np) / 2 || part >= parts) return PS_ERR_ARG;
  PS_TRY(begin_call(ctx));
  return interp_part_run(ctx, qap, (const Fr*)d_witness_mont, (const uint32_t*)nullptr, which, part, parts, lp, d_out_evals,
                         d_w_nio_out, d_status);
}

int ps_fr_upload(ps_ctx* ctx, const uint8_t* values_be, size_t count, void* d_out_mont, void* d_status) {
  if (!ctx || (count && (!values_be || !d_out_mont)) || !d_status) return PS_ERR_ARG;
  PS_TRY(begin_call(ctx));
  uint8_t* d_in = ctx->arena.take<uint8_t>(count * 32);
  uint32_t* d_err = ctx->arena.take<uint32_t>(1);
  if (!d_in || !d_err) return PS_ERR_ALLOC;
  PS_TRY(dev_memset(d_err, 0, 4, ctx->stream));
  if (count) PS_TRY(dev_h2d(d_in, values_be, count * 32, ctx->stream));
  PS_LAUNCH(FrFromBytesK, ctx->stream, count, (const uint8_t*)d_in, (uint32_t*)d_out_mont, 1, d_err);
  PS_LAUNCH(StatusMergeK, ctx->stream, 1, (const uint32_t*)d_err, (const uint32_t*)nullptr, (uint32_t*)d_status);
  return PS_OK;
}

int ps_qap_interp_finish(ps_ctx* ctx, const ps_qap* qap, size_t parts, const void* d_evals_all, void* d_out_coef) {
  if (!ctx || !qap || !d_evals_all || !d_out_coef) return PS_ERR_ARG;
  if (qap->dense) return PS_ERR_UNSUPPORTED;
  const int lp = log2_exact(parts);
  if (lp < 1 || parts > ((size_t)1 << qap->log_np) / 2) return PS_ERR_ARG;
  PS_TRY(begin_call(ctx));
  const SparseQap* sq = (const SparseQap*)qap->sparse;
  const uint32_t n = sq->np;   // leaves of the tree: d_evals_all holds 2 np values, the output the first qap->n coefficients
  Fr* E0 = ctx->arena.take<Fr>((size_t)2 * n);
  Fr* coef = ctx->arena.take<Fr>(n);
  if (!E0 || !coef) return PS_ERR_ALLOC;
  PS_TRY(dev_d2d(E0, d_evals_all, (size_t)2 * n * sizeof(Fr), ctx->stream));
  PS_TRY(interpolate_levels(ctx, sq, n, qap->log_np, 1, 0, n, qap->log_np - lp, qap->log_np, E0, (Fr*)nullptr, coef));
  return dev_d2d(d_out_coef, coef, qap->n * sizeof(Fr), ctx->stream);
}

int ps_g16_h_from_ab(ps_ctx* ctx, const ps_qap* qap, const void* d_a, const void* d_b, void* d_h_out) {
  if (!ctx || !qap || !d_a || !d_b || !d_h_out) return PS_ERR_ARG;
  if (qap->dense) return PS_ERR_UNSUPPORTED;
  PS_TRY(begin_call(ctx));
  const SparseQap* sq = (const SparseQap*)qap->sparse;
  const uint32_t n = (uint32_t)qap->n;
  Fr* h = ctx->arena.take<Fr>(n);
  if (!h) return PS_ERR_ALLOC;
  PS_TRY(quotient_series(ctx, sq, n, qap->log_np, (const Fr*)d_a, (const Fr*)d_b, h, (Fr*)nullptr));
  PS_LAUNCH(FrStdCopyK, ctx->stream, (size_t)n - 1, (const Fr*)h, (Fr*)d_h_out);
  return PS_OK;
}

int ps_g16_scalars_ab(ps_ctx* ctx, const ps_g16_key* key, const uint8_t* r_be, const uint8_t* s_be, const void* d_a,
                      const void* d_b, void* d_scA, void* d_scB, void* d_scC_tail) {
  if (!ctx || !key || !r_be || !s_be || !d_a || !d_b || !d_scA || !d_scB || !d_scC_tail) return PS_ERR_ARG;
  PS_TRY(begin_call(ctx));
  ps_stream_t st = ctx->stream;
  const size_t n = key->n;
  Fr r, s;
  PS_TRY(parse_fr(r_be, &r));
  PS_TRY(parse_fr(s_be, &s));
  const Fr rs = r * s;
  const Fr* a = (const Fr*)d_a;
  const Fr* b = (const Fr*)d_b;
  Fr* scA = (Fr*)d_scA; Fr* scB = (Fr*)d_scB; Fr* tail = (Fr*)d_scC_tail;
  PS_LAUNCH(FrStdCopyK, st, n, a, scA);
  PS_LAUNCH(FrSet3K, st, 2, r.from_mont(), Fr::one().from_mont(), Fr::zero(), 2, scA + n);
  PS_LAUNCH(FrStdCopyK, st, n, b, scB);
  PS_LAUNCH(FrSet3K, st, 2, s.from_mont(), Fr::one().from_mont(), Fr::zero(), 2, scB + n);
  PS_LAUNCH(FrAxpbyK, st, n, s, a, r, b, tail);
  PS_LAUNCH(FrFromMontK, st, n, tail);
  PS_LAUNCH(FrSet3K, st, 3, s.from_mont(), r.from_mont(), rs.from_mont(), 3, tail + n);
  return PS_OK;
}

int ps_g16_scalars(ps_ctx* ctx, const ps_g16_key* key, const ps_qap* qap, const uint8_t* witness_be, const uint8_t* r_be,
                   const uint8_t* s_be, void* d_scA, void* d_scC, void* d_scB) {
  if (!key || !qap || !witness_be || !r_be || !s_be || !d_scA || !d_scC || !d_scB) return PS_ERR_ARG;
  PS_TRY(begin_call(ctx));
  G16Scalars sc;
  PS_TRY(g16_build_scalars(ctx, key, qap, witness_be, r_be, s_be, &sc));
  ps_stream_t st = ctx->stream;
  PS_LAUNCH(FrFromMontK, st, sc.nA, sc.scA);
  PS_LAUNCH(FrFromMontK, st, sc.nC, sc.scC);
  PS_LAUNCH(FrFromMontK, st, sc.nB, sc.scB);
  PS_TRY(dev_d2d(d_scA, sc.scA, sc.nA * sizeof(Fr), st));
  PS_TRY(dev_d2d(d_scC, sc.scC, sc.nC * sizeof(Fr), st));
  PS_TRY(dev_d2d(d_scB, sc.scB, sc.nB * sizeof(Fr), st));
  PS_TRY(check_err_flag(ctx, sc.qb.enc_err, PS_ERR_ENCODING));
  return check_err_flag(ctx, sc.qb.flag, PS_ERR_REMAINDER);
}

}  // extern "C"
