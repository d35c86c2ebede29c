// Trusted setups on the device: the exponents NewGroth16TrustedSetup (groth16.go:64-101, 238-264) and
// NewPHGR13TrustedSetup (pinochio.go:93-176) put into their keys, computed from the toxic waste for a
// resident QAP.  The per-variable polynomials u_i, v_i, w_i are only ever needed at ONE point x:
//   sparse QAP:  u_i(x) = sum_j L[j][i] l_j(x),  l_j(x) = z(x) / ((x - j) z'(j))   (Lagrange basis on {1..n};
//                one inversion per gate, then a transposed SpMV) -- O(nnz + n) instead of the reference's
//                m polynomial evaluations of n coefficients each;
//   dense QAP:   Horner on the stored coefficients, one thread per variable (Poly.Eval, algebra.go:107-115).
// The points themselves are made by the fixed-base kernel of the group translation units.
#pragma once
#include "interp.cuh"

namespace ps {

// out[j] = inv_zprime[j] / (x - (j + 1))                                                    (thread per gate)
struct LagrangeDenK {
  static constexpr int BLOCK = 128;
  PS_DEV static void run(uint32_t j, Fr x, const Fr* izp, Fr* out) {
    Fr v = Fr::zero();
    v.v[0] = j + 1;
    out[j] = fr_inv(x - v.to_mont()) * izp[j];
  }
};
// partial[t] = prod_{j = t, t + T, ...  < n} (x - (j + 1))                                   (thread per partial)
struct VanishPartialK {
  static constexpr int BLOCK = 128;
  PS_DEV static void run(uint32_t t, uint32_t T, uint32_t n, Fr x, Fr* partial) {
    Fr acc = Fr::one();
    for (uint32_t j = t; j < n; j += T) {
      Fr v = Fr::zero();
      v.v[0] = j + 1;
      acc = acc * (x - v.to_mont());
    }
    partial[t] = acc;
  }
};
// out[0] = prod partial[0..T)                                                                 (one thread)
struct FrProdK {
  static constexpr int BLOCK = 32;
  PS_DEV static void run(uint32_t t, uint32_t T, const Fr* partial, Fr* out) {
    if (t) return;
    Fr acc = Fr::one();
    for (uint32_t i = 0; i < T; i++) acc = acc * partial[i];
    out[0] = acc;
  }
};
// a[j] *= s[0]
struct FrScaleByK {
  static constexpr int BLOCK = 256;
  PS_DEV static void run(uint32_t j, const Fr* s, Fr* a) { a[j] = a[j] * s[0]; }
};
// ev[mat][i] = sum_k valT[k] * lag[gateT[k]] over the entries of variable i              (thread = mat * m + i)
// Variables with more than seg_len entries are left to SpmvTSegK / SpmvTLongSumK.
struct SpmvTK {
  static constexpr int BLOCK = 128;
  PS_DEV static void run(uint32_t tid, uint32_t m, uint32_t seg_len, const uint32_t* rp0, const uint32_t* g0, const Fr* v0, const uint32_t* rp1,
                         const uint32_t* g1, const Fr* v1, const uint32_t* rp2, const uint32_t* g2, const Fr* v2, const Fr* lag, Fr* ev) {
    uint32_t mat = tid / m, i = tid % m;
    const uint32_t* rp = mat == 0 ? rp0 : (mat == 1 ? rp1 : rp2);
    const uint32_t* gate = mat == 0 ? g0 : (mat == 1 ? g1 : g2);
    const Fr* val = mat == 0 ? v0 : (mat == 1 ? v1 : v2);
    if (rp[i + 1] - rp[i] > seg_len) return;
    Fr acc = Fr::zero();
    for (uint32_t k = rp[i]; k < rp[i + 1]; k++) acc = acc + val[k] * lag[gate[k]];
    ev[tid] = acc;
  }
};
// partial[s] = sum_k valT[k] * lag[gateT[k]] over segment s of a long row                       (thread per segment)
struct SpmvTSegK {
  static constexpr int BLOCK = 128;
  PS_DEV static void run(uint32_t s, const uint32_t* seg_lo, const uint32_t* seg_hi, const uint32_t* gate, const Fr* val, const Fr* lag,
                         Fr* partial) {
    Fr acc = Fr::zero();
    for (uint32_t k = seg_lo[s]; k < seg_hi[s]; k++) acc = acc + val[k] * lag[gate[k]];
    partial[s] = acc;
  }
};
// ev[row[r]] = sum of the partial sums of long row r                                            (thread per long row)
struct SpmvTLongSumK {
  static constexpr int BLOCK = 32;
  PS_DEV static void run(uint32_t r, const uint32_t* row, const uint32_t* seg_ptr, const Fr* partial, Fr* ev) {
    Fr acc = Fr::zero();
    for (uint32_t s = seg_ptr[r]; s < seg_ptr[r + 1]; s++) acc = acc + partial[s];
    ev[row[r]] = acc;
  }
};
// ev[mat][i] = polynomial i of matrix `mat` at x (dense QAP: m x n coefficients, low degree first)   (thread = mat * m + i)
struct PolyEvalK {
  static constexpr int BLOCK = 128;
  PS_DEV static void run(uint32_t tid, uint32_t m, uint32_t n, const Fr* left, const Fr* right, const Fr* out, Fr x, Fr* ev) {
    uint32_t mat = tid / m, i = tid % m;
    const Fr* p = (mat == 0 ? left : (mat == 1 ? right : out)) + (size_t)i * n;
    Fr acc = Fr::zero();
    for (uint32_t k = n; k-- > 0;) acc = acc * x + p[k];
    ev[tid] = acc;
  }
};
// out[i] = (cu * u[i] + cv * v[i] + cw * w[i]) * (i < split ? lo_scale : hi_scale), standard form     (thread per variable)
struct LinCombStdK {
  static constexpr int BLOCK = 256;
  PS_DEV static void run(uint32_t i, uint32_t m, const Fr* ev, Fr cu, Fr cv, Fr cw, uint32_t split, Fr lo_scale, Fr hi_scale, Fr* out) {
    Fr t = cu * ev[i] + cv * ev[(size_t)m + i] + cw * ev[2 * (size_t)m + i];
    out[i] = (t * (i < split ? lo_scale : hi_scale)).from_mont();
  }
};
// out[k] = scale * base^k in standard form
struct FrPowStdK {
  static constexpr int BLOCK = 256;
  PS_DEV static void run(uint32_t i, Fr base, Fr scale, Fr* out) {
    Fr acc = scale, b = base;
    uint32_t e = i;
    while (e) {
      if (e & 1) acc = acc * b;
      b = b * b;
      e >>= 1;
    }
    out[i] = acc.from_mont();
  }
};

// u_i(x), v_i(x), w_i(x) for every variable (3 * m Montgomery values into ev) and z(x) into zx[0] (device)
inline int qap_eval_all_at(ps_ctx* ctx, const ps_qap* q, const Fr& x, Fr* ev, Fr* zx) {
  ps_stream_t st = ctx->stream;
  const uint32_t n = (uint32_t)q->n, m = (uint32_t)q->m;
  const uint32_t T = 4096;
  Fr* partial = ctx->arena.take<Fr>(T);
  if (!partial) return PS_ERR_ALLOC;
  PS_LAUNCH(VanishPartialK, st, T, T, n, x, partial);
  PS_LAUNCH(FrProdK, st, 32, T, (const Fr*)partial, zx);
  if (q->dense) {
    PS_LAUNCH(PolyEvalK, st, (size_t)3 * m, m, n, (const Fr*)q->left, (const Fr*)q->right, (const Fr*)q->out, x, ev);
    return PS_OK;
  }
  const SparseQap* sq = (const SparseQap*)q->sparse;
  Fr* lag = ctx->arena.take<Fr>(n);
  if (!lag) return PS_ERR_ALLOC;
  PS_LAUNCH(LagrangeDenK, st, n, x, (const Fr*)sq->inv_zprime, lag);
  PS_LAUNCH(FrScaleByK, st, n, (const Fr*)zx, lag);
  PS_LAUNCH(SpmvTK, st, (size_t)3 * m, m, sq->seg_len, (const uint32_t*)sq->matT[0].row_ptr, (const uint32_t*)sq->matT[0].col, (const Fr*)sq->matT[0].val,
            (const uint32_t*)sq->matT[1].row_ptr, (const uint32_t*)sq->matT[1].col, (const Fr*)sq->matT[1].val,
            (const uint32_t*)sq->matT[2].row_ptr, (const uint32_t*)sq->matT[2].col, (const Fr*)sq->matT[2].val, (const Fr*)lag, ev);
  for (int i = 0; i < 3; i++) {
    const LongRows& lr = sq->longT[i];
    if (!lr.nrow) continue;
    Fr* partial_sums = ctx->arena.take<Fr>(lr.nseg);
    if (!partial_sums) return PS_ERR_ALLOC;
    PS_LAUNCH(SpmvTSegK, st, lr.nseg, (const uint32_t*)lr.seg_lo, (const uint32_t*)lr.seg_hi, (const uint32_t*)sq->matT[i].col,
              (const Fr*)sq->matT[i].val, (const Fr*)lag, partial_sums);
    PS_LAUNCH(SpmvTLongSumK, st, lr.nrow, (const uint32_t*)lr.row, (const uint32_t*)lr.seg_ptr, (const Fr*)partial_sums, ev + (size_t)i * m);
  }
  return PS_OK;
}

}  // namespace ps
