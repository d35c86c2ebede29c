// Sparse-R1CS form of the QAP for sizes where the dense polynomials cannot exist (3*m*n*32 bytes):
// the per-variable polynomials of ToQAP (qap.go:35-93) stay implicit as interpolants on the domain
// {1..n}; computeAggregatePoly (qap.go:164-175) becomes
//   1. SpmvK          evaluations a(j) = <L_j, w>, b(j), c(j) at the gates j = 1..n
//   2. GateCheckK     a(j) b(j) == c(j) for every gate  <=>  z divides a b - c  (the "apocalypse" test,
//                     qap.go:158-160, exact: z's roots are exactly the gates)
//   3. interpolation on {1..n}: the polynomial through (j, y_j) is  sum_j w_j prod_{i != j} (x - i)
//      with w_j = y_j / z'(j), z'(j) = (-1)^(n-j) (j-1)! (n-j)!; the sum is folded up a subproduct
//      tree, N_parent = N_L Z_R + N_R Z_L, with batched NTT products.  The Z-tree depends only on n:
//      it is built once per QAP and kept resident in its NTT domain (log2(n) * 2n Fr: 1.3 GB at
//      n = 2^20 -- HBM is plentiful); its root is z(x) (qap.go:41-55).
// The coefficient vectors are identical to Interpolate's (algebra.go:254-280): the interpolant is
// unique.  n must be a power of two in this path.
#pragma once
#include "poly.cuh"

namespace ps {

struct CsrDev {
  uint32_t* row_ptr = nullptr;  // n + 1
  uint32_t* col = nullptr;      // nnz
  Fr* val = nullptr;            // nnz, Montgomery
  size_t nnz = 0;
};

struct SparseQap {
  CsrDev mat[3];                 // left, right, out
  Fr* inv_zprime = nullptr;      // 1 / z'(j), j = 1..n
  std::vector<Fr*> ztree;        // level l: NTT_{2s} of every node's Z (s = 2^l), n/s nodes x 2s
  void release() {
    for (auto& m : mat) { dev_free(m.row_ptr); dev_free(m.col); dev_free(m.val); }
    dev_free(inv_zprime);
    for (auto* p : ztree) dev_free(p);
    ztree.clear();
  }
};

// ev[mat][j] = sum_k val[k] * w[col[k]] over row j          (thread = mat * n + j)
struct SpmvK {
  static constexpr int BLOCK = 128;
  PS_DEV static void run(uint32_t tid, uint32_t n, const uint32_t* rp0, const uint32_t* c0, const Fr* v0, const uint32_t* rp1,
                         const uint32_t* c1, const Fr* v1, const uint32_t* rp2, const uint32_t* c2, const Fr* v2, const Fr* w,
                         Fr* ev) {
    uint32_t m = tid / n, j = tid % n;
    const uint32_t* rp = m == 0 ? rp0 : (m == 1 ? rp1 : rp2);
    const uint32_t* col = m == 0 ? c0 : (m == 1 ? c1 : c2);
    const Fr* val = m == 0 ? v0 : (m == 1 ? v1 : v2);
    Fr acc = Fr::zero();
    for (uint32_t k = rp[j]; k < rp[j + 1]; k++) acc = acc + val[k] * w[col[k]];
    ev[tid] = acc;
  }
};
// flag |= a(j) b(j) != c(j);  ev = [a | b | c]
struct GateCheckK {
  static constexpr int BLOCK = 256;
  PS_DEV static void run(uint32_t j, uint32_t n, const Fr* ev, uint32_t* flag) {
    if ((ev[j] * ev[n + j]) != ev[2 * (size_t)n + j]) ps_atomic_or(flag, 1u);
  }
};
// w[p][j] = ev[p][j] * inv_zprime[j]
struct InterpWeightK {
  static constexpr int BLOCK = 256;
  PS_DEV static void run(uint32_t idx, uint32_t n, const Fr* ev, const Fr* izp, Fr* w) { w[idx] = ev[idx] * izp[idx % n]; }
};
// leaves of the Z-tree: node i is (x - (i+1)) -> [-(i+1), 1]
struct TreeLeafK {
  static constexpr int BLOCK = 256;
  PS_DEV static void run(uint32_t i, Fr* W) {
    Fr v = Fr::zero();
    v.v[0] = i + 1;
    W[2 * (size_t)i] = v.to_mont().neg();
    W[2 * (size_t)i + 1] = Fr::one();
  }
};
// T[p][e] = Zhat[2p][e] * Zhat[2p+1][e] * scale, e < 2s                          (thread over n)
struct TreeZMulK {
  static constexpr int BLOCK = 256;
  PS_DEV static void run(uint32_t idx, uint32_t two_s, const Fr* Zhat, Fr scale, Fr* T) {
    uint32_t p = idx / two_s, e = idx % two_s;
    const Fr* L = Zhat + (size_t)(2 * p) * two_s;
    T[idx] = L[e] * L[two_s + e] * scale;
  }
};
// cyclic product coefficients c[p][0..2s) of a monic degree-2s polynomial: c0 has the wrapped leading 1.
// Writes W[p][0..4s) = [c0 - 1, c1, ..., c_{2s-1}, 1, 0, ...]                      (thread over 2n)
struct TreeZLiftK {
  static constexpr int BLOCK = 256;
  PS_DEV static void run(uint32_t idx, uint32_t two_s, const Fr* T, Fr* W) {
    uint32_t four_s = 2 * two_s;
    uint32_t p = idx / four_s, e = idx % four_s;
    Fr v = Fr::zero();
    if (e < two_s) { v = T[(size_t)p * two_s + e]; if (e == 0) v = v - Fr::one(); }
    else if (e == two_s) v = Fr::one();
    W[idx] = v;
  }
};
// root: z[0..n) = c - [1,0,...], z[n] = 1                                         (thread over n+1)
struct TreeRootK {
  static constexpr int BLOCK = 256;
  PS_DEV static void run(uint32_t e, uint32_t n, const Fr* T, Fr* z) {
    if (e == n) { z[e] = Fr::one(); return; }
    Fr v = T[e];
    if (e == 0) v = v - Fr::one();
    z[e] = v;
  }
};
// W[poly][node][0..2s) = [N[poly][node][0..s), 0...]                              (thread over P*2n)
struct InterpPadK {
  static constexpr int BLOCK = 256;
  PS_DEV static void run(uint32_t idx, uint32_t n, uint32_t s, const Fr* N, Fr* W) {
    uint32_t poly = idx / (2 * n), r = idx % (2 * n);
    uint32_t node = r / (2 * s), e = r % (2 * s);
    W[idx] = e < s ? N[(size_t)poly * n + (size_t)node * s + e] : Fr::zero();
  }
};
// O[poly][p][e] = (W[poly][2p][e] Zhat[2p+1][e] + W[poly][2p+1][e] Zhat[2p][e]) * inv2s   (thread over P*n)
struct InterpCombineK {
  static constexpr int BLOCK = 256;
  PS_DEV static void run(uint32_t idx, uint32_t n, uint32_t two_s, const Fr* W, const Fr* Zhat, Fr inv2s, Fr* O) {
    uint32_t poly = idx / n, r = idx % n;
    uint32_t p = r / two_s, e = r % two_s;
    const Fr* Wl = W + (size_t)poly * 2 * n + (size_t)(2 * p) * two_s;
    const Fr* Zl = Zhat + (size_t)(2 * p) * two_s;
    O[idx] = (Wl[e] * Zl[two_s + e] + Wl[two_s + e] * Zl[e]) * inv2s;
  }
};

// Builds the Z-tree for n = 2^k (device), writes z (n+1 coefficients, Montgomery) to d_z.
inline int ztree_build(ps_ctx* ctx, SparseQap* sq, uint32_t n, int k, Fr* d_z) {
  ps_stream_t st = ctx->stream;
  const NttTables* tabs = nullptr;
  PS_TRY(ctx_ntt_tables(ctx, k + 1, &tabs));
  const uint32_t n_tw = 2 * n;
  Fr* T = ctx->arena.take<Fr>(n);
  if (!T) return PS_ERR_ALLOC;
  sq->ztree.assign(k, nullptr);
  for (int l = 0; l < k; l++) PS_TRY(dev_alloc((void**)&sq->ztree[l], (size_t)2 * n * sizeof(Fr)));
  PS_LAUNCH(TreeLeafK, st, n, sq->ztree[0]);
  PS_TRY(ntt_forward_blocks(st, sq->ztree[0], (size_t)2 * n, 1, tabs->tw, n_tw));
  for (int l = 0; l < k; l++) {
    const uint32_t two_s = 2u << l;
    Fr scale = fr_inv(fr_host_from_u64(two_s));
    PS_LAUNCH(TreeZMulK, st, n, two_s, (const Fr*)sq->ztree[l], scale, T);
    PS_TRY(ntt_inverse_blocks_unscaled(st, T, n, l + 1, tabs->tw_inv, n_tw));
    if (l + 1 == k) {
      PS_LAUNCH(TreeRootK, st, (size_t)n + 1, n, (const Fr*)T, d_z);
    } else {
      PS_LAUNCH(TreeZLiftK, st, (size_t)2 * n, two_s, (const Fr*)T, sq->ztree[l + 1]);
      PS_TRY(ntt_forward_blocks(st, sq->ztree[l + 1], (size_t)2 * n, l + 2, tabs->tw, n_tw));
    }
  }
  return PS_OK;
}

// 1 / z'(j) for j = 1..n, computed on the host (once per QAP) and uploaded
inline int inv_zprime_build(ps_ctx* ctx, SparseQap* sq, uint32_t n) {
  std::vector<Fr> fact(n + 1), invfact(n + 1), out(n);
  fact[0] = Fr::one();
  for (uint32_t i = 1; i <= n; i++) fact[i] = fact[i - 1] * fr_host_from_u64(i);
  invfact[n] = fr_inv(fact[n]);
  for (uint32_t i = n; i > 0; i--) invfact[i - 1] = invfact[i] * fr_host_from_u64(i);
  for (uint32_t j = 1; j <= n; j++) {
    Fr v = invfact[j - 1] * invfact[n - j];
    out[j - 1] = ((n - j) & 1) ? v.neg() : v;
  }
  PS_TRY(dev_alloc((void**)&sq->inv_zprime, (size_t)n * sizeof(Fr)));
  PS_TRY(dev_h2d(sq->inv_zprime, out.data(), (size_t)n * sizeof(Fr), ctx->stream));
  return dev_sync(ctx->stream);
}

// ev: [a | b | c] evaluations (3n) -> coef: [a | b | c] coefficients (3n).  Scratch from the arena.
inline int interpolate3(ps_ctx* ctx, const SparseQap* sq, uint32_t n, int k, const Fr* ev, Fr* coef) {
  ps_stream_t st = ctx->stream;
  const NttTables* tabs = nullptr;
  PS_TRY(ctx_ntt_tables(ctx, k + 1, &tabs));
  const uint32_t n_tw = 2 * n;
  const size_t P = 3;
  Fr* W = ctx->arena.take<Fr>(P * 2 * n);
  Fr* N = ctx->arena.take<Fr>(P * n);
  if (!W || !N) return PS_ERR_ALLOC;
  PS_LAUNCH(InterpWeightK, st, P * n, n, ev, (const Fr*)sq->inv_zprime, N);
  for (int l = 0; l < k; l++) {
    const uint32_t s = 1u << l, two_s = 2 * s;
    Fr inv2s = fr_inv(fr_host_from_u64(two_s));
    Fr* dst = (l + 1 == k) ? coef : N;
    PS_LAUNCH(InterpPadK, st, P * 2 * n, n, s, (const Fr*)N, W);
    PS_TRY(ntt_forward_blocks(st, W, P * 2 * n, l + 1, tabs->tw, n_tw));
    PS_LAUNCH(InterpCombineK, st, P * n, n, two_s, (const Fr*)W, (const Fr*)sq->ztree[l], inv2s, dst);
    PS_TRY(ntt_inverse_blocks_unscaled(st, dst, P * n, l + 1, tabs->tw_inv, n_tw));
  }
  return PS_OK;
}

}  // namespace ps
