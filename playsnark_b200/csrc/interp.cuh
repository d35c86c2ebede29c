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
//      Each level reuses the evaluations it has just produced as the even half of the next level's
//      (twice as large) transform and only computes the odd half (coefficients * omega_{4s}^t).
//   4. h = floor(a b / z) (= (a b - c)/z since c = a b mod z): the top n-1 coefficients of a b,
//      reversed, times the power-series inverse of rev(z) (precomputed per QAP by Newton iteration)
//      -- so c never has to be interpolated.
// The coefficient vectors are identical to Interpolate's (algebra.go:254-280) and Div2's
// (algebra.go:140-159): interpolant and quotient are unique.
// Any n >= 2: the tree is built over np = 2^k >= n leaves; the leaves above n are DUMMIES with node polynomial 1
// and weight 0 (so every node is the product over its real leaves only, the root is z of degree n, and the
// interpolant has degree < n with exact zeros above).  The matrices are padded with empty rows up to np, so
// every per-gate kernel simply runs over np gates; only the factorials in 1/z'(j), the leaf / lift / root
// kernels of the Z-tree and the reversals of the series division know the real n.
#pragma once
#include "poly.cuh"

namespace ps {

struct CsrDev {
  uint32_t* row_ptr = nullptr;  // n + 1
  uint32_t* col = nullptr;      // nnz
  Fr* val = nullptr;            // nnz, Montgomery
  size_t nnz = 0;
};

// Variables that occur in more than seg_len gates (the constant 1, a public input wired everywhere): their sums in the
// transposed SpMV are cut into segments of seg_len entries -- a thread per segment, then a thread per variable over
// its partial sums -- instead of one thread walking 2^20 entries (0.53 s at 2^20 gates, a third of the device setup).
struct LongRows {
  uint32_t* seg_lo = nullptr;   // nseg: first entry of the segment in matT.col / matT.val
  uint32_t* seg_hi = nullptr;   // nseg: one past its last entry
  uint32_t* row = nullptr;      // nrow: the variable
  uint32_t* seg_ptr = nullptr;  // nrow + 1: its segments
  uint32_t nseg = 0, nrow = 0;
};

struct SparseQap {
  CsrDev mat[3];                 // left, right, out
  CsrDev matT[3];                // their transposes (row_ptr over the variables, col = gate): the trusted setups sum over gates
  LongRows longT[3];             // the long rows of the transposes
  uint32_t seg_len = 512;
  Fr* inv_zprime = nullptr;      // 1 / z'(j), j = 1..n
  std::vector<Fr*> ztree;        // level l: NTT_{2s} of every node's Z (s = 2^l), n/s nodes x 2s
  Fr* twist = nullptr;           // level l at offset 2s - 2 (s = 2^l): (1/2s) * omega_{4s}^t, t < 2s; l < k - 1
  Fr* s_hat = nullptr;           // NTT_{2n} of the series inverse of rev(z) mod x^(n-1)   (bit-reversed)
  Fr* z_hat = nullptr;           // NTT_{2n} of z                                          (bit-reversed)
  uint32_t np = 0;               // leaves of the tree: the power of two >= n (all the sizes "n" above are np)
  void release() {
    for (auto& m : mat) { dev_free(m.row_ptr); dev_free(m.col); dev_free(m.val); }
    for (auto& m : matT) { dev_free(m.row_ptr); dev_free(m.col); dev_free(m.val); }
    for (auto& l : longT) { dev_free(l.seg_lo); dev_free(l.seg_hi); dev_free(l.row); dev_free(l.seg_ptr); }
    dev_free(inv_zprime); dev_free(s_hat); dev_free(z_hat); dev_free(twist);
    for (auto* p : ztree) dev_free(p);
    ztree.clear();
  }
};

// ev[mat][j] = sum_k val[k] * w[col[k]] over row row0 + j, j < n          (thread = mat * n + j)
struct SpmvK {
  static constexpr int BLOCK = 128;
  PS_DEV static void run(uint32_t tid, uint32_t n, uint32_t row0, const uint32_t* rp0, const uint32_t* c0, const Fr* v0, const uint32_t* rp1,
                         const uint32_t* c1, const Fr* v1, const uint32_t* rp2, const uint32_t* c2, const Fr* v2, const Fr* w,
                         Fr* ev) {
    uint32_t m = tid / n, j = row0 + tid % n;
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
// leaves of the Z-tree: node i < n is (x - (i+1)) -> [-(i+1), 1]; the dummy leaves above are the polynomial 1
struct TreeLeafK {
  static constexpr int BLOCK = 256;
  PS_DEV static void run(uint32_t i, uint32_t n, Fr* W) {
    if (i >= n) { W[2 * (size_t)i] = Fr::one(); W[2 * (size_t)i + 1] = Fr::zero(); return; }
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
// cyclic product coefficients c[p][0..2s) of a node polynomial of degree <= 2s.  A FULL node (all its 2s leaves are
// real: p < n_full) is monic of degree exactly 2s and c0 carries the wrapped leading 1:
//   W[p][0..4s) = [c0 - 1, c1, ..., c_{2s-1}, 1, 0, ...];
// a node with dummy leaves has degree < 2s, nothing wraps: W[p] = [c0, ..., c_{2s-1}, 0, ...]       (thread over 2 np)
struct TreeZLiftK {
  static constexpr int BLOCK = 256;
  PS_DEV static void run(uint32_t idx, uint32_t two_s, uint32_t n_full, const Fr* T, Fr* W) {
    uint32_t four_s = 2 * two_s;
    uint32_t p = idx / four_s, e = idx % four_s;
    const bool full = p < n_full;
    Fr v = Fr::zero();
    if (e < two_s) { v = T[(size_t)p * two_s + e]; if (e == 0 && full) v = v - Fr::one(); }
    else if (e == two_s && full) v = Fr::one();
    W[idx] = v;
  }
};
// root: n == np: z[0..n) = c - [1,0,...], z[n] = 1; n < np: the degree-n product did not wrap, z[0..n] = c[0..n]
//                                                                                  (thread over n+1)
struct TreeRootK {
  static constexpr int BLOCK = 256;
  PS_DEV static void run(uint32_t e, uint32_t n, uint32_t np, const Fr* T, Fr* z) {
    if (n < np) { z[e] = T[e]; return; }
    if (e == n) { z[e] = Fr::one(); return; }
    Fr v = T[e];
    if (e == 0) v = v - Fr::one();
    z[e] = v;
  }
};
// Builds the Z-tree over np = 2^k leaves, the first n of them real (device); writes z (n+1 coefficients, Montgomery) to d_z.
inline int ztree_build(ps_ctx* ctx, SparseQap* sq, uint32_t n, uint32_t np, int k, Fr* d_z) {
  ps_stream_t st = ctx->stream;
  const NttTables* tabs = nullptr;
  PS_TRY(ctx_ntt_tables(ctx, k + 1, &tabs));
  const uint32_t n_tw = 2 * np;
  Fr* T = ctx->arena.take<Fr>(np);
  if (!T) return PS_ERR_ALLOC;
  sq->ztree.assign(k, nullptr);
  for (int l = 0; l < k; l++) PS_TRY(dev_alloc((void**)&sq->ztree[l], (size_t)2 * np * sizeof(Fr)));
  PS_LAUNCH(TreeLeafK, st, np, n, sq->ztree[0]);
  PS_TRY(ntt_forward_blocks(st, sq->ztree[0], (size_t)2 * np, 1, tabs->tw, n_tw));
  for (int l = 0; l < k; l++) {
    const uint32_t two_s = 2u << l;
    Fr scale = fr_inv(fr_host_from_u64(two_s));
    PS_LAUNCH(TreeZMulK, st, np, two_s, (const Fr*)sq->ztree[l], scale, T);
    PS_TRY(ntt_inverse_blocks_unscaled(st, T, np, l + 1, tabs->tw_inv, n_tw));
    if (l + 1 == k) {
      PS_LAUNCH(TreeRootK, st, (size_t)n + 1, n, np, (const Fr*)T, d_z);
    } else {
      PS_LAUNCH(TreeZLiftK, st, (size_t)2 * np, two_s, n / two_s, (const Fr*)T, sq->ztree[l + 1]);
      PS_TRY(ntt_forward_blocks(st, sq->ztree[l + 1], (size_t)2 * np, l + 2, tabs->tw, n_tw));
    }
  }
  return PS_OK;
}

// 1 / z'(j) for j = 1..n (zero for the dummy leaves n < j <= np), computed on the host (once per QAP) and uploaded
inline int inv_zprime_build(ps_ctx* ctx, SparseQap* sq, uint32_t n, uint32_t np) {
  std::vector<Fr> fact(n + 1), invfact(n + 1), out(np, Fr::zero());
  fact[0] = Fr::one();
  for (uint32_t i = 1; i <= n; i++) fact[i] = fact[i - 1] * fr_host_from_u64(i);
  invfact[n] = fr_inv(fact[n]);
  for (uint32_t i = n; i > 0; i--) invfact[i - 1] = invfact[i] * fr_host_from_u64(i);
  for (uint32_t j = 1; j <= n; j++) {
    Fr v = invfact[j - 1] * invfact[n - j];
    out[j - 1] = ((n - j) & 1) ? v.neg() : v;
  }
  PS_TRY(dev_alloc((void**)&sq->inv_zprime, (size_t)np * sizeof(Fr)));
  PS_TRY(dev_h2d(sq->inv_zprime, out.data(), (size_t)np * sizeof(Fr), ctx->stream));
  return dev_sync(ctx->stream);
}

// T[t] = scale * tw[t * step]                                                          (thread over 2s)
struct TwistTableK {
  static constexpr int BLOCK = 256;
  PS_DEV static void run(uint32_t t, uint32_t step, Fr scale, const Fr* tw, Fr* T) { T[t] = scale * fe_ld(tw + (size_t)t * step); }
};
// per-level twist factors of the interpolation tree: turning the coefficients of a node polynomial
// (degree < 2s) into its values on the odd 4s-th roots of unity, with the 1/2s of the inverse transform
inline int twist_tables_build(ps_ctx* ctx, SparseQap* sq, uint32_t n, int k) {
  const NttTables* tabs = nullptr;
  PS_TRY(ctx_ntt_tables(ctx, k + 1, &tabs));
  const uint32_t n_tw = 2 * n;
  PS_TRY(dev_alloc((void**)&sq->twist, (size_t)(n > 2 ? n : 2) * sizeof(Fr)));
  for (int l = 0; l + 1 < k; l++) {
    const uint32_t two_s = 2u << l;
    PS_LAUNCH(TwistTableK, ctx->stream, two_s, n_tw / (2 * two_s), fr_inv(fr_host_from_u64(two_s)), (const Fr*)tabs->tw,
              sq->twist + (two_s - 2));
  }
  return PS_OK;
}

// E0[poly][j][0..2) = [w_j, w_j]: the size-2 transform of the constant polynomial w_j   (thread over P*n)
struct InterpLeafK {
  static constexpr int BLOCK = 256;
  PS_DEV static void run(uint32_t idx, uint32_t n, const Fr* ev, const Fr* izp, Fr* E) {
    Fr w = ev[idx] * izp[idx % n];
    E[2 * (size_t)idx] = w;
    E[2 * (size_t)idx + 1] = w;
  }
};
// O = E_L Zhat_R + E_R Zhat_L per parent (2s evaluations).  Written twice: into the even half of the
// parent's next-level block (Enext, 4s per parent) and into the contiguous scratch C.   (thread over P*n)
struct InterpCombine2K {
  static constexpr int BLOCK = 256;
  PS_DEV static void run(uint32_t idx, uint32_t n, uint32_t two_s, const Fr* E, const Fr* Zhat, Fr* Enext, Fr* Cbuf) {
    uint32_t poly = idx / n, r = idx % n;
    uint32_t p = r / two_s, e = r % two_s;
    const Fr* El = E + (size_t)poly * 2 * n + (size_t)(2 * p) * two_s;
    const Fr* Zl = Zhat + (size_t)(2 * p) * two_s;
    Fr o = El[e] * Zl[two_s + e] + El[two_s + e] * Zl[e];
    if (Enext) Enext[(size_t)poly * 2 * n + (size_t)p * 2 * two_s + e] = o;
    Cbuf[idx] = o;
  }
};
// C[t] *= scale * omega_{4s}^t  (t = position inside the 2s-block);  tw = omega_{n_tw}^i   (thread over P*n)
struct InterpTwistK {
  static constexpr int BLOCK = 256;
  PS_DEV static void run(uint32_t idx, uint32_t two_s, uint32_t tw_step, Fr scale, const Fr* tw, Fr* Cbuf) {
    uint32_t t = idx % two_s;
    Cbuf[idx] = Cbuf[idx] * scale * fe_ld(tw + (size_t)t * tw_step);
  }
};
// odd halves: Enext[poly][p][2s + e] = C[poly][p][e]                                     (thread over P*n)
struct InterpOddK {
  static constexpr int BLOCK = 256;
  PS_DEV static void run(uint32_t idx, uint32_t n, uint32_t two_s, const Fr* Cbuf, Fr* Enext) {
    uint32_t poly = idx / n, r = idx % n;
    uint32_t p = r / two_s, e = r % two_s;
    Enext[(size_t)poly * 2 * n + (size_t)p * 2 * two_s + two_s + e] = Cbuf[idx];
  }
};
// a[i] *= scale
struct FrScaleK {
  static constexpr int BLOCK = 256;
  PS_DEV static void run(uint32_t i, Fr scale, Fr* a) { a[i] = a[i] * scale; }
};

// ---- the lowest levels of the tree in ONE kernel --------------------------------------------------------------
// Levels 0 .. LOW_LEVELS-1 work on nodes of at most LOW_G = 2^LOW_LEVELS leaves: a thread block keeps the 2 LOW_G
// evaluations of its group of leaves in shared memory and runs leaf weights, combines, inverse transforms, twists and
// forward transforms of all those levels without leaving the SM; only the Z-tree slices stream in from HBM.  Per
// launch-bound subtree (a rank of the sharded prover works on n/parts gates; 2^16-gate circuits) this replaces ~50
// grid launches of a few microseconds each; the arithmetic is the same (same twiddle tables, same order), so the
// coefficients are bit-identical to the level-by-level path, which host emulation keeps using.
constexpr int LOW_LEVELS = 9;
constexpr uint32_t LOW_G = 1u << LOW_LEVELS;
struct LowTreePtrs { const Fr* z[LOW_LEVELS]; };
#if PS_GPU
static __global__ void __launch_bounds__(256) k_interp_low(uint32_t ns, uint32_t lo, const Fr* ev, const Fr* izp, LowTreePtrs zt,
                                                           const Fr* twist, const Fr* tw, const Fr* tw_inv, uint32_t n_tw, Fr* Eout) {
  extern __shared__ uint4 sm_raw[];
  Fr* E = reinterpret_cast<Fr*>(sm_raw);   // 2 LOW_G evaluations (children, then parents, in place)
  Fr* Cb = E + 2 * LOW_G;                  // LOW_G values: one transform per parent
  const uint32_t groups = ns / LOW_G;
  const uint32_t poly = blockIdx.x / groups, grp = blockIdx.x % groups;
  const uint32_t leaf0 = grp * LOW_G;
  const uint32_t T = blockDim.x, tid = threadIdx.x;
  for (uint32_t i = tid; i < LOW_G; i += T) {
    Fr w = ev[(size_t)poly * ns + leaf0 + i] * izp[leaf0 + i];
    E[2 * i] = w; E[2 * i + 1] = w;
  }
  __syncthreads();
#pragma unroll 1
  for (int l = 0; l < LOW_LEVELS; l++) {
    const uint32_t two_s = 2u << l;
    const Fr* Zhat = zt.z[l] + 2 * (size_t)(lo + leaf0);
    // combine: parent p occupies the range of its two children; every thread overwrites what only it has read
    for (uint32_t idx = tid; idx < LOW_G; idx += T) {
      const uint32_t p = idx / two_s, e = idx % two_s;
      Fr* El = E + (size_t)(2 * p) * two_s;
      const Fr* Zl = Zhat + (size_t)(2 * p) * two_s;
      Fr o = El[e] * fe_ld(Zl + two_s + e) + El[two_s + e] * fe_ld(Zl + e);
      El[e] = o;
      Cb[idx] = o;
    }
    __syncthreads();
    // inverse transforms (DIT, bit-reversed -> natural), all parents at once: blocks of two_s are multiples of 2h
    for (uint32_t h = 1; h < two_s; h <<= 1) {
      for (uint32_t b = tid; b < LOW_G / 2; b += T) {
        const uint32_t j = b % h, i0 = (b / h) * 2 * h + j;
        Fr u = Cb[i0], v = Cb[i0 + h];
        if (j) v = v * ntt_twiddle(tw_inv, n_tw, 2 * h, j);
        Cb[i0] = u + v; Cb[i0 + h] = u - v;
      }
      __syncthreads();
    }
    // twist: coefficient t times (1 / two_s) omega_{2 two_s}^t
    const Fr* tws = twist + (two_s - 2);
    for (uint32_t idx = tid; idx < LOW_G; idx += T) Cb[idx] = Cb[idx] * fe_ld(tws + (idx & (two_s - 1)));
    __syncthreads();
    // forward transforms (DIF, natural -> bit-reversed)
    for (uint32_t h = two_s / 2; h >= 1; h >>= 1) {
      for (uint32_t b = tid; b < LOW_G / 2; b += T) {
        const uint32_t j = b % h, i0 = (b / h) * 2 * h + j;
        Fr u = Cb[i0], v = Cb[i0 + h];
        Cb[i0] = u + v;
        Fr d = u - v;
        Cb[i0 + h] = j ? d * ntt_twiddle(tw, n_tw, 2 * h, j) : d;
      }
      __syncthreads();
    }
    // odd halves of the parents' next-level blocks
    for (uint32_t idx = tid; idx < LOW_G; idx += T) {
      const uint32_t p = idx / two_s, e = idx % two_s;
      E[(size_t)p * 2 * two_s + two_s + e] = Cb[idx];
    }
    __syncthreads();
  }
  Fr* out = Eout + (size_t)poly * 2 * ns + (size_t)grp * 2 * LOW_G;
  for (uint32_t i = tid; i < 2 * LOW_G; i += T) out[i] = E[i];
}
#endif

// Levels [l0, l1) of the interpolation tree over the leaves [lo, lo + ns) of the domain {1..n} (ns a power
// of two that divides lo; the whole tree is lo = 0, ns = n, l0 = 0, l1 = k).  E0 holds the evaluations
// entering level l0: P polynomials x ns/s nodes x 2s values (s = 2^l0); it is used as scratch.
// l1 == k (only with ns == n): `coef` receives the P * n coefficients.  l1 < k: `Eout` receives the
// evaluations entering level l1 (P * 2 ns values), e.g. to be gathered from several GPUs.
inline int interpolate_levels(ps_ctx* ctx, const SparseQap* sq, uint32_t n, int k, int P, uint32_t lo, uint32_t ns, int l0,
                              int l1, Fr* E0, Fr* Eout, Fr* coef) {
  ps_stream_t st = ctx->stream;
  if (l0 < 0 || l1 > k || l0 > l1 || (l1 == k && (ns != n || !coef)) || (l1 < k && !Eout)) return PS_ERR_ARG;
  const NttTables* tabs = nullptr;
  PS_TRY(ctx_ntt_tables(ctx, k + 1, &tabs));
  const uint32_t n_tw = 2 * n;
  Fr* E[2] = {E0, ctx->arena.take<Fr>((size_t)P * 2 * ns)};
  Fr* Cbuf = ctx->arena.take<Fr>((size_t)P * ns);
  if (!E[1] || !Cbuf) return PS_ERR_ALLOC;
  int cur = 0;
  for (int l = l0; l < l1; l++) {
    const uint32_t two_s = 2u << l;
    const bool last = (l + 1 == k);
    Fr* dst = last ? coef : Cbuf;
    // The combine O = E_L Zhat_R + E_R Zhat_L is its own (perfectly coalesced) pass: fused into the
    // loads of the inverse transform's first pass it was measured slower (2^20: quotient 21.1 vs 17.6 ms;
    // that pass reads thread-contiguous runs of 8 elements, and the combine multiplies the badly
    // coalesced streams by five).  The other element-wise steps ride on the transforms' last passes
    // (NttIO), whose accesses are coalesced:
    //   twist    coefficient t times (1/2s) omega_{4s}^t while the inverse transform stores its output,
    //   the forward transform stores straight into the odd half of the parent's block.
    NttIO io;
    io.ns = ns;
    io.two_s = two_s;
    io.twist = sq->twist + (two_s - 2);
    io.odd_dst = E[cur ^ 1];
    io.scale = fr_inv(fr_host_from_u64(two_s));
    const Fr* Zhat = sq->ztree[l] + 2 * (size_t)lo;  // node j of level l starts at j * 2s = 2 * (its first leaf)
    PS_LAUNCH(InterpCombine2K, st, (size_t)P * ns, ns, two_s, (const Fr*)E[cur], Zhat, last ? (Fr*)nullptr : E[cur ^ 1], dst);
    PS_TRY(ntt_inverse_blocks_unscaled(st, dst, (size_t)P * ns, l + 1, tabs->tw_inv, n_tw, last ? NTT_ST_SCALE : NTT_ST_TWIST, &io));
    if (!last) {
      PS_TRY(ntt_forward_blocks(st, Cbuf, (size_t)P * ns, l + 1, tabs->tw, n_tw, NTT_ST_ODD, &io));
      cur ^= 1;
    }
  }
  if (l1 < k && E[cur] != Eout) PS_TRY(dev_d2d(Eout, E[cur], (size_t)P * 2 * ns * sizeof(Fr), st));
  return PS_OK;
}

// The tree from the leaves: ev holds P x ns gate evaluations (polynomial-major) of the gates [lo, lo + ns); levels
// [0, l1) as interpolate_levels.  On the device the lowest LOW_LEVELS levels run fused (k_interp_low) when the subtree is
// large enough; `fused` = 0 forces the level-by-level path (option "interp_fused", tests).
inline int interpolate_from_leaves(ps_ctx* ctx, const SparseQap* sq, uint32_t n, int k, int P, uint32_t lo, uint32_t ns, int l1,
                                   const Fr* ev, Fr* Eout, Fr* coef) {
  Fr* E0 = ctx->arena.take<Fr>((size_t)P * 2 * ns);
  if (!E0) return PS_ERR_ALLOC;
#if PS_GPU
  if (ctx->interp_fused && ns >= LOW_G && l1 >= LOW_LEVELS && k > LOW_LEVELS) {
    const NttTables* tabs = nullptr;
    PS_TRY(ctx_ntt_tables(ctx, k + 1, &tabs));
    LowTreePtrs zt;
    for (int l = 0; l < LOW_LEVELS; l++) zt.z[l] = sq->ztree[l];
    const size_t smem = (size_t)3 * LOW_G * sizeof(Fr);
    static bool attr_set = false;
    if (!attr_set && smem > 48 * 1024) {
      PS_CUDA_TRY(cudaFuncSetAttribute(k_interp_low, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      attr_set = true;
    }
    k_interp_low<<<(uint32_t)P * (ns / LOW_G), 256, smem, ctx->stream>>>(ns, lo, ev, sq->inv_zprime + lo, zt, sq->twist, tabs->tw, tabs->tw_inv,
                                                                          2 * n, E0);
    PS_CUDA_TRY(cudaGetLastError());
    launch_counter()++;
    if (l1 == LOW_LEVELS) return dev_d2d(Eout, E0, (size_t)P * 2 * ns * sizeof(Fr), ctx->stream);
    return interpolate_levels(ctx, sq, n, k, P, lo, ns, LOW_LEVELS, l1, E0, Eout, coef);
  }
#endif
  PS_LAUNCH(InterpLeafK, ctx->stream, (size_t)P * ns, ns, ev, (const Fr*)(sq->inv_zprime + lo), E0);
  return interpolate_levels(ctx, sq, n, k, P, lo, ns, 0, l1, E0, Eout, coef);
}

// ev: P polynomials' evaluations on {1..n} (P*n) -> coef: their coefficients (P*n).
inline int interpolate_ap(ps_ctx* ctx, const SparseQap* sq, uint32_t n, int k, int P, const Fr* ev, Fr* coef) {
  return interpolate_from_leaves(ctx, sq, n, k, P, 0, n, k, ev, (Fr*)nullptr, coef);
}

// ---- division by z through the power-series inverse of rev(z) -----------------------------------------
// F[i] = z[n - i] for i < lim (and i <= n), 0 above                                   (thread over len)
struct RevZPadK {
  static constexpr int BLOCK = 256;
  PS_DEV static void run(uint32_t i, uint32_t n, uint32_t lim, const Fr* z, Fr* F) {
    F[i] = (i < lim && i <= n) ? z[n - i] : Fr::zero();
  }
};
// E[i] = F[i] * G[i] * G[i] * scale
struct MulSqK {
  static constexpr int BLOCK = 256;
  PS_DEV static void run(uint32_t i, const Fr* F, const Fr* G, Fr scale, Fr* E) { E[i] = F[i] * G[i] * G[i] * scale; }
};
// S_new[i] = 2 S_old[i] (i < m) - e[i], i < 2m
struct NewtonUpdateK {
  static constexpr int BLOCK = 256;
  PS_DEV static void run(uint32_t i, uint32_t m, const Fr* S_old, const Fr* e, Fr* S_new) {
    Fr v = i < m ? S_old[i].dbl() : Fr::zero();
    S_new[i] = v - e[i];
  }
};
// X[i] = A[i] * B[i] * scale
struct FrMul3K {
  static constexpr int BLOCK = 256;
  PS_DEV static void run(uint32_t i, const Fr* A, const Fr* B, Fr scale, Fr* X) { X[i] = A[i] * B[i] * scale; }
};
// T[i] = P[2n-2-i] for i <= n-2, 0 above                                               (thread over 2n)
struct RevTopK {
  static constexpr int BLOCK = 256;
  PS_DEV static void run(uint32_t i, uint32_t n, const Fr* Pc, Fr* T) { T[i] = (i + 2 <= n) ? Pc[2 * (size_t)n - 2 - i] : Fr::zero(); }
};
// h[k] = revh[n-2-k] for k <= n-2, h[n-1] = 0                                          (thread over n)
struct RevOutK {
  static constexpr int BLOCK = 256;
  PS_DEV static void run(uint32_t k, uint32_t n, const Fr* revh, Fr* h) { h[k] = (k + 2 <= n) ? revh[n - 2 - k] : Fr::zero(); }
};
// c[i] = P[i] - Q[i], i < n
struct FrSubK {
  static constexpr int BLOCK = 256;
  PS_DEV static void run(uint32_t i, const Fr* Pc, const Fr* Q, Fr* c) { c[i] = Pc[i] - Q[i]; }
};

// Precomputes s_hat and z_hat from z (n+1 coefficients on the device); transforms of size 2 np = 2^(k+1) >= 2n.
inline int series_tables_build(ps_ctx* ctx, SparseQap* sq, uint32_t n, uint32_t np, int k, const Fr* d_z) {
  ps_stream_t st = ctx->stream;
  Arena& ar = ctx->arena;
  const size_t N2 = (size_t)2 * np;
  Fr* S[2] = {ar.take<Fr>(N2), ar.take<Fr>(N2)};
  Fr* F = ar.take<Fr>(2 * N2); Fr* G = ar.take<Fr>(2 * N2); Fr* E = ar.take<Fr>(2 * N2);
  if (!S[0] || !S[1] || !F || !G || !E) return PS_ERR_ALLOC;
  PS_TRY(dev_memset(S[0], 0, N2 * sizeof(Fr), st));
  Fr one = Fr::one();
  PS_TRY(dev_h2d(S[0], &one, sizeof(Fr), st));  // S = 1 mod x  (rev(z)[0] = 1: z is monic)
  PS_TRY(dev_sync(st));
  int cur = 0;
  for (uint32_t m = 1; m < n; m <<= 1) {
    int lg = 0;
    while ((1u << lg) < 4 * m) lg++;
    const NttTables* t = nullptr;
    PS_TRY(ctx_ntt_tables(ctx, lg, &t));
    const uint32_t len = 4 * m;
    PS_LAUNCH(RevZPadK, st, len, n, 2 * m, d_z, F);
    PS_LAUNCH(ScaleCopyK, st, len, (const Fr*)S[cur], m, (const Fr*)nullptr, G);
    PS_TRY(ntt_forward(st, F, lg, t->tw));
    PS_TRY(ntt_forward(st, G, lg, t->tw));
    PS_LAUNCH(MulSqK, st, len, (const Fr*)F, (const Fr*)G, fr_inv(fr_host_from_u64(len)), E);
    PS_TRY(ntt_inverse_unscaled(st, E, lg, t->tw_inv));
    PS_LAUNCH(NewtonUpdateK, st, 2 * m, m, (const Fr*)S[cur], (const Fr*)E, S[cur ^ 1]);
    cur ^= 1;
  }
  // precision is now >= n >= n-1 coefficients; keep the first n-1, transform at size 2n
  const NttTables* t2 = nullptr;
  PS_TRY(ctx_ntt_tables(ctx, k + 1, &t2));
  PS_TRY(dev_alloc((void**)&sq->s_hat, N2 * sizeof(Fr)));
  PS_TRY(dev_alloc((void**)&sq->z_hat, N2 * sizeof(Fr)));
  PS_LAUNCH(ScaleCopyK, st, N2, (const Fr*)S[cur], n - 1, (const Fr*)nullptr, sq->s_hat);
  PS_TRY(ntt_forward(st, sq->s_hat, k + 1, t2->tw));
  PS_LAUNCH(ScaleCopyK, st, N2, d_z, n + 1, (const Fr*)nullptr, sq->z_hat);
  PS_TRY(ntt_forward(st, sq->z_hat, k + 1, t2->tw));
  return PS_OK;
}

// h = floor(a*b / z): a, b have n coefficients (device); h receives n entries (h[n-1] = 0).
// c_out (optional, n entries) = a*b - h*z, the polynomial computeAggregatePoly would have returned.
// Transforms of size 2 np = 2^(k+1) >= 2n (np = sq->np).
inline int quotient_series(ps_ctx* ctx, const SparseQap* sq, uint32_t n, int k, const Fr* a, const Fr* b, Fr* h, Fr* c_out) {
  ps_stream_t st = ctx->stream;
  Arena& ar = ctx->arena;
  const size_t N2 = (size_t)2 * sq->np;
  const NttTables* t = nullptr;
  PS_TRY(ctx_ntt_tables(ctx, k + 1, &t));
  Fr* A = ar.take<Fr>(N2); Fr* B = ar.take<Fr>(N2); Fr* Pc = ar.take<Fr>(N2); Fr* T = ar.take<Fr>(N2);
  if (!A || !B || !Pc || !T) return PS_ERR_ALLOC;
  Fr invN = fr_inv(fr_host_from_u64(N2));
  PS_LAUNCH(ScaleCopyK, st, N2, a, n, (const Fr*)nullptr, A);
  PS_LAUNCH(ScaleCopyK, st, N2, b, n, (const Fr*)nullptr, B);
  PS_TRY(ntt_forward(st, A, k + 1, t->tw));
  PS_TRY(ntt_forward(st, B, k + 1, t->tw));
  PS_LAUNCH(FrMul3K, st, N2, (const Fr*)A, (const Fr*)B, invN, Pc);
  PS_TRY(ntt_inverse_unscaled(st, Pc, k + 1, t->tw_inv));          // Pc = a*b, 2n-1 coefficients
  PS_LAUNCH(RevTopK, st, N2, n, (const Fr*)Pc, T);
  PS_TRY(ntt_forward(st, T, k + 1, t->tw));
  PS_LAUNCH(FrMul3K, st, N2, (const Fr*)T, (const Fr*)sq->s_hat, invN, T);
  PS_TRY(ntt_inverse_unscaled(st, T, k + 1, t->tw_inv));           // T[0..n-1) = rev(h)
  PS_LAUNCH(RevOutK, st, n, n, (const Fr*)T, h);
  if (c_out) {
    PS_LAUNCH(ScaleCopyK, st, N2, (const Fr*)h, n, (const Fr*)nullptr, A);
    PS_TRY(ntt_forward(st, A, k + 1, t->tw));
    PS_LAUNCH(FrMul3K, st, N2, (const Fr*)A, (const Fr*)sq->z_hat, invN, A);
    PS_TRY(ntt_inverse_unscaled(st, A, k + 1, t->tw_inv));         // A = h*z
    PS_LAUNCH(FrSubK, st, n, (const Fr*)Pc, (const Fr*)A, c_out);
  }
  return PS_OK;
}

}  // namespace ps
