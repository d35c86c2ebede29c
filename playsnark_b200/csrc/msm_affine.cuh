// Bucket accumulation by batched affine additions (alternative to MsmAccumK's XYZZ chains).
//
// An affine addition costs one field inversion plus 2M + 1S; with Montgomery's simultaneous
// inversion the inversion is shared by a whole batch (3M per element), so one addition costs about
// 5M + 1S instead of the 8M + 2S of an XYZZ mixed addition -- the IMAD work per bucket entry drops
// by about 40 %.  The sum inside a bucket is re-associated as a balanced tree: in every round the
// entries of a bucket segment are added in pairs (rank 2k with rank 2k+1; an odd last entry is
// carried), all pairs of all buckets form one batch, and the segment lengths halve.  The group
// element per bucket is the same whatever the association order.
//
// One round (all kernels are position-computable, nothing depends on a per-thread chunking):
//   AffHalfLenK     cnt[b] = ceil(len[b] / 2)         -> exclusive scan = output offsets
//   AffBidFillK     bucket id of every output slot (binary search in the output offsets)
//   AffPairDenK     denominator of every pair (x2 - x1, or 2y for a doubling, or 1 for a carry /
//                   infinity case) and the running product inside each thread's K outputs
//   AffInvGroupK    inverts the thread totals, GROUP totals per inversion
//   AffPairFinishK  back-substitution -> 1/den per pair, lambda, the sum; writes the output points
// After ceil(log2(max segment length)) rounds every bucket holds at most one point
// (AffFinalScatterK writes it to the bucket array in XYZZ form for the usual reduction).
#pragma once
#include "msm.cuh"

namespace ps {

template <class F>
struct AffSrc {              // where a round reads its input points from
  const Affine<F>* tab;      // round 0: the base tables ...
  const uint32_t* ent;       // ... indexed by the sorted entries (sign in bit 31)
  const Affine<F>* pin;      // later rounds: the previous round's output (tab == nullptr)
};
template <class F>
PS_DEV Affine<F> aff_load(const AffSrc<F>& s, uint32_t i) {
  return s.pin ? s.pin[i] : msm_load_point(s.tab, s.ent[i]);
}
template <class F>
PS_DEV F aff_load_x(const AffSrc<F>& s, uint32_t i) {
  return s.pin ? s.pin[i].x : s.tab[s.ent[i] & 0x7FFFFFFFu].x;
}

enum { AFF_LEFT = 0, AFF_RIGHT = 1, AFF_ADD = 2, AFF_DBL = 3, AFF_INF = 4 };

// classification of the pair (p1, p2) and its denominator; never returns a zero denominator
template <class F>
PS_DEV int aff_classify(const Affine<F>& p1, const Affine<F>& p2, bool has_right, F& den) {
  den = F::one();
  if (!has_right || p2.is_inf()) return AFF_LEFT;
  if (p1.is_inf()) return AFF_RIGHT;
  F dx = p2.x - p1.x;
  if (!dx.is_zero()) { den = dx; return AFF_ADD; }
  if (p1.y == p2.y && !p1.y.is_zero()) { den = p1.y.dbl(); return AFF_DBL; }
  return AFF_INF;
}

struct AffHalfLenK {   // cnt[b] = ceil(len/2) for b < nb, cnt[nb] = 0; maxlen = max len
  static constexpr int BLOCK = 256;
  PS_DEV static void run(uint32_t b, uint32_t nb, const uint32_t* off_in, uint32_t* cnt, uint32_t* maxlen) {
    if (b == nb) { cnt[b] = 0; return; }
    uint32_t len = off_in[b + 1] - off_in[b];
    cnt[b] = (len + 1) / 2;
    if (maxlen && len > 1) ps_atomic_max(maxlen, len);
  }
};

struct AffBidFillK {   // bid[j] = bucket of output slot j, j < off_out[nb]
  static constexpr int BLOCK = 256;
  PS_DEV static void run(uint32_t j, uint32_t nb, const uint32_t* off_out, uint32_t* bid) {
    if (j >= off_out[nb]) return;
    bid[j] = msm_find_bucket(off_out, nb, j);
  }
};

template <class F>
struct AffPairDenK {
  static constexpr int BLOCK = 128;
  PS_DEV static void run(uint32_t t, uint32_t K, uint32_t nb, AffSrc<F> src, const uint32_t* off_in, const uint32_t* off_out,
                         const uint32_t* bid, F* prefix, F* totals) {
    const uint32_t E = off_out[nb];
    F run = F::one();
    for (uint32_t k = 0; k < K; k++) {
      uint64_t j64 = (uint64_t)t * K + k;
      if (j64 >= E) break;
      uint32_t j = (uint32_t)j64;
      uint32_t b = bid[j];
      uint32_t i = off_in[b] + 2 * (j - off_out[b]);
      bool has_right = i + 1 < off_in[b + 1];
      F den = F::one();
      if (has_right) {
        F x1 = aff_load_x(src, i), x2 = aff_load_x(src, i + 1);
        F dx = x2 - x1;
        if (!dx.is_zero() && !x1.is_zero() && !x2.is_zero()) den = dx;            // the common case: no y needed
        else aff_classify(aff_load(src, i), aff_load(src, i + 1), true, den);       // equal x or a possible (0,0)
      }
      run = run * den;
      prefix[j] = run;
    }
    totals[t] = run;
  }
};

template <class F>
struct AffInvGroupK {
  static constexpr int BLOCK = 64;
  static constexpr uint32_t GROUP = 32;
  // thread u: inverse of every totals[idx], idx in [u*GROUP, (u+1)*GROUP), with one field inversion
  PS_DEV static void run(uint32_t u, uint32_t n_tot, const F* totals, F* pre2, F* invtot) {
    uint32_t lo = u * GROUP, hi = lo + GROUP < n_tot ? lo + GROUP : n_tot;
    F acc = F::one();
    for (uint32_t idx = lo; idx < hi; idx++) { acc = acc * totals[idx]; pre2[idx] = acc; }
    F inv = FieldInv<F>::inv(acc);
    for (uint32_t idx = hi; idx-- > lo;) {
      invtot[idx] = idx > lo ? inv * pre2[idx - 1] : inv;
      inv = inv * totals[idx];
    }
  }
};

template <class F>
struct AffPairFinishK {
  static constexpr int BLOCK = 128;
  PS_DEV static void run(uint32_t t, uint32_t K, uint32_t nb, AffSrc<F> src, const uint32_t* off_in, const uint32_t* off_out,
                         const uint32_t* bid, const F* prefix, const F* invtot, Affine<F>* pout) {
    const uint32_t E = off_out[nb];
    uint64_t j0 = (uint64_t)t * K;
    if (j0 >= E) return;
    uint32_t cnt = (E - j0 < K) ? (uint32_t)(E - j0) : K;
    F inv_run = invtot[t];   // 1 / (product of this thread's denominators)
    for (uint32_t k = cnt; k-- > 0;) {
      uint32_t j = (uint32_t)j0 + k;
      uint32_t b = bid[j];
      uint32_t i = off_in[b] + 2 * (j - off_out[b]);
      bool has_right = i + 1 < off_in[b + 1];
      Affine<F> p1 = aff_load(src, i);
      Affine<F> p2 = has_right ? aff_load(src, i + 1) : Affine<F>::inf();
      F den;
      int mode = aff_classify(p1, p2, has_right, den);
      F inv_j = k > 0 ? inv_run * prefix[j - 1] : inv_run;
      inv_run = inv_run * den;
      Affine<F> out;
      if (mode == AFF_LEFT) out = p1;
      else if (mode == AFF_RIGHT) out = p2;
      else if (mode == AFF_INF) out = Affine<F>::inf();
      else {
        F num;
        if (mode == AFF_ADD) num = p2.y - p1.y;
        else { F xx = p1.x.sqr(); num = xx.dbl() + xx; }
        F lam = num * inv_j;
        F x3 = lam.sqr() - p1.x - p2.x;
        F y3 = lam * (p1.x - x3) - p1.y;
        out = Affine<F>{x3, y3};
      }
      pout[j] = out;
    }
  }
};

template <class F>
struct AffFinalScatterK {   // buckets[bid[j]] = point j (at most one per bucket is left)
  static constexpr int BLOCK = 128;
  PS_DEV static void run(uint32_t j, uint32_t nb, AffSrc<F> src, const uint32_t* off_cur, const uint32_t* bid, XYZZ<F>* buckets) {
    if (j >= off_cur[nb]) return;
    buckets[bid[j]] = XYZZ<F>::from_affine(aff_load(src, j));
  }
};

// Accumulates the sorted entries into `buckets` (pre-zeroed).  off0: nb+1 offsets of the sorted entries.
template <class F>
int msm_accumulate_affine(ps_ctx* ctx, uint32_t nb, size_t max_ent, const Affine<F>* tab, const uint32_t* ent, const uint32_t* off0,
                          XYZZ<F>* buckets) {
  ps_stream_t st = ctx->stream;
  Arena& ar = ctx->arena;
  const uint32_t K = 16;
  // upper bounds of the element count per round
  auto next_bound = [&](size_t e) { size_t nz = e < nb ? e : nb; return (e + nz + 1) / 2; };
  const size_t e1 = next_bound(max_ent), e2 = next_bound(e1);
  uint32_t* offs[2] = {ar.take<uint32_t>((size_t)nb + 1), ar.take<uint32_t>((size_t)nb + 1)};
  uint32_t* cnt = ar.take<uint32_t>((size_t)nb + 1);
  uint32_t* tile_sums = ar.take<uint32_t>((size_t)nb / 2048 + 2);
  uint32_t* bids[2] = {ar.take<uint32_t>(e1), ar.take<uint32_t>(e2 > max_ent ? e2 : (e2 > e1 ? e2 : e1))};
  Affine<F>* pts[2] = {ar.take<Affine<F>>(e1), ar.take<Affine<F>>(e2)};
  F* prefix = ar.take<F>(e1);
  const size_t n_tot_max = (e1 + K - 1) / K;
  F* totals = ar.take<F>(n_tot_max);
  F* pre2 = ar.take<F>(n_tot_max);
  F* invtot = ar.take<F>(n_tot_max);
  uint32_t* d_maxlen = ar.take<uint32_t>(1);
  if (!offs[0] || !offs[1] || !cnt || !tile_sums || !bids[0] || !bids[1] || !pts[0] || !pts[1] || !prefix || !totals || !pre2 ||
      !invtot || !d_maxlen)
    return PS_ERR_ALLOC;

  // round 0 offsets + the longest segment (decides the number of rounds)
  PS_TRY(dev_memset(d_maxlen, 0, 4, st));
  PS_LAUNCH(AffHalfLenK, st, (size_t)nb + 1, nb, off0, cnt, d_maxlen);
  uint32_t maxlen = 0;
  PS_TRY(dev_d2h(&maxlen, d_maxlen, 4, st));
  PS_TRY(dev_sync(st));
  int rounds = 0;
  while (((uint64_t)1 << rounds) < maxlen) rounds++;

  AffSrc<F> src{tab, ent, nullptr};
  const uint32_t* off_in = off0;
  const uint32_t* bid_cur = nullptr;
  size_t bound = max_ent;
  for (int r = 0; r < rounds; r++) {
    uint32_t* off_out = offs[r & 1];
    uint32_t* bid_out = bids[r & 1];
    Affine<F>* pout = pts[r & 1];
    if (r > 0) PS_LAUNCH(AffHalfLenK, st, (size_t)nb + 1, nb, off_in, cnt, (uint32_t*)nullptr);
    PS_TRY(exclusive_scan_u32(st, cnt, off_out, tile_sums, nb + 1));
    const size_t e_out = next_bound(bound);
    const size_t n_thr = (e_out + K - 1) / K;
    PS_LAUNCH(AffBidFillK, st, e_out, nb, (const uint32_t*)off_out, bid_out);
    PS_LAUNCH(AffPairDenK<F>, st, n_thr, K, nb, src, off_in, (const uint32_t*)off_out, (const uint32_t*)bid_out, prefix, totals);
    PS_LAUNCH(AffInvGroupK<F>, st, (n_thr + AffInvGroupK<F>::GROUP - 1) / AffInvGroupK<F>::GROUP, (uint32_t)n_thr, (const F*)totals, pre2,
              invtot);
    PS_LAUNCH(AffPairFinishK<F>, st, n_thr, K, nb, src, off_in, (const uint32_t*)off_out, (const uint32_t*)bid_out, (const F*)prefix,
              (const F*)invtot, pout);
    src = AffSrc<F>{nullptr, nullptr, pout};
    off_in = off_out;
    bid_cur = bid_out;
    bound = e_out;
  }
  if (!bid_cur) {  // no round ran: every bucket already holds at most one entry
    uint32_t* b0 = bids[0];  // the kernel only writes j < off[nb] <= nb <= e1 slots
    PS_LAUNCH(AffBidFillK, st, bound, nb, off_in, b0);
    bid_cur = b0;
  }
  PS_LAUNCH(AffFinalScatterK<F>, st, bound, nb, src, off_in, bid_cur, buckets);
  return PS_OK;
}

}  // namespace ps
