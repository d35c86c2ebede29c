// Lane-cooperative group law for the latency-bound tails of the MSM (merge of boundary partials, bucket
// reduction, final Horner): a TEAM of 4 adjacent lanes of a warp works on ONE point addition / doubling.
//
// Why: a single thread that adds two XYZZ points runs 14 dependent-ish field products back to back
// (12M + 2S); the integer multiplier issues one IMAD.WIDE per 4 cycles for a whole warp whether 1 or
// 32 lanes are active, so a one-lane addition leaves 31/32 of every issue slot idle and costs the full
// 14 product latencies (~8 us in G1, ~28 us in G2).  The addition formula has only 4 levels of
// dependent products (4 + 4 + 3 + 3), the doubling 3 (2 + 4 + 3): with the four products of a level on
// four lanes the serial depth of every tail kernel shrinks by ~3.5x at (almost) unchanged pipe work.
//
// Every lane of a team holds identical copies of the operands and of the result (operands are loaded
// redundantly: the same addresses inside a warp are one transaction); after each level the four
// products are exchanged with width-4 shuffles.  All branches (infinity, equal points) depend only on
// the shared operands, hence are uniform inside a team; teams of the same warp may diverge freely
// because every shuffle names only its own four lanes.
//
// Host emulation (tests): lanes are run one after the other, so the team functions fall back to the
// plain serial formulas there; the result held by each lane is the same by construction.
#pragma once
#include "curve.cuh"

namespace ps {

constexpr int TEAM = 4;

struct TeamCtx {
  uint32_t tl;    // lane inside the team, 0..3
  uint32_t mask;  // shuffle mask naming the team's four lanes
  PS_DEV static TeamCtx of(uint32_t tid) {
    TeamCtx t;
    t.tl = tid & 3u;
    t.mask = 0xFu << (tid & 28u);  // blocks are multiples of 32 threads: tid & 31 is the lane id
    return t;
  }
  PS_DEV void sync() const {
#ifdef __CUDA_ARCH__
    __syncwarp(mask);
#endif
  }
};

#ifdef __CUDA_ARCH__
// the value held by team lane `src`
template <class F>
PS_DEV F team_get(const F& v, int src, uint32_t mask) {
  F r;
  const uint32_t* in = reinterpret_cast<const uint32_t*>(&v);
  uint32_t* out = reinterpret_cast<uint32_t*>(&r);
#pragma unroll
  for (int i = 0; i < (int)(sizeof(F) / 4); i++) out[i] = __shfl_sync(mask, in[i], src, TEAM);
  return r;
}
template <class F>
PS_DEV F team_sel(uint32_t tl, const F& a, const F& b, const F& c, const F& d) {
  F r;
  const uint32_t* pa = reinterpret_cast<const uint32_t*>(&a);
  const uint32_t* pb = reinterpret_cast<const uint32_t*>(&b);
  const uint32_t* pc = reinterpret_cast<const uint32_t*>(&c);
  const uint32_t* pd = reinterpret_cast<const uint32_t*>(&d);
  uint32_t* out = reinterpret_cast<uint32_t*>(&r);
#pragma unroll
  for (int i = 0; i < (int)(sizeof(F) / 4); i++) {
    uint32_t lo = (tl & 1u) ? pb[i] : pa[i];
    uint32_t hi = (tl & 1u) ? pd[i] : pc[i];
    out[i] = (tl & 2u) ? hi : lo;
  }
  return r;
}
#endif

// p = 2 p   (dbl-2008-s-1, a = 0): levels {V, X2} {W, S, ZZ3, MM} {t1, t2, ZZZ3}
template <class F>
PS_NOINLINE void xyzz_dbl_t(XYZZ<F>& p, TeamCtx tc) {
#ifdef __CUDA_ARCH__
  if (p.is_inf() || p.y.is_zero()) { p = XYZZ<F>::inf(); return; }
  const uint32_t tl = tc.tl, mk = tc.mask;
  F U = p.y.dbl();
  F a = team_sel(tl, U, p.x, U, p.x);
  F m = a * a;
  F V = team_get(m, 0, mk), X2 = team_get(m, 1, mk);
  F M = X2.dbl() + X2;
  a = team_sel(tl, U, p.x, V, M);
  F b = team_sel(tl, V, V, p.zz, M);
  m = a * b;
  F W = team_get(m, 0, mk), S = team_get(m, 1, mk), ZZ3 = team_get(m, 2, mk), MM = team_get(m, 3, mk);
  F X3 = MM - S.dbl();
  a = team_sel(tl, M, W, W, W);
  b = team_sel(tl, S - X3, p.y, p.zzz, p.zzz);
  m = a * b;
  F t1 = team_get(m, 0, mk), t2 = team_get(m, 1, mk), ZZZ3 = team_get(m, 2, mk);
  p.x = X3; p.y = t1 - t2; p.zz = ZZ3; p.zzz = ZZZ3;
#else
  (void)tc;
  p = xyzz_dbl(p);
#endif
}

// acc += q   (add-2008-s): levels {U1, U2, S1, S2} {PP, RR, ZZ12, ZZZ12} {PPP, Q, ZZ3} {t1, t2, ZZZ3}
template <class F>
PS_NOINLINE void xyzz_add_t(XYZZ<F>& acc, const XYZZ<F>& q, TeamCtx tc) {
#ifdef __CUDA_ARCH__
  if (q.is_inf()) return;
  if (acc.is_inf()) { acc = q; return; }
  const uint32_t tl = tc.tl, mk = tc.mask;
  F a = team_sel(tl, acc.x, q.x, acc.y, q.y);
  F b = team_sel(tl, q.zz, acc.zz, q.zzz, acc.zzz);
  F m = a * b;
  F U1 = team_get(m, 0, mk), U2 = team_get(m, 1, mk), S1 = team_get(m, 2, mk), S2 = team_get(m, 3, mk);
  F Pd = U2 - U1, Rd = S2 - S1;
  if (Pd.is_zero()) {
    if (Rd.is_zero()) xyzz_dbl_t(acc, tc); else acc = XYZZ<F>::inf();
    return;
  }
  a = team_sel(tl, Pd, Rd, acc.zz, acc.zzz);
  b = team_sel(tl, Pd, Rd, q.zz, q.zzz);
  m = a * b;
  F PP = team_get(m, 0, mk), RR = team_get(m, 1, mk), ZZ12 = team_get(m, 2, mk), ZZZ12 = team_get(m, 3, mk);
  a = team_sel(tl, Pd, U1, ZZ12, ZZ12);
  m = a * PP;
  F PPP = team_get(m, 0, mk), Q = team_get(m, 1, mk), ZZ3 = team_get(m, 2, mk);
  F X3 = RR - PPP - Q.dbl();
  a = team_sel(tl, Rd, S1, ZZZ12, ZZZ12);
  b = team_sel(tl, Q - X3, PPP, PPP, PPP);
  m = a * b;
  F t1 = team_get(m, 0, mk), t2 = team_get(m, 1, mk), ZZZ3 = team_get(m, 2, mk);
  acc.x = X3; acc.y = t1 - t2; acc.zz = ZZ3; acc.zzz = ZZZ3;
#else
  (void)tc;
  xyzz_add(acc, q);
#endif
}

}  // namespace ps
