// C ABI, part 4: the verifiers -- pairing-product checks on the device and, on top of them and of
// the MSM entry points, Groth16Verify (groth16.go:214-233) and PHGR13Verify (pinochio.go:281-375).
//
// Every equation the reference tests, "left.Equal(right)" over GT values, is rewritten as  prod_i e(P_i, Q_i) == 1
// with the G1 points of one side negated; the sums over the public inputs (sum_i io_i IoLP_i, computeCommitIOSolution
// pinochio.go:390-407) and the point additions around them are MSMs with the extra points given the scalar 1.
// One thread per Miller loop, one per final exponentiation (pairing.cuh): the work per proof is constant, the
// parallelism is across pairs, equations and the proofs of a batch.
#include "group_ops.cuh"
#include "pairing.cuh"
#include "poly_api.cuh"

#include <vector>

using namespace ps;

namespace {

inline size_t g1_bytes(int format) { return format == PS_FMT_COMPRESSED ? 48 : 96; }
inline size_t g2_bytes(int format) { return format == PS_FMT_COMPRESSED ? 96 : 192; }

// -P for a point in wire format: the sign flag of the compressed form (bit 5 of byte 0; infinity has none),
// p - y for the uncompressed one (big-endian subtraction; y = 0 cannot occur on these curves)
void negate_point(uint8_t* pt, int group, int format) {
  if (pt[0] & 0x40) return;                       // infinity
  if (format == PS_FMT_COMPRESSED) { pt[0] ^= 0x20; return; }
  static const uint8_t P_BE[48] = {0x1a, 0x01, 0x11, 0xea, 0x39, 0x7f, 0xe6, 0x9a, 0x4b, 0x1b, 0xa7, 0xb6, 0x43, 0x4b, 0xac, 0xd7,
                                   0x64, 0x77, 0x4b, 0x84, 0xf3, 0x85, 0x12, 0xbf, 0x67, 0x30, 0xd2, 0xa0, 0xf6, 0xb0, 0xf6, 0x24,
                                   0x1e, 0xab, 0xff, 0xfe, 0xb1, 0x53, 0xff, 0xff, 0xb9, 0xfe, 0xff, 0xff, 0xff, 0xff, 0xaa, 0xab};
  const int coords = group == PS_G1 ? 1 : 2;      // y is the second half: 48 B (G1) or 2 x 48 B (G2: c1 | c0)
  uint8_t* y = pt + 48 * coords;
  for (int c = 0; c < coords; c++, y += 48) {
    bool zero = true;
    for (int i = 0; i < 48; i++) zero = zero && y[i] == 0;
    if (zero) continue;
    int borrow = 0;
    for (int i = 47; i >= 0; i--) {
      int d = (int)P_BE[i] - (int)y[i] - borrow;
      borrow = d < 0;
      y[i] = (uint8_t)(d + (borrow ? 256 : 0));
    }
  }
}

// sum_i scalars[i] * points[i] through the public MSM path; result as wire bytes of the same format's COMPRESSED size
// (ps_msm returns compressed points)
int msm_bytes(ps_ctx* ctx, int group, const std::vector<uint8_t>& points, const std::vector<uint8_t>& scalars_be, size_t n, int format,
              uint8_t* out) {
  ps_bases* b = nullptr;
  PS_TRY(ps_bases_load(ctx, group, points.data(), n, format, 0, 1, &b));
  int rc = ps_msm(ctx, b, scalars_be.data(), n, out);
  ps_bases_free(b);
  return rc;
}

const uint8_t ONE_BE[32] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1};

// the generator of G2 in compressed wire format (NewG2().Base(), curve.go:30)
int g2_generator_bytes(ps_ctx* ctx, uint8_t out[96]) {
  ps_bases* b = nullptr;
  PS_TRY(ps_bases_from_scalars(ctx, PS_G2, ONE_BE, 1, 0, 1, &b));
  int rc = ps_bases_export(ctx, b, 0, 1, PS_FMT_COMPRESSED, out);
  ps_bases_free(b);
  return rc;
}

}  // namespace

extern "C" {

int ps_pairing_check_batch(ps_ctx* ctx, const uint8_t* g1_points, const uint8_t* g2_points, const uint32_t* counts, size_t n_checks,
                           int format, uint8_t* ok) {
  if (!ctx || !counts || !ok || (format != PS_FMT_COMPRESSED && format != PS_FMT_AFFINE)) return PS_ERR_ARG;
  if (n_checks == 0) return PS_OK;
  if (n_checks > (1u << 24)) return PS_ERR_UNSUPPORTED;
  std::vector<uint32_t> first(n_checks + 1, 0);
  for (size_t t = 0; t < n_checks; t++) {
    if (counts[t] > (1u << 20)) return PS_ERR_UNSUPPORTED;
    first[t + 1] = first[t] + counts[t];
    if (first[t + 1] > (1u << 26)) return PS_ERR_UNSUPPORTED;
  }
  const size_t total = first[n_checks];
  if (total && (!g1_points || !g2_points)) return PS_ERR_ARG;
  PS_TRY(begin_call(ctx));
  ps_stream_t st = ctx->stream;
  Arena& ar = ctx->arena;
  const size_t b1 = total * g1_bytes(format), b2 = total * g2_bytes(format);
  uint8_t* d_in1 = ar.take<uint8_t>(b1);
  uint8_t* d_in2 = ar.take<uint8_t>(b2);
  G1Affine* P = ar.take<G1Affine>(total);
  G2Affine* Q = ar.take<G2Affine>(total);
  Fp12* f = ar.take<Fp12>(total);
  uint32_t* d_first = ar.take<uint32_t>(n_checks + 1);
  uint8_t* d_ok = ar.take<uint8_t>(n_checks);
  uint32_t* d_err = ar.take<uint32_t>(1);
  if (!d_in1 || !d_in2 || !P || !Q || !f || !d_first || !d_ok || !d_err) return PS_ERR_ALLOC;
  PS_TRY(dev_memset(d_err, 0, 4, st));
  if (total) {
    PS_TRY(dev_h2d(d_in1, g1_points, b1, st));
    PS_TRY(dev_h2d(d_in2, g2_points, b2, st));
    PS_TRY(GroupOps<Fp>::decode(ctx, d_in1, total, format, P, d_err, ctx->subgroup_check != 0));
    PS_TRY(GroupOps<Fp2>::decode(ctx, d_in2, total, format, Q, d_err, ctx->subgroup_check != 0));
  }
  PS_TRY(dev_h2d(d_first, first.data(), (n_checks + 1) * 4, st));
  PS_LAUNCH(PairingMillerK, st, total, (const G1Affine*)P, (const G2Affine*)Q, f);
  PS_LAUNCH(PairingCheckK, st, n_checks, (const Fp12*)f, (const uint32_t*)d_first, d_ok);
  PS_TRY(dev_d2h(ok, d_ok, n_checks, st));
  return check_err_flag(ctx, d_err, PS_ERR_ENCODING);   // synchronises: `first` and `ok` are settled
}

int ps_g16_verify(ps_ctx* ctx, const uint8_t* alpha, const uint8_t* beta2, const uint8_t* gamma2, const uint8_t* delta2,
                  const uint8_t* iolp, size_t n_io, const uint8_t* io_be, const uint8_t* A, const uint8_t* B, const uint8_t* C, int* ok) {
  if (!ctx || !alpha || !beta2 || !gamma2 || !delta2 || !A || !B || !C || !ok || (n_io && (!iolp || !io_be))) return PS_ERR_ARG;
  *ok = 0;
  // b1 = sum_i io[i] * IoLP[i]   (groth16.go:224-227)
  uint8_t b1[48];
  if (n_io) {
    std::vector<uint8_t> pts(iolp, iolp + n_io * 48), sc(io_be, io_be + n_io * 32);
    PS_TRY(msm_bytes(ctx, PS_G1, pts, sc, n_io, PS_FMT_COMPRESSED, b1));
  } else {
    memset(b1, 0, 48);
    b1[0] = 0xc0;
  }
  // e(A, B) == e(Alpha, Beta2) e(b1, Gamma) e(C, Delta2)   <=>   e(-A, B) e(Alpha, Beta2) e(b1, Gamma) e(C, Delta2) == 1
  uint8_t g1[4 * 48], g2[4 * 96];
  memcpy(g1, A, 48); negate_point(g1, PS_G1, PS_FMT_COMPRESSED);
  memcpy(g1 + 48, alpha, 48); memcpy(g1 + 96, b1, 48); memcpy(g1 + 144, C, 48);
  memcpy(g2, B, 96); memcpy(g2 + 96, beta2, 96); memcpy(g2 + 192, gamma2, 96); memcpy(g2 + 288, delta2, 96);
  const uint32_t count = 4;
  uint8_t res = 0;
  PS_TRY(ps_pairing_check_batch(ctx, g1, g2, &count, 1, PS_FMT_COMPRESSED, &res));
  *ok = res ? 1 : 0;
  return PS_OK;
}

int ps_phgr13_verify(ps_ctx* ctx, const uint8_t* vk_fixed, const uint8_t* vs, const uint8_t* ws, const uint8_t* ys, size_t n_io,
                     const uint8_t* io_be, const uint8_t* proof, int* ok) {
  if (!ctx || !vk_fixed || !proof || !ok || (n_io && (!vs || !ws || !ys || !io_be))) return PS_ERR_ARG;
  *ok = 0;
  // vk_fixed: av(G2) aw(G1) ay(G2) gamma(G2) bgamma(G1) bgamma2(G2) yts(G2)     (layout of ps_phgr13_setup)
  const uint8_t *av = vk_fixed, *aw = vk_fixed + 96, *ay = vk_fixed + 144, *gamma = vk_fixed + 240, *bgamma = vk_fixed + 336,
                *bgamma2 = vk_fixed + 384, *yts = vk_fixed + 480;
  // proof: the 432 bytes ps_phgr13_prove writes -- hs vss yss vass wass yass gz (G1) then wss (G2)
  const uint8_t *hs = proof, *vss = proof + 48, *yss = proof + 96, *vass = proof + 144, *wass = proof + 192, *yass = proof + 240,
                *gz = proof + 288, *wss = proof + 336;
  // gv = sum io_k vs_k + vss, gw = sum io_k ws_k + wss, gy = sum io_k ys_k + yss      (pinochio.go:293-308)
  std::vector<uint8_t> sc(io_be, io_be + n_io * 32);
  sc.insert(sc.end(), ONE_BE, ONE_BE + 32);
  uint8_t gv[48], gw[96], gy[48], lt1[48], g2gen[96];
  {
    std::vector<uint8_t> pts(vs, vs + n_io * 48);
    pts.insert(pts.end(), vss, vss + 48);
    PS_TRY(msm_bytes(ctx, PS_G1, pts, sc, n_io + 1, PS_FMT_COMPRESSED, gv));
  }
  {
    std::vector<uint8_t> pts(ws, ws + n_io * 96);
    pts.insert(pts.end(), wss, wss + 96);
    PS_TRY(msm_bytes(ctx, PS_G2, pts, sc, n_io + 1, PS_FMT_COMPRESSED, gw));
  }
  {
    std::vector<uint8_t> pts(ys, ys + n_io * 48);
    pts.insert(pts.end(), yss, yss + 48);
    PS_TRY(msm_bytes(ctx, PS_G1, pts, sc, n_io + 1, PS_FMT_COMPRESSED, gy));
  }
  {   // lt1 = vss + yss   (pinochio.go:365)
    std::vector<uint8_t> pts(vss, vss + 48), one2(ONE_BE, ONE_BE + 32);
    pts.insert(pts.end(), yss, yss + 48);
    one2.insert(one2.end(), ONE_BE, ONE_BE + 32);
    PS_TRY(msm_bytes(ctx, PS_G1, pts, one2, 2, PS_FMT_COMPRESSED, lt1));
  }
  PS_TRY(g2_generator_bytes(ctx, g2gen));
  // the five equations of pinochio.go:312-372, each as a product that must be 1
  std::vector<uint8_t> g1, g2;
  auto pair = [&](const uint8_t* p, bool neg, const uint8_t* q) {
    uint8_t t[48];
    memcpy(t, p, 48);
    if (neg) negate_point(t, PS_G1, PS_FMT_COMPRESSED);
    g1.insert(g1.end(), t, t + 48);
    g2.insert(g2.end(), q, q + 96);
  };
  pair(gv, false, gw); pair(hs, true, yts); pair(gy, true, g2gen);            // division check
  pair(vass, false, g2gen); pair(vss, true, av);                                // CRS checks
  pair(wass, false, g2gen); pair(aw, true, wss);
  pair(yass, false, g2gen); pair(yss, true, ay);
  pair(gz, false, gamma); pair(lt1, true, bgamma2); pair(bgamma, true, wss);    // linear check
  const uint32_t counts[5] = {3, 2, 2, 2, 3};
  uint8_t res[5] = {0, 0, 0, 0, 0};
  PS_TRY(ps_pairing_check_batch(ctx, g1.data(), g2.data(), counts, 5, PS_FMT_COMPRESSED, res));
  *ok = (res[0] && res[1] && res[2] && res[3] && res[4]) ? 1 : 0;
  return PS_OK;
}

}  // extern "C"
