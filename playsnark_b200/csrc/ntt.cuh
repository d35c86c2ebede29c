// Radix-2 number-theoretic transforms over Fr and the element-wise polynomial kernels around them.
//
// They replace the reference's schoolbook Poly.Mul (algebra.go:92-105) and long division Poly.Div2
// (algebra.go:140-159) inside QAP.Quotient (qap.go:151-162); see poly.cuh for how.
// Forward transforms are decimation-in-frequency (natural order in, bit-reversed out), inverse
// transforms decimation-in-time (bit-reversed in, natural out), so no permutation pass is needed
// between them; each launch fuses R <= 3 butterfly stages in registers (2^R elements per thread).
// omega_n = 7^((r-1)/n); Fr has 2-adicity 32.
#pragma once
#include "context.cuh"

namespace ps {

PS_DEV Fr fr_load(const Fr* p) { return *p; }

// out[i] = scale * base^i   (square-and-multiply per element; tables are built once per size)
struct FrPowTableK {
  static constexpr int BLOCK = 256;
  PS_DEV static void run(uint32_t i, Fr base, Fr scale, Fr* out) {
    Fr acc = scale, b = base;
    uint32_t e = i;
    while (e) {
      if (e & 1) acc = acc * b;
      b = b * b;
      e >>= 1;
    }
    out[i] = acc;
  }
};

// a[i] *= t[i]
struct FrMulTableK {
  static constexpr int BLOCK = 256;
  PS_DEV static void run(uint32_t i, Fr* a, const Fr* t) { a[i] = a[i] * t[i]; }
};

// Twiddle addressing.  The flat table tw[i] = omega_n^i makes a butterfly stage on sub-transforms of size M read
// tw[p * (n / M)]: consecutive lanes (consecutive p) then touch addresses n/M * 32 B apart -- a separate 128-byte line
// per lane for every twiddle load of the middle passes (the radix-8 pass profiled at 71 % L1/TEX throughput next to
// 60 % multiplier activity).  The LAYERED table stored right behind the flat one holds, for every M = 2^s <= n,
// omega_M^p for p < M/2 contiguously (offset M/2 - 1): lanes read consecutive 32-byte entries.  Same values, same
// products: results are bit-identical.  PS_NTT_LAYERED = 0 keeps the flat addressing (A/B builds).
#ifndef PS_NTT_LAYERED
#define PS_NTT_LAYERED 1
#endif
// twiddle omega_M^p of a table built for n: M = n / tw_mul
PS_DEV Fr ntt_twiddle(const Fr* tw, uint32_t n, uint32_t M, uint32_t p) {
#if PS_NTT_LAYERED
  return fe_ld(tw + (size_t)(n >> 1) + ((M >> 1) - 1) + p);
#else
  return fe_ld(tw + (size_t)p * (n / M));
#endif
}

#ifndef PS_NTT_BLOCK
#define PS_NTT_BLOCK 128
#endif
#ifndef PS_NTT_MINB
#define PS_NTT_MINB 4
#endif
#ifndef PS_NTT_SMEM_PAD
#define PS_NTT_SMEM_PAD 0      // A/B builds only: see SmemPad in backend.cuh
#endif

// How the LAST pass of a batched transform writes its outputs (compile-time mode, so that every variant
// stays as lean as the plain kernel).  The interpolation tree (interp.cuh) fuses its element-wise steps
// into these stores:
//   NTT_ST_PLAIN  a[idx] = v
//   NTT_ST_TWIST  a[idx] = v * io.twist[idx mod two_s]
//   NTT_ST_SCALE  a[idx] = v * io.scale
//   NTT_ST_ODD    v goes to the odd half of its parent block in io.odd_dst (not to a)
enum { NTT_ST_PLAIN = 0, NTT_ST_TWIST = 1, NTT_ST_SCALE = 2, NTT_ST_ODD = 3 };
struct NttIO {
  const Fr* twist = nullptr;
  Fr* odd_dst = nullptr;
  uint32_t ns = 0, two_s = 0;
  Fr scale;
};
template <int MODE>
PS_DEV void ntt_io_store(const NttIO& io, Fr* a, size_t idx, const Fr& v) {
  if (MODE == NTT_ST_ODD) {
    const uint32_t poly = (uint32_t)(idx / io.ns), r = (uint32_t)(idx % io.ns);
    const uint32_t p = r / io.two_s, e = r % io.two_s;
    io.odd_dst[(size_t)poly * 2 * io.ns + (size_t)p * 2 * io.two_s + io.two_s + e] = v;
  } else if (MODE == NTT_ST_TWIST) {
    a[idx] = v * fe_ld(io.twist + ((uint32_t)idx & (io.two_s - 1)));
  } else if (MODE == NTT_ST_SCALE) {
    a[idx] = v * io.scale;
  } else {
    a[idx] = v;
  }
}

// TRIV: the pass in which the sub-transforms have shrunk to 2^R elements (q == 1, the last pass of a forward transform):
// there the twiddle of every butterfly with i == 0 is omega^0 = 1 and its product is skipped -- 7 of the 12 products of
// a radix-8 pass, which is most of the work of the small transforms at the bottom of the interpolation tree.
template <int R, int MODE = NTT_ST_PLAIN, bool TRIV = false>
struct NttDifK {
  static constexpr int BLOCK = PS_NTT_BLOCK;
  static constexpr int MIN_BLOCKS = PS_NTT_MINB;   // 8 field elements per thread: cap registers for 4 warps / scheduler
  static constexpr int SMEM_PAD = PS_NTT_SMEM_PAD;
  // one launch = R stages on sub-transforms of size B (B >= 2^R); n/2^R threads
  PS_DEV static void run(uint32_t tid, Fr* a, uint32_t n, uint32_t B, const Fr* tw, NttIO io) {
    const uint32_t q = B >> R;
    const uint32_t blk = tid / q, j = tid % q;
    const size_t base = (size_t)blk * B + j;
    Fr x[1 << R];
#pragma unroll
    for (int k = 0; k < (1 << R); k++) x[k] = a[base + (size_t)k * q];
#pragma unroll
    for (int t = 0; t < R; t++) {
      const int hl = 1 << (R - 1 - t);           // half size in units of k
      const uint32_t M = B >> t;                 // size of the sub-transforms of this stage
#pragma unroll
      for (int grp = 0; grp < (1 << t); grp++) {
#pragma unroll
        for (int i = 0; i < hl; i++) {
          const int k = grp * 2 * hl + i, k2 = k + hl;
          uint32_t p = j + (uint32_t)i * q;
          Fr u = x[k], v = x[k2];
          x[k] = u + v;
          if (TRIV && i == 0) x[k2] = u - v;
          else x[k2] = (u - v) * ntt_twiddle(tw, n, M, p);
        }
      }
    }
#pragma unroll
    for (int k = 0; k < (1 << R); k++) ntt_io_store<MODE>(io, a, base + (size_t)k * q, x[k]);
  }
};

// TRIV: the first pass of an inverse transform (B0 == 1): twiddles with i == 0 are 1, as above
template <int R, int MODE = NTT_ST_PLAIN, bool TRIV = false>
struct NttDitK {
  static constexpr int BLOCK = PS_NTT_BLOCK;
  static constexpr int MIN_BLOCKS = PS_NTT_MINB;
  static constexpr int SMEM_PAD = PS_NTT_SMEM_PAD;
  // one launch = R stages that grow finished sub-transforms of size B0 to B0 * 2^R
  PS_DEV static void run(uint32_t tid, Fr* a, uint32_t n, uint32_t B0, const Fr* tw_inv, NttIO io) {
    const uint32_t blk = tid / B0, j = tid % B0;
    const size_t base = ((size_t)blk * B0 << R) + j;
    Fr x[1 << R];
#pragma unroll
    for (int k = 0; k < (1 << R); k++) x[k] = a[base + (size_t)k * B0];
#pragma unroll
    for (int t = 0; t < R; t++) {
      const int hl = 1 << t;
      const uint32_t M = (2 * B0) << t;          // size of the sub-transforms this stage completes
#pragma unroll
      for (int grp = 0; grp < (1 << (R - 1 - t)); grp++) {
#pragma unroll
        for (int i = 0; i < hl; i++) {
          const int k = grp * 2 * hl + i, k2 = k + hl;
          uint32_t p = j + (uint32_t)i * B0;
          Fr u = x[k], v = x[k2];
          if (!(TRIV && i == 0)) v = v * ntt_twiddle(tw_inv, n, M, p);
          x[k] = u + v;
          x[k2] = u - v;
        }
      }
    }
#pragma unroll
    for (int k = 0; k < (1 << R); k++) ntt_io_store<MODE>(io, a, base + (size_t)k * B0, x[k]);
  }
};

struct BitRevK {
  static constexpr int BLOCK = 256;
  PS_DEV static void run(uint32_t i, const Fr* in, Fr* out, int log_n) {
    uint32_t r = 0, v = i;
    for (int b = 0; b < log_n; b++) { r = (r << 1) | (v & 1); v >>= 1; }
    out[r] = in[i];
  }
};

inline Fr fr_host_from_u64(uint64_t v) {
  Fr x = Fr::zero();
  x.v[0] = (uint32_t)v; x.v[1] = (uint32_t)(v >> 32);
  return x.to_mont();
}
inline Fr fr_host_pow(Fr b, uint64_t e) {
  Fr acc = Fr::one();
  while (e) { if (e & 1) acc = acc * b; b = b * b; e >>= 1; }
  return acc;
}
// primitive 2^log_n-th root of unity (Montgomery form), computed on the host from the 2^32-th root
inline Fr fr_root_of_unity(int log_n) {
  Fr w = Fr::from_const<FrParams::ROOT_2_32>();
  for (int i = log_n; i < 32; i++) w = w * w;
  return w;
}

// Batched forward DIF: `len` elements = len / 2^log_block independent transforms of size 2^log_block
// laid out back to back (natural -> bit-reversed inside each block).  tw[i] = omega_{n_tw}^i for
// i < n_tw/2 with n_tw >= 2^log_block (a block transform is the tail of a larger transform's stages).
#ifndef PS_NTT_MAXR
#define PS_NTT_MAXR 3
#endif
template <template <int, int, bool> class K, int MODE, bool TRIV>
inline int ntt_launch_pass(ps_stream_t st, int r, size_t len, Fr* a, uint32_t n_tw, uint32_t B, const Fr* tw, const NttIO& io) {
  if (r == 3) PS_LAUNCH(K<3 PS_COMMA MODE PS_COMMA TRIV>, st, len >> 3, a, n_tw, B, tw, io);
  else if (r == 2) PS_LAUNCH(K<2 PS_COMMA MODE PS_COMMA TRIV>, st, len >> 2, a, n_tw, B, tw, io);
  else PS_LAUNCH(K<1 PS_COMMA MODE PS_COMMA TRIV>, st, len >> 1, a, n_tw, B, tw, io);
  return PS_OK;
}
template <template <int, int, bool> class K, bool TRIV>
inline int ntt_launch_pass_mode(ps_stream_t st, int mode, int r, size_t len, Fr* a, uint32_t n_tw, uint32_t B, const Fr* tw, const NttIO& io) {
  switch (mode) {
    case NTT_ST_TWIST: return ntt_launch_pass<K, NTT_ST_TWIST, TRIV>(st, r, len, a, n_tw, B, tw, io);
    case NTT_ST_SCALE: return ntt_launch_pass<K, NTT_ST_SCALE, TRIV>(st, r, len, a, n_tw, B, tw, io);
    case NTT_ST_ODD: return ntt_launch_pass<K, NTT_ST_ODD, TRIV>(st, r, len, a, n_tw, B, tw, io);
    default: return ntt_launch_pass<K, NTT_ST_PLAIN, TRIV>(st, r, len, a, n_tw, B, tw, io);
  }
}
// `store_mode` / `io`: output handling of the LAST pass (the earlier passes store in place).
inline int ntt_forward_blocks(ps_stream_t st, Fr* a, size_t len, int log_block, const Fr* tw, uint32_t n_tw,
                              int store_mode = NTT_ST_PLAIN, const NttIO* io = nullptr) {
  int done = 0;
  NttIO none;
  while (done < log_block) {
    int r = log_block - done >= PS_NTT_MAXR ? PS_NTT_MAXR : log_block - done;
    uint32_t B = 1u << (log_block - done);
    const bool lastp = done + r == log_block;
    // the last pass works on sub-transforms of exactly 2^r elements: its i == 0 twiddles are 1
    if (lastp) PS_TRY((ntt_launch_pass_mode<NttDifK, true>(st, store_mode, r, len, a, n_tw, B, tw, io ? *io : none)));
    else PS_TRY((ntt_launch_pass_mode<NttDifK, false>(st, NTT_ST_PLAIN, r, len, a, n_tw, B, tw, none)));
    done += r;
  }
  return PS_OK;
}
// Batched inverse DIT (bit-reversed -> natural inside each block), WITHOUT the 1/2^log_block factor.
inline int ntt_inverse_blocks_unscaled(ps_stream_t st, Fr* a, size_t len, int log_block, const Fr* tw_inv, uint32_t n_tw,
                                       int store_mode = NTT_ST_PLAIN, const NttIO* io = nullptr) {
  int done = 0;
  NttIO none;
  while (done < log_block) {
    int r = log_block - done >= PS_NTT_MAXR ? PS_NTT_MAXR : log_block - done;
    uint32_t B0 = 1u << done;
    const bool lastp = done + r == log_block;
    // the first pass (B0 == 1) has unit twiddles for i == 0
    if (done == 0) PS_TRY((ntt_launch_pass_mode<NttDitK, true>(st, lastp ? store_mode : NTT_ST_PLAIN, r, len, a, n_tw, B0, tw_inv, (lastp && io) ? *io : none)));
    else PS_TRY((ntt_launch_pass_mode<NttDitK, false>(st, lastp ? store_mode : NTT_ST_PLAIN, r, len, a, n_tw, B0, tw_inv, (lastp && io) ? *io : none)));
    done += r;
  }
  return PS_OK;
}
// single transform of size n = 2^log_n with its own table
inline int ntt_forward(ps_stream_t st, Fr* a, int log_n, const Fr* tw) {
  return ntt_forward_blocks(st, a, (size_t)1 << log_n, log_n, tw, 1u << log_n);
}
inline int ntt_inverse_unscaled(ps_stream_t st, Fr* a, int log_n, const Fr* tw_inv) {
  return ntt_inverse_blocks_unscaled(st, a, (size_t)1 << log_n, log_n, tw_inv, 1u << log_n);
}

// layered[(M/2 - 1) + p] = flat[p * (n / M)] for M = 2^s <= n, p < M/2                  (thread over n - 1 entries)
struct TwLayerK {
  static constexpr int BLOCK = 256;
  PS_DEV static void run(uint32_t idx, int log_n, const Fr* flat, Fr* layered) {
    int s = 1;
    while ((2u << (s - 1)) - 1 <= idx) s++;        // layer s starts at 2^(s-1) - 1
    const uint32_t p = idx - ((1u << (s - 1)) - 1);
    layered[idx] = flat[(size_t)p << (log_n - s)];
  }
};

// flat table omega^i (i < n/2) followed by the layered table (n - 1 entries, see PS_NTT_LAYERED), in one allocation
inline int ntt_tables_build(ps_stream_t st, int log_n, NttTables* t) {
  if (log_n < 0 || log_n > 30) return PS_ERR_ARG;
  t->release();
  const size_t n = (size_t)1 << log_n;
  size_t half = log_n ? n >> 1 : 1;
  PS_TRY(dev_alloc((void**)&t->tw, (half + n) * sizeof(Fr)));
  PS_TRY(dev_alloc((void**)&t->tw_inv, (half + n) * sizeof(Fr)));
  Fr w = fr_root_of_unity(log_n);
  Fr wi = fr_host_pow(w, ((uint64_t)1 << log_n) - 1);  // w^-1 = w^(n-1)
  PS_LAUNCH(FrPowTableK, st, half, w, Fr::one(), t->tw);
  PS_LAUNCH(FrPowTableK, st, half, wi, Fr::one(), t->tw_inv);
  if (log_n >= 1) {
    PS_LAUNCH(TwLayerK, st, n - 1, log_n, (const Fr*)t->tw, t->tw + half);
    PS_LAUNCH(TwLayerK, st, n - 1, log_n, (const Fr*)t->tw_inv, t->tw_inv + half);
  }
  t->log_n = log_n;   // only a completely built entry is marked valid
  return PS_OK;
}

}  // namespace ps
