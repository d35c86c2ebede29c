// C ABI of libplaysnark_b200.so (see include/playsnark_b200.h for the contract and the reference
// interfaces each entry point stands in for).  Host orchestration only: every arithmetic step is a
// kernel from msm.cuh / ntt.cuh / poly.cuh / codec.cuh on the context's stream.
#include "codec.cuh"
#include "microbench.cuh"
#include "msm.cuh"
#include "msm_affine.cuh"
#include "poly.cuh"
#include "interp.cuh"

#include <new>

using namespace ps;

struct ps_bases {
  int group = 0;
  size_t n = 0;
  int c = 0;  // fixed window (0 = choose per call)
  int T = 1;  // precomputed tables
  void* tab = nullptr;
};

struct ps_g16_key {
  size_t n = 0, n_nio = 0;
  ps_bases *A = nullptr, *B = nullptr, *C = nullptr;
};

struct ps_phgr13_key {
  size_t n = 0, n_mid = 0;
  ps_bases* g1[7] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};  // gsi vs ys vas was yas [vbs|wbs|ybs]
  ps_bases* ws = nullptr;
};

namespace {

// ---- setup-side kernels ---------------------------------------------------------------------------
template <class F> PS_DEV Affine<F> generator();
template <> PS_DEV Affine<Fp> generator<Fp>() {
  return Affine<Fp>{Fp::from_const<FpParams::G1X>(), Fp::from_const<FpParams::G1Y>()};
}
template <> PS_DEV Affine<Fp2> generator<Fp2>() {
  return Affine<Fp2>{Fp2{Fp::from_const<FpParams::G2X0>(), Fp::from_const<FpParams::G2X1>()},
                     Fp2{Fp::from_const<FpParams::G2Y0>(), Fp::from_const<FpParams::G2Y1>()}};
}

// table[w*255 + d-1] = d * 2^(8w) * generator, w < 32, 1 <= d <= 255
template <class F>
struct FixedBaseTableK {
  static constexpr int BLOCK = 64;
  PS_DEV static void run(uint32_t tid, Affine<F>* table) {
    uint32_t w = tid / 255, d = tid % 255 + 1;
    uint32_t k[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    k[w / 4] = d << (8 * (w % 4));
    XYZZ<F> r = xyzz_scalar_mul(XYZZ<F>::from_affine(generator<F>()), k, 8);
    table[tid] = xyzz_to_affine_c(r);
  }
};
// out[i] = scalar[i] * generator  (scalars: standard-form limbs), left in XYZZ form
template <class F>
struct FixedBaseMulK {
  static constexpr int BLOCK = 128;
  PS_DEV static void run(uint32_t i, const uint32_t* scalars, const Affine<F>* table, XYZZ<F>* out) {
    XYZZ<F> acc = XYZZ<F>::inf();
    for (int j = 0; j < 8; j++) {
      uint32_t limb = scalars[(size_t)i * 8 + j];
      for (int b = 0; b < 4; b++) {
        uint32_t d = (limb >> (8 * b)) & 0xFF;
        if (d) xyzz_madd_c(acc, table[(uint32_t)(4 * j + b) * 255 + d - 1]);
      }
    }
    out[i] = acc;
  }
};
// next[i] = 2^c * prev[i], left in XYZZ form
template <class F>
struct ShiftTableK {
  static constexpr int BLOCK = 128;
  PS_DEV static void run(uint32_t i, const Affine<F>* prev, XYZZ<F>* next, int c) {
    XYZZ<F> r = XYZZ<F>::from_affine(prev[i]);
    for (int d = 0; d < c; d++) r = xyzz_dbl_c(r);
    next[i] = r;
  }
};
// XYZZ -> affine for K consecutive points per thread with ONE field inversion (Montgomery's trick on ZZZ)
template <class F>
struct BatchToAffineK {
  static constexpr int BLOCK = 64;
  static constexpr uint32_t K = sizeof(F) == sizeof(Fp) ? 8 : 4;
  PS_DEV static void run(uint32_t t, uint32_t n, const XYZZ<F>* in, Affine<F>* out) {
    const uint32_t i0 = t * K;
    F pre[K];
    F acc = F::one();
#pragma unroll
    for (uint32_t k = 0; k < K; k++) {
      if (i0 + k < n) { F z = in[i0 + k].zzz; if (!in[i0 + k].zz.is_zero()) acc = acc * z; }
      pre[k] = acc;
    }
    F inv = FieldInv<F>::inv(acc);
#pragma unroll
    for (uint32_t kk = K; kk-- > 0;) {
      if (i0 + kk >= n) continue;
      XYZZ<F> p = in[i0 + kk];
      if (p.zz.is_zero()) { out[i0 + kk] = Affine<F>::inf(); continue; }
      F zzz_inv = kk > 0 ? inv * pre[kk - 1] : inv;
      inv = inv * p.zzz;
      F tt = zzz_inv * p.zz;
      out[i0 + kk] = Affine<F>{p.x * tt.sqr(), p.y * zzz_inv};
    }
  }
};

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

template <class F> struct GroupOf;
template <> struct GroupOf<Fp> { static constexpr int ID = PS_G1; };
template <> struct GroupOf<Fp2> { static constexpr int ID = PS_G2; };

inline size_t point_bytes(int group, int format) {
  return group == PS_G1 ? (format == PS_FMT_COMPRESSED ? 48 : 96) : (format == PS_FMT_COMPRESSED ? 96 : 192);
}

int begin_call(ps_ctx* ctx) {
  if (!ctx) return PS_ERR_ARG;
#if PS_GPU
  PS_CUDA_TRY(cudaSetDevice(ctx->device));
#endif
  PS_TRY(ctx->arena2.reset());
  return ctx->arena.reset();
}

// fills tables 1..T-1 from table 0
template <class F>
int bases_finish(ps_ctx* ctx, ps_bases* b) {
  Affine<F>* tab = (Affine<F>*)b->tab;
  if (b->T <= 1) return PS_OK;
  XYZZ<F>* tmp = ctx->arena.take<XYZZ<F>>(b->n);
  if (!tmp) return PS_ERR_ALLOC;
  const size_t groups = (b->n + BatchToAffineK<F>::K - 1) / BatchToAffineK<F>::K;
  for (int t = 1; t < b->T; t++) {
    PS_LAUNCH(ShiftTableK<F>, ctx->stream, b->n, (const Affine<F>*)(tab + (size_t)(t - 1) * b->n), tmp, b->c);
    PS_LAUNCH(BatchToAffineK<F>, ctx->stream, groups, (uint32_t)b->n, (const XYZZ<F>*)tmp, tab + (size_t)t * b->n);
  }
  return PS_OK;
}

// `shards`: the base set will be summed in `shards` index ranges (one per GPU of a sharded proof); the
// automatic window is sized for n / shards points per call so that each rank's bucket set (whose
// merge and reduction are a fixed cost per call) matches its share
int bases_alloc(int group, size_t n, int window_bits, int tables, int shards, int bucket_cost, ps_bases** out) {
  if (window_bits < 0 || window_bits > 24 || (window_bits > 0 && window_bits < 2)) return PS_ERR_ARG;
  if (tables < 0) {  // automatic: all windows precomputed when the tables fit comfortably in HBM
    if (window_bits == 0) window_bits = msm_pick_window_full(n / (size_t)(shards > 0 ? shards : 1) + 1, (double)bucket_cost);
    tables = msm_windows(window_bits);
    size_t need = n * (size_t)tables * (group == PS_G1 ? sizeof(G1Affine) : sizeof(G2Affine));
    size_t free_b = need * 8, total_b = 0;
#if PS_GPU
    if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) free_b = 0;
#endif
    (void)total_b;
    if (need > free_b / 4) { tables = 1; window_bits = 0; }
  }
  if (tables < 1) tables = 1;
  if (tables > 1 && window_bits == 0) return PS_ERR_ARG;
  if (window_bits && tables > msm_windows(window_bits)) tables = msm_windows(window_bits);
  if ((unsigned long long)n * (unsigned long long)tables >= 0x7FFFFFFFull) return PS_ERR_UNSUPPORTED;
  ps_bases* b = new (std::nothrow) ps_bases();
  if (!b) return PS_ERR_ALLOC;
  b->group = group; b->n = n; b->c = window_bits; b->T = tables;
  size_t pt = group == PS_G1 ? sizeof(G1Affine) : sizeof(G2Affine);
  int rc = dev_alloc(&b->tab, n * tables * pt);
  if (rc != PS_OK) { delete b; return rc; }
  *out = b;
  return PS_OK;
}

template <class F, class DecodeK>
int bases_load_t(ps_ctx* ctx, const uint8_t* points, size_t n, int format, int window_bits, int tables, ps_bases** out) {
  ps_bases* b = nullptr;
  PS_TRY(bases_alloc(GroupOf<F>::ID, n, window_bits, tables, ctx->msm_shards, ctx->msm_bucket_cost, &b));
  ps_stream_t st = ctx->stream;
  size_t bytes = n * point_bytes(GroupOf<F>::ID, format);
  uint8_t* d_in = ctx->arena.take<uint8_t>(bytes);
  uint32_t* d_err = ctx->arena.take<uint32_t>(1);
  int rc = (!d_in || !d_err) ? PS_ERR_ALLOC : PS_OK;
  if (rc == PS_OK) rc = dev_h2d(d_in, points, bytes, st);
  if (rc == PS_OK) rc = dev_memset(d_err, 0, 4, st);
  if (rc == PS_OK) rc = ps_launch<DecodeK>(st, n, (const uint8_t*)d_in, format, (Affine<F>*)b->tab, d_err);
  if (rc == PS_OK) rc = bases_finish<F>(ctx, b);
  uint32_t herr = 0;
  if (rc == PS_OK) rc = dev_d2h(&herr, d_err, 4, st);
  if (rc == PS_OK) rc = dev_sync(st);
  if (rc == PS_OK && herr) rc = PS_ERR_ENCODING;
  if (rc != PS_OK) { ps_bases_free(b); return rc; }
  *out = b;
  return PS_OK;
}

template <class F>
int fixed_base_table(ps_ctx* ctx, const Affine<F>** out) {
  int slot = GroupOf<F>::ID == PS_G1 ? 0 : 1;
  if (!ctx->fixed_base[slot]) {
    void* p = nullptr;
    PS_TRY(dev_alloc(&p, (size_t)32 * 255 * sizeof(Affine<F>)));
    ctx->fixed_base[slot] = p;
    PS_LAUNCH(FixedBaseTableK<F>, ctx->stream, (size_t)32 * 255, (Affine<F>*)p);
  }
  *out = (const Affine<F>*)ctx->fixed_base[slot];
  return PS_OK;
}

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

template <class F>
int msm_on_bases(ps_ctx* ctx, const ps_bases* b, size_t first, const uint32_t* d_scalars, size_t n, int mont, XYZZ<F>* d_out) {
  if (first + n > b->n) return PS_ERR_LENGTH;
  MsmGeom g;
  g.n = (uint32_t)n; g.nbase = (uint32_t)b->n; g.first = (uint32_t)first;
  g.c = b->c ? b->c : msm_pick_window(n);
  g.W = msm_windows(g.c);
  g.T = b->T;
  g.S = (g.W + g.T - 1) / g.T;
  g.D = 1u << (g.c - 1);
  return msm_run<F>(ctx, g, (const Affine<F>*)b->tab, d_scalars, mont, d_out);
}

template <class F>
int encode_points(ps_ctx* ctx, const XYZZ<F>* d_pts, size_t count, uint8_t* host_out) {
  size_t per = PointBytes<F>::COMP;
  uint8_t* d_bytes = ctx->arena.take<uint8_t>(count * per);
  if (!d_bytes) return PS_ERR_ALLOC;
  PS_LAUNCH(XyzzEncodeK<F>, ctx->stream, count, d_pts, (int)PS_FMT_COMPRESSED, d_bytes);
  PS_TRY(dev_d2h(host_out, d_bytes, count * per, ctx->stream));
  return PS_OK;
}

// the same into the context's page-locked staging area (offset in bytes): does not block the host
template <class F>
int encode_points_staged(ps_ctx* ctx, const XYZZ<F>* d_pts, size_t count, size_t stage_off) {
  if (stage_off + count * PointBytes<F>::COMP > ps_ctx::H_STAGE_BYTES) return PS_ERR_ARG;
  return encode_points<F>(ctx, d_pts, count, ctx->h_stage + stage_off);
}

int check_err_flag(ps_ctx* ctx, const uint32_t* d_err, int code) {
  uint32_t h = 0;
  PS_TRY(dev_d2h(&h, d_err, 4, ctx->stream));
  PS_TRY(dev_sync(ctx->stream));
  return h ? code : PS_OK;
}

// witness -> device Montgomery; a, b, c, h on the device (n' entries each)
struct QuotientBufs { Fr *w, *a, *b, *c, *h; uint32_t* flag; uint32_t* enc_err; };
int run_quotient_sparse(ps_ctx* ctx, const ps_qap* q, const uint8_t* witness_be, QuotientBufs* o, bool want_c) {
  const SparseQap* sq = (const SparseQap*)q->sparse;
  const uint32_t n = (uint32_t)q->n;
  ps_stream_t st = ctx->stream;
  uint32_t* d_w = nullptr;
  PS_TRY(stage_scalars(ctx, witness_be, q->m, 1, &d_w, &o->enc_err));
  o->w = (Fr*)d_w;
  Fr* ev = ctx->arena.take<Fr>((size_t)3 * n);
  Fr* coef = ctx->arena.take<Fr>((size_t)3 * n);
  o->h = ctx->arena.take<Fr>(n);
  o->flag = ctx->arena.take<uint32_t>(1);
  if (!ev || !coef || !o->h || !o->flag) return PS_ERR_ALLOC;
  PS_TRY(dev_memset(o->flag, 0, 4, st));
  PS_LAUNCH(SpmvK, st, (size_t)3 * n, n, 0u, (const uint32_t*)sq->mat[0].row_ptr, (const uint32_t*)sq->mat[0].col, (const Fr*)sq->mat[0].val,
            (const uint32_t*)sq->mat[1].row_ptr, (const uint32_t*)sq->mat[1].col, (const Fr*)sq->mat[1].val,
            (const uint32_t*)sq->mat[2].row_ptr, (const uint32_t*)sq->mat[2].col, (const Fr*)sq->mat[2].val, (const Fr*)o->w, ev);
  PS_LAUNCH(GateCheckK, st, n, n, (const Fr*)ev, o->flag);
  // only a and b are interpolated: c = a*b mod z never has to exist for the proof
  PS_TRY(interpolate_ap(ctx, sq, n, q->log_np, 2, ev, coef));
  o->a = coef; o->b = coef + n; o->c = coef + 2 * (size_t)n;
  PS_TRY(quotient_series(ctx, sq, n, q->log_np, o->a, o->b, o->h, want_c ? o->c : (Fr*)nullptr));
  return PS_OK;
}

int run_quotient(ps_ctx* ctx, const ps_qap* q, const uint8_t* witness_be, QuotientBufs* o, bool want_c = false) {
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

// concatenates host point arrays into one device base set
template <class F, class DecodeK>
int bases_concat(ps_ctx* ctx, int format, const uint8_t* const* parts, const size_t* counts, int nparts, ps_bases** out) {
  size_t per = point_bytes(GroupOf<F>::ID, format), total = 0;
  for (int i = 0; i < nparts; i++) total += counts[i];
  std::vector<uint8_t> buf(total * per);
  size_t o = 0;
  for (int i = 0; i < nparts; i++) {
    if (counts[i] && !parts[i]) return PS_ERR_ARG;
    memcpy(buf.data() + o, parts[i], counts[i] * per);
    o += counts[i] * per;
  }
  return bases_load_t<F, DecodeK>(ctx, buf.data(), total, format, 0, -1, out);
}

}  // namespace

extern "C" {

const char* ps_strerror(int status) {
  switch (status) {
    case PS_OK: return "ok";
    case PS_ERR_ARG: return "bad argument";
    case PS_ERR_LENGTH: return "mismatch of length between poly and blinded eval points";
    case PS_ERR_REMAINDER: return "apocalypse";
    case PS_ERR_ENCODING: return "bad scalar or point encoding";
    case PS_ERR_CUDA: return "CUDA error";
    case PS_ERR_ALLOC: return "out of memory";
    case PS_ERR_UNSUPPORTED: return "unsupported";
    default: return "unknown status";
  }
}

const char* ps_version(void) {
#if PS_GPU
  return "playsnark_b200 0.1 (sm_100a)";
#else
  return "playsnark_b200 0.1 (HOST EMULATION - tests only)";
#endif
}

uint64_t ps_launch_count(void) { return launch_counter(); }

int ps_ctx_create(int device, ps_ctx** out) {
  if (!out) return PS_ERR_ARG;
  ps_ctx* ctx = new (std::nothrow) ps_ctx();
  if (!ctx) return PS_ERR_ALLOC;
  ctx->device = device;
#if PS_GPU
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) {
    fprintf(stderr, "playsnark_b200: no CUDA device %d (found %d); there is no CPU fallback\n", device, count);
    delete ctx;
    return PS_ERR_CUDA;
  }
  PS_CUDA_TRY(cudaSetDevice(device));
  cudaDeviceProp prop;
  PS_CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  ctx->sm_count = prop.multiProcessorCount;
  cudaStream_t st;
  PS_CUDA_TRY(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  ctx->stream = st;
  ctx->own_stream = true;
  cudaStream_t st2;
  PS_CUDA_TRY(cudaStreamCreateWithFlags(&st2, cudaStreamNonBlocking));
  ctx->stream2 = st2;
  {
    cudaEvent_t e;
    PS_CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); ctx->ev_fork = e;
    PS_CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); ctx->ev_join = e;
  }
  for (int i = 0; i < 5; i++) {
    cudaEvent_t e;
    PS_CUDA_TRY(cudaEventCreate(&e));
    ctx->ev[i] = e;
  }
  for (int i = 0; i < 6; i++) {
    cudaEvent_t e;
    PS_CUDA_TRY(cudaEventCreate(&e));
    ctx->evp[i] = e;
  }
#endif
  ctx->arena.stream = ctx->stream;
  ctx->arena2.stream = ctx->stream2;
  {
    void* hp = nullptr;
    int rc = ps_host_alloc(ps_ctx::H_STAGE_BYTES, &hp);
    if (rc != PS_OK) { ps_ctx_destroy(ctx); return rc; }
    ctx->h_stage = (uint8_t*)hp;
  }
  *out = ctx;
  return PS_OK;
}

int ps_ctx_set_stream(ps_ctx* ctx, void* cuda_stream) {
  if (!ctx) return PS_ERR_ARG;
#if PS_GPU
  PS_TRY(dev_sync(ctx->stream));
  if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
  ctx->stream = (cudaStream_t)cuda_stream;
  ctx->own_stream = false;
  ctx->arena.stream = ctx->stream;
#else
  (void)cuda_stream;
#endif
  return PS_OK;
}

int ps_ctx_set_option(ps_ctx* ctx, const char* name, int value) {
  if (!ctx || !name) return PS_ERR_ARG;
  if (!strcmp(name, "msm_accumulate")) {
    if (value != 0 && value != 1) return PS_ERR_ARG;
    ctx->accum_mode = value;
    return PS_OK;
  }
  if (!strcmp(name, "msm_shards")) {
    if (value < 1 || value > 1024) return PS_ERR_ARG;
    ctx->msm_shards = value;
    return PS_OK;
  }
  if (!strcmp(name, "msm_bucket_cost")) {
    if (value < 1 || value > 100000) return PS_ERR_ARG;
    ctx->msm_bucket_cost = value;
    return PS_OK;
  }
  if (!strcmp(name, "msm_scatter")) {
    if (value < 0 || value > 2) return PS_ERR_ARG;
    ctx->msm_scatter = value;
    return PS_OK;
  }
  if (!strcmp(name, "msm_team")) {
    if (value != 0 && value != 1) return PS_ERR_ARG;
    ctx->msm_team = value;
    return PS_OK;
  }
  return PS_ERR_ARG;
}

int ps_ctx_sync(ps_ctx* ctx) { return ctx ? dev_sync(ctx->stream) : PS_ERR_ARG; }

void ps_ctx_destroy(ps_ctx* ctx) {
  if (!ctx) return;
  dev_sync(ctx->stream);
  dev_sync(ctx->stream2);
  ctx->arena.release();
  ctx->arena2.release();
  for (auto& t : ctx->ntt_cache) t.release();
  dev_free(ctx->fixed_base[0]);
  dev_free(ctx->fixed_base[1]);
  ps_host_free(ctx->h_stage);
#if PS_GPU
  for (int i = 0; i < 5; i++) if (ctx->ev[i]) cudaEventDestroy((cudaEvent_t)ctx->ev[i]);
  for (int i = 0; i < 6; i++) if (ctx->evp[i]) cudaEventDestroy((cudaEvent_t)ctx->evp[i]);
  if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
  if (ctx->stream2) cudaStreamDestroy(ctx->stream2);
  if (ctx->ev_fork) cudaEventDestroy((cudaEvent_t)ctx->ev_fork);
  if (ctx->ev_join) cudaEventDestroy((cudaEvent_t)ctx->ev_join);
#endif
  delete ctx;
}

// ---- bases ------------------------------------------------------------------------------------------
int ps_bases_load(ps_ctx* ctx, int group, const uint8_t* points, size_t n, int format, int window_bits,
                  int precompute_tables, ps_bases** out) {
  if (!out || (n && !points) || (format != PS_FMT_COMPRESSED && format != PS_FMT_AFFINE)) return PS_ERR_ARG;
  PS_TRY(begin_call(ctx));
  if (group == PS_G1) return bases_load_t<Fp, G1DecodeK>(ctx, points, n, format, window_bits, precompute_tables, out);
  if (group == PS_G2) return bases_load_t<Fp2, G2DecodeK>(ctx, points, n, format, window_bits, precompute_tables, out);
  return PS_ERR_ARG;
}

size_t ps_bases_len(const ps_bases* b) { return b ? b->n : 0; }

int ps_bases_info(const ps_bases* b, int out[4]) {
  if (!b || !out) return PS_ERR_ARG;
  int c = b->c ? b->c : msm_pick_window(b->n ? b->n : 1);
  out[0] = c; out[1] = msm_windows(c); out[2] = b->T; out[3] = b->group;
  return PS_OK;
}

void ps_bases_free(ps_bases* b) {
  if (!b) return;
  dev_free(b->tab);
  delete b;
}

int ps_bases_from_scalars(ps_ctx* ctx, int group, const uint8_t* scalars_be, size_t n, int window_bits,
                          int precompute_tables, ps_bases** out) {
  if (!out || (n && !scalars_be) || (group != PS_G1 && group != PS_G2)) return PS_ERR_ARG;
  PS_TRY(begin_call(ctx));
  ps_bases* b = nullptr;
  PS_TRY(bases_alloc(group, n, window_bits, precompute_tables, ctx->msm_shards, ctx->msm_bucket_cost, &b));
  uint32_t *d_sc = nullptr, *d_err = nullptr;
  int rc = stage_scalars(ctx, scalars_be, n, 0, &d_sc, &d_err);
  if (rc == PS_OK) {
    if (group == PS_G1) {
      const G1Affine* tbl = nullptr;
      rc = fixed_base_table<Fp>(ctx, &tbl);
      G1XYZZ* tmp = ctx->arena.take<G1XYZZ>(n);
      if (rc == PS_OK && !tmp) rc = PS_ERR_ALLOC;
      if (rc == PS_OK) rc = ps_launch<FixedBaseMulK<Fp>>(ctx->stream, n, (const uint32_t*)d_sc, tbl, tmp);
      if (rc == PS_OK) rc = ps_launch<BatchToAffineK<Fp>>(ctx->stream, (n + BatchToAffineK<Fp>::K - 1) / BatchToAffineK<Fp>::K, (uint32_t)n,
                                                       (const G1XYZZ*)tmp, (G1Affine*)b->tab);
      if (rc == PS_OK) rc = bases_finish<Fp>(ctx, b);
    } else {
      const G2Affine* tbl = nullptr;
      rc = fixed_base_table<Fp2>(ctx, &tbl);
      G2XYZZ* tmp = ctx->arena.take<G2XYZZ>(n);
      if (rc == PS_OK && !tmp) rc = PS_ERR_ALLOC;
      if (rc == PS_OK) rc = ps_launch<FixedBaseMulK<Fp2>>(ctx->stream, n, (const uint32_t*)d_sc, tbl, tmp);
      if (rc == PS_OK) rc = ps_launch<BatchToAffineK<Fp2>>(ctx->stream, (n + BatchToAffineK<Fp2>::K - 1) / BatchToAffineK<Fp2>::K, (uint32_t)n,
                                                        (const G2XYZZ*)tmp, (G2Affine*)b->tab);
      if (rc == PS_OK) rc = bases_finish<Fp2>(ctx, b);
    }
  }
  if (rc == PS_OK) rc = check_err_flag(ctx, d_err, PS_ERR_ENCODING);
  if (rc != PS_OK) { ps_bases_free(b); return rc; }
  *out = b;
  return PS_OK;
}

int ps_bases_export(ps_ctx* ctx, const ps_bases* b, size_t first, size_t count, int format, uint8_t* out) {
  if (!b || !out || first + count > b->n || (format != PS_FMT_COMPRESSED && format != PS_FMT_AFFINE)) return PS_ERR_ARG;
  PS_TRY(begin_call(ctx));
  size_t per = point_bytes(b->group, format);
  uint8_t* d_bytes = ctx->arena.take<uint8_t>(count * per);
  if (!d_bytes) return PS_ERR_ALLOC;
  if (b->group == PS_G1) PS_LAUNCH(AffineEncodeK<Fp>, ctx->stream, count, (const G1Affine*)b->tab + first, format, d_bytes);
  else PS_LAUNCH(AffineEncodeK<Fp2>, ctx->stream, count, (const G2Affine*)b->tab + first, format, d_bytes);
  PS_TRY(dev_d2h(out, d_bytes, count * per, ctx->stream));
  return dev_sync(ctx->stream);
}

// ---- MSM --------------------------------------------------------------------------------------------
int ps_msm(ps_ctx* ctx, const ps_bases* b, const uint8_t* scalars_be, size_t n, uint8_t* out) {
  if (!b || !out || (n && !scalars_be)) return PS_ERR_ARG;
  if (n != b->n) return PS_ERR_LENGTH;
  PS_TRY(begin_call(ctx));
  uint32_t *d_sc = nullptr, *d_err = nullptr;
  PS_TRY(stage_scalars(ctx, scalars_be, n, 0, &d_sc, &d_err));
  if (b->group == PS_G1) {
    G1XYZZ* d_res = ctx->arena.take<G1XYZZ>(1);
    if (!d_res) return PS_ERR_ALLOC;
    PS_TRY(msm_on_bases<Fp>(ctx, b, 0, d_sc, n, 0, d_res));
    PS_TRY(encode_points<Fp>(ctx, d_res, 1, out));
  } else {
    G2XYZZ* d_res = ctx->arena.take<G2XYZZ>(1);
    if (!d_res) return PS_ERR_ALLOC;
    PS_TRY(msm_on_bases<Fp2>(ctx, b, 0, d_sc, n, 0, d_res));
    PS_TRY(encode_points<Fp2>(ctx, d_res, 1, out));
  }
  return check_err_flag(ctx, d_err, PS_ERR_ENCODING);
}

int ps_msm_device(ps_ctx* ctx, const ps_bases* b, size_t first, const void* d_scalars_le, size_t n, void* d_out_xyzz) {
  if (!b || !d_out_xyzz || (n && !d_scalars_le)) return PS_ERR_ARG;
  PS_TRY(begin_call(ctx));
  if (b->group == PS_G1) return msm_on_bases<Fp>(ctx, b, first, (const uint32_t*)d_scalars_le, n, 0, (G1XYZZ*)d_out_xyzz);
  return msm_on_bases<Fp2>(ctx, b, first, (const uint32_t*)d_scalars_le, n, 0, (G2XYZZ*)d_out_xyzz);
}

int ps_msm_combine(ps_ctx* ctx, int group, const void* d_partials_xyzz, size_t count, uint8_t* out) {
  if (!d_partials_xyzz || !out || (group != PS_G1 && group != PS_G2)) return PS_ERR_ARG;
  PS_TRY(begin_call(ctx));
  if (group == PS_G1) {
    G1XYZZ* d_res = ctx->arena.take<G1XYZZ>(1);
    if (!d_res) return PS_ERR_ALLOC;
    PS_TRY((launch_coop<MsmSumK, Fp>(ctx->msm_team != 0, ctx->stream, 1, (uint32_t)count, (const G1XYZZ*)d_partials_xyzz, d_res)));
    PS_TRY(encode_points<Fp>(ctx, d_res, 1, out));
  } else {
    G2XYZZ* d_res = ctx->arena.take<G2XYZZ>(1);
    if (!d_res) return PS_ERR_ALLOC;
    PS_TRY((launch_coop<MsmSumK, Fp2>(ctx->msm_team != 0, ctx->stream, 1, (uint32_t)count, (const G2XYZZ*)d_partials_xyzz, d_res)));
    PS_TRY(encode_points<Fp2>(ctx, d_res, 1, out));
  }
  return dev_sync(ctx->stream);
}

int ps_last_msm_timing(ps_ctx* ctx, float out_ms[5]) {
  if (!ctx || !out_ms) return PS_ERR_ARG;
  for (int i = 0; i < 5; i++) out_ms[i] = 0.f;
#if PS_GPU
  if (!ctx->ev_valid) return PS_ERR_ARG;
  PS_CUDA_TRY(cudaEventSynchronize((cudaEvent_t)ctx->ev[4]));
  for (int i = 0; i < 4; i++)
    PS_CUDA_TRY(cudaEventElapsedTime(&out_ms[i], (cudaEvent_t)ctx->ev[i], (cudaEvent_t)ctx->ev[i + 1]));
  PS_CUDA_TRY(cudaEventElapsedTime(&out_ms[4], (cudaEvent_t)ctx->ev[0], (cudaEvent_t)ctx->ev[4]));
#endif
  return PS_OK;
}

int ps_last_prove_timing(ps_ctx* ctx, float out_ms[6]) {
  if (!ctx || !out_ms) return PS_ERR_ARG;
  for (int i = 0; i < 6; i++) out_ms[i] = 0.f;
#if PS_GPU
  if (!ctx->evp_valid) return PS_ERR_ARG;
  PS_CUDA_TRY(cudaEventSynchronize((cudaEvent_t)ctx->evp[5]));
  for (int i = 0; i < 5; i++)
    PS_CUDA_TRY(cudaEventElapsedTime(&out_ms[i], (cudaEvent_t)ctx->evp[i], (cudaEvent_t)ctx->evp[i + 1]));
  PS_CUDA_TRY(cudaEventElapsedTime(&out_ms[5], (cudaEvent_t)ctx->evp[0], (cudaEvent_t)ctx->evp[5]));
#endif
  return PS_OK;
}

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
  if (n_gates & (n_gates - 1)) return PS_ERR_UNSUPPORTED;  // the interpolation tree needs n = 2^k
  if (n_gates > (1u << 26) || n_vars > (1u << 28)) return PS_ERR_UNSUPPORTED;
  PS_TRY(begin_call(ctx));
  ps_stream_t st = ctx->stream;
  ps_qap* q = new (std::nothrow) ps_qap();
  SparseQap* sq = new (std::nothrow) SparseQap();
  if (!q || !sq) { delete q; delete sq; return PS_ERR_ALLOC; }
  q->n = n_gates; q->m = n_vars; q->n_io = n_io; q->dense = false; q->sparse = sq;
  int k = 0;
  while (((size_t)1 << k) < n_gates) k++;
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
    rc = dev_alloc((void**)&m.row_ptr, (n_gates + 1) * 4);
    if (rc == PS_OK) rc = dev_alloc((void**)&m.col, nnz * 4);
    if (rc == PS_OK) rc = dev_alloc((void**)&m.val, nnz * sizeof(Fr));
    if (rc == PS_OK) rc = dev_h2d(m.row_ptr, rps[i], (n_gates + 1) * 4, st);
    if (rc == PS_OK && nnz) rc = dev_h2d(m.col, cols[i], nnz * 4, st);
    uint8_t* d_bytes = ctx->arena.take<uint8_t>(nnz * 32);
    if (rc == PS_OK && !d_bytes) rc = PS_ERR_ALLOC;
    if (rc == PS_OK && nnz) rc = dev_h2d(d_bytes, vals[i], nnz * 32, st);
    if (rc == PS_OK) rc = ps_launch<FrFromBytesK>(st, nnz, (const uint8_t*)d_bytes, (uint32_t*)m.val, 1, d_err);
  }
  Fr* d_z = ctx->arena.take<Fr>(n_gates + 1);
  if (rc == PS_OK && !d_z) rc = PS_ERR_ALLOC;
  if (rc == PS_OK) rc = inv_zprime_build(ctx, sq, (uint32_t)n_gates);
  if (rc == PS_OK) rc = ztree_build(ctx, sq, (uint32_t)n_gates, k, d_z);
  if (rc == PS_OK) rc = twist_tables_build(ctx, sq, (uint32_t)n_gates, k);
  if (rc == PS_OK) rc = series_tables_build(ctx, sq, (uint32_t)n_gates, k, d_z);
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

// ---- Groth16 ------------------------------------------------------------------------------------------
// The proof elements are assembled as three MSMs over concatenated base sets (same group elements as
// groth16.go:146-200, which adds the pieces one scalar multiplication at a time):
//   A = <[a | r | 1],            [Xi  | Delta  | Alpha]>
//   B = <[b | s | 1],            [Xi2 | Delta2 | Beta2]>
//   C = <[w_nio | h | s a + r b | s | r | r s], [NioLP | XiT | Xi | Alpha | Beta | Delta]>
// using  s A + r B1 - r s Delta = sum_k (s a_k + r b_k) Xi_k + s Alpha + r Beta + r s Delta.
int ps_g16_key_load(ps_ctx* ctx, size_t n_gates, size_t n_nio, int format, const uint8_t* xi, const uint8_t* xi2,
                    const uint8_t* xit, const uint8_t* niolp, const uint8_t* alpha, const uint8_t* beta,
                    const uint8_t* delta, const uint8_t* beta2, const uint8_t* delta2, ps_g16_key** key) {
  if (!key || !xi || !xi2 || !xit || (n_nio && !niolp) || !alpha || !beta || !delta || !beta2 || !delta2 || n_gates < 2)
    return PS_ERR_ARG;
  if (format != PS_FMT_COMPRESSED && format != PS_FMT_AFFINE) return PS_ERR_ARG;
  PS_TRY(begin_call(ctx));
  ps_g16_key* k = new (std::nothrow) ps_g16_key();
  if (!k) return PS_ERR_ALLOC;
  k->n = n_gates; k->n_nio = n_nio;
  const uint8_t* pa[3] = {xi, delta, alpha};
  const size_t ca[3] = {n_gates, 1, 1};
  const uint8_t* pb[3] = {xi2, delta2, beta2};
  const uint8_t* pc[6] = {niolp, xit, xi, alpha, beta, delta};
  const size_t cc[6] = {n_nio, n_gates - 1, n_gates, 1, 1, 1};
  int rc = bases_concat<Fp, G1DecodeK>(ctx, format, pa, ca, 3, &k->A);
  if (rc == PS_OK) rc = ctx->arena.reset();
  if (rc == PS_OK) rc = bases_concat<Fp2, G2DecodeK>(ctx, format, pb, ca, 3, &k->B);
  if (rc == PS_OK) rc = ctx->arena.reset();
  if (rc == PS_OK) rc = bases_concat<Fp, G1DecodeK>(ctx, format, pc, cc, 6, &k->C);
  if (rc != PS_OK) { ps_g16_key_free(k); return rc; }
  *key = k;
  return PS_OK;
}

void ps_g16_key_free(ps_g16_key* key) {
  if (!key) return;
  ps_bases_free(key->A); ps_bases_free(key->B); ps_bases_free(key->C);
  delete key;
}

namespace {
// Fr vectors in Montgomery form -> standard form, in place
struct FrFromMontK {
  static constexpr int BLOCK = 256;
  PS_DEV static void run(uint32_t i, Fr* a) { a[i] = a[i].from_mont(); }
};

struct G16Scalars { Fr *scA, *scB, *scC; size_t nA, nB, nC; QuotientBufs qb; };

// quotient + the three MSM scalar vectors of the Groth16 proof (Montgomery form, arena memory)
int g16_build_scalars(ps_ctx* ctx, const ps_g16_key* key, const ps_qap* qap, const uint8_t* witness_be, const uint8_t* r_be,
                      const uint8_t* s_be, G16Scalars* o) {
  if (key->n != qap->n || key->n_nio != qap->n_io) return PS_ERR_LENGTH;
  ps_stream_t st = ctx->stream;
  const size_t n = qap->n, nio = qap->n_io, diff = qap->m - qap->n_io;
  PS_TRY(run_quotient(ctx, qap, witness_be, &o->qb));
  Fr hrs[2];
  for (int t = 0; t < 2; t++) {
    const uint8_t* src = t == 0 ? r_be : s_be;
    Fr x;
    for (int j = 0; j < 8; j++) {
      const uint8_t* p = src + 4 * (7 - j);
      x.v[j] = ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | (uint32_t)p[3];
    }
    if (!limbs_lt_mod<FrParams>(x.v)) return PS_ERR_ENCODING;
    hrs[t] = x.to_mont();
  }
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
}  // namespace

int ps_g16_prove(ps_ctx* ctx, const ps_g16_key* key, const ps_qap* qap, const uint8_t* witness_be, const uint8_t* r_be,
                 const uint8_t* s_be, uint8_t* outA, uint8_t* outB, uint8_t* outC, uint8_t* out_h) {
  if (!key || !qap || !witness_be || !r_be || !s_be || !outA || !outB || !outC) return PS_ERR_ARG;
  PS_TRY(begin_call(ctx));
  G16Scalars sc;
  ctx->evp_valid = false;
  PS_TRY(ctx_prove_event(ctx, 0));
  PS_TRY(g16_build_scalars(ctx, key, qap, witness_be, r_be, s_be, &sc));
  PS_TRY(ctx_prove_event(ctx, 1));
  G1XYZZ* resG1 = ctx->arena.take<G1XYZZ>(2);
  G2XYZZ* resG2 = ctx->arena.take<G2XYZZ>(1);
  if (!resG1 || !resG2) return PS_ERR_ALLOC;
  // the G2 MSM is independent of the two G1 ones: it runs on the secondary stream so that its serial
  // tails and its register-bound accumulate kernel overlap with the G1 work
  PS_TRY(ctx_fork(ctx));
  {
    SecondaryScope scope(ctx);
    PS_TRY(msm_on_bases<Fp2>(ctx, key->B, 0, (const uint32_t*)sc.scB, sc.nB, 1, resG2));
    PS_TRY(encode_points_staged<Fp2>(ctx, resG2, 1, 96));   // its inversion chain overlaps the G1 work too
  }
  PS_TRY(msm_on_bases<Fp>(ctx, key->A, 0, (const uint32_t*)sc.scA, sc.nA, 1, resG1));
  PS_TRY(ctx_prove_event(ctx, 2));
  PS_TRY(msm_on_bases<Fp>(ctx, key->C, 0, (const uint32_t*)sc.scC, sc.nC, 1, resG1 + 1));
  PS_TRY(ctx_prove_event(ctx, 3));
  PS_TRY(encode_points_staged<Fp>(ctx, resG1, 2, 0));
  PS_TRY(ctx_prove_event(ctx, 4));
  PS_TRY(ctx_join(ctx));
  PS_TRY(ctx_prove_event(ctx, 5));
  ctx->evp_valid = true;
  if (out_h) PS_TRY(export_fr(ctx, sc.qb.h, qap->n - 1, out_h));
  PS_TRY(check_err_flag(ctx, sc.qb.enc_err, PS_ERR_ENCODING));
  PS_TRY(check_err_flag(ctx, sc.qb.flag, PS_ERR_REMAINDER));   // synchronises the primary stream, which has joined the second
  memcpy(outA, ctx->h_stage, 48);
  memcpy(outC, ctx->h_stage + 48, 48);
  memcpy(outB, ctx->h_stage + 96, 96);
  return PS_OK;
}

size_t ps_g16_scalar_count(const ps_g16_key* key, int which) {
  if (!key) return 0;
  const ps_bases* b = which == 0 ? key->A : (which == 1 ? key->C : (which == 2 ? key->B : nullptr));
  return b ? b->n : 0;
}

const ps_bases* ps_g16_key_bases(const ps_g16_key* key, int which) {
  if (!key) return nullptr;
  return which == 0 ? key->A : (which == 1 ? key->C : (which == 2 ? key->B : nullptr));
}

int ps_g16_msm_partials(ps_ctx* ctx, const ps_g16_key* key, const void* d_scA, const void* d_scC, const void* d_scB,
                        const size_t first[3], const size_t count[3], void* d_partials) {
  if (!ctx || !key || !d_scA || !d_scC || !d_scB || !first || !count || !d_partials) return PS_ERR_ARG;
  PS_TRY(begin_call(ctx));
  uint8_t* out = (uint8_t*)d_partials;  // [A: 192 B | C: 192 B | B: 384 B]
  PS_TRY(ctx_fork(ctx));
  {
    SecondaryScope scope(ctx);
    PS_TRY(msm_on_bases<Fp2>(ctx, key->B, first[2], (const uint32_t*)d_scB, count[2], 0, (G2XYZZ*)(out + 384)));
  }
  PS_TRY(msm_on_bases<Fp>(ctx, key->A, first[0], (const uint32_t*)d_scA, count[0], 0, (G1XYZZ*)out));
  PS_TRY(msm_on_bases<Fp>(ctx, key->C, first[1], (const uint32_t*)d_scC, count[1], 0, (G1XYZZ*)(out + 192)));
  return ctx_join(ctx);
}

// One aggregate polynomial of the sparse QAP in coefficient form (which = 0: a, 1: b): the SpMV and
// the interpolation on {1..n} for that polynomial only, so that two GPUs can share the work.
int ps_qap_aggregate_one(ps_ctx* ctx, const ps_qap* qap, const uint8_t* witness_be, int which, void* d_out_coef) {
  if (!ctx || !qap || !witness_be || !d_out_coef || which < 0 || which > 1) return PS_ERR_ARG;
  if (qap->dense) return PS_ERR_UNSUPPORTED;
  PS_TRY(begin_call(ctx));
  const SparseQap* sq = (const SparseQap*)qap->sparse;
  const uint32_t n = (uint32_t)qap->n;
  ps_stream_t st = ctx->stream;
  uint32_t *d_w = nullptr, *d_err = nullptr;
  PS_TRY(stage_scalars(ctx, witness_be, qap->m, 1, &d_w, &d_err));
  Fr* ev = ctx->arena.take<Fr>((size_t)3 * n);
  if (!ev) return PS_ERR_ALLOC;
  PS_LAUNCH(SpmvK, st, (size_t)3 * n, n, 0u, (const uint32_t*)sq->mat[0].row_ptr, (const uint32_t*)sq->mat[0].col, (const Fr*)sq->mat[0].val,
            (const uint32_t*)sq->mat[1].row_ptr, (const uint32_t*)sq->mat[1].col, (const Fr*)sq->mat[1].val,
            (const uint32_t*)sq->mat[2].row_ptr, (const uint32_t*)sq->mat[2].col, (const Fr*)sq->mat[2].val, (const Fr*)d_w, ev);
  PS_TRY(interpolate_ap(ctx, sq, n, qap->log_np, 1, ev + (size_t)which * n, (Fr*)d_out_coef));
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
  const uint32_t n = (uint32_t)qap->n;
  const size_t nio = qap->n_io, diff = qap->m - qap->n_io;
  uint32_t *d_w = nullptr, *d_err = nullptr;
  PS_TRY(stage_scalars(ctx, witness_be, qap->m, 1, &d_w, &d_err));
  Fr* w = (Fr*)d_w;
  Fr* ev = ctx->arena.take<Fr>((size_t)3 * n);
  Fr* h = ctx->arena.take<Fr>(n);
  uint32_t* flag = ctx->arena.take<uint32_t>(1);
  if (!ev || !h || !flag) return PS_ERR_ALLOC;
  PS_TRY(dev_memset(flag, 0, 4, st));
  PS_LAUNCH(SpmvK, st, (size_t)3 * n, n, 0u, (const uint32_t*)sq->mat[0].row_ptr, (const uint32_t*)sq->mat[0].col, (const Fr*)sq->mat[0].val,
            (const uint32_t*)sq->mat[1].row_ptr, (const uint32_t*)sq->mat[1].col, (const Fr*)sq->mat[1].val,
            (const uint32_t*)sq->mat[2].row_ptr, (const uint32_t*)sq->mat[2].col, (const Fr*)sq->mat[2].val, (const Fr*)w, ev);
  PS_LAUNCH(GateCheckK, st, n, n, (const Fr*)ev, flag);
  const Fr* a = (const Fr*)d_a;
  const Fr* b = (const Fr*)d_b;
  PS_TRY(quotient_series(ctx, sq, n, qap->log_np, a, b, h, (Fr*)nullptr));
  Fr hrs[2];
  for (int t = 0; t < 2; t++) {
    const uint8_t* src = t == 0 ? r_be : s_be;
    Fr x;
    for (int j = 0; j < 8; j++) {
      const uint8_t* p = src + 4 * (7 - j);
      x.v[j] = ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | (uint32_t)p[3];
    }
    if (!limbs_lt_mod<FrParams>(x.v)) return PS_ERR_ENCODING;
    hrs[t] = x.to_mont();
  }
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

// ---- Groth16 over several GPUs: quotient split by subtree, MSMs overlapped with the division ---------------
namespace {
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
int parse_fr(const uint8_t* src, Fr* out) {
  Fr x;
  for (int j = 0; j < 8; j++) {
    const uint8_t* p = src + 4 * (7 - j);
    x.v[j] = ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | (uint32_t)p[3];
  }
  if (!limbs_lt_mod<FrParams>(x.v)) return PS_ERR_ENCODING;
  *out = x.to_mont();
  return PS_OK;
}
int log2_exact(size_t v) {
  int l = 0;
  while (((size_t)1 << l) < v) l++;
  return ((size_t)1 << l) == v ? l : -1;
}
}  // namespace

int ps_host_alloc(size_t bytes, void** out) {
  if (!out) return PS_ERR_ARG;
#if PS_GPU
  PS_CUDA_TRY(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocPortable));
#else
  *out = malloc(bytes ? bytes : 1);
  if (!*out) return PS_ERR_ALLOC;
#endif
  return PS_OK;
}

void ps_host_free(void* p) {
  if (!p) return;
#if PS_GPU
  cudaFreeHost(p);
#else
  free(p);
#endif
}

namespace {
// body of ps_qap_interp_part once the witness is on the device (Montgomery form); d_err may be null
int interp_part_run(ps_ctx* ctx, const ps_qap* qap, const Fr* d_w, const uint32_t* d_err, int which, size_t part, size_t parts,
                    int lp, void* d_out_evals, void* d_w_nio_out, void* d_status) {
  const SparseQap* sq = (const SparseQap*)qap->sparse;
  const uint32_t n = (uint32_t)qap->n, ns = (uint32_t)(qap->n / parts), lo = (uint32_t)part * ns;
  ps_stream_t st = ctx->stream;
  if (d_w_nio_out) PS_LAUNCH(FrStdCopyK, st, qap->n_io, d_w + (qap->m - qap->n_io), (Fr*)d_w_nio_out);
  Fr* ev = ctx->arena.take<Fr>((size_t)3 * ns);
  Fr* E0 = ctx->arena.take<Fr>((size_t)2 * ns);
  uint32_t* flag = ctx->arena.take<uint32_t>(1);
  if (!ev || !E0 || !flag) return PS_ERR_ALLOC;
  PS_TRY(dev_memset(flag, 0, 4, st));
  PS_LAUNCH(SpmvK, st, (size_t)3 * ns, ns, lo, (const uint32_t*)sq->mat[0].row_ptr, (const uint32_t*)sq->mat[0].col, (const Fr*)sq->mat[0].val,
            (const uint32_t*)sq->mat[1].row_ptr, (const uint32_t*)sq->mat[1].col, (const Fr*)sq->mat[1].val,
            (const uint32_t*)sq->mat[2].row_ptr, (const uint32_t*)sq->mat[2].col, (const Fr*)sq->mat[2].val, d_w, ev);
  PS_LAUNCH(GateCheckK, st, ns, ns, (const Fr*)ev, flag);   // this rank's gates; every gate is checked by some rank
  PS_LAUNCH(InterpLeafK, st, ns, ns, (const Fr*)(ev + (size_t)which * ns), (const Fr*)(sq->inv_zprime + lo), E0);
  // parts == 1: the whole tree, d_out_evals receives the n coefficients
  PS_TRY(interpolate_levels(ctx, sq, n, qap->log_np, 1, lo, ns, 0, qap->log_np - lp, E0, lp ? (Fr*)d_out_evals : (Fr*)nullptr,
                            lp ? (Fr*)nullptr : (Fr*)d_out_evals));
  PS_LAUNCH(StatusMergeK, st, 1, d_err, (const uint32_t*)flag, (uint32_t*)d_status);
  return PS_OK;
}
}  // namespace

int ps_qap_interp_part(ps_ctx* ctx, const ps_qap* qap, const uint8_t* witness_be, int which, size_t part, size_t parts,
                       void* d_out_evals, void* d_w_nio_out, void* d_status) {
  if (!ctx || !qap || !witness_be || !d_out_evals || !d_status || which < 0 || which > 1) return PS_ERR_ARG;
  if (qap->dense) return PS_ERR_UNSUPPORTED;
  const int lp = log2_exact(parts);
  if (lp < 0 || parts > qap->n / 2 || part >= parts) return PS_ERR_ARG;
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
  if (lp < 0 || parts > qap->n / 2 || part >= parts) return PS_ERR_ARG;
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
  if (lp < 1 || parts > qap->n / 2) return PS_ERR_ARG;
  PS_TRY(begin_call(ctx));
  const SparseQap* sq = (const SparseQap*)qap->sparse;
  const uint32_t n = (uint32_t)qap->n;
  Fr* E0 = ctx->arena.take<Fr>((size_t)2 * n);
  if (!E0) return PS_ERR_ALLOC;
  PS_TRY(dev_d2d(E0, d_evals_all, (size_t)2 * n * sizeof(Fr), ctx->stream));
  return interpolate_levels(ctx, sq, n, qap->log_np, 1, 0, n, qap->log_np - lp, qap->log_np, E0, (Fr*)nullptr, (Fr*)d_out_coef);
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

extern "C++" {
namespace {
// Sums of the per-rank partial points of a sharded proof, read in place from the gathered records
// (stride bytes apart): item i adds, over all records, the points at byte offsets off0[i] and (if
// >= 0) off1[i].  One team of four lanes per item.
template <class F>
struct RecordSumK {
  static constexpr int BLOCK = 32;
  PS_DEV static void run(uint32_t tid, uint32_t count, const uint8_t* recs, uint32_t stride, int a0, int a1, int b0, int b1,
                         XYZZ<F>* out) {
    Coop<true> co(tid);
    if (co.idle()) return;
    const int o0 = tid == 0 ? a0 : b0, o1 = tid == 0 ? a1 : b1;
    XYZZ<F> r = XYZZ<F>::inf();
    for (uint32_t i = 0; i < count; i++) {
      const uint8_t* rec = recs + (size_t)i * stride;
      co.add(r, *(const XYZZ<F>*)(rec + o0));
      if (o1 >= 0) co.add(r, *(const XYZZ<F>*)(rec + o1));
    }
    if (co.writer()) out[tid] = r;
  }
};
}  // namespace
}  // extern "C++"

int ps_g16_combine(ps_ctx* ctx, const void* d_records, size_t count, size_t stride, uint8_t* outA, uint8_t* outB, uint8_t* outC) {
  if (!ctx || !d_records || !count || !outA || !outB || !outC || stride < 960 || (stride & 15)) return PS_ERR_ARG;
  PS_TRY(begin_call(ctx));
  const uint8_t* recs = (const uint8_t*)d_records;
  G1XYZZ* resG1 = ctx->arena.take<G1XYZZ>(2);
  if (!resG1) return PS_ERR_ALLOC;
  PS_TRY(ctx_fork(ctx));
  {
    SecondaryScope scope(ctx);
    G2XYZZ* resG2 = ctx->arena.take<G2XYZZ>(1);
    if (!resG2) return PS_ERR_ALLOC;
    PS_LAUNCH(RecordSumK<Fp2>, ctx->stream, (size_t)TEAM, (uint32_t)count, recs, (uint32_t)stride, 384, -1, 384, -1, resG2);
    PS_TRY(encode_points_staged<Fp2>(ctx, resG2, 1, 96));
  }
  PS_LAUNCH(RecordSumK<Fp>, ctx->stream, (size_t)2 * TEAM, (uint32_t)count, recs, (uint32_t)stride, 0, -1, 192, 768, resG1);
  PS_TRY(encode_points_staged<Fp>(ctx, resG1, 2, 0));
  PS_TRY(ctx_join(ctx));
  PS_TRY(dev_sync(ctx->stream));
  memcpy(outA, ctx->h_stage, 48);
  memcpy(outC, ctx->h_stage + 48, 48);
  memcpy(outB, ctx->h_stage + 96, 96);
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

// ---- PHGR13 -------------------------------------------------------------------------------------------
int ps_phgr13_key_load(ps_ctx* ctx, size_t n_gates, size_t n_mid, int format, const uint8_t* gsi, const uint8_t* vs,
                       const uint8_t* ws, const uint8_t* ys, const uint8_t* vas, const uint8_t* was, const uint8_t* yas,
                       const uint8_t* vbs, const uint8_t* wbs, const uint8_t* ybs, ps_phgr13_key** key) {
  if (!key || !gsi || n_gates < 2) return PS_ERR_ARG;
  if (n_mid && (!vs || !ws || !ys || !vas || !was || !yas || !vbs || !wbs || !ybs)) return PS_ERR_ARG;
  if (format != PS_FMT_COMPRESSED && format != PS_FMT_AFFINE) return PS_ERR_ARG;
  PS_TRY(begin_call(ctx));
  ps_phgr13_key* k = new (std::nothrow) ps_phgr13_key();
  if (!k) return PS_ERR_ALLOC;
  k->n = n_gates; k->n_mid = n_mid;
  const uint8_t* singles[6] = {gsi, vs, ys, vas, was, yas};
  const size_t counts[6] = {n_gates - 1, n_mid, n_mid, n_mid, n_mid, n_mid};
  int rc = PS_OK;
  for (int i = 0; i < 6 && rc == PS_OK; i++) {
    rc = bases_concat<Fp, G1DecodeK>(ctx, format, &singles[i], &counts[i], 1, &k->g1[i]);
    if (rc == PS_OK) rc = ctx->arena.reset();
  }
  const uint8_t* zs[3] = {vbs, wbs, ybs};
  const size_t zc[3] = {n_mid, n_mid, n_mid};
  if (rc == PS_OK) rc = bases_concat<Fp, G1DecodeK>(ctx, format, zs, zc, 3, &k->g1[6]);
  if (rc == PS_OK) rc = ctx->arena.reset();
  if (rc == PS_OK) rc = bases_concat<Fp2, G2DecodeK>(ctx, format, &ws, &n_mid, 1, &k->ws);
  if (rc != PS_OK) { ps_phgr13_key_free(k); return rc; }
  *key = k;
  return PS_OK;
}

void ps_phgr13_key_free(ps_phgr13_key* key) {
  if (!key) return;
  for (auto* b : key->g1) ps_bases_free(b);
  ps_bases_free(key->ws);
  delete key;
}

int ps_phgr13_prove(ps_ctx* ctx, const ps_phgr13_key* key, const ps_qap* qap, const uint8_t* witness_be, uint8_t* out432,
                    uint8_t* out_h) {
  if (!key || !qap || !witness_be || !out432) return PS_ERR_ARG;
  if (key->n != qap->n || key->n_mid != qap->n_io) return PS_ERR_LENGTH;
  PS_TRY(begin_call(ctx));
  ps_stream_t st = ctx->stream;
  const size_t n = qap->n, nmid = qap->n_io, diff = qap->m - qap->n_io;
  QuotientBufs qb;
  PS_TRY(run_quotient(ctx, qap, witness_be, &qb));
  Fr* w3 = ctx->arena.take<Fr>(3 * nmid);
  G1XYZZ* resG1 = ctx->arena.take<G1XYZZ>(7);
  G2XYZZ* resG2 = ctx->arena.take<G2XYZZ>(1);
  if (!w3 || !resG1 || !resG2) return PS_ERR_ALLOC;
  const Fr* wmid = qb.w + diff;
  for (int t = 0; t < 3; t++) PS_LAUNCH(FrCopyK, st, nmid, wmid, w3 + t * nmid);
  // wss (G2) is independent of the seven G1 sums: second stream, like Groth16's B
  PS_TRY(ctx_fork(ctx));
  {
    SecondaryScope scope(ctx);
    PS_TRY(msm_on_bases<Fp2>(ctx, key->ws, 0, (const uint32_t*)wmid, nmid, 1, resG2));               // wss
    PS_TRY(encode_points_staged<Fp2>(ctx, resG2, 1, 7 * 48));
  }
  PS_TRY(msm_on_bases<Fp>(ctx, key->g1[0], 0, (const uint32_t*)qb.h, n - 1, 1, resG1));            // hs
  for (int i = 1; i < 6; i++)                                                                         // vss yss vass wass yass
    PS_TRY(msm_on_bases<Fp>(ctx, key->g1[i], 0, (const uint32_t*)wmid, nmid, 1, resG1 + i));
  PS_TRY(msm_on_bases<Fp>(ctx, key->g1[6], 0, (const uint32_t*)w3, 3 * nmid, 1, resG1 + 6));         // gz
  PS_TRY(encode_points_staged<Fp>(ctx, resG1, 7, 0));
  PS_TRY(ctx_join(ctx));
  if (out_h) PS_TRY(export_fr(ctx, qb.h, n - 1, out_h));
  PS_TRY(check_err_flag(ctx, qb.enc_err, PS_ERR_ENCODING));
  PS_TRY(check_err_flag(ctx, qb.flag, PS_ERR_REMAINDER));
  memcpy(out432, ctx->h_stage, 432);
  return PS_OK;
}

// ---- measurement ----------------------------------------------------------------------------------------
int ps_bench_intpipe(ps_ctx* ctx, int variant, int iters, double* inst_per_s, double* ms_out) {
  if (!ctx || !inst_per_s || variant < 0 || variant > 3 || iters < 1) return PS_ERR_ARG;
#if PS_GPU
  PS_TRY(begin_call(ctx));
  uint32_t* d_out = ctx->arena.take<uint32_t>(4);
  if (!d_out) return PS_ERR_ALLOC;
  const int blocks = ctx->sm_count * 8, threads = 256;
  cudaEvent_t e0, e1;
  PS_CUDA_TRY(cudaEventCreate(&e0)); PS_CUDA_TRY(cudaEventCreate(&e1));
  for (int rep = 0; rep < 2; rep++) {  // first pass warms up
    PS_CUDA_TRY(cudaEventRecord(e0, ctx->stream));
    switch (variant) {
      case 0: k_intpipe<0><<<blocks, threads, 0, ctx->stream>>>(d_out, iters, 12345u); break;
      case 1: k_intpipe<1><<<blocks, threads, 0, ctx->stream>>>(d_out, iters, 12345u); break;
      case 2: k_intpipe<2><<<blocks, threads, 0, ctx->stream>>>(d_out, iters, 12345u); break;
      default: k_intpipe<3><<<blocks, threads, 0, ctx->stream>>>(d_out, iters, 12345u); break;
    }
    PS_CUDA_TRY(cudaEventRecord(e1, ctx->stream));
    PS_CUDA_TRY(cudaEventSynchronize(e1));
    launch_counter()++;
  }
  float ms = 0;
  PS_CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  double inst = (double)blocks * threads * (double)iters * intpipe_inst_per_iter(variant);
  *inst_per_s = inst / (ms * 1e-3);
  if (ms_out) *ms_out = ms;
  return PS_OK;
#else
  (void)ms_out;
  return PS_ERR_UNSUPPORTED;
#endif
}

int ps_bench_fieldmul(ps_ctx* ctx, int field, int iters, double* mul_per_s, double* ms_out) {
  if (!ctx || !mul_per_s || field < 0 || field > 1 || iters < 1) return PS_ERR_ARG;
#if PS_GPU
  PS_TRY(begin_call(ctx));
  const int blocks = ctx->sm_count * 4, threads = 256;
  const size_t nthreads = (size_t)blocks * threads;
  void* d_io = ctx->arena.take<Fp>(2 * nthreads);
  if (!d_io) return PS_ERR_ALLOC;
  PS_TRY(dev_memset(d_io, 0x5a, 2 * nthreads * sizeof(Fp), ctx->stream));
  cudaEvent_t e0, e1;
  PS_CUDA_TRY(cudaEventCreate(&e0)); PS_CUDA_TRY(cudaEventCreate(&e1));
  for (int rep = 0; rep < 2; rep++) {
    PS_CUDA_TRY(cudaEventRecord(e0, ctx->stream));
    if (field == 0) k_fieldmul<Fr><<<blocks, threads, 0, ctx->stream>>>((Fr*)d_io, iters);
    else k_fieldmul<Fp><<<blocks, threads, 0, ctx->stream>>>((Fp*)d_io, iters);
    PS_CUDA_TRY(cudaEventRecord(e1, ctx->stream));
    PS_CUDA_TRY(cudaEventSynchronize(e1));
    launch_counter()++;
  }
  float ms = 0;
  PS_CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  *mul_per_s = (double)nthreads * 2.0 * iters / (ms * 1e-3);
  if (ms_out) *ms_out = ms;
  return PS_OK;
#else
  (void)ms_out;
  return PS_ERR_UNSUPPORTED;
#endif
}

}  // extern "C"
