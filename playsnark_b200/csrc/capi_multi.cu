// C ABI, part 3: ONE host call that proves (or sums an MSM) on several GPUs of one box.
//
// A ps_mctx owns one ps_ctx and one worker thread per device.  A multi-GPU call posts the same job to
// every worker; the devices exchange data directly over NVLink / NVSwitch peer memory
// (cudaMemcpyPeerAsync into buffers the receiver owns, ordered by CUDA events that the receivers wait
// on in-stream) -- no NCCL, no second process, nothing for the Go caller of Groth16Prove
// (groth16.go:122) to choreograph.  The worker threads meet at host barriers only to make sure an event
// has been RECORDED before a peer enqueues its wait; the GPUs themselves never wait for the host.
//
// Groth16 over N = 2 * parts devices (parts a power of two; sparse QAP).  Device d works on polynomial
// g = d / parts (0: a, 1: b), subtree part = d % parts:
//   W  every device uploads 1/N of the witness and pushes its slice to all peers (one upload per box)
//   I  SpMV + gate check on its n/parts gates, interpolation subtree up to one node (ps_qap_interp_part_dev)
//   R  roots pushed to the devices of the same polynomial; each of them runs the top log2(parts) levels
//      (redundantly: same latency as one leader, and no broadcast of 32 MB vectors afterwards)
//   X  pairwise swap d <-> d + parts: afterwards every device holds a and b
//   H  device 0 divides (h = floor(a b / z)) with a smaller MSM share and pushes each device its slice of h
//   S  its slices of the scalar vectors; B_d (G2) starts at once on the second, high-priority stream and runs while
//      device 0 divides; when h_d has arrived, A_d and the whole of C_d ([w_nio | h | s a + r b | s r rs] against
//      [NioLP_d | XiT_d | Xi_d | ...]) run as ONE G1 pipeline with two outputs, B_d's tail hidden beneath it
//   L  the 976-byte records are pushed to device 0, which adds and encodes.
// Other device counts (odd, or a dense QAP): device 0 computes all scalar vectors and pushes slices.
// The key is SHARDED: device d holds only its index ranges of Xi, Xi2, XiT, NioLP (with all window tables).
#include "group_ops.cuh"
#include "multi_api.cuh"

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <new>
#include <thread>
#include <vector>

using namespace ps;

namespace {

struct HostBarrier {
  std::mutex mu;
  std::condition_variable cv;
  int count = 0, waiting = 0;
  uint64_t gen = 0;
  void arrive() {
    std::unique_lock<std::mutex> lk(mu);
    const uint64_t g = gen;
    if (++waiting == count) { waiting = 0; gen++; cv.notify_all(); return; }
    cv.wait(lk, [&] { return gen != g; });
  }
};

struct Worker {
  std::thread th;
  std::mutex mu;
  std::condition_variable cv;
  std::function<int(int)> job;
  bool has_job = false, done = true, quit = false;
  int rc = PS_OK;
};

constexpr int TL_MAX = 12;   // timeline marks per device

struct ProveWs {            // per-device workspace of the sharded prover (allocated once per (n, m) shape)
  size_t n = 0, m = 0, parts = 0, nx = 0, nt = 0, nn = 0;   // shape the buffers were sized for
  Fr* wfull = nullptr;      // witness, Montgomery (m rounded up to a multiple of ndev)
  Fr* e_all = nullptr;      // subtree roots of this device's polynomial, parts x rows
  Fr* coef[2] = {nullptr, nullptr};
  Fr *scA = nullptr, *scB = nullptr, *scC = nullptr;
  Fr* h_full = nullptr;     // device 0
  uint8_t* recs = nullptr;  // device 0: ndev records of 976 B
  uint8_t* rec = nullptr;   // this device's record
  uint32_t* status = nullptr;
  void release() {
    dev_free(wfull); dev_free(e_all); dev_free(coef[0]); dev_free(coef[1]); dev_free(scA); dev_free(scB); dev_free(scC);
    dev_free(h_full); dev_free(recs); dev_free(rec); dev_free(status);
    *this = ProveWs();
  }
};

constexpr size_t REC_BYTES = 976;   // [A 192 | C early 192 | B 384 | C late 192 | status 4 | pad 12]

}  // namespace

struct ps_mctx {
  int ndev = 0;
  std::vector<int> devs;
  std::vector<ps_ctx*> ctx;
  std::vector<Worker*> workers;
  HostBarrier bar;
  std::atomic<int> failed{0};
  std::vector<ProveWs> ws;
  // exchange events, one set per device (created on that device): witness, roots, coefficients, h, record
  std::vector<void*> ev[5];
  // timeline of the last proof: timing events per device
  std::vector<void*> tl[TL_MAX];
  int tl_count = 0;
  float rank0_share = 0.f;   // 0 = automatic
};

struct ps_mg16_key {
  size_t n = 0, n_nio = 0;
  std::vector<ps_g16_key*> part;
  std::vector<KeySlice> slice;
};

struct ps_mqap {
  size_t n = 0, m = 0, n_io = 0;
  bool dense = false;
  std::vector<ps_qap*> part;   // replicated: one per device (dense: device 0 only)
};

struct ps_mbases {
  int group = 0;
  size_t n = 0;
  std::vector<ps_bases*> part;
  std::vector<size_t> lo;      // ndev + 1 cut points
};

namespace {

// (the Worker is handed over by pointer: m->workers is still growing in ps_mctx_create while the first threads start,
// and indexing it from here raced with its reallocation -- a worker could end up waiting on a stale object)
void worker_main(ps_mctx* m, int d, Worker* w) {
#if PS_GPU
  cudaSetDevice(m->devs[d]);
#endif
  for (;;) {
    std::function<int(int)> job;
    {
      std::unique_lock<std::mutex> lk(w->mu);
      w->cv.wait(lk, [&] { return w->has_job || w->quit; });
      if (w->quit) return;
      job = w->job;
      w->has_job = false;
    }
    int rc = job(d);
    {
      std::lock_guard<std::mutex> lk(w->mu);
      w->rc = rc;
      w->done = true;
    }
    w->cv.notify_all();
  }
}

// runs fn(d) on every device's worker thread; returns the first non-zero status
int run_all(ps_mctx* m, const std::function<int(int)>& fn) {
  m->failed.store(0);
  for (int d = 0; d < m->ndev; d++) {
    Worker* w = m->workers[d];
    {
      std::lock_guard<std::mutex> lk(w->mu);
      w->job = fn; w->has_job = true; w->done = false;
    }
    w->cv.notify_all();
  }
  int rc = PS_OK;
  for (int d = 0; d < m->ndev; d++) {
    Worker* w = m->workers[d];
    std::unique_lock<std::mutex> lk(w->mu);
    w->cv.wait(lk, [&] { return w->done; });
    if (rc == PS_OK && w->rc != PS_OK) rc = w->rc;
  }
  return rc;
}

// a stage of a multi-device job: skipped once any device has failed (every thread still reaches every barrier)
#define STAGE(expr)                                              \
  do {                                                           \
    if (!m->failed.load()) {                                     \
      int _rc = (expr);                                          \
      if (_rc != PS_OK) { my_rc = _rc; m->failed.store(1); }     \
    }                                                            \
  } while (0)

int ev_record(ps_mctx* m, int which, int d) {
#if PS_GPU
  PS_CUDA_TRY(cudaEventRecord((cudaEvent_t)m->ev[which][d], m->ctx[d]->stream));
#else
  (void)m; (void)which; (void)d;
#endif
  return PS_OK;
}
int ev_wait(ps_mctx* m, int which, int from, int d) {
#if PS_GPU
  PS_CUDA_TRY(cudaStreamWaitEvent(m->ctx[d]->stream, (cudaEvent_t)m->ev[which][from], 0));
#else
  (void)m; (void)which; (void)from; (void)d;
#endif
  return PS_OK;
}
int tl_mark(ps_mctx* m, int k, int d) {
#if PS_GPU
  if (k < TL_MAX) PS_CUDA_TRY(cudaEventRecord((cudaEvent_t)m->tl[k][d], m->ctx[d]->stream));
#else
  (void)m; (void)k; (void)d;
#endif
  return PS_OK;
}
// device-to-device copy into memory owned by device `to`, enqueued on device `from`'s stream
int peer_copy(ps_mctx* m, void* dst, int to, const void* src, int from, size_t bytes) {
  if (!bytes) return PS_OK;
#if PS_GPU
  if (to == from) PS_CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, m->ctx[from]->stream));
  else PS_CUDA_TRY(cudaMemcpyPeerAsync(dst, m->devs[to], src, m->devs[from], bytes, m->ctx[from]->stream));
#else
  (void)m; (void)to; (void)from;
  memmove(dst, src, bytes);
#endif
  return PS_OK;
}

// contiguous slices of [0, count) with sizes proportional to the weights
std::vector<size_t> weighted_cuts(size_t count, const std::vector<double>& w) {
  double total = 0, acc = 0;
  for (double x : w) total += x;
  std::vector<size_t> cuts(w.size() + 1, 0);
  for (size_t i = 0; i < w.size(); i++) {
    acc += w[i];
    size_t c = (size_t)((double)count * acc / total + 0.5);
    if (c > count) c = count;
    if (c < cuts[i]) c = cuts[i];
    cuts[i + 1] = c;
  }
  cuts[w.size()] = count;
  return cuts;
}

std::vector<double> shard_weights(const ps_mctx* m) {
  // device 0 also divides (about 1/16 of the single-GPU MSM time) while the others already run B_d: its MSM share
  // shrinks with the device count (8 GPUs, 2^20 constraints: 11.70 / 11.59 / 11.50 ms at 25 / 40 / 55 %, 13.1 at 70 %)
  std::vector<double> w((size_t)m->ndev, 1.0);
  if (m->ndev > 1) {
    double s0 = m->rank0_share > 0.f ? (double)m->rank0_share : 1.0 - 0.06 * m->ndev;
    if (s0 < 0.2) s0 = 0.2;
    w[0] = s0;
  }
  return w;
}

int ws_prepare(ps_mctx* m, int d, size_t n, size_t np, size_t mvars, size_t nio, const KeySlice& sl, size_t parts) {
  ProveWs& w = m->ws[d];
  const size_t nx = sl.x_hi - sl.x_lo, nt = sl.t_hi - sl.t_lo, nn = sl.n_hi - sl.n_lo;
  if (w.n == n && w.m == mvars && w.parts == parts && w.nx == nx && w.nt == nt && w.nn == nn) return PS_OK;
  w.release();
  const size_t chunk = (mvars + m->ndev - 1) / m->ndev;
  const size_t rows = parts > 1 ? 2 * np / parts : n;
  PS_TRY(dev_alloc((void**)&w.wfull, chunk * m->ndev * sizeof(Fr)));
  PS_TRY(dev_alloc((void**)&w.e_all, (parts > 1 ? parts : 1) * rows * sizeof(Fr)));
  PS_TRY(dev_alloc((void**)&w.coef[0], n * sizeof(Fr)));
  PS_TRY(dev_alloc((void**)&w.coef[1], n * sizeof(Fr)));
  PS_TRY(dev_alloc((void**)&w.scA, (nx + 2) * sizeof(Fr)));
  PS_TRY(dev_alloc((void**)&w.scB, (nx + 2) * sizeof(Fr)));
  PS_TRY(dev_alloc((void**)&w.scC, (nn + nt + nx + 3) * sizeof(Fr)));
  if (d == 0) {
    PS_TRY(dev_alloc((void**)&w.h_full, n * sizeof(Fr)));
    PS_TRY(dev_alloc((void**)&w.recs, (size_t)m->ndev * REC_BYTES));
  }
  PS_TRY(dev_alloc((void**)&w.rec, REC_BYTES));
  PS_TRY(dev_alloc((void**)&w.status, 16));
  (void)nio;
  w.n = n; w.m = mvars; w.parts = parts; w.nx = nx; w.nt = nt; w.nn = nn;
  return PS_OK;
}

}  // namespace

extern "C" {

int ps_mctx_create(const int* devices, int ndev, ps_mctx** out) {
  if (!out || !devices || ndev < 1 || ndev > 64) return PS_ERR_ARG;
#if PS_GPU
  for (int i = 0; i < ndev; i++)
    for (int j = 0; j < i; j++)
      if (devices[i] == devices[j]) return PS_ERR_ARG;
#endif
  ps_mctx* m = new (std::nothrow) ps_mctx();
  if (!m) return PS_ERR_ALLOC;
  m->ndev = ndev;
  m->devs.assign(devices, devices + ndev);
  m->ctx.assign(ndev, nullptr);
  m->ws.resize(ndev);
  m->bar.count = ndev;
  int rc = PS_OK;
  for (int d = 0; d < ndev && rc == PS_OK; d++) rc = ps_ctx_create(devices[d], &m->ctx[d]);
#if PS_GPU
  for (int d = 0; d < ndev && rc == PS_OK; d++) {
    if (cudaSetDevice(devices[d]) != cudaSuccess) { rc = PS_ERR_CUDA; break; }
    for (int p = 0; p < ndev; p++) {
      if (p == d) continue;
      int can = 0;
      cudaDeviceCanAccessPeer(&can, devices[d], devices[p]);
      if (can) {
        cudaError_t e = cudaDeviceEnablePeerAccess(devices[p], 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { rc = PS_ERR_CUDA; break; }
        cudaGetLastError();   // clear "already enabled"
      }   // without peer access cudaMemcpyPeerAsync still works (staged through the host)
    }
    for (int k = 0; k < 5 && rc == PS_OK; k++) {
      cudaEvent_t e;
      if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) { rc = PS_ERR_CUDA; break; }
      m->ev[k].push_back(e);
    }
    for (int k = 0; k < TL_MAX && rc == PS_OK; k++) {
      cudaEvent_t e;
      if (cudaEventCreate(&e) != cudaSuccess) { rc = PS_ERR_CUDA; break; }
      m->tl[k].push_back(e);
    }
  }
#else
  for (int k = 0; k < 5; k++) m->ev[k].assign(ndev, nullptr);
  for (int k = 0; k < TL_MAX; k++) m->tl[k].assign(ndev, nullptr);
#endif
  if (rc != PS_OK) { ps_mctx_destroy(m); return rc; }
  m->workers.reserve((size_t)ndev);
  for (int d = 0; d < ndev; d++) {
    Worker* w = new (std::nothrow) Worker();
    if (!w) { ps_mctx_destroy(m); return PS_ERR_ALLOC; }
    m->workers.push_back(w);
    w->th = std::thread(worker_main, m, d, w);
  }
  *out = m;
  return PS_OK;
}

void ps_mctx_destroy(ps_mctx* m) {
  if (!m) return;
  for (Worker* w : m->workers) {
    { std::lock_guard<std::mutex> lk(w->mu); w->quit = true; }
    w->cv.notify_all();
    if (w->th.joinable()) w->th.join();
    delete w;
  }
  for (int d = 0; d < m->ndev; d++) {
#if PS_GPU
    cudaSetDevice(m->devs[d]);
    if (m->ctx[d]) cudaStreamSynchronize(m->ctx[d]->stream);
    for (int k = 0; k < 5; k++) if ((int)m->ev[k].size() > d && m->ev[k][d]) cudaEventDestroy((cudaEvent_t)m->ev[k][d]);
    for (int k = 0; k < TL_MAX; k++) if ((int)m->tl[k].size() > d && m->tl[k][d]) cudaEventDestroy((cudaEvent_t)m->tl[k][d]);
#endif
    if ((int)m->ws.size() > d) m->ws[d].release();
    if (m->ctx[d]) ps_ctx_destroy(m->ctx[d]);
  }
  delete m;
}

int ps_mctx_size(const ps_mctx* m) { return m ? m->ndev : 0; }
ps_ctx* ps_mctx_ctx(ps_mctx* m, int i) { return (m && i >= 0 && i < m->ndev) ? m->ctx[i] : nullptr; }

int ps_mctx_set_option(ps_mctx* m, const char* name, int value) {
  if (!m || !name) return PS_ERR_ARG;
  if (!strcmp(name, "rank0_share_percent")) {
    if (value < 0 || value > 100) return PS_ERR_ARG;
    m->rank0_share = (float)value / 100.f;
    return PS_OK;
  }
  for (int d = 0; d < m->ndev; d++) PS_TRY(ps_ctx_set_option(m->ctx[d], name, value));
  return PS_OK;
}

// ---- sharded base sets and MSM ---------------------------------------------------------------------------
int ps_mbases_from_scalars(ps_mctx* m, int group, const uint8_t* scalars_be, size_t n, int window_bits, int precompute_tables,
                           ps_mbases** out) {
  if (!m || !out || (n && !scalars_be) || (group != PS_G1 && group != PS_G2)) return PS_ERR_ARG;
  ps_mbases* b = new (std::nothrow) ps_mbases();
  if (!b) return PS_ERR_ALLOC;
  b->group = group; b->n = n;
  b->part.assign(m->ndev, nullptr);
  b->lo = weighted_cuts(n, std::vector<double>((size_t)m->ndev, 1.0));
  int rc = run_all(m, [&](int d) -> int {
    return ps_bases_from_scalars(m->ctx[d], group, scalars_be + 32 * b->lo[d], b->lo[d + 1] - b->lo[d], window_bits, precompute_tables,
                                 &b->part[d]);
  });
  if (rc != PS_OK) { ps_mbases_free(b); return rc; }
  *out = b;
  return PS_OK;
}

int ps_mbases_load(ps_mctx* m, int group, const uint8_t* points, size_t n, int format, int window_bits, int precompute_tables,
                   ps_mbases** out) {
  if (!m || !out || (n && !points) || (group != PS_G1 && group != PS_G2)) return PS_ERR_ARG;
  if (format != PS_FMT_COMPRESSED && format != PS_FMT_AFFINE) return PS_ERR_ARG;
  ps_mbases* b = new (std::nothrow) ps_mbases();
  if (!b) return PS_ERR_ALLOC;
  b->group = group; b->n = n;
  b->part.assign(m->ndev, nullptr);
  b->lo = weighted_cuts(n, std::vector<double>((size_t)m->ndev, 1.0));
  const size_t per = group == PS_G1 ? (format == PS_FMT_COMPRESSED ? 48 : 96) : (format == PS_FMT_COMPRESSED ? 96 : 192);
  int rc = run_all(m, [&](int d) -> int {
    return ps_bases_load(m->ctx[d], group, points + per * b->lo[d], b->lo[d + 1] - b->lo[d], format, window_bits, precompute_tables,
                         &b->part[d]);
  });
  if (rc != PS_OK) { ps_mbases_free(b); return rc; }
  *out = b;
  return PS_OK;
}

void ps_mbases_free(ps_mbases* b) {
  if (!b) return;
  for (ps_bases* p : b->part) ps_bases_free(p);
  delete b;
}

size_t ps_mbases_len(const ps_mbases* b) { return b ? b->n : 0; }

// Poly.BlindEval (algebra.go:348-359) over a sharded base set: every device sums its point range (own buckets,
// own reduction); the partial points (192 / 384 B) are pushed to device 0, which adds them and encodes.
int ps_mmsm(ps_mctx* m, const ps_mbases* b, const uint8_t* scalars_be, size_t n, uint8_t* out) {
  if (!m || !b || !out || (n && !scalars_be)) return PS_ERR_ARG;
  if (n != b->n) return PS_ERR_LENGTH;
  if ((int)b->part.size() != m->ndev) return PS_ERR_ARG;
  const size_t pb = b->group == PS_G1 ? 192 : 384;
  // record buffers: reuse the prover's (allocate on first use)
  int rc = run_all(m, [&](int d) -> int {
    ProveWs& w = m->ws[d];
    if (!w.rec) PS_TRY(dev_alloc((void**)&w.rec, REC_BYTES));
    if (d == 0 && !w.recs) PS_TRY(dev_alloc((void**)&w.recs, (size_t)m->ndev * REC_BYTES));
    return PS_OK;
  });
  if (rc != PS_OK) return rc;
  return run_all(m, [&](int d) -> int {
    int my_rc = PS_OK;
    ps_ctx* ctx = m->ctx[d];
    const size_t lo = b->lo[d], cnt = b->lo[d + 1] - lo;
    uint32_t* d_err = nullptr;
    STAGE(msm_partial_host_scalars(ctx, b->part[d], scalars_be + 32 * lo, cnt, m->ws[d].rec, &d_err));
    // the encoding flag lives in this call's arena: fetch it NOW, in stream order -- the combine below starts a new call
    // on device 0, and a new call may merge (free) the arena blocks the flag sits in
    uint32_t h = 0;
    if (d_err) STAGE(dev_d2h(&h, d_err, 4, ctx->stream));
    STAGE(peer_copy(m, m->ws[0].recs + (size_t)d * pb, 0, m->ws[d].rec, d, pb));
    STAGE(ev_record(m, 4, d));
    m->bar.arrive();
    if (d == 0) {
      for (int p = 1; p < m->ndev; p++) STAGE(ev_wait(m, 4, p, 0));
      STAGE(ps_msm_combine(ctx, b->group, m->ws[0].recs, (size_t)m->ndev, out));
    }
    dev_sync(ctx->stream);
    if (h && my_rc == PS_OK) { my_rc = PS_ERR_ENCODING; m->failed.store(1); }
    return my_rc;
  });
}

// ---- sharded Groth16 key / replicated QAP -------------------------------------------------------------------
int ps_mg16_key_load(ps_mctx* m, size_t n_gates, size_t n_nio, int format, const uint8_t* xi, const uint8_t* xi2, const uint8_t* xit,
                     const uint8_t* niolp, const uint8_t* alpha, const uint8_t* beta, const uint8_t* delta, const uint8_t* beta2,
                     const uint8_t* delta2, ps_mg16_key** key) {
  if (!m || !key || !xi || !xi2 || !xit || (n_nio && !niolp) || !alpha || !beta || !delta || !beta2 || !delta2 || n_gates < 2)
    return PS_ERR_ARG;
  ps_mg16_key* k = new (std::nothrow) ps_mg16_key();
  if (!k) return PS_ERR_ALLOC;
  k->n = n_gates; k->n_nio = n_nio;
  k->part.assign(m->ndev, nullptr);
  k->slice.resize(m->ndev);
  const std::vector<double> w = shard_weights(m);
  const std::vector<size_t> cx = weighted_cuts(n_gates, w), ct = weighted_cuts(n_gates - 1, w), cn = weighted_cuts(n_nio, w);
  for (int d = 0; d < m->ndev; d++) k->slice[d] = KeySlice{cx[d], cx[d + 1], ct[d], ct[d + 1], cn[d], cn[d + 1], d == 0};
  // windows sized for what one device runs: the early G1 batch (A_d and the h-free pieces of C_d: two outputs) and B_d
  const size_t nd = (size_t)m->ndev;
  const double bc = (double)m->ctx[0]->msm_bucket_cost;
  auto window = [&](size_t points, int nsets) {
    int best = 4; double best_cost = 1e300;
    for (int c = 4; c <= 24; c++) {
      double W = msm_windows(c), cost = (double)(points + 1) * W * 10.0 + (double)(1u << (c - 1)) * bc * nsets;
      if ((double)points * W >= 2.0e9) continue;
      if (cost < best_cost) { best_cost = cost; best = c; }
    }
    return best;
  };
  const int c_g1 = window((2 * n_gates + n_nio) / nd, 2), c_g2 = window(n_gates / nd, 1);
  int rc = run_all(m, [&](int d) -> int {
    return g16_key_load_slice(m->ctx[d], k->slice[d], format, c_g1, c_g2, xi, xi2, xit, niolp, alpha, beta, delta, beta2, delta2, &k->part[d]);
  });
  if (rc != PS_OK) { ps_mg16_key_free(k); return rc; }
  *key = k;
  return PS_OK;
}

void ps_mg16_key_free(ps_mg16_key* k) {
  if (!k) return;
  for (ps_g16_key* p : k->part) ps_g16_key_free(p);
  delete k;
}

int ps_mqap_load_r1cs(ps_mctx* m, size_t n_gates, size_t n_vars, size_t n_io, const uint32_t* l_row_ptr, const uint32_t* l_col,
                      const uint8_t* l_val, const uint32_t* r_row_ptr, const uint32_t* r_col, const uint8_t* r_val,
                      const uint32_t* o_row_ptr, const uint32_t* o_col, const uint8_t* o_val, ps_mqap** qap) {
  if (!m || !qap) return PS_ERR_ARG;
  ps_mqap* q = new (std::nothrow) ps_mqap();
  if (!q) return PS_ERR_ALLOC;
  q->n = n_gates; q->m = n_vars; q->n_io = n_io; q->dense = false;
  q->part.assign(m->ndev, nullptr);
  int rc = run_all(m, [&](int d) -> int {
    return ps_qap_load_r1cs(m->ctx[d], n_gates, n_vars, n_io, l_row_ptr, l_col, l_val, r_row_ptr, r_col, r_val, o_row_ptr, o_col, o_val,
                            &q->part[d]);
  });
  if (rc != PS_OK) { ps_mqap_free(q); return rc; }
  *qap = q;
  return PS_OK;
}

int ps_mqap_load_dense(ps_mctx* m, size_t n_gates, size_t n_vars, size_t n_io, const uint8_t* left, const uint8_t* right,
                       const uint8_t* out, const uint8_t* z, ps_mqap** qap) {
  if (!m || !qap) return PS_ERR_ARG;
  ps_mqap* q = new (std::nothrow) ps_mqap();
  if (!q) return PS_ERR_ALLOC;
  q->n = n_gates; q->m = n_vars; q->n_io = n_io; q->dense = true;
  q->part.assign(m->ndev, nullptr);
  int rc = ps_qap_load_dense(m->ctx[0], n_gates, n_vars, n_io, left, right, out, z, &q->part[0]);   // the quotient runs on device 0
  if (rc != PS_OK) { ps_mqap_free(q); return rc; }
  *qap = q;
  return PS_OK;
}

void ps_mqap_free(ps_mqap* q) {
  if (!q) return;
  for (ps_qap* p : q->part) if (p) ps_qap_free(p);
  delete q;
}

// ---- Groth16Prove (groth16.go:122-211) on all devices of the context, one call ------------------------------------
int ps_mg16_prove(ps_mctx* m, const ps_mg16_key* key, const ps_mqap* qap, const uint8_t* witness_be, const uint8_t* r_be,
                  const uint8_t* s_be, uint8_t* outA, uint8_t* outB, uint8_t* outC) {
  if (!m || !key || !qap || !witness_be || !r_be || !s_be || !outA || !outB || !outC) return PS_ERR_ARG;
  if (key->n != qap->n || key->n_nio != qap->n_io) return PS_ERR_LENGTH;
  if ((int)key->part.size() != m->ndev || (int)qap->part.size() != m->ndev) return PS_ERR_ARG;
  const int N = m->ndev;
  const size_t n = qap->n, mv = qap->m, nio = qap->n_io, diff = mv - nio;
  const size_t parts = (size_t)N / 2;
  size_t np = 2;                       // leaves of the interpolation tree (sparse QAP): the power of two >= n
  while (np < n) np <<= 1;
  const bool pipelined = !qap->dense && N >= 2 && N % 2 == 0 && (parts & (parts - 1)) == 0 && parts <= np / 2;
  const size_t rows = parts > 1 ? 2 * np / parts : n;
  const size_t chunk = (mv + N - 1) / N;
  m->tl_count = 0;

  int rc = run_all(m, [&](int d) -> int { return ws_prepare(m, d, n, np, mv, nio, key->slice[d], pipelined ? parts : 1); });
  if (rc != PS_OK) return rc;

  rc = run_all(m, [&](int d) -> int {
    int my_rc = PS_OK;
    ps_ctx* ctx = m->ctx[d];
    ProveWs& w = m->ws[d];
    const KeySlice& sl = key->slice[d];
    const size_t nx = sl.x_hi - sl.x_lo, nt = sl.t_hi - sl.t_lo, nn = sl.n_hi - sl.n_lo;
    int mark = 0;
    STAGE(dev_memset(w.status, 0, 16, ctx->stream));
    STAGE(tl_mark(m, mark++, d));
    if (pipelined) {
      const int g = d / (int)parts;
      const size_t part = (size_t)d % parts;
      // W: this device's slice of the witness, pushed to every peer
      const size_t lo = std::min(mv, (size_t)d * chunk), hi = std::min(mv, (size_t)(d + 1) * chunk);
      STAGE(ps_fr_upload(ctx, witness_be + 32 * lo, hi - lo, w.wfull + lo, w.status));
      for (int p = 0; p < N; p++)
        if (p != d) STAGE(peer_copy(m, m->ws[p].wfull + lo, p, w.wfull + lo, d, (hi - lo) * sizeof(Fr)));
      STAGE(ev_record(m, 0, d));
      m->bar.arrive();
      for (int p = 0; p < N; p++)
        if (p != d) STAGE(ev_wait(m, 0, p, d));
      STAGE(tl_mark(m, mark++, d));
      // I: subtree of polynomial g over this device's gates (parts == 1: the whole polynomial, coefficients)
      Fr* my_root = parts > 1 ? w.e_all + part * rows : w.coef[g];
      STAGE(ps_qap_interp_part_dev(ctx, qap->part[d], w.wfull, g, part, parts, my_root, nullptr, w.status));
      STAGE(tl_mark(m, mark++, d));
      if (parts > 1) {
        // R: roots to the other devices of the same polynomial; everyone of them folds the top levels
        for (size_t q = 0; q < parts; q++) {
          const int p = g * (int)parts + (int)q;
          if (p != d) STAGE(peer_copy(m, m->ws[p].e_all + part * rows, p, my_root, d, rows * sizeof(Fr)));
        }
        STAGE(ev_record(m, 1, d));
        m->bar.arrive();
        for (size_t q = 0; q < parts; q++) {
          const int p = g * (int)parts + (int)q;
          if (p != d) STAGE(ev_wait(m, 1, p, d));
        }
        STAGE(tl_mark(m, mark++, d));
        STAGE(ps_qap_interp_finish(ctx, qap->part[d], parts, w.e_all, w.coef[g]));
      } else {
        STAGE(tl_mark(m, mark++, d));
      }
      STAGE(tl_mark(m, mark++, d));
      // X: swap with the device that holds the other polynomial
      const int partner = (d + (int)parts) % N;
      STAGE(peer_copy(m, m->ws[partner].coef[g], partner, w.coef[g], d, n * sizeof(Fr)));
      STAGE(ev_record(m, 2, d));
      m->bar.arrive();
      STAGE(ev_wait(m, 2, partner, d));
      STAGE(tl_mark(m, mark++, d));
      // S + early MSMs; device 0 divides first so that h leaves as early as possible
      if (d == 0) {
        STAGE(ps_g16_h_from_ab(ctx, qap->part[0], w.coef[0], w.coef[1], w.h_full));
        for (int p = 0; p < N; p++) {
          const KeySlice& ps_ = key->slice[p];
          STAGE(peer_copy(m, m->ws[p].scC + (ps_.n_hi - ps_.n_lo), p, w.h_full + ps_.t_lo, 0, (ps_.t_hi - ps_.t_lo) * sizeof(Fr)));
        }
        STAGE(ev_record(m, 3, 0));
      }
      STAGE(g16_slice_scalars(ctx, sl, r_be, s_be, w.coef[0], w.coef[1], w.wfull, diff, w.scA, w.scB, w.scC));
      STAGE(tl_mark(m, mark++, d));
      // MSMs: B_d starts at once on the second stream; the G1 pipeline (A_d and all of C_d) is enqueued behind the
      // arrival of h.  The host barrier (device 0 has recorded its event) sits inside the call, after B_d's launches.
      bool arrived = false;
      STAGE(g16_slice_msm_all(ctx, key->part[d], sl, w.scA, w.scB, w.scC, w.rec, [&]() -> int {
        m->bar.arrive();
        arrived = true;
        return d != 0 ? ev_wait(m, 3, 0, d) : PS_OK;
      }));
      if (!arrived) m->bar.arrive();   // a failed device still meets the others
      STAGE(tl_mark(m, mark++, d));
    } else {
      // device 0 computes the three scalar vectors (quotient included) and pushes every device its slices
      if (d == 0) {
        const size_t nA = n + 2, nC = nio + (n - 1) + n + 3;
        Fr *fa = nullptr, *fc = nullptr, *fb = nullptr;
        STAGE(dev_alloc((void**)&fa, nA * sizeof(Fr)));
        STAGE(dev_alloc((void**)&fc, nC * sizeof(Fr)));
        STAGE(dev_alloc((void**)&fb, nA * sizeof(Fr)));
        // ps_g16_scalars checks key sizes against the QAP: hand it a key header with the global sizes
        ps_g16_key hdr;
        hdr.n = n; hdr.n_nio = nio;
        STAGE(ps_g16_scalars(ctx, &hdr, qap->part[0], witness_be, r_be, s_be, fa, fc, fb));
        for (int p = 0; p < N; p++) {
          const KeySlice& ps_ = key->slice[p];
          const size_t px = ps_.x_hi - ps_.x_lo, pt = ps_.t_hi - ps_.t_lo, pn = ps_.n_hi - ps_.n_lo;
          ProveWs& pw = m->ws[p];
          STAGE(peer_copy(m, pw.scA, p, fa + ps_.x_lo, 0, px * sizeof(Fr)));
          STAGE(peer_copy(m, pw.scB, p, fb + ps_.x_lo, 0, px * sizeof(Fr)));
          STAGE(peer_copy(m, pw.scC, p, fc + ps_.n_lo, 0, pn * sizeof(Fr)));
          STAGE(peer_copy(m, pw.scC + pn, p, fc + nio + ps_.t_lo, 0, pt * sizeof(Fr)));
          STAGE(peer_copy(m, pw.scC + pn + pt, p, fc + nio + (n - 1) + ps_.x_lo, 0, px * sizeof(Fr)));
          if (ps_.consts) {
            STAGE(peer_copy(m, pw.scA + px, p, fa + n, 0, 2 * sizeof(Fr)));
            STAGE(peer_copy(m, pw.scB + px, p, fb + n, 0, 2 * sizeof(Fr)));
            STAGE(peer_copy(m, pw.scC + pn + pt + px, p, fc + nio + (n - 1) + n, 0, 3 * sizeof(Fr)));
          }
        }
        STAGE(ev_record(m, 3, 0));
        dev_sync(ctx->stream);   // fa / fc / fb are freed below
        dev_free(fa); dev_free(fc); dev_free(fb);
      }
      m->bar.arrive();
      if (d != 0) STAGE(ev_wait(m, 3, 0, d));
      STAGE(tl_mark(m, mark++, d));
      STAGE(g16_slice_msm_all(ctx, key->part[d], sl, w.scA, w.scB, w.scC, w.rec, nullptr));
      STAGE(tl_mark(m, mark++, d));
    }
    // the record goes to device 0
    STAGE(dev_d2d(w.rec + 960, w.status, 16, ctx->stream));
    STAGE(tl_mark(m, mark++, d));
    STAGE(peer_copy(m, m->ws[0].recs + (size_t)d * REC_BYTES, 0, w.rec, d, REC_BYTES));
    STAGE(ev_record(m, 4, d));
    m->bar.arrive();
    if (d == 0) {
      for (int p = 1; p < N; p++) STAGE(ev_wait(m, 4, p, 0));
      STAGE(ps_g16_combine(ctx, w.recs, (size_t)N, REC_BYTES, outA, outB, outC));
      STAGE(tl_mark(m, mark++, d));
      // status words of all devices travelled with the records
      if (!m->failed.load()) {
        std::vector<uint8_t> host((size_t)N * REC_BYTES);
        STAGE(dev_d2h(host.data(), w.recs, host.size(), ctx->stream));
        STAGE(dev_sync(ctx->stream));
        uint32_t st = 0;
        for (int p = 0; p < N; p++) { uint32_t v; memcpy(&v, host.data() + (size_t)p * REC_BYTES + 960, 4); st |= v; }
        if (my_rc == PS_OK && (st & 2)) my_rc = PS_ERR_REMAINDER;
        else if (my_rc == PS_OK && (st & 1)) my_rc = PS_ERR_ENCODING;
      }
      m->tl_count = mark;
    } else {
      dev_sync(ctx->stream);
    }
    (void)nx; (void)nt; (void)nn;
    return my_rc;
  });
  return rc;
}

// stage marks of the last ps_mg16_prove on device `dev`: out_ms[k] = time from the start mark to mark k + 1
// (CUDA events on that device's stream); returns the number of values written in *count
int ps_mg16_last_timeline(ps_mctx* m, int dev, float* out_ms, int max, int* count) {
  if (!m || !out_ms || !count || dev < 0 || dev >= m->ndev) return PS_ERR_ARG;
  *count = 0;
#if PS_GPU
  PS_CUDA_TRY(cudaSetDevice(m->devs[dev]));
  // device 0 records one more mark (after the combine) than the others
  const int marks = dev == 0 ? m->tl_count : m->tl_count - 1;
  for (int k = 1; k < marks && k <= max; k++) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, (cudaEvent_t)m->tl[0][dev], (cudaEvent_t)m->tl[k][dev]) != cudaSuccess) { cudaGetLastError(); break; }
    out_ms[k - 1] = ms;
    *count = k;
  }
#else
  (void)max;
#endif
  return PS_OK;
}

}  // extern "C"
