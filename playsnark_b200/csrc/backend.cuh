// Launch / memory backend.
//
// Every kernel in this library is a functor with a `run(tid, args...)` body.  In the product build
// (nvcc, sm_100a) PS_LAUNCH starts it as a __global__ grid on the context's stream.  With
// -DPS_HOST_EMU (tests only, built by tests/conftest.py into tests/_build/) the same body is driven
// by a serial loop on the CPU so that the host orchestration, index arithmetic and the arithmetic
// templates can be checked against the oracle on a machine without a GPU.  The emulation is never
// linked into libplaysnark_b200.so: there is no CPU fallback in the product path.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <vector>

#include "../../include/playsnark_b200.h"

#if defined(__CUDACC__) && !defined(PS_HOST_EMU)
#include <cuda_runtime.h>
#define PS_GPU 1
#else
#define PS_GPU 0
#endif

#include "field.cuh"

namespace ps {

#if PS_GPU
#define PS_CUDA_TRY(expr)                                                                     \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess) {                                                                  \
      fprintf(stderr, "playsnark_b200: CUDA error %s at %s:%d\n", cudaGetErrorString(_e), __FILE__, __LINE__); \
      return PS_ERR_CUDA;                                                                     \
    }                                                                                         \
  } while (0)
typedef cudaStream_t ps_stream_t;
#else
#define PS_CUDA_TRY(expr) do { (void)(expr); } while (0)
typedef void* ps_stream_t;
#endif

#define PS_TRY(expr)             \
  do {                           \
    int _rc = (expr);            \
    if (_rc != PS_OK) return _rc; \
  } while (0)

PS_DEV uint32_t ps_atomic_add(uint32_t* p, uint32_t v) {
#ifdef __CUDA_ARCH__
  return atomicAdd(p, v);
#else
  uint32_t o = *p; *p = o + v; return o;
#endif
}
PS_DEV void ps_atomic_or(uint32_t* p, uint32_t v) {
#ifdef __CUDA_ARCH__
  atomicOr(p, v);
#else
  *p |= v;
#endif
}

PS_DEV void ps_atomic_max(uint32_t* p, uint32_t v) {
#ifdef __CUDA_ARCH__
  atomicMax(p, v);
#else
  if (v > *p) *p = v;
#endif
}

inline std::atomic<uint64_t>& launch_counter() { static std::atomic<uint64_t> c{0}; return c; }

// optional K::MIN_BLOCKS (resident blocks per SM the register allocator must allow)
template <class K, class = void> struct MinBlocks { static constexpr int V = 1; };
template <class K> struct MinBlocks<K, decltype((void)K::MIN_BLOCKS)> { static constexpr int V = K::MIN_BLOCKS; };
// optional K::SMEM_PAD: bytes of (unused) dynamic shared memory requested per block, to hold the number of resident
// blocks per SM BELOW what the registers alone would allow (wave-count tuning of multiplier-bound kernels); <= 48 KB
template <class K, class = void> struct SmemPad { static constexpr int V = 0; };
template <class K> struct SmemPad<K, decltype((void)K::SMEM_PAD)> { static constexpr int V = K::SMEM_PAD; };

#if PS_GPU
template <class K, class... Args>
__global__ void __launch_bounds__(K::BLOCK, MinBlocks<K>::V) ps_kernel(uint32_t n, Args... args) {
  uint32_t tid = blockIdx.x * (uint32_t)K::BLOCK + threadIdx.x;
  if (tid < n) K::run(tid, args...);
}
template <class K, class... Args>
inline int ps_launch(ps_stream_t st, size_t n, Args... args) {
  if (n == 0) return PS_OK;
  if (n > 0xFFFFFFFFull) return PS_ERR_ARG;
  uint32_t blocks = (uint32_t)((n + K::BLOCK - 1) / K::BLOCK);
  ps_kernel<K, Args...><<<blocks, K::BLOCK, SmemPad<K>::V, st>>>((uint32_t)n, args...);
  PS_CUDA_TRY(cudaGetLastError());
  launch_counter()++;
  return PS_OK;
}
#else
// Host emulation (tests only).  Kernels whose threads write disjoint outputs and use no atomics may be marked
// EmuParallel (specialisations next to the kernels, inside #if !PS_GPU): their serial loop is cut over the host's
// cores, which is what keeps the CPU suite at a few minutes (the fixed-base tables alone are 16 320 scalar
// multiplications per context).  PS_EMU_PROFILE=1 prints the seconds spent per kernel at exit.
}  // namespace ps
#include <chrono>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <typeinfo>
namespace ps {
template <class K> struct EmuParallel { static constexpr bool V = false; };
struct EmuProfile {
  std::map<std::string, double> secs;
  std::mutex mu;
  bool on = getenv("PS_EMU_PROFILE") != nullptr;
  ~EmuProfile() {
    if (on) for (auto& kv : secs) if (kv.second > 0.2) fprintf(stderr, "[emu] %8.2f s  %s\n", kv.second, kv.first.c_str());
  }
};
inline EmuProfile& emu_profile() { static EmuProfile p; return p; }
template <class K, class... Args>
inline int ps_launch(ps_stream_t, size_t n, Args... args) {
  auto t0 = std::chrono::steady_clock::now();
  unsigned hw = std::thread::hardware_concurrency();
  if (EmuParallel<K>::V && n >= 64 && hw > 1) {
    const size_t parts = hw < 16 ? hw : 16;
    std::vector<std::thread> th;
    for (size_t p = 0; p < parts; p++)
      th.emplace_back([=]() { for (size_t t = n * p / parts; t < n * (p + 1) / parts; t++) K::run((uint32_t)t, args...); });
    for (auto& x : th) x.join();
  } else {
    for (size_t t = 0; t < n; t++) K::run((uint32_t)t, args...);
  }
  EmuProfile& pr = emu_profile();
  if (pr.on) {
    std::lock_guard<std::mutex> g(pr.mu);
    pr.secs[typeid(K).name()] += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  }
  return PS_OK;
}
#endif
#define PS_LAUNCH(K, st, n, ...) PS_TRY((ps_launch<K>(st, n, __VA_ARGS__)))
#define PS_COMMA ,

// ---- device memory ----------------------------------------------------------------------------
inline int dev_alloc(void** p, size_t bytes) {
  if (bytes == 0) bytes = 16;
#if PS_GPU
  PS_CUDA_TRY(cudaMalloc(p, bytes));
#else
  *p = malloc(bytes);
  if (!*p) return PS_ERR_ALLOC;
#endif
  return PS_OK;
}
inline void dev_free(void* p) {
  if (!p) return;
#if PS_GPU
  cudaFree(p);
#else
  free(p);
#endif
}
inline int dev_memset(void* p, int v, size_t bytes, ps_stream_t st) {
#if PS_GPU
  PS_CUDA_TRY(cudaMemsetAsync(p, v, bytes, st));
#else
  (void)st; memset(p, v, bytes);
#endif
  return PS_OK;
}
inline int dev_h2d(void* d, const void* h, size_t bytes, ps_stream_t st) {
#if PS_GPU
  PS_CUDA_TRY(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, st));
#else
  (void)st; memcpy(d, h, bytes);
#endif
  return PS_OK;
}
inline int dev_d2h(void* h, const void* d, size_t bytes, ps_stream_t st) {
#if PS_GPU
  PS_CUDA_TRY(cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, st));
#else
  (void)st; memcpy(h, d, bytes);
#endif
  return PS_OK;
}
inline int dev_d2d(void* d, const void* s, size_t bytes, ps_stream_t st) {
#if PS_GPU
  PS_CUDA_TRY(cudaMemcpyAsync(d, s, bytes, cudaMemcpyDeviceToDevice, st));
#else
  (void)st; memmove(d, s, bytes);
#endif
  return PS_OK;
}
inline int dev_sync(ps_stream_t st) {
#if PS_GPU
  PS_CUDA_TRY(cudaStreamSynchronize(st));
#else
  (void)st;
#endif
  return PS_OK;
}

// Scratch arena: bump allocation out of a few large device blocks that persist across calls, so a
// steady-state call performs no cudaMalloc.  reset() (start of a call) merges the blocks of the
// previous call into one of their combined size.
struct Arena {
  struct Block { char* base; size_t cap, off; };
  std::vector<Block> blocks;
  ps_stream_t stream = nullptr;
  static size_t pad(size_t b) { return (b + 255) & ~(size_t)255; }
  int reset() {
    if (blocks.size() > 1) {
      size_t total = 0;
      for (auto& b : blocks) total += b.cap;
      PS_TRY(dev_sync(stream));
      for (auto& b : blocks) dev_free(b.base);
      blocks.clear();
      Block nb{nullptr, total, 0};
      PS_TRY(dev_alloc((void**)&nb.base, total));
      blocks.push_back(nb);
    }
    for (auto& b : blocks) b.off = 0;
    return PS_OK;
  }
  void* take_bytes(size_t bytes) {
    size_t need = pad(bytes ? bytes : 1);
    for (auto& b : blocks)
      if (b.off + need <= b.cap) { void* p = b.base + b.off; b.off += need; return p; }
    Block nb{nullptr, need > (size_t(32) << 20) ? need : (size_t(32) << 20), 0};
    if (dev_alloc((void**)&nb.base, nb.cap) != PS_OK) return nullptr;
    nb.off = need;
    blocks.push_back(nb);
    return nb.base;
  }
  template <class T>
  T* take(size_t count) { return (T*)take_bytes(count * sizeof(T)); }
  void release() { for (auto& b : blocks) dev_free(b.base); blocks.clear(); }
};

}  // namespace ps

