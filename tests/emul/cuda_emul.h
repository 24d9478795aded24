/*
 * cuda_emul.h -- TEST INFRASTRUCTURE ONLY.
 *
 * A tiny CUDA-on-CPU shim so that simudo_b200/csrc/pdd_kernels.cu can be
 * compiled with g++ and its kernel logic exercised in this GPU-less container
 * (pytest -m "not gpu").  One OS thread per CUDA thread of a block, blocks run
 * one after another; __syncthreads / __syncwarp are std::barrier's.  The
 * product never loads the resulting library: simudo_b200/_lib.py only opens
 * the nvcc-built libsimudo_b200.so and insists on a CUDA device.
 */
#ifndef SB_CUDA_EMUL_H
#define SB_CUDA_EMUL_H

#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __constant__ static
#define __shared__ static
#define __align__(n) alignas(n)
#define __launch_bounds__(...)

struct dim3e { unsigned x, y, z; };
struct alignas(16) uint4 { unsigned x, y, z, w; };
static thread_local dim3e threadIdx, blockIdx, blockDim, gridDim;

typedef int cudaError_t;
typedef void* cudaStream_t;
typedef void* cudaEvent_t;
enum { cudaSuccess = 0 };
enum { cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2 };
enum { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
static inline const char* cudaGetErrorString(cudaError_t) { return "emulated"; }
static inline cudaError_t cudaMalloc(void** p, size_t n) { *p = malloc(n); return *p ? 0 : 2; }
static inline cudaError_t cudaFree(void* p) { free(p); return 0; }
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, int) { memcpy(d, s, n); return 0; }
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) { memset(d, v, n); return 0; }
static inline cudaError_t cudaGetLastError() { return 0; }
template <typename F>
static inline cudaError_t cudaFuncSetAttribute(F, int, int) { return 0; }
#define cudaMemcpyToSymbolAsync(sym, src, n, off, kind, s) (memcpy(&(sym), (src), (n)), 0)

namespace emul {
struct BlockCtx {
  std::unique_ptr<std::barrier<>> block_bar;
  std::vector<std::unique_ptr<std::barrier<>>> warp_bar;
  std::vector<unsigned char> smem;
  std::vector<double> shfl;   /* one slot per thread */
};
static BlockCtx* g_ctx = nullptr;
static thread_local unsigned t_tid = 0;

template <typename F>
static void launch(unsigned grid, unsigned block, size_t smem, F body) {
  BlockCtx ctx;
  ctx.smem.resize(smem + 64);
  ctx.shfl.resize(block);
  g_ctx = &ctx;
  for (unsigned b = 0; b < grid; ++b) {
    ctx.block_bar.reset(new std::barrier<>(block));
    ctx.warp_bar.clear();
    for (unsigned w = 0; w * 32 < block; ++w) {
      unsigned n = block - w * 32 < 32 ? block - w * 32 : 32;
      ctx.warp_bar.emplace_back(new std::barrier<>(n));
    }
    std::vector<std::thread> th;
    for (unsigned t = 0; t < block; ++t)
      th.emplace_back([&, t, b] {
        threadIdx = {t, 0, 0}; blockIdx = {b, 0, 0};
        blockDim = {block, 1, 1}; gridDim = {grid, 1, 1};
        t_tid = t;
        body();
      });
    for (auto& x : th) x.join();
  }
  g_ctx = nullptr;
}
static inline unsigned char* dyn_smem() {
  uintptr_t p = (uintptr_t)g_ctx->smem.data();
  return (unsigned char*)((p + 63) & ~(uintptr_t)63);
}
}  // namespace emul

static inline void __syncthreads() { emul::g_ctx->block_bar->arrive_and_wait(); }
static inline void __syncwarp() { emul::g_ctx->warp_bar[emul::t_tid / 32]->arrive_and_wait(); }

static inline double atomicAdd(double* a, double v) {
  uint64_t* p = reinterpret_cast<uint64_t*>(a);
  uint64_t old = __atomic_load_n(p, __ATOMIC_RELAXED), des;
  double od;
  do {
    memcpy(&od, &old, 8);
    const double nd = od + v;
    memcpy(&des, &nd, 8);
  } while (!__atomic_compare_exchange_n(p, &old, des, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED));
  return od;
}
static inline int atomicAdd(int* a, int v) { return __atomic_fetch_add(a, v, __ATOMIC_RELAXED); }
static inline unsigned long long atomicMax(unsigned long long* a, unsigned long long v) {
  unsigned long long old = __atomic_load_n(a, __ATOMIC_RELAXED);
  while (old < v && !__atomic_compare_exchange_n(a, &old, v, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
  return old;
}
static inline long long __double_as_longlong(double d) { long long r; memcpy(&r, &d, 8); return r; }
static inline double __shfl_down_sync(unsigned, double v, int o) {
  auto& c = *emul::g_ctx;
  unsigned t = emul::t_tid;
  c.shfl[t] = v;
  __syncwarp();
  unsigned src = t + o;
  double r = ((src / 32) == (t / 32) && src < c.shfl.size()) ? c.shfl[src] : v;
  __syncwarp();
  return r;
}
static inline int __any_sync(unsigned, int pred) {
  auto& c = *emul::g_ctx;
  const unsigned t = emul::t_tid;
  c.shfl[t] = pred ? 1.0 : 0.0;
  __syncwarp();
  int r = 0;
  for (unsigned q = (t / 32) * 32; q < (t / 32) * 32 + 32 && q < c.shfl.size(); ++q) r |= c.shfl[q] != 0.0;
  __syncwarp();
  return r;
}
using std::isnan;

#define SB_EMUL 1
#endif
