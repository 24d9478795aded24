// Micro-benchmarks that size the assembly kernels' design (B200, sm_100a):
//  (1) DFMA peak: independent FMA chains per thread, warps per SM swept.
//  (2) store paths for many small contiguous runs (120 B .. 1 KB) scattered in a big array:
//      a) per-thread scalar 8-byte stores   b) per-thread 16-byte stores
//      c) warp-cooperative coalesced stores from shared memory (lane = entry)
//      d) per-thread cp.async.bulk shared->global (1-D TMA)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mb_fp64_store mb_fp64_store.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

template <int ILP>
__global__ void k_dfma(double* out, int iters, double a, double b) {
  double v[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) v[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) v[i] = fma(v[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += v[i];
  if (s == 1.2345) out[0] = s;
}

template <int ILP>
static void run_dfma(int threads, int ctas_per_sm) {
  double* d; CK(cudaMalloc(&d, 8));
  const int iters = 20000;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int grid = 148 * ctas_per_sm;
  k_dfma<ILP><<<grid, threads>>>(d, 100, 1.0000001, 1e-9);
  CK(cudaDeviceSynchronize());
  cudaEventRecord(e0);
  k_dfma<ILP><<<grid, threads>>>(d, iters, 1.0000001, 1e-9);
  cudaEventRecord(e1); CK(cudaDeviceSynchronize());
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double fl = 2.0 * ILP * (double)iters * threads * grid;
  printf("dfma ilp=%d threads/cta=%d ctas/sm=%d warps/sm=%d : %.2f TFLOP/s (%.3f ms)\n", ILP, threads,
         ctas_per_sm, threads / 32 * ctas_per_sm, fl / ms * 1e-9, ms);
  cudaFree(d);
}

// ---------------- store paths ----------------
// Each "cell" (thread) owns NR runs of RL doubles; run r of cell c starts at
// base(c, r) = (c * NR + r) * STRIDE + (c & 1)   (odd/even phases like the CSR rows).
template <int RL, int NR>
__global__ void k_store_scalar(double* out, int64_t ncell, int stride) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncell) return;
  for (int r = 0; r < NR; ++r) {
    double* dst = out + (c * NR + r) * stride + (c & 1);
#pragma unroll
    for (int e = 0; e < RL; ++e) dst[e] = (double)(c + e);
  }
}

template <int RL, int NR>
__global__ void k_store_coop(double* out, int64_t ncell, int stride) {
  extern __shared__ double sm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double* st = sm + warp * 32 * (RL + 1);
  const int64_t c0 = (int64_t)blockIdx.x * blockDim.x + warp * 32;
  const int64_t c = c0 + lane;
  for (int r = 0; r < NR; ++r) {
#pragma unroll
    for (int e = 0; e < RL; ++e) st[lane * (RL + 1) + e] = (double)(c + e);
    __syncwarp();
    // flush: 32 cells x RL entries, lane = entry within a pair of cells
    constexpr int PER = 32 / RL > 0 ? 32 / RL : 1;     // runs per instruction
    if (RL <= 32) {
      const int sub = lane / RL, e = lane - sub * RL;
      for (int i = 0; i < 32; i += PER) {
        const int ci = i + sub;
        if (sub < PER && ci < 32 && c0 + ci < ncell)
          out[((c0 + ci) * NR + r) * stride + ((c0 + ci) & 1) + e] = st[ci * (RL + 1) + e];
      }
    } else {
      for (int i = 0; i < 32; ++i) {
        if (c0 + i >= ncell) break;
        double* dst = out + ((c0 + i) * NR + r) * stride + ((c0 + i) & 1);
        for (int e = lane; e < RL; e += 32) dst[e] = st[i * (RL + 1) + e];
      }
    }
    __syncwarp();
  }
}

template <int RL, int NR>
__global__ void k_store_bulk(double* out, int64_t ncell, int stride) {
  extern __shared__ __align__(16) double sm[];
  constexpr int SL = (RL + 3) & ~1;              // per-lane slot, even, room for the phase shift
  double* st = sm + (size_t)threadIdx.x * SL;
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool act = c < ncell;
  for (int r = 0; r < NR; ++r) {
    const int phase = (int)(((c * NR + r) * stride + (c & 1)) & 1);
    if (r) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    if (act) {
#pragma unroll
      for (int e = 0; e < RL; ++e) st[phase + e] = (double)(c + e);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      double* dst = out + (c * NR + r) * stride + (c & 1);
      const double* s0 = st + phase;
      const int head = phase;
      const int nb = (RL - head) & ~1;
      if (head) dst[0] = s0[0];
      if (head + nb < RL) dst[RL - 1] = s0[RL - 1];
      const uint32_t saddr = (uint32_t)__cvta_generic_to_shared(s0 + head);
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                   :: "l"(dst + head), "r"(saddr), "r"(nb * 8) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

static int g_gap = -1;     // >= 0: doubles left unwritten between two runs (0 = dense output)
template <int RL, int NR>
static void run_store(int64_t ncell) {
  const int stride = g_gap >= 0 ? RL + g_gap
                                : RL + (RL > 60 ? 55 : 17) + ((RL & 1) ? 0 : 1);   // gaps between runs (other cells' columns)
  const size_t n = (size_t)ncell * NR * stride + 8;
  double* d; CK(cudaMalloc(&d, n * 8)); CK(cudaMemset(d, 0, n * 8));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const double bytes = (double)ncell * NR * RL * 8;
  const int T = 128;
  const unsigned grid = (unsigned)((ncell + T - 1) / T);
  float ms;
  for (int v = 0; v < 3; ++v) {
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      if (v == 0) k_store_scalar<RL, NR><<<grid, T>>>(d, ncell, stride);
      if (v == 1) { size_t sh = (size_t)(T / 32) * 32 * (RL + 1) * 8;
        cudaFuncSetAttribute(k_store_coop<RL, NR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh);
        k_store_coop<RL, NR><<<grid, T, sh>>>(d, ncell, stride); }
      if (v == 2) { size_t sh = (size_t)T * ((RL + 3) & ~1) * 8;
        cudaFuncSetAttribute(k_store_bulk<RL, NR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh);
        k_store_bulk<RL, NR><<<grid, T, sh>>>(d, ncell, stride); }
      cudaEventRecord(e1); CK(cudaDeviceSynchronize());
      cudaEventElapsedTime(&ms, e0, e1);
    }
    printf("store RL=%d NR=%d cells=%lld %s: %.3f ms  %.0f GB/s useful (%.0f GB/s incl. gaps)\n", RL, NR,
           (long long)ncell, v == 0 ? "scalar" : v == 1 ? "coop  " : "bulk  ", ms, bytes / ms * 1e-6,
           (double)n * 8 / ms * 1e-6);
  }
  cudaFree(d);
}

__global__ void k_fill(double* p, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = 1.0;
}

int main() {
  run_dfma<1>(128, 1); run_dfma<2>(128, 1); run_dfma<4>(128, 1); run_dfma<8>(128, 1);
  run_dfma<1>(128, 2); run_dfma<2>(128, 2); run_dfma<4>(128, 2); run_dfma<8>(128, 2);
  run_dfma<1>(128, 3); run_dfma<4>(128, 3); run_dfma<8>(128, 3);
  run_dfma<1>(256, 2); run_dfma<4>(256, 2); run_dfma<8>(256, 4); run_dfma<8>(1024, 2);
  { // plain fill reference
    const int64_t n = 1ll << 30; double* d; CK(cudaMalloc(&d, n * 8));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1); float ms;
    for (int r = 0; r < 2; ++r) { cudaEventRecord(e0); k_fill<<<148 * 8, 256>>>(d, n); cudaEventRecord(e1); CK(cudaDeviceSynchronize()); cudaEventElapsedTime(&ms, e0, e1); }
    printf("fill 8 GiB: %.3f ms %.0f GB/s\n", ms, n * 8.0 / ms * 1e-6); cudaFree(d);
  }
  run_store<15, 36>(1000000);       // facet runs, CSR
  run_store<45, 12>(1000000);       // facet runs, 3 x 3 block rows
  run_store<54, 12>(1000000);
  run_store<72, 4>(1000000);        // DG block
  run_store<126, 4>(1000000);       // owned block per sub-problem
  // the same runs with no gaps (what a sub-problem-major layout gives one launch), and with
  // the 3 : 1 gaps of the cell-major layout (a launch writes 1/4 of every cell's block)
  g_gap = 0;
  run_store<45, 12>(1000000); run_store<72, 4>(1000000); run_store<126, 4>(1000000);
  g_gap = 369;
  run_store<126, 4>(1000000);
  return 0;
}
