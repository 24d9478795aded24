/*
 * band_solver.cu -- on-device sparse direct solve for the Newton systems of
 * "1D" (Xs x {0,1}) and thin-strip product meshes.
 *
 * Replaces dolfin.PETScLUSolver("mumps").solve(A, Qdu, b)
 * (simudo/fem/newton_solver.py:124-181).  cuDSS is not installed in this image,
 * so the solve is a hand-written banded LU: with dofs ordered by the
 * x-position of their mesh entity the Jacobian of a product mesh is banded
 * (half bandwidth ~ 27 P for nY = 2), and LU with partial pivoting inside the
 * band (LAPACK gbtf2 organisation) is exact sparse LU for that ordering.
 * Row/column equilibration stands in for MUMPS' ICNTL(6)=5 / ICNTL(8)=77
 * scaling (newton_solver.py:140-142).  One CTA factors and solves one system;
 * a grid of CTAs handles a batch (parameter-sweep ensembles).
 */
#ifdef SB_EMUL_BUILD
#include "cuda_emul.h"
#define SBB_LAUNCH(kernel, grid, block, smem, stream, ...) \
  emul::launch((grid), (block), (smem), [&] { kernel(__VA_ARGS__); })
#define SBB_THREADS 128
#else
#include <cuda_runtime.h>
#define SBB_LAUNCH(kernel, grid, block, smem, stream, ...) \
  kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#define SBB_THREADS 1024
#endif
#include <math.h>
#include <algorithm>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#include "../../include/simudo_b200.h"

extern thread_local char g_sb_err[512];

struct sb_band_solver {
  int64_t n, nnz;
  int32_t kl, ku, ldab;
  int32_t batch;
  int32_t n_refine;    /* iterative-refinement steps with double-double residual */
  int32_t n_ruiz;      /* equilibration sweeps */
  double* rmax;        /* [n * batch] scratch of the equilibration */
  double* cmax;
  int64_t* rowptr;     /* device */
  int32_t* colidx;     /* device */
  int32_t* perm;       /* device: new index of each dof */
  double* ab;          /* [batch][n][ldab] column-major band storage */
  double* rs;          /* [batch][n] row scale */
  double* cs;          /* [batch][n] column scale */
  double* work;        /* [batch][n] permuted right-hand side / solution */
  int32_t* ipiv;       /* [batch][n] */
  int32_t* info;       /* [batch] first zero pivot + 1, 0 = ok */
  double* rdiag;       /* [batch][n] reciprocal pivots (window kernels) */
  bool factored = false;   /* ab / ipiv / rs / cs hold the factors of a completed sb_band_solve */
#ifndef SB_EMUL_BUILD
  cudaGraphExec_t gexec = nullptr;   /* the captured launch sequence of a solve ...         */
  const double* g_vals = nullptr;    /* ... for exactly these arguments                      */
  const double* g_b = nullptr;
  double* g_x = nullptr;
  int32_t g_nsys = 0, g_refine = -1;
  int64_t g_vstride = 0;
  const double* c_vals = nullptr;    /* arguments of the previous call (capture candidates)  */
  const double* c_b = nullptr;
  double* c_x = nullptr;
  int32_t c_nsys = 0, c_refine = -1;
  int64_t c_vstride = 0;
#endif
  double* arow;        /* [batch][n][kl + ku + 1] row-major copy of the scaled band, or null:
                          the shared-memory window kernels are used when it exists */
  size_t window_smem;
  std::vector<void*> allocs;
};

#define BTRY(x)                                                                   \
  do {                                                                            \
    cudaError_t e_ = (x);                                                         \
    if (e_ != cudaSuccess) {                                                      \
      snprintf(g_sb_err, sizeof g_sb_err, "%s failed: %s (%s:%d)", #x,           \
               cudaGetErrorString(e_), __FILE__, __LINE__);                       \
      return SB_ERR_CUDA;                                                         \
    }                                                                             \
  } while (0)

__device__ __forceinline__ void atomic_max_pos_b(double* addr, double v) {
  atomicMax(reinterpret_cast<unsigned long long*>(addr),
            static_cast<unsigned long long>(__double_as_longlong(v)));
}

/* row scale = 1 / max_j |a_ij| */
__global__ void k_band_rowscale(int64_t n, const int64_t* rowptr, const double* vals,
                                int64_t vstride, double* rs, double* cs) {
  const int64_t sys = blockIdx.y;
  const double* v = vals + sys * vstride;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n;
       r += (int64_t)gridDim.x * blockDim.x) {
    double m = 0.0;
    for (int64_t k = rowptr[r]; k < rowptr[r + 1]; ++k) m = fmax(m, fabs(v[k]));
    rs[sys * n + r] = m > 0.0 ? 1.0 / m : 1.0;
    cs[sys * n + r] = 0.0;
  }
}

/* column max of the row-scaled matrix */
__global__ void k_band_colmax(int64_t n, const int64_t* rowptr, const int32_t* colidx,
                              const double* vals, int64_t vstride, const double* rs, double* cs) {
  const int64_t sys = blockIdx.y;
  const double* v = vals + sys * vstride;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n;
       r += (int64_t)gridDim.x * blockDim.x) {
    const double s = rs[sys * n + r];
    for (int64_t k = rowptr[r]; k < rowptr[r + 1]; ++k)
      atomic_max_pos_b(&cs[sys * n + colidx[k]], fabs(v[k]) * s);
  }
}

/* Ruiz equilibration (iterated: r_i /= sqrt(max_j |a_ij|), c_j /= sqrt(max_i |a_ij|) of the
 * current scaled matrix).  One row pass followed by one column pass is not enough for the
 * Newton matrices of wide-gap devices: a continuity row mixes O(1) divergence entries with
 * generation entries ~ carrier density (1e-26 um^-3 for minority carriers at a 2 eV gap), and
 * the weakly determined minority quasi-Fermi modes then come out as noise of size 1e9 eV.
 * A few sweeps balance rows and columns simultaneously -- the job MUMPS' MC64 scaling does in
 * the reference (newton_solver.py:140-142, ICNTL(6) = 5, ICNTL(8) = 77).                    */
__global__ void k_band_ones(int64_t total, double* rs, double* cs) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) { rs[i] = 1.0; cs[i] = 1.0; }
}
__global__ void k_band_ruiz_max(int64_t n, const int64_t* rowptr, const int32_t* colidx,
                                const double* vals, int64_t vstride, const double* rs,
                                const double* cs, double* rmax, double* cmax) {
  const int64_t sys = blockIdx.y;
  const double* v = vals + sys * vstride;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n;
       r += (int64_t)gridDim.x * blockDim.x) {
    const double s = rs[sys * n + r];
    double m = 0.0;
    for (int64_t k = rowptr[r]; k < rowptr[r + 1]; ++k) {
      const double a = fabs(v[k]) * s * cs[sys * n + colidx[k]];
      m = fmax(m, a);
      atomic_max_pos_b(&cmax[sys * n + colidx[k]], a);
    }
    rmax[sys * n + r] = m;
  }
}
__global__ void k_band_ruiz_apply(int64_t total, const double* rmax, const double* cmax,
                                  double* rs, double* cs) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    if (rmax[i] > 0.0) rs[i] /= sqrt(rmax[i]);
    if (cmax[i] > 0.0) cs[i] /= sqrt(cmax[i]);
  }
}

__global__ void k_band_colscale(int64_t n, double* cs, int64_t total) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x)
    cs[i] = cs[i] > 0.0 ? 1.0 / cs[i] : 1.0;
}

/* scatter the scaled CSR into band storage; permuted scaled rhs into work */
__global__ void k_band_fill(int64_t n, const int64_t* rowptr, const int32_t* colidx,
                            const int32_t* perm, const double* vals, int64_t vstride,
                            const double* b, int64_t bstride, const double* rs, const double* cs,
                            double* ab, int32_t ldab, int32_t kv, double* work,
                            double* arow /* [n][kv + 1] row-major band copy, or null */, int32_t kl) {
  const int64_t sys = blockIdx.y;
  const double* v = vals + sys * vstride;
  double* A = ab + sys * n * ldab;
  double* AR = arow ? arow + sys * n * (kv + 1) : nullptr;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n;
       r += (int64_t)gridDim.x * blockDim.x) {
    const double s = rs[sys * n + r];
    const int64_t pr = perm[r];
    for (int64_t k = rowptr[r]; k < rowptr[r + 1]; ++k) {
      const int32_t c = colidx[k];
      const int64_t pc = perm[c];
      const double a = v[k] * s * cs[sys * n + c];
      A[pc * ldab + kv + pr - pc] = a;
      if (AR) AR[pr * (kv + 1) + (pc - pr + kl)] = a;
    }
    work[sys * n + pr] = b[sys * bstride + r] * s;
  }
}

/* permuted, row-scaled right-hand side into work (a solve with stored factors) */
__global__ void k_band_rhs(int64_t n, const int32_t* perm, const double* b, int64_t bstride,
                           const double* rs, double* work) {
  const int64_t sys = blockIdx.y;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n;
       r += (int64_t)gridDim.x * blockDim.x)
    work[sys * n + perm[r]] = b[sys * bstride + r] * rs[sys * n + r];
}

/* one CTA: gbtf2 + gbtrs on one system */
__global__ void __launch_bounds__(SBB_THREADS)
k_band_factor_solve(int64_t n, int32_t kl, int32_t ku, int32_t ldab, double* ab_all,
                    double* work_all, int32_t* ipiv_all, int32_t* info_all) {
  const int64_t sys = blockIdx.x;
  double* AB = ab_all + sys * n * ldab;
  double* x = work_all + sys * n;
  int32_t* ipiv = ipiv_all + sys * n;
  const int kv = kl + ku;
  const int tid = threadIdx.x, nt = blockDim.x;
  __shared__ double s_val[SBB_THREADS];
  __shared__ int s_idx[SBB_THREADS];
  __shared__ int s_jp;
  __shared__ double s_piv;
  int ju = 0;
  int info = 0;
#define ABE(i, j) AB[(int64_t)(j) * ldab + kv + (i) - (j)]
  for (int64_t j = 0; j < n; ++j) {
    const int km = (int)((kl < n - 1 - j) ? kl : (n - 1 - j));
    /* pivot search over rows j .. j+km of column j */
    double best = -1.0; int bi = 0;
    for (int i = tid; i <= km; i += nt) {
      const double a = fabs(ABE(j + i, j));
      if (a > best) { best = a; bi = i; }
    }
    s_val[tid] = best; s_idx[tid] = bi;
    __syncthreads();
    for (int o = nt >> 1; o > 0; o >>= 1) {
      if (tid < o) {
        const double a = s_val[tid + o];
        if (a > s_val[tid] || (a == s_val[tid] && s_idx[tid + o] < s_idx[tid])) {
          s_val[tid] = a; s_idx[tid] = s_idx[tid + o];
        }
      }
      __syncthreads();
    }
    if (tid == 0) {
      s_jp = s_idx[0];
      ipiv[j] = (int32_t)(j + s_idx[0]);
      s_piv = ABE(j + s_idx[0], j);
    }
    __syncthreads();
    const int jp = s_jp;
    const double piv = s_piv;
    {
      const int64_t cand = j + ku + jp;
      const int juc = (int)(cand < n - 1 ? cand : n - 1);
      if (juc > ju) ju = juc;
    }
    if (piv == 0.0) { if (!info) info = (int)(j + 1); continue; }
    /* swap rows j and j+jp over columns j .. ju, swap rhs */
    if (jp != 0) {
      for (int64_t c = j + tid; c <= ju; c += nt) {
        const double t = ABE(j, c);
        ABE(j, c) = ABE(j + jp, c);
        ABE(j + jp, c) = t;
      }
      if (tid == 0) { const double t = x[j]; x[j] = x[j + jp]; x[j + jp] = t; }
    }
    __syncthreads();
    /* multipliers + forward substitution of the rhs */
    const double rp = 1.0 / piv;
    const double xj = x[j];
    for (int i = 1 + tid; i <= km; i += nt) {
      const double l = ABE(j + i, j) * rp;
      ABE(j + i, j) = l;
      x[j + i] -= l * xj;
    }
    __syncthreads();
    /* trailing update */
    const int ncol = ju - (int)j;
    const int total = ncol * km;
    for (int e = tid; e < total; e += nt) {
      const int cc = e / km, i = e - cc * km + 1;
      const int64_t c = j + 1 + cc;
      ABE(j + i, c) -= ABE(j + i, j) * ABE(j, c);
    }
    __syncthreads();
  }
  /* back substitution with U (bandwidth kv) */
  for (int64_t j = n - 1; j >= 0; --j) {
    if (tid == 0) x[j] = x[j] / ABE(j, j);
    __syncthreads();
    const double xj = x[j];
    const int up = (int)((kv < j) ? kv : j);
    for (int i = 1 + tid; i <= up; i += nt) x[j - i] -= ABE(j - i, j) * xj;
    __syncthreads();
  }
#undef ABE
  if (tid == 0) info_all[sys] = info;
}

/* one CTA: solve with the stored factors (multipliers below the diagonal of AB,
 * row interchanges in ipiv) -- used by the iterative-refinement steps */
__global__ void __launch_bounds__(SBB_THREADS)
k_band_resolve(int64_t n, int32_t kl, int32_t ku, int32_t ldab, const double* ab_all,
               double* work_all, const int32_t* ipiv_all) {
  const int64_t sys = blockIdx.x;
  const double* AB = ab_all + sys * n * ldab;
  double* x = work_all + sys * n;
  const int32_t* ipiv = ipiv_all + sys * n;
  const int kv = kl + ku;
  const int tid = threadIdx.x, nt = blockDim.x;
#define ABE(i, j) AB[(int64_t)(j) * ldab + kv + (i) - (j)]
  for (int64_t j = 0; j < n; ++j) {
    const int km = (int)((kl < n - 1 - j) ? kl : (n - 1 - j));
    const int64_t jp = ipiv[j];
    if (tid == 0 && jp != j) { const double t = x[j]; x[j] = x[jp]; x[jp] = t; }
    __syncthreads();
    const double xj = x[j];
    for (int i = 1 + tid; i <= km; i += nt) x[j + i] -= ABE(j + i, j) * xj;
    __syncthreads();
  }
  for (int64_t j = n - 1; j >= 0; --j) {
    if (tid == 0) x[j] = x[j] / ABE(j, j);
    __syncthreads();
    const double xj = x[j];
    const int up = (int)((kv < j) ? kv : j);
    for (int i = 1 + tid; i <= up; i += nt) x[j - i] -= ABE(j - i, j) * xj;
    __syncthreads();
  }
#undef ABE
}

/* ---- shared-memory window variants ----------------------------------------------------------
 * The kernels above walk the band in global memory with one CTA-wide barrier per phase (about
 * fifteen per column, 1024 threads): measured 224 ms per Newton solve of the quasi-1D IB cell
 * (27 k unknowns, kl = ku = 62), i.e. the whole J-V sweep was waiting on it.  Same algorithm
 * (gbtf2: partial pivoting inside the band, multipliers below the diagonal of AB, row
 * interchanges in ipiv -- the factors are interchangeable), but
 *  - the active (kl + 1) x (kv + 1) part of the band lives in shared memory as a ring in both
 *    directions (one spare row and column slot, so the row entering at the next step is stored
 *    while the current one is still read); it is fed from a row-major copy of the scaled band
 *    (k_band_fill), SBB_STAGE rows at a time, and finished L columns / U rows stream out to AB;
 *  - the pivot search is one warp with shuffles; a step has three barriers of a 256-thread CTA;
 *  - the triangular solves (back substitution here, both sweeps in k_band_resolve_window) run
 *    in ONE warp over blocks of SBB_TB columns that the whole CTA has loaded into shared memory
 *    (coalesced, double-buffered): __syncwarp per column instead of __syncthreads.          */
#ifdef SB_EMUL_BUILD
#define SBB_WT 128
static inline int sbb_shfl_down_i(int v, int o) { return (int)__shfl_down_sync(0xffffffffu, (double)v, o); }
static inline double sbb_shfl_down_d(double v, int o) { return __shfl_down_sync(0xffffffffu, v, o); }
static inline int sbb_shfl_idx_i(int v, int src) {       /* broadcast of lane src (emulation: via the shuffle slots) */
  auto& c = *emul::g_ctx;
  const unsigned t = emul::t_tid;
  c.shfl[t] = (double)v;
  __syncwarp();
  const int r = (int)c.shfl[(t / 32) * 32 + src];
  __syncwarp();
  return r;
}
#else
#define SBB_WT 512
__device__ __forceinline__ int sbb_shfl_down_i(int v, int o) { return __shfl_down_sync(0xffffffffu, v, o); }
__device__ __forceinline__ double sbb_shfl_down_d(double v, int o) { return __shfl_down_sync(0xffffffffu, v, o); }
__device__ __forceinline__ int sbb_shfl_idx_i(int v, int src) { return __shfl_sync(0xffffffffu, v, src); }
#endif
#define SBB_STAGE 16            /* rows of the row-major band staged per refill            */
#define SBB_TB 32               /* columns per block of the triangular solves              */

/* back substitution U x = y by warp 0 of the CTA (U: bandwidth kv above the diagonal, column j
 * of U contiguous in AB); blk: shared memory [2][SBB_TB][kv + 1], ring: [kv + 1].           */
__device__ void sbb_back_substitution(int64_t n, int kv, int ldab, const double* __restrict__ AB,
                                      double* __restrict__ x, double* blk, double* ring,
                                      double* xin /* [2][SBB_TB]: entries entering the ring */,
                                      const double* __restrict__ rdiag /* 1 / U(j, j) */) {
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5;
  const int RW = kv + 1;
  const int64_t nblk = (n + SBB_TB - 1) / SBB_TB;
  /* block q covers columns j in (hi - SBB_TB, hi], hi = n - 1 - q SBB_TB; column j stored as
     blk[..][j_local][i] = U(j - i, j), i = 0 .. kv */
  /* t0, tn: rank and number of the threads that take part in the load */
  auto load_block = [&](int64_t q, double* dst, int t0, int tn) {
    const int64_t hi = n - 1 - q * SBB_TB;
    for (int e = t0; e < SBB_TB * RW; e += tn) {
      const int jl = e / RW, i = e - jl * RW;
      const int64_t j = hi - jl;
      dst[e] = (j >= 0 && i <= j) ? (i == 0 ? rdiag[j] : AB[j * ldab + kv - i]) : 0.0;
    }
    /* the entry that enters the ring when column hi - jl retires (already final: y) */
    for (int e = t0; e < SBB_TB; e += tn) {
      const int64_t rin = hi - e - kv - 1;
      xin[(q & 1) * SBB_TB + e] = rin >= 0 ? x[rin] : 0.0;
    }
  };
  load_block(0, blk, tid, nt);
  for (int e = tid; e < RW; e += nt) { const int64_t r = n - 1 - e; ring[e] = r >= 0 ? x[r] : 0.0; }
  /* ring slot of row r: (n - 1 - r) mod RW */
  __syncthreads();
  int slot = 0;                                   /* slot of row j */
  for (int64_t q = 0; q < nblk; ++q) {
    double* cur = blk + (size_t)(q & 1) * SBB_TB * RW;
    if (warp != 0) {
      if (q + 1 < nblk) load_block(q + 1, blk + (size_t)((q + 1) & 1) * SBB_TB * RW, tid - 32, nt - 32);
    } else {
      const int64_t hi = n - 1 - q * SBB_TB;
      for (int jl = 0; jl < SBB_TB; ++jl) {
        const int64_t j = hi - jl;
        if (j < 0) break;
        const double* col = cur + jl * RW;
        const double xj = ring[slot] * col[0];    /* col[0] = 1 / U(j, j) */
        const int up = (int)(kv < j ? kv : j);
        __syncwarp();                             /* every lane has read ring[slot] */
        for (int i = 1 + lane; i <= up; i += 32) {
          int sl = slot + i; if (sl >= RW) sl -= RW;
          ring[sl] -= col[i] * xj;
        }
        if (lane == 0) {
          x[j] = xj;
          ring[slot] = xin[(q & 1) * SBB_TB + jl]; /* row j - kv - 1 enters: takes the slot of row j */
        }
        __syncwarp();
        if (++slot >= RW) slot = 0;
      }
    }
    __syncthreads();
  }
}

/* forward substitution with the stored multipliers and interchanges, same organisation;
 * blk: [2][SBB_TB][kl + 1] (entry 0 unused), piv: int [2][SBB_TB], ring: [kl + 1]         */
__device__ void sbb_forward_substitution(int64_t n, int kl, int kv, int ldab,
                                         const double* __restrict__ AB, const int32_t* __restrict__ ipiv,
                                         double* __restrict__ x, double* blk, int* piv, double* ring,
                                         double* xin) {
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5;
  const int RW = kl + 1;
  const int64_t nblk = (n + SBB_TB - 1) / SBB_TB;
  auto load_block = [&](int64_t q, double* dst, int* pdst, int t0, int tn) {
    const int64_t lo = q * SBB_TB;
    for (int e = t0; e < SBB_TB * RW; e += tn) {
      const int jl = e / RW, i = e - jl * RW;
      const int64_t j = lo + jl;
      dst[e] = (j < n && i >= 1 && j + i < n) ? AB[j * ldab + kv + i] : 0.0;
    }
    for (int e = t0; e < SBB_TB; e += tn) {
      pdst[e] = lo + e < n ? (int)(ipiv[lo + e] - (lo + e)) : 0;
      const int64_t rin = lo + e + kl + 1;          /* untouched right-hand side entry */
      xin[(q & 1) * SBB_TB + e] = rin < n ? x[rin] : 0.0;
    }
  };
  load_block(0, blk, piv, tid, nt);
  for (int e = tid; e < RW; e += nt) ring[e] = e < n ? x[e] : 0.0;      /* slot of row r: r mod RW */
  __syncthreads();
  int slot = 0;
  for (int64_t q = 0; q < nblk; ++q) {
    double* cur = blk + (size_t)(q & 1) * SBB_TB * RW;
    int* pc = piv + (q & 1) * SBB_TB;
    if (warp != 0) {
      if (q + 1 < nblk)
        load_block(q + 1, blk + (size_t)((q + 1) & 1) * SBB_TB * RW, piv + ((q + 1) & 1) * SBB_TB,
                   tid - 32, nt - 32);
    } else {
      const int64_t lo = q * SBB_TB;
      for (int jl = 0; jl < SBB_TB; ++jl) {
        const int64_t j = lo + jl;
        if (j >= n) break;
        const double* col = cur + jl * RW;
        const int jp = pc[jl];
        if (jp != 0 && lane == 0) {
          int sp = slot + jp; if (sp >= RW) sp -= RW;
          const double t = ring[slot]; ring[slot] = ring[sp]; ring[sp] = t;
        }
        __syncwarp();
        const double xj = ring[slot];
        const int km = (int)(kl < n - 1 - j ? kl : n - 1 - j);
        __syncwarp();                             /* every lane has read ring[slot] */
        for (int i = 1 + lane; i <= km; i += 32) {
          int sl = slot + i; if (sl >= RW) sl -= RW;
          ring[sl] -= col[i] * xj;
        }
        if (lane == 0) {
          x[j] = xj;
          ring[slot] = xin[(q & 1) * SBB_TB + jl];
        }
        __syncwarp();
        if (++slot >= RW) slot = 0;
      }
    }
    __syncthreads();
  }
}

static size_t sbb_window_smem(int kl, int ku) {
  const size_t kv = (size_t)kl + ku, WR = kl + 2, WC = kv + 2, LD = WR | 1;
  const size_t fac = WC * LD + WR + (size_t)SBB_STAGE * (kv + 2);
  const size_t tri = 2 * (size_t)SBB_TB * (kv + 1) + (kv + 1) + SBB_TB + 2 * SBB_TB;   /* ints as doubles */
  return 8 * (fac > tri ? fac : tri) + 64;
}

__global__ void __launch_bounds__(SBB_WT)
k_band_factor_window(int64_t n, int32_t kl, int32_t ku, int32_t ldab, const double* arow_all,
                     double* ab_all, double* work_all, int32_t* ipiv_all, int32_t* info_all,
                     double* rdiag_all) {
#ifdef SB_EMUL_BUILD
  double* sm = reinterpret_cast<double*>(emul::dyn_smem());
#else
  extern __shared__ __align__(16) double sm[];
#endif
  const int64_t sys = blockIdx.x;
  const int kv = kl + ku, WR = kl + 2, WC = kv + 2, RS = kv + 2;
  const int LD = WR | 1;     /* odd column stride of the window: accesses along a row (the
                                interchange, the U row, the entering row) hit 32 banks */
  double* AB = ab_all + sys * n * ldab;
  const double* AR = arow_all + sys * n * (kv + 1);
  double* x = work_all + sys * n;
  int32_t* ipiv = ipiv_all + sys * n;
  double* S = sm;                                 /* window [WC][LD]: S[(c mod WC) LD + (r mod WR)] */
  double* xs = S + (size_t)WC * LD;               /* rhs of the window rows [WR]                     */
  double* stage = xs + WR;                        /* [SBB_STAGE][kv + 2]: band row + its rhs         */
  __shared__ int s_jp;
  __shared__ double s_piv, s_rp, s_diag, s_xj, s_xold;
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarp = nt >> 5;
#define ABE(i, j) AB[(int64_t)(j) * ldab + kv + (i) - (j)]
  for (int e = tid; e < WC * LD + WR; e += nt) S[e] = 0.0;
  __syncthreads();
  /* rows 0 .. kl enter directly */
  for (int64_t r = 0; r <= kl && r < n; ++r) {
    for (int t = tid; t <= kv; t += nt) {
      const int64_t c = r - kl + t;
      if (c >= 0 && c < n) S[(int)(c % WC) * LD + (int)(r % WR)] = AR[r * (kv + 1) + t];
    }
    if (tid == 0) xs[r % WR] = x[r];
  }
  /* rows kl + 1 + [0, SBB_STAGE) wait in the staging area */
  int64_t stage_row0 = kl + 1;
  for (int e = tid; e < SBB_STAGE * RS; e += nt) {
    const int q = e / RS, t = e - q * RS;
    const int64_t r = stage_row0 + q;
    stage[e] = r < n ? (t <= kv ? AR[r * (kv + 1) + t] : x[r]) : 0.0;
  }
  __syncthreads();
  int jr = 0, jc = 0, ju = 0, info = 0;
  double* abj = AB + kv;                          /* &ABE(j, j): ABE(j + i, j) = abj[i],
                                                     ABE(j, j + t) = abj[t (ldab - 1)]      */
  double* rdiag = rdiag_all + sys * n;            /* 1 / U(j, j) for the triangular solves   */
  for (int64_t j = 0; j < n; ++j, abj += ldab) {
    const int km = (int)((kl < n - 1 - j) ? kl : (n - 1 - j));
    /* phase 1: pivot search over rows j .. j + km of column j, one warp with shuffles */
    if (warp == 0) {
      double best = -1.0; int bi = 0;
      for (int i = lane; i <= km; i += 32) {
        int rs = jr + i; if (rs >= WR) rs -= WR;
        const double a = fabs(S[jc * LD + rs]);
        if (a > best) { best = a; bi = i; }
      }
      for (int o = 16; o > 0; o >>= 1) {
        const double ob = sbb_shfl_down_d(best, o);
        const int oi = sbb_shfl_down_i(bi, o);
        if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
      }
      if (lane == 0) {
        int rsp0 = jr + bi; if (rsp0 >= WR) rsp0 -= WR;
        const double pv = S[jc * LD + rsp0];
        s_jp = bi;
        s_piv = pv;
        s_rp = 1.0 / pv;
        s_diag = S[jc * LD + jr];
        s_xj = xs[rsp0];
        s_xold = xs[jr];
        ipiv[j] = (int32_t)(j + bi);
        rdiag[j] = s_rp;
      }
    }
    __syncthreads();
    const int jp = s_jp;
    const double piv = s_piv;
    {
      const int64_t cand = j + ku + jp;
      const int juc = (int)(cand < n - 1 ? cand : n - 1);
      if (juc > ju) ju = juc;
    }
    const bool sing = (piv == 0.0);
    if (sing && !info) info = (int)(j + 1);
    int rsp = jr + jp; if (rsp >= WR) rsp -= WR;
    /* phase 2: multipliers of column j + forward elimination of the rhs (threads over rows);
       interchange of rows j and j + jp in the columns right of j (threads over columns) */
    if (!sing) {
      const double rp = s_rp, xjn = s_xj;
      for (int i = 1 + tid; i <= km; i += nt) {
        int rs = jr + i; if (rs >= WR) rs -= WR;
        const double a = (i == jp) ? s_diag : S[jc * LD + rs];
        const double l = a * rp;
        S[jc * LD + rs] = l;
        const double xi = (i == jp) ? s_xold : xs[rs];
        xs[rs] = xi - l * xjn;
      }
      if (tid == 0) { S[jc * LD + jr] = piv; xs[jr] = xjn; }
      if (jp != 0) {
        /* the upper warps take the interchange: the lower ones are busy with the column */
        for (int cc = nt - 1 - tid; cc < ju - (int)j; cc += nt) {
          int cs = jc + 1 + cc; if (cs >= WC) cs -= WC;
          const double t = S[cs * LD + jr];
          S[cs * LD + jr] = S[cs * LD + rsp];
          S[cs * LD + rsp] = t;
        }
      }
    }
    __syncthreads();
    /* phase 3: trailing update; the finished L column and U row stream out; the row entering
       at the next step goes to the spare row / column slots */
    if (!sing) {
      const int ncol = ju - (int)j;
      /* a lane keeps the multipliers of its rows in registers (<= 4 rows: kl <= 128) and
         streams over the columns of its warp */
      const int nq = (km + 31) >> 5;
      double lv[4]; int ro[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int i = 1 + lane + 32 * q;
        int rs = jr + (i <= km ? i : 0); if (rs >= WR) rs -= WR;
        ro[q] = rs;
        lv[q] = (q < nq && i <= km) ? S[jc * LD + rs] : 0.0;
      }
      /* 4 columns in flight per warp: the shared-memory round trips are the cost */
      for (int c0 = warp * 4; c0 < ncol; c0 += nwarp * 4) {
        double* colp[4]; double u[4]; double a[4][4];
#pragma unroll
        for (int m = 0; m < 4; ++m) {
          int cs = jc + 1 + (c0 + m < ncol ? c0 + m : c0); if (cs >= WC) cs -= WC;
          colp[m] = S + cs * LD;
          u[m] = colp[m][jr];
        }
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (q < nq) {
#pragma unroll
            for (int m = 0; m < 4; ++m) a[q][m] = colp[m][ro[q]];
          }
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (q < nq && 1 + lane + 32 * q <= km) {
#pragma unroll
            for (int m = 0; m < 4; ++m)
              if (c0 + m < ncol) colp[m][ro[q]] = a[q][m] - lv[q] * u[m];
          }
      }
    }
    for (int i = 1 + tid; i <= km; i += nt) {
      int rs = jr + i; if (rs >= WR) rs -= WR;
      abj[i] = S[jc * LD + rs];
    }
    for (int t = tid; t <= kv && j + t < n; t += nt) {
      int cs = jc + t; if (cs >= WC) cs -= WC;
      abj[(int64_t)t * (ldab - 1)] = S[cs * LD + jr];
    }
    if (tid == 0) x[j] = xs[jr];
    {
      const int64_t r = j + kl + 1;               /* entering row; its columns are j + 1 .. j + kv + 1 */
      int rspare = jr + kl + 1; if (rspare >= WR) rspare -= WR;
      int cspare = jc + kv + 1; if (cspare >= WC) cspare -= WC;
      const double* srow = stage + (int)(r - stage_row0) * RS;
      /* the spare column slot held a retired column: clear it, then its corner arrives below */
      for (int e = tid; e < WR; e += nt) if (e != rspare) S[cspare * LD + e] = 0.0;
      for (int t = tid; t <= kv; t += nt) {
        int cs = jc + 1 + t; if (cs >= WC) cs -= WC;
        S[cs * LD + rspare] = (r < n && j + 1 + t < n) ? srow[t] : 0.0;
      }
      if (tid == 0) xs[rspare] = r < n ? srow[kv + 1] : 0.0;
    }
    __syncthreads();
    if (j + kl + 1 - stage_row0 == SBB_STAGE - 1) {       /* staging area consumed: refill */
      stage_row0 += SBB_STAGE;
      for (int e = tid; e < SBB_STAGE * RS; e += nt) {
        const int q = e / RS, t = e - q * RS;
        const int64_t r = stage_row0 + q;
        stage[e] = r < n ? (t <= kv ? AR[r * (kv + 1) + t] : x[r]) : 0.0;
      }
      __syncthreads();
    }
    if (++jr >= WR) jr = 0;
    if (++jc >= WC) jc = 0;
  }
#undef ABE
  if (tid == 0) info_all[sys] = info;
  __syncthreads();
  sbb_back_substitution(n, kv, ldab, AB, x, sm, sm + 2 * (size_t)SBB_TB * (kv + 1),
                        sm + 2 * (size_t)SBB_TB * (kv + 1) + (kv + 1) + SBB_TB, rdiag);
}

/* forward + back substitution with the stored factors (iterative refinement) */
__global__ void __launch_bounds__(SBB_WT)
k_band_resolve_window(int64_t n, int32_t kl, int32_t ku, int32_t ldab, const double* ab_all,
                      double* work_all, const int32_t* ipiv_all, const double* rdiag_all) {
#ifdef SB_EMUL_BUILD
  double* sm = reinterpret_cast<double*>(emul::dyn_smem());
#else
  extern __shared__ __align__(16) double sm[];
#endif
  const int64_t sys = blockIdx.x;
  const int kv = kl + ku;
  const double* AB = ab_all + sys * n * ldab;
  double* x = work_all + sys * n;
  double* ring = sm + 2 * (size_t)SBB_TB * (kv + 1);
  double* xin = ring + (kv + 1) + SBB_TB;
  sbb_forward_substitution(n, kl, kv, ldab, AB, ipiv_all + sys * n, x, sm,
                           reinterpret_cast<int*>(ring + kv + 1), ring, xin);
  __syncthreads();
  sbb_back_substitution(n, kv, ldab, AB, x, sm, ring, xin, rdiag_all + sys * n);
}

/* r = b - A x with the row sums accumulated in double-double (error-free
 * products via FMA, TwoSum accumulation); r is written pre-scaled and permuted
 * into the solver's work vector: work[perm[row]] = rs[row] * r[row].
 * The Newton matrices mix entries of size 1 and 1e28 in one row; a plain double
 * residual is too noisy to refine the weakly determined quasi-Fermi modes. */
__global__ void k_band_residual_dd(int64_t n, const int64_t* rowptr, const int32_t* colidx,
                                   const int32_t* perm, const double* vals, int64_t vstride,
                                   const double* b, const double* x, int64_t xstride,
                                   const double* rs, double* work) {
  const int64_t sys = blockIdx.y;
  const double* v = vals + sys * vstride;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n;
       r += (int64_t)gridDim.x * blockDim.x) {
    double hi = b[sys * xstride + r], lo = 0.0;
    for (int64_t k = rowptr[r]; k < rowptr[r + 1]; ++k) {
      const double a = -v[k], xv = x[sys * xstride + colidx[k]];
      const double p = a * xv;
      const double pe = fma(a, xv, -p);
      const double s = hi + p;
      const double bb = s - hi;
      const double se = (hi - (s - bb)) + (p - bb);
      hi = s;
      lo += se + pe;
    }
    work[sys * n + perm[r]] = (hi + lo) * rs[sys * n + r];
  }
}

/* r = b - A x for a general CSR matrix, same compensated accumulation (used by the
 * multifrontal solver's iterative refinement, simudo_b200/fem/multifrontal.py) */
__global__ void k_csr_residual_dd(int64_t n, const int64_t* rowptr, const int32_t* colidx,
                                  const double* vals, const double* x, const double* b, double* r) {
  for (int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; row < n;
       row += (int64_t)gridDim.x * blockDim.x) {
    double hi = b[row], lo = 0.0;
    for (int64_t k = rowptr[row]; k < rowptr[row + 1]; ++k) {
      const double a = -vals[k], xv = x[colidx[k]];
      const double p = a * xv;
      const double pe = fma(a, xv, -p);
      const double s = hi + p;
      const double bb = s - hi;
      const double se = (hi - (s - bb)) + (p - bb);
      hi = s;
      lo += se + pe;
    }
    r[row] = hi + lo;
  }
}

/* x[r] += cs[r] * work[perm[r]] */
__global__ void k_band_accumulate(int64_t n, const int32_t* perm, const double* cs,
                                  const double* work, double* x, int64_t xstride) {
  const int64_t sys = blockIdx.y;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n;
       r += (int64_t)gridDim.x * blockDim.x)
    x[sys * xstride + r] += cs[sys * n + r] * work[sys * n + perm[r]];
}

/* x[r] = cs[r] * work[perm[r]] */
__global__ void k_band_unpermute(int64_t n, const int32_t* perm, const double* cs,
                                 const double* work, double* x, int64_t xstride) {
  const int64_t sys = blockIdx.y;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n;
       r += (int64_t)gridDim.x * blockDim.x)
    x[sys * xstride + r] = cs[sys * n + r] * work[sys * n + perm[r]];
}

template <typename T>
static int balloc(sb_band_solver* s, size_t n, T** out) {
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, (n ? n : 1) * sizeof(T));
  if (e != cudaSuccess) {
    snprintf(g_sb_err, sizeof g_sb_err, "cudaMalloc(%zu bytes) failed: %s", n * sizeof(T),
             cudaGetErrorString(e));
    return SB_ERR_ALLOC;
  }
  s->allocs.push_back(p);
  *out = static_cast<T*>(p);
  return SB_OK;
}

extern "C" void sb_band_destroy(sb_band_solver* s) {
  if (!s) return;
#ifndef SB_EMUL_BUILD
  if (s->gexec) cudaGraphExecDestroy(s->gexec);
#endif
  for (void* p : s->allocs) cudaFree(p);
  delete s;
}

extern "C" int sb_band_create(int64_t n, const int64_t* h_rowptr, const int32_t* h_colidx,
                              const int32_t* h_perm, int32_t batch, int64_t max_band_bytes,
                              sb_band_solver** out) {
  if (!h_rowptr || !h_colidx || !h_perm || !out || n <= 0 || batch <= 0) {
    snprintf(g_sb_err, sizeof g_sb_err, "bad argument to sb_band_create");
    return SB_ERR_ARG;
  }
  int64_t kl = 0, ku = 0;
  for (int64_t r = 0; r < n; ++r)
    for (int64_t k = h_rowptr[r]; k < h_rowptr[r + 1]; ++k) {
      const int64_t d = (int64_t)h_perm[r] - (int64_t)h_perm[h_colidx[k]];
      if (d > kl) kl = d;
      if (-d > ku) ku = -d;
    }
  const int64_t ldab = 2 * kl + ku + 1;
  const int64_t bytes = ldab * n * 8 * batch;
  if (max_band_bytes > 0 && bytes > max_band_bytes) {
    snprintf(g_sb_err, sizeof g_sb_err,
             "banded LU needs %lld bytes (n=%lld, kl=%lld, ku=%lld): mesh is not a thin "
             "strip; a general sparse direct solver (cuDSS) is not available in this build",
             (long long)bytes, (long long)n, (long long)kl, (long long)ku);
    return SB_ERR_UNSUPPORTED;
  }
  sb_band_solver* s = new sb_band_solver();
  s->n = n; s->nnz = h_rowptr[n]; s->kl = (int32_t)kl; s->ku = (int32_t)ku;
  s->ldab = (int32_t)ldab; s->batch = batch; s->n_refine = 2;
  int rc = SB_OK;
  if (rc == SB_OK) rc = balloc(s, (size_t)n + 1, &s->rowptr);
  if (rc == SB_OK) rc = balloc(s, (size_t)s->nnz, &s->colidx);
  if (rc == SB_OK) rc = balloc(s, (size_t)n, &s->perm);
  if (rc == SB_OK) rc = balloc(s, (size_t)(ldab * n * batch), &s->ab);
  if (rc == SB_OK) rc = balloc(s, (size_t)(n * batch), &s->rs);
  if (rc == SB_OK) rc = balloc(s, (size_t)(n * batch), &s->cs);
  if (rc == SB_OK) rc = balloc(s, (size_t)(n * batch), &s->work);
  if (rc == SB_OK) rc = balloc(s, (size_t)(n * batch), &s->rmax);
  if (rc == SB_OK) rc = balloc(s, (size_t)(n * batch), &s->cmax);
  s->n_ruiz = 8;
  if (rc == SB_OK) rc = balloc(s, (size_t)(n * batch), &s->ipiv);
  if (rc == SB_OK) rc = balloc(s, (size_t)batch, &s->info);
  /* the window kernels need (kl + 2)(kl + ku + 2) doubles of shared memory (227 KB per CTA on
     sm_100a); SIMUDO_B200_BAND_GLOBAL=1 keeps the global-memory kernels (tests compare both) */
  s->arow = nullptr; s->rdiag = nullptr;
  s->window_smem = sbb_window_smem((int)kl, (int)ku);
  size_t max_dyn = 227 * 1024;
#ifndef SB_EMUL_BUILD
  {
    /* dynamic shared memory a launch of the window kernels may ask for on this device: the
       opt-in maximum minus their static shared memory.  The attribute is per kernel, not per
       solver, and solvers of different bandwidths (PDD system, CG2 optics matrices) coexist
       in one process, possibly on several threads: opt in to the maximum once and for all */
    int dev = 0, optin = 0;
    cudaFuncAttributes fa1, fa2;
    BTRY(cudaGetDevice(&dev));
    BTRY(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    BTRY(cudaFuncGetAttributes(&fa1, k_band_factor_window));
    BTRY(cudaFuncGetAttributes(&fa2, k_band_resolve_window));
    const size_t st = fa1.sharedSizeBytes > fa2.sharedSizeBytes ? fa1.sharedSizeBytes : fa2.sharedSizeBytes;
    max_dyn = (size_t)optin > st ? (size_t)optin - st : 0;
    BTRY(cudaFuncSetAttribute(k_band_factor_window, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              (int)((size_t)optin - fa1.sharedSizeBytes)));
    BTRY(cudaFuncSetAttribute(k_band_resolve_window, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              (int)((size_t)optin - fa2.sharedSizeBytes)));
  }
#endif
  if (rc == SB_OK && s->window_smem <= max_dyn && kl <= 128 && !getenv("SIMUDO_B200_BAND_GLOBAL")) {
    rc = balloc(s, (size_t)(n * (kl + ku + 1) * batch), &s->arow);
    if (rc == SB_OK) rc = balloc(s, (size_t)(n * batch), &s->rdiag);
  }
  if (rc != SB_OK) { sb_band_destroy(s); return rc; }
  if (cudaMemcpy(s->rowptr, h_rowptr, (n + 1) * sizeof(int64_t), cudaMemcpyHostToDevice) != cudaSuccess ||
      cudaMemcpy(s->colidx, h_colidx, s->nnz * sizeof(int32_t), cudaMemcpyHostToDevice) != cudaSuccess ||
      cudaMemcpy(s->perm, h_perm, n * sizeof(int32_t), cudaMemcpyHostToDevice) != cudaSuccess) {
    snprintf(g_sb_err, sizeof g_sb_err, "sb_band_create: upload of the sparsity structure failed: %s",
             cudaGetErrorString(cudaGetLastError()));
    sb_band_destroy(s);
    return SB_ERR_CUDA;
  }
  *out = s;
  return SB_OK;
}

extern "C" int sb_band_set_refinement(sb_band_solver* s, int32_t steps) {
  if (!s || steps < 0 || steps > 16) return SB_ERR_ARG;
  s->n_refine = steps;
  return SB_OK;
}

extern "C" int sb_band_info(const sb_band_solver* s, int64_t* n, int32_t* kl, int32_t* ku) {
  if (!s) return SB_ERR_ARG;
  if (n) *n = s->n;
  if (kl) *kl = s->kl;
  if (ku) *ku = s->ku;
  return SB_OK;
}

/* Solve A x = b for `nsys` <= batch systems sharing the sparsity pattern:
 * d_vals [nsys][vstride], d_b / d_x [nsys][n].  h_info (optional, host) receives the
 * LAPACK-style info per system (synchronises the stream when given).            */
/* the launch sequence of one solve (about 40 launches), enqueued on `st` */
static int band_enqueue(sb_band_solver* s, int32_t nsys, const double* d_vals, int64_t vstride,
                        const double* d_b, double* d_x, cudaStream_t st) {
  const int64_t n = s->n;
  const unsigned gx = (unsigned)((n + 255) / 256 < 1184 ? (n + 255) / 256 : 1184);
#ifdef SB_EMUL_BUILD
  for (int sys = 0; sys < nsys; ++sys) {
    /* the emulator has a 1D grid: run the batch entries one after another */
    (void)gx;
  }
#endif
  BTRY(cudaMemsetAsync(s->ab, 0, (size_t)s->ldab * n * nsys * sizeof(double), st));
  if (s->arow)
    BTRY(cudaMemsetAsync(s->arow, 0, (size_t)(s->kl + s->ku + 1) * n * nsys * sizeof(double), st));
#ifdef SB_EMUL_BUILD
#define GRID2(gx_, ny_) (gx_)
#else
#define GRID2(gx_, ny_) dim3((gx_), (ny_))
#endif
#ifndef SB_EMUL_BUILD
  SBB_LAUNCH(k_band_ones, gx, 256, 0, st, n * nsys, s->rs, s->cs);
  for (int it = 0; it < s->n_ruiz; ++it) {
    BTRY(cudaMemsetAsync(s->cmax, 0, (size_t)n * nsys * sizeof(double), st));
    SBB_LAUNCH(k_band_ruiz_max, GRID2(gx, nsys), 256, 0, st, n, s->rowptr, s->colidx, d_vals,
               vstride, s->rs, s->cs, s->rmax, s->cmax);
    SBB_LAUNCH(k_band_ruiz_apply, gx, 256, 0, st, n * nsys, s->rmax, s->cmax, s->rs, s->cs);
  }
  SBB_LAUNCH(k_band_fill, GRID2(gx, nsys), 256, 0, st, n, s->rowptr, s->colidx, s->perm, d_vals,
             vstride, d_b, n, s->rs, s->cs, s->ab, s->ldab, s->kl + s->ku, s->work, s->arow, s->kl);
  if (s->arow)
    SBB_LAUNCH(k_band_factor_window, (unsigned)nsys, SBB_WT, s->window_smem, st, n, s->kl, s->ku,
               s->ldab, s->arow, s->ab, s->work, s->ipiv, s->info, s->rdiag);
  else
    SBB_LAUNCH(k_band_factor_solve, (unsigned)nsys, SBB_THREADS, 0, st, n, s->kl, s->ku, s->ldab,
               s->ab, s->work, s->ipiv, s->info);
  SBB_LAUNCH(k_band_unpermute, GRID2(gx, nsys), 256, 0, st, n, s->perm, s->cs, s->work, d_x, n);
  for (int it = 0; it < s->n_refine; ++it) {
    SBB_LAUNCH(k_band_residual_dd, GRID2(gx, nsys), 256, 0, st, n, s->rowptr, s->colidx, s->perm,
               d_vals, vstride, d_b, d_x, n, s->rs, s->work);
    if (s->arow)
      SBB_LAUNCH(k_band_resolve_window, (unsigned)nsys, SBB_WT, s->window_smem, st, n, s->kl, s->ku,
                 s->ldab, s->ab, s->work, s->ipiv, s->rdiag);
    else
      SBB_LAUNCH(k_band_resolve, (unsigned)nsys, SBB_THREADS, 0, st, n, s->kl, s->ku, s->ldab,
                 s->ab, s->work, s->ipiv);
    SBB_LAUNCH(k_band_accumulate, GRID2(gx, nsys), 256, 0, st, n, s->perm, s->cs, s->work, d_x, n);
  }
#else
  if (nsys != 1) {
    snprintf(g_sb_err, sizeof g_sb_err, "emulation build solves one system at a time");
    return SB_ERR_UNSUPPORTED;
  }
  SBB_LAUNCH(k_band_ones, 4u, 64, 0, st, n * nsys, s->rs, s->cs);
  for (int it = 0; it < s->n_ruiz; ++it) {
    BTRY(cudaMemsetAsync(s->cmax, 0, (size_t)n * nsys * sizeof(double), st));
    SBB_LAUNCH(k_band_ruiz_max, 1u, 1, 0, st, n, s->rowptr, s->colidx, d_vals, vstride, s->rs,
               s->cs, s->rmax, s->cmax);
    SBB_LAUNCH(k_band_ruiz_apply, 4u, 64, 0, st, n * nsys, s->rmax, s->cmax, s->rs, s->cs);
  }
  SBB_LAUNCH(k_band_fill, 4u, 64, 0, st, n, s->rowptr, s->colidx, s->perm, d_vals, vstride, d_b,
             n, s->rs, s->cs, s->ab, s->ldab, s->kl + s->ku, s->work, s->arow, s->kl);
  if (s->arow)
    SBB_LAUNCH(k_band_factor_window, 1u, SBB_WT, s->window_smem, st, n, s->kl, s->ku, s->ldab,
               s->arow, s->ab, s->work, s->ipiv, s->info, s->rdiag);
  else
    SBB_LAUNCH(k_band_factor_solve, 1u, SBB_THREADS, 0, st, n, s->kl, s->ku, s->ldab, s->ab,
               s->work, s->ipiv, s->info);
  SBB_LAUNCH(k_band_unpermute, 4u, 64, 0, st, n, s->perm, s->cs, s->work, d_x, n);
  for (int it = 0; it < s->n_refine; ++it) {
    SBB_LAUNCH(k_band_residual_dd, 4u, 64, 0, st, n, s->rowptr, s->colidx, s->perm, d_vals,
               vstride, d_b, d_x, n, s->rs, s->work);
    if (s->arow)
      SBB_LAUNCH(k_band_resolve_window, 1u, SBB_WT, s->window_smem, st, n, s->kl, s->ku, s->ldab,
                 s->ab, s->work, s->ipiv, s->rdiag);
    else
      SBB_LAUNCH(k_band_resolve, 1u, SBB_THREADS, 0, st, n, s->kl, s->ku, s->ldab, s->ab, s->work,
                 s->ipiv);
    SBB_LAUNCH(k_band_accumulate, 4u, 64, 0, st, n, s->perm, s->cs, s->work, d_x, n);
  }
#endif
  BTRY(cudaGetLastError());
  return SB_OK;
}

extern "C" int sb_band_solve(sb_band_solver* s, int32_t nsys, const double* d_vals,
                             int64_t vstride, const double* d_b, double* d_x,
                             int32_t* h_info, void* stream) {
  if (!s || !d_vals || !d_b || !d_x || nsys <= 0 || nsys > s->batch) {
    snprintf(g_sb_err, sizeof g_sb_err, "bad argument to sb_band_solve");
    return SB_ERR_ARG;
  }
  cudaStream_t st = (cudaStream_t)stream;
#ifndef SB_EMUL_BUILD
  /* A Newton loop calls this with the same buffers every iteration: the ~40 launches of a
     solve are captured once into a CUDA graph and replayed (one launch per solve; with eight
     continuation chains in flight per GPU the launch rate of the process was the limit of a
     J-V sweep).  The legacy default stream cannot be captured: direct launches there. */
  const bool same = s->gexec && s->g_vals == d_vals && s->g_b == d_b && s->g_x == d_x &&
                    s->g_nsys == nsys && s->g_vstride == vstride && s->g_refine == s->n_refine;
  if (same) {
    BTRY(cudaGraphLaunch(s->gexec, st));
  } else {
    if (s->gexec) { cudaGraphExecDestroy(s->gexec); s->gexec = nullptr; }
    bool captured = false;
    /* capture only what has been seen twice in a row: one-off argument sets are launched
       directly (capturing + instantiating costs more than the launches) */
    const bool again = s->c_vals == d_vals && s->c_b == d_b && s->c_x == d_x &&
                       s->c_nsys == nsys && s->c_vstride == vstride && s->c_refine == s->n_refine;
    s->c_vals = d_vals; s->c_b = d_b; s->c_x = d_x; s->c_nsys = nsys; s->c_vstride = vstride;
    s->c_refine = s->n_refine;
    if (again && st != nullptr && !getenv("SIMUDO_B200_NO_GRAPH") &&
        cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
      const int rc = band_enqueue(s, nsys, d_vals, vstride, d_b, d_x, st);
      cudaGraph_t g = nullptr;
      const cudaError_t e = cudaStreamEndCapture(st, &g);
      if (rc == SB_OK && e == cudaSuccess && g &&
          cudaGraphInstantiate(&s->gexec, g, 0) == cudaSuccess) {
        s->g_vals = d_vals; s->g_b = d_b; s->g_x = d_x; s->g_nsys = nsys;
        s->g_vstride = vstride; s->g_refine = s->n_refine;
        captured = true;
      } else {
        s->gexec = nullptr;
      }
      if (g) cudaGraphDestroy(g);
      cudaGetLastError();                         /* a failed capture leaves a sticky-free error */
    } else {
      cudaGetLastError();
    }
    if (captured) {
      BTRY(cudaGraphLaunch(s->gexec, st));
    } else {
      const int rc = band_enqueue(s, nsys, d_vals, vstride, d_b, d_x, st);
      if (rc != SB_OK) return rc;
    }
  }
#else
  {
    const int rc = band_enqueue(s, nsys, d_vals, vstride, d_b, d_x, st);
    if (rc != SB_OK) return rc;
  }
#endif
  const int64_t n = s->n;
  (void)n;
  s->factored = true;
  if (h_info) {
#ifndef SB_EMUL_BUILD
    BTRY(cudaMemcpyAsync(h_info, s->info, nsys * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    BTRY(cudaStreamSynchronize(st));
#else
    memcpy(h_info, s->info, nsys * sizeof(int32_t));
#endif
  }
  return SB_OK;
}

/* x = A^-1 b with the factors (and scalings) of the last sb_band_solve: for a matrix that does
 * not change between solves -- the CG2 mass matrix of the optical projection
 * (dolfin.project, simudo/fem/assign.py:12-47, refactors it on every call).            */
extern "C" int sb_band_solve_factored(sb_band_solver* s, int32_t nsys, const double* d_b,
                                      double* d_x, void* stream) {
  if (!s || !d_b || !d_x || nsys <= 0 || nsys > s->batch || !s->factored) {
    snprintf(g_sb_err, sizeof g_sb_err, !s || s->factored ? "bad argument to sb_band_solve_factored"
                                                          : "sb_band_solve_factored before sb_band_solve");
    return SB_ERR_ARG;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t n = s->n;
#ifdef SB_EMUL_BUILD
  if (nsys != 1) return SB_ERR_UNSUPPORTED;
  SBB_LAUNCH(k_band_rhs, 4u, 64, 0, st, n, s->perm, d_b, n, s->rs, s->work);
  if (s->arow)
    SBB_LAUNCH(k_band_resolve_window, 1u, SBB_WT, s->window_smem, st, n, s->kl, s->ku, s->ldab,
               s->ab, s->work, s->ipiv, s->rdiag);
  else
    SBB_LAUNCH(k_band_resolve, 1u, SBB_THREADS, 0, st, n, s->kl, s->ku, s->ldab, s->ab, s->work,
               s->ipiv);
  SBB_LAUNCH(k_band_unpermute, 4u, 64, 0, st, n, s->perm, s->cs, s->work, d_x, n);
#else
  const unsigned gx = (unsigned)((n + 255) / 256 < 1184 ? (n + 255) / 256 : 1184);
  SBB_LAUNCH(k_band_rhs, dim3(gx, nsys), 256, 0, st, n, s->perm, d_b, n, s->rs, s->work);
  if (s->arow)
    SBB_LAUNCH(k_band_resolve_window, (unsigned)nsys, SBB_WT, s->window_smem, st, n, s->kl, s->ku,
               s->ldab, s->ab, s->work, s->ipiv, s->rdiag);
  else
    SBB_LAUNCH(k_band_resolve, (unsigned)nsys, SBB_THREADS, 0, st, n, s->kl, s->ku, s->ldab,
               s->ab, s->work, s->ipiv);
  SBB_LAUNCH(k_band_unpermute, dim3(gx, nsys), 256, 0, st, n, s->perm, s->cs, s->work, d_x, n);
#endif
  BTRY(cudaGetLastError());
  return SB_OK;
}

/* r = b - A x in double-double accumulation for a device CSR matrix (rows contiguous,
 * any column order).  Replaces nothing in the reference (MUMPS refines internally,
 * newton_solver.py:140-146); used by the multifrontal solver's refinement steps.   */
extern "C" int sb_csr_residual_dd(int64_t n, const int64_t* d_rowptr, const int32_t* d_colidx,
                                  const double* d_vals, const double* d_x, const double* d_b,
                                  double* d_r, void* stream) {
  if (n < 0 || !d_rowptr || !d_colidx || !d_vals || !d_x || !d_b || !d_r) {
    snprintf(g_sb_err, sizeof g_sb_err, "bad argument to sb_csr_residual_dd");
    return SB_ERR_ARG;
  }
  cudaStream_t st = (cudaStream_t)stream;
  (void)st;
#ifdef SB_EMUL_BUILD
  SBB_LAUNCH(k_csr_residual_dd, 2u, 32, 0, st, n, d_rowptr, d_colidx, d_vals, d_x, d_b, d_r);
#else
  const unsigned gx = (unsigned)std::min<int64_t>((n + 255) / 256, 148 * 16);
  SBB_LAUNCH(k_csr_residual_dd, gx ? gx : 1u, 256, 0, st, n, d_rowptr, d_colidx, d_vals, d_x, d_b, d_r);
#endif
  BTRY(cudaGetLastError());
  return SB_OK;
}
