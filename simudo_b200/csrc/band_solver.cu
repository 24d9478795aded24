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
                            double* ab, int32_t ldab, int32_t kv, double* work) {
  const int64_t sys = blockIdx.y;
  const double* v = vals + sys * vstride;
  double* A = ab + sys * n * ldab;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n;
       r += (int64_t)gridDim.x * blockDim.x) {
    const double s = rs[sys * n + r];
    const int64_t pr = perm[r];
    for (int64_t k = rowptr[r]; k < rowptr[r + 1]; ++k) {
      const int32_t c = colidx[k];
      const int64_t pc = perm[c];
      A[pc * ldab + kv + pr - pc] = v[k] * s * cs[sys * n + c];
    }
    work[sys * n + pr] = b[sys * bstride + r] * s;
  }
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
  if (rc != SB_OK) { sb_band_destroy(s); return rc; }
  cudaMemcpy(s->rowptr, h_rowptr, (n + 1) * sizeof(int64_t), cudaMemcpyHostToDevice);
  cudaMemcpy(s->colidx, h_colidx, s->nnz * sizeof(int32_t), cudaMemcpyHostToDevice);
  cudaMemcpy(s->perm, h_perm, n * sizeof(int32_t), cudaMemcpyHostToDevice);
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
extern "C" int sb_band_solve(sb_band_solver* s, int32_t nsys, const double* d_vals,
                             int64_t vstride, const double* d_b, double* d_x,
                             int32_t* h_info, void* stream) {
  if (!s || !d_vals || !d_b || !d_x || nsys <= 0 || nsys > s->batch) {
    snprintf(g_sb_err, sizeof g_sb_err, "bad argument to sb_band_solve");
    return SB_ERR_ARG;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t n = s->n;
  const unsigned gx = (unsigned)((n + 255) / 256 < 1184 ? (n + 255) / 256 : 1184);
#ifdef SB_EMUL_BUILD
  for (int sys = 0; sys < nsys; ++sys) {
    /* the emulator has a 1D grid: run the batch entries one after another */
    (void)gx;
  }
#endif
  BTRY(cudaMemsetAsync(s->ab, 0, (size_t)s->ldab * n * nsys * sizeof(double), st));
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
             vstride, d_b, n, s->rs, s->cs, s->ab, s->ldab, s->kl + s->ku, s->work);
  SBB_LAUNCH(k_band_factor_solve, (unsigned)nsys, SBB_THREADS, 0, st, n, s->kl, s->ku, s->ldab,
             s->ab, s->work, s->ipiv, s->info);
  SBB_LAUNCH(k_band_unpermute, GRID2(gx, nsys), 256, 0, st, n, s->perm, s->cs, s->work, d_x, n);
  for (int it = 0; it < s->n_refine; ++it) {
    SBB_LAUNCH(k_band_residual_dd, GRID2(gx, nsys), 256, 0, st, n, s->rowptr, s->colidx, s->perm,
               d_vals, vstride, d_b, d_x, n, s->rs, s->work);
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
             n, s->rs, s->cs, s->ab, s->ldab, s->kl + s->ku, s->work);
  SBB_LAUNCH(k_band_factor_solve, 1u, SBB_THREADS, 0, st, n, s->kl, s->ku, s->ldab, s->ab,
             s->work, s->ipiv, s->info);
  SBB_LAUNCH(k_band_unpermute, 4u, 64, 0, st, n, s->perm, s->cs, s->work, d_x, n);
  for (int it = 0; it < s->n_refine; ++it) {
    SBB_LAUNCH(k_band_residual_dd, 4u, 64, 0, st, n, s->rowptr, s->colidx, s->perm, d_vals,
               vstride, d_b, d_x, n, s->rs, s->work);
    SBB_LAUNCH(k_band_resolve, 1u, SBB_THREADS, 0, st, n, s->kl, s->ku, s->ldab, s->ab, s->work,
               s->ipiv);
    SBB_LAUNCH(k_band_accumulate, 4u, 64, 0, st, n, s->perm, s->cs, s->work, d_x, n);
  }
#endif
  BTRY(cudaGetLastError());
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
