/*
 * pdd_kernels.cu -- sm_100a kernels + C ABI (include/simudo_b200.h).
 *
 * k_assemble: one thread = one cell; coefficient moments in registers,
 * expansion tensors as warp-uniform constant-bank operands, direct scatter
 * into the fixed CSR (see pdd_device.cuh for the math and the citations).
 * No tensor cores: the path is FP64 gather / pointwise transcendental /
 * scatter (BASELINE.json north_star).
 */
#ifdef SB_EMUL_BUILD
#include "cuda_emul.h"      /* tests/emul: CPU shim, test infrastructure only */
#define SB_LAUNCH(kernel, grid, block, smem, stream, ...) \
  emul::launch((grid), (block), (smem), [&] { kernel(__VA_ARGS__); })
#else
#include <cuda_runtime.h>
#define SB_LAUNCH(kernel, grid, block, smem, stream, ...) \
  kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#endif
#include <stdio.h>
#include <string.h>
#include <vector>
#include <algorithm>
#include <atomic>
#include <chrono>
#include <mutex>
#include <type_traits>

#define SB_TILE 128                /* threads (= cells) per CTA */
#include "pdd_device.cuh"

#define SB_MAX_SMEM_PATTERNS 16

__constant__ SbTables c_T;
__constant__ sb_model c_M;

thread_local char g_sb_err[512] = "";   /* shared with band_solver.cu */
#define g_err g_sb_err
static std::atomic<long long> g_launches{0};
/* which plan's tables sit in the __constant__ copies, per device (one copy per device) */
#define SB_MAX_DEVICES 64
static const void* g_tables_owner[SB_MAX_DEVICES] = {nullptr};
static int current_device_slot() {
#ifdef SB_EMUL_BUILD
  return 0;
#else
  int dev = 0;
  cudaGetDevice(&dev);
  return (dev >= 0 && dev < SB_MAX_DEVICES) ? dev : 0;
#endif
}

#define CUDA_TRY(x)                                                          \
  do {                                                                       \
    cudaError_t e_ = (x);                                                    \
    if (e_ != cudaSuccess) {                                                 \
      snprintf(g_err, sizeof g_err, "%s failed: %s (%s:%d)", #x,            \
               cudaGetErrorString(e_), __FILE__, __LINE__);                  \
      return SB_ERR_CUDA;                                                    \
    }                                                                        \
  } while (0)

struct PlanDev {
  int64_t n_cells, N, nnz;
  int32_t P, L, nb, nb_dof;
  const double* cell_vertices;
  const int32_t* cell_tag;
  const int32_t* cell_dofs;
  const int32_t* cell_neighbors;
  const uint8_t* cell_jump_mask;
  const int64_t* rowptr;
  const uint8_t* patterns;       /* [n_patterns][L][L]                       */
  const int32_t* cell_pattern;   /* [n_cells]                                */
  int64_t n_patterns;
  int stage_patk;                /* rank tables fit the shared-memory staging area */
  int owned_bulk;                /* DG / interior rows have the structural lengths: bulk copy */
  const uint8_t* patk;           /* [n_patterns][L][32] packed emission-order ranks */
  const int32_t* sp_cells;       /* [n_special] cells with Dirichlet / natural-BC rows */
  const int64_t* rowbase;        /* [5P][n_cells] vals offset of the first row of each
                                    (sub-problem, entity) group: DG, facet 0..2, interior */
  const uint16_t* rowlen;        /* [5P][n_cells] length of the rows of that group     */
  int64_t n_special;
  const double* tag_params;
  const int32_t* cell_special;   /* [n_cells] -1 or special index            */
  const uint64_t* sp_mask_lo;
  const uint32_t* sp_mask_hi;
  const int32_t* sp_bcidx;       /* [n_special][L] index into bc_values | -1 */
  const int32_t* sp_nat_begin;   /* [n_special+1]                            */
  const int32_t* natbc_row;      /* sorted by cell                           */
  const SbTables* tables;        /* global copy for lane-varying indexing    */
  double* scratch;               /* [(nb*44 + nb_dof*SB_NG)][n_cells] moments  */
  /* closed-form layout (native.inl): 'b200' numbering + 'native' pattern      */
  int native;
  const int32_t* nat_fid;        /* [n_cells][3] facet ids                    */
  const int32_t* nat_finfo;      /* [n_cells] per local facet f, bits 4f..: 1 boundary,
                                    2 this cell is the facet's second cell, (4|8) the
                                    facet's local number inside the neighbour; bits
                                    12 + 4f + k: base_w jump of band k on facet f ('+' side);
                                    bits 24 + k: cell inside the subdomain of band k */
  const int64_t* nat_fbase;      /* [P][n_facets] vals index of the facet's 3 rows of sub-problem p */
  int64_t nat_dofF0;             /* first facet dof = 6 P n_cells             */
  int64_t nat_nf;                /* number of facets                          */
  int* nat_tile_ctr;             /* [8] tile counters of the persistent row kernels, one per
                                    sub-problem launch, zeroed once per pass     */
};

struct sb_plan {
  PlanDev d;
  sb_model model;
  SbTables tables;
  std::vector<void*> allocs;
  int64_t n_bc, n_natbc, n_special;
  double* tmp_cells;             /* [n_cells] scratch for rebase             */
#ifndef SB_EMUL_BUILD
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;   /* around the assembly pass when profiling */
  cudaEvent_t evs[3] = {nullptr, nullptr, nullptr};  /* after moments, generation, fast rows */
#endif
  int profiling = 0;
#ifndef SB_EMUL_BUILD
  /* the boundary-cell kernel runs beside the row kernels on a stream of its own */
  cudaStream_t side = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
#endif
  int gen_sig = 0;               /* 0: interpret the process list; 1: SbSigIB5; 2: SbSigPnSRH */
};

struct StateDev {
  const double* u;
  const double* base_w;
  const double* thmeq;
  const double* phi_opt;
  const double* bc_values;
  const double* natbc_values;
};

/* ------------------------------------------------------------------------- */
/* zero the slots that two cells add into: residual entries of facet dofs and
 * the (facet dof, facet dof) Jacobian entries                                */
__global__ void k_zero_shared(PlanDev pl, double* vals, double* b) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= pl.n_cells) return;
  const int L = pl.L;
  const int32_t* dofs = pl.cell_dofs + c * L;
  const uint8_t* pat = pl.patterns + (int64_t)pl.cell_pattern[c] * L * L;
  for (int p = 0; p < pl.P; ++p)
    for (int f = 0; f < 3; ++f)
      for (int k = 0; k < 3; ++k) {
        const int li = sb_lbdm(p, 3 * f + k);
        const int64_t row = dofs[li];
        if (b) b[row] = 0.0;
        if (vals) {
          const int64_t r0 = pl.rowptr[row];
          for (int k2 = 0; k2 < 3; ++k2) {
            const uint8_t rk = pat[li * L + sb_lbdm(p, 3 * f + k2)];
            if (rk != 0xFF) vals[r0 + rk] = 0.0;
          }
        }
      }
}

/* ------------------------------------------------------------------------- */
/* The assembly is three launches with small code footprints and independent
 * occupancy instead of one fused kernel (the fused variant spent a third of
 * its cycles waiting for instructions, profiles/README.md):
 *   k_moments     thread = (cell, band):   phase A, coefficient moments
 *   k_generation  thread = cell:           phase B, generation moments
 *   k_rows        thread = (cell, sub-problem): phase C, entries + scatter
 * Moments travel through an L2-friendly scratch array laid out [moment][cell]
 * so that both the producer and the consumer access it coalesced.            */
#define SB_NMOMA (SB_NMC + SB_NQ + SB_NRX + SB_NR)     /* 44 per (cell, band) */

__device__ __forceinline__ double* momA_ptr(const PlanDev& pl, int k, int64_t c) {
  return pl.scratch + (int64_t)k * SB_NMOMA * pl.n_cells + c;
}
__device__ __forceinline__ double* momG_ptr(const PlanDev& pl, int k, int64_t c) {
  return pl.scratch + ((int64_t)pl.nb * SB_NMOMA + (int64_t)k * SB_NG) * pl.n_cells + c;
}

#ifndef SB_MB_MOM
#define SB_MB_MOM 3
#endif
#ifndef SB_MB_GEN
#define SB_MB_GEN 2
#endif
#ifndef SB_MB_ROWS
#define SB_MB_ROWS 2
#endif
#include "native.inl"

/* zero the (facet dof, facet dof) Jacobian slots of sub-problem p on the facets this
 * cell owns (boundary facet, or lower cell index of the pair): both cells of a facet
 * atomically add into them later in the stream */
__device__ __forceinline__ void zero_shared_slots(const PlanDev& pl, int64_t c, int p, double* vals) {
  const int L = pl.L;
  const int64_t nc = pl.n_cells;
  const uint8_t* patk = pl.patk + (int64_t)pl.cell_pattern[c] * L * 32;
#pragma unroll
  for (int f = 0; f < 3; ++f) {
    const int64_t cn = pl.cell_neighbors[c * 3 + f];
    if (cn >= 0 && cn < c) continue;
    /* the three rows of a facet are adjacent: base + kk * length (coalesced tables) */
    const int64_t rb = pl.rowbase[(int64_t)(5 * p + 1 + f) * nc + c];
    const int64_t rl = pl.rowlen[(int64_t)(5 * p + 1 + f) * nc + c];
#pragma unroll
    for (int kk = 0; kk < 3; ++kk) {
      const int li = sb_lbdm(p, 3 * f + kk);
      const uint32_t* q = reinterpret_cast<const uint32_t*>(patk + li * 32);
#pragma unroll
      for (int k2 = 0; k2 < 3; ++k2) {
        const int e = 3 * f + k2;
        const uint32_t rk = (q[e >> 2] >> (8 * (e & 3))) & 0xFFu;
        if (rk != 0xFFu) vals[rb + kk * rl + rk] = 0.0;
      }
    }
  }
}

__global__ void __launch_bounds__(SB_TILE, SB_MB_MOM)
k_moments(PlanDev pl, StateDev st, int nblk, double* vals) {
  const int k = blockIdx.x / nblk;
  const int64_t c = (int64_t)(blockIdx.x % nblk) * SB_TILE + threadIdx.x;
  if (c >= pl.n_cells) return;
  const int L = pl.L;
  const bool has_dofs = k < pl.nb_dof;
  const SbCellGeom g = sb_geom(pl.cell_vertices + c * 6);
  const double* tp = pl.tag_params + (int64_t)pl.cell_tag[c] * SB_TAG_STRIDE;
  const int32_t* dofs = pl.cell_dofs + c * L;
  const SbBandK K = sb_band_consts(c_M, tp, k);
  const bool inside = tp[SB_TP_BAND0 + 5 * k + SB_TPB_IN_SUBDOMAIN] != 0.0;
  const bool dd = inside && has_dofs;
  double phi[3], dw[3] = {0.0, 0.0, 0.0}, w0 = 0.0;
#pragma unroll
  for (int i = 0; i < 3; ++i) phi[i] = st.u[dofs[i]];
  double gc[2][6];
  if (has_dofs) {
    w0 = st.base_w[(int64_t)k * pl.n_cells + c];
#pragma unroll
    for (int i = 0; i < 3; ++i) dw[i] = st.u[dofs[sb_ldg(1 + k, i)]];
  }
  if (dd) {
    double jc[2][6];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int r = 0; r < 6; ++r) jc[a][r] = 0.0;
#pragma unroll
    for (int m = 0; m < 12; ++m) {
      const double fm = st.u[dofs[sb_lbdm(1 + k, m)]];
#pragma unroll
      for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int r = 0; r < 6; ++r) jc[a][r] += fm * c_T.bdmC[m][a][r];
    }
#pragma unroll
    for (int r = 0; r < 6; ++r) {
      gc[0][r] = g.G11 * jc[0][r] + g.G12 * jc[1][r];
      gc[1][r] = g.G12 * jc[0][r] + g.G22 * jc[1][r];
    }
  }
  const double chi0 = phi[0] + dw[0] + w0;
  const double chix = (phi[1] - phi[0]) + (dw[1] - dw[0]);
  const double chiy = (phi[2] - phi[0]) + (dw[2] - dw[0]);
  double mC[SB_NMC], Q[SB_NQ], mRx[SB_NRX], mR[SB_NR] = {0.0, 0.0, 0.0};
#pragma unroll
  for (int t = 0; t < SB_NMC; ++t) mC[t] = 0.0;
#pragma unroll
  for (int t = 0; t < SB_NQ; ++t) Q[t] = 0.0;
#pragma unroll
  for (int t = 0; t < SB_NRX; ++t) mRx[t] = 0.0;
  const bool rho = (K.N != 0.0);
  const bool same_rule = (c_M.quad_m_rho == c_M.quad_m_super);
  if (same_rule && c_M.quad_m_super == 11) {
    /* common cases as straight-line instances (see sb_phaseA_band) */
#ifndef SB_MOM_Q_REGS
    /* the 20 Q accumulators are touched once per rule row: in shared memory they free 40
       registers for the interleaved exp chains of the row (fewer spills, 2.65 -> 2.50 ms) */
    __shared__ double s_q[SB_NQ][SB_TILE];
    double* qs = &s_q[0][threadIdx.x];
#else
    double* qs = nullptr;
#endif
    bool fastm = false;
    if (rho && dd && K.kind == SB_BAND_NONDEGENERATE) {
      sb_phaseA_band<11, 1>(c_T.sup, rho, dd, K, chi0, chix, chiy, gc, mC, Q, mR, mRx, qs);
      fastm = true;
    } else if (rho && dd && K.const_mob) {
      sb_phaseA_band<11, 2>(c_T.sup, rho, dd, K, chi0, chix, chiy, gc, mC, Q, mR, mRx, qs);
      fastm = true;
    }
    if (fastm) {
      if (qs) {
#pragma unroll
        for (int t = 0; t < SB_NQ; ++t) Q[t] = qs[t * SB_TILE];
      }
    }
    else if (rho || dd) sb_phaseA_band<11>(c_T.sup, rho, dd, K, chi0, chix, chiy, gc, mC, Q, mR, mRx);
  } else if (same_rule) {
    if (rho || dd) sb_phaseA_band<0>(c_T.sup, rho, dd, K, chi0, chix, chiy, gc, mC, Q, mR, mRx);
  } else {
    if (rho) sb_phaseA_band<0>(c_T.rho, true, false, K, chi0, chix, chiy, gc, mC, Q, mR, mRx);
    if (dd) sb_phaseA_band<0>(c_T.sup, false, true, K, chi0, chix, chiy, gc, mC, Q, mR, mRx);
  }
  double* o = momA_ptr(pl, k, c);
  const int64_t nc = pl.n_cells;
  if (dd) {
#pragma unroll
    for (int t = 0; t < SB_NMC; ++t) o[(int64_t)t * nc] = mC[t];
#pragma unroll
    for (int t = 0; t < SB_NQ; ++t) o[(int64_t)(SB_NMC + t) * nc] = Q[t];
  }
#pragma unroll
  for (int t = 0; t < SB_NRX; ++t) o[(int64_t)(SB_NMC + SB_NQ + t) * nc] = mRx[t];
#pragma unroll
  for (int t = 0; t < SB_NR; ++t) o[(int64_t)(SB_NMC + SB_NQ + SB_NRX + t) * nc] = mR[t];
  /* last, so that no load of the moment computation queues behind these stores */
  if (vals) {
    if (k < pl.nb_dof) zero_shared_slots(pl, c, 1 + k, vals);
    if (k == 0) zero_shared_slots(pl, c, 0, vals);
  }
}

template <int NB>
__global__ void __launch_bounds__(SB_TILE, SB_MB_GEN)
k_generation(PlanDev pl, StateDev st) {
  const int64_t c = (int64_t)blockIdx.x * SB_TILE + threadIdx.x;
  if (c >= pl.n_cells) return;
  const int L = pl.L;
  const double* tp = pl.tag_params + (int64_t)pl.cell_tag[c] * SB_TAG_STRIDE;
  const int32_t* dofs = pl.cell_dofs + c * L;
  double dgv[1 + SB_MAX_BANDS][3];
  double w0[SB_MAX_BANDS];
  bool inside_k[SB_MAX_BANDS];
#pragma unroll
  for (int i = 0; i < 3; ++i) dgv[0][i] = st.u[dofs[i]];
#pragma unroll
  for (int k = 0; k < SB_MAX_BANDS; ++k) {
    w0[k] = 0.0; inside_k[k] = false;
#pragma unroll
    for (int i = 0; i < 3; ++i) dgv[1 + k][i] = 0.0;
    if (k < NB) {
      w0[k] = st.base_w[(int64_t)k * pl.n_cells + c];
      inside_k[k] = tp[SB_TP_BAND0 + 5 * k + SB_TPB_IN_SUBDOMAIN] != 0.0;
#pragma unroll
      for (int i = 0; i < 3; ++i) dgv[1 + k][i] = st.u[dofs[sb_ldg(1 + k, i)]];
    }
  }
  double mG[NB > 0 ? NB : 1][SB_NG];
  sb_phaseB_all<NB>(c_T, c_M, tp, dgv, w0, inside_k, st.thmeq ? st.thmeq + c * 6 : nullptr,
                    st.phi_opt ? st.phi_opt + c * 6 : nullptr, pl.n_cells * 6, mG);
  const int64_t nc = pl.n_cells;
#pragma unroll
  for (int k = 0; k < NB; ++k) {
    double* o = momG_ptr(pl, k, c);
#pragma unroll
    for (int t = 0; t < 9 + 6 * NB; ++t) o[(int64_t)t * nc] = mG[k][t];
  }
}

/* boundary layer: cells with Dirichlet dofs or natural-BC rows (generic path) */
template <int NB>
__global__ void __launch_bounds__(SB_TILE)
k_rows_special(PlanDev pl, StateDev st, double* vals, double* b, int nblk) {
  const int p = blockIdx.x / nblk;
  const int64_t is = (int64_t)(blockIdx.x % nblk) * SB_TILE + threadIdx.x;
  if (is >= pl.n_special) return;
  const int64_t c = pl.sp_cells[is];
  const int L = pl.L;
  const int64_t nc = pl.n_cells;
  SbIO io;
  io.L = L;
  io.dofs = pl.cell_dofs + c * L;
  io.pat = pl.patterns + (int64_t)pl.cell_pattern[c] * L * L;
  io.rowptr = pl.rowptr;
  io.u = st.u; io.bc_values = st.bc_values;
  io.vals = vals; io.b = b;
  io.mlo = 0; io.mhi = 0; io.bcidx = nullptr;
  io.nat = pl.native; io.finfo = 0; io.snb = nullptr;
  double snb[18];
  if (pl.native) {
    /* the neighbours' (facet, facet) blocks of this sub-problem, for the facets this
       cell writes (native.inl) */
    io.finfo = pl.nat_finfo[c];
    for (int f = 0; f < 3; ++f) {
      const bool mine = ((io.finfo >> (4 * f)) & 3) == 0;
      const int64_t cn = mine ? pl.cell_neighbors[c * 3 + f] : c;
      const int f2 = (io.finfo >> (4 * f + 2)) & 3;
      const double* sn = (p == 0) ? natS0_ptr(pl, cn) : natA_ptr(pl, p - 1, cn) + (int64_t)SB_NAT_S * SB_NQS;
      for (int t = 0; t < 6; ++t) snb[6 * f + t] = mine ? sn[(int64_t)(6 * f2 + t) * SB_NQS] : 0.0;
    }
    io.snb = snb;
  }
  int nb0 = 0, nb1 = 0;
  const int32_t spc = pl.cell_special[c];
  if (spc >= 0) {
    io.mlo = pl.sp_mask_lo[spc]; io.mhi = pl.sp_mask_hi[spc];
    io.bcidx = pl.sp_bcidx + (int64_t)spc * L;
    nb0 = pl.sp_nat_begin[spc]; nb1 = pl.sp_nat_begin[spc + 1];
  }
  const bool sp = (io.mlo | io.mhi) != 0;
  const SbCellGeom g = sb_geom(pl.cell_vertices + c * 6);
  const double* tp = pl.tag_params + (int64_t)pl.cell_tag[c] * SB_TAG_STRIDE;
  const int32_t* dofs = io.dofs;
  double fl[12], sl[3];
#pragma unroll
  for (int m = 0; m < 12; ++m) fl[m] = st.u[dofs[sb_lbdm(p, m)]];
#pragma unroll
  for (int i = 0; i < 3; ++i) sl[i] = st.u[dofs[sb_ldg(p, i)]];

  if (p == 0) {
    /* Poisson rows: rho moments summed over the bands */
    double mR[SB_NR] = {0.0, 0.0, 0.0};
    double mX[1 + SB_MAX_BANDS][SB_NRX];
#pragma unroll
    for (int q = 0; q < 1 + SB_MAX_BANDS; ++q)
#pragma unroll
      for (int t = 0; t < SB_NRX; ++t) mX[q][t] = 0.0;
    for (int k = 0; k < pl.nb; ++k) {
      const int64_t qs = pl.native ? SB_NQS : nc;           /* stride between quantities */
      const double* o = pl.native ? natA_ptr(pl, k, c) + (int64_t)SB_NAT_RX * qs
                                  : momA_ptr(pl, k, c) + (int64_t)(SB_NMC + SB_NQ) * qs;
#pragma unroll
      for (int t = 0; t < SB_NRX; ++t) {
        const double v = o[(int64_t)t * qs];
        mX[0][t] += v;
        if (k < NB) mX[1 + k][t] = v;
      }
#pragma unroll
      for (int t = 0; t < SB_NR; ++t) mR[t] += o[(int64_t)(SB_NRX + t) * qs];
    }
    /* static charge: constant => only the Psi_0 = sqrt(2) moment, int Psi_0 = 1/sqrt(2) */
    mR[0] += tp[SB_TP_RHO_STATIC] / tp[SB_TP_EPS] * 0.70710678118654752440;
    if (sp) {
      sb_rows_scalar<true, 1 + NB>(c_T, c_M, io, g, 0, true, fl, sl, -g.adet, mR, &mX[0][0],
                                   pl.natbc_row, st.natbc_values, nb0, nb1);
      sb_rows_flux<true>(c_T, c_M, io, g, 0, 0, nullptr, nullptr, fl, sl, 0.0, 0.0, 0.0,
                         pl.natbc_row, st.natbc_values, nb0, nb1);
    } else {
      sb_rows_scalar<false, 1 + NB>(c_T, c_M, io, g, 0, true, fl, sl, -g.adet, mR, &mX[0][0],
                                    pl.natbc_row, st.natbc_values, nb0, nb1);
      sb_rows_flux<false>(c_T, c_M, io, g, 0, 0, nullptr, nullptr, fl, sl, 0.0, 0.0, 0.0,
                          pl.natbc_row, st.natbc_values, nb0, nb1);
    }
    return;
  }

  const int k = p - 1;
  const bool inside = tp[SB_TP_BAND0 + 5 * k + SB_TPB_IN_SUBDOMAIN] != 0.0;
  {
    double mC[21], Q[36];                          /* native: W (21 values), d[3 i + l] (36) */
    if (inside && pl.native) {
      const double* o = natA_ptr(pl, k, c);
#pragma unroll
      for (int t = 0; t < 21; ++t) mC[t] = o[(int64_t)(SB_NAT_W + t) * SB_NQS];
#pragma unroll
      for (int t = 0; t < 36; ++t) Q[t] = o[(int64_t)(SB_NAT_D + t) * SB_NQS];
    } else if (inside) {
      const double* o = momA_ptr(pl, k, c);
#pragma unroll
      for (int t = 0; t < SB_NMC; ++t) mC[t] = o[(int64_t)t * nc];
#pragma unroll
      for (int t = 0; t < SB_NQ; ++t) Q[t] = o[(int64_t)(SB_NMC + t) * nc];
    }
    /* base_w jump term on interior facets where this cell is the '+' side */
    double jumpf[3];
    const double w0 = st.base_w[(int64_t)k * nc + c];
#pragma unroll
    for (int f = 0; f < 3; ++f) {
      jumpf[f] = 0.0;
      if ((pl.cell_jump_mask[c * 3 + f] >> k) & 1) {
        const int64_t cn = pl.cell_neighbors[c * 3 + f];
        const double ref_out = (f == 1) ? -1.0 : 1.0;
        jumpf[f] = (st.base_w[(int64_t)k * nc + cn] - w0) * ref_out * g.sdet;
      }
    }
    const int mode = inside ? 1 : 2;
    if (sp)
      sb_rows_flux<true>(c_T, c_M, io, g, p, mode, mC, Q, fl, sl, jumpf[0], jumpf[1], jumpf[2],
                         pl.natbc_row, st.natbc_values, nb0, nb1, pl.native != 0);
    else
      sb_rows_flux<false>(c_T, c_M, io, g, p, mode, mC, Q, fl, sl, jumpf[0], jumpf[1], jumpf[2],
                          pl.natbc_row, st.natbc_values, nb0, nb1, pl.native != 0);
  }
  {
    double mG[SB_NG];
    if (inside) {
      const double* o = pl.native ? natG_ptr(pl, k, c) : momG_ptr(pl, k, c);
      const int64_t qs = pl.native ? SB_NQS : nc;
#pragma unroll
      for (int t = 0; t < 9 + 6 * NB; ++t) mG[t] = o[(int64_t)t * qs];
    }
    const double coef = -(double)c_M.band_sign[k] * c_M.e_charge * g.adet;
    if (sp)
      sb_rows_scalar<true, 1 + NB>(c_T, c_M, io, g, p, inside, fl, sl, coef, mG, mG + 3,
                                   pl.natbc_row, st.natbc_values, nb0, nb1);
    else
      sb_rows_scalar<false, 1 + NB>(c_T, c_M, io, g, p, inside, fl, sl, coef, mG, mG + 3,
                                    pl.natbc_row, st.natbc_values, nb0, nb1);
  }
}

/* interior cells (no Dirichlet dofs, no natural-BC rows): compile-time
 * sub-problem, packed ranks, no boundary logic */
template <int NB, int PS>
__global__ void __launch_bounds__(SB_TILE, SB_MB_ROWS)
k_rows_fast(PlanDev pl, StateDev st, double* vals, double* b) {
  /* packed ranks of this sub-problem's 15 rows for every pattern, staged in shared
     memory (a product mesh has 8 patterns -> 3.8 KB) */
  __align__(16) __shared__ uint8_t s_patk[SB_MAX_SMEM_PATTERNS][15][32];
  const int L = pl.L;
  const bool stage = pl.stage_patk != 0;
  if (stage) {
    const int nw = (int)pl.n_patterns * 15 * 8;            /* 32-bit words */
    for (int w = threadIdx.x; w < nw; w += SB_TILE) {
      const int q = w / 120, rem = w % 120, r = rem / 8, e = rem % 8;
      reinterpret_cast<uint32_t*>(&s_patk[q][r][0])[e] =
          reinterpret_cast<const uint32_t*>(pl.patk + ((int64_t)q * L + 15 * PS + r) * 32)[e];
    }
  }
  __syncthreads();
  const int64_t c = (int64_t)blockIdx.x * SB_TILE + threadIdx.x;
  if (c >= pl.n_cells) return;
  if (pl.cell_special[c] >= 0) return;           /* handled by k_rows_special */
  const int64_t nc = pl.n_cells;
  SbFastIO io;
  io.dofs = pl.cell_dofs + c * L;
  /* io.patk is indexed with the cell-local row li = 15 PS + r */
  io.patk = stage ? (&s_patk[pl.cell_pattern[c]][0][0] - 15 * PS * 32)
                  : (pl.patk + (int64_t)pl.cell_pattern[c] * L * 32);
  io.vals = vals; io.b = b;
#pragma unroll
  for (int e = 0; e < 5; ++e) {
    io.rb[e] = pl.rowbase[(int64_t)(5 * PS + e) * nc + c];
    io.rl[e] = pl.rowlen[(int64_t)(5 * PS + e) * nc + c];
    io.d0[e] = io.dofs[15 * PS + (e == 0 ? 0 : 3 * e)];
  }
  const SbCellGeom g = sb_geom(pl.cell_vertices + c * 6);
  const double* tp = pl.tag_params + (int64_t)pl.cell_tag[c] * SB_TAG_STRIDE;
  const int32_t* dofs = io.dofs;
  double fl[12], sl[3];
#pragma unroll
  for (int m = 0; m < 12; ++m) fl[m] = st.u[dofs[sb_lbdm(PS, m)]];
#pragma unroll
  for (int i = 0; i < 3; ++i) sl[i] = st.u[dofs[sb_ldg(PS, i)]];
  if (PS == 0) {
    double mR[SB_NR] = {0.0, 0.0, 0.0};
    double mX[1 + SB_MAX_BANDS][SB_NRX];
#pragma unroll
    for (int q = 0; q < 1 + SB_MAX_BANDS; ++q)
#pragma unroll
      for (int t = 0; t < SB_NRX; ++t) mX[q][t] = 0.0;
#pragma unroll
    for (int k = 0; k < SB_MAX_BANDS; ++k) {
      if (k >= pl.nb) break;
      const double* o = momA_ptr(pl, k, c);
#pragma unroll
      for (int t = 0; t < SB_NRX; ++t) {
        const double v = o[(int64_t)(SB_NMC + SB_NQ + t) * nc];
        mX[0][t] += v;
        if (k < NB) mX[1 + k][t] = v;
      }
#pragma unroll
      for (int t = 0; t < SB_NR; ++t) mR[t] += o[(int64_t)(SB_NMC + SB_NQ + SB_NRX + t) * nc];
    }
    mR[0] += tp[SB_TP_RHO_STATIC] / tp[SB_TP_EPS] * 0.70710678118654752440;
    sb_rows_scalar_fast<0, 1 + NB>(c_T, c_M, io, g, true, fl, sl, -g.adet, mR, &mX[0][0]);
    sb_rows_flux_fast<0, 0>(c_T, c_M, io, g, nullptr, nullptr, fl, sl, 0.0, 0.0, 0.0);
  } else {
    constexpr int k = PS > 0 ? PS - 1 : 0;
    const bool inside = tp[SB_TP_BAND0 + 5 * k + SB_TPB_IN_SUBDOMAIN] != 0.0;
    if (inside) {
      double mC[SB_NMC], Q[SB_NQ];
      const double* o = momA_ptr(pl, k, c);
#pragma unroll
      for (int t = 0; t < SB_NMC; ++t) mC[t] = o[(int64_t)t * nc];
#pragma unroll
      for (int t = 0; t < SB_NQ; ++t) Q[t] = o[(int64_t)(SB_NMC + t) * nc];
      double jumpf[3];
      const double w0 = st.base_w[(int64_t)k * nc + c];
#pragma unroll
      for (int f = 0; f < 3; ++f) {
        jumpf[f] = 0.0;
        if ((pl.cell_jump_mask[c * 3 + f] >> k) & 1) {
          const int64_t cn = pl.cell_neighbors[c * 3 + f];
          const double ref_out = (f == 1) ? -1.0 : 1.0;
          jumpf[f] = (st.base_w[(int64_t)k * nc + cn] - w0) * ref_out * g.sdet;
        }
      }
      sb_rows_flux_fast<PS, 1>(c_T, c_M, io, g, mC, Q, fl, sl, jumpf[0], jumpf[1], jumpf[2]);
      double mG[SB_NG];
      const double* og = momG_ptr(pl, k, c);
#pragma unroll
      for (int t = 0; t < 9 + 6 * NB; ++t) mG[t] = og[(int64_t)t * nc];
      const double coef = -(double)c_M.band_sign[k] * c_M.e_charge * g.adet;
      sb_rows_scalar_fast<PS, 1 + NB>(c_T, c_M, io, g, true, fl, sl, coef, mG, mG + 3);
    } else {
      sb_rows_flux_fast<PS, 2>(c_T, c_M, io, g, nullptr, nullptr, fl, sl, 0.0, 0.0, 0.0);
      sb_rows_scalar_fast<PS, 1 + NB>(c_T, c_M, io, g, false, fl, sl, 0.0, nullptr, nullptr);
    }
  }
}

/* Staged variant of k_rows_fast (pdd_device.cuh, "Staged fast path"): threads stage
 * one entity group at a time in shared memory, the warp flushes it with lane = entry.
 * Needs the rank tables in shared memory (pl.stage_patk); no thread leaves before the
 * last flush.                                                                     */
template <int NB, int PS>
__global__ void __launch_bounds__(SB_TILE, SB_MB_ROWS)
k_rows_staged(PlanDev pl, StateDev st, double* vals, double* b) {
  /* dynamic shared memory (95 KB > the 48 KB static limit): staging | meta | patk */
#ifdef SB_EMUL_BUILD
  unsigned char* s_dyn = emul::dyn_smem();
#else
  extern __shared__ __align__(16) unsigned char s_dyn[];
#endif
  typedef double StageRow[32 * SB_STAGE_STRIDE];
  typedef uint8_t PatkTab[15][32];
  StageRow* s_stage = reinterpret_cast<StageRow*>(s_dyn);
  SbWarpMeta* s_meta = reinterpret_cast<SbWarpMeta*>(s_dyn + sizeof(StageRow) * (SB_TILE / 32));
  PatkTab* s_patk = reinterpret_cast<PatkTab*>(s_dyn + (sizeof(StageRow) + sizeof(SbWarpMeta))
                                                * (SB_TILE / 32));
  const int L = pl.L;
  {
    const int nw = (int)pl.n_patterns * 15 * 8;            /* 32-bit words */
    for (int w = threadIdx.x; w < nw; w += SB_TILE) {
      const int q = w / 120, rem = w % 120, r = rem / 8, e = rem % 8;
      reinterpret_cast<uint32_t*>(&s_patk[q][r][0])[e] =
          reinterpret_cast<const uint32_t*>(pl.patk + ((int64_t)q * L + 15 * PS + r) * 32)[e];
    }
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t nc = pl.n_cells;
  const int64_t c = (int64_t)blockIdx.x * SB_TILE + threadIdx.x;
  const bool active = (c < nc) && (pl.cell_special[c < nc ? c : 0] < 0);
  const int64_t cc = active ? c : 0;                 /* inactive lanes read cell 0, emit nothing */
  SbWarpMeta& meta = s_meta[warp];
  double* stg = &s_stage[warp][lane * SB_STAGE_STRIDE];
  const int32_t* dofs = pl.cell_dofs + cc * L;
  const double* tp = pl.tag_params + (int64_t)pl.cell_tag[cc] * SB_TAG_STRIDE;
  constexpr int k = PS > 0 ? PS - 1 : 0;
  const bool inside = (PS == 0) ? true : (tp[SB_TP_BAND0 + 5 * k + SB_TPB_IN_SUBDOMAIN] != 0.0);
#pragma unroll
  for (int e = 0; e < 5; ++e) {
    meta.rb[e][lane] = pl.rowbase[(int64_t)(5 * PS + e) * nc + cc];
    meta.rl[e][lane] = pl.rowlen[(int64_t)(5 * PS + e) * nc + cc];
    meta.d0[e][lane] = dofs[15 * PS + (e == 0 ? 0 : 3 * e)];
  }
  meta.info[lane] = (active ? 1 : 0) | (inside ? 2 : 0) | ((int32_t)pl.cell_pattern[cc] << 8);
  /* owned rows: bulk copy needs the structural row lengths and a Jacobian to write */
  const bool bulk = pl.owned_bulk && vals != nullptr;
  const int64_t rb0 = meta.rb[0][lane], rb4 = meta.rb[4][lane];
  const int ph0 = (int)(rb0 & 1), ph4 = (int)(rb4 & 1);
  const uint8_t* pk_own = &s_patk[pl.cell_pattern[cc]][0][0];      /* rows 0..2 DG, 12..14 interior */
  const SbCellGeom g = sb_geom(pl.cell_vertices + cc * 6);
  double fl[12], sl[3];
#pragma unroll
  for (int m = 0; m < 12; ++m) fl[m] = st.u[dofs[sb_lbdm(PS, m)]];
#pragma unroll
  for (int i = 0; i < 3; ++i) sl[i] = st.u[dofs[sb_ldg(PS, i)]];
  double W[21];
  double V[2][6][3];
  double jumpf[3] = {0.0, 0.0, 0.0};
  /* ---- group 0: DG rows ---------------------------------------------------- */
  if (PS == 0) {
    double mR[SB_NR] = {0.0, 0.0, 0.0};
    double mX[1 + SB_MAX_BANDS][SB_NRX];
#pragma unroll
    for (int q = 0; q < 1 + SB_MAX_BANDS; ++q)
#pragma unroll
      for (int t = 0; t < SB_NRX; ++t) mX[q][t] = 0.0;
#pragma unroll
    for (int kk = 0; kk < SB_MAX_BANDS; ++kk) {
      if (kk >= pl.nb) break;
      const double* o = momA_ptr(pl, kk, cc);
#pragma unroll
      for (int t = 0; t < SB_NRX; ++t) {
        const double v = o[(int64_t)(SB_NMC + SB_NQ + t) * nc];
        mX[0][t] += v;
        if (kk < NB) mX[1 + kk][t] = v;
      }
#pragma unroll
      for (int t = 0; t < SB_NR; ++t) mR[t] += o[(int64_t)(SB_NMC + SB_NQ + SB_NRX + t) * nc];
    }
    mR[0] += tp[SB_TP_RHO_STATIC] / tp[SB_TP_EPS] * 0.70710678118654752440;
    if (active) {
      if (bulk)
        sb_stage_scalar_rows<0, 1 + NB>(c_T, c_M, g, true, fl, sl, -g.adet, mR, &mX[0][0],
                                        stg + ph0, pk_own, 12 + 3 * (1 + NB),
                                        b ? b + dofs[0] : nullptr);
      else
        sb_stage_scalar_rows<0, 1 + NB>(c_T, c_M, g, true, fl, sl, -g.adet, mR, &mX[0][0], stg);
    }
  } else {
    double mG[SB_NG];
    if (inside) {
      const double* og = momG_ptr(pl, k, cc);
#pragma unroll
      for (int t = 0; t < 9 + 6 * NB; ++t) mG[t] = og[(int64_t)t * nc];
    }
    const double coef = -(double)c_M.band_sign[k] * c_M.e_charge * g.adet;
    if (active) {
      if (bulk)
        sb_stage_scalar_rows<PS, 1 + NB>(c_T, c_M, g, inside, fl, sl, coef, mG, mG + 3,
                                         stg + ph0, pk_own, 12 + 3 * (1 + NB),
                                         b ? b + dofs[15 * PS] : nullptr);
      else
        sb_stage_scalar_rows<PS, 1 + NB>(c_T, c_M, g, inside, fl, sl, coef, mG, mG + 3, stg);
    }
    if (inside) {
      /* moments for the flux rows: loaded before the first flush */
      double mC[SB_NMC], Q[SB_NQ];
      const double* o = momA_ptr(pl, k, cc);
#pragma unroll
      for (int t = 0; t < SB_NMC; ++t) mC[t] = o[(int64_t)t * nc];
#pragma unroll
      for (int t = 0; t < SB_NQ; ++t) Q[t] = o[(int64_t)(SB_NMC + t) * nc];
      const double w0 = st.base_w[(int64_t)k * nc + cc];
#pragma unroll
      for (int f = 0; f < 3; ++f) {
        if ((pl.cell_jump_mask[cc * 3 + f] >> k) & 1) {
          const int64_t cn = pl.cell_neighbors[cc * 3 + f];
          const double ref_out = (f == 1) ? -1.0 : 1.0;
          jumpf[f] = (st.base_w[(int64_t)k * nc + cn] - w0) * ref_out * g.sdet;
        }
      }
#pragma unroll
      for (int e = 0; e < 21; ++e) {
        double v = 0.0;
#pragma unroll
        for (int t = 0; t < 15; ++t) v += c_T.G22u[e][t] * mC[t];
        W[e] = v;
      }
#pragma unroll
      for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int r = 0; r < 6; ++r)
#pragma unroll
          for (int l = 0; l < 3; ++l) {
            double v = 0.0;
#pragma unroll
            for (int t = 0; t < 10; ++t) v += c_T.G21[r][l][t] * Q[a * 10 + t];
            V[a][r][l] = v;
          }
    }
  }
  const uint8_t* patk0 = &s_patk[0][0][0];
  if (bulk) {
    if (active) sb_bulk_store_owned(vals + rb0, stg, ph0, 3 * (12 + 3 * (1 + NB)));
    sb_bulk_wait_read();
  } else {
    __syncwarp();
    sb_flush_group<24>(meta, &s_stage[warp][0], patk0, 15 * 32, 0, 0, -1,
                       0, 12 + 3 * (1 + NB), 12 + 3 * PS, 12 + 3 * PS + 3, vals, b, lane);
    __syncwarp();
  }
  /* ---- groups 1..4: facet 0, 1, 2 and interior BDM rows ------------------------- */
#pragma unroll 1
  for (int grp = 0; grp < 4; ++grp) {
#ifdef SB_EXP_NOFACET
    if (grp < 3) continue;
#endif
    /* the interior rows (grp 3) are owned: rank-ordered staging + bulk copy */
    const bool own = bulk && grp == 3;
    double* sdst = own ? stg + ph4 : stg;
    const uint8_t* pkg = own ? pk_own + 12 * 32 : nullptr;
    double* bdst = (own && b) ? b + dofs[15 * PS + 12] : nullptr;
    constexpr int rl4 = PS == 0 ? 15 : 18;
#ifdef SB_EXP_STOREONLY
    if (false) {
#else
    if (active) {
#endif
      if (PS == 0)
        sb_stage_flux_group<PS, 0>(c_T, c_M, g, W, V, fl, sl, 0.0, grp, sdst, pkg, rl4, bdst);
      else if (inside)
        sb_stage_flux_group<PS, 1>(c_T, c_M, g, W, V, fl, sl, grp < 3 ? jumpf[grp < 3 ? grp : 0] : 0.0,
                                   grp, sdst, pkg, rl4, bdst);
      else
        sb_stage_flux_group<PS, 2>(c_T, c_M, g, W, V, fl, sl, 0.0, grp, sdst, pkg, rl4, bdst);
    }
    if (own) {
      if (active) sb_bulk_store_owned(vals + rb4, stg, ph4, 3 * rl4);
      sb_bulk_wait_read();
    } else {
      __syncwarp();
      sb_flush_group<18>(meta, &s_stage[warp][0], patk0, 15 * 32, 3 + 3 * grp, 1 + grp,
                         grp < 3 ? grp : -1, 0, PS == 0 ? 15 : 18, 0, 12, vals, b, lane);
      __syncwarp();
    }
  }
}

/* Opt a kernel in to `smem` bytes of dynamic shared memory on the current device, once:
 * cudaFuncSetAttribute is not a cheap call -- it was made on every launch, and with several
 * sweep chains in flight (one thread + stream each) it waited for the other chains' running
 * kernels: 23 ms of host time per sb_assemble, 45 % of the wall time of an 8-chain ensemble
 * (tools/prof_ensemble.py).  The attribute is per kernel and device: remembered per pair. */
static void ensure_dynamic_smem(const void* fn, size_t smem) {
#ifdef SB_EMUL_BUILD
  (void)fn; (void)smem;
#else
  struct Entry { const void* fn; int dev; size_t smem; };
  static std::vector<Entry> cache;
  static std::mutex mtx;
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lk(mtx);
  for (Entry& e : cache)
    if (e.fn == fn && e.dev == dev) {
      if (e.smem >= smem) return;
      cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      e.smem = smem;
      return;
    }
  cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cache.push_back({fn, dev, smem});
#endif
}

/* launch k_rows_fast<NB, PS> for PS = 0 .. NB */
template <int NB, int PS>
struct RowsLauncher {
  static void go(const PlanDev& d, const StateDev& sd, double* vals, double* b, unsigned nblk,
                 cudaStream_t s) {
    if (d.stage_patk) {
      const size_t smem = (sizeof(double) * 32 * SB_STAGE_STRIDE + sizeof(SbWarpMeta))
                          * (SB_TILE / 32) + (size_t)SB_MAX_SMEM_PATTERNS * 15 * 32;
      ensure_dynamic_smem((const void*)k_rows_staged<NB, PS>, smem);
      SB_LAUNCH((k_rows_staged<NB, PS>), nblk, SB_TILE, smem, s, d, sd, vals, b);
    } else
      SB_LAUNCH((k_rows_fast<NB, PS>), nblk, SB_TILE, 0, s, d, sd, vals, b);
    RowsLauncher<NB, PS - 1>::go(d, sd, vals, b, nblk, s);
  }
};
template <int NB>
struct RowsLauncher<NB, -1> {
  static void go(const PlanDev&, const StateDev&, double*, double*, unsigned, cudaStream_t) {}
};

template <int NB>
static void launch_rows(const PlanDev& d, const StateDev& sd, double* vals, double* b,
                        unsigned nblk, cudaStream_t s, cudaEvent_t after_gen,
                        cudaEvent_t after_fast) {
#ifndef SB_EMUL_BUILD
  if (after_gen) cudaEventRecord(after_gen, s);
#endif
  RowsLauncher<NB, NB>::go(d, sd, vals, b, nblk, s);
#ifndef SB_EMUL_BUILD
  if (after_fast) cudaEventRecord(after_fast, s);
#endif
  if (d.n_special > 0) {
    const unsigned nsb = (unsigned)((d.n_special + SB_TILE - 1) / SB_TILE);
    SB_LAUNCH(k_rows_special<NB>, nsb * (unsigned)d.P, SB_TILE, 0, s, d, sd, vals, b, (int)nsb);
  }
}

/* ------------------------------------------------------------------------- */
__global__ void k_colidx(PlanDev pl, int32_t* col) {
  const int64_t c = blockIdx.x;
  const int L = pl.L;
  const int32_t* dofs = pl.cell_dofs + c * L;
  for (int e = threadIdx.x; e < L * L; e += blockDim.x) {
    const int i = e / L, j = e % L;
    const uint8_t rk = pl.patterns[((int64_t)pl.cell_pattern[c] * L + i) * L + j];
    if (rk != 0xFF) col[pl.rowptr[dofs[i]] + rk] = dofs[j];
  }
}

/* ---- a9 rebase --------------------------------------------------------------- */
__global__ void k_rebase_mean(PlanDev pl, const double* u, const double* base_w, int band,
                              double* tmp) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= pl.n_cells) return;
  const int32_t* dofs = pl.cell_dofs + c * pl.L + sb_ldg(1 + band, 0);
  const double m = (u[dofs[0]] + u[dofs[1]] + u[dofs[2]]) / 3.0;
  tmp[c] = base_w[(int64_t)band * pl.n_cells + c] + m;
}
__global__ void k_rebase_override(int64_t n, const int32_t* cells, const double* avg, double* tmp) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) tmp[cells[i]] = avg[i];
}
__global__ void k_rebase_apply(PlanDev pl, double* u, double* base_w, int band, const double* tmp) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= pl.n_cells) return;
  const int32_t* dofs = pl.cell_dofs + c * pl.L + sb_ldg(1 + band, 0);
  double* bw = base_w + (int64_t)band * pl.n_cells + c;
  const double shift = *bw - tmp[c];
  for (int i = 0; i < 3; ++i) u[dofs[i]] += shift;
  *bw = tmp[c];
}

/* ---- a11 Newton update + metrics -------------------------------------------- */
__device__ __forceinline__ void atomic_max_pos(double* addr, double v) {
  /* non-negative doubles order like their bit patterns */
  atomicMax(reinterpret_cast<unsigned long long*>(addr),
            static_cast<unsigned long long>(__double_as_longlong(v)));
}

__global__ void k_newton_update(int64_t N, double* u, const double* du, const double* b,
                                const double* tol, double omega, double max_du,
                                double rel_tol, double abs_tol, double log_alpha, double* out) {
  double bmax = 0, dmax = 0, rsum = 0, cmax = 0, nanf = 0, cnt = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N;
       i += (int64_t)gridDim.x * blockDim.x) {
    const double ui = u[i], di = du[i];
    const double bi = b ? b[i] : 0.0;
    double step = di;
    if (max_du > 0.0) step = fmin(fmax(step, -max_du), max_du);
    /* NewtonSolverLogDamping (newton_solver.py:542-547): sign(du) log1p(alpha |du|) / alpha */
    if (log_alpha > 0.0) step = copysign(log1p(fabs(step) * log_alpha) / log_alpha, step);
    const double un = ui - omega * step;
    u[i] = un;
    /* the reference evaluates its metrics after the update, i.e. against the
       new u_ (newton_solver.py:306-317 then :94-117) */
    if (isnan(un) || isnan(bi)) nanf = 1.0;
    const double ab = fabs(bi), ad = fabs(di);
    if (ab > bmax) bmax = ab;
    if (ad > dmax) dmax = ad;
    const double r = fabs(di / un);
    if (!isnan(r)) { rsum += r * r; cnt += 1.0; }
    const double cc = ad / (fabs(un) * rel_tol + abs_tol * (tol ? tol[i] : 1.0));
    if (cc > cmax) cmax = cc;
    if (isnan(ad)) nanf = 1.0;
  }
  __shared__ double sh[6][32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  double v[6] = {bmax, dmax, rsum, cmax, nanf, cnt};
  for (int o = 16; o > 0; o >>= 1) {
    for (int q = 0; q < 6; ++q) {
      const double other = __shfl_down_sync(0xffffffffu, v[q], o);
      v[q] = (q == 2 || q == 5) ? v[q] + other : fmax(v[q], other);
    }
  }
  if (lane == 0) for (int q = 0; q < 6; ++q) sh[q][w] = v[q];
  __syncthreads();
  if (w == 0) {
    const int nw = blockDim.x >> 5;
    for (int q = 0; q < 6; ++q) v[q] = lane < nw ? sh[q][lane] : 0.0;
    for (int o = 16; o > 0; o >>= 1)
      for (int q = 0; q < 6; ++q) {
        const double other = __shfl_down_sync(0xffffffffu, v[q], o);
        v[q] = (q == 2 || q == 5) ? v[q] + other : fmax(v[q], other);
      }
    if (lane == 0) {
      atomic_max_pos(&out[0], v[0]);
      atomic_max_pos(&out[1], v[1]);
      atomicAdd(&out[2], v[2]);
      atomic_max_pos(&out[3], v[3]);
      atomic_max_pos(&out[4], v[4]);
      atomicAdd(&out[5], v[5]);
    }
  }
}

/* ---- terminal current ----------------------------------------------------------- */
__global__ void k_surface_flux(PlanDev pl, const double* u, int p, int64_t n,
                               const int32_t* fcell, const int8_t* flocal,
                               const double* fsign, double* out) {
  __shared__ double sh[2][256];
  double flux = 0.0, len = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
    const int64_t c = fcell[i];
    const int f = flocal[i];
    const int32_t* dofs = pl.cell_dofs + c * pl.L + sb_lbdm(p, 3 * f);
    const double* xv = pl.cell_vertices + c * 6;
    const SbCellGeom g = sb_geom(xv);
    const double ref_out = (f == 1) ? -1.0 : 1.0;
    const double t = c_T.trI[0] * u[dofs[0]] + c_T.trI[1] * u[dofs[1]] + c_T.trI[2] * u[dofs[2]];
    flux += t * ref_out * g.sdet * fsign[i];
    /* edge opposite vertex f */
    const int a = (f == 0) ? 1 : 0, bb = (f == 2) ? 1 : 2;
    const double dx = xv[2 * bb] - xv[2 * a], dy = xv[2 * bb + 1] - xv[2 * a + 1];
    len += sqrt(dx * dx + dy * dy);
  }
  sh[0][threadIdx.x] = flux; sh[1][threadIdx.x] = len;
  __syncthreads();
  for (int o = blockDim.x / 2; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) {
      sh[0][threadIdx.x] += sh[0][threadIdx.x + o];
      sh[1][threadIdx.x] += sh[1][threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) { out[0] = sh[0][0]; out[1] = sh[1][0]; }
}

__global__ void k_fill(double* p, int64_t n, double v) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) p[i] = v;
}

/* ========================================================================= */
/* host side of the C ABI                                                     */
/* ========================================================================= */
template <typename T>
static int upload(sb_plan* pl, const T* h, size_t n, const T** d) {
  *d = nullptr;
  if (n == 0) { n = 1; }
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, n * sizeof(T));
  if (e != cudaSuccess) {
    snprintf(g_err, sizeof g_err, "cudaMalloc(%zu bytes) failed: %s", n * sizeof(T),
             cudaGetErrorString(e));
    return SB_ERR_ALLOC;
  }
  pl->allocs.push_back(p);
  if (h) {
    e = cudaMemcpy(p, h, n * sizeof(T), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
      snprintf(g_err, sizeof g_err, "cudaMemcpy H2D failed: %s", cudaGetErrorString(e));
      return SB_ERR_CUDA;
    }
  }
  *d = static_cast<const T*>(p);
  return SB_OK;
}

extern "C" const char* sb_last_error(void) { return g_err; }
extern "C" int sb_version(void) { return 100; }
extern "C" int64_t sb_launch_count(void) { return g_launches.load(); }
extern "C" int64_t sb_tables_size(void) { return (int64_t)sizeof(SbTables); }
extern "C" int64_t sb_plan_nnz(const sb_plan* p) { return p ? p->d.nnz : -1; }
extern "C" int sb_plan_is_native(const sb_plan* p) { return p ? p->d.native : 0; }
extern "C" int sb_plan_generation_signature(const sb_plan* p) {
  return (p && p->d.native) ? p->gen_sig : 0;
}

extern "C" void sb_plan_destroy(sb_plan* pl) {
  if (!pl) return;
#ifndef SB_EMUL_BUILD
  if (pl->ev_fork) cudaEventDestroy(pl->ev_fork);
  if (pl->ev_join) cudaEventDestroy(pl->ev_join);
  if (pl->side) cudaStreamDestroy(pl->side);
#endif
  for (void* p : pl->allocs) cudaFree(p);
#ifndef SB_EMUL_BUILD
  if (pl->ev0) {
    cudaEventDestroy(pl->ev0); cudaEventDestroy(pl->ev1);
    for (int i = 0; i < 3; ++i) cudaEventDestroy(pl->evs[i]);
  }
#endif
  for (int i = 0; i < SB_MAX_DEVICES; ++i)
    if (g_tables_owner[i] == pl) g_tables_owner[i] = nullptr;
  delete pl;
}

extern "C" int sb_plan_create(const sb_model* model, const sb_problem* pr, sb_plan** out) {
  if (!model || !pr || !out) { snprintf(g_err, sizeof g_err, "null argument"); return SB_ERR_ARG; }
  const int nbd = model->bands_have_dofs ? model->n_bands : 0;
  if (pr->P != 1 + nbd || model->n_bands > SB_MAX_BANDS || model->n_procs > SB_MAX_PROCS ||
      model->n_fields > SB_MAX_FIELDS) {
    snprintf(g_err, sizeof g_err, "inconsistent model/problem (P=%d, n_bands=%d)", pr->P,
             model->n_bands);
    return SB_ERR_ARG;
  }
  if (pr->n_tables != (int64_t)(sizeof(SbTables) / sizeof(double)) &&
      pr->n_tables * 8 != (int64_t)sizeof(SbTables)) {
    snprintf(g_err, sizeof g_err, "table blob has %lld doubles, expected %zu bytes",
             (long long)pr->n_tables, sizeof(SbTables));
    return SB_ERR_ARG;
  }
  sb_plan* pl = new sb_plan();
  pl->model = *model;
  /* process list known as a type?  (static lists, pdd_device.cuh; SIMUDO_B200_GEN_INTERPRET=1
     forces the interpreter: the parity tests run both) */
  pl->gen_sig = 0;
  if (!getenv("SIMUDO_B200_GEN_INTERPRET")) {
    if (sb_sig_matches<SbSigIB5>(*model, 3)) pl->gen_sig = 1;
    else if (sb_sig_matches<SbSigPnSRH>(*model, 2)) pl->gen_sig = 2;
  }
  memcpy(&pl->tables, pr->h_tables, sizeof(SbTables));
  const int L = 15 * pr->P;
  const int64_t nc = pr->n_cells, N = pr->n_dofs;
  PlanDev& d = pl->d;
  d.n_cells = nc; d.N = N; d.P = pr->P; d.L = L;
  d.nb = model->n_bands; d.nb_dof = nbd;
  d.nnz = pr->h_rowptr[N];
  int rc = SB_OK;
#define UP(field, src, n) if (rc == SB_OK) rc = upload(pl, src, (size_t)(n), &d.field)
  UP(cell_vertices, pr->h_cell_vertices, nc * 6);
  UP(cell_tag, pr->h_cell_tag, nc);
  UP(cell_dofs, pr->h_cell_dofs, nc * L);
  UP(cell_neighbors, pr->h_cell_neighbors, nc * 3);
  UP(cell_jump_mask, pr->h_cell_jump_mask, nc * 3);
  UP(rowptr, pr->h_rowptr, N + 1);
  UP(patterns, pr->h_patterns, pr->n_patterns * L * L);
  UP(cell_pattern, pr->h_cell_pattern, nc);
  UP(tag_params, pr->h_tag_params, (size_t)model->n_tags * SB_TAG_STRIDE);
  UP(tables, &pl->tables, 1);
  /* special cells: Dirichlet dofs and natural-BC rows */
  std::vector<int32_t> special(nc, -1);
  std::vector<uint64_t> mlo; std::vector<uint32_t> mhi; std::vector<int32_t> bcidx;
  std::vector<int32_t> natbeg;
  {
    std::vector<uint8_t> isbc;
    if (pr->n_bc) {
      isbc.assign((size_t)N, 0);
      for (int64_t i = 0; i < pr->n_bc; ++i) {
        const int32_t dof = pr->h_bc_dofs[i];
        if (dof < 0 || dof >= N || (i && pr->h_bc_dofs[i - 1] >= dof)) {
          snprintf(g_err, sizeof g_err, "bc dofs must be sorted, unique and in range");
          sb_plan_destroy(pl); return SB_ERR_ARG;
        }
        isbc[dof] = 1;
      }
    }
    std::vector<uint8_t> hasnat(nc, 0);
    for (int64_t i = 0; i < pr->n_natbc; ++i) {
      const int32_t c = pr->h_natbc_cell[i];
      if (c < 0 || c >= nc || (i && pr->h_natbc_cell[i - 1] > c) ||
          pr->h_natbc_row[i] < 0 || pr->h_natbc_row[i] >= L) {
        snprintf(g_err, sizeof g_err, "natbc entries must be sorted by cell and in range");
        sb_plan_destroy(pl); return SB_ERR_ARG;
      }
      hasnat[c] = 1;
    }
    int64_t natpos = 0;
    int32_t ns = 0;
    for (int64_t c = 0; c < nc; ++c) {
      uint64_t lo = 0; uint32_t hi = 0;
      if (pr->n_bc)
        for (int l = 0; l < L; ++l)
          if (isbc[pr->h_cell_dofs[c * L + l]]) {
            if (l < 64) lo |= (1ull << l); else hi |= (1u << (l - 64));
          }
      if (!(lo | hi) && !hasnat[c]) continue;
      special[c] = ns++;
      mlo.push_back(lo); mhi.push_back(hi);
      for (int l = 0; l < L; ++l) {
        int32_t idx = -1;
        const bool on = l < 64 ? ((lo >> l) & 1ull) : ((hi >> (l - 64)) & 1u);
        if (on) {
          const int32_t* bd = pr->h_bc_dofs;
          idx = (int32_t)(std::lower_bound(bd, bd + pr->n_bc, pr->h_cell_dofs[c * L + l]) - bd);
        }
        bcidx.push_back(idx);
      }
      while (natpos < pr->n_natbc && pr->h_natbc_cell[natpos] < c) ++natpos;
      natbeg.push_back((int32_t)natpos);
      while (natpos < pr->n_natbc && pr->h_natbc_cell[natpos] == c) ++natpos;
    }
    natbeg.push_back((int32_t)natpos);
    pl->n_special = ns;
  }
  pl->n_bc = pr->n_bc; pl->n_natbc = pr->n_natbc;
  UP(cell_special, special.data(), nc);
  /* closed-form layout?  (native.inl): 1 = CSR 'native', 2 = block CSR 'bsr3' */
  d.native = 0; d.nat_fid = nullptr; d.nat_finfo = nullptr; d.nat_fbase = nullptr; d.nat_dofF0 = 0; d.nat_nf = 0;
  {
    std::vector<int32_t> fid, finfo;
    std::vector<int64_t> fbase;
    int64_t F0 = 0, nfac = 0;
    int kind = 0;
    if (rc == SB_OK && detect_native(model, pr, false, fid, finfo, fbase, &F0, &nfac)) kind = 1;
    else if (rc == SB_OK && detect_native(model, pr, true, fid, finfo, fbase, &F0, &nfac)) kind = 2;
    if (kind) {
      UP(nat_fid, fid.data(), fid.size());
      UP(nat_finfo, finfo.data(), finfo.size());
      UP(nat_fbase, fbase.data(), fbase.size());
      d.nat_dofF0 = F0;
      d.nat_nf = nfac;
      d.native = kind;
    }
  }
  d.stage_patk = 0; d.owned_bulk = 0; d.patk = nullptr; d.rowbase = nullptr; d.rowlen = nullptr;
  d.n_patterns = pr->n_patterns;
  if (!d.native) {
    /* packed emission-order ranks for the fast row kernels */
    const int64_t np = pr->n_patterns;
    d.n_patterns = np;
    /* SIMUDO_B200_NO_SMEM_PATTERNS: debugging knob, forces the global-memory rank path */
    d.stage_patk = (np <= SB_MAX_SMEM_PATTERNS) && !getenv("SIMUDO_B200_NO_SMEM_PATTERNS");
    const int Pn = pr->P;
    std::vector<uint8_t> patk((size_t)np * L * 32, 0xFF);
    for (int64_t q = 0; q < np; ++q)
      for (int li = 0; li < L; ++li) {
        const uint8_t* row = pr->h_patterns + ((size_t)q * L + li) * L;
        uint8_t* o = patk.data() + ((size_t)q * L + li) * 32;
        const int p = li / 15, r = li % 15;
        for (int e = 0; e < 12; ++e) o[e] = row[15 * p + 3 + e];
        if (r < 3) {
          for (int qq = 0; qq < Pn; ++qq)
            for (int l = 0; l < 3; ++l) o[12 + 3 * qq + l] = row[15 * qq + l];
        } else {
          for (int l = 0; l < 3; ++l) { o[12 + l] = row[15 * p + l]; o[15 + l] = row[l]; }
        }
      }
    UP(patk, patk.data(), patk.size());
    /* coalesced row bases: the 3 dofs of one (sub-problem, entity) are consecutive in
       both supported numberings, so one offset + one length serve 3 rows */
    {
      const int ng = 5 * Pn;
      std::vector<int64_t> rbase((size_t)ng * nc);
      std::vector<uint16_t> rlen((size_t)ng * nc);
      bool ok = true;
      for (int64_t c = 0; c < nc && ok; ++c)
        for (int gq = 0; gq < ng; ++gq) {
          const int p = gq / 5, e = gq % 5;
          const int li0 = 15 * p + (e == 0 ? 0 : 3 * e);
          const int32_t* dd = pr->h_cell_dofs + c * L + li0;
          const int64_t r0 = pr->h_rowptr[dd[0]];
          const int64_t len = pr->h_rowptr[dd[0] + 1] - r0;
          if (dd[1] != dd[0] + 1 || dd[2] != dd[0] + 2 || len > 65535 ||
              pr->h_rowptr[dd[0] + 2] - pr->h_rowptr[dd[0] + 1] != len ||
              pr->h_rowptr[dd[0] + 3] - pr->h_rowptr[dd[0] + 2] != len) { ok = false; break; }
          rbase[(size_t)gq * nc + c] = r0;
          rlen[(size_t)gq * nc + c] = (uint16_t)len;
        }
      if (!ok) {
        snprintf(g_err, sizeof g_err, "dof numbering must keep the three dofs of each mesh entity "
                 "consecutive (the 'ufc' and 'b200' numberings do)");
        sb_plan_destroy(pl); return SB_ERR_UNSUPPORTED;
      }
      UP(rowbase, rbase.data(), rbase.size());
      UP(rowlen, rlen.data(), rlen.size());
      /* owned rows (DG, interior BDM) of the structural layout: fixed lengths, even */
      bool bulk = true;
      for (int p = 0; p < Pn && bulk; ++p)
        for (int64_t c = 0; c < nc; ++c)
          if (rlen[(size_t)(5 * p) * nc + c] != 12 + 3 * Pn ||
              rlen[(size_t)(5 * p + 4) * nc + c] != (p == 0 ? 15 : 18)) { bulk = false; break; }
      d.owned_bulk = bulk && !getenv("SIMUDO_B200_NO_BULK_STORE");
    }
  }
  {
    std::vector<int32_t> spc;
    for (int64_t c = 0; c < nc; ++c) if (special[c] >= 0) spc.push_back((int32_t)c);
    UP(sp_cells, spc.data(), spc.size());
    d.n_special = (int64_t)spc.size();
  }
  UP(sp_mask_lo, mlo.data(), mlo.size());
  UP(sp_mask_hi, mhi.data(), mhi.size());
  UP(sp_bcidx, bcidx.data(), bcidx.size());
  UP(sp_nat_begin, natbeg.data(), natbeg.size());
  UP(natbc_row, pr->h_natbc_row, pr->n_natbc);
  const double* tmp = nullptr;
  if (rc == SB_OK) rc = upload<double>(pl, nullptr, (size_t)nc, &tmp);
  pl->tmp_cells = const_cast<double*>(tmp);
  const double* scr = nullptr;
  const size_t per_cell = d.native ? ((size_t)model->n_bands * (SB_NAT_NA + SB_NG) + 18)
                                   : ((size_t)model->n_bands * 44 + (size_t)nbd * SB_NG);
  /* native: tile-blocked by 32 cells (natQ_ptr) */
  if (rc == SB_OK) rc = upload<double>(pl, nullptr, (size_t)((nc + 31) / 32 * 32) * per_cell, &scr);
  d.scratch = const_cast<double*>(scr);
  const int* ctr = nullptr;
  if (rc == SB_OK) rc = upload<int>(pl, nullptr, 8, &ctr);
  d.nat_tile_ctr = const_cast<int*>(ctr);
#undef UP
  if (rc != SB_OK) { sb_plan_destroy(pl); return rc; }
  *out = pl;
  return SB_OK;
}

/* The tables are shared __constant__ state: plans of one device must be used from one
 * stream at a time (the copy is stream-ordered on the caller's stream).              */
static SbTables g_bound_tables[SB_MAX_DEVICES];
static sb_model g_bound_model[SB_MAX_DEVICES];
static bool g_bound_valid[SB_MAX_DEVICES] = {false};

static std::atomic<int64_t> g_table_uploads{0}, g_table_upload_us{0};
static int bind_tables(sb_plan* pl, cudaStream_t s) {
  const int slot = current_device_slot();
  if (g_tables_owner[slot] == pl) return SB_OK;
  /* another plan with the same tables and model (the members of a sweep ensemble): nothing to
     copy, and no kernel of a concurrent stream ever sees the constants change */
  if (g_bound_valid[slot] && memcmp(&g_bound_tables[slot], &pl->tables, sizeof(SbTables)) == 0 &&
      memcmp(&g_bound_model[slot], &pl->model, sizeof(sb_model)) == 0) {
    g_tables_owner[slot] = pl;
    return SB_OK;
  }
#ifndef SB_EMUL_BUILD
  const auto t0_ = std::chrono::steady_clock::now();
#endif
  CUDA_TRY(cudaMemcpyToSymbolAsync(c_T, &pl->tables, sizeof(SbTables), 0,
                                   cudaMemcpyHostToDevice, s));
  CUDA_TRY(cudaMemcpyToSymbolAsync(c_M, &pl->model, sizeof(sb_model), 0,
                                   cudaMemcpyHostToDevice, s));
#ifndef SB_EMUL_BUILD
  g_table_uploads += 1;
  g_table_upload_us += (int64_t)std::chrono::duration_cast<std::chrono::microseconds>(
      std::chrono::steady_clock::now() - t0_).count();
#endif
  g_tables_owner[slot] = pl;
  g_bound_tables[slot] = pl->tables;
  g_bound_model[slot] = pl->model;
  g_bound_valid[slot] = true;
  return SB_OK;
}

/* side stream of a plan: everything enqueued on `s` so far happens before what follows on the
   returned stream; join_side_stream makes `s` wait for it again */
static inline cudaStream_t fork_side_stream(sb_plan* pl, cudaStream_t s) {
#ifdef SB_EMUL_BUILD
  (void)pl;
  return s;
#else
  if (getenv("SIMUDO_B200_NO_SIDE_STREAM")) return s;
  if (!pl->side) {
    if (cudaStreamCreateWithFlags(&pl->side, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&pl->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&pl->ev_join, cudaEventDisableTiming) != cudaSuccess) {
      cudaGetLastError();
      pl->side = nullptr;
      return s;
    }
  }
  cudaEventRecord(pl->ev_fork, s);
  cudaStreamWaitEvent(pl->side, pl->ev_fork, 0);
  return pl->side;
#endif
}
static inline void join_side_stream(sb_plan* pl, cudaStream_t s) {
#ifdef SB_EMUL_BUILD
  (void)pl; (void)s;
#else
  if (!pl->side || getenv("SIMUDO_B200_NO_SIDE_STREAM")) return;
  cudaEventRecord(pl->ev_join, pl->side);
  cudaStreamWaitEvent(s, pl->ev_join, 0);
#endif
}
static inline void rec_stage_event(sb_plan* pl, int i, cudaStream_t s) {
#ifndef SB_EMUL_BUILD
  if (pl->profiling) cudaEventRecord(pl->evs[i], s);
#else
  (void)pl; (void)i; (void)s;
#endif
}

extern "C" int sb_plan_zero_values(sb_plan* pl, double* d_vals, void* stream) {
  if (!pl || !d_vals) { snprintf(g_err, sizeof g_err, "null argument"); return SB_ERR_ARG; }
  CUDA_TRY(cudaMemsetAsync(d_vals, 0, (size_t)pl->d.nnz * sizeof(double), (cudaStream_t)stream));
  return SB_OK;
}

extern "C" int sb_assemble(sb_plan* pl, const sb_state* st, double* d_vals, double* d_b,
                           void* stream) {
  if (!pl || !st || !st->d_u) { snprintf(g_err, sizeof g_err, "null argument"); return SB_ERR_ARG; }
  if (pl->d.nb_dof && !st->d_base_w) {
    snprintf(g_err, sizeof g_err, "base_w required when bands have dofs"); return SB_ERR_ARG;
  }
  if (pl->model.n_fields && !st->d_phi_opt) {
    snprintf(g_err, sizeof g_err, "phi_opt required when n_fields > 0"); return SB_ERR_ARG;
  }
  if ((pl->n_bc && !st->d_bc_values) || (pl->n_natbc && !st->d_natbc_values)) {
    snprintf(g_err, sizeof g_err, "bc_values / natbc_values required"); return SB_ERR_ARG;
  }
  /* rows that one cell owns leave the SM as cp.async.bulk copies whose 16-byte phase is
     taken from the value index: the array itself must be 16-byte aligned */
  if (d_vals && (pl->d.native || pl->d.owned_bulk) && ((uintptr_t)d_vals & 15)) {
    snprintf(g_err, sizeof g_err, "d_vals must be 16-byte aligned"); return SB_ERR_ARG;
  }
  cudaStream_t s = (cudaStream_t)stream;
  int rc = bind_tables(pl, s);
  if (rc != SB_OK) return rc;
  StateDev sd{st->d_u, st->d_base_w, st->d_thmeq_phi, st->d_phi_opt, st->d_bc_values,
              st->d_natbc_values};
  const int64_t nc = pl->d.n_cells;
  /* residual: facet rows receive two atomic adds -> start from zero (one memset) */
  if (d_b) CUDA_TRY(cudaMemsetAsync(d_b, 0, (size_t)pl->d.N * sizeof(double), s));
  if (pl->d.nb == 0)     /* no band => no k_moments launch to carry the slot zeroing */
    SB_LAUNCH(k_zero_shared, (unsigned)((nc + 255) / 256), 256, 0, s, pl->d, d_vals, (double*)nullptr);
#ifndef SB_EMUL_BUILD
  if (pl->profiling) CUDA_TRY(cudaEventRecord(pl->ev0, s));
#endif
  const unsigned nblk = (unsigned)((nc + SB_TILE - 1) / SB_TILE);
  if (pl->d.native) {
    /* closed-form path (native.inl): moments, generation, one row launch per sub-problem,
       boundary cells through the generic emitter; no slot is zeroed, no atomic on vals */
    const PlanDev& d = pl->d;
    CUDA_TRY(cudaMemsetAsync(d.nat_tile_ctr, 0, 8 * sizeof(int), s));
    SB_LAUNCH(k_moments_nat, nblk * (unsigned)d.nb, SB_TILE, 0, s, d, sd, (int)nblk);
#ifndef SB_EMUL_BUILD
    if (pl->profiling) CUDA_TRY(cudaEventRecord(pl->evs[0], s));
#endif
#define SB_NAT_CASE(NBV)                                                                      \
    case NBV:                                                                                 \
      if (NBV == 3 && pl->gen_sig == 1)                                                       \
        SB_LAUNCH((k_generation_nat<3, SbSigIB5>), nblk, SB_TILE, 0, s, d, sd);                \
      else if (NBV == 2 && pl->gen_sig == 2)                                                  \
        SB_LAUNCH((k_generation_nat<2, SbSigPnSRH>), nblk, SB_TILE, 0, s, d, sd);              \
      else                                                                                    \
        SB_LAUNCH(k_generation_nat<NBV>, nblk, SB_TILE, 0, s, d, sd);                         \
      rec_stage_event(pl, 1, s);                                                              \
      if (d.n_special > 0) {                                                                  \
        /* boundary cells (0.3 %): a latency-bound kernel of ~100 CTAs that touches no value   \
           or residual slot the row kernels store to (facet residual rows: commutative         \
           two-addend atomics) -- launched first, on the side stream, it runs beside them     \
           instead of 0.15 ms after them */                                                   \
        const unsigned nsb = (unsigned)((d.n_special + SB_TILE - 1) / SB_TILE);               \
        cudaStream_t ss = fork_side_stream(pl, s);                                            \
        SB_LAUNCH(k_rows_special<NBV>, nsb * (unsigned)d.P, SB_TILE, 0, ss, d, sd, d_vals,    \
                  d_b, (int)nsb);                                                             \
      }                                                                                       \
      NatRowsLauncher<NBV, NBV>::go(d, sd, d_vals, d_b, nblk, s);                             \
      rec_stage_event(pl, 2, s);                                                              \
      if (d.n_special > 0) join_side_stream(pl, s);                                           \
      break;
#ifdef SB_PROBE_NB          /* developer builds (tools/probe_build.sh): one band count only */
    switch (d.nb_dof) { default: SB_NAT_CASE(SB_PROBE_NB) }
#else
    switch (d.nb_dof) {
      SB_NAT_CASE(1) SB_NAT_CASE(2) SB_NAT_CASE(3)
      default: SB_NAT_CASE(4)
    }
#endif
#undef SB_NAT_CASE
#ifndef SB_EMUL_BUILD
    if (pl->profiling) CUDA_TRY(cudaEventRecord(pl->ev1, s));
#endif
    g_launches += 2 + d.P + (d.n_special > 0 ? 1 : 0);
    CUDA_TRY(cudaGetLastError());
    return SB_OK;
  }
  int nl = 1;
  if (pl->d.nb > 0) {
    SB_LAUNCH(k_moments, nblk * (unsigned)pl->d.nb, SB_TILE, 0, s, pl->d, sd, (int)nblk, d_vals);
    ++nl;
  }
  cudaEvent_t eg = nullptr, ef = nullptr;
#ifndef SB_EMUL_BUILD
  if (pl->profiling) { CUDA_TRY(cudaEventRecord(pl->evs[0], s)); eg = pl->evs[1]; ef = pl->evs[2]; }
#endif
#ifdef SB_PROBE_NB
  switch (0) {
#else
  switch (pl->d.nb_dof) {
#endif
    case 0: launch_rows<0>(pl->d, sd, d_vals, d_b, nblk, s, eg, ef); break;
#ifndef SB_PROBE_NB
    case 1: SB_LAUNCH(k_generation<1>, nblk, SB_TILE, 0, s, pl->d, sd);
            launch_rows<1>(pl->d, sd, d_vals, d_b, nblk, s, eg, ef); ++nl; break;
    case 2: SB_LAUNCH(k_generation<2>, nblk, SB_TILE, 0, s, pl->d, sd);
            launch_rows<2>(pl->d, sd, d_vals, d_b, nblk, s, eg, ef); ++nl; break;
    case 3: SB_LAUNCH(k_generation<3>, nblk, SB_TILE, 0, s, pl->d, sd);
            launch_rows<3>(pl->d, sd, d_vals, d_b, nblk, s, eg, ef); ++nl; break;
    default: SB_LAUNCH(k_generation<4>, nblk, SB_TILE, 0, s, pl->d, sd);
             launch_rows<4>(pl->d, sd, d_vals, d_b, nblk, s, eg, ef); ++nl; break;
#endif
  }
  nl += pl->d.P + (pl->d.n_special > 0 ? 1 : 0) - 1;
#ifndef SB_EMUL_BUILD
  if (pl->profiling) CUDA_TRY(cudaEventRecord(pl->ev1, s));
#endif
  g_launches += nl;

  CUDA_TRY(cudaGetLastError());
  return SB_OK;
}

/* CUDA-event timing of the dominant kernel (k_assemble) of the last sb_assemble */
extern "C" int sb_plan_set_profiling(sb_plan* pl, int on) {
  if (!pl) return SB_ERR_ARG;
#ifndef SB_EMUL_BUILD
  if (on && !pl->ev0) {
    CUDA_TRY(cudaEventCreate(&pl->ev0));
    CUDA_TRY(cudaEventCreate(&pl->ev1));
    for (int i = 0; i < 3; ++i) CUDA_TRY(cudaEventCreate(&pl->evs[i]));
  }
#endif
  pl->profiling = on;
  return SB_OK;
}
extern "C" int sb_plan_last_kernel_ms(sb_plan* pl, double* ms) {
  if (!pl || !ms) return SB_ERR_ARG;
  *ms = 0.0;
#ifndef SB_EMUL_BUILD
  if (!pl->profiling || !pl->ev1) { snprintf(g_err, sizeof g_err, "profiling is off"); return SB_ERR_ARG; }
  CUDA_TRY(cudaEventSynchronize(pl->ev1));
  float f = 0.f;
  CUDA_TRY(cudaEventElapsedTime(&f, pl->ev0, pl->ev1));
  *ms = f;
#endif
  return SB_OK;
}

extern "C" int sb_plan_last_stage_ms(sb_plan* pl, double* ms4) {
  if (!pl || !ms4) return SB_ERR_ARG;
  for (int i = 0; i < 4; ++i) ms4[i] = 0.0;
#ifndef SB_EMUL_BUILD
  if (!pl->profiling || !pl->ev1) { snprintf(g_err, sizeof g_err, "profiling is off"); return SB_ERR_ARG; }
  CUDA_TRY(cudaEventSynchronize(pl->ev1));
  cudaEvent_t e[5] = {pl->ev0, pl->evs[0], pl->evs[1], pl->evs[2], pl->ev1};
  for (int i = 0; i < 4; ++i) {
    float f = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&f, e[i], e[i + 1]));
    ms4[i] = f;
  }
#endif
  return SB_OK;
}

/* ---- a8: exterior-facet integrals of the natural boundary conditions -----------------
 * + oint g psi_i . n ds for the three facet dofs i of every listed exterior facet
 * (the `ds` term of poisson_drift_diffusion1.py1:47 built by fem/spatial.py:237-273),
 * g = c0 + (P2 function on the facet's cell), 3 Gauss-Legendre points (degree 4, SURVEY
 * Appendix A6).  One thread per facet; results go straight to their slots of the
 * natural-BC value array sb_assemble consumes.                                        */
struct SbNatTables {
  double w[3];            /* GL weights on [0, 1]                                     */
  double p2[3][3][6];     /* [local facet][point][P2 node] Lagrange values            */
  double tr[3][3][3];     /* [local facet][facet dof][point] BDM scaled-normal trace  */
};

__global__ void k_natbc_eval(PlanDev pl, int64_t n, const int32_t* cell, const int8_t* local,
                             const int32_t* slot, const double* c0, const double* p2vals,
                             SbNatTables T, double* out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t c = cell[i];
  const int l = local[i];
  const SbCellGeom g = sb_geom(pl.cell_vertices + c * 6);
  const double sgn = ((l == 1) ? -1.0 : 1.0) * g.sdet;       /* outward / reference normal */
  double acc[3] = {0.0, 0.0, 0.0};
  for (int q = 0; q < 3; ++q) {
    double v = c0 ? c0[i] : 0.0;
    if (p2vals)
      for (int a = 0; a < 6; ++a) v += p2vals[i * 6 + a] * T.p2[l][q][a];
    for (int k = 0; k < 3; ++k) acc[k] += T.w[q] * v * T.tr[l][k][q];
  }
  for (int k = 0; k < 3; ++k) out[slot[3 * i + k]] = sgn * acc[k];
}

extern "C" int sb_natbc_eval(sb_plan* pl, int64_t n_facets, const int32_t* d_cell,
                             const int8_t* d_local, const int32_t* d_slot, const double* d_c0,
                             const double* d_p2, const double* h_tables, int64_t n_tables,
                             double* d_natbc_values, void* stream) {
  if (!pl || n_facets < 0 || !d_cell || !d_local || !d_slot || !h_tables || !d_natbc_values ||
      n_tables != (int64_t)(sizeof(SbNatTables) / sizeof(double))) {
    snprintf(g_err, sizeof g_err, "bad argument to sb_natbc_eval"); return SB_ERR_ARG;
  }
  if (n_facets == 0) return SB_OK;
  SbNatTables T;
  memcpy(&T, h_tables, sizeof T);
  SB_LAUNCH(k_natbc_eval, (unsigned)((n_facets + 127) / 128), 128, 0, (cudaStream_t)stream, pl->d,
            n_facets, d_cell, d_local, d_slot, d_c0, d_p2, T, d_natbc_values);
  g_launches += 1;
  CUDA_TRY(cudaGetLastError());
  return SB_OK;
}

/* ---- FP64 roofline denominator -----------------------------------------------------
 * DFMA throughput of the device, measured: 8 independent FMA chains per thread, enough
 * resident warps to cover the pipe latency (1024 threads x 2 CTAs per SM).             */
__global__ void __launch_bounds__(1024, 2) k_dfma_peak(double* out, int iters, double a, double b) {
  double v[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = fma(v[i], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += v[i];
  if (s == 1.2345) out[0] = s;
}

extern "C" int sb_measure_fp64_peak(double* tflops) {
  if (!tflops) return SB_ERR_ARG;
  *tflops = 0.0;
#ifndef SB_EMUL_BUILD
  int dev = 0, sms = 0;
  CUDA_TRY(cudaGetDevice(&dev));
  CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  double* d = nullptr;
  CUDA_TRY(cudaMalloc(&d, 8));
  cudaEvent_t e0, e1;
  CUDA_TRY(cudaEventCreate(&e0));
  CUDA_TRY(cudaEventCreate(&e1));
  const int iters = 20000, grid = 2 * sms;
  k_dfma_peak<<<grid, 1024>>>(d, 200, 1.0000001, 1e-9);
  double best = 0.0;
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0);
    k_dfma_peak<<<grid, 1024>>>(d, iters, 1.0000001, 1e-9);
    cudaEventRecord(e1);
    CUDA_TRY(cudaEventSynchronize(e1));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    const double t = 2.0 * 8 * (double)iters * 1024.0 * grid / (ms * 1e-3) * 1e-12;
    if (t > best) best = t;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d);
  g_launches += 4;
  *tflops = best;
#endif
  return SB_OK;
}

extern "C" int sb_plan_column_indices(sb_plan* pl, int32_t* d_colidx, void* stream) {
  if (!pl || !d_colidx) { snprintf(g_err, sizeof g_err, "null argument"); return SB_ERR_ARG; }
  SB_LAUNCH(k_colidx, (unsigned)pl->d.n_cells, 256, 0, (cudaStream_t)stream, pl->d, d_colidx);
  g_launches += 1;
  CUDA_TRY(cudaGetLastError());
  return SB_OK;
}

extern "C" int sb_rebase_w(sb_plan* pl, double* d_u, double* d_base_w, int32_t band,
                           int64_t n_bc_cells, const int32_t* d_bc_cells,
                           const double* d_bc_avg, void* stream) {
  if (!pl || !d_u || !d_base_w || band < 0 || band >= pl->d.nb_dof) {
    snprintf(g_err, sizeof g_err, "bad argument to sb_rebase_w"); return SB_ERR_ARG;
  }
  cudaStream_t s = (cudaStream_t)stream;
  const unsigned nblk = (unsigned)((pl->d.n_cells + 255) / 256);
  SB_LAUNCH(k_rebase_mean, nblk, 256, 0, s, pl->d, d_u, d_base_w, band, pl->tmp_cells);
  if (n_bc_cells > 0)
    SB_LAUNCH(k_rebase_override, (unsigned)((n_bc_cells + 255) / 256), 256, 0, s, n_bc_cells, d_bc_cells, d_bc_avg, pl->tmp_cells);
  SB_LAUNCH(k_rebase_apply, nblk, 256, 0, s, pl->d, d_u, d_base_w, band, pl->tmp_cells);
  g_launches += (n_bc_cells > 0) ? 3 : 2;
  CUDA_TRY(cudaGetLastError());
  return SB_OK;
}

extern "C" int sb_newton_update_damped(sb_plan* pl, double* d_u, const double* d_du,
                                       const double* d_b, const double* d_tol, double omega,
                                       double max_du, double log_alpha, double rel_tol,
                                       double abs_tol, double* d_out, void* stream);

extern "C" int sb_newton_update(sb_plan* pl, double* d_u, const double* d_du, const double* d_b,
                                const double* d_tol, double omega, double max_du,
                                double rel_tol, double abs_tol, double* d_out, void* stream) {
  return sb_newton_update_damped(pl, d_u, d_du, d_b, d_tol, omega, max_du, 0.0, rel_tol, abs_tol,
                                 d_out, stream);
}

extern "C" int sb_newton_update_damped(sb_plan* pl, double* d_u, const double* d_du,
                                       const double* d_b, const double* d_tol, double omega,
                                       double max_du, double log_alpha, double rel_tol,
                                       double abs_tol, double* d_out, void* stream) {
  if (!pl || !d_u || !d_du || !d_out) { snprintf(g_err, sizeof g_err, "null argument"); return SB_ERR_ARG; }
  cudaStream_t s = (cudaStream_t)stream;
  CUDA_TRY(cudaMemsetAsync(d_out, 0, 6 * sizeof(double), s));
  const int64_t N = pl->d.N;
  unsigned nblk = (unsigned)std::min<int64_t>((N + 255) / 256, 148 * 8);
  SB_LAUNCH(k_newton_update, nblk, 256, 0, s, N, d_u, d_du, d_b, d_tol, omega, max_du, rel_tol,
                                       abs_tol, log_alpha, d_out);
  g_launches += 1;
  CUDA_TRY(cudaGetLastError());
  return SB_OK;
}

extern "C" int sb_surface_flux(sb_plan* pl, const double* d_u, int32_t p, int64_t n_facets,
                               const int32_t* d_facet_cell, const int8_t* d_facet_local,
                               const double* d_facet_sign, double* d_out, void* stream) {
  if (!pl || !d_u || !d_out || p < 0 || p >= pl->d.P) {
    snprintf(g_err, sizeof g_err, "bad argument to sb_surface_flux"); return SB_ERR_ARG;
  }
  cudaStream_t s = (cudaStream_t)stream;
  int rc = bind_tables(pl, s);
  if (rc != SB_OK) return rc;
  SB_LAUNCH(k_surface_flux, 1, 256, 0, s, pl->d, d_u, p, n_facets, d_facet_cell, d_facet_local,
                                   d_facet_sign, d_out);
  g_launches += 1;
  CUDA_TRY(cudaGetLastError());
  return SB_OK;
}

#include "points.inl"
#include "optics.inl"
#include "lcn.inl"
