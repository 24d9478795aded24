/*
 * pdd_device.cuh -- device-side math of the assembly hot path.
 *
 * What the reference computes with FFC-generated tabulate_tensor at
 * 6 + 25 + 121 quadrature points per cell (simudo/fem/newton_solver.py:61-77;
 * forms simudo/physics/poisson_drift_diffusion1.py1:42-47,264-303), reorganised
 * for the GPU:
 *
 *   phase A  (one thread per cell, per band): the 121-point rule is walked in
 *            its collapsed tensor form; the non-polynomial coefficients
 *            rho(phi,w)/eps, d rho/d chi, c = 1/(mu u), dc/dchi are never
 *            multiplied with basis functions point by point -- instead their
 *            moments against an orthonormal Dubiner basis are accumulated by
 *            sum factorisation (5 + 8 + 5 FMAs per point instead of ~400).
 *   phase B  (one thread per cell, per band): the 25-point rule for the
 *            generation term g and its partial derivatives, same idea.
 *   phase C  (one warp per cell): moments x constant expansion tensors x cell
 *            geometry -> structurally non-zero entries of the element tensor,
 *            Dirichlet application, and scatter into the fixed CSR.
 *
 * Everything is written as lane-strided loops so the very same source runs
 * under tests/host_emulation (lane = 0, nlanes = 1) for CPU-side debugging of
 * the kernel logic; that harness is test infrastructure only.
 */
#ifndef SB_PDD_DEVICE_CUH
#define SB_PDD_DEVICE_CUH

#include <math.h>
#include <stdint.h>
#include "../../include/simudo_b200.h"
#include "pdd_tables.h"

#if defined(__CUDACC__)
#define SB_HD __host__ __device__ __forceinline__
#define SB_D __device__ __forceinline__
#else
#define SB_HD inline
#define SB_D inline
#endif

#define SB_MAXL (15 * (1 + SB_MAX_BANDS))

/* moment block sizes */
#define SB_NMC 15      /* c against P4            */
#define SB_NQ 20       /* c' g^a against P3, a=0,1 */
#define SB_NRX 6       /* d(s u/eps)/dchi vs P2   */
#define SB_NR 3        /* rho/eps vs P1           */
#define SB_NG (3 + 6 + 6 * SB_MAX_BANDS)   /* g, dg/dphi, dg/dw_l */

struct SbCellGeom {
  double G11, G12, G22;   /* J^T J            */
  double adet, iad, sdet; /* |detJ|, 1/|detJ|, sign(detJ) */
};

SB_HD SbCellGeom sb_geom(const double* xv /* [3][2] */) {
  const double J00 = xv[2] - xv[0], J01 = xv[4] - xv[0];
  const double J10 = xv[3] - xv[1], J11 = xv[5] - xv[1];
  const double det = J00 * J11 - J01 * J10;
  SbCellGeom g;
  g.G11 = J00 * J00 + J10 * J10;
  g.G12 = J00 * J01 + J10 * J11;
  g.G22 = J01 * J01 + J11 * J11;
  g.adet = fabs(det);
  g.iad = 1.0 / g.adet;
  g.sdet = det > 0.0 ? 1.0 : -1.0;
  return g;
}

/* local dof helpers: sub-problem p, DG1 dof i | BDM dof i */
SB_HD int sb_ldg(int p, int i) { return 15 * p + i; }
SB_HD int sb_lbdm(int p, int i) { return 15 * p + 3 + i; }
/* which local facet a local dof sits on (-1: cell interior) */
SB_HD int sb_local_facet(int l) {
  const int r = l % 15;
  return (r >= 3 && r < 12) ? (r - 3) / 3 : -1;
}

/* Pointwise band quantities at chi = e phi + w (band statistics
 * poisson_drift_diffusion1.py1:170-245).  Outputs:
 *   su  = s u / eps               dsu = d(su)/dchi
 *   c   = dd_scale / (mu u)       dc  = dc/dchi                      */
struct SbBandPoint { double su, dsu, c, dc, u, du; };

SB_HD SbBandPoint sb_band_point(int kind, double s, int const_mob, double beta,
                                double E, double N, double mob, double inv_eps,
                                double dd_scale, double chi) {
  SbBandPoint r;
  const double t = s * beta * (chi - E);
  const double e = exp(t);
  if (kind == SB_BAND_NONDEGENERATE) {
    const double inv = 1.0 / e;
    r.u = N * inv;
    r.du = -s * beta * r.u;
    r.c = dd_scale * e / (mob * N);
    r.dc = s * beta * r.c;
  } else {
    const double f = 1.0 / (1.0 + e);
    const double omf = e * f;                 /* 1 - f */
    r.u = N * f;
    const double df = -s * beta * f * omf;    /* df/dchi */
    r.du = N * df;
    if (const_mob) {
      r.c = dd_scale * (1.0 + e) / (mob * N);
      r.dc = dd_scale * s * beta * e / (mob * N);
    } else {
      /* mu = mu0 [(1-f) - 1/2 f^0.33 (1-f)^1.33]  (:228-231) */
      const double fa = pow(f, 0.33), ob = pow(omf, 1.33);
      const double mu = mob * (omf - 0.5 * fa * ob);
      const double dmu = mob * (-df - 0.5 * (0.33 * fa / f * df * ob
                                             + fa * 1.33 * ob / omf * (-df)));
      r.c = dd_scale / (mu * r.u);
      r.dc = -r.c * (dmu / mu + df / f);
    }
  }
  r.su = s * r.u * inv_eps;
  r.dsu = s * r.du * inv_eps;
  return r;
}

/* ------------------------------------------------------------------------- *
 * phase A: moments of one band over the collapsed rule of R.
 *   chi(X,Y) = chi0 + chix X + chiy Y                      (P1, incl. base_w)
 *   gc[a][r]: Dubiner-P2 coefficients of g^a = (J^T J jhat)^a
 * outputs (overwritten unless noted):
 *   mC[15], Q[2][10]          if do_dd
 *   mRx[6], mR[3] (+=)        if do_rho
 * ------------------------------------------------------------------------- */
SB_HD void sb_phaseA_band(const SbRule& R, bool do_rho, bool do_dd,
                          int kind, double s, int const_mob, double beta,
                          double E, double N, double mob, double inv_eps,
                          double dd_scale, double chi0, double chix, double chiy,
                          const double (*gc)[6],
                          double* mC, double* Q, double* mR, double* mRx) {
  const int m = R.m;
  if (do_dd) {
    for (int t = 0; t < SB_NMC; ++t) mC[t] = 0.0;
    for (int t = 0; t < SB_NQ; ++t) Q[t] = 0.0;
  }
  if (do_rho) for (int t = 0; t < SB_NRX; ++t) mRx[t] = 0.0;
  for (int j = 0; j < m; ++j) {
    const double bj = R.b[j];
    const double omb = 1.0 - bj;
    /* row-wise pieces of g^a: H[a][p] = sum_q gc[a][(p,q)] h_pq(b_j) */
    double H[2][3];
    if (do_dd) {
      const double* h = R.h6[j];   /* order: (0,0) (0,1) (1,0) (0,2) (1,1) (2,0) */
      for (int a = 0; a < 2; ++a) {
        H[a][0] = gc[a][0] * h[0] + gc[a][1] * h[1] + gc[a][3] * h[3];
        H[a][1] = gc[a][2] * h[2] + gc[a][4] * h[4];
        H[a][2] = gc[a][5] * h[5];
      }
    }
    double Sc[5] = {0, 0, 0, 0, 0};
    double Sq[2][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}};
    double Sr[2] = {0, 0};
    double Sx[3] = {0, 0, 0};
    const double chij = chi0 + chiy * bj;
    const double chis = chix * omb;
    for (int i = 0; i < m; ++i) {
      const double chi = chij + chis * R.a[i];
      const SbBandPoint bp = sb_band_point(kind, s, const_mob, beta, E, N, mob,
                                           inv_eps, dd_scale, chi);
      const double* wg = R.wag[i];
      if (do_dd) {
        const double* g3 = R.g3[i];
        const double q0 = bp.dc * (g3[0] * H[0][0] + g3[1] * H[0][1] + g3[2] * H[0][2]);
        const double q1 = bp.dc * (g3[0] * H[1][0] + g3[1] * H[1][1] + g3[2] * H[1][2]);
        for (int p = 0; p < 5; ++p) Sc[p] += wg[p] * bp.c;
        for (int p = 0; p < 4; ++p) { Sq[0][p] += wg[p] * q0; Sq[1][p] += wg[p] * q1; }
      }
      if (do_rho) {
        Sr[0] += wg[0] * bp.su; Sr[1] += wg[1] * bp.su;
        Sx[0] += wg[0] * bp.dsu; Sx[1] += wg[1] * bp.dsu; Sx[2] += wg[2] * bp.dsu;
      }
    }
    /* outer accumulation: Psi_t = g_p h_pq, t in hierarchical (degree, p) order:
       t: 0:(0,0) 1:(0,1) 2:(1,0) 3:(0,2) 4:(1,1) 5:(2,0) 6:(0,3) 7:(1,2) 8:(2,1)
          9:(3,0) 10:(0,4) 11:(1,3) 12:(2,2) 13:(3,1) 14:(4,0)                  */
    const double* wh = R.wbh[j];
    if (do_dd) {
      int t = 0;
      for (int n = 0; n <= 4; ++n)
        for (int p = 0; p <= n; ++p, ++t) mC[t] += wh[t] * Sc[p];
      t = 0;
      for (int n = 0; n <= 3; ++n)
        for (int p = 0; p <= n; ++p, ++t) {
          Q[t] += wh[t] * Sq[0][p];
          Q[10 + t] += wh[t] * Sq[1][p];
        }
    }
    if (do_rho) {
      mR[0] += wh[0] * Sr[0]; mR[1] += wh[1] * Sr[0]; mR[2] += wh[2] * Sr[1];
      int t = 0;
      for (int n = 0; n <= 2; ++n)
        for (int p = 0; p <= n; ++p, ++t) mRx[t] += wh[t] * Sx[p];
    }
  }
}

/* ------------------------------------------------------------------------- *
 * phase B: generation term of band k at the points of the g-rule.
 * ------------------------------------------------------------------------- */
struct SbBandsAtPoint {
  double u[SB_MAX_BANDS], du[SB_MAX_BANDS], w[SB_MAX_BANDS], u0[SB_MAX_BANDS];
};

SB_HD int sb_gen_sign(const sb_model& M, int p, int k) {
  if (k == M.proc_dst[p]) return +1;
  if (k == M.proc_src[p]) return -(M.band_sign[M.proc_dst[p]] * M.band_sign[M.proc_src[p]]);
  return 0;
}

/* Shockley-Read kernel G = expm1(x) u_R F c and its partials
 * (electro_optical_process.py:294-312), hand-derived.                     */
SB_HD void sb_sr_trap(const sb_model& M, const double* tp, int p, double coeff,
                      const SbBandsAtPoint& B, double scale,
                      double& g, double& dphi, double* dw) {
  const int IB = M.proc_trap[p];
  const int RB = (M.proc_dst[p] == IB) ? M.proc_src[p] : M.proc_dst[p];
  const double beta = 1.0 / tp[SB_TP_KT];
  const double sR = (double)M.band_sign[RB];
  const double N_I = tp[SB_TP_BAND0 + 5 * IB + SB_TPB_N];
  const bool same = (M.band_sign[IB] == M.band_sign[RB]);
  const double F = same ? (N_I - B.u[IB]) : B.u[IB];
  const double dF = same ? -B.du[IB] : B.du[IB];
  const double x = sR * (B.w[RB] - B.w[IB]) * beta;
  const double ex = exp(x), em = expm1(x);
  const double cs = coeff * scale;
  g += cs * em * B.u[RB] * F;
  dphi += cs * em * (B.du[RB] * F + B.u[RB] * dF);
  dw[RB] += cs * F * (ex * sR * beta * B.u[RB] + em * B.du[RB]);
  dw[IB] += cs * B.u[RB] * (-ex * sR * beta * F + em * dF);
}

/* g_k and its partials w.r.t. phi and every w_l at one point */
SB_HD void sb_generation(const sb_model& M, const double* tp, int k,
                         const SbBandsAtPoint& B, const double* phi_clip,
                         double& g, double& dphi, double* dw) {
  g = 0.0; dphi = 0.0;
  for (int l = 0; l < SB_MAX_BANDS; ++l) dw[l] = 0.0;
  const double beta = 1.0 / tp[SB_TP_KT];
  for (int p = 0; p < M.n_procs; ++p) {
    const int sgi = sb_gen_sign(M, p, k);
    if (!sgi) continue;
    const double sg = (double)sgi;
    const double* pp = tp + SB_TP_PROC0 + 8 * p;
    const int dst = M.proc_dst[p], src = M.proc_src[p];
    switch (M.proc_kind[p]) {
      case SB_PROC_SRH: {
        const double D = (B.u[dst] + pp[SB_TPP_P2]) * pp[SB_TPP_P1]
                       + (B.u[src] + pp[SB_TPP_P3]) * pp[SB_TPP_P0];
        const double r = (B.u[src] * B.u[dst] - B.u0[src] * B.u0[dst]) / D;
        const double dr_dst = (B.u[src] - r * pp[SB_TPP_P1]) / D * B.du[dst];
        const double dr_src = (B.u[dst] - r * pp[SB_TPP_P0]) / D * B.du[src];
        g -= sg * r;
        dw[dst] -= sg * dr_dst;
        dw[src] -= sg * dr_src;
        dphi -= sg * (dr_dst + dr_src);
      } break;
      case SB_PROC_SR_TRAP:
        sb_sr_trap(M, tp, p, pp[SB_TPP_P0], B, sg, g, dphi, dw);
        break;
      case SB_PROC_RAD_BB: {
        double S = 0.0;
        for (int f = 0; f < M.n_fields; ++f) S += phi_clip[f] * pp[SB_TPP_FIELD0 + f];
        g += sg * S;
        if (M.proc_radiative[p]) {
          const double x = beta * (B.w[dst] - B.w[src]);
          const double ex = exp(x);
          g -= sg * pp[SB_TPP_P0] * expm1(x);
          dw[dst] -= sg * pp[SB_TPP_P0] * beta * ex;
          dw[src] += sg * pp[SB_TPP_P0] * beta * ex;
        }
      } break;
      case SB_PROC_RAD_IB: {
        const int IB = M.proc_trap[p];
        const int RB = (dst == IB) ? src : dst;
        const double N_I = tp[SB_TP_BAND0 + 5 * IB + SB_TPB_N];
        const bool same = (M.band_sign[IB] == M.band_sign[RB]);
        const double Fa = same ? B.u[IB] : (N_I - B.u[IB]);
        const double dFa = same ? B.du[IB] : -B.du[IB];
        double S = 0.0;
        for (int f = 0; f < M.n_fields; ++f) S += phi_clip[f] * pp[SB_TPP_FIELD0 + f];
        g += sg * S * Fa;
        dw[IB] += sg * S * dFa;
        dphi += sg * S * dFa;
        if (M.proc_radiative[p])
          sb_sr_trap(M, tp, p, pp[SB_TPP_P0], B, sg, g, dphi, dw);
      } break;
      default: break;
    }
  }
}

/* Moments of g_k, dg_k/dphi, dg_k/dw_l over the g-rule.
 * dgv[p][3]: P1 nodal values of (phi | delta_w_l) ; w0[l]: base_w of the cell.
 * out: mG[0..2] (P1), mG[3..8] (dphi, P2), mG[9+6l ..] (dw_l, P2)            */
SB_HD void sb_phaseB_band(const SbTables& T, const sb_model& M, const double* tp, int k,
                          const double (*dgv)[3], const double* w0,
                          const double* thmeq6, const double* phi_opt, int64_t field_stride,
                          double* mG) {
  const int nb = M.n_bands;
  for (int t = 0; t < SB_NG; ++t) mG[t] = 0.0;
  const double beta = 1.0 / tp[SB_TP_KT];
  for (int n = 0; n < T.ng; ++n) {
    const double X = T.gX[n], Y = T.gY[n];
    const double l0 = 1.0 - X - Y;
    const double phi = dgv[0][0] * l0 + dgv[0][1] * X + dgv[0][2] * Y;
    const double* n2 = T.gp2[n];
    double phi_te = 0.0;
    if (thmeq6) for (int a = 0; a < 6; ++a) phi_te += thmeq6[a] * n2[a];
    SbBandsAtPoint B;
    for (int l = 0; l < nb; ++l) {
      const double* bp = tp + SB_TP_BAND0 + 5 * l;
      B.w[l] = w0[l] + dgv[1 + l][0] * l0 + dgv[1 + l][1] * X + dgv[1 + l][2] * Y;
      const double s = (double)M.band_sign[l];
      const SbBandPoint q = sb_band_point(M.band_kind[l], s, 1, beta, bp[SB_TPB_E],
                                          bp[SB_TPB_N], 1.0, 1.0, 1.0, phi + B.w[l]);
      B.u[l] = q.u; B.du[l] = q.du;
      const SbBandPoint q0 = sb_band_point(M.band_kind[l], s, 1, beta, bp[SB_TPB_E],
                                           bp[SB_TPB_N], 1.0, 1.0, 1.0, phi_te);
      B.u0[l] = q0.u;
    }
    double phi_clip[SB_MAX_FIELDS];
    for (int f = 0; f < M.n_fields; ++f) {
      const double* pf = phi_opt + (int64_t)f * field_stride;
      double v = 0.0;
      for (int a = 0; a < 6; ++a) v += pf[a] * n2[a];
      phi_clip[f] = v < 0.0 ? 0.0 : v;      /* optical.py:201-203 */
    }
    double g, dphi, dw[SB_MAX_BANDS];
    sb_generation(M, tp, k, B, phi_clip, g, dphi, dw);
    const double* wp = T.gwpsi[n];
    for (int t = 0; t < 3; ++t) mG[t] += wp[t] * g;
    for (int t = 0; t < 6; ++t) mG[3 + t] += wp[t] * dphi;
    for (int l = 0; l < nb; ++l)
      for (int t = 0; t < 6; ++t) mG[9 + 6 * l + t] += wp[t] * dw[l];
  }
}

/* ------------------------------------------------------------------------- *
 * phase C: per-cell (warp) scratch and scatter
 * ------------------------------------------------------------------------- */
struct SbWarpScratch {
  double W[21];
  double GC[12][2][6];   /* sum_b G_ab C[m][b][s]                   */
  double Z[12][2][6];    /* sum_s W_rs GC[m][a][s]                  */
  double V[2][6][3];     /* sum_t G21[r][l][t] Q_a[t]               */
  double zr[2][6];       /* sum_s W_rs gc[a][s]                     */
  double gc[2][6];       /* Dubiner-P2 coefficients of G jhat       */
  double GM[12][12];     /* sum_ab G_ab Mhat[i][j][ab] (iad applied later) */
  double ul[SB_MAXL];    /* local dofs                              */
  double bl[SB_MAXL];    /* local residual                          */
  double gl[SB_MAXL];    /* Dirichlet lift values u_i - bc_i        */
  double lift[15][9];    /* -A_ij g_j per (row of the phase, constrained facet dof j):
                            summed in fixed order => bit-reproducible   */
};

/* everything a cell needs for scatter */
struct SbCellCtx {
  int64_t c;
  int L;
  const int32_t* dofs;       /* [L] global dofs of this cell          */
  const uint8_t* rank_own;   /* [L]                                   */
  const uint8_t* rank_fac;   /* [3][L]                                */
  const int64_t* rowptr;
  uint64_t bcmask_lo;        /* constrained local dofs, bits 0..63    */
  uint32_t bcmask_hi;        /* bits 64..74                           */
  int row0;                  /* first local row of the current phase  */
  double* vals;              /* may be NULL                           */
  double* b;                 /* may be NULL                           */
};

SB_HD bool sb_is_bc(const SbCellCtx& cx, int l) {
  return l < 64 ? ((cx.bcmask_lo >> l) & 1ull) : ((cx.bcmask_hi >> (l - 64)) & 1u);
}

#if defined(__CUDA_ARCH__) || defined(SB_EMUL_BUILD)
#define SB_ATOMIC_ADD(ptr, v) atomicAdd((ptr), (v))
#else
#define SB_ATOMIC_ADD(ptr, v) (*(ptr) += (v))
#endif

/* write one structurally non-zero Jacobian entry (local i, j) = val with the
 * element-wise symmetric Dirichlet rule.               */
SB_HD void sb_emit(const SbCellCtx& cx, SbWarpScratch& S, int i, int j, double val) {
  if (cx.bcmask_lo | cx.bcmask_hi) {
    if (sb_is_bc(cx, i)) {
      val = (i == j) ? 1.0 : 0.0;
    } else if (sb_is_bc(cx, j)) {
      /* constrained dofs are BDM facet dofs: slot = facet-dof number 0..8 */
      S.lift[i - cx.row0][(j % 15) - 3] = -val * S.gl[j];
      val = 0.0;
    }
  }
  if (!cx.vals) return;
  const int fi = sb_local_facet(i);
  const int rk = (fi < 0) ? cx.rank_own[j] : cx.rank_fac[fi * cx.L + j];
  if (rk == 0xFF) return;                 /* not in a compacted pattern */
  double* dst = cx.vals + cx.rowptr[cx.dofs[i]] + rk;
  /* two cells add into the same slot only for (facet dof, facet dof) pairs of
     one facet; two addends onto a zeroed slot commute -> deterministic       */
  if (fi >= 0 && sb_local_facet(j) == fi) SB_ATOMIC_ADD(dst, val);
  else *dst = val;
}

/* per-cell, band-independent prep: GC, GM, local dofs */
SB_HD void sb_prep_cell(const SbTables& T, const SbCellGeom& g, SbWarpScratch& S,
                        int lane, int nl) {
  for (int e = lane; e < 144; e += nl) {
    const int m = e / 12, rem = e % 12, a = rem / 6, s = rem % 6;
    const double Ga0 = a == 0 ? g.G11 : g.G12, Ga1 = a == 0 ? g.G12 : g.G22;
    S.GC[m][a][s] = Ga0 * T.bdmC[m][0][s] + Ga1 * T.bdmC[m][1][s];
  }
  for (int e = lane; e < 144; e += nl) {
    const int i = e / 12, j = e % 12;
    S.GM[i][j] = g.G11 * T.Mhat[i][j][0] + g.G12 * T.Mhat[i][j][1] + g.G22 * T.Mhat[i][j][2];
  }
}

/* gc[a][r] = sum_m j_m GC[m][a][r] for flux dofs jl[12]  (needs GC) */
SB_HD void sb_flux_coeffs(const SbWarpScratch& S, const double* jl, double (*gc)[6],
                          int lane, int nl) {
  for (int e = lane; e < 12; e += nl) {
    const int a = e / 6, r = e % 6;
    double v = 0.0;
    for (int m = 0; m < 12; ++m) v += jl[m] * S.GC[m][a][r];
    gc[a][r] = v;
  }
}

SB_HD int sb_sym21(int r, int s) {      /* packed index of (r<=s) in 6x6 */
  if (r > s) { const int t = r; r = s; s = t; }
  return r * 6 - (r * (r - 1)) / 2 + (s - r);
}

#endif /* SB_PDD_DEVICE_CUH */
