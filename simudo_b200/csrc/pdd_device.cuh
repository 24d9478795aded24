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
 *   phase C  (same thread): moments x constant expansion tensors (warp-uniform
 *            constant-bank operands) x cell geometry -> structurally non-zero
 *            entries of the element tensor, Dirichlet application, scatter
 *            into the fixed CSR.  No shared memory, no barriers: 32 cells of a
 *            warp progress in lock-step, so every table operand is uniform.
 *
 * The same source also compiles for the CPU under tests/emul (CUDA shim) for
 * debugging of the kernel logic; that harness is test infrastructure only.
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
 * phase C: scatter context of one cell (one thread)
 * ------------------------------------------------------------------------- */
struct SbIO {
  int L;
  const int32_t* dofs;       /* [L] global dofs of this cell                */
  const uint8_t* rank_own;   /* [L]                                         */
  const uint8_t* rank_fac;   /* [3][L]                                      */
  const int64_t* rowptr;
  const double* u;
  const double* bc_values;
  const int32_t* bcidx;      /* [L] index into bc_values | -1 (special cells) */
  uint64_t mlo;              /* constrained local dofs, bits 0..63          */
  uint32_t mhi;              /* bits 64..74                                 */
  double* vals;              /* may be NULL                                 */
  double* b;                 /* may be NULL                                 */
};

SB_HD bool sb_is_bc(const SbIO& io, int l) {
  return l < 64 ? ((io.mlo >> l) & 1ull) : ((io.mhi >> (l - 64)) & 1u);
}
SB_HD bool sb_special(const SbIO& io) { return (io.mlo | io.mhi) != 0; }
/* Dirichlet lift value u_j - bc_j of a constrained local dof */
SB_HD double sb_lift_value(const SbIO& io, int j) {
  return io.u[io.dofs[j]] - io.bc_values[io.bcidx[j]];
}

#if defined(__CUDA_ARCH__) || defined(SB_EMUL_BUILD)
#define SB_ATOMIC_ADD(ptr, v) atomicAdd((ptr), (v))
#else
#define SB_ATOMIC_ADD(ptr, v) (*(ptr) += (v))
#endif

/* Write one structurally non-zero Jacobian entry (local i, j) = val with the
 * element-wise symmetric Dirichlet rule; Fi is the
 * residual accumulator of row i (receives the x0 lifting, in program order,
 * hence bit-reproducible).                                                   */
SB_HD void sb_emit(const SbIO& io, int i, int j, double val, double& Fi) {
  if (sb_special(io)) {
    if (sb_is_bc(io, i)) {
      val = (i == j) ? 1.0 : 0.0;
    } else if (sb_is_bc(io, j)) {
      Fi -= val * sb_lift_value(io, j);
      val = 0.0;
    }
  }
  if (!io.vals) return;
  const int fi = sb_local_facet(i);
  const int rk = (fi < 0) ? io.rank_own[j] : io.rank_fac[fi * io.L + j];
  if (rk == 0xFF) return;                 /* not in a compacted pattern */
  double* dst = io.vals + io.rowptr[io.dofs[i]] + rk;
  /* two cells add into the same slot only for (facet dof, facet dof) pairs of
     one facet; two addends onto a zeroed slot commute -> deterministic       */
  if (fi >= 0 && sb_local_facet(j) == fi) SB_ATOMIC_ADD(dst, val);
  else *dst = val;
}

/* final residual value of local row i */
SB_HD void sb_put_row(const SbIO& io, int i, double F) {
  if (!io.b) return;
  if (sb_special(io) && sb_is_bc(io, i)) F = sb_lift_value(io, i);
  double* dst = io.b + io.dofs[i];
  if (sb_local_facet(i) >= 0) SB_ATOMIC_ADD(dst, F); else *dst = F;
}

SB_HD int sb_pair(int i, int m) {       /* packed index of (i <= m) in 12 x 12 */
  return i * 12 - (i * (i - 1)) / 2 + (m - i);
}

/* natural-BC values of this cell for rows [lo, lo+n): F[row-lo] += value */
SB_HD void sb_add_natbc(const int32_t* nat_row, const double* nat_val, int b0, int b1,
                        int lo, int n, double* F) {
  for (int k = b0; k < b1; ++k) {
    const int r = nat_row[k] - lo;
    if (r >= 0 && r < n) F[r] += nat_val[k];
  }
}

/* ---- rows xi^k (BDM test functions of band k) ------------------------------- */
SB_HD void sb_rows_flux(const SbTables& T, const sb_model& M, const SbIO& io,
                        const SbCellGeom& g, int p, bool is_band, bool inside,
                        const double* mC, const double* Q,
                        const double* fl /*[12] flux dofs*/, const double* sl /*[3] scalar dofs*/,
                        const double* jumpf /*[3] or NULL*/,
                        const int32_t* nat_row, const double* nat_val, int nb0, int nb1) {
  const int r0 = sb_lbdm(p, 0);
  double F[12];
  for (int i = 0; i < 12; ++i) F[i] = 0.0;
  if (is_band && inside) {
    /* <xi_i . xi_m c>: symmetric, 78 pairs x 45 FMAs on warp-uniform constants */
    int pair = 0;
    for (int i = 0; i < 12; ++i)
      for (int m = i; m < 12; ++m, ++pair) {
        double a0 = 0.0, a1 = 0.0, a2 = 0.0;
#pragma unroll
        for (int t = 0; t < 15; ++t) {
          a0 += mC[t] * T.TB[pair][0][t];
          a1 += mC[t] * T.TB[pair][1][t];
          a2 += mC[t] * T.TB[pair][2][t];
        }
        const double v = (g.G11 * a0 + g.G12 * a1 + g.G22 * a2) * g.iad;
        F[i] += v * fl[m];
        sb_emit(io, r0 + i, r0 + m, v, F[i]);
        if (m != i) {
          F[m] += v * fl[i];
          sb_emit(io, r0 + m, r0 + i, v, F[m]);
        }
      }
    /* <div xi_i delta_w> + <xi_i . j dc/dchi lam_l> */
    for (int i = 0; i < 12; ++i) {
      for (int l = 0; l < 3; ++l) {
        double d = 0.0;
#pragma unroll
        for (int t = 0; t < 10; ++t)
          d += T.TD[i][l][0][t] * Q[t] + T.TD[i][l][1][t] * Q[10 + t];
        d *= g.iad;
        const double dv = g.sdet * T.Dhat[i][l];
        F[i] += dv * sl[l];
        sb_emit(io, r0 + i, sb_ldg(p, l), dv + d, F[i]);
        sb_emit(io, r0 + i, sb_ldg(0, l), d, F[i]);
      }
      if (jumpf) F[i] -= jumpf[i / 3] * T.trI[i % 3];
    }
  } else {
    /* polynomial mass block: Poisson <psi.E> (scale 1) or the out-of-subdomain
       penalty <j.xi> c2 (poisson_drift_diffusion1.py1:288-289)               */
    const double sc = (is_band ? M.ext_j_coef : 1.0) * g.iad;
    for (int i = 0; i < 12; ++i)
      for (int m = i; m < 12; ++m) {
        const double v = sc * (g.G11 * T.Mhat[i][m][0] + g.G12 * T.Mhat[i][m][1]
                               + g.G22 * T.Mhat[i][m][2]);
        F[i] += v * fl[m];
        sb_emit(io, r0 + i, r0 + m, v, F[i]);
        if (m != i) {
          F[m] += v * fl[i];
          sb_emit(io, r0 + m, r0 + i, v, F[m]);
        }
      }
    if (!is_band) {
      /* - <div psi_i phi> */
      for (int i = 0; i < 12; ++i)
        for (int l = 0; l < 3; ++l) {
          const double dv = -g.sdet * T.Dhat[i][l];
          F[i] += dv * sl[l];
          sb_emit(io, r0 + i, sb_ldg(0, l), dv, F[i]);
        }
    }
  }
  if (nb1 > nb0) sb_add_natbc(nat_row, nat_val, nb0, nb1, r0, 12, F);
  for (int i = 0; i < 12; ++i) sb_put_row(io, r0 + i, F[i]);
}

/* ---- rows of the DG1 test functions: w (Poisson, p = 0) or v^k (p = 1 + k) ---- *
 * coef: multiplies the moment terms (-adet for rho, -s e adet for g)
 * mom1[3]: P1 moments of the source (rho/eps | g_k)
 * momJ[q][6]: P2 moments of d source / d (phi | delta_w_q-1), q < nq          */
SB_HD void sb_rows_scalar(const SbTables& T, const sb_model& M, const SbIO& io,
                          const SbCellGeom& g, int p, bool inside,
                          const double* fl /*[12] flux dofs of p*/, const double* sl /*[3]*/,
                          double coef, const double* mom1, const double* momJ, int nq,
                          const int32_t* nat_row, const double* nat_val, int nb0, int nb1) {
  const int r0 = sb_ldg(p, 0);
  double F[3] = {0.0, 0.0, 0.0};
  if (inside) {
    for (int i = 0; i < 3; ++i) {
      for (int m = 0; m < 12; ++m) {            /* <v_i div j> */
        const double dv = g.sdet * T.Dhat[m][i];
        F[i] += dv * fl[m];
        sb_emit(io, r0 + i, sb_lbdm(p, m), dv, F[i]);
      }
      for (int t = 0; t < 3; ++t) F[i] += coef * T.G1[i][t] * mom1[t];
      for (int q = 0; q < nq; ++q)
        for (int l = 0; l < 3; ++l) {
          double v = 0.0;
#pragma unroll
          for (int t = 0; t < 6; ++t) v += T.G11[i][l][t] * momJ[6 * q + t];
          sb_emit(io, r0 + i, sb_ldg(q, l), coef * v, F[i]);
        }
    }
  } else {
    /* out-of-subdomain penalty <delta_w v> c1 (:285-286) */
    const double c1 = M.ext_w_coef * g.adet;
    for (int i = 0; i < 3; ++i)
      for (int l = 0; l < 3; ++l) {
        const double v = c1 * T.M1[i][l];
        F[i] += v * sl[l];
        sb_emit(io, r0 + i, r0 + l, v, F[i]);
      }
  }
  if (nb1 > nb0) sb_add_natbc(nat_row, nat_val, nb0, nb1, r0, 3, F);
  for (int i = 0; i < 3; ++i) sb_put_row(io, r0 + i, F[i]);
}

#endif /* SB_PDD_DEVICE_CUH */
