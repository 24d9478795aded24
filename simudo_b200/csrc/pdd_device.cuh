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
#include <string.h>
#include "../../include/simudo_b200.h"
#include "pdd_tables.h"

#if defined(__CUDACC__)
#define SB_HD __host__ __device__ __forceinline__
#define SB_D __device__ __forceinline__
#define SB_NOINLINE __host__ __device__ __noinline__
#else
#define SB_HD inline
#define SB_D inline
#define SB_NOINLINE inline
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

SB_HD constexpr int sb_sym21(int r, int s) {      /* packed index of the (r, s) pair, 6 x 6 symmetric */
  return (r <= s) ? (r * 6 - (r * (r - 1)) / 2 + (s - r)) : (s * 6 - (s * (s - 1)) / 2 + (r - s));
}

/* Per (cell, band) constants hoisted out of the quadrature loops. */
struct SbBandK {
  int kind, const_mob;
  double s, beta, sbeta;   /* sign, 1/kT, s/kT                         */
  double E, N, mob;
  double N_eps;            /* N / eps                                   */
  double cinv;             /* dd_scale / (mob N)                        */
  double dd_scale;
};

SB_HD SbBandK sb_band_consts(const sb_model& M, const double* tp, int k) {
  const double* bp = tp + SB_TP_BAND0 + 5 * k;
  SbBandK K;
  K.kind = M.band_kind[k];
  K.const_mob = M.band_const_mobility[k];
  K.s = (double)M.band_sign[k];
  K.beta = 1.0 / tp[SB_TP_KT];
  K.sbeta = K.s * K.beta;
  K.E = bp[SB_TPB_E];
  K.N = bp[SB_TPB_N];
  K.mob = bp[SB_TPB_MOBILITY];
  K.N_eps = K.N / tp[SB_TP_EPS];
  K.dd_scale = M.dd_scale;
  K.cinv = M.dd_scale / (K.mob * K.N);
  return K;
}

/* IB with filling-dependent mobility mu = mu0 [(1-f) - 1/2 f^0.33 (1-f)^1.33]
 * (poisson_drift_diffusion1.py1:228-231); out of line: pow() is bulky and this
 * is the rarely used model (use_constant_mobility = False)                  */
struct SbCdc { double c, dc; };     /* by value: out-pointers would force a stack round trip */
SB_NOINLINE SbCdc sb_ib_fill_mobility(double mob, double N, double dd_scale, double sbeta,
                                      double f, double omf) {
  const double df = -sbeta * f * omf;       /* df/dchi */
  const double fa = pow(f, 0.33), ob = pow(omf, 1.33);
  const double mu = mob * (omf - 0.5 * fa * ob);
  const double dmu = mob * (-df - 0.5 * (0.33 * fa / f * df * ob
                                          + fa * 1.33 * ob / omf * (-df)));
  SbCdc r;
  r.c = dd_scale / (mu * N * f);
  r.dc = -r.c * (dmu / mu + df / f);
  return r;
}

/* exp(x) and exp(-x) from one range reduction: x = k ln2 + r, |r| <= ln2/2,
 * exp(+-r) = E(r^2) +- r O(r^2) (Taylor to degree 13, truncation 4e-18 relative;
 * the even and odd chains are independent, half the dependency depth of a
 * Horner exp), scaled by 2^(+-k) built from the exponent bits.  x is clamped to
 * +-700 (no inf / denormal paths).  Replaces exp() + a division per point.  */
SB_HD double sb_pow2i(int k) {
#ifdef __CUDA_ARCH__
  return __hiloint2double((1023 + k) << 20, 0);
#else
  const uint64_t u = (uint64_t)(1023 + k) << 52;
  double d;
  memcpy(&d, &u, sizeof d);
  return d;
#endif
}
/* Coefficients of the exp kernels.  On the device they sit in the constant bank so
 * that every DFMA of the chains takes its coefficient as a c[bank][offset] operand;
 * as inline literals each 64-bit coefficient cost two extra move instructions per
 * use (the kernels run at the register limit, ptxas rematerialises them), which
 * tripled the instruction count of the chains.                                   */
#define SB_EXPC_VALUES { \
  1.4426950408889634074, -6.93147180369123816490e-01, -1.90821492927058770002e-10, \
  1.0 / 479001600.0, 1.0 / 3628800.0, 1.0 / 40320.0, 1.0 / 720.0, 1.0 / 24.0, \
  1.0 / 6227020800.0, 1.0 / 39916800.0, 1.0 / 362880.0, 1.0 / 5040.0, 1.0 / 120.0, 1.0 / 6.0 }
#if defined(__CUDACC__) && !defined(SB_EMUL_BUILD)
static __constant__ double c_sb_expc[14] = SB_EXPC_VALUES;
#endif
static const double h_sb_expc[14] = SB_EXPC_VALUES;
#ifdef __CUDA_ARCH__
#define SB_EXPC(i) c_sb_expc[i]
#else
#define SB_EXPC(i) h_sb_expc[i]
#endif

/* range reduction shared by the exp kernels: x (clamped to +-700) = k ln2 + r */
SB_HD double sb_exp_reduce(double x, int& k) {
  x = fmin(fmax(x, -700.0), 700.0);
  const double magic = 6755399441055744.0;            /* 1.5 * 2^52 */
  const double t = fma(x, SB_EXPC(0), magic);
#ifdef __CUDA_ARCH__
  k = __double2loint(t);
#else
  uint64_t tb;
  memcpy(&tb, &t, sizeof tb);
  k = (int)(uint32_t)(tb & 0xFFFFFFFFu);
#endif
  const double kf = t - magic;
  double r = fma(kf, SB_EXPC(1), x);
  r = fma(kf, SB_EXPC(2), r);
  return r;
}

SB_HD void sb_exp_pair(double x, double& ep, double& em) {
  int k;
  const double r = sb_exp_reduce(x, k);
  const double z = r * r;
  double E = SB_EXPC(3), O = SB_EXPC(8);
  E = fma(E, z, SB_EXPC(4));  O = fma(O, z, SB_EXPC(9));
  E = fma(E, z, SB_EXPC(5));  O = fma(O, z, SB_EXPC(10));
  E = fma(E, z, SB_EXPC(6));  O = fma(O, z, SB_EXPC(11));
  E = fma(E, z, SB_EXPC(7));  O = fma(O, z, SB_EXPC(12));
  E = fma(E, z, 0.5);         O = fma(O, z, SB_EXPC(13));
  E = fma(E, z, 1.0);         O = fma(O, z, 1.0);
  const double ro = r * O;
  ep = (E + ro) * sb_pow2i(k);
  em = (E - ro) * sb_pow2i(-k);
}

/* exp(x) alone, same chains */
SB_HD double sb_exp(double x) {
  int k;
  const double r = sb_exp_reduce(x, k);
  const double z = r * r;
  double E = SB_EXPC(3), O = SB_EXPC(8);
  E = fma(E, z, SB_EXPC(4));  O = fma(O, z, SB_EXPC(9));
  E = fma(E, z, SB_EXPC(5));  O = fma(O, z, SB_EXPC(10));
  E = fma(E, z, SB_EXPC(6));  O = fma(O, z, SB_EXPC(11));
  E = fma(E, z, SB_EXPC(7));  O = fma(O, z, SB_EXPC(12));
  E = fma(E, z, 0.5);         O = fma(O, z, SB_EXPC(13));
  E = fma(E, z, 1.0);         O = fma(O, z, 1.0);
  return fma(r, O, E) * sb_pow2i(k);
}

/* expm1(x) with the same reduction: expm1(x) = 2^k expm1(r) + (2^k - 1), where
 * expm1(r) = r O(r^2) + r^2 E1(r^2) has no cancellation (E = 1 + z E1); for k = 0 this
 * is the series itself, so the value near thermal equilibrium (x -> 0) keeps full
 * relative accuracy; branch-free.  x is clamped to [-700, 700].                   */
SB_HD double sb_expm1(double x) {
  int k;
  const double r = sb_exp_reduce(x, k);
  const double z = r * r;
  double E1 = SB_EXPC(3), O = SB_EXPC(8);
  E1 = fma(E1, z, SB_EXPC(4));  O = fma(O, z, SB_EXPC(9));
  E1 = fma(E1, z, SB_EXPC(5));  O = fma(O, z, SB_EXPC(10));
  E1 = fma(E1, z, SB_EXPC(6));  O = fma(O, z, SB_EXPC(11));
  E1 = fma(E1, z, SB_EXPC(7));  O = fma(O, z, SB_EXPC(12));
  E1 = fma(E1, z, 0.5);         O = fma(O, z, SB_EXPC(13));
  O = fma(O, z, 1.0);
  const double em_r = fma(r, O, z * E1);
  const double s = sb_pow2i(k);
  return fma(s, em_r, s - 1.0);
}

/* 1 / d for d in [1, 1e305] (d = 1 + exp(x), x clamped): hardware approximation plus two
 * Newton steps, within 1 ulp, and no special-case branch -- the IEEE division's slow path
 * would end the basic block and with it the interleaving of neighbouring points.     */
SB_HD double sb_rcp_pos(double d) {
#ifdef __CUDA_ARCH__
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
  double e = fma(-d, y, 1.0);
  y = fma(y, e, y);
  e = fma(-d, y, 1.0);
  return fma(y, e, y);
#else
  return 1.0 / d;
#endif
}

/* Pointwise band quantities at chi = e phi + w (band statistics
 * poisson_drift_diffusion1.py1:170-245):
 *   su = s u / eps, dsu = d(su)/dchi, c = dd_scale / (mu u), dc = dc/dchi   */
/* kind / const_mob passed separately: callers that hold K in shared memory pass the
 * warp-uniform model values so that the branch stays uniform and the unrolled point
 * loops remain straight-line code (independent exp chains interleave) */
SB_HD void sb_band_point_kind(const SbBandK& K, int kind, int const_mob, double chi,
                              double& su, double& dsu, double& c, double& dc) {
  double e, einv;
  sb_exp_pair(K.sbeta * (chi - K.E), e, einv);
  if (kind == SB_BAND_NONDEGENERATE) {
    const double ue = K.N_eps * einv;         /* u / eps */
    su = K.s * ue;
    dsu = -K.beta * ue;
    c = K.cinv * e;
    dc = K.sbeta * c;
  } else {
    const double f = 1.0 / (1.0 + e);         /* IEEE division: in the moment kernel its slow-path
                                                 branch also bounds the interleaving, measured
                                                 faster (fewer spills) than sb_rcp_pos here */
    const double omf = e * f;                 /* 1 - f */
    su = K.s * K.N_eps * f;
    dsu = -K.beta * K.N_eps * f * omf;
    if (const_mob) {
      c = K.cinv * (1.0 + e);
      dc = K.cinv * K.sbeta * e;
    } else {
      const SbCdc r = sb_ib_fill_mobility(K.mob, K.N, K.dd_scale, K.sbeta, f, omf);
      c = r.c; dc = r.dc;
    }
  }
}
SB_HD void sb_band_point(const SbBandK& K, double chi,
                         double& su, double& dsu, double& c, double& dc) {
  sb_band_point_kind(K, K.kind, K.const_mob, chi, su, dsu, c, dc);
}

/* carrier density and its chi-derivative only (generation term) */
/* branch-free for both kinds (selects): the bands of a point form one basic block */
SB_HD void sb_band_u_kind(const SbBandK& K, int kind, double chi, double& u, double& du) {
  double e, einv;
  sb_exp_pair(K.sbeta * (chi - K.E), e, einv);
  const double f = sb_rcp_pos(1.0 + e);
  const bool nd = (kind == SB_BAND_NONDEGENERATE);
  u = K.N * (nd ? einv : f);
  du = -K.sbeta * u * (nd ? 1.0 : e * f);
}
SB_HD void sb_band_u(const SbBandK& K, double chi, double& u, double& du) {
  sb_band_u_kind(K, K.kind, chi, u, du);
}

/* ------------------------------------------------------------------------- *
 * phase A: moments of one band over the collapsed rule of R.
 *   chi(X,Y) = chi0 + chix X + chiy Y                      (P1, incl. base_w)
 *   gc[a][r]: Dubiner-P2 coefficients of g^a = (J^T J jhat)^a
 * outputs (overwritten unless noted):
 *   mC[15], Q[2][10]          if do_dd
 *   mRx[6], mR[3] (+=)        if do_rho
 * MM > 0: compile-time number of points per axis (inner loop unrolled, the
 * independent exp chains of neighbouring points interleave); MM = 0: R.m.
 * ------------------------------------------------------------------------- */
#ifndef SB_TILE
#define SB_TILE 128
#endif
template <int MM, int MODE = 0>
SB_HD void sb_phaseA_band(const SbRule& R, bool do_rho_, bool do_dd_, const SbBandK& K,
                          double chi0, double chix, double chiy, const double (*gc)[6],
                          double* mC, double* Q, double* mR, double* mRx,
                          double* qs = nullptr /* Q accumulators in shared memory: Q[t] at
                                                  qs[t * SB_TILE] (then Q is not touched) */) {
  const int m = MM > 0 ? MM : R.m;
  /* MODE 1 / 2: rho and dd both on, kind fixed (non-degenerate / intermediate with constant
     mobility): the unrolled loop over the 11 points of a row is straight-line code, so the
     independent exp chains of neighbouring points interleave (measured 3.19 -> 2.77 ms per
     1M cells; with the run-time branches every point was its own basic block and the warp
     ran one dependent chain at a time).  MODE 0: everything at run time. */
  const bool do_rho = MODE ? true : do_rho_;
  const bool do_dd = MODE ? true : do_dd_;
  const int kind = MODE == 1 ? SB_BAND_NONDEGENERATE : (MODE == 2 ? SB_BAND_INTERMEDIATE : K.kind);
  const int cmob = MODE == 2 ? 1 : K.const_mob;       /* MODE 2: constant-mobility IB */
  if (do_dd) {
#pragma unroll
    for (int t = 0; t < SB_NMC; ++t) mC[t] = 0.0;
#pragma unroll
    for (int t = 0; t < SB_NQ; ++t) { if (qs) qs[t * SB_TILE] = 0.0; else Q[t] = 0.0; }
  }
  if (do_rho) {
#pragma unroll
    for (int t = 0; t < SB_NRX; ++t) mRx[t] = 0.0;
  }
#pragma unroll 1
  for (int j = 0; j < m; ++j) {
    const double bj = R.b[j];
    const double omb = 1.0 - bj;
    /* row-wise pieces of g^a: H[a][p] = sum_q gc[a][(p,q)] h_pq(b_j) */
    double H[2][3];
    if (do_dd) {
      const double* h = R.h6[j];   /* order: (0,0) (0,1) (1,0) (0,2) (1,1) (2,0) */
#pragma unroll
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
#pragma unroll
    for (int i = 0; i < m; ++i) {
      const double chi = chij + chis * R.a[i];
      double su, dsu, c, dc;
      sb_band_point_kind(K, kind, cmob, chi, su, dsu, c, dc);
      const double* wg = R.wag[i];
      if (do_dd) {
        const double* g3 = R.g3[i];
        const double q0 = dc * (g3[0] * H[0][0] + g3[1] * H[0][1] + g3[2] * H[0][2]);
        const double q1 = dc * (g3[0] * H[1][0] + g3[1] * H[1][1] + g3[2] * H[1][2]);
#pragma unroll
        for (int p = 0; p < 5; ++p) Sc[p] += wg[p] * c;
#pragma unroll
        for (int p = 0; p < 4; ++p) { Sq[0][p] += wg[p] * q0; Sq[1][p] += wg[p] * q1; }
      }
      if (do_rho) {
        Sr[0] += wg[0] * su; Sr[1] += wg[1] * su;
        Sx[0] += wg[0] * dsu; Sx[1] += wg[1] * dsu; Sx[2] += wg[2] * dsu;
      }
    }
    /* outer accumulation: Psi_t = g_p h_pq, t in hierarchical (degree, p) order:
       t: 0:(0,0) 1:(0,1) 2:(1,0) 3:(0,2) 4:(1,1) 5:(2,0) 6:(0,3) 7:(1,2) 8:(2,1)
          9:(3,0) 10:(0,4) 11:(1,3) 12:(2,2) 13:(3,1) 14:(4,0)                  */
    const double* wh = R.wbh[j];
    if (do_dd) {
      int t = 0;
#pragma unroll
      for (int n = 0; n <= 4; ++n)
#pragma unroll
        for (int p = 0; p <= n; ++p, ++t) mC[t] += wh[t] * Sc[p];
      t = 0;
#pragma unroll
      for (int n = 0; n <= 3; ++n)
#pragma unroll
        for (int p = 0; p <= n; ++p, ++t) {
          if (qs) {
            qs[t * SB_TILE] += wh[t] * Sq[0][p];
            qs[(10 + t) * SB_TILE] += wh[t] * Sq[1][p];
          } else {
            Q[t] += wh[t] * Sq[0][p];
            Q[10 + t] += wh[t] * Sq[1][p];
          }
        }
    }
    if (do_rho) {
      mR[0] += wh[0] * Sr[0]; mR[1] += wh[1] * Sr[0]; mR[2] += wh[2] * Sr[1];
      int t = 0;
#pragma unroll
      for (int n = 0; n <= 2; ++n)
#pragma unroll
        for (int p = 0; p <= n; ++p, ++t) mRx[t] += wh[t] * Sx[p];
    }
  }
}

/* ------------------------------------------------------------------------- *
 * phase B: generation terms of all bands at the points of the g-rule.
 *
 * The process list is data (sb_model in the constant bank), so band indices are
 * only known at run time.  Everything per point nevertheless stays in registers:
 * arrays are indexed with compile-time indices only, a run-time band index goes
 * through sb_pick / sb_bump (select chains over the NB <= 4 candidates; the index
 * is warp-uniform).  Dynamically indexed thread-local arrays put the first
 * version of this kernel at 5.6 GB of local-memory traffic per 320 k cells.
 * ------------------------------------------------------------------------- */
/* ---- static process lists ---------------------------------------------------------------- *
 * The process list of a model is data, but a handful of lists cover the shipped device models
 * (four-layer IB cell: fourlayer.py:537-564; pn diode with SRH: pn_diode.py:161-164).  For
 * those the generation kernel is instantiated with the list as a type: SbCList<...>::operator[]
 * folds to a literal once the process loop is unrolled, so sb_pick / sb_bump / sb_gen_sign and
 * the kind dispatch disappear and the processes of a point become one straight-line block
 * (the interpreter was ~35 % of the kernel's issue slots, profiles/r2_s2_hot_lines.txt).
 * Any other list runs the interpreter (MV = sb_model).                                      */
template <int... V>
struct SbCList {
  SB_HD constexpr int operator[](int i) const {
    constexpr int v[sizeof...(V)] = {V...};
    return v[i];
  }
};
template <class MV> struct SbModelTraits { static constexpr int unroll_procs = 1; };
SB_HD const sb_model& sb_dyn(const sb_model& M) { return M; }
/* four-layer IB cell: bands CB (-), IB (-), VB (+); opt_ci, opt_iv, opt_cv, nr_top, nr_bottom */
struct SbSigIB5 {
  const sb_model& dyn;
  static constexpr int n_procs = 5, n_fields = 3;
  SbCList<SB_BAND_NONDEGENERATE, SB_BAND_INTERMEDIATE, SB_BAND_NONDEGENERATE> band_kind;
  SbCList<-1, -1, 1> band_sign;
  SbCList<SB_PROC_RAD_IB, SB_PROC_RAD_IB, SB_PROC_RAD_BB, SB_PROC_SR_TRAP, SB_PROC_SR_TRAP> proc_kind;
  SbCList<0, 1, 0, 0, 1> proc_dst;
  SbCList<1, 2, 2, 1, 2> proc_src;
  SbCList<1, 1, -1, 1, 1> proc_trap;
  SbCList<1, 1, 1, 1, 1> proc_radiative;
};
/* pn diode: bands CB (-), VB (+); SRH */
struct SbSigPnSRH {
  const sb_model& dyn;
  static constexpr int n_procs = 1, n_fields = 0;
  SbCList<SB_BAND_NONDEGENERATE, SB_BAND_NONDEGENERATE> band_kind;
  SbCList<-1, 1> band_sign;
  SbCList<SB_PROC_SRH> proc_kind;
  SbCList<0> proc_dst;
  SbCList<1> proc_src;
  SbCList<-1> proc_trap;
  SbCList<1> proc_radiative;
};
template <> struct SbModelTraits<SbSigIB5> { static constexpr int unroll_procs = 8; };
template <> struct SbModelTraits<SbSigPnSRH> { static constexpr int unroll_procs = 8; };
SB_HD const sb_model& sb_dyn(const SbSigIB5& M) { return M.dyn; }
SB_HD const sb_model& sb_dyn(const SbSigPnSRH& M) { return M.dyn; }
/* host: does `m` have the list of signature S? */
template <class S>
static inline bool sb_sig_matches(const sb_model& m, int nb) {
  const sb_model dummy = m;
  const S sg{dummy};
  if (m.n_bands != nb || m.n_procs != S::n_procs || m.n_fields != S::n_fields) return false;
  for (int l = 0; l < nb; ++l)
    if (m.band_kind[l] != sg.band_kind[l] || m.band_sign[l] != sg.band_sign[l]) return false;
  for (int p = 0; p < S::n_procs; ++p)
    if (m.proc_kind[p] != sg.proc_kind[p] || m.proc_dst[p] != sg.proc_dst[p] ||
        m.proc_src[p] != sg.proc_src[p] || m.proc_trap[p] != sg.proc_trap[p] ||
        (m.proc_radiative[p] != 0) != (sg.proc_radiative[p] != 0)) return false;
  return true;
}

template <int NB>
SB_HD double sb_pick(const double* a, int i) {
  double v = a[0];
#pragma unroll
  for (int l = 1; l < NB; ++l) v = (i == l) ? a[l] : v;
  return v;
}
template <int NB>
SB_HD void sb_bump(double* a, int i, double v) {
#pragma unroll
  for (int l = 0; l < NB; ++l) a[l] = (i == l) ? a[l] + v : a[l];
}

template <int NB>
struct SbBandsAtPoint {
  double u[NB > 0 ? NB : 1], du[NB > 0 ? NB : 1], w[NB > 0 ? NB : 1], u0[NB > 0 ? NB : 1];
  /* expm1(s_R (w_R - w_I) / kT) for every band R against the trap band `trap`
     (-1: none): shared by the Shockley-Read and the radiative IB process of a pair */
  double em[NB > 0 ? NB : 1];
  int trap;
};

template <class MV>
SB_HD int sb_gen_sign(const MV& M, int p, int k) {
  if (k == M.proc_dst[p]) return +1;
  if (k == M.proc_src[p]) return -(M.band_sign[M.proc_dst[p]] * M.band_sign[M.proc_src[p]]);
  return 0;
}

/* Shockley-Read kernel G = expm1(x) u_R F c and its partials
 * (electro_optical_process.py:294-312), hand-derived.                     */
template <int NB, class MV>
SB_HD void sb_sr_trap(const MV& M, const double* tp, int p, double coeff, double beta,
                      const SbBandsAtPoint<NB>& B, double& g, double& dphi, double* dw) {
  const int IB = M.proc_trap[p];
  const int RB = (M.proc_dst[p] == IB) ? M.proc_src[p] : M.proc_dst[p];
  const double sR = (double)M.band_sign[RB];
  const double N_I = tp[SB_TP_BAND0 + 5 * IB + SB_TPB_N];
  const bool same = (M.band_sign[IB] == M.band_sign[RB]);
  const double uI = sb_pick<NB>(B.u, IB), duI = sb_pick<NB>(B.du, IB);
  const double uR = sb_pick<NB>(B.u, RB), duR = sb_pick<NB>(B.du, RB);
  const double F = same ? (N_I - uI) : uI;
  const double dF = same ? -duI : duI;
  double em;
  if (B.trap == IB) {
    em = sb_pick<NB>(B.em, RB);
  } else {
    em = sb_expm1(sR * (sb_pick<NB>(B.w, RB) - sb_pick<NB>(B.w, IB)) * beta);
  }
  const double ex = em + 1.0;
  g += coeff * em * uR * F;
  dphi += coeff * em * (duR * F + uR * dF);
  sb_bump<NB>(dw, RB, coeff * F * (ex * sR * beta * uR + em * duR));
  sb_bump<NB>(dw, IB, coeff * uR * (-ex * sR * beta * F + em * dF));
}

/* One process at one point: G_p (positive = the destination band gains
 * carriers) and its partials w.r.t. phi and every w_l.  Band k then receives
 * sign(p, k) * G_p (TwoBandEOPMixin, electro_optical_process.py:193-205).    */
template <int NB, class MV>
SB_HD void sb_process_point(const MV& M, const double* tp, int p, double beta,
                            const SbBandsAtPoint<NB>& B, const double* phi_clip,
                            double& g, double& dphi, double* dw) {
  g = 0.0; dphi = 0.0;
#pragma unroll
  for (int l = 0; l < NB; ++l) dw[l] = 0.0;
  const double* pp = tp + SB_TP_PROC0 + 8 * p;
  const int dst = M.proc_dst[p], src = M.proc_src[p];
  const int kind = M.proc_kind[p];
  if (kind == SB_PROC_SRH) {
    const double ud = sb_pick<NB>(B.u, dst), us = sb_pick<NB>(B.u, src);
    const double D = (ud + pp[SB_TPP_P2]) * pp[SB_TPP_P1] + (us + pp[SB_TPP_P3]) * pp[SB_TPP_P0];
    const double r = (us * ud - sb_pick<NB>(B.u0, src) * sb_pick<NB>(B.u0, dst)) / D;
    const double dr_dst = (us - r * pp[SB_TPP_P1]) / D * sb_pick<NB>(B.du, dst);
    const double dr_src = (ud - r * pp[SB_TPP_P0]) / D * sb_pick<NB>(B.du, src);
    g = -r;
    sb_bump<NB>(dw, dst, -dr_dst);
    sb_bump<NB>(dw, src, -dr_src);
    dphi = -(dr_dst + dr_src);
  } else if (kind == SB_PROC_SR_TRAP) {
    sb_sr_trap<NB>(M, tp, p, pp[SB_TPP_P0], beta, B, g, dphi, dw);
  } else if (kind == SB_PROC_RAD_BB) {
    double S = 0.0;
#pragma unroll
    for (int f = 0; f < SB_MAX_FIELDS; ++f)
      if (f < M.n_fields) S += phi_clip[f] * pp[SB_TPP_FIELD0 + f];
    g = S;
    if (M.proc_radiative[p]) {
      const double x = beta * (sb_pick<NB>(B.w, dst) - sb_pick<NB>(B.w, src));
      const double em = sb_expm1(x), ex = em + 1.0;
      g -= pp[SB_TPP_P0] * em;
      sb_bump<NB>(dw, dst, -pp[SB_TPP_P0] * beta * ex);
      sb_bump<NB>(dw, src, pp[SB_TPP_P0] * beta * ex);
    }
  } else if (kind == SB_PROC_RAD_IB) {
    const int IB = M.proc_trap[p];
    const int RB = (dst == IB) ? src : dst;
    const double N_I = tp[SB_TP_BAND0 + 5 * IB + SB_TPB_N];
    const bool same = (M.band_sign[IB] == M.band_sign[RB]);
    const double uI = sb_pick<NB>(B.u, IB), duI = sb_pick<NB>(B.du, IB);
    const double Fa = same ? uI : (N_I - uI);
    const double dFa = same ? duI : -duI;
    double S = 0.0;
#pragma unroll
    for (int f = 0; f < SB_MAX_FIELDS; ++f)
      if (f < M.n_fields) S += phi_clip[f] * pp[SB_TPP_FIELD0 + f];
    g = S * Fa;
    sb_bump<NB>(dw, IB, S * dFa);
    dphi = S * dFa;
    if (M.proc_radiative[p])
      sb_sr_trap<NB>(M, tp, p, pp[SB_TPP_P0], beta, B, g, dphi, dw);
  }
}

/* which processes need what, hoisted out of the point loop */
template <class MV>
SB_HD void sb_phaseB_scan(const MV& M, bool& need_u0, int& trap) {
  need_u0 = false;
  trap = -1;
  constexpr int unroll_procs = SbModelTraits<MV>::unroll_procs;
#pragma unroll unroll_procs
  for (int p = 0; p < M.n_procs; ++p) {
    need_u0 = need_u0 || (M.proc_kind[p] == SB_PROC_SRH);
    if (trap < 0 && (M.proc_kind[p] == SB_PROC_SR_TRAP ||
                     (M.proc_kind[p] == SB_PROC_RAD_IB && M.proc_radiative[p])))
      trap = M.proc_trap[p];
  }
}

/* Point n of the g-rule: totals per band gk[k] = g_k, dphik[k] = dg_k/dphi,
 * dwk[k * NB + l] = dg_k/dw_l.  dgv[3 p + i]: P1 nodal values of (phi | delta_w_l);
 * w0[l]: base_w of the cell; K[l]: band constants.                                */
template <int NB, class MV>
SB_HD void sb_phaseB_point(const SbTables& T, const MV& M, const double* tp, int n,
                           const double* dgv, const double* w0, const SbBandK* K,
                           bool need_u0, int trap, double beta_pt,
                           const double* thmeq6, const double* phi_opt, int64_t field_stride,
                           double* gk, double* dphik, double* dwk, uint32_t pmask = 0xFFFFFFFFu) {
  const double X = T.gX[n], Y = T.gY[n];
  const double l0 = 1.0 - X - Y;
  const double phi = dgv[0] * l0 + dgv[1] * X + dgv[2] * Y;
  const double* n2 = T.gp2[n];
  SbBandsAtPoint<NB> B;
#pragma unroll
  for (int l = 0; l < NB; ++l) {
    B.w[l] = w0[l] + dgv[3 + 3 * l] * l0 + dgv[4 + 3 * l] * X + dgv[5 + 3 * l] * Y;
    /* (skipping the bands without states in this region, N = 0, was measured slower: the
       branch splits the point's basic block and its exp chains no longer interleave) */
    sb_band_u_kind(K[l], M.band_kind[l], phi + B.w[l], B.u[l], B.du[l]);
    B.u0[l] = 0.0;
  }
  if (need_u0) {
    double phi_te = 0.0;
#pragma unroll
    for (int a = 0; a < 6; ++a) phi_te += (thmeq6 ? thmeq6[a] : 0.0) * n2[a];
#pragma unroll
    for (int l = 0; l < NB; ++l) { double d_; sb_band_u_kind(K[l], M.band_kind[l], phi_te, B.u0[l], d_); }
  }
  B.trap = trap;
  if (trap >= 0) {
    const double wt = sb_pick<NB>(B.w, trap);
#pragma unroll
    for (int l = 0; l < NB; ++l)
      B.em[l] = sb_expm1((double)M.band_sign[l] * (B.w[l] - wt) * beta_pt);   /* l == trap: expm1(0) = 0 */
  } else {
#pragma unroll
    for (int l = 0; l < NB; ++l) B.em[l] = 0.0;
  }
  double phi_clip[SB_MAX_FIELDS];
#pragma unroll
  for (int f = 0; f < SB_MAX_FIELDS; ++f) {
    double v = 0.0;
    if (f < M.n_fields) {
      const double* pf = phi_opt + (int64_t)f * field_stride;
#pragma unroll
      for (int a = 0; a < 6; ++a) v += pf[a] * n2[a];
    }
    phi_clip[f] = v < 0.0 ? 0.0 : v;      /* optical.py:201-203 */
  }
#pragma unroll
  for (int k = 0; k < NB; ++k) {
    gk[k] = 0.0; dphik[k] = 0.0;
#pragma unroll
    for (int l = 0; l < NB; ++l) dwk[k * NB + l] = 0.0;
  }
  constexpr int unroll_procs = SbModelTraits<MV>::unroll_procs;
#pragma unroll unroll_procs
  for (int p = 0; p < M.n_procs; ++p) {
    if (!((pmask >> p) & 1u)) continue;         /* process switched off in this cell's region */
    double g, dphi, dw[NB > 0 ? NB : 1];
    sb_process_point<NB>(M, tp, p, beta_pt, B, phi_clip, g, dphi, dw);
#pragma unroll
    for (int k = 0; k < NB; ++k) {
      const int sgi = sb_gen_sign(M, p, k);
      if (!sgi) continue;                       /* warp-uniform */
      const double sg = (double)sgi;
      gk[k] += sg * g; dphik[k] += sg * dphi;
#pragma unroll
      for (int l = 0; l < NB; ++l) dwk[k * NB + l] += sg * dw[l];
    }
  }
}

/* Moments of g_k, dg_k/dphi, dg_k/dw_l for every band k at once: per point each
 * band's density and each process are evaluated once.  dgv[p][3]: P1 nodal
 * values of (phi | delta_w_l); w0[l]: base_w of the cell.
 * out mG[k][.]: [0..2] g (P1), [3..8] d/dphi (P2), [9+6l ..] d/dw_l (P2).       */
template <int NB, class MV>
SB_HD void sb_phaseB_all(const SbTables& T, const MV& M, const double* tp,
                         const double (*dgv)[3], const double* w0, const bool* inside,
                         const double* thmeq6, const double* phi_opt, int64_t field_stride,
                         double (*mG)[SB_NG], bool region_mask = false) {
#pragma unroll
  for (int k = 0; k < NB; ++k)
#pragma unroll
    for (int t = 0; t < 9 + 6 * NB; ++t) mG[k][t] = 0.0;

  SbBandK K[NB > 0 ? NB : 1];
  bool need_u0;
  int trap;
  sb_phaseB_scan(M, need_u0, trap);
  /* Processes whose parameters are all zero in this cell's region (missing Spatial rule =
     0.0, fem/expr.py:148-153: e.g. the IB processes outside the IB layer) contribute exact
     zeros: skip them.  SRH divides by its lifetimes and is never skipped.              */
  uint32_t pmask = 0xFFFFFFFFu;
  if (region_mask) {
    pmask = 0u;
    bool any_trap = false;
    constexpr int unroll_procs = SbModelTraits<MV>::unroll_procs;
#pragma unroll unroll_procs
    for (int p = 0; p < M.n_procs; ++p) {
      const double* pp = tp + SB_TP_PROC0 + 8 * p;
      bool on = (M.proc_kind[p] == SB_PROC_SRH) || pp[SB_TPP_P0] != 0.0;
#pragma unroll
      for (int f = 0; f < SB_MAX_FIELDS; ++f) on = on || pp[SB_TPP_FIELD0 + f] != 0.0;
      if (on) {
        pmask |= 1u << p;
        any_trap = any_trap || M.proc_kind[p] == SB_PROC_SR_TRAP ||
                   (M.proc_kind[p] == SB_PROC_RAD_IB && M.proc_radiative[p]);
      }
    }
    if (!any_trap) trap = -1;                   /* no expm1 against the trap band needed */
  }
  const double beta_pt = 1.0 / tp[SB_TP_KT];
#pragma unroll
  for (int l = 0; l < NB; ++l) K[l] = sb_band_consts(sb_dyn(M), tp, l);
#ifndef SB_GEN_UNROLL
#define SB_GEN_UNROLL 1
#endif
  constexpr int gen_unroll = SB_GEN_UNROLL;
#pragma unroll gen_unroll
  for (int n = 0; n < T.ng; ++n) {
    double gk[NB > 0 ? NB : 1], dphik[NB > 0 ? NB : 1], dwk[NB > 0 ? NB * NB : 1];
    sb_phaseB_point<NB>(T, M, tp, n, &dgv[0][0], w0, K, need_u0, trap, beta_pt, thmeq6, phi_opt,
                        field_stride, gk, dphik, dwk, pmask);
    const double* wp = T.gwpsi[n];
#pragma unroll
    for (int k = 0; k < NB; ++k) {
      if (!inside[k]) continue;
#pragma unroll
      for (int t = 0; t < 3; ++t) mG[k][t] += wp[t] * gk[k];
#pragma unroll
      for (int t = 0; t < 6; ++t) mG[k][3 + t] += wp[t] * dphik[k];
#pragma unroll
      for (int l = 0; l < NB; ++l)
#pragma unroll
        for (int t = 0; t < 6; ++t) mG[k][9 + 6 * l + t] += wp[t] * dwk[k * NB + l];
    }
  }
}

/* ------------------------------------------------------------------------- *
 * phase C: scatter context of one cell (one thread)
 * ------------------------------------------------------------------------- */
struct SbIO {
  int L;
  const int32_t* dofs;       /* [L] global dofs of this cell                */
  const uint8_t* pat;        /* [L][L] rank pattern of this cell            */
  const int64_t* rowptr;
  const double* u;
  const double* bc_values;
  const int32_t* bcidx;      /* [L] index into bc_values | -1 (special cells) */
  uint64_t mlo;              /* constrained local dofs, bits 0..63          */
  uint32_t mhi;              /* bits 64..74                                 */
  double* vals;              /* may be NULL                                 */
  double* b;                 /* may be NULL                                 */
  /* native layout (csrc/native.inl): the (facet dof, facet dof) block of an interior facet
     is written once, by the facet's first cell, as own + neighbour contribution         */
  int nat;                   /* 0: two cells add atomically onto zeroed slots          */
  int32_t finfo;             /* per local facet f, bits 4f..: 1 boundary, 2 second cell */
  const double* snb;         /* [3][6] neighbour's block of this sub-problem per facet */
};

SB_HD constexpr int sb_sym3(int k, int k2) {      /* packed index of a symmetric 3 x 3 */
  return (k <= k2) ? (k * 3 - (k * (k - 1)) / 2 + (k2 - k)) : (k2 * 3 - (k2 * (k2 - 1)) / 2 + (k - k2));
}

SB_HD bool sb_is_bc(const SbIO& io, int l) {
  return l < 64 ? ((io.mlo >> l) & 1ull) : ((io.mhi >> (l - 64)) & 1u);
}
/* Dirichlet lift value u_j - bc_j of a constrained local dof */
SB_HD double sb_lift_value(const SbIO& io, int j) {
  return io.u[io.dofs[j]] - io.bc_values[io.bcidx[j]];
}

#if defined(__CUDA_ARCH__) || defined(SB_EMUL_BUILD)
#define SB_ATOMIC_ADD(ptr, v) atomicAdd((ptr), (v))
#else
#define SB_ATOMIC_ADD(ptr, v) (*(ptr) += (v))
#endif

/* Write one structurally non-zero Jacobian entry (local i, j) = val.
 * rowbase = rowptr[dofs[i]].  SP: the cell has Dirichlet dofs -> element-wise
 * symmetric rule; Fi is the residual accumulator of row i
 * and receives the x0 lifting in program order (bit-reproducible).  With
 * compile-time (i, j) every facet test folds away.                            */
template <bool SP>
SB_HD void sb_emit(const SbIO& io, int64_t rowbase, int i, int j, double val, double& Fi) {
  if (SP) {
    if (sb_is_bc(io, i)) {
      val = (i == j) ? 1.0 : 0.0;
    } else if (sb_is_bc(io, j)) {
      Fi -= val * sb_lift_value(io, j);
      val = 0.0;
    }
  }
  if (!io.vals) return;
  const int fi = sb_local_facet(i);
  const int rk = io.pat[i * io.L + j];
  if (rk == 0xFF) return;                 /* not in a compacted pattern */
  double* dst = io.vals + rowbase + rk;
  /* two cells add into the same slot only for (facet dof, facet dof) pairs of
     one facet; two addends onto a zeroed slot commute -> deterministic       */
  if (fi >= 0 && sb_local_facet(j) == fi) {
    if (io.nat) {
      const int bits = (io.finfo >> (4 * fi)) & 3;
      if (bits & 2) return;                                   /* the first cell writes it */
      if (!(bits & 1)) {
        /* the neighbour's share of the block: its cell block from the moment scratch -- unless
           the pair involves a Dirichlet dof of this (interior) facet: the neighbour applies the
           same element-wise rule, so its share is the identity entry */
        if (SP && (sb_is_bc(io, i) || sb_is_bc(io, j))) val += (i == j) ? 1.0 : 0.0;
        else val += io.snb[6 * fi + sb_sym3((i % 15 - 3) % 3, (j % 15 - 3) % 3)];
      }
      *dst = val;
    } else {
      SB_ATOMIC_ADD(dst, val);
    }
  } else {
    *dst = val;
  }
}

/* final residual value of local row i */
template <bool SP>
SB_HD void sb_put_row(const SbIO& io, int i, double F) {
  if (!io.b) return;
  if (SP && sb_is_bc(io, i)) F = sb_lift_value(io, i);
  double* dst = io.b + io.dofs[i];
  if (sb_local_facet(i) >= 0) SB_ATOMIC_ADD(dst, F); else *dst = F;
}

/* natural-BC values of this cell for rows [lo, lo+n): F[row-lo] += value */
SB_HD void sb_add_natbc(const int32_t* nat_row, const double* nat_val, int b0, int b1,
                        int lo, int n, double* F) {
  for (int k = b0; k < b1; ++k) {
    const int r = nat_row[k] - lo;
    if (r >= 0 && r < n) F[r] += nat_val[k];
  }
}

/* Loops over local rows/columns are fully unrolled on the fast path (all
 * indices compile-time: table operands become constant-bank immediates, rank
 * loads get immediate offsets) and kept rolled on the rare Dirichlet path.    */

/* ---- rows psi (Poisson) / xi^k (band k): BDM test functions ------------------ *
 * One row at a time, nothing indexed dynamically except the row itself:
 *   <xi_i . xi_m c> = iad sum_{b,s} GY_i[b][s] C[m][b][s],
 *   GY_i = G (C_i W),  W_rs = int c pi_r pi_s = sum_t G22u[rs][t] mC[t]
 * (orthonormal Dubiner basis: W = identity reproduces the plain BDM mass
 * matrix, used for the Poisson <psi.E> block and the out-of-subdomain penalty
 * <j.xi> c2, poisson_drift_diffusion1.py1:288-289).
 *   <xi_i . j dc/dchi lam_l> = iad sum_{a,r} C_i[a][r] V_a[r][l],
 *   V_a[r][l] = sum_t G21[r][l][t] Q_a[t].
 * mode 0: Poisson psi rows; 1: band inside its subdomain; 2: band outside.    */
template <bool SP>
SB_HD void sb_rows_flux(const SbTables& T, const sb_model& M, const SbIO& io,
                        const SbCellGeom& g, int p, int mode,
                        const double* mC, const double* Q,
                        const double* fl /*[12] flux dofs*/, const double* sl /*[3] scalar dofs*/,
                        double jump0, double jump1, double jump2,
                        const int32_t* nat_row, const double* nat_val, int nb0, int nb1,
                        bool mC_is_W = false /* native scratch: mC holds the 21 values of W and
                                                Q the 36 finished couplings d[3 i + l]        */) {
  const int r0 = sb_lbdm(p, 0);
  double W[21];
  double V[2][6][3];
  if (mode == 1) {
#pragma unroll
    for (int e = 0; e < 21; ++e) {
      double v = 0.0;
      if (mC_is_W) {
        v = mC[e];
      } else {
#pragma unroll
        for (int t = 0; t < 15; ++t) v += T.G22u[e][t] * mC[t];
      }
      W[e] = v;
    }
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int r = 0; r < 6; ++r)
#pragma unroll
        for (int l = 0; l < 3; ++l) {
          double v = 0.0;
          if (!mC_is_W) {
#pragma unroll
            for (int t = 0; t < 10; ++t) v += T.G21[r][l][t] * Q[a * 10 + t];
          }
          V[a][r][l] = v;
        }
  } else {
#pragma unroll
    for (int r = 0; r < 6; ++r)
#pragma unroll
      for (int s2 = r; s2 < 6; ++s2) W[sb_sym21(r, s2)] = (r == s2) ? 1.0 : 0.0;
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int r = 0; r < 6; ++r)
#pragma unroll
        for (int l = 0; l < 3; ++l) V[a][r][l] = 0.0;
  }
  const double sc = (mode == 2 ? M.ext_j_coef : 1.0) * g.iad;
  const double divsign = (mode == 0) ? -g.sdet : g.sdet;   /* -<div psi phi> | +<div xi dw> */
#pragma unroll 1
  for (int i = 0; i < 12; ++i) {
    double Ci[2][6];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int r = 0; r < 6; ++r) Ci[a][r] = T.bdmC[i][a][r];
    /* GY[b][s] = sum_a G_ab sum_r C_i[a][r] W_rs */
    double GY[2][6];
#pragma unroll
    for (int s2 = 0; s2 < 6; ++s2) {
      double y0 = 0.0, y1 = 0.0;
#pragma unroll
      for (int r = 0; r < 6; ++r) {
        y0 += Ci[0][r] * W[sb_sym21(r, s2)];
        y1 += Ci[1][r] * W[sb_sym21(r, s2)];
      }
      GY[0][s2] = (g.G11 * y0 + g.G12 * y1) * sc;
      GY[1][s2] = (g.G12 * y0 + g.G22 * y1) * sc;
    }
    const int li = r0 + i;
    const int64_t rowbase = io.rowptr[io.dofs[li]];
    double Fi = 0.0;
#pragma unroll
    for (int m = 0; m < 12; ++m) {
      double v = 0.0;
#pragma unroll
      for (int s2 = 0; s2 < 6; ++s2)
        v += GY[0][s2] * T.bdmC[m][0][s2] + GY[1][s2] * T.bdmC[m][1][s2];
      Fi += v * fl[m];
      sb_emit<SP>(io, rowbase, li, r0 + m, v, Fi);
    }
    if (mode != 2) {
#pragma unroll
      for (int l = 0; l < 3; ++l) {
        const double dv = divsign * T.Dhat[i][l];
        Fi += dv * sl[l];
        if (mode == 1) {
          double d = 0.0;
#pragma unroll
          for (int r = 0; r < 6; ++r) d += Ci[0][r] * V[0][r][l] + Ci[1][r] * V[1][r][l];
          d *= g.iad;
          if (mC_is_W) d = Q[3 * i + l];
          sb_emit<SP>(io, rowbase, li, sb_ldg(p, l), dv + d, Fi);
          sb_emit<SP>(io, rowbase, li, sb_ldg(0, l), d, Fi);
        } else {
          sb_emit<SP>(io, rowbase, li, sb_ldg(0, l), dv, Fi);
        }
      }
    }
    if (mode == 1 && i < 9) {                      /* base_w jump: facet dofs only */
      const double jf = (i < 3) ? jump0 : (i < 6 ? jump1 : jump2);
      Fi -= jf * T.trI[i % 3];
    }
    if (nb1 > nb0) sb_add_natbc(nat_row, nat_val, nb0, nb1, li, 1, &Fi);
    sb_put_row<SP>(io, li, Fi);
  }
}

/* ---- rows of the DG1 test functions: w (Poisson, p = 0) or v^k (p = 1 + k) ---- *
 * coef: multiplies the moment terms (-adet for rho, -s e adet for g)
 * mom1[3]: P1 moments of the source (rho/eps | g_k)
 * momJ[q][6]: P2 moments of d source / d (phi | delta_w_q-1), q < NQ          */
template <bool SP, int NQ>
SB_HD void sb_rows_scalar(const SbTables& T, const sb_model& M, const SbIO& io,
                          const SbCellGeom& g, int p, bool inside,
                          const double* fl /*[12] flux dofs of p*/, const double* sl /*[3]*/,
                          double coef, const double* mom1, const double* momJ,
                          const int32_t* nat_row, const double* nat_val, int nb0, int nb1) {
  const int r0 = sb_ldg(p, 0);
  double F[3] = {0.0, 0.0, 0.0};
  int64_t rb[3];
#pragma unroll 1
  for (int i = 0; i < 3; ++i) rb[i] = io.rowptr[io.dofs[r0 + i]];
  if (inside) {
#pragma unroll 1
    for (int i = 0; i < 3; ++i) {
#pragma unroll 1
      for (int m = 0; m < 12; ++m) {            /* <v_i div j> */
        const double dv = g.sdet * T.Dhat[m][i];
        F[i] += dv * fl[m];
        sb_emit<SP>(io, rb[i], r0 + i, sb_lbdm(p, m), dv, F[i]);
      }
#pragma unroll
      for (int t = 0; t < 3; ++t) F[i] += coef * T.G1[i][t] * mom1[t];
#pragma unroll 1
      for (int q = 0; q < NQ; ++q)
#pragma unroll 1
        for (int l = 0; l < 3; ++l) {
          double v = 0.0;
#pragma unroll
          for (int t = 0; t < 6; ++t) v += T.G11[i][l][t] * momJ[6 * q + t];
          sb_emit<SP>(io, rb[i], r0 + i, sb_ldg(q, l), coef * v, F[i]);
        }
    }
  } else {
    /* out-of-subdomain penalty <delta_w v> c1 (:285-286) */
    const double c1 = M.ext_w_coef * g.adet;
#pragma unroll 1
    for (int i = 0; i < 3; ++i)
#pragma unroll 1
      for (int l = 0; l < 3; ++l) {
        const double v = c1 * T.M1[i][l];
        F[i] += v * sl[l];
        sb_emit<SP>(io, rb[i], r0 + i, r0 + l, v, F[i]);
      }
  }
  if (nb1 > nb0) sb_add_natbc(nat_row, nat_val, nb0, nb1, r0, 3, F);
#pragma unroll 1
  for (int i = 0; i < 3; ++i) sb_put_row<SP>(io, r0 + i, F[i]);
}

/* ========================================================================= *
 * Fast row path: cells without Dirichlet dofs / natural-BC rows (all but the
 * boundary layer), compile-time sub-problem PS.  Ranks come from the packed
 * emission-order table patk[pattern][row][32] (two 16-byte loads per row):
 *   BDM row: [0..11] BDM cols of PS, [12..14] DG cols of PS, [15..17] DG cols of 0
 *   DG  row: [0..11] BDM cols of PS, [12 + 3 q + l] DG col l of sub-problem q
 * ========================================================================= */
struct SbFastIO {
  const int32_t* dofs;
  const uint8_t* patk;       /* [L][32] packed ranks of this cell's pattern */
  double* vals;
  double* b;
  /* per (entity) group of this sub-problem: 0 DG rows, 1..3 facets, 4 interior */
  int64_t rb[5];             /* vals offset of the group's first row            */
  int32_t rl[5];             /* row length (the 3 rows of a group are adjacent) */
  int32_t d0[5];             /* global dof of the group's first row             */
};

struct SbRanks { uint32_t w[8]; };

SB_HD SbRanks sb_load_ranks(const uint8_t* patk, int li) {
  SbRanks r;
  const uint4* q = reinterpret_cast<const uint4*>(patk + li * 32);
  const uint4 a = q[0], b = q[1];
  r.w[0] = a.x; r.w[1] = a.y; r.w[2] = a.z; r.w[3] = a.w;
  r.w[4] = b.x; r.w[5] = b.y; r.w[6] = b.z; r.w[7] = b.w;
  return r;
}
SB_HD int sb_rank(const SbRanks& r, int e) { return (int)((r.w[e >> 2] >> (8 * (e & 3))) & 0xFFu); }

SB_HD void sb_store(double* vals, int64_t rowbase, int rk, double v, bool shared_slot) {
  if (!vals || rk == 0xFF) return;
#if defined(SB_EXP_NOSTORE)          /* experiment: keep the math, drop the scatter */
  if (v == 1.2345e-300) vals[0] = v;
  return;
#elif defined(SB_EXP_STOREONLY)      /* experiment: keep the scatter, drop the math */
  v = 1.0;
#endif
  double* dst = vals + rowbase + rk;
  if (shared_slot) SB_ATOMIC_ADD(dst, v); else *dst = v;
}

/* BDM rows of sub-problem PS; MODE 0: Poisson psi rows, 1: band inside its
 * subdomain, 2: band outside (see sb_rows_flux for the algebra)               */
template <int PS, int MODE>
SB_HD void sb_rows_flux_fast(const SbTables& T, const sb_model& M, const SbFastIO& io,
                             const SbCellGeom& g, const double* mC, const double* Q,
                             const double* fl, const double* sl,
                             double jump0, double jump1, double jump2) {
  constexpr int r0 = 15 * PS + 3;
  double W[21];
  double V[2][6][3];
  if (MODE == 1) {
#pragma unroll
    for (int e = 0; e < 21; ++e) {
      double v = 0.0;
#pragma unroll
      for (int t = 0; t < 15; ++t) v += T.G22u[e][t] * mC[t];
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
          for (int t = 0; t < 10; ++t) v += T.G21[r][l][t] * Q[a * 10 + t];
          V[a][r][l] = v;
        }
  }
  const double sc = (MODE == 2 ? M.ext_j_coef : 1.0) * g.iad;
  const double divsign = (MODE == 0) ? -g.sdet : g.sdet;
#pragma unroll 1
  for (int i = 0; i < 12; ++i) {
    const int grp = i / 3, r3 = i - 3 * grp;   /* facet 0, 1, 2, interior */
    const int64_t gbase = grp == 0 ? io.rb[1] : (grp == 1 ? io.rb[2] : (grp == 2 ? io.rb[3] : io.rb[4]));
    const int32_t glen = grp == 0 ? io.rl[1] : (grp == 1 ? io.rl[2] : (grp == 2 ? io.rl[3] : io.rl[4]));
    const int32_t gdof = grp == 0 ? io.d0[1] : (grp == 1 ? io.d0[2] : (grp == 2 ? io.d0[3] : io.d0[4]));
    const int fi = (grp < 3) ? grp : -1;
    const double jf = (grp == 0) ? jump0 : (grp == 1 ? jump1 : jump2);
    {
      double Ci[2][6];
#pragma unroll
      for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int r = 0; r < 6; ++r) Ci[a][r] = T.bdmC[i][a][r];
      double GY[2][6];
#pragma unroll
      for (int s2 = 0; s2 < 6; ++s2) {
        double y0, y1;
        if (MODE == 1) {
          y0 = 0.0; y1 = 0.0;
#pragma unroll
          for (int r = 0; r < 6; ++r) {
            y0 += Ci[0][r] * W[sb_sym21(r, s2)];
            y1 += Ci[1][r] * W[sb_sym21(r, s2)];
          }
        } else {                       /* W = identity: plain BDM mass matrix */
          y0 = Ci[0][s2]; y1 = Ci[1][s2];
        }
        GY[0][s2] = (g.G11 * y0 + g.G12 * y1) * sc;
        GY[1][s2] = (g.G12 * y0 + g.G22 * y1) * sc;
      }
      const int li = r0 + i;
      const int64_t rowbase = gbase + (int64_t)r3 * glen;
      const SbRanks rk = sb_load_ranks(io.patk, li);
      double Fi = 0.0;
#pragma unroll
      for (int m = 0; m < 12; ++m) {
        double v = 0.0;
#pragma unroll
        for (int s2 = 0; s2 < 6; ++s2)
          v = fma(GY[1][s2], T.bdmC[m][1][s2], fma(GY[0][s2], T.bdmC[m][0][s2], v));
        Fi += v * fl[m];
        sb_store(io.vals, rowbase, sb_rank(rk, m), v, (m < 9) && (m / 3 == fi));
      }
      if (MODE != 2) {
#pragma unroll
        for (int l = 0; l < 3; ++l) {
          const double dv = divsign * T.Dhat[i][l];
          Fi += dv * sl[l];
          if (MODE == 1) {
            double d = 0.0;
#pragma unroll
            for (int r = 0; r < 6; ++r) d = fma(Ci[1][r], V[1][r][l], fma(Ci[0][r], V[0][r][l], d));
            d *= g.iad;
            sb_store(io.vals, rowbase, sb_rank(rk, 12 + l), dv + d, false);
            sb_store(io.vals, rowbase, sb_rank(rk, 15 + l), d, false);
          } else {
            sb_store(io.vals, rowbase, sb_rank(rk, 12 + l), dv, false);
          }
        }
      }
      if (MODE == 1 && grp < 3) Fi -= jf * T.trI[r3];     /* base_w jump: facet dofs */
      if (io.b) {
        double* dst = io.b + (gdof + r3);
        if (fi >= 0) SB_ATOMIC_ADD(dst, Fi); else *dst = Fi;
      }
    }
  }
}

/* DG1 rows of sub-problem PS (w for PS = 0, v^k otherwise) */
template <int PS, int NQ>
SB_HD void sb_rows_scalar_fast(const SbTables& T, const sb_model& M, const SbFastIO& io,
                               const SbCellGeom& g, bool inside, const double* fl,
                               const double* sl, double coef, const double* mom1,
                               const double* momJ) {
  constexpr int r0 = 15 * PS;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const int li = r0 + i;
    const int64_t rowbase = io.rb[0] + (int64_t)i * io.rl[0];
    const SbRanks rk = sb_load_ranks(io.patk, li);
    double Fi = 0.0;
    if (inside) {
#pragma unroll
      for (int m = 0; m < 12; ++m) {
        const double dv = g.sdet * T.Dhat[m][i];
        Fi += dv * fl[m];
        sb_store(io.vals, rowbase, sb_rank(rk, m), dv, false);
      }
#pragma unroll
      for (int t = 0; t < 3; ++t) Fi += coef * T.G1[i][t] * mom1[t];
#pragma unroll
      for (int q = 0; q < NQ; ++q)
#pragma unroll
        for (int l = 0; l < 3; ++l) {
          double v = 0.0;
#pragma unroll
          for (int t = 0; t < 6; ++t) v += T.G11[i][l][t] * momJ[6 * q + t];
          sb_store(io.vals, rowbase, sb_rank(rk, 12 + 3 * q + l), coef * v, false);
        }
    } else {
      const double c1 = M.ext_w_coef * g.adet;
#pragma unroll
      for (int l = 0; l < 3; ++l) {
        const double v = c1 * T.M1[i][l];
        Fi += v * sl[l];
        sb_store(io.vals, rowbase, sb_rank(rk, 12 + 3 * PS + l), v, false);
      }
    }
    if (io.b) io.b[io.d0[0] + i] = Fi;
  }
}


/* =========================================================================
 * Staged fast path.  Measured on B200 (profiles/README.md, "store-bound"): with
 * one thread per cell every structural entry is its own 8-byte store -- a 32-byte
 * sector transaction each -- and the row kernels spend 10.2 of 10.5 ms in that
 * scatter; the arithmetic alone takes 2.9 ms.  Here a thread only *stages* the
 * entries of one entity group (3 adjacent rows) in shared memory, in emission
 * order; the warp then flushes cell after cell with lane = entry, so the stores of
 * one instruction fall into the same one or two CSR rows and coalesce into full
 * sectors (DG and interior rows are written as whole contiguous runs).
 *
 * Staging slot of lane:  [r3 * NK + k] value of emission slot k of row r3
 *                        [3 * NK + r3] residual of row r3
 * NK = 24 for the DG group (k: 0..11 BDM cols, 12 + 3 q + l DG col l of q),
 * NK = 18 for the BDM groups (k: 0..11 BDM, 12..14 DG(PS), 15..17 DG(0)) --
 * the same order as the packed rank table patk.
 * ========================================================================= */
#define SB_STAGE_STRIDE 78          /* doubles per lane: 1 + 3 * 24 + 3 used; 78 * 8 B keeps every
                                       lane's base 16-byte aligned for the bulk copies */

/* per-warp flush metadata, one entry per lane (cell) and group */
struct SbWarpMeta {
  int64_t rb[5][32];                /* vals offset of the group's first row   */
  int32_t rl[5][32];                /* row length                             */
  int32_t d0[5][32];                /* global dof of the group's first row    */
  int32_t info[32];                 /* bit 0 active, bit 1 inside, bits 8.. pattern */
};

/* Rows owned by one cell (its DG rows, its interior BDM rows) are a contiguous run
 * of 3 * rl values that this thread alone writes.  They are staged in *rank* order
 * (final memory order) at stg[phase + ...], phase = parity of the run's first index,
 * so that shared and global addresses have the same 16-byte phase, and leave the SM
 * as one asynchronous bulk copy (cp.async.bulk shared -> global, the 1-D TMA path);
 * an odd start or an odd length contributes one scalar store at that end.                          */
SB_HD void sb_bulk_store_owned(double* gdst, const double* stg, int phase, int n) {
#if defined(__CUDA_ARCH__)
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  const double* s0 = stg + phase;
  const int head = phase;                         /* one scalar up to the 16-byte boundary */
  const int nb = (n - head) & ~1;                 /* bulk part: whole 16-byte units        */
  if (head) gdst[0] = s0[0];
  if (head + nb < n) gdst[n - 1] = s0[n - 1];
  const uint32_t saddr = (uint32_t)__cvta_generic_to_shared(s0 + head);
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
               :: "l"(gdst + head), "r"(saddr), "r"(nb * 8) : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
#else
  for (int e = 0; e < n; ++e) gdst[e] = stg[phase + e];
#endif
}
/* the staging area may be rewritten once the bulk engine has read it */
SB_HD void sb_bulk_wait_read() {
#if defined(__CUDA_ARCH__)
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
#endif
}

/* ---- asynchronous input staging of the persistent row kernels (native.inl) ---------------- *
 * cp.async (LDGSTS): global -> shared copies that occupy no register while in flight, one
 * group per cell tile; L2 prefetches for what the next tile will load with plain loads.    */
SB_HD void sb_cp_async4(void* sdst, const void* gsrc) {
#if defined(__CUDA_ARCH__)
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;"
               :: "r"((uint32_t)__cvta_generic_to_shared(sdst)), "l"(gsrc) : "memory");
#else
  memcpy(sdst, gsrc, 4);
#endif
}
SB_HD void sb_cp_async8(void* sdst, const void* gsrc) {
#if defined(__CUDA_ARCH__)
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;"
               :: "r"((uint32_t)__cvta_generic_to_shared(sdst)), "l"(gsrc) : "memory");
#else
  memcpy(sdst, gsrc, 8);
#endif
}
SB_HD void sb_cp_async16(void* sdst, const void* gsrc) {
#if defined(__CUDA_ARCH__)
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;"
               :: "r"((uint32_t)__cvta_generic_to_shared(sdst)), "l"(gsrc) : "memory");
#else
  memcpy(sdst, gsrc, 16);
#endif
}
SB_HD void sb_cp_async_commit() {
#if defined(__CUDA_ARCH__)
  asm volatile("cp.async.commit_group;" ::: "memory");
#endif
}
SB_HD void sb_cp_async_wait_all() {
#if defined(__CUDA_ARCH__)
  asm volatile("cp.async.wait_group 0;" ::: "memory");
#endif
}
SB_HD void sb_prefetch_l2(const void* g) {
#if defined(__CUDA_ARCH__)
  asm volatile("prefetch.global.L2 [%0];" :: "l"(g));
#else
  (void)g;
#endif
}
/* one instruction for a contiguous block (1-D TMA prefetch): g 16-byte aligned, bytes % 16 == 0 */
SB_HD void sb_prefetch_l2_bulk(const void* g, uint32_t bytes) {
#if defined(__CUDA_ARCH__)
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" :: "l"(g), "r"(bytes) : "memory");
#else
  (void)g; (void)bytes;
#endif
}

/* DG1 rows of sub-problem PS into the staging area (see sb_rows_scalar) */
/* pk != nullptr: rank-ordered staging for the bulk copy (pk = the cell's rank rows of
 * this group, rl = row length, st already offset by the phase); the residual then goes
 * straight to bdst (global, may be null) */
template <int PS, int NQ>
SB_HD void sb_stage_scalar_rows(const SbTables& T, const sb_model& M, const SbCellGeom& g,
                                bool inside, const double* fl, const double* sl, double coef,
                                const double* mom1, const double* momJ, double* st,
                                const uint8_t* pk = nullptr, int rl = 24, double* bdst = nullptr) {
#define SB_DG_SLOT(i, k) (pk ? (i) * rl + (int)pk[(i) * 32 + (k)] : (i) * 24 + (k))
  if (pk && !inside) {               /* not emitted entries of owned rows stay zero */
    for (int e = 0; e < 3 * rl; ++e) st[e] = 0.0;
  }
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    double Fi = 0.0;
    if (inside) {
#pragma unroll
      for (int m = 0; m < 12; ++m) {
        const double dv = g.sdet * T.Dhat[m][i];
        Fi += dv * fl[m];
        st[SB_DG_SLOT(i, m)] = dv;
      }
#pragma unroll
      for (int t = 0; t < 3; ++t) Fi += coef * T.G1[i][t] * mom1[t];
#pragma unroll
      for (int q = 0; q < NQ; ++q)
#pragma unroll
        for (int l = 0; l < 3; ++l) {
          double v = 0.0;
#pragma unroll
          for (int t = 0; t < 6; ++t) v += T.G11[i][l][t] * momJ[6 * q + t];
          st[SB_DG_SLOT(i, 12 + 3 * q + l)] = coef * v;
        }
    } else {
      const double c1 = M.ext_w_coef * g.adet;
#pragma unroll
      for (int l = 0; l < 3; ++l) {
        const double v = c1 * T.M1[i][l];
        Fi += v * sl[l];
        st[SB_DG_SLOT(i, 12 + 3 * PS + l)] = v;
      }
    }
    if (pk) { if (bdst) bdst[i] = Fi; } else st[72 + i] = Fi;
  }
#undef SB_DG_SLOT
}

/* the three BDM rows of entity group grp (facet 0, 1, 2, interior) into the staging
 * area; MODE as in sb_rows_flux_fast                                             */
template <int PS, int MODE>
SB_HD void sb_stage_flux_group(const SbTables& T, const sb_model& M, const SbCellGeom& g,
                               const double* W, const double (*V)[6][3], const double* fl,
                               const double* sl, double jf, int grp, double* st,
                               const uint8_t* pk = nullptr, int rl = 18, double* bdst = nullptr) {
  const double sc = (MODE == 2 ? M.ext_j_coef : 1.0) * g.iad;
  const double divsign = (MODE == 0) ? -g.sdet : g.sdet;
  if (pk && MODE == 2) {             /* not emitted entries of owned rows stay zero */
    for (int e = 0; e < 3 * rl; ++e) st[e] = 0.0;
  }
#pragma unroll 1
  for (int r3 = 0; r3 < 3; ++r3) {
    const int i = 3 * grp + r3;
    double Ci[2][6];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int r = 0; r < 6; ++r) Ci[a][r] = T.bdmC[i][a][r];
    double GY[2][6];
#pragma unroll
    for (int s2 = 0; s2 < 6; ++s2) {
      double y0, y1;
      if (MODE == 1) {
        y0 = 0.0; y1 = 0.0;
#pragma unroll
        for (int r = 0; r < 6; ++r) {
          y0 += Ci[0][r] * W[sb_sym21(r, s2)];
          y1 += Ci[1][r] * W[sb_sym21(r, s2)];
        }
      } else {                       /* W = identity: plain BDM mass matrix */
        y0 = Ci[0][s2]; y1 = Ci[1][s2];
      }
      GY[0][s2] = (g.G11 * y0 + g.G12 * y1) * sc;
      GY[1][s2] = (g.G12 * y0 + g.G22 * y1) * sc;
    }
    double* row = st + (pk ? r3 * rl : r3 * 18);
    const uint8_t* prow = pk ? pk + r3 * 32 : nullptr;
#define SB_BDM_SLOT(k) (prow ? (int)prow[(k)] : (k))
    double Fi = 0.0;
#pragma unroll
    for (int m = 0; m < 12; ++m) {
      double v = 0.0;
#pragma unroll
      for (int s2 = 0; s2 < 6; ++s2)
        v = fma(GY[1][s2], T.bdmC[m][1][s2], fma(GY[0][s2], T.bdmC[m][0][s2], v));
      Fi += v * fl[m];
      row[SB_BDM_SLOT(m)] = v;
    }
    if (MODE != 2) {
#pragma unroll
      for (int l = 0; l < 3; ++l) {
        const double dv = divsign * T.Dhat[i][l];
        Fi += dv * sl[l];
        if (MODE == 1) {
          double d = 0.0;
#pragma unroll
          for (int r = 0; r < 6; ++r) d = fma(Ci[1][r], V[1][r][l], fma(Ci[0][r], V[0][r][l], d));
          d *= g.iad;
          row[SB_BDM_SLOT(12 + l)] = dv + d;
          row[SB_BDM_SLOT(15 + l)] = d;
        } else {
          row[SB_BDM_SLOT(12 + l)] = dv;
        }
      }
    }
    if (MODE == 1 && grp < 3) Fi -= jf * T.trI[r3];     /* base_w jump: facet dofs */
    if (pk) { if (bdst) bdst[r3] = Fi; } else st[54 + r3] = Fi;
#undef SB_BDM_SLOT
  }
}

/* Warp-cooperative flush of one staged group: cell after cell, lane = entry.
 * NK: slots per row; klo/khi: emitted slot range of a cell given its inside flag;
 * facet >= 0: group of facet `facet` (its (facet, facet) entries and residual rows
 * are shared with the neighbour -> atomic adds onto zeroed slots).              */
template <int NK>
SB_HD void sb_flush_group(const SbWarpMeta& meta, const double* __restrict__ stage0,
                          const uint8_t* __restrict__ patk0, int patk_pattern_stride,
                          int patk_row0, int grp5, int facet, int klo_in, int khi_in,
                          int klo_out, int khi_out, double* __restrict__ vals,
                          double* __restrict__ b, int lane) {
  constexpr int NE = 3 * NK + 3;
  constexpr int NIT = (NE + 31) / 32;
  /* Everything that depends only on the lane is hoisted out of the cell loop: slot
     e = lane + 32 it -> row r, emission slot k, rank-table offset, and (as bit `it` of
     a mask) whether inside / outside cells emit it, whether it is a shared (atomic)
     slot, whether it is a residual slot.                                             */
  int soff[NIT], rsel[NIT];
  uint32_t inbits = 0, outbits = 0, shbits = 0, resbits = 0;
#pragma unroll
  for (int it = 0; it < NIT; ++it) {
    const int e = lane + 32 * it;
    const bool ent = e < 3 * NK;
    const int r = ent ? e / NK : (e < NE ? e - 3 * NK : 0);
    const int k = ent ? e - r * NK : 0;
    rsel[it] = r;
    soff[it] = ent ? r * 32 + k : 0;
    if (ent && vals != nullptr && k >= klo_in && k < khi_in) inbits |= 1u << it;
    if (ent && vals != nullptr && k >= klo_out && k < khi_out) outbits |= 1u << it;
    if (facet >= 0 && k < 9 && k / 3 == facet) shbits |= 1u << it;
    if (!ent && e < NE && b != nullptr) resbits |= 1u << it;
  }
  const uint8_t* pk0 = patk0 + patk_row0 * 32;
  /* independent cells: unrolled so that the shared-memory reads of several cells are
     in flight together */
#pragma unroll 8
  for (int i = 0; i < 32; ++i) {
    const int32_t info = meta.info[i];
    const uint32_t onb = (info & 1) ? ((info & 2) ? inbits : outbits) : 0u;
    const uint32_t resb = (info & 1) ? resbits : 0u;
    const int rl = meta.rl[grp5][i];
    double* base = vals + meta.rb[grp5][i];
    double* bbase = b + meta.d0[grp5][i];
    const uint8_t* pk = pk0 + (info >> 8) * patk_pattern_stride;
    const double* src = stage0 + i * SB_STAGE_STRIDE + lane;
#pragma unroll
    for (int it = 0; it < NIT; ++it) {
      const double v = src[32 * it];
      const int off = rsel[it] * rl + (int)pk[soff[it]];      /* < 3 * 255 + 255 */
      double* dst = base + off;
      const bool on = (onb >> it) & 1u;
#ifdef SB_EXP_NOATOMIC
      if (on) *dst = v;
#else
      const bool sh = (shbits >> it) & 1u;
      if (on && sh) SB_ATOMIC_ADD(dst, v);
      if (on && !sh) *dst = v;
#endif
      if ((resb >> it) & 1u) {
        double* rdst = bbase + rsel[it];
        if (facet >= 0) SB_ATOMIC_ADD(rdst, v); else *rdst = v;
      }
    }
  }
}

#endif /* SB_PDD_DEVICE_CUH */
