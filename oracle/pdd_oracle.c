/*
 * ORACLE -- test infrastructure, NOT product code.  PARITY UNPINNED.
 *
 * CPU restatement of what FFC-generated tabulate_tensor + dolfin
 * SystemAssembler compute for Simudo's coupled Poisson / drift-diffusion form
 * (the reference's arithmetic lives in un-vendored legacy FEniCS, see
 * oracle/fiat_like.py).  Structure mirrors the generated code on purpose: one
 * dense L x L element tensor and one element vector per cell, filled by plain
 * loops over quadrature points x test x trial functions, then element-wise
 * Dirichlet application and insertion into the global CSR by column search
 * (MatSetValues-style).  Pointwise Gateaux derivatives are obtained by
 * forward-mode dual numbers (the analogue of UFL's symbolic derivative,
 * simudo/fem/newton_solver.py:275), NOT by the hand-derived closed forms the
 * CUDA kernels use.
 *
 * Forms restated (paths relative to /root/reference):
 *   mixed Poisson            simudo/physics/poisson_drift_diffusion1.py1:17-30,42-47
 *   band statistics          :170-197 (nondegenerate), :199-245 (intermediate)
 *   continuity + drift-diff. :259-303, subdomain measures :320-327
 *   generation processes     simudo/physics/electro_optical_process.py:141-157,
 *                            :294-312, :420-431, :448-492, :516-536, :553-562
 *   clipped photon flux      simudo/physics/optical.py:201-203
 *   Dirichlet with x0        simudo/fem/newton_solver.py:70-77
 *
 * Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline leg and
 * --impl reference) may load this library.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "../include/simudo_b200.h"   /* input data layout only (sb_model, SB_TP_*) */

#define ND (1 + SB_MAX_BANDS)          /* derivative slots: phi, w_0 .. w_3 */
#define MAXL (15 * (1 + SB_MAX_BANDS))

typedef struct { double v; double d[ND]; } dual;

static dual dc(double c) { dual r; r.v = c; memset(r.d, 0, sizeof r.d); return r; }
static dual dadd(dual a, dual b) { for (int i = 0; i < ND; ++i) a.d[i] += b.d[i]; a.v += b.v; return a; }
static dual dsub(dual a, dual b) { for (int i = 0; i < ND; ++i) a.d[i] -= b.d[i]; a.v -= b.v; return a; }
static dual dscale(dual a, double s) { for (int i = 0; i < ND; ++i) a.d[i] *= s; a.v *= s; return a; }
static dual dmul(dual a, dual b) {
  dual r; r.v = a.v * b.v;
  for (int i = 0; i < ND; ++i) r.d[i] = a.d[i] * b.v + a.v * b.d[i];
  return r;
}
static dual ddiv(dual a, dual b) {
  dual r; r.v = a.v / b.v;
  for (int i = 0; i < ND; ++i) r.d[i] = (a.d[i] - r.v * b.d[i]) / b.v;
  return r;
}
static dual dexp(dual a) { double e = exp(a.v); for (int i = 0; i < ND; ++i) a.d[i] *= e; a.v = e; return a; }
/* d/dx expm1 = exp  (simudo/fem/ufl_extra.py:32-61) */
static dual dexpm1(dual a) {
  const double e = exp(a.v), m = expm1(a.v);
  for (int i = 0; i < ND; ++i) a.d[i] *= e;
  a.v = m;
  return a;
}
static dual dpow(dual a, double p) {
  dual r; r.v = pow(a.v, p); double f = p * pow(a.v, p - 1.0);
  for (int i = 0; i < ND; ++i) r.d[i] = f * a.d[i];
  return r;
}

/* quadrature-point tabulation of the reference bases (from oracle/fiat_like.py) */
typedef struct {
  int32_t n;
  const double* w;       /* [n]          reference weights (sum 1/2)        */
  const double* p1;      /* [n][3]       P1 Lagrange                        */
  const double* p2;      /* [n][6]       P2 Lagrange                        */
  const double* bdm;     /* [n][12][2]   reference BDM2 values              */
  const double* bdmdiv;  /* [n][12]      reference BDM2 divergence          */
} oracle_rule;

typedef struct {
  sb_model model;
  int64_t n_cells, n_dofs;
  int32_t P;
  const double*  cell_vertices;   /* [nc][3][2] */
  const int32_t* cell_tag;
  const int32_t* cell_dofs;       /* [nc][L] */
  const int32_t* cell_neighbors;  /* [nc][3] */
  const uint8_t* cell_jump_mask;  /* [nc][3] */
  const double*  tag_params;      /* [n_tags][SB_TAG_STRIDE] */
  const int64_t* rowptr;          /* [N+1] */
  const int32_t* colidx;          /* [nnz] sorted per row */
  int64_t n_bc;
  const int32_t* bc_dofs;         /* sorted */
  const double*  bc_values;
  int64_t n_natbc;
  const int32_t* natbc_cell;      /* sorted by cell */
  const int32_t* natbc_row;
  const double*  natbc_values;
  oracle_rule rule_default, rule_rho, rule_super, rule_g;
  int32_t n_facet_q;
  const double* facet_w;          /* [nq] GL weights on [0,1]               */
  const double* facet_trace;      /* [3][nq][12] BDM scaled-normal traces   */
  /* state */
  const double* u;                /* [N] */
  const double* base_w;           /* [n_bands][nc] or NULL */
  const double* thmeq_phi;        /* [nc][6] or NULL */
  const double* phi_opt;          /* [n_fields][nc][6] or NULL */
  int32_t n_threads;
  int32_t cols_sorted;            /* 1: ascending columns per row (binary search);
                                     0: any order inside a row (linear search) */
} oracle_problem;

static double fd_f(double x) { return 1.0 / (1.0 + exp(x)); }

/* carrier density u(phiqfl) as a dual; statistics per band kind */
static dual band_u(const sb_model* M, const double* tp, int k, dual phiqfl) {
  const double* bp = tp + SB_TP_BAND0 + 5 * k;
  double kT = tp[SB_TP_KT];
  double s = (double)M->band_sign[k];
  if (M->band_kind[k] == SB_BAND_NONDEGENERATE) {
    double N0 = bp[SB_TPB_N] * exp(s * bp[SB_TPB_E] / kT);
    return dscale(dexp(dscale(phiqfl, -s / kT)), N0);
  } else {
    dual x = dscale(dsub(phiqfl, dc(bp[SB_TPB_E])), s / kT);
    dual f = ddiv(dc(1.0), dadd(dc(1.0), dexp(x)));
    return dscale(f, bp[SB_TPB_N]);
  }
}

static dual band_mobility(const sb_model* M, const double* tp, int k, dual phiqfl) {
  const double* bp = tp + SB_TP_BAND0 + 5 * k;
  if (M->band_kind[k] == SB_BAND_NONDEGENERATE || M->band_const_mobility[k])
    return dc(bp[SB_TPB_MOBILITY]);
  double kT = tp[SB_TP_KT];
  double s = (double)M->band_sign[k];
  dual x = dscale(dsub(phiqfl, dc(bp[SB_TPB_E])), s / kT);
  dual fill = ddiv(dc(1.0), dadd(dc(1.0), dexp(x)));
  dual omf = ddiv(dc(1.0), dadd(dc(1.0), dexp(dscale(x, -1.0))));
  dual t = dmul(dpow(fill, 0.33), dpow(omf, 1.33));
  return dscale(dsub(omf, dscale(t, 0.5)), bp[SB_TPB_MOBILITY]);
}

/* generation sign of a two-band process for band k (TwoBandEOPMixin) */
static int gen_sign(const sb_model* M, int p, int k) {
  if (k == M->proc_dst[p]) return +1;
  if (k == M->proc_src[p]) return -(M->band_sign[M->proc_dst[p]] * M->band_sign[M->proc_src[p]]);
  return 0;
}

/* Shockley-Read trapping kernel shared by NonRadiativeTrap and the radiative
 * IB process (electro_optical_process.py:294-312) */
static dual sr_trap_generation(const sb_model* M, const double* tp, int p, double coeff,
                               const dual* u, const dual* w) {
  int IB = M->proc_trap[p];
  int RB = (M->proc_dst[p] == IB) ? M->proc_src[p] : M->proc_dst[p];
  double kT = tp[SB_TP_KT];
  double N_I = tp[SB_TP_BAND0 + 5 * IB + SB_TPB_N];
  dual fill = (M->band_sign[IB] == M->band_sign[RB]) ? dsub(dc(N_I), u[IB]) : u[IB];
  dual x = dscale(dsub(w[RB], w[IB]), (double)M->band_sign[RB] / kT);
  dual r = dscale(dmul(dmul(dscale(dexpm1(x), -1.0), u[RB]), fill), coeff);
  return dscale(r, -1.0);
}

/* g for band k at a point: sum over processes */
static dual generation(const sb_model* M, const double* tp, int k,
                       const dual* u, const dual* w, const double* u0,
                       const double* phi_clip) {
  dual g = dc(0.0);
  for (int p = 0; p < M->n_procs; ++p) {
    int sg = gen_sign(M, p, k);
    if (!sg) continue;
    const double* pp = tp + SB_TP_PROC0 + 8 * p;
    int dst = M->proc_dst[p], src = M->proc_src[p];
    double kT = tp[SB_TP_KT];
    switch (M->proc_kind[p]) {
      case SB_PROC_SRH: {
        /* r = (p n - p0 n0) / ((n + n1) tau_p + (p + p1) tau_n); dst = CB, src = VB */
        dual num = dsub(dmul(u[src], u[dst]), dc(u0[src] * u0[dst]));
        dual den = dadd(dscale(dadd(u[dst], dc(pp[SB_TPP_P2])), pp[SB_TPP_P1]),
                        dscale(dadd(u[src], dc(pp[SB_TPP_P3])), pp[SB_TPP_P0]));
        g = dadd(g, dscale(ddiv(num, den), -(double)sg));
      } break;
      case SB_PROC_SR_TRAP: {
        g = dadd(g, dscale(sr_trap_generation(M, tp, p, pp[SB_TPP_P0], u, w), (double)sg));
      } break;
      case SB_PROC_RAD_BB: {
        dual gabs = dc(0.0);
        for (int f = 0; f < M->n_fields; ++f)
          gabs.v += phi_clip[f] * pp[SB_TPP_FIELD0 + f] * (double)sg;
        dual grad = dc(0.0);
        if (M->proc_radiative[p]) {
          dual x = dscale(dsub(w[dst], w[src]), 1.0 / kT);
          grad = dscale(dexpm1(x), pp[SB_TPP_P0] * (double)sg);
        }
        g = dadd(g, dsub(gabs, grad));
      } break;
      case SB_PROC_RAD_IB: {
        int IB = M->proc_trap[p];
        int RB = (dst == IB) ? src : dst;
        double N_I = tp[SB_TP_BAND0 + 5 * IB + SB_TPB_N];
        dual fill = (M->band_sign[IB] == M->band_sign[RB]) ? u[IB] : dsub(dc(N_I), u[IB]);
        dual gabs = dc(0.0);
        for (int f = 0; f < M->n_fields; ++f)
          gabs = dadd(gabs, dscale(fill, phi_clip[f] * pp[SB_TPP_FIELD0 + f] * (double)sg));
        dual grad = dc(0.0);
        if (M->proc_radiative[p])   /* g_rad = -(g_SR with pseudo capture coeff) * sign */
          grad = dscale(sr_trap_generation(M, tp, p, pp[SB_TPP_P0], u, w), -(double)sg);
        g = dadd(g, dsub(gabs, grad));
      } break;
      default: break;
    }
  }
  return g;
}

static int64_t find_col(const oracle_problem* Q, int64_t row, int32_t col) {
  int64_t lo = Q->rowptr[row], hi = Q->rowptr[row + 1] - 1;
  if (!Q->cols_sorted) {
    for (int64_t k = lo; k <= hi; ++k)
      if (Q->colidx[k] == col) return k;
    return -1;
  }
  while (lo <= hi) {
    int64_t mid = (lo + hi) >> 1;
    int32_t c = Q->colidx[mid];
    if (c == col) return mid;
    if (c < col) lo = mid + 1; else hi = mid - 1;
  }
  return -1;
}

static int64_t find_bc(const oracle_problem* Q, int32_t dof) {
  int64_t lo = 0, hi = Q->n_bc - 1;
  while (lo <= hi) {
    int64_t mid = (lo + hi) >> 1;
    int32_t c = Q->bc_dofs[mid];
    if (c == dof) return mid;
    if (c < dof) lo = mid + 1; else hi = mid - 1;
  }
  return -1;
}

/* element tensor of one cell: A[L*L] (row-major), b[L] */
static void cell_tensor(const oracle_problem* Q, int64_t c, double* A, double* b) {
  const sb_model* M = &Q->model;
  const int P = Q->P, L = 15 * P, nb = M->n_bands;
  const int dofbands = M->bands_have_dofs ? nb : 0;
  const double* tp = Q->tag_params + (int64_t)Q->cell_tag[c] * SB_TAG_STRIDE;
  const double* xv = Q->cell_vertices + c * 6;
  const double J00 = xv[2] - xv[0], J01 = xv[4] - xv[0];
  const double J10 = xv[3] - xv[1], J11 = xv[5] - xv[1];
  const double detJ = J00 * J11 - J01 * J10;
  const double adet = fabs(detJ);
  double ul[MAXL];
  for (int i = 0; i < L; ++i) ul[i] = Q->u[Q->cell_dofs[c * L + i]];
  memset(A, 0, sizeof(double) * L * L);
  memset(b, 0, sizeof(double) * L);
  double w0[SB_MAX_BANDS] = {0, 0, 0, 0};
  if (Q->base_w && dofbands)
    for (int k = 0; k < nb; ++k) w0[k] = Q->base_w[(int64_t)k * Q->n_cells + c];

#define PHYS_BDM(rule, q, i, vx, vy, dv)                                  \
  do {                                                                     \
    double rx = (rule)->bdm[((q) * 12 + (i)) * 2 + 0];                     \
    double ry = (rule)->bdm[((q) * 12 + (i)) * 2 + 1];                     \
    vx = (J00 * rx + J01 * ry) / detJ;                                     \
    vy = (J10 * rx + J11 * ry) / detJ;                                     \
    dv = (rule)->bdmdiv[(q) * 12 + (i)] / detJ;                            \
  } while (0)

  /* ---- default-degree group (all polynomial terms) ---------------------- */
  {
    const oracle_rule* R = &Q->rule_default;
    for (int q = 0; q < R->n; ++q) {
      const double wq = R->w[q] * adet;
      const double* l = R->p1 + q * 3;
      double bx[12], by[12], bd[12];
      for (int i = 0; i < 12; ++i) PHYS_BDM(R, q, i, bx[i], by[i], bd[i]);
      for (int p = 0; p < P; ++p) {
        const int o = 15 * p;
        const int inside = (p == 0) ? 1 : (tp[SB_TP_BAND0 + 5 * (p - 1) + SB_TPB_IN_SUBDOMAIN] != 0.0);
        /* field values */
        double s = 0, vx = 0, vy = 0, dv = 0;     /* scalar (phi | delta_w), flux (E | j) */
        for (int i = 0; i < 3; ++i) s += ul[o + i] * l[i];
        for (int i = 0; i < 12; ++i) { vx += ul[o + 3 + i] * bx[i]; vy += ul[o + 3 + i] * by[i]; dv += ul[o + 3 + i] * bd[i]; }
        if (p == 0) {
          /* + E.psi - phi div psi + w div E */
          for (int i = 0; i < 12; ++i) {
            b[o + 3 + i] += wq * ((vx * bx[i] + vy * by[i]) - s * bd[i]);
            for (int j = 0; j < 12; ++j) A[(o + 3 + i) * L + o + 3 + j] += wq * (bx[i] * bx[j] + by[i] * by[j]);
            for (int j = 0; j < 3; ++j) A[(o + 3 + i) * L + o + j] += -wq * bd[i] * l[j];
          }
          for (int i = 0; i < 3; ++i) {
            b[o + i] += wq * l[i] * dv;
            for (int j = 0; j < 12; ++j) A[(o + i) * L + o + 3 + j] += wq * l[i] * bd[j];
          }
        } else if (inside) {
          /* + v div j ; + div xi delta_w */
          for (int i = 0; i < 3; ++i) {
            b[o + i] += wq * l[i] * dv;
            for (int j = 0; j < 12; ++j) A[(o + i) * L + o + 3 + j] += wq * l[i] * bd[j];
          }
          for (int i = 0; i < 12; ++i) {
            b[o + 3 + i] += wq * bd[i] * s;
            for (int j = 0; j < 3; ++j) A[(o + 3 + i) * L + o + j] += wq * bd[i] * l[j];
          }
        } else {
          /* outside the band's subdomain: delta_w v c1 ; j.xi c2 */
          for (int i = 0; i < 3; ++i) {
            b[o + i] += wq * M->ext_w_coef * s * l[i];
            for (int j = 0; j < 3; ++j) A[(o + i) * L + o + j] += wq * M->ext_w_coef * l[i] * l[j];
          }
          for (int i = 0; i < 12; ++i) {
            b[o + 3 + i] += wq * M->ext_j_coef * (vx * bx[i] + vy * by[i]);
            for (int j = 0; j < 12; ++j) A[(o + 3 + i) * L + o + 3 + j] += wq * M->ext_j_coef * (bx[i] * bx[j] + by[i] * by[j]);
          }
        }
      }
    }
  }

  /* ---- rho term: - w rho/eps at quad_degree_rho --------------------------- */
  {
    const oracle_rule* R = &Q->rule_rho;
    for (int q = 0; q < R->n; ++q) {
      const double wq = R->w[q] * adet;
      const double* l = R->p1 + q * 3;
      dual phi = dc(0.0); phi.d[0] = 1.0;
      for (int i = 0; i < 3; ++i) phi.v += ul[i] * l[i];
      dual rho = dc(tp[SB_TP_RHO_STATIC]);
      for (int k = 0; k < nb; ++k) {
        dual w = dc(0.0);
        if (dofbands) {
          w.v = w0[k]; w.d[1 + k] = 1.0;
          for (int i = 0; i < 3; ++i) w.v += ul[15 * (1 + k) + i] * l[i];
        }
        dual u = band_u(M, tp, k, dadd(phi, w));
        rho = dadd(rho, dscale(u, (double)M->band_sign[k]));
      }
      dual erho = dscale(rho, 1.0 / tp[SB_TP_EPS]);
      for (int i = 0; i < 3; ++i) {
        b[i] -= wq * l[i] * erho.v;
        for (int j = 0; j < 3; ++j) {
          A[i * L + j] -= wq * l[i] * erho.d[0] * l[j];
          for (int k = 0; k < dofbands; ++k)
            A[i * L + 15 * (1 + k) + j] -= wq * l[i] * erho.d[1 + k] * l[j];
        }
      }
    }
  }

  if (dofbands) {
    /* ---- drift-diffusion j term: xi . j / (mu u) at quad_degree_super ------- */
    const oracle_rule* R = &Q->rule_super;
    for (int q = 0; q < R->n; ++q) {
      const double wq = R->w[q] * adet;
      const double* l = R->p1 + q * 3;
      double bx[12], by[12], bd[12];
      for (int i = 0; i < 12; ++i) PHYS_BDM(R, q, i, bx[i], by[i], bd[i]);
      (void)bd;
      dual phi = dc(0.0); phi.d[0] = 1.0;
      for (int i = 0; i < 3; ++i) phi.v += ul[i] * l[i];
      for (int k = 0; k < nb; ++k) {
        if (tp[SB_TP_BAND0 + 5 * k + SB_TPB_IN_SUBDOMAIN] == 0.0) continue;
        const int o = 15 * (1 + k);
        dual w = dc(w0[k]); w.d[1 + k] = 1.0;
        for (int i = 0; i < 3; ++i) w.v += ul[o + i] * l[i];
        dual pq = dadd(phi, w);
        dual coef = ddiv(dc(M->dd_scale), dmul(band_mobility(M, tp, k, pq), band_u(M, tp, k, pq)));
        double jx = 0, jy = 0;
        for (int i = 0; i < 12; ++i) { jx += ul[o + 3 + i] * bx[i]; jy += ul[o + 3 + i] * by[i]; }
        for (int i = 0; i < 12; ++i) {
          const double xij = bx[i] * jx + by[i] * jy;
          b[o + 3 + i] += wq * xij * coef.v;
          for (int j = 0; j < 12; ++j)
            A[(o + 3 + i) * L + o + 3 + j] += wq * (bx[i] * bx[j] + by[i] * by[j]) * coef.v;
          for (int j = 0; j < 3; ++j) {
            A[(o + 3 + i) * L + j] += wq * xij * coef.d[0] * l[j];
            for (int m = 0; m < nb; ++m)
              A[(o + 3 + i) * L + 15 * (1 + m) + j] += wq * xij * coef.d[1 + m] * l[j];
          }
        }
      }
    }
    /* ---- generation term: - s e v g at quad_degree_g --------------------------- */
    R = &Q->rule_g;
    for (int q = 0; q < R->n; ++q) {
      const double wq = R->w[q] * adet;
      const double* l = R->p1 + q * 3;
      const double* n2 = R->p2 + q * 6;
      dual phi = dc(0.0); phi.d[0] = 1.0;
      for (int i = 0; i < 3; ++i) phi.v += ul[i] * l[i];
      dual w[SB_MAX_BANDS], u[SB_MAX_BANDS];
      double u0[SB_MAX_BANDS];
      double phi_te = 0.0;
      if (Q->thmeq_phi) for (int a = 0; a < 6; ++a) phi_te += Q->thmeq_phi[c * 6 + a] * n2[a];
      for (int k = 0; k < nb; ++k) {
        const int o = 15 * (1 + k);
        w[k] = dc(w0[k]); w[k].d[1 + k] = 1.0;
        for (int i = 0; i < 3; ++i) w[k].v += ul[o + i] * l[i];
        u[k] = band_u(M, tp, k, dadd(phi, w[k]));
        u0[k] = band_u(M, tp, k, dc(phi_te)).v;
      }
      double phi_clip[SB_MAX_FIELDS] = {0, 0, 0, 0};
      for (int f = 0; f < M->n_fields; ++f) {
        double v = 0.0;
        const double* pf = Q->phi_opt + ((int64_t)f * Q->n_cells + c) * 6;
        for (int a = 0; a < 6; ++a) v += pf[a] * n2[a];
        phi_clip[f] = (v < 0.0) ? 0.0 : v;
      }
      for (int k = 0; k < nb; ++k) {
        if (tp[SB_TP_BAND0 + 5 * k + SB_TPB_IN_SUBDOMAIN] == 0.0) continue;
        const int o = 15 * (1 + k);
        dual g = generation(M, tp, k, u, w, u0, phi_clip);
        const double f = -(double)M->band_sign[k] * M->e_charge * wq;
        for (int i = 0; i < 3; ++i) {
          b[o + i] += f * l[i] * g.v;
          for (int j = 0; j < 3; ++j) {
            A[(o + i) * L + j] += f * l[i] * g.d[0] * l[j];
            for (int m = 0; m < nb; ++m)
              A[(o + i) * L + 15 * (1 + m) + j] += f * l[i] * g.d[1 + m] * l[j];
          }
        }
      }
    }
    /* ---- base_w jump term on interior facets (this cell = '+' side) ------------ */
    for (int f = 0; f < 3; ++f) {
      const uint8_t mask = Q->cell_jump_mask ? Q->cell_jump_mask[c * 3 + f] : 0;
      if (!mask) continue;
      const int64_t cn = Q->cell_neighbors[c * 3 + f];
      /* outward orientation of the reference scaled normal on facet f */
      static const double ref_out[3] = {+1.0, -1.0, +1.0};
      const double sgn = ref_out[f] * (detJ > 0 ? 1.0 : -1.0);
      for (int k = 0; k < nb; ++k) {
        if (!(mask & (1u << k))) continue;
        const double jump = Q->base_w[(int64_t)k * Q->n_cells + cn] - w0[k];   /* w0(-) - w0(+) */
        for (int i = 0; i < 12; ++i) {
          double tr = 0.0;
          for (int q = 0; q < Q->n_facet_q; ++q)
            tr += Q->facet_w[q] * Q->facet_trace[(f * Q->n_facet_q + q) * 12 + i];
          b[15 * (1 + k) + 3 + i] -= jump * sgn * tr;
        }
      }
    }
  }
#undef PHYS_BDM
}

/* lower_bound on natbc_cell */
static int64_t natbc_begin(const oracle_problem* Q, int64_t c) {
  int64_t lo = 0, hi = Q->n_natbc;
  while (lo < hi) { int64_t mid = (lo + hi) >> 1; if (Q->natbc_cell[mid] < c) lo = mid + 1; else hi = mid; }
  return lo;
}

/*
 * Full assembly.  vals: [nnz] (zeroed here), b: [N] (zeroed here).
 * mode bit 0: assemble A; bit 1: assemble b.
 */
int pdd_oracle_assemble(const oracle_problem* Q, double* vals, double* b, int mode) {
  const int L = 15 * Q->P;
  const int64_t nnz = Q->rowptr[Q->n_dofs];
  if (vals && (mode & 1)) memset(vals, 0, sizeof(double) * nnz);
  if (b && (mode & 2)) memset(b, 0, sizeof(double) * Q->n_dofs);
  int nthreads = Q->n_threads > 0 ? Q->n_threads : 1;
  int err = 0;
#ifdef _OPENMP
#pragma omp parallel num_threads(nthreads)
#endif
  {
    double* A = (double*)malloc(sizeof(double) * L * L);
    double* bl = (double*)malloc(sizeof(double) * L);
#ifdef _OPENMP
#pragma omp for schedule(static)
#endif
    for (int64_t c = 0; c < Q->n_cells; ++c) {
      cell_tensor(Q, c, A, bl);
      /* exterior-facet natural terms (precomputed integrals) */
      for (int64_t k = natbc_begin(Q, c); k < Q->n_natbc && Q->natbc_cell[k] == c; ++k)
        bl[Q->natbc_row[k]] += Q->natbc_values[k];
      /* element-wise symmetric Dirichlet application with x0 (Appendix A9) */
      if (Q->n_bc) {
        for (int i = 0; i < L; ++i) {
          int64_t ib = find_bc(Q, Q->cell_dofs[c * L + i]);
          if (ib < 0) continue;
          const double gval = Q->u[Q->cell_dofs[c * L + i]] - Q->bc_values[ib];
          for (int r = 0; r < L; ++r) { bl[r] -= A[r * L + i] * gval; A[r * L + i] = 0.0; }
          for (int j = 0; j < L; ++j) A[i * L + j] = 0.0;
          A[i * L + i] = 1.0;
          bl[i] = gval;
        }
      }
      for (int i = 0; i < L; ++i) {
        const int64_t row = Q->cell_dofs[c * L + i];
        if (b && (mode & 2)) {
#ifdef _OPENMP
#pragma omp atomic
#endif
          b[row] += bl[i];
        }
        if (vals && (mode & 1)) {
          for (int j = 0; j < L; ++j) {
            const double v = A[i * L + j];
            if (v == 0.0) continue;
            int64_t pos = find_col(Q, row, Q->cell_dofs[c * L + j]);
            if (pos < 0) { err = 1; continue; }
#ifdef _OPENMP
#pragma omp atomic
#endif
            vals[pos] += v;
          }
        }
      }
    }
    free(A); free(bl);
  }
  return err ? -1 : 0;
}

/* element tensors of selected cells without insertion (for tests) */
int pdd_oracle_cell_tensors(const oracle_problem* Q, int64_t n, const int64_t* cells,
                            double* A_out, double* b_out) {
  const int L = 15 * Q->P;
  for (int64_t k = 0; k < n; ++k)
    cell_tensor(Q, cells[k], A_out + k * L * L, b_out + k * L);
  return 0;
}

int pdd_oracle_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
