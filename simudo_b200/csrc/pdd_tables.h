/*
 * pdd_tables.h -- packed constant tables shared by host packing code
 * (simudo_b200/fem/tables.py mirrors this struct with ctypes) and the kernels.
 * Content is produced by simudo_b200/fem/reference_element.py.
 */
#ifndef SB_PDD_TABLES_H
#define SB_PDD_TABLES_H

#include <stdint.h>

#define SB_QM 11            /* max collapsed-rule points per axis */
#define SB_QG (SB_QM * SB_QM)
#define SB_NPAIR 78         /* (i <= m) pairs of the 12 BDM2 functions */

/* factored collapsed Gauss-Jacobi rule + separable Dubiner factors at its
 * nodes: point (i, j) = (a_i (1 - b_j), b_j), weight wa_i wb_j,
 * Psi_(p,q)(i, j) = g_p(a_i) h_pq(b_j). */
typedef struct SbRule {
  int32_t m;
  int32_t pad_;
  double a[SB_QM];
  double wag[SB_QM][5];   /* wa_i * g_p(a_i), p = 0..4              */
  double g3[SB_QM][3];    /* g_p(a_i), p = 0..2 (unweighted)        */
  double b[SB_QM];
  double wbh[SB_QM][15];  /* wb_j * h_t(b_j), t in dubiner order    */
  double h6[SB_QM][6];    /* h_t(b_j), t < 6 (unweighted)           */
} SbRule;

typedef struct SbTables {
  double bdmC[12][2][6];  /* BDM2 nodal basis in Dubiner P2                      */
  double G22u[21][15];    /* int Psi_t pi_r pi_s, (r<=s) packed                  */
  double G21[6][3][10];   /* int Psi_t pi_r lam_l                                */
  double G11[3][3][6];    /* int Psi_t lam_i lam_j                               */
  double G1[3][3];        /* int Psi_t lam_i                                     */
  double Mhat[12][12][3]; /* int psi_i^a psi_j^b: (11, 12+21, 22)                */
  double Dhat[12][3];     /* int div psi_i lam_l                                 */
  double M1[3][3];        /* int lam_i lam_j                                     */
  double trI[3];          /* int_0^1 of the 3 edge trace functions               */
  double pad_;
  double D3[12][3][20];   /* sum_r bdmC[i][a][r] G21[r][l][t] at [i][l][10 a + t]:
                             <xi_i . j dc/dchi lam_l> |detJ| = sum D3[i][l][.] Q[.]  */
  SbRule rho;             /* rule of the rho term                                */
  SbRule sup;             /* rule of the xi.j/(mu u) term                        */
  /* generation rule, unfactored: point n, X, Y, weight*Psi_t, P2 Lagrange       */
  int32_t ng;
  int32_t pad2_;
  double gX[SB_QG];
  double gY[SB_QG];
  double gwpsi[SB_QG][6];
  double gp2[SB_QG][6];
} SbTables;

#endif
