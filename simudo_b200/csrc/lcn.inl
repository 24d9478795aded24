/* SPDX-License-Identifier: MIT
 * Local-charge-neutrality stage, included at the end
 * of pdd_kernels.cu.  The reference solves  <w, rho(phi) / eps> = 0  for a DG1
 * potential with the bands at zero quasi-Fermi level
 * (physics/poisson_drift_diffusion1.py1:89-125, solver parameters
 * physics/poisson_drift_diffusion.py:318-331).  Cells decouple: the Jacobian is one
 * 3 x 3 block per cell, so a whole Newton iteration -- residual, Jacobian, the
 * 3 x 3 solve, the clipped and damped update (NewtonSolverMaxDu,
 * fem/newton_solver.py:533-541) and the convergence metrics (:79-122) -- is one
 * thread per cell in one launch.  UFL's estimated degree for this form is 4: the
 * 6-point default rule.                                   */

__constant__ double c_lcn_w[6] = {
    0.109951743655322 / 2, 0.109951743655322 / 2, 0.109951743655322 / 2,
    0.223381589678011 / 2, 0.223381589678011 / 2, 0.223381589678011 / 2};
__constant__ double c_lcn_xy[6][2] = {
    {0.816847572980459, 0.091576213509771}, {0.091576213509771, 0.816847572980459},
    {0.091576213509771, 0.091576213509771}, {0.108103018168070, 0.445948490915965},
    {0.445948490915965, 0.108103018168070}, {0.445948490915965, 0.445948490915965}};

struct LcnModel {
  int nb;
  int kind[SB_MAX_BANDS];
  double sign[SB_MAX_BANDS];
};

__global__ void k_lcn_iteration(int64_t nc, const double* cell_vertices, const int32_t* cell_tag,
                                const double* tag_params, LcnModel Mo, double* phi,
                                const double* tol, double omega, double max_du,
                                double rel_tol, double abs_tol, double* du_out, double* b_out,
                                double* out) {
  double bmax = 0, dmax = 0, rsum = 0, cmax = 0, nanf = 0, cnt = 0;
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c < nc) {
    const double* xv = cell_vertices + c * 6;
    const double adet = fabs((xv[2] - xv[0]) * (xv[5] - xv[1]) - (xv[4] - xv[0]) * (xv[3] - xv[1]));
    const double* tp = tag_params + (int64_t)cell_tag[c] * SB_TAG_STRIDE;
    const double kT = tp[SB_TP_KT], eps = tp[SB_TP_EPS];
    double p[3] = {phi[c * 3], phi[c * 3 + 1], phi[c * 3 + 2]};
    double F[3] = {0, 0, 0}, J[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    for (int q = 0; q < 6; ++q) {
      const double l[3] = {1.0 - c_lcn_xy[q][0] - c_lcn_xy[q][1], c_lcn_xy[q][0], c_lcn_xy[q][1]};
      const double ph = l[0] * p[0] + l[1] * p[1] + l[2] * p[2];
      double rho = tp[SB_TP_RHO_STATIC], drho = 0.0;
      for (int b = 0; b < Mo.nb; ++b) {
        const double N = tp[SB_TP_BAND0 + 5 * b + SB_TPB_N], E = tp[SB_TP_BAND0 + 5 * b + SB_TPB_E];
        const double s = Mo.sign[b];
        double u, du;
        if (Mo.kind[b] == SB_BAND_NONDEGENERATE) {
          u = N * exp(-s * (ph - E) / kT);
          du = -s / kT * u;
        } else {
          const double e = exp(s * (ph - E) / kT);
          const double f = 1.0 / (1.0 + e);
          u = N * f;
          du = -s / kT * N * f * (e * f);
        }
        rho += s * u;
        drho += s * du;
      }
      const double wq = c_lcn_w[q] * adet / eps;
      for (int i = 0; i < 3; ++i) {
        F[i] += wq * rho * l[i];
        for (int j = 0; j < 3; ++j) J[i][j] += wq * drho * l[i] * l[j];
      }
    }
    /* 3 x 3 solve (Cramer; the block is a weighted P1 mass matrix) */
    const double det = J[0][0] * (J[1][1] * J[2][2] - J[1][2] * J[2][1])
                     - J[0][1] * (J[1][0] * J[2][2] - J[1][2] * J[2][0])
                     + J[0][2] * (J[1][0] * J[2][1] - J[1][1] * J[2][0]);
    double du[3];
    du[0] = (F[0] * (J[1][1] * J[2][2] - J[1][2] * J[2][1])
           - J[0][1] * (F[1] * J[2][2] - J[1][2] * F[2])
           + J[0][2] * (F[1] * J[2][1] - J[1][1] * F[2])) / det;
    du[1] = (J[0][0] * (F[1] * J[2][2] - J[1][2] * F[2])
           - F[0] * (J[1][0] * J[2][2] - J[1][2] * J[2][0])
           + J[0][2] * (J[1][0] * F[2] - F[1] * J[2][0])) / det;
    du[2] = (J[0][0] * (J[1][1] * F[2] - F[1] * J[2][1])
           - J[0][1] * (J[1][0] * F[2] - F[1] * J[2][0])
           + F[0] * (J[1][0] * J[2][1] - J[1][1] * J[2][0])) / det;
    for (int i = 0; i < 3; ++i) {
      double step = du[i];
      if (max_du > 0.0) step = fmin(fmax(step, -max_du), max_du);
      const double un = p[i] - omega * step;
      phi[c * 3 + i] = un;
      if (du_out) du_out[c * 3 + i] = du[i];
      if (b_out) b_out[c * 3 + i] = F[i];
      /* metrics against the updated potential, as in k_newton_update */
      if (isnan(un) || isnan(F[i])) nanf = 1.0;
      const double ab = fabs(F[i]), ad = fabs(du[i]);
      if (ab > bmax) bmax = ab;
      if (ad > dmax) dmax = ad;
      const double r = fabs(du[i] / un);
      if (!isnan(r)) { rsum += r * r; cnt += 1.0; }
      const double cc = ad / (fabs(un) * rel_tol + abs_tol * (tol ? tol[c * 3 + i] : 1.0));
      if (cc > cmax) cmax = cc;
      if (isnan(ad)) nanf = 1.0;
    }
  }
#ifdef SB_EMUL_BUILD
  if (c < nc) {                      /* emulated threads run concurrently: atomics */
    atomic_max_pos(out + 0, bmax); atomic_max_pos(out + 1, dmax); atomicAdd(out + 2, rsum);
    atomic_max_pos(out + 3, cmax); atomic_max_pos(out + 4, nanf); atomicAdd(out + 5, cnt);
  }
#else
  /* warp reduce, then one atomic per warp */
  for (int o = 16; o > 0; o >>= 1) {
    bmax = fmax(bmax, __shfl_xor_sync(0xffffffffu, bmax, o));
    dmax = fmax(dmax, __shfl_xor_sync(0xffffffffu, dmax, o));
    cmax = fmax(cmax, __shfl_xor_sync(0xffffffffu, cmax, o));
    nanf = fmax(nanf, __shfl_xor_sync(0xffffffffu, nanf, o));
    rsum += __shfl_xor_sync(0xffffffffu, rsum, o);
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  }
  if ((threadIdx.x & 31) == 0) {
    atomic_max_pos(out + 0, bmax); atomic_max_pos(out + 1, dmax); atomicAdd(out + 2, rsum);
    atomic_max_pos(out + 3, cmax); atomic_max_pos(out + 4, nanf); atomicAdd(out + 5, cnt);
  }
#endif
}

extern "C" int sb_lcn_iteration(const sb_model* model, int64_t n_cells,
                                const double* d_cell_vertices, const int32_t* d_cell_tag,
                                const double* d_tag_params, double* d_phi, const double* d_tol,
                                double omega, double max_du, double rel_tol, double abs_tol,
                                double* d_du, double* d_b, double* d_out, void* stream) {
  if (!model || !d_cell_vertices || !d_cell_tag || !d_tag_params || !d_phi || !d_out) {
    snprintf(g_err, sizeof g_err, "null argument"); return SB_ERR_ARG;
  }
  LcnModel Mo;
  Mo.nb = model->n_bands;
  for (int b = 0; b < SB_MAX_BANDS; ++b) {
    Mo.kind[b] = b < model->n_bands ? model->band_kind[b] : 0;
    Mo.sign[b] = b < model->n_bands ? (double)model->band_sign[b] : 0.0;
  }
  cudaStream_t s = (cudaStream_t)stream;
  CUDA_TRY(cudaMemsetAsync(d_out, 0, 6 * sizeof(double), s));
  SB_LAUNCH(k_lcn_iteration, (unsigned)((n_cells + 127) / 128), 128, 0, s, n_cells, d_cell_vertices,
            d_cell_tag, d_tag_params, Mo, d_phi, d_tol, omega, max_du, rel_tol, abs_tol, d_du,
            d_b, d_out);
  g_launches += 1;
  CUDA_TRY(cudaGetLastError());
  return SB_OK;
}
