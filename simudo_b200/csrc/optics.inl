/* SPDX-License-Identifier: MIT
 * a10: optical sub-problems, included at the end of
 * pdd_kernels.cu (same translation unit: shares PlanDev, StateDev, c_M, upload()).
 *
 *   update_input   alpha on CG2 = L2 projection of the absorption coefficient
 *                  (physics/optical.py:260-274 -> fem/assign.py:12-47,
 *                  dolfin.project): k_opt_alpha_cells (12-point degree-6 rule,
 *                  thread = cell) + k_opt_gather_sum -> right-hand side; the mass
 *                  solve is sb_band_solve on the constant CG2 mass matrix.
 *   MSORTE         second-order radiative-transfer form (physics/optical1.py1:27-45)
 *                  with S = 0:  oint -(v Omega.n) alpha Phi
 *                             + int -(Omega.grad Phi)(Omega.grad v) + v Omega.grad(alpha Phi)
 *                  k_opt_element (thread = cell, 6 x 6 element matrix from constant
 *                  reference tensors) + k_opt_gather_sum into the fixed CG2 CSR +
 *                  k_opt_dirichlet (inflow rows, optical1.py1:57-61).
 *   update_output  Phi -> Phi_pddproj (optical.py:248-258): k_opt_to_cells, an index
 *                  gather into the [cell][6] P2 layout the PDD kernels read.
 *
 * Scatter is a fixed-order segmented sum (no atomics): every CSR entry / vector
 * entry owns a host-built list of the element slots that feed it, so results are
 * bit-reproducible.  The problem is small next to the PDD pass (Nv + Nf dofs per
 * field); the kernels are plain coalesced FP64, one launch each.                */

struct SbOpticsTables {
  double K[2][2][6][6];     /* int d_a N_j d_b N_i                         */
  double T[2][6][6][6];     /* int N_i d_a (N_k N_j)                       */
  double F[3][6][6][6];     /* int_0^1 N_i N_j N_k on reference edge f     */
  double q12_w[12];         /* degree-6 default rule (weights sum to 1/2)  */
  double q12_l[12][3];      /* P1 basis at the points                      */
  double q12_n[12][6];      /* P2 basis at the points                      */
};

__constant__ SbOpticsTables c_OT;
static const void* g_otables_owner = nullptr;

struct sb_optics {
  sb_plan* plan;
  SbOpticsTables tables;
  int64_t n_dofs, nnz;
  const int32_t* cell_dofs;      /* [n_cells][6] */
  const int64_t* msrc_ptr;       /* [nnz + 1]    -> element slots (c * 36 + 6 i + j) */
  const int64_t* msrc_idx;
  const int64_t* vsrc_ptr;       /* [n_dofs + 1] -> vector slots (c * 6 + i)         */
  const int64_t* vsrc_idx;
  const int64_t* rowptr;         /* [n_dofs + 1] */
  const int64_t* diag_pos;       /* [n_dofs]     */
  double* elem;                  /* [n_cells][36] scratch */
  double* loc;                   /* [n_cells][6]  scratch */
};

/* alpha of one optical field at a point of cell c (electro_optical_process.py:
 * 415-418 band-to-band: constant inside its absorption window; :464-480 IB:
 * sigma_opt * (u_I or N_I - u_I)); the per-(process, field) coefficient already
 * carries the window (lowering.py)                                              */
SB_HD double sb_opt_alpha_point(const sb_model& M, const double* tp, int field,
                                double phi, const double* w) {
  double a = 0.0;
  for (int p = 0; p < M.n_procs; ++p) {
    const double coef = tp[SB_TP_PROC0 + 8 * p + SB_TPP_FIELD0 + field];
    if (coef == 0.0) continue;
    if (M.proc_kind[p] == SB_PROC_RAD_BB) {
      a += coef;
    } else if (M.proc_kind[p] == SB_PROC_RAD_IB) {
      const int IB = M.proc_trap[p];
      const int RB = (M.proc_dst[p] == IB) ? M.proc_src[p] : M.proc_dst[p];
      const SbBandK K = sb_band_consts(M, tp, IB);
      double u, du;
      sb_band_u(K, phi + w[IB], u, du);
      const bool same = (M.band_sign[IB] == M.band_sign[RB]);
      a += coef * (same ? u : (K.N - u));
    }
  }
  return a;
}

__global__ void k_opt_alpha_cells(PlanDev pl, StateDev st, int field, double* loc) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= pl.n_cells) return;
  const SbCellGeom g = sb_geom(pl.cell_vertices + c * 6);
  const double* tp = pl.tag_params + (int64_t)pl.cell_tag[c] * SB_TAG_STRIDE;
  const int32_t* dofs = pl.cell_dofs + c * pl.L;
  double phi[3], dw[SB_MAX_BANDS][3], w0[SB_MAX_BANDS];
  for (int i = 0; i < 3; ++i) phi[i] = st.u[dofs[i]];
  for (int k = 0; k < SB_MAX_BANDS; ++k) {
    w0[k] = 0.0;
    for (int i = 0; i < 3; ++i) dw[k][i] = 0.0;
    if (k < pl.nb_dof) {
      w0[k] = st.base_w[(int64_t)k * pl.n_cells + c];
      for (int i = 0; i < 3; ++i) dw[k][i] = st.u[dofs[sb_ldg(1 + k, i)]];
    }
  }
  double acc[6] = {0, 0, 0, 0, 0, 0};
  for (int q = 0; q < 12; ++q) {
    const double* l = c_OT.q12_l[q];
    const double ph = l[0] * phi[0] + l[1] * phi[1] + l[2] * phi[2];
    double w[SB_MAX_BANDS];
    for (int k = 0; k < SB_MAX_BANDS; ++k)
      w[k] = w0[k] + l[0] * dw[k][0] + l[1] * dw[k][1] + l[2] * dw[k][2];
    const double a = c_OT.q12_w[q] * sb_opt_alpha_point(c_M, tp, field, ph, w);
    for (int i = 0; i < 6; ++i) acc[i] += a * c_OT.q12_n[q][i];
  }
  for (int i = 0; i < 6; ++i) loc[c * 6 + i] = g.adet * acc[i];
}

/* out[e] = sum of src[idx[ptr[e] .. ptr[e+1])] in list order */
__global__ void k_opt_gather_sum(int64_t n, const int64_t* ptr, const int64_t* idx,
                                 const double* src, double* out) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  double s = 0.0;
  for (int64_t k = ptr[e]; k < ptr[e + 1]; ++k) s += src[idx[k]];
  out[e] = s;
}

__global__ void k_opt_element(PlanDev pl, const int32_t* cell_dofs, const double* alpha,
                              double ox, double oy, double* elem) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= pl.n_cells) return;
  const double* xv = pl.cell_vertices + c * 6;
  const double J00 = xv[2] - xv[0], J01 = xv[4] - xv[0];
  const double J10 = xv[3] - xv[1], J11 = xv[5] - xv[1];
  const double det = J00 * J11 - J01 * J10;
  const double adet = fabs(det);
  /* omega_hat = J^-1 Omega: Omega.grad = sum_a omega_hat_a d/dX_a */
  const double oh[2] = {(J11 * ox - J01 * oy) / det, (-J10 * ox + J00 * oy) / det};
  double al[6];
  for (int k = 0; k < 6; ++k) al[k] = alpha[cell_dofs[c * 6 + k]];
  /* exterior facets: (Omega.n_f) |f| = -|det| omega_hat . grad_ref lambda_f */
  double fw[3];
  const double gl[3][2] = {{-1.0, -1.0}, {1.0, 0.0}, {0.0, 1.0}};
  for (int f = 0; f < 3; ++f)
    fw[f] = (pl.cell_neighbors[c * 3 + f] < 0) ? adet * (oh[0] * gl[f][0] + oh[1] * gl[f][1]) : 0.0;
  for (int i = 0; i < 6; ++i)
    for (int j = 0; j < 6; ++j) {
      double a = 0.0;
      for (int p = 0; p < 2; ++p)
        for (int q = 0; q < 2; ++q) a -= oh[p] * oh[q] * c_OT.K[p][q][i][j];
      double t0 = 0.0, t1 = 0.0, tf = 0.0;
      for (int k = 0; k < 6; ++k) {
        t0 += al[k] * c_OT.T[0][i][j][k];
        t1 += al[k] * c_OT.T[1][i][j][k];
        tf += al[k] * (fw[0] * c_OT.F[0][i][j][k] + fw[1] * c_OT.F[1][i][j][k]
                       + fw[2] * c_OT.F[2][i][j][k]);
      }
      elem[c * 36 + 6 * i + j] = adet * (a + oh[0] * t0 + oh[1] * t1) + tf;
    }
}

/* inflow rows: Phi = value (identity row), all other right-hand sides are zero */
__global__ void k_opt_dirichlet(int64_t n_bc, const int32_t* bc_dofs, const double* bc_vals,
                                const int64_t* rowptr, const int64_t* diag_pos,
                                double* vals, double* rhs) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_bc) return;
  const int32_t d = bc_dofs[i];
  for (int64_t k = rowptr[d]; k < rowptr[d + 1]; ++k) vals[k] = 0.0;
  vals[diag_pos[d]] = 1.0;
  rhs[d] = bc_vals[i];
}

__global__ void k_opt_to_cells(int64_t n, const int32_t* cell_dofs, const double* nodal,
                               double* out) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < n) out[e] = nodal[cell_dofs[e]];
}

static int bind_optics_tables(sb_optics* o, cudaStream_t s) {
  if (g_otables_owner == o) return SB_OK;
  CUDA_TRY(cudaMemcpyToSymbolAsync(c_OT, &o->tables, sizeof(SbOpticsTables), 0,
                                   cudaMemcpyHostToDevice, s));
  g_otables_owner = o;
  return SB_OK;
}

extern "C" int64_t sb_optics_tables_size(void) { return (int64_t)sizeof(SbOpticsTables); }

extern "C" int sb_optics_create(sb_plan* pl, const sb_optics_problem* pr, sb_optics** out) {
  if (!pl || !pr || !out || !pr->h_cell_dofs || !pr->h_pos || !pr->h_rowptr ||
      !pr->h_diag_pos || !pr->h_tables) {
    snprintf(g_err, sizeof g_err, "null argument"); return SB_ERR_ARG;
  }
  if (pr->n_tables != (int64_t)sizeof(SbOpticsTables)) {
    snprintf(g_err, sizeof g_err, "optics tables: %lld bytes given, %zu expected",
             (long long)pr->n_tables, sizeof(SbOpticsTables));
    return SB_ERR_ARG;
  }
  const int64_t nc = pl->d.n_cells;
  if (pr->n_cells != nc) { snprintf(g_err, sizeof g_err, "cell count mismatch"); return SB_ERR_ARG; }
  sb_optics* o = new sb_optics();
  o->plan = pl;
  memcpy(&o->tables, pr->h_tables, sizeof(SbOpticsTables));
  o->n_dofs = pr->n_dofs; o->nnz = pr->nnz;
  /* inverse maps: CSR entry <- element slots, dof <- vector slots (stable order) */
  std::vector<int64_t> mptr(pr->nnz + 1, 0), midx((size_t)nc * 36);
  std::vector<int64_t> vptr(pr->n_dofs + 1, 0), vidx((size_t)nc * 6);
  for (int64_t e = 0; e < nc * 36; ++e) {
    const int32_t p = pr->h_pos[e];
    if (p < 0 || p >= pr->nnz) { snprintf(g_err, sizeof g_err, "pos out of range"); delete o; return SB_ERR_ARG; }
    ++mptr[p + 1];
  }
  for (int64_t e = 0; e < nc * 6; ++e) {
    const int32_t d = pr->h_cell_dofs[e];
    if (d < 0 || d >= pr->n_dofs) { snprintf(g_err, sizeof g_err, "dof out of range"); delete o; return SB_ERR_ARG; }
    ++vptr[d + 1];
  }
  for (int64_t i = 0; i < pr->nnz; ++i) mptr[i + 1] += mptr[i];
  for (int64_t i = 0; i < pr->n_dofs; ++i) vptr[i + 1] += vptr[i];
  {
    std::vector<int64_t> fill(mptr.begin(), mptr.end() - 1);
    for (int64_t e = 0; e < nc * 36; ++e) midx[fill[pr->h_pos[e]]++] = e;
    std::vector<int64_t> fillv(vptr.begin(), vptr.end() - 1);
    for (int64_t e = 0; e < nc * 6; ++e) vidx[fillv[pr->h_cell_dofs[e]]++] = e;
  }
  int rc = SB_OK;
#define OUP(field, src, n) if (rc == SB_OK) rc = upload(pl, src, (size_t)(n), &o->field)
  OUP(cell_dofs, pr->h_cell_dofs, nc * 6);
  OUP(msrc_ptr, mptr.data(), mptr.size());
  OUP(msrc_idx, midx.data(), midx.size());
  OUP(vsrc_ptr, vptr.data(), vptr.size());
  OUP(vsrc_idx, vidx.data(), vidx.size());
  OUP(rowptr, pr->h_rowptr, pr->n_dofs + 1);
  OUP(diag_pos, pr->h_diag_pos, pr->n_dofs);
  const double* tmp = nullptr;
  if (rc == SB_OK) rc = upload(pl, (const double*)nullptr, (size_t)nc * 36, &tmp);
  o->elem = const_cast<double*>(tmp);
  if (rc == SB_OK) rc = upload(pl, (const double*)nullptr, (size_t)nc * 6, &tmp);
  o->loc = const_cast<double*>(tmp);
#undef OUP
  if (rc != SB_OK) { delete o; return rc; }   /* device memory is owned by the plan */
  *out = o;
  return SB_OK;
}

extern "C" void sb_optics_destroy(sb_optics* o) {
  if (!o) return;
  if (g_otables_owner == o) g_otables_owner = nullptr;
  delete o;
}

extern "C" int sb_optics_alpha_rhs(sb_optics* o, const sb_state* st, int32_t field,
                                   double* d_rhs, void* stream) {
  if (!o || !st || !st->d_u || !d_rhs || field < 0 || field >= o->plan->model.n_fields) {
    snprintf(g_err, sizeof g_err, "bad argument to sb_optics_alpha_rhs"); return SB_ERR_ARG;
  }
  sb_plan* pl = o->plan;
  if (pl->d.nb_dof && !st->d_base_w) {
    snprintf(g_err, sizeof g_err, "base_w required when bands have dofs"); return SB_ERR_ARG;
  }
  cudaStream_t s = (cudaStream_t)stream;
  int rc = bind_tables(pl, s);
  if (rc == SB_OK) rc = bind_optics_tables(o, s);
  if (rc != SB_OK) return rc;
  StateDev sd{st->d_u, st->d_base_w, st->d_thmeq_phi, st->d_phi_opt, st->d_bc_values,
              st->d_natbc_values};
  const int64_t nc = pl->d.n_cells;
  SB_LAUNCH(k_opt_alpha_cells, (unsigned)((nc + 127) / 128), 128, 0, s, pl->d, sd, (int)field, o->loc);
  SB_LAUNCH(k_opt_gather_sum, (unsigned)((o->n_dofs + 255) / 256), 256, 0, s, o->n_dofs,
            o->vsrc_ptr, o->vsrc_idx, (const double*)o->loc, d_rhs);
  g_launches += 2;
  CUDA_TRY(cudaGetLastError());
  return SB_OK;
}

extern "C" int sb_optics_assemble(sb_optics* o, const double* d_alpha, double ox, double oy,
                                  int64_t n_bc, const int32_t* d_bc_dofs,
                                  const double* d_bc_vals, double* d_vals, double* d_rhs,
                                  void* stream) {
  if (!o || !d_alpha || !d_vals || !d_rhs || (n_bc > 0 && (!d_bc_dofs || !d_bc_vals))) {
    snprintf(g_err, sizeof g_err, "bad argument to sb_optics_assemble"); return SB_ERR_ARG;
  }
  sb_plan* pl = o->plan;
  cudaStream_t s = (cudaStream_t)stream;
  int rc = bind_optics_tables(o, s);
  if (rc != SB_OK) return rc;
  const int64_t nc = pl->d.n_cells;
  SB_LAUNCH(k_opt_element, (unsigned)((nc + 127) / 128), 128, 0, s, pl->d, o->cell_dofs, d_alpha,
            ox, oy, o->elem);
  SB_LAUNCH(k_opt_gather_sum, (unsigned)((o->nnz + 255) / 256), 256, 0, s, o->nnz, o->msrc_ptr,
            o->msrc_idx, (const double*)o->elem, d_vals);
  CUDA_TRY(cudaMemsetAsync(d_rhs, 0, (size_t)o->n_dofs * sizeof(double), s));
  if (n_bc > 0)
    SB_LAUNCH(k_opt_dirichlet, (unsigned)((n_bc + 127) / 128), 128, 0, s, n_bc, d_bc_dofs,
              d_bc_vals, o->rowptr, o->diag_pos, d_vals, d_rhs);
  g_launches += n_bc > 0 ? 3 : 2;
  CUDA_TRY(cudaGetLastError());
  return SB_OK;
}

extern "C" int sb_optics_to_cells(sb_optics* o, const double* d_nodal, double* d_cells,
                                  void* stream) {
  if (!o || !d_nodal || !d_cells) { snprintf(g_err, sizeof g_err, "null argument"); return SB_ERR_ARG; }
  const int64_t n = o->plan->d.n_cells * 6;
  SB_LAUNCH(k_opt_to_cells, (unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream, n,
            o->cell_dofs, d_nodal, d_cells);
  g_launches += 1;
  CUDA_TRY(cudaGetLastError());
  return SB_OK;
}
