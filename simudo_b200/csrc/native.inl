/*
 * native.inl -- the closed-form ("native") fast path of sb_assemble (included by
 * pdd_kernels.cu).
 *
 * Same arithmetic as the generic kernels (NewtonSolution.A / .b,
 * simudo/fem/newton_solver.py:61-77; forms poisson_drift_diffusion1.py1:42-47,264-303),
 * for problems whose dof numbering and CSR layout are the library's own:
 *
 *   numbering 'b200'  sub-problem major: cell c owns dofs 6 Nc p + 6 c + (0..2 DG1 | 3..5
 *                     interior BDM2); facet f owns dofs 6P Nc + 3 Nf p + 3 f + k
 *                     (fem/dofmap.py) -- with values stored in row order, everything the row
 *                     kernel of sub-problem p writes is one dense range of vals
 *   pattern 'native'  structural non-zeros only, the columns of a row in the emission
 *                     order of these kernels (fem/dofmap.py: _native_ranks_of_row):
 *                       DG1 row        [BDM(p) 0..11 | DG(q) q = 0..P-1]
 *                       interior row   [BDM(p) 0..11 | DG(p) | DG(0) if p > 0]
 *                       facet row      [3 dofs of the facet | run of cell A | run of cell B]
 *
 * so that (1) no dof map, row-pointer or rank table is read at all -- every address is a
 * closed form of the cell index and three facet ids; (2) everything a cell contributes to
 * a row is one contiguous run: the rows a cell owns (6P rows, one contiguous block of
 * vals) leave the SM as cp.async.bulk copies (1-D TMA), its runs in the facet rows are
 * flushed by the warp with lane = entry, no rank lookup; (3) the (facet dof, facet dof)
 * 3 x 3 blocks two cells share are written once, by the facet's first cell, as its own
 * block plus the neighbour's block -- k_moments_nat leaves every cell's blocks in the
 * moment scratch -- so no slot is zeroed and no atomic touches the Jacobian.
 *
 * sb_plan_create verifies the closed forms entry by entry (detect_native) and falls back
 * to the generic kernels otherwise ('ufc' numbering, dolfin / sorted patterns).
 */

/* moment scratch of the native path, [quantity][cell]:
 *   band k:   (k * SB_NAT_NA + q):  W 21 | D 36 | Rx 6 | R 3 | S 18
 *             W = packed int c pi_r pi_s; D[3 i + l] = <xi_i . j dc/dchi lam_l> (the
 *             (xi, delta_w) / (xi, phi) coupling of BDM row i); S = (facet, facet) blocks
 *   then      generation moments   (nb * SB_NAT_NA + k * SB_NG + t)
 *   then      S of the Poisson sub-problem, 18                                  */
#define SB_NAT_NA 84
#define SB_NAT_W 0
#define SB_NAT_D 21
#define SB_NAT_RX 57
#define SB_NAT_R 63
#define SB_NAT_S 66

/* tile-blocked: quantity q of cell c at ((c / 32) * NQT + q) * 32 + c % 32, NQT quantities
 * per cell -- everything a warp (32 cells) reads or writes is one contiguous block of
 * NQT * 256 bytes instead of NQT segments 8 * n_cells bytes apart (one DRAM page / TLB
 * entry per quantity)                                                                  */
#define SB_NQS 32                                   /* stride between quantities of a cell */
__device__ __forceinline__ int sb_nat_nqt(const PlanDev& pl) { return pl.nb * (SB_NAT_NA + SB_NG) + 18; }
__device__ __forceinline__ double* natQ_ptr(const PlanDev& pl, int q, int64_t c) {
  return pl.scratch + (((c >> 5) * sb_nat_nqt(pl) + q) << 5) + (c & 31);
}
__device__ __forceinline__ double* natA_ptr(const PlanDev& pl, int k, int64_t c) {
  return natQ_ptr(pl, k * SB_NAT_NA, c);
}
__device__ __forceinline__ double* natG_ptr(const PlanDev& pl, int k, int64_t c) {
  return natQ_ptr(pl, pl.nb * SB_NAT_NA + k * SB_NG, c);
}
__device__ __forceinline__ double* natS0_ptr(const PlanDev& pl, int64_t c) {
  return natQ_ptr(pl, pl.nb * (SB_NAT_NA + SB_NG), c);
}

/* (facet dof, facet dof) blocks of the cell's flux mass matrix <xi_i . xi_m c> / |detJ|:
 * S[6 f + sym3(k, k')] for i = 3f + k, m = 3f + k'.  W = the 21 packed values of
 * int c pi_r pi_s, or nullptr for c = 1 (plain BDM mass through Mhat).               */
__device__ __forceinline__ void sb_facet_blocks(const SbTables& T, const SbCellGeom& g,
                                                const double* W, double coef, double* S) {
  const double sc = coef * g.iad;
  if (!W) {
#pragma unroll
    for (int f = 0; f < 3; ++f)
#pragma unroll
      for (int k = 0; k < 3; ++k)
#pragma unroll
        for (int k2 = k; k2 < 3; ++k2) {
          const double* mh = T.Mhat[3 * f + k][3 * f + k2];
          S[6 * f + sb_sym3(k, k2)] = sc * (g.G11 * mh[0] + g.G12 * mh[1] + g.G22 * mh[2]);
        }
    return;
  }
#pragma unroll 1
  for (int f = 0; f < 3; ++f) {
    double Sf[6];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int i = 3 * f + k;
      double GY[2][6];
#pragma unroll
      for (int s2 = 0; s2 < 6; ++s2) {
        double y0 = 0.0, y1 = 0.0;
#pragma unroll
        for (int r = 0; r < 6; ++r) {
          y0 += T.bdmC[i][0][r] * W[sb_sym21(r, s2)];
          y1 += T.bdmC[i][1][r] * W[sb_sym21(r, s2)];
        }
        GY[0][s2] = (g.G11 * y0 + g.G12 * y1) * sc;
        GY[1][s2] = (g.G12 * y0 + g.G22 * y1) * sc;
      }
#pragma unroll
      for (int k2 = k; k2 < 3; ++k2) {
        const int m = 3 * f + k2;
        double v = 0.0;
#pragma unroll
        for (int s2 = 0; s2 < 6; ++s2)
          v = fma(GY[1][s2], T.bdmC[m][1][s2], fma(GY[0][s2], T.bdmC[m][0][s2], v));
        Sf[sb_sym3(k, k2)] = v;
      }
    }
#pragma unroll
    for (int t = 0; t < 6; ++t) S[6 * f + t] = Sf[t];
  }
}

/* ---- phase A, native: thread = (cell, band) -------------------------------------------- */
__global__ void __launch_bounds__(SB_TILE, SB_MB_MOM)
k_moments_nat(PlanDev pl, StateDev st, int nblk) {
  const int k = blockIdx.x / nblk;
  const int64_t c = (int64_t)(blockIdx.x % nblk) * SB_TILE + threadIdx.x;
  if (c >= pl.n_cells) return;
  const int P = pl.P;
  const int64_t nc = pl.n_cells;
  const SbCellGeom g = sb_geom(pl.cell_vertices + c * 6);
  const double* tp = pl.tag_params + (int64_t)pl.cell_tag[c] * SB_TAG_STRIDE;
  const SbBandK K = sb_band_consts(c_M, tp, k);
  const bool dd = tp[SB_TP_BAND0 + 5 * k + SB_TPB_IN_SUBDOMAIN] != 0.0;
  const double* uc = st.u + (int64_t)6 * c;                       /* sub-problem 0 */
  const double* uk = st.u + (int64_t)6 * nc * (1 + k) + 6 * c;     /* sub-problem 1 + k */
  double phi[3], dw[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) { phi[i] = uc[i]; dw[i] = uk[i]; }
  const double w0 = st.base_w[(int64_t)k * nc + c];
  double gc[2][6];
  if (dd) {
    double jc[2][6];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int r = 0; r < 6; ++r) jc[a][r] = 0.0;
#pragma unroll
    for (int m = 0; m < 12; ++m) {
      const double fm = (m < 9)
          ? st.u[pl.nat_dofF0 + (int64_t)3 * pl.nat_nf * (1 + k) + 3 * pl.nat_fid[c * 3 + m / 3] + m % 3]
          : uk[3 + (m - 9)];
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
    __shared__ double s_q[SB_NQ][SB_TILE];
    double* qs = &s_q[0][threadIdx.x];
    bool fastm = false;
    if (rho && dd && K.kind == SB_BAND_NONDEGENERATE) {
      sb_phaseA_band<11, 1>(c_T.sup, rho, dd, K, chi0, chix, chiy, gc, mC, Q, mR, mRx, qs);
      fastm = true;
    } else if (rho && dd && K.const_mob) {
      sb_phaseA_band<11, 2>(c_T.sup, rho, dd, K, chi0, chix, chiy, gc, mC, Q, mR, mRx, qs);
      fastm = true;
    }
    if (fastm) {
#pragma unroll
      for (int t = 0; t < SB_NQ; ++t) Q[t] = qs[t * SB_TILE];
    } else if (rho || dd) {
      sb_phaseA_band<11>(c_T.sup, rho, dd, K, chi0, chix, chiy, gc, mC, Q, mR, mRx);
    }
  } else if (same_rule) {
    if (rho || dd) sb_phaseA_band<0>(c_T.sup, rho, dd, K, chi0, chix, chiy, gc, mC, Q, mR, mRx);
  } else {
    if (rho) sb_phaseA_band<0>(c_T.rho, true, false, K, chi0, chix, chiy, gc, mC, Q, mR, mRx);
    if (dd) sb_phaseA_band<0>(c_T.sup, false, true, K, chi0, chix, chiy, gc, mC, Q, mR, mRx);
  }
  double* o = natA_ptr(pl, k, c);
#pragma unroll
  for (int t = 0; t < SB_NRX; ++t) o[(int64_t)(SB_NAT_RX + t) * SB_NQS] = mRx[t];
#pragma unroll
  for (int t = 0; t < SB_NR; ++t) o[(int64_t)(SB_NAT_R + t) * SB_NQS] = mR[t];
  double S[18];
  if (dd) {
    /* d_il = (1 / |detJ|) sum_t D3[i][l][t] Q[t]: the register-hungry contraction is done
       here, where nothing else is live, so that the row kernel never holds Q */
#pragma unroll 1
    for (int i = 0; i < 12; ++i)
#pragma unroll
      for (int l = 0; l < 3; ++l) {
        double d = 0.0;
#pragma unroll
        for (int t = 0; t < SB_NQ; ++t) d = fma(c_T.D3[i][l][t], Q[t], d);
        o[(int64_t)(SB_NAT_D + 3 * i + l) * SB_NQS] = d * g.iad;
      }
    double W[21];
#pragma unroll
    for (int e = 0; e < 21; ++e) {
      double v = 0.0;
#pragma unroll
      for (int t = 0; t < 15; ++t) v += c_T.G22u[e][t] * mC[t];
      W[e] = v;
      o[(int64_t)(SB_NAT_W + e) * SB_NQS] = v;
    }
    sb_facet_blocks(c_T, g, W, 1.0, S);
  } else {
    sb_facet_blocks(c_T, g, nullptr, c_M.ext_j_coef, S);
  }
#pragma unroll
  for (int t = 0; t < 18; ++t) o[(int64_t)(SB_NAT_S + t) * SB_NQS] = S[t];
  if (k == 0) {                                   /* Poisson <psi_i . psi_m> blocks */
    sb_facet_blocks(c_T, g, nullptr, 1.0, S);
    double* o0 = natS0_ptr(pl, c);
#pragma unroll
    for (int t = 0; t < 18; ++t) o0[(int64_t)t * SB_NQS] = S[t];
  }
}

/* ---- phase B, native: thread = cell ---------------------------------------------------- */
template <int NB, class MV = sb_model>
__global__ void __launch_bounds__(SB_TILE, SB_MB_GEN)
k_generation_nat(PlanDev pl, StateDev st) {
  const int64_t c = (int64_t)blockIdx.x * SB_TILE + threadIdx.x;
  if (c >= pl.n_cells) return;
  constexpr int P = NB + 1;
  const double* tp = pl.tag_params + (int64_t)pl.cell_tag[c] * SB_TAG_STRIDE;
  const double* uc = st.u + (int64_t)6 * c;
  double dgv[1 + SB_MAX_BANDS][3];
  double w0[SB_MAX_BANDS];
  bool inside_k[SB_MAX_BANDS];
#pragma unroll
  for (int i = 0; i < 3; ++i) dgv[0][i] = uc[i];
  bool any_inside = false;
#pragma unroll
  for (int k = 0; k < SB_MAX_BANDS; ++k) {
    w0[k] = 0.0; inside_k[k] = false;
#pragma unroll
    for (int i = 0; i < 3; ++i) dgv[1 + k][i] = 0.0;
    if (k < NB) {
      w0[k] = st.base_w[(int64_t)k * pl.n_cells + c];
      inside_k[k] = tp[SB_TP_BAND0 + 5 * k + SB_TPB_IN_SUBDOMAIN] != 0.0;
      any_inside = any_inside || inside_k[k];
#pragma unroll
      for (int i = 0; i < 3; ++i) dgv[1 + k][i] = uc[(int64_t)6 * pl.n_cells * (1 + k) + i];
    }
  }
  if (!any_inside) return;                        /* nobody reads this cell's g moments */
  double mG[NB > 0 ? NB : 1][SB_NG];
  /* MV: the model's process list as a type (static lists, pdd_device.cuh) or the interpreter */
  if constexpr (std::is_same<MV, sb_model>::value) {
    sb_phaseB_all<NB>(c_T, c_M, tp, dgv, w0, inside_k, st.thmeq ? st.thmeq + c * 6 : nullptr,
                      st.phi_opt ? st.phi_opt + c * 6 : nullptr, pl.n_cells * 6, mG, true);
  } else {
    const MV mv{c_M};
    sb_phaseB_all<NB>(c_T, mv, tp, dgv, w0, inside_k, st.thmeq ? st.thmeq + c * 6 : nullptr,
                      st.phi_opt ? st.phi_opt + c * 6 : nullptr, pl.n_cells * 6, mG, true);
  }
  const int64_t nc = pl.n_cells;
#pragma unroll
  for (int k = 0; k < NB; ++k) {
    if (!inside_k[k]) continue;
    double* o = natG_ptr(pl, k, c);
#pragma unroll
    for (int t = 0; t < 9 + 6 * NB; ++t) o[(int64_t)t * SB_NQS] = mG[k][t];
  }
}

/* ---- phase C, native --------------------------------------------------------------------
 * thread = cell (arithmetic), warp = 32 cells (stores).  Per sub-problem PS a cell emits
 *   group 0: its 3 DG rows, group 4: its 3 interior BDM rows -- one contiguous block of
 *            vals each, staged in final order with the 16-byte phase of the global address
 *            and copied out by cp.async.bulk;
 *   group 1..3: its run in the 3 rows of facet 0..2 (3 + RB entries when it is the facet's
 *            first cell, RB otherwise), staged [row][entry] and flushed by the warp with
 *            lane = entry: one store instruction covers up to two row runs.           */
/* experiment switch (tools/build_variant.sh -DSB_EXP_NOLOAD): the row kernels without their
   moment loads, to separate load time from compute and store time */
#ifdef SB_EXP_NOLOAD
#define SB_LDQ(x) (1.0 + 1e-3 * (double)lane)
#else
#define SB_LDQ(x) (x)
#endif
#define SB_NAT_FROW 18                              /* staging slots per facet row */
#define SB_NAT_FSTR 57                              /* block CSR: staging doubles per lane of a facet warp */
#ifndef SB_MB_ROWS_BSR
#define SB_MB_ROWS_BSR 2                            /* CTAs per SM of the block-CSR row kernels (255 registers: no spills; measured faster than 3 x 168) */
#endif
#define SB_NAT_STR(P) ((((3 * (12 + 3 * (P))) > 3 * SB_NAT_FROW ? (3 * (12 + 3 * (P))) : 3 * SB_NAT_FROW) + 2) | 1)

struct SbNatMeta {
  int64_t gb[3][32];        /* vals index of the cell's run in row k = 0 of facet f      */
  int32_t rn[3][32];        /* row length | entries per run << 16 (0: nothing to store)  */
};

/* warp flush of one staged facet group: cell after cell, lane = (row, entry) */
__device__ __forceinline__ void sb_nat_flush_facet(const SbNatMeta& meta, int f, const double* wst,
                                                   int str, double* __restrict__ vals, int lane) {
  const int k0 = lane / SB_NAT_FROW, e0 = lane - k0 * SB_NAT_FROW;           /* slot lane      */
  const int s1 = lane + 32;
  const int k1 = s1 / SB_NAT_FROW, e1 = s1 - k1 * SB_NAT_FROW;               /* slot lane + 32 */
  const bool has1 = s1 < 3 * SB_NAT_FROW;
#pragma unroll 4
  for (int i = 0; i < 32; ++i) {
    const int32_t rn = meta.rn[f][i];
    const int n = rn >> 16, rl = rn & 0xFFFF;
    if (n == 0) continue;                                   /* warp-uniform */
    double* base = vals + meta.gb[f][i];
    const double* src = wst + i * str;
    if (e0 < n) base[k0 * rl + e0] = src[lane];
    if (has1 && e1 < n) base[k1 * rl + e1] = src[s1];
  }
}

/* The 12 BDM rows of one cell and sub-problem PS: three facet groups (staged, flushed by
 * the warp with lane = entry) and the interior group (bulk copy).  One rolled loop over the
 * groups (small code: the fully unrolled variant, 11 k instructions, ran 1.6x slower out of
 * the instruction cache).  Per group the three rows are computed together:
 *   GY_k[b][s] = (coef / |detJ|) sum_a G_ab sum_r C_i[a][r] W_rs        i = 3 grp + k
 *   A_im       = sum_{b,s} GY_k[b][s] C_m[b][s]
 * so that the 12 constants of column m are fetched once for three rows (sm_100a FP64
 * instructions take no constant-bank operand: every table value costs a load).
 * MODE 0: Poisson (W = identity), 1: band, moment-driven (a lane outside the band's
 * subdomain runs it with W = identity, no divergence block, penalty coefficient: exactly
 * the out-of-subdomain form :288-289), 2: whole warp outside the subdomain.            */
template <int NB, int PS, int MODE, int LAY>
__device__ __forceinline__ void sb_nat_flux_part(
    const PlanDev& pl, const StateDev& st, double* __restrict__ vals, double* __restrict__ b,
    const SbNatMeta& meta, double* wst, double* stg, double* stg16, int lane, int64_t cc,
    bool active, bool bulk, int32_t finfo, const int32_t* fid, const SbCellGeom& g,
    const double* fl, const double* sl, bool inside, int64_t g0, int64_t dof0, int grp,
    double w0, const int32_t* nbr /* staged: base_w and the 3 neighbours of this cell */) {
  constexpr int P = NB + 1;
  constexpr int RDG = 12 + 3 * P;
  constexpr int RI = PS == 0 ? 15 : 18;
  constexpr int STR = SB_NAT_STR(P);
  constexpr int kb = PS > 0 ? PS - 1 : 0;
  const int64_t nc = pl.n_cells;
  const SbTables& T = c_T;
  const bool momdrv = (MODE == 1) && inside;      /* W, D from the moment scratch */
  double W[21];
  double jumpf[3] = {0.0, 0.0, 0.0};
  double sc = g.iad, divsign = (MODE == 0) ? -g.sdet : g.sdet;
  const double* sa_base = (PS == 0) ? natS0_ptr(pl, cc) : natA_ptr(pl, kb, cc) + (int64_t)SB_NAT_S * SB_NQS;
  const double* ob = natA_ptr(pl, kb, cc);
  if (MODE == 1) {
    if (inside) {
#pragma unroll
      for (int e = 0; e < 21; ++e) W[e] = SB_LDQ(ob[(int64_t)(SB_NAT_W + e) * SB_NQS]);
#pragma unroll
      for (int f = 0; f < 3; ++f) {
        if (f != grp) continue;
        if ((finfo >> (12 + 4 * f + kb)) & 1) {
          const int64_t cn = nbr[f];
          const double ref_out = (f == 1) ? -1.0 : 1.0;
          jumpf[f] = (st.base_w[(int64_t)kb * nc + cn] - w0) * ref_out * g.sdet;
        }
      }
    } else {
#pragma unroll
      for (int r = 0; r < 6; ++r)
#pragma unroll
        for (int s2 = r; s2 < 6; ++s2) W[sb_sym21(r, s2)] = (r == s2) ? 1.0 : 0.0;
      sc = c_M.ext_j_coef * g.iad;
      divsign = 0.0;
    }
  } else if (MODE == 2) {
    sc = c_M.ext_j_coef * g.iad;
    divsign = 0.0;
  }
  const double g11 = g.G11 * sc, g12 = g.G12 * sc, g22 = g.G22 * sc;

  {
    constexpr bool BSR = (LAY == 2);
    constexpr int RB = PS == 0 ? 12 : 15;
    const bool facet = grp < 3;
    const int bits = facet ? ((finfo >> (4 * grp)) & 3) : 2;
    const int lead = (bits & 2) ? 0 : 3;          /* run starts with the (facet, facet) entries */
    const int rowstr = facet ? SB_NAT_FROW : RI;
    /* vals index of what this thread copies out itself: the interior block, or (block CSR)
       its run in the facet's block row */
    int64_t g1 = g0 + 3 * RDG;
    if (BSR && facet) {
      const int32_t fd = grp == 0 ? fid[0] : (grp == 1 ? fid[1] : fid[2]);
      g1 = pl.nat_fbase[(int64_t)PS * pl.nat_nf + fd] + ((bits & 2) ? 3 * (3 + RB) : 0);
    }
    const int ph = (int)(g1 & 1);
    /* 16-byte phase of the bulk copy; CSR facet runs are flushed by the warp instead */
    double* st0 = (facet && !BSR) ? stg : stg16 + ph;
    /* staging index of (row k, position `slot` of the row's run): row-major run per row
       (CSR), or 3 x 3 blocks, block-major (block CSR) */
#define SB_NAT_IDX(k, slot) (BSR ? ((slot) / 3) * 9 + 3 * (k) + (slot) % 3 : (k) * rowstr + (slot))
    /* operands that come from memory: issued first, used last */
    double S[6] = {0, 0, 0, 0, 0, 0}, D[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    if (lead) {
#pragma unroll
      for (int t = 0; t < 6; ++t) S[t] = SB_LDQ(sa_base[(int64_t)(6 * grp + t) * SB_NQS]);
      if (bits == 0) {                            /* interior facet, this cell is its first cell */
        const int64_t cn = nbr[grp];
        const int f2 = (finfo >> (4 * grp + 2)) & 3;
        const double* sn = (PS == 0) ? natS0_ptr(pl, cn)
                                     : natA_ptr(pl, kb, cn) + (int64_t)SB_NAT_S * SB_NQS;
#pragma unroll
        for (int t = 0; t < 6; ++t) S[t] += SB_LDQ(sn[(int64_t)(6 * f2 + t) * SB_NQS]);
      }
    }
    if (momdrv) {
#pragma unroll
      for (int t = 0; t < 9; ++t) D[t] = SB_LDQ(ob[(int64_t)(SB_NAT_D + 9 * grp + t) * SB_NQS]);
    }
    double F[3] = {0.0, 0.0, 0.0};
    if (active) {
      double GY[3][12];
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const int i = 3 * grp + k;
        if (MODE == 1) {
#pragma unroll
          for (int s2 = 0; s2 < 6; ++s2) {
            double y0 = 0.0, y1 = 0.0;
#pragma unroll
            for (int r = 0; r < 6; ++r) {
              y0 = fma(T.bdmC[i][0][r], W[sb_sym21(r, s2)], y0);
              y1 = fma(T.bdmC[i][1][r], W[sb_sym21(r, s2)], y1);
            }
            GY[k][s2] = g11 * y0 + g12 * y1;
            GY[k][6 + s2] = g12 * y0 + g22 * y1;
          }
        } else {                                  /* W = identity */
#pragma unroll
          for (int s2 = 0; s2 < 6; ++s2) {
            const double y0 = T.bdmC[i][0][s2], y1 = T.bdmC[i][1][s2];
            GY[k][s2] = g11 * y0 + g12 * y1;
            GY[k][6 + s2] = g12 * y0 + g22 * y1;
          }
        }
      }
      sb_bulk_wait_read();                        /* an earlier bulk copy has left the slot */
#pragma unroll
      for (int m = 0; m < 12; ++m) {
        double cm[12];
#pragma unroll
        for (int q = 0; q < 12; ++q) cm[q] = T.bdmC[m][q / 6][q % 6];
        double v[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          double a = 0.0;
#pragma unroll
          for (int q = 0; q < 12; ++q) a = fma(GY[k][q], cm[q], a);
          v[k] = a;
          F[k] = fma(a, fl[m], F[k]);
        }
        const int own = facet && (m / 3 == grp);
        if (!own) {
          const int slot = lead + (facet ? (m < 3 * grp ? m : m - 3) : m);
#pragma unroll
          for (int k = 0; k < 3; ++k) st0[SB_NAT_IDX(k, slot)] = v[k];
        }
      }
      const int dg0 = lead + (facet ? 9 : 12);
#pragma unroll
      for (int k = 0; k < 3; ++k) {
#pragma unroll
        for (int l = 0; l < 3; ++l) {
          const double dv = divsign * T.Dhat[3 * grp + k][l];
          F[k] = fma(dv, sl[l], F[k]);
          st0[SB_NAT_IDX(k, dg0 + l)] = dv + D[3 * k + l];
          if (PS > 0) st0[SB_NAT_IDX(k, dg0 + 3 + l)] = D[3 * k + l];
        }
        if (lead) {
#pragma unroll
          for (int k2 = 0; k2 < 3; ++k2) st0[SB_NAT_IDX(k, k2)] = S[sb_sym3(k, k2)];
        }
        if (MODE == 1 && facet) {                 /* base_w jump: facet dofs */
          const double jf = grp == 0 ? jumpf[0] : (grp == 1 ? jumpf[1] : jumpf[2]);
          F[k] -= jf * T.trI[k];
        }
      }
      if (b) {
        if (facet) {
          const int32_t fd = grp == 0 ? fid[0] : (grp == 1 ? fid[1] : fid[2]);
          double* bd = b + pl.nat_dofF0 + (int64_t)3 * pl.nat_nf * PS + 3 * fd;
#pragma unroll
          for (int k = 0; k < 3; ++k) SB_ATOMIC_ADD(bd + k, F[k]);
        } else {
#pragma unroll
          for (int k = 0; k < 3; ++k) b[dof0 + 3 + k] = F[k];
        }
      }
    } else {
      sb_bulk_wait_read();
    }
#undef SB_NAT_IDX
    if (facet && !BSR) {
      __syncwarp();
#ifndef SB_EXP_NOFLUSH
      if (vals) sb_nat_flush_facet(meta, grp, wst, STR, vals, lane);
#endif
      __syncwarp();
    } else if (bulk) {
#ifndef SB_EXP_NOBULK
      sb_bulk_store_owned(vals + g1, stg16, ph, facet ? 3 * (RB + lead) : 3 * RI);
#endif
    }
  }
}

/* ---- staged inputs of one 32-cell tile ---------------------------------------------------
 * Everything whose address is a function of the cell index alone is copied global -> shared
 * with cp.async TWO tiles ahead (no register is held while the copy is in flight); ONE tile
 * ahead, when the facet ids and neighbours of that tile have landed, the lines the tile will
 * load through them (facet dofs of u, facet row bases, neighbour base_w and (facet, facet)
 * blocks) and the tile's own moment block are prefetched into L2.  The thread-per-cell
 * kernel spent 62 % of its warp time on long-scoreboard stalls, four to five dependent DRAM
 * round trips per tile (profiles/r2_s2_hot_lines.txt); now the chain is shared memory -> one
 * round of L2 hits.                                                                        */
#define SB_NIN_VERTS 0            /* double [32][6]                                        */
#define SB_NIN_UCELL 1536         /* double [32][6]: DG + interior dofs of sub-problem PS  */
#define SB_NIN_W0 3072            /* double [32]: base_w of band PS - 1                    */
#define SB_NIN_FID 3328           /* int32 [32][3]                                         */
#define SB_NIN_NBR 3712           /* int32 [32][3]                                         */
#define SB_NIN_FINFO 4096         /* int32 [32]                                            */
#define SB_NIN_SPECIAL 4224       /* int32 [32]                                            */
#define SB_NIN_TAG 4352           /* int32 [32]                                            */
#define SB_NIN_BYTES 4480
#define SB_NIN_STAGES 3

template <int P, int PS>
__device__ __forceinline__ void sb_nat_fetch(const PlanDev& pl, const StateDev& st,
                                             unsigned char* buf, int64_t tile, int tid) {
  constexpr int kb = PS > 0 ? PS - 1 : 0;
  const int64_t c0 = tile * 32, nc = pl.n_cells;
  if (tid < 96) {                                 /* 3 pieces per cell */
    const int i = tid / 3, part = tid - 3 * i;
    const int64_t cs = c0 + i < nc ? c0 + i : nc - 1;      /* ragged last tile: a valid cell */
    sb_cp_async16(buf + SB_NIN_VERTS + tid * 16, pl.cell_vertices + cs * 6 + part * 2);
    sb_cp_async4(buf + SB_NIN_FID + tid * 4, pl.nat_fid + cs * 3 + part);
    sb_cp_async4(buf + SB_NIN_NBR + tid * 4, pl.cell_neighbors + cs * 3 + part);
    const int32_t* src = part == 0 ? pl.nat_finfo : (part == 1 ? pl.cell_special : pl.cell_tag);
    sb_cp_async4(buf + SB_NIN_FINFO + part * 128 + i * 4, src + cs);
  } else if (PS > 0) {
    const int i = tid - 96;
    const int64_t cs = c0 + i < nc ? c0 + i : nc - 1;
    sb_cp_async8(buf + SB_NIN_W0 + i * 8, st.base_w + (int64_t)kb * nc + cs);
  }
  for (int q = tid; q < 192; q += SB_TILE) {
    const int i = q / 6, k = q - 6 * i;
    const int64_t cs = c0 + i < nc ? c0 + i : nc - 1;
    sb_cp_async8(buf + SB_NIN_UCELL + q * 8, st.u + (int64_t)6 * nc * PS + 6 * cs + k);
  }
}

/* L2 prefetch of what tile `tile` (its staged inputs are in `in`) will load with plain loads */
template <int NB, int PS>
__device__ __forceinline__ void sb_nat_prefetch(const PlanDev& pl, const StateDev& st,
                                                const unsigned char* in, int64_t tile,
                                                int role, int lane) {
  constexpr int P = NB + 1;
  constexpr int kb = PS > 0 ? PS - 1 : 0;
  const int64_t nc = pl.n_cells;
#ifdef SB_EXP_NOPF_DEP
  if (false) {
#else
  if (role < 3) {
#endif
    const int32_t fd = reinterpret_cast<const int32_t*>(in + SB_NIN_FID)[lane * 3 + role];
    const int32_t cn = reinterpret_cast<const int32_t*>(in + SB_NIN_NBR)[lane * 3 + role];
    const int32_t finfo = reinterpret_cast<const int32_t*>(in + SB_NIN_FINFO)[lane];
    const double* uf = st.u + pl.nat_dofF0 + (int64_t)3 * pl.nat_nf * PS + 3 * fd;
    sb_prefetch_l2(uf);
    sb_prefetch_l2(uf + 2);                       /* 24 bytes may straddle a line */
    sb_prefetch_l2(pl.nat_fbase + (int64_t)PS * pl.nat_nf + fd);
    if (cn >= 0) {
      if (PS > 0) sb_prefetch_l2(st.base_w + (int64_t)kb * nc + cn);
      if (((finfo >> (4 * role)) & 3) == 0) {     /* this cell writes the shared block */
        const int f2 = (finfo >> (4 * role + 2)) & 3;
        const double* sn = (PS == 0) ? natS0_ptr(pl, cn)
                                     : natA_ptr(pl, kb, cn) + (int64_t)SB_NAT_S * SB_NQS;
#pragma unroll
        for (int t = 0; t < 6; ++t) sb_prefetch_l2(sn + (int64_t)(6 * f2 + t) * SB_NQS);
      }
    }
#ifdef SB_EXP_NOPF_BULK
  } else if (false) {
#else
  } else if (role == 3 && lane == 0) {            /* the tile's own moment blocks: contiguous */
#endif
    const int64_t c0 = tile * 32;
    if (PS > 0) {
      sb_prefetch_l2_bulk(natA_ptr(pl, kb, c0), SB_NAT_NA * 32 * 8);
      sb_prefetch_l2_bulk(natG_ptr(pl, kb, c0), (9 + 6 * NB) * 32 * 8);
    } else {
#pragma unroll
      for (int kk = 0; kk < NB; ++kk)
        sb_prefetch_l2_bulk(natA_ptr(pl, kk, c0) + (int64_t)SB_NAT_RX * SB_NQS,
                            (SB_NRX + SB_NR) * 32 * 8);
      sb_prefetch_l2_bulk(natS0_ptr(pl, c0), 18 * 32 * 8);
    }
  }
}

template <int NB, int PS, int LAY>
__global__ void __launch_bounds__(SB_TILE, (LAY == 2 ? SB_MB_ROWS_BSR : SB_MB_ROWS))
k_rows_nat(PlanDev pl, StateDev st, double* __restrict__ vals, double* __restrict__ b, int ntile) {
#ifdef SB_EMUL_BUILD
  unsigned char* s_dyn = emul::dyn_smem();
#else
  extern __shared__ __align__(16) unsigned char s_dyn[];
#endif
  constexpr int P = NB + 1;
  constexpr int RDG = 12 + 3 * P;                 /* DG row length                      */
  constexpr int RI = PS == 0 ? 15 : 18;           /* interior BDM row length            */
  constexpr int RB = PS == 0 ? 12 : 15;           /* a cell's run in a facet row        */
  constexpr int STR = SB_NAT_STR(P);              /* staging doubles per lane (odd)     */
  constexpr int CBP = 3 * RDG + 3 * RI;           /* owned values per cell and sub-problem */
  constexpr int OFF_DG = PS == 0 ? 0 : (3 * RDG + 45) + (PS - 1) * (3 * RDG + 54);   /* of the ones before */
  constexpr int kb = PS > 0 ? PS - 1 : 0;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  /* staging: CSR: STR doubles per lane for every warp (+ flush metadata); block CSR: the
     three facet warps need 54 + 2 doubles per lane, warp 3 the DG block (3 RDG + 2) */
  const int lstr = (LAY == 2 && warp < 3) ? SB_NAT_FSTR : STR;
  double* wst = reinterpret_cast<double*>(s_dyn) +
                (LAY == 2 ? (size_t)warp * 32 * SB_NAT_FSTR : (size_t)warp * 32 * STR);
  SbNatMeta& meta = reinterpret_cast<SbNatMeta*>(s_dyn + sizeof(double) * 32 * STR * (SB_TILE / 32))[warp];
  unsigned char* s_in = s_dyn + (LAY == 2 ? sizeof(double) * 32 * (3 * SB_NAT_FSTR + STR)
                                          : (sizeof(double) * 32 * STR + sizeof(SbNatMeta)) * (SB_TILE / 32));
  double* stg = wst + lane * lstr;                /* this lane's slot                   */
  double* stg16 = stg + (lane & 1);               /* 16-byte aligned (odd strides)      */
  const int64_t nc = pl.n_cells;
  /* a CTA = 32 cells x 4 warps; warp `role` 0..2 emits the rows of facet `role`, warp 3 the
     DG and interior rows: four short, independent instruction streams per cell tile
     instead of one long one.  The CTA is persistent: tiles blockIdx.x, + gridDim.x, ... */
  const int role = warp;
  /* Tiles are handed out dynamically (one atomic per tile on a per-launch counter): CTAs that
     start late -- the boundary-cell kernel runs beside the row kernels on a second stream --
     take fewer tiles, and a small mesh (a strip of a decomposed mesh) does not end with some
     CTAs one tile behind.  Thread 0 claims the tile of iteration it + 3 while the CTA works
     on `it`; the claim travels through s_claim and the barrier at the top of the loop. */
  __shared__ int s_claim[2];
  int* ctr = pl.nat_tile_ctr + PS;
  if (threadIdx.x == 0) {
    s_claim[0] = atomicAdd(ctr, 1);
    s_claim[1] = atomicAdd(ctr, 1);
  }
  __syncthreads();
  int64_t tile = s_claim[0], tile1 = s_claim[1], tile2 = 0;
  __syncthreads();
  if (threadIdx.x == 0) s_claim[0] = atomicAdd(ctr, 1);          /* for iteration 0: tile of it + 2 */
  if (tile >= ntile) return;
  sb_nat_fetch<P, PS>(pl, st, s_in, tile, threadIdx.x);
  if (tile1 < ntile) sb_nat_fetch<P, PS>(pl, st, s_in + SB_NIN_BYTES, tile1, threadIdx.x);
  sb_cp_async_commit();
  for (int it = 0; tile < ntile; tile = tile1, tile1 = tile2, ++it) {
  sb_cp_async_wait_all();                         /* tiles `tile` and `tile1` have landed ...    */
  __syncthreads();                                /* ... for every thread; the previous tile is
                                                     consumed; the claim of thread 0 is visible  */
  tile2 = s_claim[it & 1];                        /* tile of iteration it + 2                    */
  if (threadIdx.x == 0 && tile2 < ntile) s_claim[(it + 1) & 1] = atomicAdd(ctr, 1);
  else if (threadIdx.x == 0) s_claim[(it + 1) & 1] = ntile;
  const unsigned char* in = s_in + (it % SB_NIN_STAGES) * SB_NIN_BYTES;
  if (tile1 < ntile) {
    sb_nat_prefetch<NB, PS>(pl, st, s_in + ((it + 1) % SB_NIN_STAGES) * SB_NIN_BYTES, tile1,
                            role, lane);
    if (tile2 < ntile)
      sb_nat_fetch<P, PS>(pl, st, s_in + ((it + 2) % SB_NIN_STAGES) * SB_NIN_BYTES, tile2,
                          threadIdx.x);
    sb_cp_async_commit();
  }
  const int64_t c = tile * 32 + lane;
  const int64_t cc = c < nc ? c : nc - 1;
  const bool active = (c < nc) && (reinterpret_cast<const int32_t*>(in + SB_NIN_SPECIAL)[lane] < 0);
  const int32_t finfo = reinterpret_cast<const int32_t*>(in + SB_NIN_FINFO)[lane];
  const bool inside = (PS == 0) ? true : (((finfo >> (24 + kb)) & 1) != 0);
  const int32_t* nbr = reinterpret_cast<const int32_t*>(in + SB_NIN_NBR) + lane * 3;
  const double w0c = reinterpret_cast<const double*>(in + SB_NIN_W0)[lane];
  int32_t fid[3];
#pragma unroll
  for (int f = 0; f < 3; ++f) fid[f] = reinterpret_cast<const int32_t*>(in + SB_NIN_FID)[lane * 3 + f];
  const SbCellGeom g = sb_geom(reinterpret_cast<const double*>(in + SB_NIN_VERTS) + lane * 6);
  const double* uc = reinterpret_cast<const double*>(in + SB_NIN_UCELL) + lane * 6;
  double fl[12], sl[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) { sl[i] = uc[i]; fl[9 + i] = uc[3 + i]; }
#pragma unroll
  for (int f = 0; f < 3; ++f) {
    const double* uf = st.u + pl.nat_dofF0 + (int64_t)3 * pl.nat_nf * PS + 3 * fid[f];
#pragma unroll
    for (int k = 0; k < 3; ++k) fl[3 * f + k] = uf[k];
  }
  /* flush metadata of this warp's facet group */
#pragma unroll
  for (int f = 0; f < 3; ++f) {
    if (f != role || LAY == 2) continue;          /* block CSR: no warp flush, no metadata */
    const int bits = (finfo >> (4 * f)) & 3;
    const bool bnd = bits & 1, second = bits & 2;
    const int rl = bnd ? (PS == 0 ? 15 : 18) : (PS == 0 ? 27 : 33);
    meta.gb[f][lane] = pl.nat_fbase[(int64_t)PS * pl.nat_nf + fid[f]] + (second ? 3 + RB : 0);
    const int n = (active && vals != nullptr) ? (second ? RB : RB + 3) : 0;
    meta.rn[f][lane] = rl | (n << 16);
  }
  /* vals index of the cell's DG block: the owned rows of sub-problem PS are one dense array
     [cell][3 RDG + 3 RI] */
  const int64_t g0 = nc * OFF_DG + cc * CBP;
  const int64_t dof0 = (int64_t)6 * nc * PS + 6 * cc;
  const bool bulk = active && vals != nullptr;

  /* ---- group 0: DG rows ----------------------------------------------------------- */
  if (role == 3) {
    const int ph = (int)(g0 & 1);
    double* sd = stg16 + ph;
    sb_bulk_wait_read();                          /* the previous tile's copy has left the slot */
    /* staging index of (row i, position k of the row): see SB_NAT_IDX */
#define SB_NAT_DGIDX(i, k) (LAY == 2 ? ((k) / 3) * 9 + 3 * (i) + (k) % 3 : (i) * RDG + (k))
    double F[3] = {0.0, 0.0, 0.0};
    if (PS == 0 || inside) {
      double mom[9 + 6 * (NB > 0 ? NB : 1)];     /* [0..2] P1 source | [3 + 6q ..] d/d(q) */
      double coef;
      if (PS == 0) {
#pragma unroll
        for (int t = 0; t < 9 + 6 * NB; ++t) mom[t] = 0.0;
#pragma unroll
        for (int kk = 0; kk < NB; ++kk) {
          const double* o = natA_ptr(pl, kk, cc);
#pragma unroll
          for (int t = 0; t < SB_NRX; ++t) {
            const double v = SB_LDQ(o[(int64_t)(SB_NAT_RX + t) * SB_NQS]);
            mom[3 + t] += v;
            mom[3 + 6 * (1 + kk) + t] = v;
          }
#pragma unroll
          for (int t = 0; t < SB_NR; ++t) mom[t] += SB_LDQ(o[(int64_t)(SB_NAT_R + t) * SB_NQS]);
        }
        const double* tp = pl.tag_params +
            (int64_t)reinterpret_cast<const int32_t*>(in + SB_NIN_TAG)[lane] * SB_TAG_STRIDE;
        mom[0] += tp[SB_TP_RHO_STATIC] / tp[SB_TP_EPS] * 0.70710678118654752440;
        coef = -g.adet;
      } else {
        const double* og = natG_ptr(pl, kb, cc);
#pragma unroll
        for (int t = 0; t < 9 + 6 * NB; ++t) mom[t] = SB_LDQ(og[(int64_t)t * SB_NQS]);
        coef = -(double)c_M.band_sign[kb] * c_M.e_charge * g.adet;
      }
      if (active) {
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          double Fi = 0.0;
#pragma unroll
          for (int m = 0; m < 12; ++m) {
            const double dv = g.sdet * c_T.Dhat[m][i];
            Fi += dv * fl[m];
            sd[SB_NAT_DGIDX(i, m)] = dv;
          }
#pragma unroll
          for (int t = 0; t < 3; ++t) Fi += coef * c_T.G1[i][t] * mom[t];
#pragma unroll
          for (int q = 0; q < 1 + NB; ++q)
#pragma unroll
            for (int l = 0; l < 3; ++l) {
              double v = 0.0;
#pragma unroll
              for (int t = 0; t < 6; ++t) v += c_T.G11[i][l][t] * mom[3 + 6 * q + t];
              sd[SB_NAT_DGIDX(i, 12 + 3 * q + l)] = coef * v;
            }
          F[i] = Fi;
        }
      }
    } else if (active) {                          /* band outside its subdomain (:285-286) */
      const double c1 = c_M.ext_w_coef * g.adet;
#pragma unroll
      for (int e = 0; e < 3 * RDG; ++e) sd[e] = 0.0;
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int l = 0; l < 3; ++l) {
          const double v = c1 * c_T.M1[i][l];
          F[i] += v * sl[l];
          sd[SB_NAT_DGIDX(i, 12 + 3 * PS + l)] = v;
        }
    }
    if (active && b) {
#pragma unroll
      for (int i = 0; i < 3; ++i) b[dof0 + i] = F[i];
    }
#undef SB_NAT_DGIDX
#ifndef SB_EXP_NOBULK
    if (bulk) sb_bulk_store_owned(vals + g0, stg16, ph, 3 * RDG);
#endif
  }

  /* ---- flux rows ------------------------------------------------------------------- *
   * PS = 0: plain BDM mass (MODE 0).  PS > 0: one moment-driven code path (MODE 1) for the
   * whole warp -- a lane outside the band's subdomain runs it with W = identity, Q = 0, no
   * divergence block and the penalty coefficient, which is exactly the out-of-subdomain
   * form (:288-289) -- except when no lane of the warp is inside: then the cheap MODE 2.
   * (A per-lane branch between the two would keep both operand sets live: spills.)       */
  if (PS == 0) {
    sb_nat_flux_part<NB, PS, 0, LAY>(pl, st, vals, b, meta, wst, stg, stg16, lane, cc, active, bulk,
                                finfo, fid, g, fl, sl, true, g0, dof0, role, w0c, nbr);
  } else {
    const bool any_in = __any_sync(0xffffffffu, inside && active);
    if (any_in)
      sb_nat_flux_part<NB, PS, 1, LAY>(pl, st, vals, b, meta, wst, stg, stg16, lane, cc, active, bulk,
                                  finfo, fid, g, fl, sl, inside, g0, dof0, role, w0c, nbr);
    else
      sb_nat_flux_part<NB, PS, 2, LAY>(pl, st, vals, b, meta, wst, stg, stg16, lane, cc, active, bulk,
                                  finfo, fid, g, fl, sl, false, g0, dof0, role, w0c, nbr);
  }
  }                                               /* tile loop */
  sb_bulk_wait_read();                            /* the slot may be reused once it was read */
}

/* Grid of a persistent row kernel: as many CTAs as are resident at once (occupancy x SMs of
 * the current device; looked up once per kernel and device), capped by the number of tiles. */
static unsigned nat_rows_grid(const void* fn, size_t smem, int ntile) {
#ifdef SB_EMUL_BUILD
  (void)fn; (void)smem;
  return (unsigned)(ntile < 3 ? ntile : 3);       /* several tiles per CTA + a ragged tail */
#else
  struct Entry { const void* fn; int dev; int resident; };
  static std::vector<Entry> cache;
  static std::mutex mtx;
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lk(mtx);
  for (const Entry& e : cache)
    if (e.fn == fn && e.dev == dev) return (unsigned)(ntile < e.resident ? ntile : e.resident);
  cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int per_sm = 0, sms = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, SB_TILE, smem);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int resident = (per_sm > 0 ? per_sm : 1) * (sms > 0 ? sms : 1);
  cache.push_back({fn, dev, resident});
  return (unsigned)(ntile < resident ? ntile : resident);
#endif
}

/* launch k_rows_nat<NB, PS> for PS = NB .. 0 */
template <int NB, int PS>
struct NatRowsLauncher {
  static void go(const PlanDev& d, const StateDev& sd, double* vals, double* b, unsigned nblk,
                 cudaStream_t s) {
    const size_t smem = (sizeof(double) * 32 * SB_NAT_STR(NB + 1) + sizeof(SbNatMeta)) * (SB_TILE / 32)
                        + SB_NIN_STAGES * SB_NIN_BYTES;
    const int ntile = (int)((d.n_cells + 31) / 32);               /* 32 cells per tile */
    if (d.native == 2) {
      const size_t smem2 = sizeof(double) * 32 * (3 * SB_NAT_FSTR + SB_NAT_STR(NB + 1))
                           + SB_NIN_STAGES * SB_NIN_BYTES;
      const unsigned grid = nat_rows_grid((const void*)k_rows_nat<NB, PS, 2>, smem2, ntile);
      SB_LAUNCH((k_rows_nat<NB, PS, 2>), grid, SB_TILE, smem2, s, d, sd, vals, b, ntile);
    } else {
      const unsigned grid = nat_rows_grid((const void*)k_rows_nat<NB, PS, 1>, smem, ntile);
      SB_LAUNCH((k_rows_nat<NB, PS, 1>), grid, SB_TILE, smem, s, d, sd, vals, b, ntile);
    }
    NatRowsLauncher<NB, PS - 1>::go(d, sd, vals, b, nblk, s);
  }
};
template <int NB>
struct NatRowsLauncher<NB, -1> {
  static void go(const PlanDev&, const StateDev&, double*, double*, unsigned, cudaStream_t) {}
};

/* ---- host: does the problem have the closed-form layout? ------------------------------- *
 * On success fills fid [nc][3], finfo [nc], fbase [nf] and returns true.                   */
/* bsr: the value layout is block CSR with 3 x 3 blocks (pattern 'bsr3'): same block rows
 * and block order, rowptr[3 T + k] = block row base + 3 k, rank = 9 (block rank) + column
 * within its triple.                                                                     */
static bool detect_native(const sb_model* model, const sb_problem* pr, bool bsr,
                          std::vector<int32_t>& fid, std::vector<int32_t>& finfo,
                          std::vector<int64_t>& fbase, int64_t* dofF0, int64_t* n_facets) {
  const int P = pr->P, L = 15 * P;
  const int64_t nc = pr->n_cells, N = pr->n_dofs;
  if (!model->bands_have_dofs || P < 2 || getenv("SIMUDO_B200_NO_NATIVE")) return false;
  const int64_t F0 = (int64_t)6 * P * nc;
  if (N <= F0 || (N - F0) % (3 * P) != 0) return false;
  const int64_t nf = (N - F0) / (3 * P);
  const int RDG = 12 + 3 * P;
  const int32_t* cd = pr->h_cell_dofs;
  const int64_t* rp = pr->h_rowptr;
  fid.assign((size_t)nc * 3, -1);
  finfo.assign((size_t)nc, 0);
  fbase.assign((size_t)nf * P, -1);
  /* dof numbering: sub-problem major (fem/dofmap.py 'b200') */
  for (int64_t c = 0; c < nc; ++c) {
    const int32_t* d = cd + c * L;
    for (int f = 0; f < 3; ++f) {
      const int64_t q = (int64_t)d[3 + 3 * f] - F0;
      if (q < 0 || q >= 3 * nf || q % 3 != 0) return false;
      fid[c * 3 + f] = (int32_t)(q / 3);
    }
    for (int p = 0; p < P; ++p) {
      for (int i = 0; i < 3; ++i) {
        if (d[15 * p + i] != 6 * nc * p + 6 * c + i) return false;
        if (d[15 * p + 12 + i] != 6 * nc * p + 6 * c + 3 + i) return false;
      }
      for (int f = 0; f < 3; ++f)
        for (int k = 0; k < 3; ++k)
          if (d[15 * p + 3 + 3 * f + k] != F0 + 3 * nf * p + (int64_t)3 * fid[c * 3 + f] + k) return false;
    }
  }
  /* facet sides: first cell = lower index; neighbour's local facet */
  std::vector<uint8_t> fbnd((size_t)nf, 0);
  for (int64_t c = 0; c < nc; ++c) {
    int32_t info = 0;
    for (int f = 0; f < 3; ++f) {
      const int64_t cn = pr->h_cell_neighbors[c * 3 + f];
      if (cn >= nc || cn == c) return false;
      if (cn < 0) { info |= 1 << (4 * f); fbnd[fid[c * 3 + f]] = 1; continue; }
      int f2 = -1;
      for (int q = 0; q < 3; ++q)
        if (fid[cn * 3 + q] == fid[c * 3 + f] && pr->h_cell_neighbors[cn * 3 + q] == c) f2 = q;
      if (f2 < 0) return false;
      if (cn < c) info |= 2 << (4 * f);
      info |= f2 << (4 * f + 2);
      info |= (int32_t)(pr->h_cell_jump_mask[c * 3 + f] & 0xF) << (12 + 4 * f);
    }
    const double* tpc = pr->h_tag_params + (size_t)pr->h_cell_tag[c] * SB_TAG_STRIDE;
    for (int k = 0; k < model->n_bands; ++k)
      if (tpc[SB_TP_BAND0 + 5 * k + SB_TPB_IN_SUBDOMAIN] != 0.0) info |= 1 << (24 + k);
    finfo[c] = info;
  }
  /* row pointers: per sub-problem the owned rows [cell][3 DG rows | 3 interior rows], after all
     of them per sub-problem the facet rows [facet][3 rows] */
  {
    int64_t off = 0;
    for (int p = 0; p < P; ++p) {
      const int ri = p == 0 ? 15 : 18;
      for (int64_t c = 0; c < nc; ++c) {
        for (int i = 0; i < 3; ++i)
          if (rp[6 * nc * p + 6 * c + i] != off + (bsr ? 3 * i : RDG * i)) return false;
        off += 3 * RDG;
        for (int i = 0; i < 3; ++i)
          if (rp[6 * nc * p + 6 * c + 3 + i] != off + (bsr ? 3 * i : ri * i)) return false;
        off += 3 * ri;
      }
    }
    for (int p = 0; p < P; ++p)
      for (int64_t f = 0; f < nf; ++f) {
        const int rl = fbnd[f] ? (p == 0 ? 15 : 18) : (p == 0 ? 27 : 33);
        fbase[(size_t)p * nf + f] = off;
        for (int k = 0; k < 3; ++k)
          if (rp[F0 + 3 * nf * p + 3 * f + k] != off + (bsr ? 3 * k : rl * k)) return false;
        off += 3 * rl;
      }
    if (rp[N] != off) return false;
  }
  /* rank patterns: once per (pattern id, side bits) */
  std::vector<uint8_t> checked((size_t)pr->n_patterns * 8, 0);
  for (int64_t c = 0; c < nc; ++c) {
    int side = 0;
    for (int f = 0; f < 3; ++f) if ((finfo[c] >> (4 * f)) & 2) side |= 1 << f;
    const int32_t q = pr->h_cell_pattern[c];
    if (q < 0 || q >= pr->n_patterns) return false;
    if (checked[(size_t)q * 8 + side]) continue;
    checked[(size_t)q * 8 + side] = 1;
    const uint8_t* pat = pr->h_patterns + (size_t)q * L * L;
    for (int i = 0; i < L; ++i) {
      const int p = i / 15, r = i % 15;
      const int fr = (r >= 3 && r < 12) ? (r - 3) / 3 : -1;
      const int nrun = p == 0 ? 12 : 15;
      for (int j = 0; j < L; ++j) {
        const int pj = j / 15, rj = j % 15;
        int want = 0xFF;
        if (rj >= 3) {                                   /* BDM column */
          if (pj == p) {
            const int m = rj - 3;
            if (fr < 0) want = m;
            else if (m / 3 == fr) want = m - 3 * fr;
            else want = 3 + (m < 3 * fr ? m : m - 3);
          }
        } else if (r < 3) {
          want = 12 + 3 * pj + rj;                       /* DG row: DG(q) of every q */
        } else if (pj == p || pj == 0) {
          want = (fr < 0 ? 12 : 12) + rj + ((pj == 0 && p > 0) ? 3 : 0);
        }
        if (want != 0xFF && fr >= 0 && ((side >> fr) & 1) && want >= 3) want += nrun;
        if (want != 0xFF && bsr) want = (want / 3) * 9 + want % 3;
        if (pat[i * L + j] != want) return false;
      }
    }
  }
  /* Dirichlet dofs: facet dofs only (boundary facets, or interior ones such as the bounds of
     an intermediate-band layer: both cells of such a facet go through k_rows_special, whose
     emitter gives the shared block the identity entries of both cells) */
  for (int64_t i = 0; i < pr->n_bc; ++i) {
    const int64_t d = pr->h_bc_dofs[i];
    if (d < F0) return false;
  }
  *dofF0 = F0;
  *n_facets = nf;
  return true;
}
