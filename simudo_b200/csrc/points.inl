/*
 * points.inl -- point evaluation of the solution / line cuts (included by
 * pdd_kernels.cu).
 *
 * Replaces Function_eval_many (simudo/fem/extra_bindings.py:27-75: for every point
 * `mesh.bounding_box_tree()->compute_first_entity_collision(point)`, then
 * `function.eval(values, x, cell, ufc_cell)`; a point outside the mesh is skipped), the
 * evaluator behind LineCutCsvPlot (simudo/io/csv.py:54-63, 5001 points along the mid line).
 *
 * Point location: a uniform grid of bins over the mesh's bounding box, built once per mesh
 * on the host (bin -> cells whose bounding box touches it, ascending cell index); a thread
 * takes one point, walks the candidates of its bin and stops at the first cell containing
 * the point (barycentric test with tolerance 1e-12), i.e. the containing cell with the
 * lowest index -- on a cell boundary DG functions are two-valued and dolfin's tree order is
 * not reproducible, this choice is deterministic.  Evaluation: DG1 = barycentric
 * interpolation of the cell's 3 dofs; BDM2 = contravariant Piola map of the reference basis
 * (bdmC: nodal basis in the Dubiner P2 basis, the table of the assembly kernels; the Dubiner
 * functions at an arbitrary point through their monomial matrix h_dubiner_mono).
 */
struct SbLocDev {
  double x0, y0, ihx, ihy;
  int32_t nbx, nby;
  const int32_t* bin_ptr;        /* [nbx nby + 1] */
  const int32_t* bin_cells;
  double mono[6][6];             /* Psi_r = sum_k mono[r][k] (1, X, Y, X^2, XY, Y^2)_k */
};
struct sb_locator {
  sb_plan* plan;
  SbLocDev d;
};

__global__ void k_eval_points(PlanDev pl, SbLocDev lc, const double* __restrict__ u, int p,
                              int what, int64_t n, const double* __restrict__ xy,
                              double* __restrict__ out, int32_t* __restrict__ cell_out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double x = xy[2 * i], y = xy[2 * i + 1];
  int bx = (int)floor((x - lc.x0) * lc.ihx), by = (int)floor((y - lc.y0) * lc.ihy);
  /* a point on the upper edge of the box belongs to the last bin; anything further out
     still looks at the nearest bin and fails the containment test */
  bx = bx < 0 ? 0 : (bx >= lc.nbx ? lc.nbx - 1 : bx);
  by = by < 0 ? 0 : (by >= lc.nby ? lc.nby - 1 : by);
  const int b = by * lc.nbx + bx;
  const double tol = 1e-12;
  int64_t found = -1;
  double X = 0.0, Y = 0.0, J00 = 0, J01 = 0, J10 = 0, J11 = 0, det = 1.0;
  for (int q = lc.bin_ptr[b]; q < lc.bin_ptr[b + 1]; ++q) {
    const int64_t c = lc.bin_cells[q];
    const double* xv = pl.cell_vertices + c * 6;
    J00 = xv[2] - xv[0]; J01 = xv[4] - xv[0];
    J10 = xv[3] - xv[1]; J11 = xv[5] - xv[1];
    det = J00 * J11 - J01 * J10;
    const double dx = x - xv[0], dy = y - xv[1];
    X = (J11 * dx - J01 * dy) / det;
    Y = (J00 * dy - J10 * dx) / det;
    if (X >= -tol && Y >= -tol && X + Y <= 1.0 + tol) { found = c; break; }
  }
  if (cell_out) cell_out[i] = (int32_t)found;
  if (found < 0) return;                                    /* "If not found, just skip" */
  const int32_t* dofs = pl.cell_dofs + found * pl.L + 15 * p;
  if (what == 0) {                                          /* DG1 scalar */
    out[i] = u[dofs[0]] * (1.0 - X - Y) + u[dofs[1]] * X + u[dofs[2]] * Y;
    return;
  }
  const double mo[6] = {1.0, X, Y, X * X, X * Y, Y * Y};
  double psi[6];
#pragma unroll
  for (int r = 0; r < 6; ++r) {
    double v = 0.0;
#pragma unroll
    for (int k = 0; k < 6; ++k) v += lc.mono[r][k] * mo[k];
    psi[r] = v;
  }
  double h0 = 0.0, h1 = 0.0;                                /* reference-space vector */
  for (int m = 0; m < 12; ++m) {
    const double um = u[dofs[3 + m]];
    double a0 = 0.0, a1 = 0.0;
#pragma unroll
    for (int r = 0; r < 6; ++r) {
      a0 += c_T.bdmC[m][0][r] * psi[r];
      a1 += c_T.bdmC[m][1][r] * psi[r];
    }
    h0 += um * a0; h1 += um * a1;
  }
  out[2 * i] = (J00 * h0 + J01 * h1) / det;                 /* contravariant Piola */
  out[2 * i + 1] = (J10 * h0 + J11 * h1) / det;
}

extern "C" int sb_locator_create(sb_plan* pl, const double* h_cell_vertices, int32_t nbx,
                                 int32_t nby, const double* h_dubiner_mono, sb_locator** out) {
  if (!pl || !h_cell_vertices || !h_dubiner_mono || !out) {
    snprintf(g_err, sizeof g_err, "bad argument to sb_locator_create"); return SB_ERR_ARG;
  }
  const int64_t nc = pl->d.n_cells;
  double lo[2] = {1e300, 1e300}, hi[2] = {-1e300, -1e300};
  for (int64_t i = 0; i < nc * 3; ++i)
    for (int a = 0; a < 2; ++a) {
      const double v = h_cell_vertices[2 * i + a];
      lo[a] = v < lo[a] ? v : lo[a]; hi[a] = v > hi[a] ? v : hi[a];
    }
  if (nbx <= 0 || nby <= 0) {                    /* about one cell per bin, aspect of the box */
    const double ex = hi[0] - lo[0], ey = hi[1] - lo[1];
    const double side = sqrt(ex * ey / (double)(nc > 0 ? nc : 1));
    nbx = (int32_t)std::min<double>(4096.0, std::max<double>(1.0, ex / (side > 0 ? side : 1.0)));
    nby = (int32_t)std::min<double>(4096.0, std::max<double>(1.0, ey / (side > 0 ? side : 1.0)));
  }
  const double hx = (hi[0] - lo[0]) / nbx, hy = (hi[1] - lo[1]) / nby;
  if (!(hx > 0.0) || !(hy > 0.0)) {
    snprintf(g_err, sizeof g_err, "sb_locator_create: degenerate mesh bounding box"); return SB_ERR_ARG;
  }
  const double pad = 1e-9;
  auto range = [&](int64_t c, int a, int nb, double l0, double h, int& b0, int& b1) {
    double mn = 1e300, mx = -1e300;
    for (int v = 0; v < 3; ++v) {
      const double t = h_cell_vertices[c * 6 + 2 * v + a];
      mn = t < mn ? t : mn; mx = t > mx ? t : mx;
    }
    b0 = (int)floor((mn - l0) / h - pad); b1 = (int)floor((mx - l0) / h + pad);
    b0 = b0 < 0 ? 0 : (b0 >= nb ? nb - 1 : b0);
    b1 = b1 < 0 ? 0 : (b1 >= nb ? nb - 1 : b1);
  };
  std::vector<int32_t> ptr((size_t)nbx * nby + 1, 0);
  for (int64_t c = 0; c < nc; ++c) {
    int x0, x1, y0, y1;
    range(c, 0, nbx, lo[0], hx, x0, x1); range(c, 1, nby, lo[1], hy, y0, y1);
    for (int by = y0; by <= y1; ++by)
      for (int bx = x0; bx <= x1; ++bx) ++ptr[(size_t)by * nbx + bx + 1];
  }
  for (size_t b = 0; b < (size_t)nbx * nby; ++b) {
    if ((int64_t)ptr[b] + ptr[b + 1] > 0x7fffffff) {
      snprintf(g_err, sizeof g_err, "sb_locator_create: bin lists exceed int32"); return SB_ERR_UNSUPPORTED;
    }
    ptr[b + 1] += ptr[b];
  }
  std::vector<int32_t> cells((size_t)ptr.back());
  std::vector<int32_t> fill(ptr.begin(), ptr.end() - 1);
  for (int64_t c = 0; c < nc; ++c) {             /* ascending cell index inside every bin */
    int x0, x1, y0, y1;
    range(c, 0, nbx, lo[0], hx, x0, x1); range(c, 1, nby, lo[1], hy, y0, y1);
    for (int by = y0; by <= y1; ++by)
      for (int bx = x0; bx <= x1; ++bx) cells[(size_t)fill[(size_t)by * nbx + bx]++] = (int32_t)c;
  }
  sb_locator* lc = new sb_locator();
  lc->plan = pl;
  lc->d.x0 = lo[0]; lc->d.y0 = lo[1]; lc->d.ihx = 1.0 / hx; lc->d.ihy = 1.0 / hy;
  lc->d.nbx = nbx; lc->d.nby = nby;
  memcpy(lc->d.mono, h_dubiner_mono, sizeof lc->d.mono);
  const int32_t *dp = nullptr, *dc = nullptr;
  int rc = upload<int32_t>(pl, ptr.data(), ptr.size(), &dp);     /* released with the plan */
  if (rc == SB_OK) rc = upload<int32_t>(pl, cells.data(), cells.size() ? cells.size() : 1, &dc);
  if (rc != SB_OK) { delete lc; return rc; }
  lc->d.bin_ptr = dp; lc->d.bin_cells = dc;
  *out = lc;
  return SB_OK;
}

extern "C" void sb_locator_destroy(sb_locator* lc) { delete lc; }

extern "C" int sb_eval_points(sb_locator* lc, const double* d_u, int32_t sub_problem, int32_t what,
                              int64_t n_points, const double* d_xy, double* d_values,
                              int32_t* d_cell, void* stream) {
  if (!lc || !d_u || !d_xy || !d_values || n_points < 0 || sub_problem < 0 ||
      sub_problem >= lc->plan->d.P || (what != 0 && what != 1)) {
    snprintf(g_err, sizeof g_err, "bad argument to sb_eval_points"); return SB_ERR_ARG;
  }
  if (n_points == 0) return SB_OK;
  cudaStream_t s = (cudaStream_t)stream;
  int rc = bind_tables(lc->plan, s);
  if (rc != SB_OK) return rc;
  SB_LAUNCH(k_eval_points, (unsigned)((n_points + 127) / 128), 128, 0, s, lc->plan->d, lc->d, d_u,
            (int)sub_problem, (int)what, n_points, d_xy, d_values, d_cell);
  g_launches += 1;
  CUDA_TRY(cudaGetLastError());
  return SB_OK;
}
