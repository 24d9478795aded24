/*
 * simudo_b200.h -- C ABI of the B200 assembly hot path.
 *
 * The reference (simudo 0.6.5) has no FFI for this path: its seam is the
 * Python duck-typed one between simudo.fem.NewtonSolver and
 * dolfin.SystemAssembler / PETScLUSolver.  Each entry point below names the
 * reference interface it replaces (paths relative to /root/reference).
 * INTEGRATION.md shows the ctypes binding a Simudo maintainer would add.
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / CUDA types in signatures;
 *   - "h_" = host pointer, "d_" = device pointer (cudaMalloc'ed, e.g. a
 *     torch tensor's data_ptr()); all device work is enqueued on `stream`
 *     (a cudaStream_t cast to void*, NULL = legacy default stream) and is
 *     stream-ordered, nothing synchronises unless stated;
 *   - return 0 on success, negative sb_status on failure; the Python host maps
 *     failures to RuntimeError exactly like dolfin/PETSc errors surface in the
 *     reference (simudo/fem/adaptive_stepper.py:212-217);
 *   - every number is in "form units": length in mesh_unit (um), potential V,
 *     energy eV, current A, time s, charge C, densities 1/um^3
 *     (simudo/fem/delayed_form.py:80-94 folds pint units into the same
 *     dimensionless form values).
 */
#ifndef SIMUDO_B200_H
#define SIMUDO_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SB_MAX_BANDS   4
#define SB_MAX_PROCS   8
#define SB_MAX_FIELDS  4
#define SB_MAX_QUAD_M  11      /* collapsed Gauss-Jacobi points per axis    */

/* band statistics: simudo/physics/poisson_drift_diffusion1.py1:170-245 */
enum { SB_BAND_NONDEGENERATE = 0, SB_BAND_INTERMEDIATE = 1 };

/* generation / recombination processes: simudo/physics/electro_optical_process.py */
enum {
  SB_PROC_SRH      = 0,  /* SRHRecombination                 :495-536 */
  SB_PROC_SR_TRAP  = 1,  /* NonRadiativeTrap                 :539-562 */
  SB_PROC_RAD_BB   = 2,  /* NonOverlappingTopHatBeerLambert  :398-431 */
  SB_PROC_RAD_IB   = 3   /* NonOverlappingTopHatBeerLambertIB:434-492 */
};

typedef enum {
  SB_OK = 0,
  SB_ERR_ARG = -1,
  SB_ERR_CUDA = -2,
  SB_ERR_ALLOC = -3,
  SB_ERR_NOT_FINITE = -4,
  SB_ERR_UNSUPPORTED = -5
} sb_status;

/* Per cell-tag parameter row (doubles).  One row per distinct cell value of
 * MeshData.cell_function; replaces the conditional tree that Spatial.get()
 * builds on the DG0 tag function (simudo/fem/spatial.py:350-376,
 * simudo/fem/expr.py:106-153; missing rule => 0.0). */
#define SB_TP_KT          0   /* k T / e                                [eV]        */
#define SB_TP_EPS         1   /* permittivity                [e / (V um)]           */
#define SB_TP_RHO_STATIC  2   /* poisson/static_rho          [e / um^3]             */
#define SB_TP_BAND0       3   /* + 5*b : band b                                     */
#define   SB_TPB_IN_SUBDOMAIN 0 /* 1 inside band.subdomain, 0 in subdomain_inv     */
#define   SB_TPB_N            1 /* eff. DOS or number_of_states [1/um^3]           */
#define   SB_TPB_E            2 /* (effective) energy level     [eV]               */
#define   SB_TPB_MOBILITY     3 /* mobility / mobility0         [um^2/(V s)]       */
#define   SB_TPB_SPARE        4
#define SB_TP_PROC0       (SB_TP_BAND0 + 5 * SB_MAX_BANDS)  /* + 8*p : process p   */
#define   SB_TPP_P0           0 /* SRH tau(dst)  | trap capture coeff [um^3/s]     */
                                /* RAD_BB alpha*I_strandberg [1/(um^3 s)]          */
                                /* RAD_IB pseudo capture coeff sigma_opt*I/u1      */
#define   SB_TPP_P1           1 /* SRH tau(src) [s]                                */
#define   SB_TPP_P2           2 /* SRH u1(dst) [1/um^3]                            */
#define   SB_TPP_P3           3 /* SRH u1(src) [1/um^3]                            */
#define   SB_TPP_FIELD0       4 /* + f : absorption of optical field f inside its  */
                                /* window: RAD_BB alpha [1/um]; RAD_IB sigma [um^2]*/
#define SB_TAG_STRIDE     (SB_TP_PROC0 + 8 * SB_MAX_PROCS)

/* Physics description (what PoissonDriftDiffusion + bands + processes lower to). */
typedef struct sb_model {
  int32_t n_bands;          /* bands entering rho, in pdd.bands order               */
  int32_t bands_have_dofs;  /* 1: goal 'full' (MixedQfl unknowns, P = 1 + n_bands); */
                            /* 0: goal 'thermal equilibrium' (qfl == 0, P = 1)      */
  int32_t n_procs;
  int32_t n_fields;         /* optical fields feeding g (Phi_pddproj, CG2 per cell) */
  int32_t n_tags;
  int32_t quad_m_rho;       /* (mixed_debug_quad_degree_rho   + 2)/2, default 11    */
  int32_t quad_m_super;     /* (mixedqfl_debug_quad_degree_super + 2)/2, default 11 */
  int32_t quad_m_g;         /* (mixedqfl_debug_quad_degree_g  + 2)/2, default 5     */
  int32_t band_kind[SB_MAX_BANDS];
  int32_t band_sign[SB_MAX_BANDS];            /* -1 electrons, +1 holes            */
  int32_t band_const_mobility[SB_MAX_BANDS];  /* IB: use_constant_mobility         */
  int32_t proc_kind[SB_MAX_PROCS];
  int32_t proc_dst[SB_MAX_PROCS];             /* band indices, -1 = none           */
  int32_t proc_src[SB_MAX_PROCS];
  int32_t proc_trap[SB_MAX_PROCS];
  int32_t proc_radiative[SB_MAX_PROCS];       /* enable_radiative_recombination    */
  double  e_charge;         /* elementary charge [C]: -s e <v g> factor            */
  double  dd_scale;         /* unit factor of <xi.j/(mu u)>  (= 1/e_charge)        */
  double  ext_w_coef;       /* external_qfl_term constant  (poisson_drift_diffusion1.py1:285-286) */
  double  ext_j_coef;       /* external_j_term constant    (:288-289)              */
} sb_model;

/* Mesh + space + boundary description handed to sb_plan_create (host pointers,
 * copied to the device).  Replaces dolfin Mesh/DofMap/SparsityPattern/
 * DirichletBC construction inside SystemAssembler(J, F, bcs)
 * (simudo/fem/newton_solver.py:293-304). */
typedef struct sb_problem {
  int64_t n_cells;
  int64_t n_dofs;              /* N                                                */
  int32_t P;                   /* sub-problems; L = 15 P local dofs                */
  const double*  h_cell_vertices; /* [n_cells][3][2] vertex-sorted cell geometry   */
  const int32_t* h_cell_tag;      /* [n_cells] row index into tag_params           */
  const int32_t* h_cell_dofs;     /* [n_cells][L]                                  */
  const int32_t* h_cell_neighbors;/* [n_cells][3] cell across local facet f, -1    */
  const uint8_t* h_cell_jump_mask;/* [n_cells][3] bit b set: base_w jump term of   */
                                  /* band b on this (interior) facet, and this     */
                                  /* cell is its '+' side                          */
  const int64_t* h_rowptr;        /* [N+1] CSR row pointers (dolfin or structural) */
  /* position of local entry (i, j) inside global row cell_dofs[c][i]:           */
  /* vals index = rowptr[row] + patterns[cell_pattern[c]][i][j]; 0xFF = the      */
  /* entry is absent (structural zero dropped from a compacted pattern)          */
  int64_t n_patterns;
  const uint8_t* h_patterns;      /* [n_patterns][L][L]                            */
  const int32_t* h_cell_pattern;  /* [n_cells]                                     */
  const double*  h_tag_params;    /* [n_tags][SB_TAG_STRIDE]                       */
  /* essential (Dirichlet) conditions: constrained global dofs, sorted ascending   */
  int64_t n_bc;
  const int32_t* h_bc_dofs;       /* [n_bc]                                        */
  /* cell-local right-hand-side additions (natural BC terms): static sparsity      */
  int64_t n_natbc;
  const int32_t* h_natbc_cell;    /* [n_natbc]                                     */
  const int32_t* h_natbc_row;     /* [n_natbc] local dof index 0..L-1              */
  /* reference-element constant tables, see simudo_b200/fem/reference_element.py   */
  const double* h_tables;         /* packed; layout = sb_tables_layout()           */
  int64_t n_tables;
} sb_problem;

typedef struct sb_plan sb_plan;   /* opaque */

/* Per-assembly inputs (device pointers). */
typedef struct sb_state {
  const double* d_u;          /* [N] current solution u_ (mixed function vector)    */
  const double* d_base_w;     /* [n_bands][n_cells] mixedqfl_base_w (DG0), or NULL  */
  const double* d_thmeq_phi;  /* [n_cells][6] thermal_equilibrium_phi (DG2), or NULL*/
  const double* d_phi_opt;    /* [n_fields][n_cells][6] Phi_pddproj per cell (P2)   */
  const double* d_bc_values;  /* [n_bc] Dirichlet value per constrained dof         */
  const double* d_natbc_values;/* [n_natbc] values added to the local residual      */
} sb_state;

/* ---- lifetime ---------------------------------------------------------------- */
int  sb_plan_create(const sb_model* model, const sb_problem* problem, sb_plan** out);
void sb_plan_destroy(sb_plan* plan);
const char* sb_last_error(void);
int  sb_version(void);
/* number of kernel launches issued through this library since load */
int64_t sb_launch_count(void);

/* ---- a1 + a2: Jacobian and residual in one traversal ---------------------------
 * Replaces NewtonSolution.A / .b, i.e. SystemAssembler.assemble(A) and
 * .assemble(b, x0=u_) (simudo/fem/newton_solver.py:48-77): cell integrals,
 * exterior-facet natural terms, the base_w interior-facet jump term, and the
 * element-wise symmetric Dirichlet application with x0 lifting.
 * d_vals [nnz] (dolfin pattern; structural zeros are written once by
 * sb_plan_zero_values and never touched again), d_b [N].  Either may be NULL.
 * d_vals must be 16-byte aligned (bulk copies; SB_ERR_ARG otherwise) -- any cudaMalloc'ed
 * array is, a slice at an odd element offset is not.
 * Stream semantics: everything is ordered after the work already on `stream` and complete
 * for what is enqueued on `stream` afterwards; internally the boundary-cell kernel runs on a
 * side stream of the plan, forked and joined with events.  One pass of a plan at a time:
 * do not call sb_assemble on the same plan from two streams concurrently (the plan owns the
 * moment scratch and the tile counters); different plans are independent.          */
int sb_assemble(sb_plan* plan, const sb_state* state,
                double* d_vals, double* d_b, void* stream);
int sb_plan_zero_values(sb_plan* plan, double* d_vals, void* stream);
int64_t sb_plan_nnz(const sb_plan* plan);
/* 1 when the plan runs the closed-form kernels (csrc/native.inl): dof numbering 'b200' and
 * CSR pattern 'native' of simudo_b200/fem/dofmap.py (structural non-zeros of the dolfin
 * pattern, columns of a row in emission order), Dirichlet dofs on facet dofs only (boundary or interior facets).
 * Any other numbering / pattern handed to sb_plan_create runs the generic kernels.      */
int sb_plan_is_native(const sb_plan* plan);
/* Which instance of the generation kernel (a6: electro_optical_process.py:141-157,294-312,
 * 420-431,482-492,516-536) a native plan runs: 0 = the process list is interpreted per
 * quadrature point; 1 = the list of the four-layer IB cell (bands CB, IB, VB; opt_ci, opt_iv,
 * opt_cv, nr_top, nr_bottom: example/fourlayer/fourlayer.py:537-564) and 2 = the pn diode with
 * SRH (example/pn_diode/pn_diode.py:161-164) are compiled as straight-line code.  Same
 * arithmetic either way (SIMUDO_B200_GEN_INTERPRET=1 at plan creation forces 0).          */
int sb_plan_generation_signature(const sb_plan* plan);
/* materialise the CSR column indices (int32) on the device: for export/solvers */
int sb_plan_column_indices(sb_plan* plan, int32_t* d_colidx, void* stream);

/* ---- a8: natural boundary conditions (exterior-facet integrals) ------------------------
 * Replaces the `ds(fv)` terms Spatial builds for natural BCs (simudo/fem/spatial.py:
 * 237-273; poisson_drift_diffusion1.py1:47: + oint phi_bc psi . n): for every listed
 * exterior facet (cell, local facet) the three integrals  oint g psi_k . n ds  of the
 * facet's BDM2 dofs with g = d_c0[i] + sum_a d_p2[i][a] N2_a (P2 nodal values of the
 * non-constant part on the facet's cell; either pointer may be NULL), 3 Gauss-Legendre
 * points.  Results are written to d_natbc_values[d_slot[3 i + k]] -- the array
 * sb_state.d_natbc_values of the next sb_assemble -- so a bias sweep only updates d_c0.
 * h_tables: 84 doubles {w[3], N2[3 facets][3 points][6], trace[3 facets][3 dofs][3 points]}
 * (simudo_b200/fem/reference_element.py: facet_gl_w, facet_p2, facet_bdm_trace).     */
int sb_natbc_eval(sb_plan* plan, int64_t n_facets, const int32_t* d_cell,
                  const int8_t* d_local, const int32_t* d_slot, const double* d_c0,
                  const double* d_p2, const double* h_tables, int64_t n_tables,
                  double* d_natbc_values, void* stream);

/* ---- a9: qfl rebasing ----------------------------------------------------------
 * Replaces MixedQflBand.mixedqfl_do_rebase_w (poisson_drift_diffusion1.py1:339-383)
 * with default thresholds (0,0): base_w += cellavg(delta_w), BC-adjacent cells
 * overridden; delta_w re-expressed so base_w + delta_w is unchanged.
 * bc_cells/bc_avg describe PartitionOffsets.bc_cells (offset_partitioning.py:95-107). */
int sb_rebase_w(sb_plan* plan, double* d_u, double* d_base_w,
                int32_t band, int64_t n_bc_cells,
                const int32_t* d_bc_cells, const double* d_bc_avg, void* stream);

/* ---- a11: Newton update + convergence metrics -----------------------------------
 * Replaces NewtonSolver.do_update_solution / NewtonSolution.{b_norm, du_norm,
 * rel_du_norm, combined_du_norm, has_nans} (newton_solver.py:79-122,306-317,533-547).
 * u -= omega * clip(du, +-max_du) (max_du <= 0: no clipping).
 * d_out[0]=|b|_inf  [1]=|du|_inf  [2]=sum (du/u)^2 over finite ratios
 * [3]=max |du| / (rel |u| + abs tol_i)  [4]=nan flag (u or b)  [5]=count of finite ratios */
int sb_newton_update(sb_plan* plan, double* d_u, const double* d_du, const double* d_b,
                     const double* d_tol, double omega, double max_du,
                     double rel_tol, double abs_tol, double* d_out, void* stream);
/* the same with NewtonSolverLogDamping's step (newton_solver.py:542-547, ref. Gaury 2018):
 * the applied step is sign(du) log1p(log_alpha |du|) / log_alpha (log_alpha <= 0: off; the
 * reference's default alpha is 1.72); the metrics use the raw du like the reference.   */
int sb_newton_update_damped(sb_plan* plan, double* d_u, const double* d_du, const double* d_b,
                            const double* d_tol, double omega, double max_du, double log_alpha,
                            double rel_tol, double abs_tol, double* d_out, void* stream);

/* ---- terminal current ----------------------------------------------
 * Replaces MetaExtractorIntegrals.add_surface_flux (simudo/io/output_writer.py:
 * 174-231): sum over listed exterior facets of the BDM normal flux of sub-problem
 * p; d_out[0] = integral j.n ds, d_out[1] = total facet length.                   */
int sb_surface_flux(sb_plan* plan, const double* d_u, int32_t p,
                    int64_t n_facets, const int32_t* d_facet_cell,
                    const int8_t* d_facet_local, const double* d_facet_sign,
                    double* d_out, void* stream);

/* ---- point evaluation / line cuts ----------------------------------------
 * Replaces Function_eval_many (simudo/fem/extra_bindings.py:27-75), the evaluator behind
 * LineCutCsvPlot (simudo/io/csv.py:54-63): for every point the first cell containing it
 * (here: the containing cell with the lowest index, found through a uniform bin grid built
 * once per mesh), then the value of a sub-function there.  sub_problem p, what = 0: the DG1
 * scalar (phi, delta_w_k) -> d_values[n]; what = 1: the BDM2 vector (E, j_k) -> d_values[n][2].
 * A point outside the mesh is skipped like in the reference (its value slot is left
 * untouched) and d_cell (optional, [n]) receives -1.
 * h_cell_vertices: the array handed to sb_plan_create; nbx, nby <= 0: chosen automatically;
 * h_dubiner_mono [6][6]: monomial matrix of the orthonormal Dubiner P2 basis
 * (simudo_b200/fem/reference_element.py: _dubiner_p2_monomial_matrix).  The device tables
 * of a locator are released with its plan.                                          */
typedef struct sb_locator sb_locator;
int  sb_locator_create(sb_plan* plan, const double* h_cell_vertices, int32_t nbx, int32_t nby,
                       const double* h_dubiner_mono, sb_locator** out);
void sb_locator_destroy(sb_locator* loc);
int  sb_eval_points(sb_locator* loc, const double* d_u, int32_t sub_problem, int32_t what,
                    int64_t n_points, const double* d_xy, double* d_values, int32_t* d_cell,
                    void* stream);

/* ---- Newton linear solve ------------------------------------------------------------
 * Replaces PETScLUSolver("mumps").solve(A, Qdu, b) (simudo/fem/newton_solver.py:
 * 124-181) for product meshes that are thin in y: banded LU with partial pivoting
 * in the ordering h_perm (new index of each dof, e.g. by x-position of its mesh
 * entity), with row/column equilibration.  cuDSS is not available in this image;
 * sb_band_create returns SB_ERR_UNSUPPORTED when the band storage would exceed
 * max_band_bytes (> 0).  One CTA per system, `batch` systems per call.            */
typedef struct sb_band_solver sb_band_solver;
int  sb_band_create(int64_t n, const int64_t* h_rowptr, const int32_t* h_colidx,
                    const int32_t* h_perm, int32_t batch, int64_t max_band_bytes,
                    sb_band_solver** out);
void sb_band_destroy(sb_band_solver* s);
int  sb_band_info(const sb_band_solver* s, int64_t* n, int32_t* kl, int32_t* ku);
/* iterative-refinement steps (default 2): residual b - A x in double-double, correction
 * through the stored factors.  Needed because a Newton row mixes entries of size 1 and
 * 1e28: the weakly determined minority quasi-Fermi modes are lost in a plain double
 * solve (MUMPS does the analogous work through its MC64 scaling, ICNTL(6)=5).      */
int  sb_band_set_refinement(sb_band_solver* s, int32_t steps);
/* d_vals [nsys][vstride] CSR values, d_b / d_x [nsys][n]; h_info (host, optional):
 * 0 = ok, j+1 = zero pivot in column j (synchronises the stream when given).        */
int  sb_band_solve(sb_band_solver* s, int32_t nsys, const double* d_vals, int64_t vstride,
                   const double* d_b, double* d_x, int32_t* h_info, void* stream);
/* x = A^-1 b with the factors and scalings of the last sb_band_solve on this solver (no
 * refinement): for matrices that do not change between solves, e.g. the CG2 mass matrix of
 * OpticalField.update_input -> dolfin.project (simudo/physics/optical.py:260-274,
 * simudo/fem/assign.py:12-47), which the reference refactors on every call.             */
int  sb_band_solve_factored(sb_band_solver* s, int32_t nsys, const double* d_b, double* d_x,
                            void* stream);

/* r = b - A x with a compensated (double-double) row sum for a device CSR matrix: the
 * residual the refinement steps of the Newton linear solve need (a Newton row mixes entries
 * of size 1 and 1e28).  MUMPS does the analogous work internally (ICNTL(10), scaling
 * ICNTL(6)/(8), simudo/fem/newton_solver.py:140-146).                               */
int sb_csr_residual_dd(int64_t n, const int64_t* d_rowptr, const int32_t* d_colidx,
                       const double* d_vals, const double* d_x, const double* d_b,
                       double* d_r, void* stream);

/* ---- a10: optical sub-problems -------------------------
 * One scalar photon-flux field per (photon energy, direction) on CG2 over the PDD
 * mesh.  Replaces, per field and per optical update (physics/steppers.py:120-136):
 *   OpticalField.update_input  (physics/optical.py:260-274; dolfin.project of the
 *       absorption coefficient onto CG2, fem/assign.py:12-47)   -> sb_optics_alpha_rhs
 *       + a sb_band_solve with the constant CG2 mass matrix;
 *   SystemAssembler on OpticalFieldMSORTE.get_weak_form()
 *       (physics/optical1.py1:27-45, Dirichlet :57-61)          -> sb_optics_assemble
 *       + a sb_band_solve;
 *   OpticalField.update_output (physics/optical.py:248-258)     -> sb_optics_to_cells.
 * CG2 dofs: vertex v -> v, facet f -> n_vertices + f; local order = 3 vertices, then
 * the midpoint of the edge opposite vertex i (the layout of sb_state.d_phi_opt).    */
typedef struct sb_optics sb_optics;
typedef struct sb_optics_problem {
  int64_t n_cells, n_dofs, nnz;
  const int32_t* h_cell_dofs;   /* [n_cells][6]                                      */
  const int32_t* h_pos;         /* [n_cells][36] CSR position of local entry (i, j)  */
  const int64_t* h_rowptr;      /* [n_dofs + 1]                                      */
  const int64_t* h_diag_pos;    /* [n_dofs] CSR position of the diagonal             */
  const void* h_tables;         /* packed SbOpticsTables (simudo_b200/fem/cg2.py)    */
  int64_t n_tables;             /* its size in bytes, checked                        */
} sb_optics_problem;
int  sb_optics_create(sb_plan* plan, const sb_optics_problem* pr, sb_optics** out);
void sb_optics_destroy(sb_optics* o);   /* device buffers are released with the plan */
int64_t sb_optics_tables_size(void);
/* d_rhs[n_dofs] = int alpha_field N_i dx with alpha evaluated from the PDD state
 * (electro_optical_process.py:415-418, 464-480) at the 12 points of the degree-6 rule */
int sb_optics_alpha_rhs(sb_optics* o, const sb_state* st, int32_t field, double* d_rhs,
                        void* stream);
/* d_vals[nnz], d_rhs[n_dofs]: MSORTE matrix for direction (ox, oy) with CG2 absorption
 * d_alpha[n_dofs]; rows of d_bc_dofs become Phi = d_bc_vals (inflow Dirichlet)          */
int sb_optics_assemble(sb_optics* o, const double* d_alpha, double ox, double oy,
                       int64_t n_bc, const int32_t* d_bc_dofs, const double* d_bc_vals,
                       double* d_vals, double* d_rhs, void* stream);
/* d_cells[n_cells][6] = d_nodal[cell_dofs]: one field's slice of sb_state.d_phi_opt    */
int sb_optics_to_cells(sb_optics* o, const double* d_nodal, double* d_cells, void* stream);

/* ---- local-charge-neutrality stage ----------------------
 * One Newton iteration of  <w, rho(phi) / eps> = 0  for the DG1 potential d_phi
 * [n_cells][3] with every band at zero quasi-Fermi level: the initial-guess stage
 * PoissonDriftDiffusion.easy_auto_pre_solve() runs for goal 'local charge neutrality'
 * (physics/poisson_drift_diffusion1.py1:89-125; LocalChargeNeutralitySolver,
 * physics/poisson_drift_diffusion.py:318-331 = NewtonSolverMaxDu,
 * fem/newton_solver.py:533-541).  Residual, the 3 x 3 block Jacobian, its solve, the
 * clipped/damped update and the metrics of sb_newton_update (same d_out layout) in one
 * launch; d_du / d_b (optional, [n_cells][3]) receive the raw Newton step and residual. */
int sb_lcn_iteration(const sb_model* model, int64_t n_cells, const double* d_cell_vertices,
                     const int32_t* d_cell_tag, const double* d_tag_params, double* d_phi,
                     const double* d_tol, double omega, double max_du, double rel_tol,
                     double abs_tol, double* d_du, double* d_b, double* d_out, void* stream);

/* ---- measurement ------------------------------------------------------------------
 * CUDA-event timing (on the launching stream) of the dominant kernel of the last
 * sb_assemble call; sb_plan_last_kernel_ms synchronises on that event.            */
int sb_plan_set_profiling(sb_plan* plan, int on);
int sb_plan_last_kernel_ms(sb_plan* plan, double* ms);
/* the same pass split by stage: ms4 = {k_moments, k_generation, k_rows_fast x P,
 * k_rows_special}                                                                  */
int sb_plan_last_stage_ms(sb_plan* plan, double* ms4);

/* measured DFMA throughput of the current device in TFLOP/s (the FP64 roofline
 * denominator bench.py reports beside the HBM one; synchronises)                        */
int sb_measure_fp64_peak(double* tflops);

/* layout of the packed reference tables expected in sb_problem.h_tables */
int64_t sb_tables_size(void);

#ifdef __cplusplus
}
#endif
#endif /* SIMUDO_B200_H */
