"""Parity cases shared by the CPU-emulated run (tests/test_emulated_kernels.py)
and the real B200 run (tests/test_gpu_parity.py).  Every case drives the
product through its C ABI (AssemblyPlan -> libsimudo_b200) and checks it
against the oracle on the same seeded inputs.

Tolerances: the north star asks for 1e-12 relative on assembled entries;
"relative" is taken per CSR row (max-norm of the oracle's row) because single
entries of a row span > 20 decades.
"""
import numpy as np
import torch

from oracle import oracle as orc_mod
from oracle.oracle import Oracle
from simudo_b200 import synthetic as syn
from simudo_b200.fem import model as M
from simudo_b200.fem.assembly import AssemblyPlan
from simudo_b200.fem.problem_arrays import ProblemArrays
from simudo_b200.mesh import Product2DMesh

from parity_util import row_scaled_error, vector_error, run_plan, entry_relative_error

A_TOL = 1e-12
B_TOL = 1e-12


def _oracle(pa, st):
    return Oracle(pa).assemble(st['u'], st['base_w'], st['thmeq_phi'], st['phi_opt'],
                               st['bc_values'], st['natbc_values'])


def pattern_of(kw):
    return kw.get('pattern', 'structural')


def assembly_parity(maker, nx, ny, numbering='b200', with_bcs=True, **kw):
    pa = maker(nx=nx, ny=ny, numbering=numbering, with_bcs=with_bcs, **kw)
    st = syn.synthetic_state(pa)
    vals0, b0 = _oracle(pa, st)
    plan = AssemblyPlan(pa)
    assert plan.is_native() == (pattern_of(kw) in ('native', 'bsr3') and numbering == 'b200')
    vals, b = run_plan(plan, st)
    vals, b = vals.cpu().numpy(), b.cpu().numpy()
    assert np.isfinite(vals).all() and np.isfinite(b).all()
    ea, eb = row_scaled_error(pa, vals, vals0), vector_error(b, b0)
    assert ea < A_TOL, 'Jacobian row-scaled error %.3e' % ea
    assert eb < B_TOL, 'residual error %.3e' % eb
    # bit-exact sparsity: device-built column indices == host CSR structure
    col = plan.column_indices().cpu().numpy()
    assert (col == pa.csr.column_indices()).all()
    # ... and that structure is the reference's: same (row, column) set as the structural
    # (or dolfin) pattern built independently from the cell dof lists
    rp, cs, perm = pa.csr.sorted_csr()
    assert (np.diff(rp) > 0).all() and len(np.unique(perm)) == pa.csr.nnz
    if pa.pattern == 'dolfin':
        # structural zeros of the dolfin pattern are never touched: exactly zero
        pos = pa.csr.entry_positions()
        from simudo_b200.fem.dofmap import structural_mask
        S = structural_mask(pa.P)
        assert (vals[pos[:, ~S]] == 0.0).all()
    return pa, st, plan, vals, b, vals0, b0


def generation_signature_matches_interpreter():
    """The compiled-in process lists (sb_plan_generation_signature 1: IB cell, 2: pn + SRH)
    against the oracle and against the interpreted list of the same model."""
    import os
    for maker, nx, ny, want in ((syn.ib_problem, 9, 4, 1), (syn.pn_problem, 9, 4, 2)):
        pa = maker(nx=nx, ny=ny, pattern='bsr3')
        st = syn.synthetic_state(pa)
        vals0, b0 = _oracle(pa, st)
        plan = AssemblyPlan(pa)
        assert plan.is_native() and plan.generation_signature() == want
        vals, b = (x.cpu().numpy() for x in run_plan(plan, st))
        os.environ['SIMUDO_B200_GEN_INTERPRET'] = '1'
        try:
            plan_i = AssemblyPlan(pa)
        finally:
            del os.environ['SIMUDO_B200_GEN_INTERPRET']
        assert plan_i.generation_signature() == 0
        vals_i, b_i = (x.cpu().numpy() for x in run_plan(plan_i, st))
        for v, bb in ((vals, b), (vals_i, b_i)):
            assert row_scaled_error(pa, v, vals0) < A_TOL and vector_error(bb, b0) < B_TOL
        assert row_scaled_error(pa, vals, vals_i) < 1e-13
    # a list that is not compiled in (the IB cell without the band-to-band process) is interpreted
    pa = syn.ib_problem(nx=6, ny=3, pattern='bsr3')
    pa.model.n_procs = 2
    plan = AssemblyPlan(pa)
    assert plan.is_native() and plan.generation_signature() == 0
    st = syn.synthetic_state(pa)
    vals0, b0 = _oracle(pa, st)
    vals, b = (x.cpu().numpy() for x in run_plan(plan, st))
    assert row_scaled_error(pa, vals, vals0) < A_TOL and vector_error(b, b0) < B_TOL


def band_window_kernels_match_global_kernels():
    """The shared-memory-window LU / triangular-solve kernels (csrc/band_solver.cu) against the
    global-memory gbtf2 kernels they replace (same pivots -> identical factors) and against
    scipy, on random banded systems: several staging refills, several solve blocks, kl != ku,
    n smaller than the window."""
    import os
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    from simudo_b200.fem.linear_solver import BandSolver
    dev = 'cuda' if torch.cuda.is_available() else 'cpu'
    # solvers of very different bandwidths coexist in one process (PDD system + CG2 optics
    # matrices): the wide one is created first and used after the narrow ones exist
    wide = None
    for n, kl, ku, seed in ((400, 60, 60, 9), (5, 1, 1, 0), (40, 3, 2, 1), (100, 7, 5, 2),
                            (200, 20, 11, 3), (37, 4, 9, 4), (300, 1, 1, 5), (0, 0, 0, 0)):
        if n == 0:                                  # last: solve with the first (wide) solver again
            bs, A, b = wide
            x = bs.solve(torch.as_tensor(A.data.copy(), device=dev),
                         torch.as_tensor(b, device=dev)).cpu().numpy()
            xr = spla.spsolve(A.tocsc(), b)
            assert np.abs(x - xr).max() <= 1e-9 * np.abs(xr).max()
            break
        rng = np.random.default_rng(seed)
        rows, cols, vals = [], [], []
        for i in range(n):
            for j in range(max(0, i - kl), min(n, i + ku + 1)):
                if rng.random() < 0.7 or i == j:
                    rows.append(i); cols.append(j)
                    vals.append(rng.normal() + (0.1 if i == j else 0.0))
        A = sp.csr_matrix((vals, (rows, cols)), shape=(n, n))
        A.sort_indices()
        b = rng.normal(size=n)
        x = {}
        for mode in ('window', 'global'):
            if mode == 'global':
                os.environ['SIMUDO_B200_BAND_GLOBAL'] = '1'
            try:
                bs = BandSolver(None, dev, csr=(n, A.indptr.astype(np.int64),
                                                A.indices.astype(np.int32),
                                                np.arange(n, dtype=np.int32)))
            finally:
                os.environ.pop('SIMUDO_B200_BAND_GLOBAL', None)
            x[mode] = bs.solve(torch.as_tensor(A.data.copy(), device=dev),
                               torch.as_tensor(b, device=dev)).cpu().numpy()
            # a second right-hand side through the stored factors (sb_band_solve_factored)
            b2 = rng.normal(size=n)
            x2 = bs.solve_factored(torch.as_tensor(b2, device=dev)).cpu().numpy()
            xr2 = spla.spsolve(A.tocsc(), b2)
            assert np.abs(x2 - xr2).max() <= 1e-8 * np.abs(xr2).max(), (mode, n, kl, ku)
            if mode == 'window':
                bs_keep = bs
        if wide is None:
            wide = (bs_keep, A, b)
        xr = spla.spsolve(A.tocsc(), b)
        assert np.abs(x['window'] - x['global']).max() <= 1e-13 * np.abs(xr).max(), (n, kl, ku)
        assert np.abs(x['window'] - xr).max() <= 1e-10 * np.abs(xr).max(), (n, kl, ku)


def eval_points_parity():
    """sb_eval_points (Function_eval_many, extra_bindings.py:27-75) against the oracle's
    brute-force restatement: interior points, points on cell edges and vertices, points
    outside the mesh, the mid-line cut of LineCutCsvPlot (io/csv.py:30-41)."""
    pa = syn.ib_problem(nx=11, ny=6, pattern='bsr3')
    st = syn.synthetic_state(pa)
    plan = AssemblyPlan(pa)
    u = plan.tensor(st['u'])
    rng = np.random.default_rng(7)
    (x0, y0), (x1, y1) = pa.cell_vertices.reshape(-1, 2).min(0), pa.cell_vertices.reshape(-1, 2).max(0)
    pts = np.column_stack([rng.uniform(x0, x1, 300), rng.uniform(y0, y1, 300)])
    t = np.linspace(0, 1, 101)
    cut = np.column_stack([x0 + t * (x1 - x0), np.full_like(t, 0.5 * (y0 + y1))])
    verts = pa.cell_vertices.reshape(-1, 2)[::7]
    outside = np.array([[x0 - 1.0, y0], [x1 + 1e-3, y1], [0.5 * (x0 + x1), y1 + 2.0]])
    allp = np.concatenate([pts, cut, verts, outside])
    for p in range(pa.P):
        for vector in (False, True):
            ref, cref = orc_mod.eval_points(pa, st['u'], p, allp, vector=vector)
            got, cg = plan.eval_points(u, p, allp, vector=vector, return_cells=True)
            got, cg = got.cpu().numpy(), cg.cpu().numpy()
            assert (cg == cref).all()
            assert (cg[-3:] == -1).all() and np.isnan(got[-3:]).all()
            inside = cref >= 0
            scale = np.abs(ref[inside]).max()
            assert np.abs(got[inside] - ref[inside]).max() <= 1e-12 * scale


def assembly_parity_at_scale(nx, ny):
    """The benchmarked configuration (bench.py: IB cell, nx x ny points, BCs on) against the
    oracle: what small cases cannot see -- 64-bit value offsets past 2^31, full grids, the
    closed forms at 1e9 non-zeros.  CSR 'native' layout against the oracle (all host
    threads, ~15 s per 1M cells); the block layout 'bsr3' bit for bit against 'native'
    entry by entry.  Returns (cells, row-relative, entry-relative, residual) errors."""
    import os
    pa = syn.ib_problem(nx=nx, ny=ny, pattern='native')
    st = syn.synthetic_state(pa)
    plan = AssemblyPlan(pa)
    assert plan.is_native()
    vals, b = run_plan(plan, st)
    col = plan.column_indices().cpu().numpy()
    # structure: the device-built columns against the host pattern on a sample of cells
    # (entry-wise on the whole mesh at the small sizes of assembly_parity)
    rng = np.random.default_rng(0)
    cells = np.unique(np.concatenate([np.arange(64), pa.n_cells - 1 - np.arange(64),
                                      rng.integers(0, pa.n_cells, 4096)]))
    pos = pa.csr.entry_positions(cells)
    cols = np.broadcast_to(pa.cell_dofs[cells][:, None, :], pos.shape)
    ok = pos >= 0
    assert (col[pos[ok]] == cols[ok]).all()
    nthr = orc_mod.lib().pdd_oracle_max_threads()
    orc = Oracle(pa, n_threads=nthr, colidx=col)
    vals0, b0 = orc.assemble(st['u'], st['base_w'], st['thmeq_phi'], st['phi_opt'],
                             st['bc_values'], st['natbc_values'])
    v = vals.cpu().numpy()
    assert np.isfinite(v).all()
    ea = row_scaled_error(pa, v, vals0)
    ee = entry_relative_error(v, vals0)
    eb = vector_error(b.cpu().numpy(), b0)
    assert ea < A_TOL, 'Jacobian row-scaled error %.3e' % ea
    assert eb < B_TOL, 'residual error %.3e' % eb
    assert ee < 1e-9, 'worst entry-relative error %.3e' % ee
    del v, vals0, orc
    # block layout: same arithmetic, other placement -> identical bits, entry by entry
    pa2 = syn.ib_problem(nx=nx, ny=ny, pattern='bsr3')
    plan2 = AssemblyPlan(pa2)
    assert plan2.is_native()
    vals2, b2 = run_plan(plan2, st)
    assert torch.equal(b2, b)
    T = plan.tensor
    step = 1 << 14
    for c0 in range(0, pa.n_cells, step):
        cs = np.arange(c0, min(pa.n_cells, c0 + step))
        p1 = pa.csr.entry_positions(cs).ravel()
        p2 = pa2.csr.entry_positions(cs).ravel()
        ok = p1 >= 0
        assert ((p2 >= 0) == ok).all()
        assert torch.equal(vals[T(p1[ok], torch.int64)], vals2[T(p2[ok], torch.int64)])
    return pa.n_cells, ea, ee, eb


def thermal_equilibrium_mode():
    """goal 'thermal equilibrium': bands feed rho with qfl == 0, no band dofs
    (poisson_drift_diffusion.py:221-223)."""
    pa_full = syn.pn_problem(nx=7, ny=3, with_bcs=False)
    model = M.make_model([dict(kind=0, sign=-1), dict(kind=0, sign=+1)], [],
                         n_tags=2, bands_have_dofs=False)
    pa = ProblemArrays(pa_full.mesh, pa_full.cell_tag, model, pa_full.tag_params)
    bf = pa.mesh.boundary_facets()
    pa.set_bc_dofs(pa.facet_dofs(0, bf[::2]).ravel())
    st = syn.synthetic_state(pa)
    st['base_w'] = None
    st['bc_values'] = 0.01 * np.cos(np.arange(len(pa.bc_dofs)))
    st['natbc_values'] = np.zeros(0)
    vals0, b0 = Oracle(pa).assemble(st['u'], None, None, None, st['bc_values'])
    plan = AssemblyPlan(pa)
    vals, b = run_plan(plan, st)
    assert row_scaled_error(pa, vals.cpu().numpy(), vals0) < A_TOL
    assert vector_error(b.cpu().numpy(), b0) < B_TOL


def fill_dependent_ib_mobility():
    pa, st, plan, vals, b, vals0, b0 = assembly_parity(
        syn.ib_problem, 6, 3, const_mobility=False)


def non_default_quadrature_degrees():
    """mixedqfl_debug_quad_degree_* are user-settable
    (example/pn_benchmark/pn_benchmark.py:209-222)."""
    pa0 = syn.pn_problem(nx=6, ny=3)
    model = M.make_model([dict(kind=0, sign=-1), dict(kind=0, sign=+1)],
                         [dict(kind=M.PROC_SRH, dst=0, src=1)], n_tags=2,
                         quad_degree_rho=12, quad_degree_super=16, quad_degree_g=10)
    pa = ProblemArrays(pa0.mesh, pa0.cell_tag, model, pa0.tag_params,
                       jump_facet_mask=np.full(pa0.mesh.num_facets, 3, np.uint8))
    st = syn.synthetic_state(pa)
    st['bc_values'] = np.zeros(0)
    st['natbc_values'] = np.zeros(0)
    vals0, b0 = _oracle(pa, st)
    vals, b = run_plan(AssemblyPlan(pa), st)
    assert row_scaled_error(pa, vals.cpu().numpy(), vals0) < A_TOL
    assert vector_error(b.cpu().numpy(), b0) < B_TOL


def assembly_is_deterministic_and_idempotent():
    pa = syn.ib_problem(nx=8, ny=4)
    st = syn.synthetic_state(pa)
    plan = AssemblyPlan(pa)
    v1, b1 = run_plan(plan, st)
    T = plan.tensor
    args = (T(st['u']), T(st['base_w']), T(st['thmeq_phi']), T(st['phi_opt']),
            T(st['bc_values']), T(plan.natbc_device_order(st['natbc_values'])))
    v2, b2 = v1.clone(), b1.clone()
    plan.assemble(*args, v2, b2)          # re-assemble into the used buffers
    assert torch.equal(v1, v2) and torch.equal(b1, b2)
    v3, b3 = run_plan(AssemblyPlan(pa), st)
    assert torch.equal(v1, v3) and torch.equal(b1, b3)


def residual_only_and_matrix_only():
    pa = syn.pn_problem(nx=6, ny=3)
    st = syn.synthetic_state(pa)
    plan = AssemblyPlan(pa)
    vals, b = run_plan(plan, st)
    T = plan.tensor
    args = (T(st['u']), T(st['base_w']), T(st['thmeq_phi']), None,
            T(st['bc_values']), T(plan.natbc_device_order(st['natbc_values'])))
    b_only = plan.new_vector()
    plan.assemble(*args, None, b_only)
    v_only = plan.new_values()
    plan.assemble(*args, v_only, None)
    assert torch.equal(b, b_only) and torch.equal(vals, v_only)


def rebase_w_parity():
    pa = syn.ib_problem(nx=7, ny=3)
    st = syn.synthetic_state(pa)
    plan = AssemblyPlan(pa)
    T = plan.tensor
    u, bw = T(st['u']).clone(), T(st['base_w']).clone()
    u_ref, bw_ref = st['u'].copy(), st['base_w'].copy()
    bc_cells = np.array([0, 3, 7], dtype=np.int32)
    bc_avg = np.array([0.11, -0.2, 0.05])
    lay = pa.layout()
    tot_before = u_ref[pa.cell_dofs[:, lay['dg'][2]]] + bw_ref[1][:, None]
    for band in range(3):
        u_ref, bw_ref = orc_mod.rebase_w(pa, u_ref, bw_ref, band, bc_cells, bc_avg)
        plan.rebase_w(u, bw, band, T(bc_cells, torch.int32), T(bc_avg))
    assert np.abs(u.cpu().numpy() - u_ref).max() < 1e-15
    assert np.abs(bw.cpu().numpy() - bw_ref).max() < 1e-15
    # size-independent property: base_w + delta_w is invariant under rebasing
    tot_after = u.cpu().numpy()[pa.cell_dofs[:, lay['dg'][2]]] + bw.cpu().numpy()[1][:, None]
    assert np.abs(tot_after - tot_before).max() < 1e-14
    assert np.abs(bw.cpu().numpy()[:, bc_cells] - bc_avg[None, :]).max() == 0.0


def newton_update_parity():
    pa = syn.pn_problem(nx=9, ny=4)
    plan = AssemblyPlan(pa)
    rng = np.random.default_rng(5)
    u = rng.standard_normal(pa.N)
    u[::17] = 0.0                                  # |du/u| = inf / nan cases
    du = rng.standard_normal(pa.N) * 1e-3
    du[::34] = 0.0
    b = rng.standard_normal(pa.N)
    tol = pa.tolerance_vector([1e-6] * pa.P, [1e-4] + [1e-10] * (pa.P - 1))
    expect = u - 0.5 * np.clip(du, -2e-3, 2e-3)
    ref = orc_mod.newton_metrics(expect, du, b, tol, 1e-4, 1.0)
    T = plan.tensor
    ud = T(u).clone()
    out = plan.newton_update(ud, T(du), T(b), T(tol), 0.5, 2e-3, 1e-4, 1.0).cpu().numpy()
    assert out[0] == ref['b_norm'] and out[1] == ref['du_norm']
    assert np.sqrt(out[2]) == np.float64(ref['rel_du_norm']) or \
        abs(np.sqrt(out[2]) - ref['rel_du_norm']) < 1e-10 * ref['rel_du_norm'] or \
        not np.isfinite(ref['rel_du_norm'])
    assert abs(out[3] - ref['combined_du_norm']) <= 1e-15 * ref['combined_du_norm']
    assert out[4] == 0.0
    assert np.abs(ud.cpu().numpy() - expect).max() == 0.0
    # NewtonSolverLogDamping (newton_solver.py:542-547): damped step, metrics from the raw du
    alpha = 1.72
    ud = T(u).clone()
    out = plan.newton_update(ud, T(du), T(b), T(tol), 0.7, None, 1e-4, 1.0,
                             log_alpha=alpha).cpu().numpy()
    expect = u - 0.7 * np.sign(du) * np.log1p(np.abs(du) * alpha) / alpha
    assert np.abs(ud.cpu().numpy() - expect).max() <= 4e-16 * np.abs(expect).max()
    assert out[1] == np.abs(du).max()
    # NaN poisoning is detected, not raised (newton_solver.py:84-89)
    u2 = u.copy(); u2[5] = np.nan
    out = plan.newton_update(T(u2), T(du), T(b), T(tol), 1.0, None, 1e-4, 1.0).cpu().numpy()
    assert out[4] > 0.0


def natbc_eval_parity():
    """a8: the exterior-facet kernel (sb_natbc_eval) against the oracle's physical-space
    restatement of the `ds` integrals, on both contacts of a graded 2D mesh (both signs of
    detJ, all three local facet numbers), constant + P2 boundary function."""
    pa = syn.ib_problem(nx=7, ny=6, pattern='native')
    plan = AssemblyPlan(pa)
    T = plan.tensor
    mesh = pa.mesh
    facets = np.concatenate([pa.contacts['left'], pa.contacts['right'], pa.contacts['horiz']])
    # the boundary of a product mesh only has local facets 0 and 2: add a few interior
    # facets seen from the cell in which they are local facet 1 (the kernel integrates over
    # whatever facet of whatever cell it is given, with that cell's outward normal)
    inner = np.nonzero((mesh.facet_cells[:, 1] >= 0) & (mesh.facet_local[:, 0] == 1))[0][:7]
    facets = np.concatenate([facets, inner])
    cells = mesh.facet_cells[facets, 0].astype(np.int32)
    local = mesh.facet_local[facets, 0].astype(np.int8)
    assert set(local.tolist()) == {0, 1, 2}
    rng = np.random.default_rng(11)
    c0 = rng.standard_normal(len(facets))
    p2 = rng.standard_normal((len(facets), 6))
    slot = rng.permutation(3 * len(facets)).astype(np.int32)
    out = torch.zeros(3 * len(facets), dtype=torch.float64, device=plan.device)
    plan.natbc_eval(T(cells, torch.int32), T(local, torch.int8), T(slot, torch.int32), T(c0),
                    T(p2), out)
    ref = orc_mod.natural_bc_integrals(pa, cells, local, c0, p2).ravel()
    got = out.cpu().numpy()[slot]
    assert np.abs(got - ref).max() <= 1e-13 * np.abs(ref).max()
    # constant-only and function-only variants (NULL pointers)
    plan.natbc_eval(T(cells, torch.int32), T(local, torch.int8), T(slot, torch.int32), T(c0),
                    None, out)
    ref = orc_mod.natural_bc_integrals(pa, cells, local, c0, np.zeros_like(p2)).ravel()
    assert np.abs(out.cpu().numpy()[slot] - ref).max() <= 1e-13 * np.abs(ref).max()


def surface_flux_parity():
    pa = syn.pn_problem(nx=6, ny=5)
    st = syn.synthetic_state(pa)
    plan = AssemblyPlan(pa)
    T = plan.tensor
    mesh = pa.mesh
    for name in ('left', 'right'):
        facets = pa.contacts[name]
        for p in range(pa.P):
            ref_flux, ref_len = orc_mod.surface_flux(pa, st['u'], p, facets)
            out = plan.surface_flux(
                T(st['u']), p, T(mesh.facet_cells[facets, 0], torch.int32),
                T(mesh.facet_local[facets, 0], torch.int8),
                T(np.ones(len(facets)))).cpu().numpy()
            assert abs(out[0] - ref_flux) <= 1e-13 * max(abs(ref_flux), 1e-300)
            assert abs(out[1] - ref_len) <= 1e-14 * ref_len


def jacobian_consistent_with_residual_large(maker, nx, ny):
    """Size-independent property for meshes the oracle cannot afford:
    J v == d/dh F(u + h v) (central differences) through the product alone."""
    pa = maker(nx=nx, ny=ny, with_bcs=False)
    st = syn.synthetic_state(pa)
    plan = AssemblyPlan(pa)
    T = plan.tensor
    lay = pa.layout()
    scale = np.ones(pa.N)
    for p in range(pa.P):
        scale[pa.cell_dofs[:, lay['dg'][p]].ravel()] = 1e-3
        scale[np.unique(pa.cell_dofs[:, lay['bdm'][p]])] = 1e-3 if p == 0 else 1e-12
    v = np.random.default_rng(2).standard_normal(pa.N) * scale
    fixed = (T(st['base_w']), T(st['thmeq_phi']), T(st['phi_opt']), None, None)
    vals, b0 = plan.new_values(), plan.new_vector()
    plan.assemble(T(st['u']), *fixed, vals, b0)
    h = 1e-4
    bp, bm = plan.new_vector(), plan.new_vector()
    plan.assemble(T(st['u'] + h * v), *fixed, None, bp)
    plan.assemble(T(st['u'] - h * v), *fixed, None, bm)
    rp, col, perm = pa.csr.sorted_csr()
    A = torch.sparse_csr_tensor(T(rp, torch.int64), T(col, torch.int64), vals[T(perm, torch.int64)],
                                size=(pa.N, pa.N))
    an = (A @ T(v).unsqueeze(1)).squeeze(1)
    fd = (bp - bm) / (2 * h)
    err = (fd - an).abs().max().item() / an.abs().max().item()
    assert err < 1e-6, err
    return pa.n_cells


def unstaged_rank_tables_match_staged():
    """Meshes with more rank patterns than fit in shared memory take the
    global-memory rank path of k_rows_fast; forced here with the library's
    debugging knob and compared bit for bit."""
    import os
    pa = syn.ib_problem(nx=6, ny=3)
    st = syn.synthetic_state(pa)
    v0, b0 = run_plan(AssemblyPlan(pa), st)
    os.environ['SIMUDO_B200_NO_SMEM_PATTERNS'] = '1'
    try:
        v1, b1 = run_plan(AssemblyPlan(pa), st)
    finally:
        del os.environ['SIMUDO_B200_NO_SMEM_PATTERNS']
    assert torch.equal(v0, v1) and torch.equal(b0, b1)


def optics_parity(nx=9, ny=3):
    """a10: alpha projection, MSORTE matrix / right-hand side and the solved photon
    flux of every optical field against oracle/optics_oracle.py."""
    import scipy.sparse as sps
    from oracle import optics_oracle as oo
    from simudo_b200.fem.optics_solver import OpticsPlan
    pa = syn.ib_problem(nx=nx, ny=ny)
    st = syn.synthetic_state(pa)
    plan = AssemblyPlan(pa)
    op = OpticsPlan(plan)
    sp = op.space
    u, bw = plan.tensor(st['u']), plan.tensor(st['base_w'])
    mesh = pa.mesh
    bf = mesh.boundary_facets()
    nrm = mesh.boundary_outward_normals(bf)
    for field, omega in ((0, (1.0, 0.0)), (1, (-1.0, 0.0)), (2, (0.6, 0.8))):
        inflow = bf[nrm @ np.asarray(omega) < -1e-12]          # Omega . n < 0
        bc_dofs = sp.facet_dofs(inflow)
        bc_vals = 1e5 * (1.0 + 0.1 * np.arange(len(bc_dofs)) / len(bc_dofs))
        # oracle
        rhs0 = oo.alpha_rhs(pa, sp, st['u'], st['base_w'], field)
        alpha0, phi0 = oo.solve_field(pa, sp, st['u'], st['base_w'], field, omega,
                                      bc_dofs, bc_vals)
        A0, b0 = oo.msorte_system(sp, alpha0, omega, bc_dofs, bc_vals)
        # device
        rhs = op.alpha_rhs(field, u, bw).cpu().numpy()
        assert np.abs(rhs - rhs0).max() <= 1e-13 * np.abs(rhs0).max()
        alpha = op.project_alpha(field, u, bw)
        assert np.abs(alpha.cpu().numpy() - alpha0).max() <= 1e-11 * np.abs(alpha0).max()
        assert alpha0.max() > 0
        vals, b = op.assemble(plan.tensor(alpha0), omega, plan.tensor(bc_dofs, torch.int32),
                              plan.tensor(bc_vals))
        A = sps.csr_matrix((vals.cpu().numpy(), sp.colidx, sp.rowptr),
                           shape=(sp.n_dofs, sp.n_dofs))
        dA = abs(A - A0)
        scale = abs(A0).max(axis=1).toarray().ravel()
        assert (dA.max(axis=1).toarray().ravel() <= 1e-12 * scale).all()
        assert np.array_equal(b.cpu().numpy(), b0)
        phi, cells = op.solve_field(field, u, bw, omega, plan.tensor(bc_dofs, torch.int32),
                                    plan.tensor(bc_vals))
        phi = phi.cpu().numpy()
        err = np.abs(phi - phi0).max() / np.abs(phi0).max()
        assert err <= 1e-9, err
        assert np.array_equal(cells.cpu().numpy(), phi[sp.cell_dofs])
        # bit-reproducible (fixed-order segmented sums, no atomics)
        vals2, _ = op.assemble(plan.tensor(alpha0), omega, plan.tensor(bc_dofs, torch.int32),
                               plan.tensor(bc_vals))
        assert torch.equal(vals2, vals)


def lcn_iteration_parity():
    """(f-2) one local-charge-neutrality Newton iteration (sb_lcn_iteration) against the
    numpy restatement, on the four-layer IB device's first stage."""
    from oracle_backend import OracleBackend, OracleLCNBackend
    from simudo_b200.fem.backend import CudaLCNBackend
    from simudo_b200.example import fourlayer
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), 'golden'))
    from make_golden import FOURLAYER_SMALL

    class Stop(Exception):
        pass

    captured = {}

    class Capture(OracleBackend):
        pass

    def lcn_factory(system):
        captured['system'] = system
        raise Stop()
    Capture.lcn_backend = staticmethod(lcn_factory)
    try:
        fourlayer.run(P=FOURLAYER_SMALL, backend_factory=Capture)
    except Stop:
        pass
    sysm = captured['system']
    rng = np.random.RandomState(3)
    phi0 = 0.3 + 0.2 * rng.rand(sysm.mesh.num_cells, 3)
    ref, dev = OracleLCNBackend(sysm), CudaLCNBackend(sysm)
    ref.set_phi(phi0)
    dev.set_phi(phi0)
    tol = sysm.tolerance_vector()
    for it in range(3):
        m0 = ref.iterate(tol, 0.2, 0.02, 1e-10, 0.1)
        m1 = dev.iterate(tol, 0.2, 0.02, 1e-10, 0.1)
        assert np.abs(dev.get_phi() - ref.get_phi()).max() < 1e-13
        scale = np.abs(ref.get_b()).max()
        assert np.abs(dev.get_b() - ref.get_b()).max() <= 1e-12 * scale
        assert np.abs(dev.get_du() / ref.get_du() - 1).max() < 1e-9
        for k in ('b_norm', 'du_norm', 'rel_du_norm', 'combined_du_norm'):
            assert abs(m1[k] - m0[k]) <= 1e-9 * abs(m0[k]), (k, m0[k], m1[k])
        assert m1['has_nans'] == m0['has_nans']


def bulk_store_matches_entry_flush():
    """Owned rows leave the SM through cp.async.bulk (1-D TMA) copies of rank-ordered
    staging; the library's debugging knob forces the lane = entry flush instead.  Both
    must give the same bits (pn: odd run lengths, ib: even)."""
    import os
    for maker, nx, ny in ((syn.ib_problem, 9, 4), (syn.pn_problem, 8, 3)):
        pa = maker(nx=nx, ny=ny)
        st = syn.synthetic_state(pa)
        v0, b0 = run_plan(AssemblyPlan(pa), st)
        os.environ['SIMUDO_B200_NO_BULK_STORE'] = '1'
        try:
            v1, b1 = run_plan(AssemblyPlan(pa), st)
        finally:
            del os.environ['SIMUDO_B200_NO_BULK_STORE']
        assert torch.equal(v0, v1) and torch.equal(b0, b1)
