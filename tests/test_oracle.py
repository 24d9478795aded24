"""The oracle checked against itself and against closed forms (CPU)."""
import numpy as np
import pytest

from oracle import fiat_like as fl
from oracle.oracle import Oracle
from simudo_b200 import synthetic as syn
from simudo_b200.fem import reference_element as re


def test_gauss_jacobi_exactness():
    for m in (5, 11):
        x, w = fl.gauss_jacobi(m, 1.0, 0.0)
        for k in range(2 * m):
            # int_-1^1 (1-x) x^k dx
            exact = (1 - (-1) ** (k + 1)) / (k + 1) - (1 - (-1) ** (k + 2)) / (k + 2)
            assert np.dot(w, x ** k) == pytest.approx(exact, abs=1e-13)


def test_collapsed_rule_integrates_degree_20():
    pts, w = fl.collapsed_triangle_rule(11)
    assert len(w) == 121 and w.sum() == pytest.approx(0.5, abs=1e-14)
    from math import factorial
    for a in range(0, 21, 5):
        for b in range(0, 21 - a, 4):
            exact = factorial(a) * factorial(b) / factorial(a + b + 2)
            assert np.dot(w, pts[:, 0] ** a * pts[:, 1] ** b) == pytest.approx(exact, rel=1e-12)


def test_two_independent_element_derivations_agree():
    """sympy-exact FIAT-style tables (oracle) vs Dubiner-based tables (product)."""
    for m in (5, 11):
        p1, w1 = fl.collapsed_triangle_rule(m)
        p2, w2 = re.collapsed_rule(m)
        assert np.abs(p1 - p2).max() < 1e-15 and np.abs(w1 - w2).max() < 1e-16
    pts, _ = fl.collapsed_triangle_rule(11)
    v1, d1 = fl.tabulate_bdm2(pts)
    v2, d2 = re.bdm2_eval(pts)
    assert np.abs(v1 - v2.transpose(1, 0, 2)).max() < 1e-12
    assert np.abs(d1 - d2.T).max() < 1e-12
    assert np.abs(fl.tabulate_p2(pts) - re.p2_lagrange_eval(pts).T).max() < 1e-15


def test_bdm2_dual_basis_definition():
    """Nodal property against FIAT's dual set (Appendix A3)."""
    import sympy as sp
    (X, Y), basis = fl.bdm2_symbolic()
    # first basis function in closed form (exact rationals)
    assert sp.expand(basis[0][0] - (7 * X ** 2 + 2 * X * Y - 4 * X)) == 0
    assert sp.expand(basis[0][1] - (-2 * X * Y + Y ** 2)) == 0
    _, _, trace, _, _ = fl.facet_tabulation(3)
    # normal traces of interior functions vanish on every edge
    assert np.abs(trace[:, :, 9:]).max() < 1e-14
    # facet functions have zero normal trace on the other edges
    for f in range(3):
        others = [i for i in range(9) if i // 3 != f]
        assert np.abs(trace[f][:, others]).max() < 1e-14


@pytest.mark.parametrize('maker', [syn.pn_problem, syn.ib_problem])
def test_oracle_jacobian_matches_finite_differences(maker):
    pa = maker(nx=5, ny=3, with_bcs=False)
    st = syn.synthetic_state(pa)
    orc = Oracle(pa)
    args = (st['base_w'], st['thmeq_phi'], st['phi_opt'])
    vals, _ = orc.assemble(st['u'], *args)
    A = orc.to_scipy(vals)
    lay = pa.layout()
    scale = np.ones(pa.N)
    for p in range(pa.P):
        scale[pa.cell_dofs[:, lay['dg'][p]].ravel()] = 1e-3
        scale[np.unique(pa.cell_dofs[:, lay['bdm'][p]])] = 1e-3 if p == 0 else 1e-12
    rng = np.random.default_rng(1)
    v = rng.standard_normal(pa.N) * scale
    h = 1e-4
    _, bp = orc.assemble(st['u'] + h * v, *args, want_A=False)
    _, bm = orc.assemble(st['u'] - h * v, *args, want_A=False)
    fd = (bp - bm) / (2 * h)
    an = A @ v
    assert np.abs(fd - an).max() <= 1e-6 * np.abs(an).max()


def test_oracle_thermal_equilibrium_has_zero_generation_and_current_residual():
    """g == 0 and the drift-diffusion rows vanish when qfl == 0, j == 0 and
    phi equals the thermal-equilibrium potential."""
    pa = syn.pn_problem(nx=6, ny=3, with_bcs=False)
    st = syn.synthetic_state(pa)
    lay = pa.layout()
    u = st['u'].copy()
    for p in (1, 2):
        u[pa.cell_dofs[:, lay['dg'][p]].ravel()] = 0.0
        u[np.unique(pa.cell_dofs[:, lay['bdm'][p]])] = 0.0
    # thermal-equilibrium potential = the (P1) potential itself, as P2 nodal values
    phi_v = u[pa.cell_dofs[:, lay['dg'][0]]]
    mid = 0.5 * (phi_v[:, [1, 0, 0]] + phi_v[:, [2, 2, 1]])
    thmeq = np.concatenate([phi_v, mid], axis=1)
    orc = Oracle(pa)
    _, b = orc.assemble(u, np.zeros_like(st['base_w']), thmeq, None, want_A=False)
    for p in (1, 2):
        rows = np.concatenate([pa.cell_dofs[:, lay['dg'][p]].ravel(),
                               np.unique(pa.cell_dofs[:, lay['bdm'][p]])])
        # generation term exactly zero (n p = n0 p0), no current
        assert np.abs(b[rows]).max() < 1e-9 * np.abs(b).max()


def test_oracle_manufactured_mixed_poisson_converges():
    """Mixed Poisson alone (no carriers): -div(eps grad phi) = rho with
    phi = 0 on the contacts has the exact parabola phi = rho x (L - x) / (2 eps)."""
    import scipy.sparse.linalg as spla
    from simudo_b200.fem import model as M
    from simudo_b200.fem.problem_arrays import ProblemArrays
    from simudo_b200.mesh import Product2DMesh
    errs = []
    for nx in (5, 9):
        mesh = Product2DMesh(np.linspace(0, 1, nx), np.linspace(0, 0.5, 3)).mesh
        model = M.make_model([], [], n_tags=1, bands_have_dofs=False)
        tp = M.new_tag_params(1)
        tp[0, M.TP_KT] = 0.025
        tp[0, M.TP_EPS] = 2.0
        tp[0, M.TP_RHO_STATIC] = 3.0
        pa = ProblemArrays(mesh, np.zeros(mesh.num_cells, np.int32), model, tp)
        bf = mesh.boundary_facets()
        n = mesh.boundary_outward_normals(bf)
        horiz = bf[np.abs(n[:, 1]) > 0.5]
        pa.set_bc_dofs(pa.facet_dofs(0, horiz).ravel())      # E.n = 0 top/bottom
        orc = Oracle(pa)
        u = np.zeros(pa.N)
        vals, b = orc.assemble(u, bc_values=np.zeros(len(pa.bc_dofs)))
        du = spla.spsolve(orc.to_scipy(vals).tocsc(), b)
        u -= du                                                # linear: one Newton step
        lay = pa.layout()
        phi = u[pa.cell_dofs[:, lay['dg'][0]]]
        x = pa.cell_vertices[:, :, 0]
        exact = 3.0 * x * (1 - x) / (2 * 2.0)
        errs.append(np.abs(phi - exact).max())
        _, b2 = orc.assemble(u, bc_values=np.zeros(len(pa.bc_dofs)), want_A=False)
        assert np.abs(b2).max() < 1e-10
    assert errs[1] < errs[0] / 3.0          # at least second-order in h


def test_multifrontal_solver_matches_sparse_lu_on_2d_meshes():
    """The nested-dissection multifrontal LU (simudo_b200/fem/multifrontal.py; numeric phase
    in batched torch ops, here on host tensors) on genuinely two-dimensional Newton matrices
    against scipy's SuperLU: residual of the equilibrated system and solution."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    import torch
    from simudo_b200.fem.multifrontal import MultifrontalSolver
    for mk, nx, ny, leaf in ((syn.pn_problem, 12, 9, 8), (syn.ib_problem, 14, 7, 16)):
        pa = mk(nx=nx, ny=ny, pattern='native')
        st = syn.synthetic_state(pa)
        orc = Oracle(pa, n_threads=4)
        vals, b = orc.assemble(st['u'], st['base_w'], st['thmeq_phi'], st['phi_opt'],
                               st['bc_values'], st['natbc_values'])
        mf = MultifrontalSolver(pa, 'cpu', leaf_cells=leaf)
        x = mf.solve(torch.as_tensor(vals), torch.as_tensor(b)).numpy()
        A = orc.to_scipy(vals).tocsc()
        r = 1.0 / np.abs(A).max(axis=1).toarray().ravel()
        As = sp.diags(r) @ A
        c = 1.0 / np.abs(As).max(axis=0).toarray().ravel()
        x0 = c * spla.splu((As @ sp.diags(c)).tocsc()).solve(r * b)
        res = np.abs(r * (A @ x - b)).max() / np.abs(r * b).max()
        assert res < 1e-11, res
        rs = mf.rs.numpy()                       # backward error in the solver's own scaling
        assert np.abs(rs * (A @ x - b)).max() / np.abs(rs * b).max() < 1e-14
        assert np.abs(x - x0).max() <= 1e-4 * np.abs(x0).max()
        # second solve with the same symbolic phase and another matrix / right-hand side
        vals2 = vals * (1.0 + 0.01 * np.cos(np.arange(len(vals))))
        x2 = mf.solve(torch.as_tensor(vals2), torch.as_tensor(b)).numpy()
        A2 = orc.to_scipy(vals2).tocsr()
        assert np.abs(r * (A2 @ x2 - b)).max() / np.abs(r * b).max() < 1e-11


def test_oracle_manufactured_qfl_with_dg0_offsets_and_jump_term():
    """The reference's own manufactured solution for exactly this discretisation
    (simudo/trash/historical/qflop_test.py:11-68): mixed quasi-Fermi-level form in DG1 x BDM2
    with a DG0 offset ``w0 = 1.0 if x > 0.5 else 0.3`` and the interior-facet jump term

        v div c - v f + exp(w + w0) phi.c + div(phi) w - (w_BC - w0) phi.n ds
                                                - (w0(-) - w0(+)) phi(+).n(+) dS = 0,

    exact solution ``W = w + w0 = x^2 (1 - x)^2 / 10``, ``c = grad W exp(-W)``, ``f = div c``.
    In the oracle this is one Boltzmann band with s = +1, kT = 1 eV, E = 0, mu N = dd_scale
    (so 1/(mu u) = exp(chi)), an electrically dead Poisson block (eps -> inf, phi_bc = 0,
    hence phi = 0) and no processes; the prescribed source f enters as a constant vector on
    the band's DG1 rows (it does not depend on the unknowns).  Checks: Newton converges, the
    smooth total W is recovered through the discontinuous offset (the jump term does its job),
    O(h^2) convergence of W in DG1 and of the flux."""
    import scipy.sparse.linalg as spla
    import sympy as sp
    from oracle import fiat_like as fl
    from oracle.oracle import natural_bc_integrals
    from simudo_b200.fem import model as M
    from simudo_b200.fem.problem_arrays import ProblemArrays
    from simudo_b200.mesh import Product2DMesh
    X, Y = sp.symbols('x y')
    Wx = X ** 2 * (1 - X) ** 2 / 10
    cx = sp.diff(Wx, X) * sp.exp(-Wx)
    f_ex = sp.lambdify((X, Y), sp.diff(cx, X), 'numpy')
    W_ex = sp.lambdify((X, Y), Wx, 'numpy')
    c_ex = sp.lambdify((X, Y), cx, 'numpy')
    qp, qw = fl.default_triangle_scheme(8)                 # reference points / weights (5 x 5)
    p1q = fl.tabulate_p1(qp)                               # [q, 3]
    bdq, _ = fl.tabulate_bdm2(qp)                          # [q, 12, 2]
    errs_w, errs_c = [], []
    for nx in (8, 16, 32):
        mesh = Product2DMesh(np.linspace(0, 1, nx + 1), np.linspace(0, 1, 4)).mesh
        model = M.make_model([dict(kind=M.BAND_NONDEGENERATE, sign=+1)], [], n_tags=1)
        tp = M.new_tag_params(1)
        tp[0, M.TP_KT] = 1.0
        tp[0, M.TP_EPS] = 1e30
        o = M.TP_BAND0
        tp[0, o + M.TPB_IN_SUBDOMAIN] = 1.0
        tp[0, o + M.TPB_N] = 1.0
        tp[0, o + M.TPB_E] = 0.0
        tp[0, o + M.TPB_MOBILITY] = model.dd_scale        # cinv = dd_scale / (mob N) = 1
        mesh.init()
        jmask = np.full(mesh.num_facets, 1, dtype=np.uint8)
        pa = ProblemArrays(mesh, np.zeros(mesh.num_cells, np.int32), model, tp,
                           jump_facet_mask=jmask)
        xv = pa.cell_vertices                               # [nc, 3, 2]
        xc = xv.mean(axis=1)
        base_w = np.where(xc[:, 0] > 0.5, 1.0, 0.3)[None, :]
        lay = pa.layout()
        # natural BC of the band on the whole boundary: - oint (W_ex - w0) xi.n ds
        bf = mesh.boundary_facets()
        bc = mesh.facet_cells[bf, 0]
        bl = mesh.facet_local[bf, 0].astype(np.int64)
        nodes = np.concatenate([xv, 0.5 * (xv[:, [1, 0, 0]] + xv[:, [2, 2, 1]])], axis=1)   # P2 nodes
        p2 = W_ex(nodes[bc][:, :, 0], nodes[bc][:, :, 1])
        integ = natural_bc_integrals(pa, bc, bl, -base_w[0, bc], p2)        # [nb, 3]
        cells = np.repeat(bc, 3).astype(np.int32)
        rows = lay['facet'][1][bl].ravel().astype(np.int32)
        order = np.argsort(cells, kind='stable')
        pa.natbc_order = np.arange(len(cells))
        pa.natbc_cell = np.ascontiguousarray(cells[order])
        pa.natbc_row = np.ascontiguousarray(rows[order])
        natbc_values = (-integ.ravel())[order]
        # source vector: int v_i f dx on the band's DG1 rows
        J = np.stack([xv[:, 1] - xv[:, 0], xv[:, 2] - xv[:, 0]], axis=2)     # [nc, xy, XY]
        det = np.abs(np.linalg.det(J))
        xq = xv[:, 0][:, None, :] + np.einsum('cab,qb->cqa', J, qp)          # [nc, q, 2]
        fq = f_ex(xq[:, :, 0], xq[:, :, 1])
        src = np.zeros(pa.N)
        dg = pa.cell_dofs[:, lay['dg'][1]]
        src[dg] = np.einsum('cq,q,qi,c->ci', fq, qw, p1q, det)
        orc = Oracle(pa)
        u = np.zeros(pa.N)
        for it in range(12):
            vals, b = orc.assemble(u, base_w=base_w, natbc_values=natbc_values)
            b = b - src
            du = spla.spsolve(orc.to_scipy(vals).tocsc(), b)
            u -= du
            if np.abs(du).max() < 1e-13:
                break
        assert np.abs(du).max() < 1e-11, 'Newton did not converge'
        # phi stays zero (dead Poisson block)
        assert np.abs(u[pa.cell_dofs[:, lay['dg'][0]]]).max() < 1e-12
        # total W = base_w + delta_w at the quadrature points vs exact, L2
        Wq = base_w[0][:, None] + np.einsum('ci,qi->cq', u[dg], p1q)
        errs_w.append(np.sqrt(np.einsum('cq,q,c->', (Wq - W_ex(xq[:, :, 0], xq[:, :, 1])) ** 2, qw, det)))
        jh = np.einsum('cm,qma->cqa', u[pa.cell_dofs[:, lay['bdm'][1]]], bdq)
        sdet = np.linalg.det(J)
        jq = np.einsum('cab,cqb->cqa', J, jh) / sdet[:, None, None]
        ec = (jq[:, :, 0] - c_ex(xq[:, :, 0], xq[:, :, 1])) ** 2 + jq[:, :, 1] ** 2
        errs_c.append(np.sqrt(np.einsum('cq,q,c->', ec, qw, det)))
        # the offsets differ by 0.7 across x = 0.5; the recovered W is smooth there
        assert np.abs(Wq - W_ex(xq[:, :, 0], xq[:, :, 1])).max() < 2e-3
    for e in (errs_w, errs_c):
        assert e[1] < e[0] / 3.5 and e[2] < e[1] / 3.5, e       # second order (BDM2: >= 2)
