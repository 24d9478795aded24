"""a10: the optical sub-problem.  CPU tests pin the oracle
(oracle/optics_oracle.py) against the analytic Beer-Lambert law and exactness
properties; the CUDA kernels are checked against the oracle in
test_emulated_kernels.py / test_gpu_parity.py (hotpath_cases.optics_parity)."""
import numpy as np
import scipy.sparse.linalg as spla

from oracle import optics_oracle as oo
from simudo_b200.fem.cg2 import CG2Space, optics_tables, p2_gradients
from simudo_b200.fem.reference_element import p2_lagrange_eval, triangle_default_scheme
from simudo_b200.mesh import Product2DMesh


def _space(nx, ny, length=2.0):
    m = Product2DMesh(np.linspace(0.0, length, nx), np.linspace(0.0, 0.3, ny)).mesh
    return CG2Space(m)


def _inflow(sp, left=True):
    m = sp.mesh
    bf = m.boundary_facets()
    xm = m.facet_midpoints()[bf, 0]
    return sp.facet_dofs(bf[np.isclose(xm, xm.min() if left else xm.max())])


def test_degree6_rule_is_exact_to_degree_6():
    pts, w = triangle_default_scheme(6)
    from math import factorial
    for a in range(7):
        for b in range(7 - a):
            exact = factorial(a) * factorial(b) / factorial(a + b + 2)
            assert abs((w * pts[:, 0] ** a * pts[:, 1] ** b).sum() - exact) < 2e-15


def test_p2_gradients_match_finite_differences():
    pts = np.array([[0.2, 0.3], [0.6, 0.1]])
    h = 1e-6
    for a in range(2):
        e = np.zeros(2)
        e[a] = h
        fd = (p2_lagrange_eval(pts + e) - p2_lagrange_eval(pts - e)) / (2 * h)
        assert np.abs(fd - p2_gradients(pts)[:, :, a]).max() < 1e-8


def test_projection_of_quadratic_is_exact():
    sp = _space(6, 3)
    x = sp.node_coordinates()
    f = 1.0 + 2 * x[:, 0] - x[:, 1] + 0.5 * x[:, 0] * x[:, 1] + x[:, 0] ** 2
    Mm = oo.mass_matrix(sp)
    sol = spla.spsolve(Mm.tocsc(), Mm @ f)
    assert np.abs(sol - f).max() < 1e-11
    assert np.allclose(sp.mass_values(), Mm.data, rtol=1e-12, atol=1e-14 * abs(Mm.data).max())
    assert np.array_equal(Mm.indices, sp.colidx) and np.array_equal(Mm.indptr, sp.rowptr)


def test_msorte_reproduces_beer_lambert():
    """Omega = (+-1, 0), constant alpha: Phi = Phi0 exp(-alpha s); the relative error at
    the far end (accumulated attenuation) converges with second order in h."""
    alpha, phi0 = 1.7, 3.0e4
    errs = []
    for nx in (9, 17, 33):
        sp = _space(nx, 3)
        x = sp.node_coordinates()[:, 0]
        bc = _inflow(sp, left=True)
        A, b = oo.msorte_system(sp, np.full(sp.n_dofs, alpha), (1.0, 0.0), bc,
                                np.full(len(bc), phi0))
        phi = spla.spsolve(A.tocsc(), b)
        errs.append(np.abs(phi / (phi0 * np.exp(-alpha * x)) - 1).max())
    assert errs[-1] < 3e-3
    assert errs[0] / errs[1] > 3.5 and errs[1] / errs[2] > 3.5
    sp = _space(17, 3)
    x = sp.node_coordinates()[:, 0]
    bc = _inflow(sp, left=False)
    A, b = oo.msorte_system(sp, np.full(sp.n_dofs, alpha), (-1.0, 0.0), bc,
                            np.full(len(bc), phi0))
    phi = spla.spsolve(A.tocsc(), b)
    assert np.abs(phi / (phi0 * np.exp(-alpha * (2.0 - x))) - 1).max() < 1e-2


def test_piecewise_absorption_layers():
    """alpha switching between layers (CG2-interpolated): flux stays monotone and the
    total attenuation matches the integral of the interpolated alpha."""
    sp = _space(41, 3)
    x = sp.node_coordinates()[:, 0]
    a = np.where(x < 1.0, 0.4, 2.5)
    bc = _inflow(sp)
    A, b = oo.msorte_system(sp, a, (1.0, 0.0), bc, np.ones(len(bc)))
    phi = spla.spsolve(A.tocsc(), b)
    right = np.isclose(x, 2.0)
    assert abs(np.log(phi[right]).mean() + (0.4 * 1.0 + 2.5 * 1.0)) < 0.08
    assert phi.min() > 0


def test_optics_tables_identities():
    t = optics_tables()
    K = np.ctypeslib.as_array(t.K)
    T = np.ctypeslib.as_array(t.T)
    F = np.ctypeslib.as_array(t.F)
    # constants are in the kernel of the stiffness part; sum_k T[a,i,j,k] = int N_i d_a N_j
    assert np.abs(K.sum(axis=3)).max() < 1e-13 and np.abs(K.sum(axis=2)).max() < 1e-13
    assert np.abs(F.sum(axis=(1, 2, 3)) - 1.0).max() < 1e-13
    assert np.abs(T.sum(axis=(2, 3))).max() < 1e-13         # int N_i d_a(1) = 0
