"""ORACLE (test infrastructure, not product code): FIAT-style element tables.

PARITY UNPINNED: the arithmetic of the reference's hot path lives in legacy
FEniCS (FIAT / FFC / UFL / dolfin 2018.1-2019.x), which is neither vendored in
/root/reference nor importable here, and no reference test pins a number at
that boundary.  This module restates FIAT's published
construction of the elements and quadrature rules the reference selects at
  simudo/physics/poisson_drift_diffusion1.py1:59-72,412-425  (DG1 x BDM2)
  simudo/physics/poisson_drift_diffusion.py:461,699-700      (degree 20 / 8)
  simudo/physics/optical1.py1:48                              (CG2)
by a route deliberately different from simudo_b200/fem/reference_element.py:

* BDM2: exact rational arithmetic with sympy -- monomial primal space,
  FIAT's dual set (scaled-normal point evaluations at the edge lattice points
  1/4, 1/2, 3/4; integral moments against the degree-1 Nedelec basis), exact
  inversion.  (FIAT brezzi_douglas_marini.py, BDMDualSet.)
* Gauss-Jacobi nodes: Newton iteration from Chebyshev guesses with deflation
  on the three-term Jacobi recurrence (FIAT jacobi.py / gaussjacobi.py), not
  scipy's eigenvalue solver.
* Collapsed triangle rule: FIAT CollapsedQuadratureTriangleRule -- x from
  GJ(0,0), y from GJ(1,0), xi = ((1+eta1)(1-eta2)/2 - 1, eta2), affine map to
  the UFC triangle, weight wx wy / 8, x index outer.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this file.
"""
import functools
import math

import numpy as np
import sympy as sp

__all__ = ['jacobi_eval', 'gauss_jacobi', 'collapsed_triangle_rule',
           'default_triangle_scheme', 'gauss_legendre_01', 'bdm2_symbolic',
           'tabulate_p1', 'tabulate_p2', 'tabulate_bdm2', 'facet_tabulation',
           'REF_EDGES', 'REF_SCALED_NORMALS']


# ---------------------------------------------------------------------------
# Jacobi polynomials by recurrence; Gauss-Jacobi by Newton + deflation
# ---------------------------------------------------------------------------
def jacobi_eval(n, a, b, x):
    """P_n^{(a,b)}(x) and its derivative."""
    if n == 0:
        return 1.0, 0.0
    p0 = 1.0
    p1 = 0.5 * (a - b + (a + b + 2.0) * x)
    for k in range(1, n):
        a1 = 2.0 * (k + 1.0) * (k + a + b + 1.0) * (2.0 * k + a + b)
        a2 = (2.0 * k + a + b + 1.0) * (a * a - b * b)
        a3 = (2.0 * k + a + b) * (2.0 * k + a + b + 1.0) * (2.0 * k + a + b + 2.0)
        a4 = 2.0 * (k + a) * (k + b) * (2.0 * k + a + b + 2.0)
        p0, p1 = p1, ((a2 + a3 * x) * p1 - a4 * p0) / a1
    # derivative: d/dx P_n^{a,b} = (n+a+b+1)/2 P_{n-1}^{a+1,b+1}
    d, _ = jacobi_eval(n - 1, a + 1.0, b + 1.0, x)
    return p1, 0.5 * (n + a + b + 1.0) * d


def gauss_jacobi(m, a, b):
    """m-point rule on [-1, 1] for the weight (1-x)^a (1+x)^b."""
    x = []
    for k in range(m):
        r = -math.cos((2.0 * k + 1.0) * math.pi / (2.0 * m))
        if k > 0:
            r = 0.5 * (r + x[k - 1])
        for _ in range(100):
            s = sum(1.0 / (r - xi) for xi in x)
            f, fp = jacobi_eval(m, a, b, r)
            delta = f / (fp - f * s)
            r -= delta
            if abs(delta) < 1e-16:
                break
        x.append(r)
    x = np.array(x)
    # weights (standard closed form)
    lg = math.lgamma
    c = (math.exp(lg(a + m + 1.0) + lg(b + m + 1.0) - lg(a + b + m + 1.0) - lg(m + 1.0))
         * 2.0 ** (a + b + 1.0))
    w = np.array([c / ((1.0 - xi * xi) * jacobi_eval(m, a, b, xi)[1] ** 2) for xi in x])
    return x, w


def collapsed_triangle_rule(m):
    """Points [m*m, 2] and weights [m*m] on the UFC triangle, FIAT order."""
    ptx, wx = gauss_jacobi(m, 0.0, 0.0)
    pty, wy = gauss_jacobi(m, 1.0, 0.0)
    pts, wts = [], []
    for i in range(m):
        for j in range(m):
            xi1 = 0.5 * (1.0 + ptx[i]) * (1.0 - pty[j]) - 1.0
            xi2 = pty[j]
            pts.append((0.5 * (xi1 + 1.0), 0.5 * (xi2 + 1.0)))
            wts.append(0.5 * 0.25 * wx[i] * wy[j])
    return np.array(pts), np.array(wts)


def default_triangle_scheme(degree):
    """FIAT ``create_quadrature(triangle, degree, 'default')`` for the degrees
    the reference's forms hit: 4 -> 6 points, > 6 -> collapsed rule."""
    if degree == 4:
        x = np.array([[0.816847572980459, 0.091576213509771],
                      [0.091576213509771, 0.816847572980459],
                      [0.091576213509771, 0.091576213509771],
                      [0.108103018168070, 0.445948490915965],
                      [0.445948490915965, 0.108103018168070],
                      [0.445948490915965, 0.445948490915965]])
        w = np.empty(6)
        w[0:3] = 0.109951743655322
        w[3:6] = 0.223381589678011
        return x, w / 2.0
    if degree > 6:
        return collapsed_triangle_rule((degree + 2) // 2)
    raise NotImplementedError(degree)


def gauss_legendre_01(n):
    x, w = gauss_jacobi(n, 0.0, 0.0)
    return 0.5 * (x + 1.0), 0.5 * w


# ---------------------------------------------------------------------------
# BDM2, exact
# ---------------------------------------------------------------------------
REF_VERTS = [(0, 0), (1, 0), (0, 1)]
REF_EDGES = [(1, 2), (0, 2), (0, 1)]            # edge i opposite vertex i
REF_SCALED_NORMALS = []
for _a, _b in REF_EDGES:
    _t = (REF_VERTS[_b][0] - REF_VERTS[_a][0], REF_VERTS[_b][1] - REF_VERTS[_a][1])
    REF_SCALED_NORMALS.append((_t[1], -_t[0]))  # tangent rotated clockwise


@functools.lru_cache(maxsize=None)
def bdm2_symbolic():
    """Exact BDM2 nodal basis: list of 12 pairs (fx, fy) of sympy polynomials."""
    X, Y = sp.symbols('X Y')
    mono = [sp.Integer(1), X, Y, X * X, X * Y, Y * Y]
    prim = [(m, 0) for m in mono] + [(0, m) for m in mono]

    def tri_int(e):
        return sp.integrate(sp.integrate(e, (Y, 0, 1 - X)), (X, 0, 1))

    lam = [1 - X - Y, X, Y]
    grad = [(-1, -1), (1, 0), (0, 1)]
    ned = []
    for a, b in REF_EDGES:
        ned.append((lam[a] * grad[b][0] - lam[b] * grad[a][0],
                    lam[a] * grad[b][1] - lam[b] * grad[a][1]))

    def functionals(f):
        out = []
        for (a, b), n in zip(REF_EDGES, REF_SCALED_NORMALS):
            for t in (sp.Rational(1, 4), sp.Rational(1, 2), sp.Rational(3, 4)):
                px = REF_VERTS[a][0] + t * (REF_VERTS[b][0] - REF_VERTS[a][0])
                py = REF_VERTS[a][1] + t * (REF_VERTS[b][1] - REF_VERTS[a][1])
                sub = {X: px, Y: py}
                out.append(sp.sympify(f[0]).subs(sub) * n[0]
                           + sp.sympify(f[1]).subs(sub) * n[1])
        for N in ned:
            out.append(tri_int(f[0] * N[0] + f[1] * N[1]))
        return out

    V = sp.Matrix([functionals(f) for f in prim])        # V[k, i] = l_i(prim_k)
    Cmat = V.inv()                                        # basis_j = sum_k Cmat[j,k] prim_k
    basis = []
    for j in range(12):
        fx = sum(Cmat[j, k] * prim[k][0] for k in range(12))
        fy = sum(Cmat[j, k] * prim[k][1] for k in range(12))
        basis.append((sp.expand(fx), sp.expand(fy)))
    return (X, Y), basis


@functools.lru_cache(maxsize=None)
def _bdm2_lambdas():
    (X, Y), basis = bdm2_symbolic()
    fx = [sp.lambdify((X, Y), b[0], 'numpy') for b in basis]
    fy = [sp.lambdify((X, Y), b[1], 'numpy') for b in basis]
    dv = [sp.lambdify((X, Y), sp.diff(b[0], X) + sp.diff(b[1], Y), 'numpy') for b in basis]
    return fx, fy, dv


def tabulate_bdm2(pts):
    """val[n, 12, 2], div[n, 12] of the reference basis at pts[n, 2]."""
    pts = np.asarray(pts, dtype=float)
    fx, fy, dv = _bdm2_lambdas()
    x, y = pts[:, 0], pts[:, 1]
    one = np.ones_like(x)
    val = np.empty((pts.shape[0], 12, 2))
    div = np.empty((pts.shape[0], 12))
    for i in range(12):
        val[:, i, 0] = fx[i](x, y) * one
        val[:, i, 1] = fy[i](x, y) * one
        div[:, i] = dv[i](x, y) * one
    return val, div


def tabulate_p1(pts):
    pts = np.asarray(pts, dtype=float)
    return np.stack([1.0 - pts[:, 0] - pts[:, 1], pts[:, 0], pts[:, 1]], axis=1)


def tabulate_p2(pts):
    """P2 Lagrange: vertices 0,1,2 then midpoints of edges 0,1,2."""
    l = tabulate_p1(pts)
    return np.stack([l[:, 0] * (2 * l[:, 0] - 1), l[:, 1] * (2 * l[:, 1] - 1),
                     l[:, 2] * (2 * l[:, 2] - 1), 4 * l[:, 1] * l[:, 2],
                     4 * l[:, 0] * l[:, 2], 4 * l[:, 0] * l[:, 1]], axis=1)


def facet_tabulation(npts):
    """Per reference edge f: Gauss-Legendre points on it, the BDM2 scaled-normal
    traces, P1 and P2 values there.  trace[f, q, i] = psi_i(p_q) . n_f."""
    t, w = gauss_legendre_01(npts)
    trace = np.zeros((3, npts, 12))
    p1 = np.zeros((3, npts, 3))
    p2 = np.zeros((3, npts, 6))
    for f, ((a, b), n) in enumerate(zip(REF_EDGES, REF_SCALED_NORMALS)):
        va, vb = np.array(REF_VERTS[a], float), np.array(REF_VERTS[b], float)
        pts = va[None, :] + t[:, None] * (vb - va)[None, :]
        val, _ = tabulate_bdm2(pts)
        trace[f] = val[:, :, 0] * n[0] + val[:, :, 1] * n[1]
        p1[f] = tabulate_p1(pts)
        p2[f] = tabulate_p2(pts)
    return t, w, trace, p1, p2
