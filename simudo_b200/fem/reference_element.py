"""Reference-triangle tables for the assembly kernels (host side, numpy).

Everything the CUDA path needs about the reference's finite elements and
quadrature, derived from first principles (nothing here imports ``oracle/``):

* UFC reference triangle (0,0),(1,0),(0,1); local facet i opposite vertex i
  (upstream FIAT/UFC, not in /root/reference).
* DG1 / P2 Lagrange nodal bases (``poisson_drift_diffusion1.py1:60,413``;
  ``poisson_drift_diffusion.py:405``; ``optical1.py1:48``).
* BDM2 nodal basis: three scaled-normal point evaluations per edge at
  t = 1/4, 1/2, 3/4 plus three interior moments against the lowest-order
  Nedelec (Whitney) functions (``poisson_drift_diffusion1.py1:67,420``).
* Collapsed Gauss-Jacobi rules for degree > 6 (``m = (deg+2)//2`` points per
  axis: degree 20 -> 11x11, degree 8 -> 5x5;
  ``poisson_drift_diffusion.py:461,699-700``), in the *factored* form
  ``x = a_i (1 - b_j), y = b_j`` that the kernels use for sum factorisation.
* An orthonormal Dubiner basis on the triangle, separable in the collapsed
  coordinates, against which the kernels take moments of the non-polynomial
  coefficients; and the constant expansion tensors that turn those moments
  into element-matrix entries.
"""
import functools

import numpy as np
from scipy.special import roots_jacobi, eval_jacobi, eval_legendre

__all__ = ['gauss_jacobi_rule', 'collapsed_rule_factors', 'collapsed_rule',
           'gauss_legendre_01', 'triangle_default_scheme', 'dubiner_index',
           'dubiner_factors', 'dubiner_eval', 'P1_LAGRANGE', 'p2_lagrange_eval',
           'bdm2_coefficients', 'ReferenceTables', 'reference_tables']


# ---------------------------------------------------------------------------
# quadrature
# ---------------------------------------------------------------------------
def gauss_jacobi_rule(m, a, b):
    """Nodes/weights on [-1,1] for weight (1-x)^a (1+x)^b."""
    x, w = roots_jacobi(m, a, b)
    return np.asarray(x, dtype=np.float64), np.asarray(w, dtype=np.float64)


def collapsed_rule_factors(m):
    """1D factors of the m x m collapsed rule on the UFC triangle.

    Returns ``(a, wa, b, wb)`` with points ``(a_i (1-b_j), b_j)`` and weights
    ``wa_i * wb_j`` (sum = 1/2); point order is i outer, j inner
    (SURVEY.md Appendix A5).
    """
    x, wx = gauss_jacobi_rule(m, 0.0, 0.0)
    y, wy = gauss_jacobi_rule(m, 1.0, 0.0)
    return 0.5 * (x + 1.0), 0.5 * wx, 0.5 * (y + 1.0), 0.25 * wy


def collapsed_rule(m):
    a, wa, b, wb = collapsed_rule_factors(m)
    X = (a[:, None] * (1.0 - b[None, :])).ravel()
    Y = np.broadcast_to(b[None, :], (m, m)).ravel()
    W = (wa[:, None] * wb[None, :]).ravel()
    return np.stack([X, Y], axis=1), W


def gauss_legendre_01(n):
    x, w = np.polynomial.legendre.leggauss(n)
    return 0.5 * (x + 1.0), 0.5 * w


def triangle_default_scheme(degree):
    """FIAT/FFC ``scheme='default'`` point sets on the UFC triangle
    (Appendix A5): tabulated low-degree rules, collapsed GJ above degree 6.
    The digits are the published 15-digit Strang-Fix/Dunavant values."""
    if degree <= 1:
        return np.array([[1.0 / 3.0, 1.0 / 3.0]]), np.array([0.5])
    if degree == 2:
        x = np.array([[1 / 6, 1 / 6], [1 / 6, 2 / 3], [2 / 3, 1 / 6]])
        return x, np.full(3, 1.0 / 6.0)
    if degree == 3:
        x = np.array([[0.659027622374092, 0.231933368553031],
                      [0.659027622374092, 0.109039009072877],
                      [0.231933368553031, 0.659027622374092],
                      [0.231933368553031, 0.109039009072877],
                      [0.109039009072877, 0.659027622374092],
                      [0.109039009072877, 0.231933368553031]])
        return x, np.full(6, 1.0 / 12.0)
    if degree == 4:
        x = np.array([[0.816847572980459, 0.091576213509771],
                      [0.091576213509771, 0.816847572980459],
                      [0.091576213509771, 0.091576213509771],
                      [0.108103018168070, 0.445948490915965],
                      [0.445948490915965, 0.108103018168070],
                      [0.445948490915965, 0.445948490915965]])
        w = np.array([0.109951743655322] * 3 + [0.223381589678011] * 3) / 2.0
        return x, w
    if degree == 5:
        x = np.array([[0.33333333333333333, 0.33333333333333333],
                      [0.79742698535308720, 0.10128650732345633],
                      [0.10128650732345633, 0.79742698535308720],
                      [0.10128650732345633, 0.10128650732345633],
                      [0.05971587178976981, 0.47014206410511505],
                      [0.47014206410511505, 0.05971587178976981],
                      [0.47014206410511505, 0.47014206410511505]])
        w = np.array([0.22500000000000000] + [0.12593918054482717] * 3
                     + [0.13239415278850616] * 3) / 2.0
        return x, w
    if degree == 6:
        x = np.array([[0.873821971016996, 0.063089014491502],
                      [0.063089014491502, 0.873821971016996],
                      [0.063089014491502, 0.063089014491502],
                      [0.501426509658179, 0.249286745170910],
                      [0.249286745170910, 0.501426509658179],
                      [0.249286745170910, 0.249286745170910],
                      [0.636502499121399, 0.310352451033785],
                      [0.636502499121399, 0.053145049844816],
                      [0.310352451033785, 0.636502499121399],
                      [0.310352451033785, 0.053145049844816],
                      [0.053145049844816, 0.636502499121399],
                      [0.053145049844816, 0.310352451033785]])
        w = np.array([0.050844906370207] * 3 + [0.116786275726379] * 3
                     + [0.082851075618374] * 6) / 2.0
        return x, w
    return collapsed_rule((degree + 2) // 2)


# ---------------------------------------------------------------------------
# orthonormal Dubiner basis, separable in collapsed coordinates
#   Psi_pq(x, y) = g_p(a) * h_pq(b),  a = x/(1-y), b = y
# ---------------------------------------------------------------------------
def dubiner_index(degree):
    """[(p, q)] ordered by total degree then p: hierarchical, so the first
    (d+1)(d+2)/2 functions span P_d."""
    idx = []
    for n in range(degree + 1):
        for p in range(n + 1):
            idx.append((p, n - p))
    return idx


def dubiner_factors(degree, a, b):
    """g[p, i] = g_p(a_i); h[(p,q)] = h_pq(b_j) as array h[k, j] in
    ``dubiner_index`` order."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    g = np.stack([np.sqrt(2 * p + 1.0) * eval_legendre(p, 2 * a - 1)
                  for p in range(degree + 1)])
    h = np.stack([np.sqrt(2.0 * (p + q + 1)) * (1 - b) ** p
                  * eval_jacobi(q, 2 * p + 1, 0, 2 * b - 1)
                  for p, q in dubiner_index(degree)])
    return g, h


def dubiner_eval(degree, pts):
    """Psi[k, n] at arbitrary points (x, y), y < 1."""
    pts = np.asarray(pts, dtype=np.float64)
    x, y = pts[:, 0], pts[:, 1]
    one_m = 1.0 - y
    safe = np.where(np.abs(one_m) > 1e-300, one_m, 1.0)
    a = np.where(np.abs(one_m) > 1e-300, x / safe, 0.0)
    out = []
    for p, q in dubiner_index(degree):
        # (1-y)^p P_p(2a-1) is a polynomial in (x, y); evaluate stably
        gp = np.sqrt(2 * p + 1.0) * eval_legendre(p, 2 * a - 1) * one_m ** p
        hq = np.sqrt(2.0 * (p + q + 1)) * eval_jacobi(q, 2 * p + 1, 0, 2 * y - 1)
        out.append(gp * hq)
    return np.stack(out)


# ---------------------------------------------------------------------------
# nodal bases
# ---------------------------------------------------------------------------
# lambda_i = c0 + cx X + cy Y  (DG1 local dof i = vertex i, Appendix A2)
P1_LAGRANGE = np.array([[1.0, -1.0, -1.0], [0.0, 1.0, 0.0], [0.0, 0.0, 1.0]])


def p1_eval(pts):
    pts = np.asarray(pts, dtype=np.float64)
    return np.stack([1 - pts[:, 0] - pts[:, 1], pts[:, 0], pts[:, 1]])


def p2_lagrange_eval(pts):
    """P2 Lagrange (DG2/CG2): 3 vertex nodes then edge midpoints, edge i
    opposite vertex i (Appendix A2). Returns N[6, n]."""
    l = p1_eval(pts)
    return np.stack([l[0] * (2 * l[0] - 1), l[1] * (2 * l[1] - 1),
                     l[2] * (2 * l[2] - 1),
                     4 * l[1] * l[2], 4 * l[0] * l[2], 4 * l[0] * l[1]])


P2_NODES = np.array([[0, 0], [1, 0], [0, 1], [.5, .5], [0, .5], [.5, 0]], dtype=float)

# reference edges: (vertex a, vertex b), a < b; scaled normal = rot_cw(b - a)
_REF_VERTS = np.array([[0.0, 0.0], [1.0, 0.0], [0.0, 1.0]])
_REF_EDGES = [(1, 2), (0, 2), (0, 1)]


def reference_edge_points(f, t):
    a, b = _REF_EDGES[f]
    return _REF_VERTS[a] + np.asarray(t)[:, None] * (_REF_VERTS[b] - _REF_VERTS[a])


def reference_scaled_normal(f):
    a, b = _REF_EDGES[f]
    t = _REF_VERTS[b] - _REF_VERTS[a]
    return np.array([t[1], -t[0]])


def _whitney(pts):
    """Lowest-order Nedelec (first kind) basis on the UFC triangle, one per
    edge with unit tangential moment: N_e = l_a grad l_b - l_b grad l_a."""
    X, Y = pts[:, 0], pts[:, 1]
    return np.stack([np.stack([-Y, X], axis=1),
                     np.stack([Y, 1 - X], axis=1),
                     np.stack([1 - Y, X], axis=1)])


@functools.lru_cache(maxsize=None)
def bdm2_coefficients():
    """C[i, comp, r]: BDM2 nodal basis function i, component comp, expanded in
    the orthonormal Dubiner P2 basis (6 functions).

    Dual basis (Appendix A3): dofs 3f+k = psi(p_fk) . n_f for edge f, points at
    1/4, 1/2, 3/4 from the lower to the higher local vertex, n_f the clockwise
    rotated edge vector (length-scaled); dofs 9..11 = integral of psi . N_e.
    """
    nb = 12
    # generic vector P2 function = sum_{comp, r} coef[comp, r] Psi_r e_comp
    V = np.zeros((nb, 2, 6))
    row = 0
    for f in range(3):
        pts = reference_edge_points(f, [0.25, 0.5, 0.75])
        n = reference_scaled_normal(f)
        psi = dubiner_eval(2, pts)            # [6, 3]
        for k in range(3):
            V[row, 0, :] = n[0] * psi[:, k]
            V[row, 1, :] = n[1] * psi[:, k]
            row += 1
    qp, qw = collapsed_rule(6)
    psi = dubiner_eval(2, qp)                 # [6, nq]
    Ned = _whitney(qp)                        # [3, nq, 2]
    for e in range(3):
        for comp in range(2):
            V[row, comp, :] = psi @ (qw * Ned[e, :, comp])
        row += 1
    Vm = V.reshape(nb, 12)
    Cm = np.linalg.inv(Vm)                    # columns = nodal basis coefficients
    return np.ascontiguousarray(Cm.T.reshape(nb, 2, 6))


def bdm2_eval(pts):
    """Reference BDM2 basis psi_hat[i, n, comp] and div_hat[i, n]."""
    C = bdm2_coefficients()
    pts = np.asarray(pts, dtype=np.float64)
    psi = dubiner_eval(2, pts)
    val = np.einsum('icr,rn->inc', C, psi)
    # divergence via finite-difference-free route: differentiate the P2
    # expansion exactly by re-projecting monomials
    dpsi = _dubiner_p2_gradients(pts)         # [6, n, 2]
    div = np.einsum('icr,rnc->in', C, dpsi)
    return val, div


@functools.lru_cache(maxsize=None)
def _dubiner_p2_monomial_matrix():
    """M[r, k]: Psi_r = sum_k M[r,k] mono_k, mono = 1, x, y, x^2, xy, y^2."""
    pts, _ = collapsed_rule(4)
    x, y = pts[:, 0], pts[:, 1]
    mono = np.stack([np.ones_like(x), x, y, x * x, x * y, y * y])
    psi = dubiner_eval(2, pts)
    M, *_ = np.linalg.lstsq(mono.T, psi.T, rcond=None)
    return np.ascontiguousarray(M.T)


def _dubiner_p2_gradients(pts):
    M = _dubiner_p2_monomial_matrix()
    x, y = pts[:, 0], pts[:, 1]
    z, o = np.zeros_like(x), np.ones_like(x)
    dmx = np.stack([z, o, z, 2 * x, y, z])
    dmy = np.stack([z, z, o, z, x, 2 * y])
    return np.stack([M @ dmx, M @ dmy], axis=2)


# ---------------------------------------------------------------------------
# table bundle handed to the C-ABI plan
# ---------------------------------------------------------------------------
class ReferenceTables(object):
    """Constant tensors used by the kernels (all float64, C-contiguous).

    Moments are taken against the orthonormal Dubiner basis ``Psi_t``:
    ``m_t = sum_pts w f Psi_t``.  Notation: ``pi_r`` = the first 6 (P2),
    ``lam_i`` = P1 Lagrange.

    ``W_rs = int f pi_r pi_s = sum_t G22[r,s,t] m_t``          (t < 15, P4)
    ``int f pi_r lam_l      = sum_t G21[r,l,t] m_t``          (t < 10, P3)
    ``int f lam_i lam_j     = sum_t G11[i,j,t] m_t``          (t < 6,  P2)
    ``int f lam_i           = sum_t G1[i,t] m_t``             (t < 3,  P1)
    ``int f N2_a lam_i ...`` similar for the P2-Lagrange coefficient fields.
    """

    def __init__(self):
        self.bdm_C = bdm2_coefficients()                  # [12, 2, 6]
        qp, qw = collapsed_rule(8)                        # exact to degree 15
        psi4 = dubiner_eval(4, qp)                        # [15, n]
        lam = p1_eval(qp)                                 # [3, n]
        pi = psi4[:6]
        self.G22 = np.einsum('rn,sn,tn,n->rst', pi, pi, psi4, qw)          # [6,6,15]
        self.G21 = np.einsum('rn,ln,tn,n->rlt', pi, lam, psi4[:10], qw)    # [6,3,10]
        self.G11 = np.einsum('in,jn,tn,n->ijt', lam, lam, psi4[:6], qw)    # [3,3,6]
        self.G1 = np.einsum('in,tn,n->it', lam, psi4[:3], qw)              # [3,3]
        # polynomial (exactly integrated) reference matrices
        val, div = bdm2_eval(qp)                          # [12,n,2], [12,n]
        self.bdm_mass = np.einsum('ina,jnb,n->ijab', val, val, qw)         # [12,12,2,2]
        self.bdm_div_p1 = np.einsum('in,jn,n->ij', div, lam, qw)           # [12,3]
        self.p1_mass = np.einsum('in,jn,n->ij', lam, lam, qw)              # [3,3]
        # divergence of each BDM function in the P1 Lagrange basis
        self.bdm_div_coef = np.linalg.solve(self.p1_mass, self.bdm_div_p1.T).T  # [12,3]
        # P2 Lagrange <-> Dubiner P2
        N2 = p2_lagrange_eval(qp)
        self.p2lag_to_dub = np.einsum('an,rn,n->ar', N2, pi, qw)           # [6,6]
        self.p1_to_dub = np.einsum('in,tn,n->it', lam, psi4[:3], qw)       # [3,3]
        # facet tables: BDM normal traces on each reference edge at GL points
        self.facet_gl_t, self.facet_gl_w = gauss_legendre_01(3)
        tr = np.zeros((3, 12, 3))          # [facet, bdm dof, gl point]
        p1f = np.zeros((3, 3, 3))          # [facet, p1 dof, gl point]
        p2f = np.zeros((3, 6, 3))          # [facet, p2 dof, gl point]
        for f in range(3):
            pts = reference_edge_points(f, self.facet_gl_t)
            v, _ = bdm2_eval(pts)
            n = reference_scaled_normal(f)
            tr[f] = v[:, :, 0] * n[0] + v[:, :, 1] * n[1]
            p1f[f] = p1_eval(pts)
            p2f[f] = p2_lagrange_eval(pts)
        self.facet_bdm_trace = tr
        self.facet_p1 = p1f
        self.facet_p2 = p2f

    def rule(self, m):
        """Factored collapsed rule + separable Dubiner factors at its nodes."""
        a, wa, b, wb = collapsed_rule_factors(m)
        g, h = dubiner_factors(4, a, b)
        return dict(a=a, wa=wa, b=b, wb=wb, g=g, h=h)


@functools.lru_cache(maxsize=None)
def reference_tables():
    return ReferenceTables()
