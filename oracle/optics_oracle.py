"""ORACLE -- TEST INFRASTRUCTURE ONLY (never imported by the product).

CPU restatement (numpy / scipy, physical-space quadrature) of the optical
sub-problem, SURVEY.md section 8 row a10:

* absorption coefficient per optical field on the PDD mesh,
  ``simudo/physics/optical.py:213-224`` summing
  ``electro_optical_process.py:415-418`` (band-to-band: constant alpha inside the
  process' absorption window) and ``:464-480`` (intermediate band:
  ``sigma_opt * (u_I  or  N_I - u_I)``);
* its L2 projection onto CG2 (``optical.py:260-274`` -> ``fem/assign.py:12-47``,
  ``dolfin.project``), right-hand side with the degree-6 12-point default rule
  (SURVEY Appendix A5/A6);
* the MSORTE form ``physics/optical1.py1:27-45`` with S = 0 and the inflow
  Dirichlet condition ``:57-61``.

Unlike the CUDA kernels (constant reference tensors contracted with J^-1 Omega)
everything here is evaluated point by point in physical space: physical
gradients ``J^-T grad N``, alpha and grad alpha at the quadrature points,
outward unit normals from the facet geometry.

Shared with the product are only elementary tables: the P2 Lagrange basis and its
gradients and the published quadrature point sets
(``simudo_b200.fem.reference_element`` / ``cg2.p2_gradients``), themselves pinned in
tests/test_optics.py (exactness of the degree-6 rule, finite-difference check of the
gradients).  Mass / MSORTE matrices and right-hand sides are assembled here from
scratch.

Parity status: unpinned against dolfin (FEniCS is not installable here); pinned
against the analytic Beer-Lambert solution exp(-alpha x), which the second-order
form reproduces to discretisation error (tests/test_optics.py).
"""
import numpy as np
import scipy.sparse as sps
import scipy.sparse.linalg as spla

from simudo_b200.fem import model as M
from simudo_b200.fem.cg2 import p2_gradients
from simudo_b200.fem.reference_element import (collapsed_rule, gauss_legendre_01,
                                               p2_lagrange_eval, reference_edge_points,
                                               triangle_default_scheme)

__all__ = ['alpha_at_points', 'alpha_rhs', 'mass_matrix', 'msorte_system', 'solve_field']


def _band_u(kind, sign, N, E, kT, chi):
    x = sign * (chi - E) / kT
    if kind == M.BAND_NONDEGENERATE:
        return N * np.exp(-x)
    return N / (1.0 + np.exp(x))


def alpha_at_points(pa, u, base_w, field, ref_pts):
    """alpha_field at reference points of every cell: [Nc, n_pts]."""
    model, tp = pa.model, pa.tag_params[pa.cell_tag]          # [Nc, stride]
    l = np.stack([1 - ref_pts[:, 0] - ref_pts[:, 1], ref_pts[:, 0], ref_pts[:, 1]])  # [3, n]
    lay = pa.layout()
    phi = u[pa.cell_dofs[:, lay['dg'][0]]] @ l                 # [Nc, n]
    out = np.zeros_like(phi)
    kT = tp[:, M.TP_KT][:, None]
    for p in range(model.n_procs):
        coef = tp[:, M.TP_PROC0 + 8 * p + M.TPP_FIELD0 + field][:, None]
        kind = model.proc_kind[p]
        if kind == M.PROC_RAD_BB:
            out += coef
        elif kind == M.PROC_RAD_IB:
            IB = model.proc_trap[p]
            RB = model.proc_src[p] if model.proc_dst[p] == IB else model.proc_dst[p]
            w = base_w[IB][:, None] + u[pa.cell_dofs[:, lay['dg'][1 + IB]]] @ l
            N = tp[:, M.TP_BAND0 + 5 * IB + M.TPB_N][:, None]
            E = tp[:, M.TP_BAND0 + 5 * IB + M.TPB_E][:, None]
            uI = _band_u(model.band_kind[IB], float(model.band_sign[IB]), N, E, kT, phi + w)
            same = model.band_sign[IB] == model.band_sign[RB]
            out += coef * (uI if same else (N - uI))
    return out


def alpha_rhs(pa, space, u, base_w, field):
    pts, w = triangle_default_scheme(6)
    a = alpha_at_points(pa, u, base_w, field, pts)             # [Nc, 12]
    N = p2_lagrange_eval(pts)                                  # [6, 12]
    _, det = pa.mesh.cell_jacobians()
    loc = np.abs(det)[:, None] * np.einsum('cq,q,iq->ci', a, w, N)
    rhs = np.zeros(space.n_dofs)
    np.add.at(rhs, space.cell_dofs.ravel(), loc.ravel())
    return rhs


def _csr(space, elem):
    rows = np.repeat(space.cell_dofs, 6, axis=1).ravel()
    cols = np.tile(space.cell_dofs, (1, 6)).ravel()
    A = sps.coo_matrix((elem.ravel(), (rows, cols)), shape=(space.n_dofs,) * 2).tocsr()
    A.sum_duplicates()
    A.sort_indices()
    return A


def mass_matrix(space):
    pts, w = collapsed_rule(3)                                 # exact to degree 5 >= 4
    N = p2_lagrange_eval(pts)
    _, det = space.mesh.cell_jacobians()
    elem = np.abs(det)[:, None, None] * np.einsum('q,iq,jq->ij', w, N, N)[None]
    return _csr(space, elem)


def msorte_system(space, alpha_nodal, omega, bc_dofs, bc_vals):
    """(A csr, b) of the MSORTE problem, Dirichlet rows replaced by identity."""
    mesh = space.mesh
    omega = np.asarray(omega, dtype=float)
    pts, w = collapsed_rule(4)                                 # exact to degree 7 >= 5
    N = p2_lagrange_eval(pts)                                  # [6, q]
    dN = p2_gradients(pts)                                     # [6, q, 2]
    J, det = mesh.cell_jacobians()
    Jinv = np.linalg.inv(J)                                    # [c, 2, 2]
    # physical gradients: grad N = J^-T grad_ref N
    G = np.einsum('cab,iqa->ciqb', Jinv, dN)                   # [c, 6, q, 2]
    oG = np.einsum('ciqb,b->ciq', G, omega)                    # Omega . grad N_i
    al = alpha_nodal[space.cell_dofs]                          # [c, 6]
    a_q = al @ N                                               # alpha at points [c, q]
    oga_q = np.einsum('ck,ckq->cq', al, oG)                    # Omega . grad alpha
    # Omega . grad(alpha N_j) = alpha Omega.grad N_j + N_j Omega.grad alpha
    ograd = a_q[:, None, :] * oG + N[None] * oga_q[:, None, :]  # [c, j, q]
    elem = np.abs(det)[:, None, None] * (
        -np.einsum('q,cjq,ciq->cij', w, oG, oG) + np.einsum('q,iq,cjq->cij', w, N, ograd))
    # exterior facets: - oint (Omega . n) v alpha Phi
    bf = mesh.boundary_facets()
    nrm = mesh.boundary_outward_normals(bf)
    length = mesh.facet_lengths()[bf]
    gt, gw = gauss_legendre_01(4)
    for f, n, ln in zip(bf, nrm, length):
        c, lf = mesh.facet_cells[f, 0], mesh.facet_local[f, 0]
        Nf = p2_lagrange_eval(reference_edge_points(int(lf), gt))      # [6, g]
        af = al[c] @ Nf
        elem[c] -= float(omega @ n) * ln * np.einsum('g,ig,g,jg->ij', gw, Nf, af, Nf)
    A = _csr(space, elem).tolil()
    b = np.zeros(space.n_dofs)
    for d, v in zip(bc_dofs, bc_vals):
        A.rows[d] = [int(d)]
        A.data[d] = [1.0]
        b[d] = v
    return A.tocsr(), b


def solve_field(pa, space, u, base_w, field, omega, bc_dofs, bc_vals):
    """(alpha nodal, Phi nodal) the way run_optical_subsolvers computes them."""
    alpha = spla.spsolve(mass_matrix(space).tocsc(), alpha_rhs(pa, space, u, base_w, field))
    A, b = msorte_system(space, alpha, omega, bc_dofs, bc_vals)
    return alpha, spla.spsolve(A.tocsc(), b)
