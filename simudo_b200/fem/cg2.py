"""CG2 (continuous P2) space on the PDD mesh for the optical sub-problems.

The reference solves one scalar photon-flux field per (energy, direction) on
``CG2`` (``simudo/physics/optical1.py1:46-52``) and projects the absorption
coefficient onto the same space with ``dolfin.project``
(``physics/optical.py:260-274`` -> ``fem/assign.py:12-47``).  This module is
the host-side setup for that path: dof map, fixed CSR pattern, the constant
mass matrix, the x-major permutation for the banded solve, and the reference
tensors the CUDA kernels contract with the cell geometry
(``csrc/optics.cu``).

Dof numbering: vertex ``v`` -> ``v``; facet (edge) ``f`` -> ``Nv + f``.  Local
order = P2 nodal order of ``reference_element.P2_NODES`` (3 vertices, then the
midpoint of the edge opposite vertex i), i.e. exactly the layout of
``sb_state.d_phi_opt`` ``[field][cell][6]``, so ``update_output``
(``optical.py:248-258``) is an index gather.
"""
import ctypes as C

import numpy as np

from .reference_element import (collapsed_rule, p2_lagrange_eval, triangle_default_scheme,
                                gauss_legendre_01, reference_edge_points)

__all__ = ['CG2Space', 'OpticsTables', 'optics_tables', 'p2_gradients']


def p2_gradients(pts):
    """Reference gradients of the P2 nodal basis: dN[6, n, 2]."""
    X, Y = pts[:, 0], pts[:, 1]
    l0 = 1 - X - Y
    dl = np.array([[-1.0, -1.0], [1.0, 0.0], [0.0, 1.0]])        # grad lambda_i
    lam = [l0, X, Y]
    out = np.zeros((6, len(X), 2))
    for i in range(3):
        out[i] = (4 * lam[i] - 1)[:, None] * dl[i][None, :]
    for k, (a, b) in enumerate(((1, 2), (0, 2), (0, 1))):
        out[3 + k] = 4 * (lam[a][:, None] * dl[b][None, :] + lam[b][:, None] * dl[a][None, :])
    return out


N_Q12 = 12


class OpticsTables(C.Structure):
    """Mirror of ``SbOpticsTables`` (``csrc/optics.cu``).

    K[a][b][i][j]  = int d_a N_j d_b N_i            (reference derivatives)
    T[a][i][j][k]  = int N_i d_a (N_k N_j)
    F[f][i][j][k]  = int_0^1 N_i N_j N_k on reference edge f
    q12_*: the degree-6, 12-point default rule used for the projection RHS
    (FFC ``scheme='default'``).
    """
    _fields_ = [('K', (((C.c_double * 6) * 6) * 2) * 2),
                ('T', (((C.c_double * 6) * 6) * 6) * 2),
                ('F', (((C.c_double * 6) * 6) * 6) * 3),
                ('q12_w', C.c_double * N_Q12),
                ('q12_l', (C.c_double * 3) * N_Q12),      # P1 basis (1-X-Y, X, Y)
                ('q12_n', (C.c_double * 6) * N_Q12)]      # P2 basis


def optics_tables():
    t = OpticsTables()
    pts, w = collapsed_rule(6)                 # exact to degree 11
    N = p2_lagrange_eval(pts)                  # [6, n]
    dN = p2_gradients(pts)                     # [6, n, 2]
    K = np.einsum('n,jna,inb->abij', w, dN, dN)
    # d_a (N_k N_j) = dN_k N_j + N_k dN_j
    T = np.einsum('n,in,kna,jn->aijk', w, N, dN, N) + np.einsum('n,in,kn,jna->aijk', w, N, N, dN)
    gt, gw = gauss_legendre_01(4)              # degree 7 >= 6
    F = np.zeros((3, 6, 6, 6))
    for f in range(3):
        Nf = p2_lagrange_eval(reference_edge_points(f, gt))
        F[f] = np.einsum('n,in,jn,kn->ijk', gw, Nf, Nf, Nf)
    qp, qw = triangle_default_scheme(6)

    def put(dst, src):
        flat = np.ascontiguousarray(src, dtype=np.float64).ravel()
        C.memmove(C.addressof(dst), flat.ctypes.data, flat.nbytes)
    put(t.K, K)
    put(t.T, T)
    put(t.F, F)
    put(t.q12_w, qw)
    put(t.q12_l, np.stack([1 - qp[:, 0] - qp[:, 1], qp[:, 0], qp[:, 1]], axis=1))
    put(t.q12_n, p2_lagrange_eval(qp).T)
    return t


class CG2Space(object):
    def __init__(self, mesh):
        mesh.init()
        self.mesh = mesh
        nv, nf, nc = mesh.num_vertices, mesh.num_facets, mesh.num_cells
        self.n_dofs = nv + nf
        cd = np.empty((nc, 6), dtype=np.int32)
        cd[:, :3] = mesh.cells
        cd[:, 3:] = nv + mesh.cell_facets
        self.cell_dofs = np.ascontiguousarray(cd)
        # CSR pattern: union over cells of the 6 x 6 couplings, columns sorted
        rows = np.repeat(cd.astype(np.int64), 6, axis=1).ravel()
        cols = np.tile(cd.astype(np.int64), (1, 6)).ravel()
        key = rows * self.n_dofs + cols
        ukey, inv = np.unique(key, return_inverse=True)
        self.nnz = int(ukey.shape[0])
        urow = ukey // self.n_dofs
        self.colidx = np.ascontiguousarray((ukey % self.n_dofs).astype(np.int32))
        self.rowptr = np.zeros(self.n_dofs + 1, dtype=np.int64)
        np.add.at(self.rowptr, urow + 1, 1)
        self.rowptr = np.cumsum(self.rowptr)
        self.pos = np.ascontiguousarray(inv.reshape(nc, 36).astype(np.int32))   # [c][6 i + j]
        self.diag_pos = np.searchsorted(ukey, np.arange(self.n_dofs, dtype=np.int64)
                                        * (self.n_dofs + 1)).astype(np.int64)

    # -- geometry helpers --------------------------------------------------------------
    def node_coordinates(self):
        m = self.mesh
        return np.concatenate([m.coordinates, m.facet_midpoints()], axis=0)

    def x_major_permutation(self):
        order = np.argsort(self.node_coordinates()[:, 0], kind='stable')
        perm = np.empty(self.n_dofs, dtype=np.int32)
        perm[order] = np.arange(self.n_dofs, dtype=np.int32)
        return perm

    def facet_dofs(self, facets):
        """All CG2 dofs on the given facets (2 vertices + midpoint), unique."""
        m = self.mesh
        facets = np.asarray(facets, dtype=np.int64)
        d = np.concatenate([m.facet_vertices[facets].ravel().astype(np.int64),
                            m.num_vertices + facets])
        return np.unique(d).astype(np.int32)

    # -- constant matrices -------------------------------------------------------------
    def mass_values(self):
        """CSR values of the CG2 mass matrix (the projection's left-hand side)."""
        pts, w = collapsed_rule(4)
        N = p2_lagrange_eval(pts)
        M0 = np.einsum('n,in,jn->ij', w, N, N)
        _, det = self.mesh.cell_jacobians()
        vals = np.zeros(self.nnz)
        np.add.at(vals, self.pos.ravel(),
                  (np.abs(det)[:, None, None] * M0[None]).reshape(-1))
        return vals

    def cell_values(self, nodal):
        """[Nc, 6] P2 nodal values per cell of a CG2 vector (update_output)."""
        return np.asarray(nodal)[self.cell_dofs]
