"""Lowered problem description: the numpy arrays behind ``sb_problem``.

This is the data the reference hands to dolfin implicitly through
``SystemAssembler(J, F, bcs)`` (``simudo/fem/newton_solver.py:293-304``):
mesh geometry, dof maps, sparsity, subdomain tags and parameters, Dirichlet
dofs, and the exterior-facet natural terms.  It is plain input data -- the
CUDA plan (``simudo_b200.fem.assembly``) and the CPU oracle (``oracle/``)
both consume it, neither depends on the other.
"""
import numpy as np

from . import model as M
from .dofmap import build_cell_dofs, CsrStructure, local_layout

__all__ = ['ProblemArrays']


class ProblemArrays(object):
    def __init__(self, mesh, cell_tag, model, tag_params, numbering='b200',
                 jump_facet_mask=None, bc_dofs=None, natbc=None, pattern='structural'):
        """
        mesh:        simudo_b200.mesh.Mesh
        cell_tag:    [Nc] int row index into tag_params
        model:       SbModel
        tag_params:  [n_tags, TAG_STRIDE]
        jump_facet_mask: [Nf] uint8, bit b set = base_w jump term of band b
                     applies on that (interior) facet
        bc_dofs:     constrained global dofs
        natbc:       (cells, local_rows) of cell-local residual additions
        """
        mesh.init()
        self.mesh = mesh
        self.model = model
        self.P = M.n_subproblems(model)
        self.L = 15 * self.P
        self.numbering = numbering
        self.cell_tag = np.ascontiguousarray(cell_tag, dtype=np.int32)
        self.tag_params = np.ascontiguousarray(tag_params, dtype=np.float64)
        assert self.tag_params.shape == (model.n_tags, M.TAG_STRIDE)
        self.cell_vertices = np.ascontiguousarray(
            mesh.coordinates[mesh.cells], dtype=np.float64)        # [Nc,3,2]
        self.cell_dofs, self.N = build_cell_dofs(mesh, self.P, numbering)
        self.pattern = pattern
        self.csr = CsrStructure(mesh, self.cell_dofs, self.P, self.N, pattern=pattern)
        self.cell_neighbors = np.ascontiguousarray(
            mesh.cell_to_cells_adjacency_via_facet(), dtype=np.int32)
        nc = mesh.num_cells
        jm = np.zeros((nc, 3), dtype=np.uint8)
        if jump_facet_mask is not None:
            fmask = np.asarray(jump_facet_mask, dtype=np.uint8)
            fc, fl = mesh.facet_cells, mesh.facet_local
            interior = fc[:, 1] >= 0
            plus = fc[interior, 0]       # '+' = first (lower-numbered) cell
            jm[plus, fl[interior, 0].astype(np.int64)] = fmask[interior]
        self.cell_jump_mask = jm
        self.set_bc_dofs(bc_dofs if bc_dofs is not None else [])
        if natbc is None:
            natbc = (np.zeros(0, np.int32), np.zeros(0, np.int32))
        cells, rows = (np.asarray(a, dtype=np.int32) for a in natbc)
        o = np.argsort(cells, kind='stable')
        self.natbc_order = o
        self.natbc_cell = np.ascontiguousarray(cells[o])
        self.natbc_row = np.ascontiguousarray(rows[o])

    def set_bc_dofs(self, dofs):
        dofs = np.unique(np.asarray(dofs, dtype=np.int32))
        self.bc_dofs = np.ascontiguousarray(dofs)

    @property
    def n_cells(self):
        return self.mesh.num_cells

    @property
    def n_bands(self):
        return self.model.n_bands

    def layout(self):
        return local_layout(self.P)

    # -- helpers used by host code and tests -----------------------------------
    def facet_dofs(self, p, facets):
        """Global dofs (3 per facet) of sub-problem p's BDM space on facets."""
        mesh = self.mesh
        lay = self.layout()
        c = mesh.facet_cells[facets, 0]
        l = mesh.facet_local[facets, 0].astype(np.int64)
        loc = lay['facet'][p][l]                     # [n, 3]
        return self.cell_dofs[c[:, None], loc]

    def subfunction_dofs(self, p, kind):
        """All global dofs of sub-problem p's 'dg' or 'bdm' component."""
        lay = self.layout()
        return np.unique(self.cell_dofs[:, lay[kind][p]])

    def tolerance_vector(self, tol_dg, tol_bdm):
        """absolute_tolerance_vector (``function_space.py:197-218``): per
        sub-problem trial tolerances assigned dof-wise."""
        t = np.zeros(self.N)
        lay = self.layout()
        for p in range(self.P):
            t[self.cell_dofs[:, lay['dg'][p]].ravel()] = tol_dg[p]
            t[self.cell_dofs[:, lay['bdm'][p]].ravel()] = tol_bdm[p]
        return t
