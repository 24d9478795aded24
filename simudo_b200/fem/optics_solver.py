"""Device-side optical sub-solver.

One ``OpticsPlan`` per PDD ``AssemblyPlan``: the CG2 space of the mesh, the
constant mass matrix, two banded LU solvers (projection, MSORTE) and the
``sb_optics_*`` entry points.  ``solve_field`` is what the reference does per
field inside ``run_optical_subsolvers`` (``simudo/physics/steppers.py:120-136``):

    update_input   alpha <- project(alpha(PDD state), CG2)   optical.py:260-274
    Newton solve   MSORTE form, linear in Phi                 optical1.py1:27-45
    update_output  Phi -> Phi_pddproj                         optical.py:248-258

The form is linear in Phi, so the reference's two Newton iterations are one
direct solve here (the second iteration only polishes round-off).
"""
import ctypes as C

import numpy as np
import torch

from .. import _lib
from .cg2 import CG2Space, optics_tables
from .linear_solver import BandSolver

__all__ = ['OpticsPlan']


def _p(t):
    return None if t is None else C.c_void_p(t.data_ptr())


class OpticsPlan(object):
    def __init__(self, plan, space=None):
        self.plan = plan
        self.lib = plan.lib
        self.device = plan.device
        self.space = space if space is not None else CG2Space(plan.pa.mesh)
        sp = self.space
        tables = optics_tables()
        if C.sizeof(tables) != self.lib.sb_optics_tables_size():
            raise RuntimeError('SbOpticsTables layout mismatch between Python and C')
        pr = _lib.SbOpticsProblem()
        pr.n_cells, pr.n_dofs, pr.nnz = plan.n_cells, sp.n_dofs, sp.nnz
        keep = [sp.cell_dofs, sp.pos, sp.rowptr, sp.diag_pos]
        pr.h_cell_dofs, pr.h_pos, pr.h_rowptr, pr.h_diag_pos = [
            a.ctypes.data_as(C.c_void_p) for a in keep]
        pr.h_tables = C.cast(C.pointer(tables), C.c_void_p)
        pr.n_tables = C.sizeof(tables)
        h = C.c_void_p()
        _lib.check(self.lib.sb_optics_create(plan._h, C.byref(pr), C.byref(h)))
        self._h = h
        self._keep = (keep, tables)
        csr = (sp.n_dofs, sp.rowptr, sp.colidx, sp.x_major_permutation())
        # one solver per system (mass projection, MSORTE) with fixed argument buffers: a solve
        # whose arguments repeat is replayed as a captured CUDA graph (csrc/band_solver.cu)
        self.band = BandSolver(None, self.device, csr=csr)
        self.band_mass = BandSolver(None, self.device, csr=csr)
        self.mass = plan.tensor(sp.mass_values())
        z = lambda n: torch.zeros(n, dtype=torch.float64, device=self.device)
        self._vals, self._rhs = z(sp.nnz), z(sp.n_dofs)
        self._arhs, self._alpha, self._phi = z(sp.n_dofs), z(sp.n_dofs), z(sp.n_dofs)
        self._mass_factored = False

    def __del__(self):
        h = getattr(self, '_h', None)
        if h:
            try:
                self.lib.sb_optics_destroy(h)
            except Exception:
                pass
            self._h = None

    # -- the three steps ------------------------------------------------------------
    def alpha_rhs(self, field, u, base_w):
        st = _lib.SbState()
        st.d_u, st.d_base_w = _p(u), _p(base_w)
        rhs = self._arhs
        _lib.check(self.lib.sb_optics_alpha_rhs(self._h, C.byref(st), int(field), _p(rhs),
                                                self.plan._stream(None)))
        return rhs

    def project_alpha(self, field, u, base_w):
        """CG2 nodal values of the absorption coefficient of one field."""
        rhs = self.alpha_rhs(field, u, base_w)
        if not self._mass_factored:                 # the mass matrix never changes: factor once
            self._mass_factored = True
            return self.band_mass.solve(self.mass, rhs, x=self._alpha)
        return self.band_mass.solve_factored(rhs, x=self._alpha)

    def assemble(self, alpha, direction, bc_dofs, bc_vals):
        n_bc = 0 if bc_dofs is None else int(bc_dofs.numel())
        _lib.check(self.lib.sb_optics_assemble(
            self._h, _p(alpha), float(direction[0]), float(direction[1]), n_bc,
            _p(bc_dofs) if n_bc else None, _p(bc_vals) if n_bc else None,
            _p(self._vals), _p(self._rhs), self.plan._stream(None)))
        return self._vals, self._rhs

    def to_cells(self, nodal, out=None):
        if out is None:
            out = torch.empty((self.plan.n_cells, 6), dtype=torch.float64, device=self.device)
        _lib.check(self.lib.sb_optics_to_cells(self._h, _p(nodal), _p(out),
                                               self.plan._stream(None)))
        return out

    def solve_field(self, field, u, base_w, direction, bc_dofs, bc_vals, out=None):
        """Returns (Phi nodal [n_dofs], Phi per cell [Nc, 6])."""
        alpha = self.project_alpha(field, u, base_w)
        vals, rhs = self.assemble(alpha, direction, bc_dofs, bc_vals)
        phi = self.band.solve(vals, rhs, x=self._phi).clone()    # the caller keeps it per field
        return phi, self.to_cells(phi, out)
