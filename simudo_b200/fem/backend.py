"""CUDA backend of the Newton hot loop.

Owns the device-resident state of one lowered problem and implements the
calls ``NewtonSolver`` makes per iteration
(``simudo/fem/newton_solver.py:319-326``): assemble (A, b), linear solve,
update + metrics, plus the rebasing hook and the terminal-current functional.
The device copy of the solution is the single source of truth during
``solve()``; host copies are made only when a hook / the stepper asks
("data ownership").  No CPU fallback: constructing this
backend without a CUDA device raises.
"""
import os

import numpy as np
import torch

from .assembly import AssemblyPlan
from .linear_solver import BandSolver, make_solver

__all__ = ['CudaBackend', 'CudaLCNBackend']


def _nvtx(name):
    """NVTX range around a device phase of the Newton loop when SIMUDO_B200_NVTX=1 (for
    `ncu --nvtx --nvtx-include` / nsys timelines); a no-op otherwise."""
    def deco(f):
        if os.environ.get('SIMUDO_B200_NVTX') != '1':
            return f

        def g(*a, **k):
            torch.cuda.nvtx.range_push('simudo_b200:' + name)
            try:
                return f(*a, **k)
            finally:
                torch.cuda.nvtx.range_pop()
        g.__name__, g.__doc__ = f.__name__, f.__doc__
        return g
    return deco


class CudaBackend(object):
    name = 'cuda'

    def __init__(self, pa, device=None):
        self.pa = pa
        self.plan = AssemblyPlan(pa, device=device)
        self.device = self.plan.device
        T = self.plan.tensor
        self.T = T
        self.u = None
        self.base_w = None
        self.thmeq = None
        self.phi_opt = None
        self.bc_values = None
        self.natbc = None
        self.vals = self.plan.new_values()
        self.b = self.plan.new_vector()
        self.du = self.plan.new_vector()
        self._solver = None
        self._metrics = torch.zeros(6, dtype=torch.float64, device=self.device)

    # -- state ----------------------------------------------------------------
    def set_state(self, u, base_w, thmeq, phi_opt):
        T = self.T
        self.u = T(u).clone()
        self.base_w = None if base_w is None else T(base_w).clone()
        self.thmeq = None if thmeq is None else T(thmeq).clone()
        self.phi_opt = None if phi_opt is None else T(phi_opt).clone()

    def set_u(self, u):
        self.u = self.T(u).clone()

    def set_phi_opt(self, phi_opt):
        self.phi_opt = self.T(phi_opt).clone()

    def get_u(self):
        return self.u.cpu().numpy()

    def get_base_w(self):
        return self.base_w.cpu().numpy()

    def get_du(self):
        return self.du.cpu().numpy()

    def get_b(self):
        return self.b.cpu().numpy()

    def set_bc_values(self, v):
        self.bc_values = self.T(v) if len(v) else None

    def set_natbc_values(self, v):
        self.natbc = self.T(self.plan.natbc_device_order(v)) if len(v) else None

    def set_natbc(self, spec):
        """Natural BCs as facet data (``PDDSystem.natbc_spec``): the exterior-facet integrals
        are evaluated on the device (``sb_natbc_eval``); only what changed is uploaded -- in a
        bias sweep the ``c0`` vector."""
        if spec is None:
            self.natbc = None
            return
        T = self.T
        c = getattr(self, '_natbc_dev', None)
        if c is None:
            c = self._natbc_dev = dict(
                cell=T(spec['cell'], torch.int32), local=T(spec['local'], torch.int8),
                slot=T(spec['slot'], torch.int32), c0=None, p2=None, h_c0=None, h_p2=None)
            self.natbc = torch.zeros(len(spec['slot']), dtype=torch.float64, device=self.device)
        dirty = False
        for k in ('c0', 'p2'):
            if c['h_' + k] is None or not np.array_equal(c['h_' + k], spec[k]):
                c['h_' + k] = spec[k].copy()
                c[k] = T(spec[k])
                dirty = True
        if dirty:
            self.plan.natbc_eval(c['cell'], c['local'], c['slot'], c['c0'], c['p2'], self.natbc)

    # -- hot loop ----------------------------------------------------------------
    @_nvtx('assemble')
    def assemble(self):
        self.plan.assemble(self.u, self.base_w, self.thmeq, self.phi_opt, self.bc_values,
                           self.natbc, self.vals, self.b)

    @_nvtx('linear_solve')
    def solve(self):
        if self._solver is None:
            # symbolic phase once per lowered problem, reused by every Newton step
            self._solver = make_solver(self.pa, self.device)
        self._solver.solve(self.vals, self.b, self.du)

    @_nvtx('newton_update')
    def newton_update(self, tol, omega, max_du, rel_tol, abs_tol, log_alpha=None):
        if not isinstance(tol, torch.Tensor):
            tol = self.T(tol)
            self._tol = tol
        out = self.plan.newton_update(self.u, self.du, self.b, tol, omega, max_du, rel_tol,
                                      abs_tol, self._metrics, log_alpha=log_alpha).cpu().numpy()
        return dict(b_norm=float(out[0]), du_norm=float(out[1]),
                    rel_du_norm=float(np.sqrt(out[2])) if out[5] > 0 else float('nan'),
                    combined_du_norm=float(out[3]), has_nans=bool(out[4] > 0))

    @_nvtx('rebase_w')
    def rebase(self, band, ohmic, thm_avg):
        if ohmic is None:
            self.plan.rebase_w(self.u, self.base_w, band)
            return
        cache = ohmic.setdefault('_dev', {})
        if 'cells' not in cache:
            T = self.T
            cache['cells'] = T(ohmic['cells'], torch.int32)
            cache['dof_a'] = T(ohmic['dof_a'], torch.int64)
            cache['dof_b'] = T(ohmic['dof_b'], torch.int64)
            cache['inv'] = T(ohmic['inv'], torch.int64)
            cache['w'] = T(ohmic['w'])
            cache['thm'] = T(thm_avg)
        # contact value e (phi_thmeq - phi), facet average of the P1 potential
        phi_avg = 0.5 * (self.u[cache['dof_a']] + self.u[cache['dof_b']]) * cache['w']
        avg = cache['thm'] - torch.zeros_like(cache['thm']).index_add_(0, cache['inv'], phi_avg)
        self.plan.rebase_w(self.u, self.base_w, band, cache['cells'], avg)

    # -- a10: optical sub-solves (physics/steppers.py:120-136) ------------------------
    @_nvtx('optical_update')
    def optical_update(self, space, specs):
        """``specs``: [(field index, direction, bc_dofs, bc_values)]; solves every
        field for the current PDD state and writes its flux into ``phi_opt``."""
        if getattr(self, '_optics', None) is None:
            from .optics_solver import OpticsPlan
            self._optics = OpticsPlan(self.plan, space)
            self._optics_bc = {}
        T = self.T
        self.phi_nodal = getattr(self, 'phi_nodal', {})
        for f, direction, bc_dofs, bc_vals in specs:
            if f not in self._optics_bc:
                self._optics_bc[f] = T(bc_dofs, torch.int32)
            phi, _ = self._optics.solve_field(f, self.u, self.base_w, direction,
                                              self._optics_bc[f], T(bc_vals),
                                              out=self.phi_opt[f])
            self.phi_nodal[f] = phi

    def get_phi_opt(self):
        return self.phi_opt.cpu().numpy()

    def surface_flux(self, p, cells, locals_):
        T = self.T
        out = self.plan.surface_flux(self.u, p, T(cells, torch.int32), T(locals_, torch.int8),
                                     T(np.ones(len(cells)))).cpu().numpy()
        return float(out[0]), float(out[1])


    def eval_points(self, p, points, vector=False):
        """Sub-function ``p`` of the current solution at ``points`` [n, 2] -> numpy."""
        return self.plan.eval_points(self.u, p, points, vector=vector).cpu().numpy()

    def locate_points(self, points):
        """Index of the containing cell of every point (-1: outside the mesh) -> numpy."""
        _, cells = self.plan.eval_points(self.u, 0, points, return_cells=True)
        return cells.cpu().numpy().astype(np.int64)


class CudaLCNBackend(object):
    """Device state + ``sb_lcn_iteration`` for the local-charge-neutrality stage
    (``LCNSystem``).  No CPU fallback."""
    name = 'cuda'

    def __init__(self, system, device=None):
        import ctypes as C
        from .. import _lib
        self._C, self._lib_mod = C, _lib
        self.lib = _lib.get()
        emul = _lib.is_emulated()
        if device is None:
            # the process' current CUDA device (one process per GPU sets it once)
            device = 'cpu' if emul else 'cuda:%d' % (
                torch.cuda.current_device() if torch.cuda.is_available() else 0)
        self.device = torch.device(device)
        if not emul and (self.device.type != 'cuda' or not torch.cuda.is_available()):
            raise RuntimeError('simudo_b200 needs a CUDA device (no CPU fallback)')
        self.system = system
        self.model = system.model

        def T(a, dtype=torch.float64):
            return torch.as_tensor(np.ascontiguousarray(a), dtype=dtype).to(self.device)
        self.T = T
        self.cell_vertices = T(system.cell_vertices)
        self.cell_tag = T(system.cell_tag, torch.int32)
        self.tag_params = T(system.tag_params)
        nc = system.cell_vertices.shape[0]
        self.n_cells = nc
        self.phi = torch.zeros((nc, 3), dtype=torch.float64, device=self.device)
        self.du = torch.zeros_like(self.phi)
        self.b = torch.zeros_like(self.phi)
        self._out = torch.zeros(6, dtype=torch.float64, device=self.device)
        self._tol = None

    def set_phi(self, phi):
        self.phi = self.T(phi).reshape(self.n_cells, 3).clone()

    def get_phi(self):
        return self.phi.cpu().numpy()

    def get_du(self):
        return self.du.cpu().numpy().ravel()

    def get_b(self):
        return self.b.cpu().numpy().ravel()

    def iterate(self, tol, omega, max_du, rel_tol, abs_tol):
        C = self._C
        if self._tol is None:
            self._tol = self.T(tol)
        p = lambda t: C.c_void_p(t.data_ptr())
        stream = (C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
                  if self.device.type == 'cuda' else None)
        self._lib_mod.check(self.lib.sb_lcn_iteration(
            C.byref(self.model), self.n_cells, p(self.cell_vertices), p(self.cell_tag),
            p(self.tag_params), p(self.phi), p(self._tol), float(omega),
            float(max_du) if max_du else 0.0, float(rel_tol), float(abs_tol),
            p(self.du), p(self.b), p(self._out), stream))
        out = self._out.cpu().numpy()
        return dict(b_norm=float(out[0]), du_norm=float(out[1]),
                    rel_du_norm=float(np.sqrt(out[2])) if out[5] > 0 else float('nan'),
                    combined_du_norm=float(out[3]), has_nans=bool(out[4] > 0))
