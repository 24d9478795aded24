"""``AssemblyPlan``: the device-resident replacement of dolfin's
``SystemAssembler(J, F, bcs)`` (``simudo/fem/newton_solver.py:293-304``).

Holds the plan handle created by ``sb_plan_create`` and exposes the hot-path
operations on torch CUDA tensors (torch is only the allocator / stream
provider; every kernel is in ``csrc/``):

``assemble``        a1 + a2  (``NewtonSolution.A`` / ``.b``, newton_solver.py:61-77)
``rebase_w``        a9       (``mixedqfl_do_rebase_w``, poisson_drift_diffusion1.py1:339-383)
``newton_update``   a11      (newton_solver.py:79-122, 306-317, 533-547)
``surface_flux``    terminal current (io/output_writer.py:174-231)
``eval_points``     point evaluation / line cuts (fem/extra_bindings.py:27-75, io/csv.py:54-63)
"""
import ctypes as C

import numpy as np
import torch

from .. import _lib
from .tables import pack_tables

__all__ = ['AssemblyPlan']


def _p(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _hp(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None and a.size else None


class AssemblyPlan(object):
    def __init__(self, pa, device=None):
        self.lib = _lib.get()
        emul = _lib.is_emulated()
        if device is None:
            # the process' current CUDA device (one process per GPU sets it once)
            device = 'cpu' if emul else 'cuda:%d' % (
                torch.cuda.current_device() if torch.cuda.is_available() else 0)
        self.device = torch.device(device)
        if not emul:
            if self.device.type != 'cuda' or not torch.cuda.is_available():
                raise RuntimeError('simudo_b200 assembly needs a CUDA device '
                                   '(no CPU fallback)')
            torch.cuda.set_device(self.device)
        self.pa = pa
        self.N, self.L, self.P = pa.N, pa.L, pa.P
        self.n_cells = pa.n_cells
        self.nnz = pa.csr.nnz
        tables = pack_tables(pa.model)
        if C.sizeof(tables) != self.lib.sb_tables_size():
            raise RuntimeError('SbTables layout mismatch between Python and C')
        pr = _lib.SbProblem()
        pr.n_cells, pr.n_dofs, pr.P = pa.n_cells, pa.N, pa.P
        keep = [pa.cell_vertices, pa.cell_tag, pa.cell_dofs, pa.cell_neighbors,
                pa.cell_jump_mask, pa.csr.rowptr, pa.csr.patterns, pa.csr.cell_pattern,
                pa.tag_params, pa.bc_dofs, pa.natbc_cell, pa.natbc_row]
        (pr.h_cell_vertices, pr.h_cell_tag, pr.h_cell_dofs, pr.h_cell_neighbors,
         pr.h_cell_jump_mask, pr.h_rowptr, pr.h_patterns, pr.h_cell_pattern,
         pr.h_tag_params, pr.h_bc_dofs, pr.h_natbc_cell, pr.h_natbc_row) = \
            [_hp(np.ascontiguousarray(a)) for a in keep]
        pr.n_patterns = pa.csr.patterns.shape[0]
        pr.n_bc = len(pa.bc_dofs)
        pr.n_natbc = len(pa.natbc_cell)
        pr.h_tables = C.cast(C.pointer(tables), C.c_void_p)
        pr.n_tables = C.sizeof(tables) // 8
        handle = C.c_void_p()
        _lib.check(self.lib.sb_plan_create(C.byref(pa.model), C.byref(pr), C.byref(handle)))
        self._h = handle
        self._keep = (keep, tables)
        self.n_bc = pr.n_bc
        self.n_natbc = pr.n_natbc

    def __del__(self):
        h = getattr(self, '_h', None)
        loc = getattr(self, '_locator', None)
        if loc:
            try:
                self.lib.sb_locator_destroy(loc)
            except Exception:
                pass
            self._locator = None
        if h:
            try:
                self.lib.sb_plan_destroy(h)
            except Exception:
                pass
            self._h = None

    # -- buffers ---------------------------------------------------------------
    def tensor(self, a, dtype=torch.float64):
        if a is None:
            return None
        if isinstance(a, torch.Tensor):
            return a.to(device=self.device, dtype=dtype).contiguous()
        return torch.as_tensor(np.ascontiguousarray(a), dtype=dtype).to(self.device)

    def new_values(self):
        """CSR value array with the structural zeros of the dolfin pattern."""
        return torch.zeros(self.nnz, dtype=torch.float64, device=self.device)

    def new_vector(self):
        return torch.zeros(self.N, dtype=torch.float64, device=self.device)

    def natbc_device_order(self, values):
        """Reorder user-order natural-BC values to the plan's cell-sorted order."""
        return np.asarray(values, dtype=np.float64)[self.pa.natbc_order]

    def _stream(self, stream):
        if stream is not None:
            return C.c_void_p(stream.cuda_stream)
        if self.device.type == 'cuda':
            return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        return None

    # -- hot path ---------------------------------------------------------------
    def assemble(self, u, base_w=None, thmeq_phi=None, phi_opt=None, bc_values=None,
                 natbc_values=None, vals=None, b=None, stream=None):
        """Enqueue one residual+Jacobian assembly.  All arguments are device
        tensors (float64, contiguous); ``vals`` / ``b`` are written in place."""
        st = _lib.SbState()
        st.d_u = _p(u)
        st.d_base_w = _p(base_w)
        st.d_thmeq_phi = _p(thmeq_phi)
        st.d_phi_opt = _p(phi_opt)
        st.d_bc_values = _p(bc_values)
        st.d_natbc_values = _p(natbc_values)
        _lib.check(self.lib.sb_assemble(self._h, C.byref(st), _p(vals), _p(b),
                                        self._stream(stream)))
        return vals, b

    def column_indices(self, stream=None):
        col = torch.empty(self.nnz, dtype=torch.int32, device=self.device)
        _lib.check(self.lib.sb_plan_column_indices(self._h, _p(col), self._stream(stream)))
        return col

    def natbc_eval(self, cell, local, slot, c0, p2, out, stream=None):
        """a8: exterior-facet integrals of the natural BCs into ``out`` (the
        ``natbc_values`` argument of ``assemble``); see ``sb_natbc_eval``."""
        from .reference_element import reference_tables
        if getattr(self, '_nat_tables', None) is None:
            T = reference_tables()
            tab = np.concatenate([np.asarray(T.facet_gl_w, dtype=np.float64).ravel(),
                                  np.transpose(T.facet_p2, (0, 2, 1)).ravel(),       # [l][q][a]
                                  np.stack([T.facet_bdm_trace[l, 3 * l:3 * l + 3, :]
                                            for l in range(3)]).ravel()])             # [l][k][q]
            self._nat_tables = np.ascontiguousarray(tab, dtype=np.float64)
        tab = self._nat_tables
        _lib.check(self.lib.sb_natbc_eval(
            self._h, int(cell.numel()), _p(cell), _p(local), _p(slot), _p(c0), _p(p2),
            tab.ctypes.data_as(C.c_void_p), len(tab), _p(out), self._stream(stream)))
        return out

    def rebase_w(self, u, base_w, band, bc_cells=None, bc_avg=None, stream=None):
        n = 0 if bc_cells is None else int(bc_cells.numel())
        _lib.check(self.lib.sb_rebase_w(self._h, _p(u), _p(base_w), band, n,
                                        _p(bc_cells), _p(bc_avg), self._stream(stream)))

    def newton_update(self, u, du, b, tol, omega, max_du, rel_tol, abs_tol, out=None,
                      stream=None, log_alpha=None):
        if out is None:
            out = torch.zeros(6, dtype=torch.float64, device=self.device)
        _lib.check(self.lib.sb_newton_update_damped(
            self._h, _p(u), _p(du), _p(b), _p(tol), float(omega),
            float(max_du if max_du is not None else -1.0),
            float(log_alpha if log_alpha is not None else 0.0), float(rel_tol),
            float(abs_tol), _p(out), self._stream(stream)))
        return out

    def surface_flux(self, u, p, facet_cell, facet_local, facet_sign, out=None,
                     stream=None):
        if out is None:
            out = torch.zeros(2, dtype=torch.float64, device=self.device)
        _lib.check(self.lib.sb_surface_flux(
            self._h, _p(u), p, int(facet_cell.numel()), _p(facet_cell),
            _p(facet_local), _p(facet_sign), _p(out), self._stream(stream)))
        return out

    def eval_points(self, u, sub_problem, points, vector=False, out=None, stream=None,
                    return_cells=False):
        """Values of a sub-function at ``points`` [n, 2] (device tensor or array):
        ``vector=False`` the DG1 scalar of ``sub_problem`` -> [n]; ``vector=True`` its BDM2
        vector -> [n, 2].  Points outside the mesh keep the value of ``out`` (NaN when ``out``
        is not given) -- ``Function_eval_many`` skips them (extra_bindings.py:62-64)."""
        if getattr(self, '_locator', None) is None:
            from .reference_element import _dubiner_p2_monomial_matrix
            mono = np.ascontiguousarray(_dubiner_p2_monomial_matrix(), dtype=np.float64)
            cv = np.ascontiguousarray(self.pa.cell_vertices, dtype=np.float64)
            h = C.c_void_p()
            _lib.check(self.lib.sb_locator_create(self._h, _hp(cv), 0, 0, _hp(mono), C.byref(h)))
            self._locator = h
        xy = self.tensor(points).reshape(-1, 2)
        n = xy.shape[0]
        if out is None:
            out = torch.full((n, 2) if vector else (n,), float('nan'), dtype=torch.float64,
                             device=self.device)
        cells = torch.empty(n, dtype=torch.int32, device=self.device) if return_cells else None
        _lib.check(self.lib.sb_eval_points(self._locator, _p(u), int(sub_problem),
                                           1 if vector else 0, n, _p(xy), _p(out), _p(cells),
                                           self._stream(stream)))
        return (out, cells) if return_cells else out

    def set_profiling(self, on=True):
        _lib.check(self.lib.sb_plan_set_profiling(self._h, 1 if on else 0))

    def last_kernel_ms(self):
        ms = C.c_double(0.0)
        _lib.check(self.lib.sb_plan_last_kernel_ms(self._h, C.byref(ms)))
        return ms.value

    def last_stage_ms(self):
        """{stage: ms} of the last pass: moments, generation, rows_fast, rows_special."""
        ms = (C.c_double * 4)()
        _lib.check(self.lib.sb_plan_last_stage_ms(self._h, ms))
        return dict(zip(('k_moments', 'k_generation', 'k_rows_fast', 'k_rows_special'),
                        [float(x) for x in ms]))

    def is_native(self):
        """True when the plan runs the closed-form kernels (csrc/native.inl)."""
        return bool(self.lib.sb_plan_is_native(self._h))

    def generation_signature(self):
        """0: the process list is interpreted; 1 / 2: compiled-in list (IB cell / pn + SRH)."""
        return int(self.lib.sb_plan_generation_signature(self._h))

    def launch_count(self):
        return int(self.lib.sb_launch_count())
