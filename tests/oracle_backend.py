"""TEST INFRASTRUCTURE: a backend with the same interface as
``simudo_b200.fem.backend.CudaBackend`` built from the oracle (C restatement
of the assembly) and scipy's sparse LU.  It lets the GPU-less container run
the full three-stage initialisation and J-V sweeps to (a) test the host
logic (lowering, Newton control flow, steppers, output) and (b) produce the
golden J-V fixtures under tests/golden/.  The product never imports it."""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from oracle import oracle as orc_mod
from oracle.oracle import Oracle


class OracleBackend(object):
    name = 'oracle'

    def __init__(self, pa, n_threads=1):
        self.pa = pa
        self.oracle = Oracle(pa, n_threads=n_threads)
        self.u = None
        self.base_w = None
        self.thmeq = None
        self.phi_opt = None
        self.bc_values = np.zeros(len(pa.bc_dofs))
        self.natbc = np.zeros(len(pa.natbc_cell))
        self.vals = None
        self.b = None
        self.du = None
        self.n_solves = 0

    def T(self, a, dtype=None):
        return np.asarray(a)

    def set_state(self, u, base_w, thmeq, phi_opt):
        self.u = np.array(u, dtype=np.float64)
        self.base_w = None if base_w is None else np.array(base_w, dtype=np.float64)
        self.thmeq = None if thmeq is None else np.array(thmeq, dtype=np.float64)
        self.phi_opt = None if phi_opt is None else np.array(phi_opt, dtype=np.float64)

    def set_u(self, u):
        self.u = np.array(u, dtype=np.float64)

    def set_phi_opt(self, p):
        self.phi_opt = np.array(p, dtype=np.float64)

    def get_u(self):
        return self.u

    def get_base_w(self):
        return self.base_w

    def get_du(self):
        return self.du

    def get_b(self):
        return self.b

    def set_bc_values(self, v):
        self.bc_values = np.asarray(v, dtype=np.float64)

    def set_natbc_values(self, v):
        self.natbc = np.asarray(v, dtype=np.float64)

    def assemble(self):
        self.vals, self.b = self.oracle.assemble(
            self.u, self.base_w, self.thmeq, self.phi_opt, self.bc_values, self.natbc)

    def solve(self):
        A = self.oracle.to_scipy(self.vals).tocsc()
        # equilibrate like the device solver; MUMPS scales too (newton_solver.py:140-142)
        r = 1.0 / np.maximum(np.abs(A).max(axis=1).toarray().ravel(), 1e-300)
        A = sp.diags(r) @ A
        c = 1.0 / np.maximum(np.abs(A).max(axis=0).toarray().ravel(), 1e-300)
        A = (A @ sp.diags(c)).tocsc()
        lu = spla.splu(A)
        x = c * lu.solve(r * self.b)
        # two refinement steps with an extended-precision residual, like the device
        # solver (csrc/band_solver.cu: k_band_residual_dd)
        A0 = self.oracle.to_scipy(self.vals).tocsr()
        data = A0.data.astype(np.longdouble)
        xl = x.astype(np.longdouble)
        bl = self.b.astype(np.longdouble)
        for _ in range(2):
            prod = data * xl[A0.indices]
            res = bl - np.add.reduceat(prod, A0.indptr[:-1])
            xl = xl + c * lu.solve(r * res.astype(np.float64))
        self.du = xl.astype(np.float64)
        self.n_solves += 1
        if not np.isfinite(self.du).all():
            raise RuntimeError('singular Jacobian')

    def newton_update(self, tol, omega, max_du, rel_tol, abs_tol):
        step = self.du if (max_du is None or max_du <= 0) else np.clip(self.du, -max_du, max_du)
        self.u = self.u - omega * step
        m = orc_mod.newton_metrics(self.u, self.du, self.b, np.asarray(tol), rel_tol, abs_tol)
        return {k: (bool(v) if k == 'has_nans' else float(v)) for k, v in m.items()}

    def rebase(self, band, ohmic, thm_avg):
        if ohmic is None:
            self.u, self.base_w = orc_mod.rebase_w(self.pa, self.u, self.base_w, band)
            return
        phi_avg = 0.5 * (self.u[ohmic['dof_a']] + self.u[ohmic['dof_b']]) * ohmic['w']
        avg = thm_avg - np.bincount(ohmic['inv'], weights=phi_avg)
        self.u, self.base_w = orc_mod.rebase_w(self.pa, self.u, self.base_w, band,
                                               ohmic['cells'], avg)

    def optical_update(self, space, specs):
        from oracle import optics_oracle as oo
        self.phi_nodal = getattr(self, 'phi_nodal', {})
        for f, direction, bc_dofs, bc_vals in specs:
            _, phi = oo.solve_field(self.pa, space, self.u, self.base_w, f, direction,
                                    bc_dofs, bc_vals)
            self.phi_opt[f] = phi[space.cell_dofs]
            self.phi_nodal[f] = phi

    def get_phi_opt(self):
        return self.phi_opt

    def surface_flux(self, p, cells, locals_):
        mesh = self.pa.mesh
        facets = mesh.cell_facets[np.asarray(cells), np.asarray(locals_).astype(np.int64)]
        return orc_mod.surface_flux(self.pa, self.u, p, facets)
