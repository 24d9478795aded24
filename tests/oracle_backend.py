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

    def eval_points(self, p, points, vector=False):
        from oracle import oracle as _orc
        return _orc.eval_points(self.pa, self.u, p, points, vector=vector)[0]

    def locate_points(self, points):
        from oracle import oracle as _orc
        return _orc.eval_points(self.pa, self.u, 0, points)[1]

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

    def set_natbc(self, spec):
        """Natural BCs as facet data: integrated by the oracle's own exterior-facet
        restatement (``oracle.natural_bc_integrals``), not handed over as numbers."""
        if spec is None:
            self.natbc = np.zeros(0)
            return
        self.natbc = orc_mod.natural_bc_integrals(self.pa, spec['cell'], spec['local'],
                                                  spec['c0'], spec['p2']).ravel()

    def assemble(self):
        self.vals, self.b = self.oracle.assemble(
            self.u, self.base_w, self.thmeq, self.phi_opt, self.bc_values, self.natbc)

    def solve(self):
        A = self.oracle.to_scipy(self.vals).tocsr()
        # equilibrate like the device solvers (Ruiz sweeps); MUMPS scales too (MC64,
        # newton_solver.py:140-142)
        n = A.shape[0]
        r, c = np.ones(n), np.ones(n)
        for _ in range(8):
            rm = np.abs(A).max(axis=1).toarray().ravel()
            cm = np.abs(A).max(axis=0).toarray().ravel()
            rr = 1.0 / np.sqrt(np.where(rm > 0, rm, 1.0))
            cc = 1.0 / np.sqrt(np.where(cm > 0, cm, 1.0))
            A = (sp.diags(rr) @ A @ sp.diags(cc)).tocsr()
            r *= rr
            c *= cc
        A = A.tocsc()
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

    def newton_update(self, tol, omega, max_du, rel_tol, abs_tol, log_alpha=None):
        step = self.du if (max_du is None or max_du <= 0) else np.clip(self.du, -max_du, max_du)
        if log_alpha:
            step = np.sign(step) * np.log1p(np.abs(step) * log_alpha) / log_alpha
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


class OracleLCNBackend(object):
    """Numpy restatement of the local-charge-neutrality Newton iteration
    (``poisson_drift_diffusion1.py1:89-125``; solver ``poisson_drift_diffusion.py:
    318-331``): vectorised over cells, generic 3x3 solves, metrics from
    ``oracle.newton_metrics``.  Test infrastructure only."""
    name = 'oracle'

    def __init__(self, system):
        from simudo_b200.fem import model as M
        from simudo_b200.fem.reference_element import triangle_default_scheme, p1_eval
        self.M = M
        self.system = system
        pts, w = triangle_default_scheme(4)
        self.qw = w
        self.ql = p1_eval(pts).T                    # [q, 3]
        self.phi = None
        self.du = self.b = None

    def set_phi(self, phi):
        self.phi = np.array(phi, dtype=np.float64).reshape(-1, 3)

    def get_phi(self):
        return self.phi

    def get_du(self):
        return self.du.ravel()

    def get_b(self):
        return self.b.ravel()

    def residual_jacobian(self):
        M, sysm = self.M, self.system
        tp = sysm.tag_params[sysm.cell_tag]          # [Nc, stride]
        kT = tp[:, M.TP_KT][:, None]
        phiq = self.phi @ self.ql.T                  # [Nc, q]
        rho = np.repeat(tp[:, M.TP_RHO_STATIC][:, None], phiq.shape[1], axis=1)
        drho = np.zeros_like(rho)
        for b, band in enumerate(sysm.bands):
            o = M.TP_BAND0 + 5 * b
            N, E = tp[:, o + M.TPB_N][:, None], tp[:, o + M.TPB_E][:, None]
            s = band.sign
            if band.kind == M.BAND_NONDEGENERATE:
                u = N * np.exp(s * E / kT) * np.exp(-s * phiq / kT)
                du = -s / kT * u
            else:
                e = np.exp(s * (phiq - E) / kT)
                f = 1.0 / (1.0 + e)
                u = N * f
                du = -s / kT * N * f * (e * f)
            rho += s * u
            drho += s * du
        eps = tp[:, M.TP_EPS][:, None]
        _, det = sysm.mesh.cell_jacobians()
        wq = self.qw[None, :] * np.abs(det)[:, None]
        F = np.einsum('cq,qi->ci', wq * rho / eps, self.ql)
        J = np.einsum('cq,qi,qj->cij', wq * drho / eps, self.ql, self.ql)
        return F, J

    def iterate(self, tol, omega, max_du, rel_tol, abs_tol):
        F, J = self.residual_jacobian()
        du = np.linalg.solve(J, F[:, :, None])[:, :, 0]
        step = du if not max_du else np.clip(du, -max_du, max_du)
        self.phi = self.phi - omega * step
        self.du, self.b = du, F
        m = orc_mod.newton_metrics(self.phi.ravel(), du.ravel(), F.ravel(),
                                   np.asarray(tol).ravel(), rel_tol, abs_tol)
        return {k: (bool(v) if k == 'has_nans' else float(v)) for k, v in m.items()}


OracleBackend.lcn_backend = OracleLCNBackend
