"""Lowering of a declared ``PoissonDriftDiffusion`` problem to kernel tables.

What UFL/FFC/dolfin do implicitly for the reference
(``NewtonSolver.from_nice_obj`` -> ``SystemAssembler(J, F, bcs)``,
``simudo/fem/newton_solver.py:266-304``) happens here explicitly and once:

* bands / processes  -> ``SbModel`` + per-tag parameter rows
  (``poisson_drift_diffusion1.py1:170-245``, ``electro_optical_process.py``)
* unit folding       -> "form units" um, V, eV, A, s (``delayed_form.py:80-94``)
* essential BCs      -> constrained BDM facet dofs (``spatial.py:275-306``)
* natural BCs        -> exterior-facet integrals added to cell residuals
  (``spatial.py:237-273``; ``poisson_drift_diffusion1.py1:47,270``)
* Ohmic contacts     -> BC-adjacent cells of the qfl rebasing
  (``spatial.py:139-235``, ``poisson_drift_diffusion1.py1:362-367``)
* interior facets    -> base_w jump-term mask (``poisson_drift_diffusion.py:709-713``)

``PDDSystem`` then owns the solution state and forwards the hot-path calls to a
backend (CUDA by default; tests inject an oracle-based one).
"""
import numpy as np

from ..fem import model as M
from ..fem.expr import Expr, ThermalEquilibriumU, as_expr, p2_from_p1, vector_components
from ..fem.problem_arrays import ProblemArrays
from ..fem.reference_element import reference_tables, gauss_legendre_01
from .optical import AbsorptionRangesHelper

__all__ = ['PDDSystem', 'LCNSystem', 'strandberg_I']

_REF_OUT = np.array([1.0, -1.0, 1.0])            # outward sign of the reference scaled normals
_REF_EDGES = [(1, 2), (0, 2), (0, 1)]            # local facet i opposite vertex i


def strandberg_I(E_L, E_U, kT, U):
    """Strandberg radiative integral K * int_{E_L}^{E_U} E^2 exp(-E/kT) dE in
    1/(mesh_unit^2 s) (``electro_optical_process.py:324-353``), closed form."""
    h = U('1 planck_constant').m_as('eV*s')
    c = U('1 speed_of_light').m_as('mesh_unit/s')
    K = 8.0 * np.pi / (h ** 3 * c ** 2)
    F = lambda E: -kT * (2 * kT ** 2 + 2 * E * kT + E ** 2) * np.exp(-E / kT)
    sign = 1.0 if E_U > E_L else -1.0
    return K * (F(E_U) - F(E_L)) * sign


def _band_u(kind, s, N, E, kT, phiqfl):
    if kind == M.BAND_NONDEGENERATE:
        return N * np.exp(s * E / kT) * np.exp(-s * phiqfl / kT)
    return N / (1.0 + np.exp(s * (phiqfl - E) / kT))


class _SystemBase(object):
    def __init__(self, pdd):
        self.pdd = pdd
        self.U = pdd.unit_registry
        self.mesh_data = pdd.mesh_data
        self.mesh = pdd.mesh_data.mesh.init()
        cf = np.asarray(self.mesh_data.cell_function)
        self.cvs = sorted(set(np.unique(cf).tolist()))
        self.cv_to_tag = {cv: i for i, cv in enumerate(self.cvs)}
        self.cell_tag = np.array([self.cv_to_tag[v] for v in cf.tolist()], dtype=np.int32) \
            if cf.size < 200000 else np.searchsorted(np.array(self.cvs), cf).astype(np.int32)
        self.bands = list(pdd.bands)

    # -- parameter tables ------------------------------------------------------------
    def _band_rows(self, tp):
        pdd, U = self.pdd, self.U
        sp = pdd.spatial
        kT = pdd.kT
        eps = sp.get_table('poisson/permittivity', 'elementary_charge/V/mesh_unit', required=True)
        rho = sp.get_table('poisson/static_rho', 'elementary_charge/mesh_unit^3')
        for cv, t in self.cv_to_tag.items():
            tp[t, M.TP_KT] = kT[cv]
            tp[t, M.TP_EPS] = eps[cv]
            tp[t, M.TP_RHO_STATIC] = rho[cv]
        self.band_E = []
        self.band_N = []
        for b, band in enumerate(self.bands):
            o = M.TP_BAND0 + 5 * b
            inside = pdd.mesh_data.evaluate_topology(band.subdomain)
            N = sp.get_table(band.spatial_prefix + band.density_key, '1/mesh_unit^3')
            E = sp.get_table(band.spatial_prefix + 'energy_level', 'eV')
            mkey = band.spatial_prefix + band.mobility_key
            if band.kind == M.BAND_INTERMEDIATE and not sp.has_rule(mkey):
                mkey = band.spatial_prefix + 'empty_band_mobility'     # deprecated alias
            mob = sp.get_table(mkey, 'mesh_unit^2/V/s')
            for cv, t in self.cv_to_tag.items():
                tp[t, o + M.TPB_IN_SUBDOMAIN] = 1.0 if cv in inside else 0.0
                tp[t, o + M.TPB_N] = N[cv]
                tp[t, o + M.TPB_E] = E[cv]
                tp[t, o + M.TPB_MOBILITY] = mob[cv]
            self.band_E.append(E)
            self.band_N.append(N)

    def band_model_list(self):
        return [dict(kind=b.kind, sign=b.sign,
                     const_mobility=bool(getattr(b, 'use_constant_mobility', False)))
                for b in self.bands]


class PDDSystem(_SystemBase):
    """goal 'thermal equilibrium' (Poisson unknowns only, bands at qfl = 0) or
    'full' (Poisson + one MixedQfl pair per band)."""

    def __init__(self, pdd, numbering='b200', pattern='native'):
        super().__init__(pdd)
        goal = pdd.problem_data.goal
        self.full = (goal == 'full')
        procs = list(pdd.electro_optical_processes) if self.full else []
        self.procs = procs
        fields = list(pdd.problem_data.optical.fields) if self.full else []
        self.fields = fields
        bidx = {id(b): i for i, b in enumerate(self.bands)}
        plist = []
        for p in procs:
            plist.append(dict(
                kind=p.kind, dst=bidx[id(p.dst_band)], src=bidx[id(p.src_band)],
                trap=bidx[id(p.trap_band)] if getattr(p, 'trap_band', None) is not None else -1,
                radiative=getattr(p, 'enable_radiative_recombination', True)))
        qs = dict(quad_degree_rho=pdd.poisson.mixed_debug_quad_degree_rho)
        if self.full and self.bands:
            qs['quad_degree_super'] = max(b.mixedqfl_debug_quad_degree_super for b in self.bands)
            qs['quad_degree_g'] = max(b.mixedqfl_debug_quad_degree_g for b in self.bands)
        self.model = M.make_model(self.band_model_list(), plist, n_fields=len(fields),
                                  n_tags=len(self.cvs), bands_have_dofs=self.full, **qs)
        tp = M.new_tag_params(len(self.cvs))
        self._band_rows(tp)
        self._proc_rows(tp, plist)
        self.tag_params = tp

        mesh = self.mesh
        jmask = np.zeros(mesh.num_facets, dtype=np.uint8)
        if self.full:
            ff = np.asarray(self.mesh_data.facet_function)
            for b, band in enumerate(self.bands):
                fvs = {fv for fv, _ in self.mesh_data.evaluate_topology(
                    band.subdomain.internal_facets())}
                jmask[np.isin(ff, list(fvs))] |= np.uint8(1 << b)
        self.pa = ProblemArrays(mesh, self.cell_tag, self.model, tp, numbering=numbering,
                                jump_facet_mask=jmask, pattern=pattern)
        self._setup_bcs()
        nc = mesh.num_cells
        self._u_host = np.zeros(self.pa.N)
        self.thmeq_phi = np.zeros((nc, 6))
        self._base_w0 = np.zeros((len(self.bands), nc))
        self._phi_opt_host = np.zeros((len(fields), nc, 6)) if fields else None
        self._phi_opt_stale = False      # True: the backend holds a newer flux
        self._backend = None
        self._dirty_host = False

    # -- processes -----------------------------------------------------------------
    def _proc_rows(self, tp, plist):
        sp, U = self.pdd.spatial, self.U
        kT = self.pdd.kT
        for pi, (p, d) in enumerate(zip(self.procs, plist)):
            o = M.TP_PROC0 + 8 * pi
            dst, src = self.bands[d['dst']], self.bands[d['src']]
            if p.kind == M.PROC_SRH:
                tau_d = sp.get_table('/'.join((p.name, dst.name, 'tau')), 's')
                tau_s = sp.get_table('/'.join((p.name, src.name, 'tau')), 's')
                Et = sp.get_table('/'.join((p.name, 'energy_level')), 'eV')
                for cv, t in self.cv_to_tag.items():
                    tp[t, o + M.TPP_P0] = tau_d[cv]
                    tp[t, o + M.TPP_P1] = tau_s[cv]
                    for slot, bi in ((M.TPP_P2, d['dst']), (M.TPP_P3, d['src'])):
                        b = self.bands[bi]
                        tp[t, o + slot] = _band_u(b.kind, b.sign, self.band_N[bi][cv],
                                                  self.band_E[bi][cv], kT[cv], Et[cv])
            elif p.kind == M.PROC_SR_TRAP:
                sig = sp.get_table(p.name + '/sigma_th', 'mesh_unit^2')
                vth = sp.get_table(p.name + '/v_th', 'mesh_unit/s')
                for cv, t in self.cv_to_tag.items():
                    tp[t, o + M.TPP_P0] = sig[cv] * vth[cv]
            elif p.kind in (M.PROC_RAD_BB, M.PROC_RAD_IB):
                key = p.name + ('/alpha' if p.kind == M.PROC_RAD_BB else '/sigma_opt')
                unit = '1/mesh_unit' if p.kind == M.PROC_RAD_BB else 'mesh_unit^2'
                coef = sp.get_table(key, unit)
                for cv, t in self.cv_to_tag.items():
                    levels = {b.name: self.band_E[i][cv] for i, b in enumerate(self.bands)}
                    bounds = AbsorptionRangesHelper.transition_bounds(levels)
                    E_L, E_U = bounds[frozenset((src.name, dst.name))]
                    I = strandberg_I(E_L, E_U, kT[cv], U)
                    if p.kind == M.PROC_RAD_BB:
                        tp[t, o + M.TPP_P0] = coef[cv] * I
                    else:
                        IB = self.bands[d['trap']]
                        RB = src if dst is IB else dst
                        ri = self.bands.index(RB)
                        E_I = self.band_E[d['trap']][cv]
                        if IB.kind == M.BAND_INTERMEDIATE and self.pdd.spatial.has_rule(
                                IB.spatial_prefix + 'degeneracy'):
                            deg = sp.get_table(IB.spatial_prefix + 'degeneracy', default=1.0)[cv]
                            if deg > 0:
                                E_I = E_I - IB.sign * kT[cv] * np.log(deg)   # DegeneracyMixin
                        u1 = _band_u(RB.kind, RB.sign, self.band_N[ri][cv],
                                     self.band_E[ri][cv], kT[cv], E_I)
                        tp[t, o + M.TPP_P0] = coef[cv] * I / u1 if u1 > 0 else 0.0
                    for f, fld in enumerate(self.fields):
                        Eph = fld.photon_energy.m_as('eV')
                        tp[t, o + M.TPP_FIELD0 + f] = coef[cv] if (E_L <= Eph < E_U) else 0.0

    # -- boundary conditions ----------------------------------------------------------
    def _setup_bcs(self):
        pa, mesh, sp = self.pa, self.mesh, self.pdd.spatial
        lay = pa.layout()
        subs = ['poisson'] + ([b.name for b in self.bands] if self.full else [])
        flux_key = lambda p: subs[p] + ('/E' if p == 0 else '/j')
        nat_key = lambda p: subs[p] + ('/phi' if p == 0 else '/u')
        bc_dofs, bc_vals = [], []
        for p in range(pa.P):
            sp.check_no_overlap(flux_key(p), nat_key(p))
            for bc, facets in sp.bc_facets(flux_key(p)):
                if len(facets) == 0:
                    continue
                d = pa.facet_dofs(p, facets)                       # [n, 3]
                val = bc.value
                # value of the BDM2 dofs (DirichletBC on the flux sub-space, spatial.py:296-304):
                # the dof functional is v(x_k) . n with n the edge tangent (lower -> higher
                # vertex) rotated clockwise, length-scaled; for a constant
                # vector field all three dofs of a facet are vx t_y - vy t_x
                unit = 'V/mesh_unit' if p == 0 else 'A/mesh_unit^2'
                mag = val.m_as(unit) if hasattr(val, 'm_as') else val
                vx, vy = vector_components(mag)
                xf = mesh.coordinates[mesh.facet_vertices[facets]]  # [n, 2 vertices, 2]
                t = xf[:, 1] - xf[:, 0]
                g = vx * t[:, 1] - vy * t[:, 0]
                bc_dofs.append(d.ravel())
                bc_vals.append(np.repeat(g, 3))
        if bc_dofs:
            dofs = np.concatenate(bc_dofs)
            vals = np.concatenate(bc_vals)
            ud, first = np.unique(dofs, return_index=True)
            pa.set_bc_dofs(ud)
            self.bc_values = vals[first]
        else:
            self.bc_values = np.zeros(0)
        # natural BC of the potential: + oint phi_bc psi.n on tagged exterior facets
        self._nat = []
        cells, rows = [], []
        for bc, facets in sp.bc_facets('poisson/phi', exterior_only=True):
            if len(facets) == 0:
                continue
            c = mesh.facet_cells[facets, 0].astype(np.int64)
            l = mesh.facet_local[facets, 0].astype(np.int64)
            cells.append(np.repeat(c, 3))
            rows.append(lay['facet'][0][l].ravel())
            self._nat.append((bc, facets, c, l))
        if cells:
            cells = np.concatenate(cells).astype(np.int32)
            rows = np.concatenate(rows).astype(np.int32)
            o = np.argsort(cells, kind='stable')
            pa.natbc_order = o
            pa.natbc_cell = np.ascontiguousarray(cells[o])
            pa.natbc_row = np.ascontiguousarray(rows[o])
        # Ohmic contacts: cells adjacent to '<band>/u' facets (rebasing)
        self.ohmic = []
        if self.full:
            for b, band in enumerate(self.bands):
                entry = None
                groups = sp.bc_facets(band.name + '/u', exterior_only=True)
                facets = np.concatenate([f for _, f in groups]) if groups else np.zeros(0, int)
                for bc, _ in groups:
                    mag = bc.value.magnitude if hasattr(bc.value, 'magnitude') else bc.value
                    leafs = [l for _, l in as_expr(mag).terms] if isinstance(mag, Expr) else []
                    if not any(isinstance(l, ThermalEquilibriumU) for l in leafs):
                        raise NotImplementedError(
                            "'<band>/u' contact values other than band.thermal_equilibrium_u "
                            "are not supported yet")
                if len(facets):
                    entry = self._ohmic_tables(facets)
                self.ohmic.append(entry)

    def _ohmic_tables(self, facets):
        """For the DG0 'cell average of the contact value' solve
        (``spatial.py:185-235``): per BC-adjacent cell, length-weighted facet
        averages.  phi is P1 on a facet (mean of its two vertex values), the
        thermal-equilibrium potential P2 (Simpson)."""
        mesh, pa = self.mesh, self.pa
        lay = pa.layout()
        c = mesh.facet_cells[facets, 0].astype(np.int64)
        l = mesh.facet_local[facets, 0].astype(np.int64)
        length = mesh.facet_lengths()[facets]
        cells, inv = np.unique(c, return_inverse=True)
        wsum = np.bincount(inv, weights=length)
        w = length / wsum[inv]
        va = np.array([_REF_EDGES[f][0] for f in l])
        vb = np.array([_REF_EDGES[f][1] for f in l])
        return dict(cells=cells.astype(np.int32), inv=inv, w=w, c=c, va=va, vb=vb, l=l,
                    dof_a=pa.cell_dofs[c, lay['dg'][0][va]], dof_b=pa.cell_dofs[c, lay['dg'][0][vb]])

    def natbc_spec(self):
        """The natural boundary conditions as data for the exterior-facet kernel
        (``sb_natbc_eval``; ``fem/spatial.py:237-273``): per 'poisson/phi' facet its cell and
        local facet, the slots of its three integrals in the plan's natural-BC array, the
        constant part ``c0`` of the BC value (floats and ``Constant`` leaves: the swept bias)
        and the P2 nodal values ``p2`` of its cell-function part on the facet's cell.  The
        integrals themselves are done by the backend (device kernel / oracle)."""
        if not self._nat:
            return None
        if getattr(self, '_nat_static', None) is None:
            inv = np.empty(len(self.pa.natbc_order), dtype=np.int32)
            inv[self.pa.natbc_order] = np.arange(len(inv), dtype=np.int32)
            cells = np.concatenate([c for _, _, c, _ in self._nat]).astype(np.int32)
            local = np.concatenate([l for _, _, _, l in self._nat]).astype(np.int8)
            self._nat_static = (cells, local, np.ascontiguousarray(inv))
        cells, local, slot = self._nat_static
        c0, p2 = [], []
        for bc, facets, c, l in self._nat:
            val = bc.value
            mag = as_expr(val.m_as('V') if hasattr(val, 'm_as') else val)
            k0 = mag.offset
            acc = np.zeros((len(c), 6))
            for a, leaf in mag.terms:
                if hasattr(leaf, 'assign'):                       # Constant
                    k0 += a * float(leaf)
                elif hasattr(leaf, 'values') and not isinstance(leaf, ThermalEquilibriumU):
                    v = np.asarray(leaf.values)[c]
                    v6 = (np.repeat(v, 6, axis=1) if v.shape[1] == 1
                          else p2_from_p1(v) if v.shape[1] == 3 else v)
                    acc += a * v6
                else:
                    raise TypeError('unsupported leaf in a natural-BC value: %r' % (leaf,))
            c0.append(np.full(len(c), k0))
            p2.append(acc)
        return dict(cell=cells, local=local, slot=slot, c0=np.concatenate(c0),
                    p2=np.ascontiguousarray(np.concatenate(p2)))

    def natbc_values(self):
        """Integrals oint phi_bc psi_i . n ds for the three facet dofs of every
        'poisson/phi' facet, in the order the plan was given (host evaluation; kept for
        backends without an exterior-facet routine and as a cross-check in the tests)."""
        if not self._nat:
            return np.zeros(0)
        T = reference_tables()
        t, w = T.facet_gl_t, T.facet_gl_w
        _, det = self.mesh.cell_jacobians()
        out = []
        for bc, facets, c, l in self._nat:
            val = bc.value
            mag = as_expr(val.m_as('V') if hasattr(val, 'm_as') else val)
            ref_pts = np.stack([np.array([[0, 0], [1, 0], [0, 1]], float)[_REF_EDGES[f][0]][None, :]
                                + t[:, None] * (np.array([[0, 0], [1, 0], [0, 1]], float)[_REF_EDGES[f][1]]
                                                - np.array([[0, 0], [1, 0], [0, 1]], float)[_REF_EDGES[f][0]])[None, :]
                                for f in l])                                  # [n, q, 2]
            phi = mag.eval_points(c, ref_pts)                                  # [n, q]
            sgn = _REF_OUT[l] * np.sign(det[c])
            tr = T.facet_bdm_trace[l]                                          # [n, 12, q]
            loc = np.stack([3 * l, 3 * l + 1, 3 * l + 2], axis=1)              # facet dofs
            trf = np.take_along_axis(tr, loc[:, :, None], axis=1)              # [n, 3, q]
            out.append((sgn[:, None] * np.einsum('nq,nkq,q->nk', phi, trf, w)).ravel())
        return np.concatenate(out)

    # -- state -----------------------------------------------------------------------
    @property
    def backend(self):
        if self._backend is None:
            factory = self.pdd.problem_data.backend_factory
            if factory is None:
                from ..fem.backend import CudaBackend
                factory = CudaBackend
            self._backend = factory(self.pa)
            self._backend.set_state(self._u_host, self._base_w0 if self.full else None,
                                    self.thmeq_phi if self.full else None, self.phi_opt)
            self._backend.set_bc_values(self.bc_values)
        return self._backend

    @property
    def N(self):
        return self.pa.N

    def u_host(self):
        if self._backend is not None:
            return self._backend.get_u()
        return self._u_host

    def set_u_host(self, u):
        self._u_host = np.array(u, dtype=np.float64)
        if self._backend is not None:
            self._backend.set_u(self._u_host)

    def base_w_host(self):
        if self._backend is not None and self.full:
            return self._backend.get_base_w()
        return self._base_w0

    def phi_nodal(self):
        lay = self.pa.layout()
        return self.u_host()[self.pa.cell_dofs[:, lay['dg'][0]]]

    def tolerance_vector(self):
        """trial tolerances (``poisson_drift_diffusion1.py1:63,70,416,423``):
        phi 1e-6, E 1e-4, delta_w 1e-6, j 1e-10."""
        P = self.pa.P
        return self.pa.tolerance_vector([1e-6] * P, [1e-4] + [1e-10] * (P - 1))

    def initialize_from(self, other):
        """``initialize_from_attributes`` (``poisson_drift_diffusion.py:418-442,
        :712-746``): phi, E, thermal_equilibrium_phi; bands start at qfl = 0, j = 0."""
        lay = self.pa.layout()
        u = np.zeros(self.pa.N)
        phi = other.phi_nodal()
        u[self.pa.cell_dofs[:, lay['dg'][0]]] = phi
        if isinstance(other, PDDSystem):
            ou = other.u_host()
            olay = other.pa.layout()
            u[self.pa.cell_dofs[:, lay['bdm'][0]]] = ou[other.pa.cell_dofs[:, olay['bdm'][0]]]
            if self.full:
                self.thmeq_phi = p2_from_p1(phi) if not other.full else other.thmeq_phi.copy()
        self._base_w0 = np.zeros_like(self._base_w0)
        self._u_host = u
        if self._backend is not None:
            self._backend.set_state(self._u_host, self._base_w0 if self.full else None,
                                    self.thmeq_phi if self.full else None, self.phi_opt)

    # -- a10 optics ------------------------------------------------------------------
    @property
    def phi_opt(self):
        if self._phi_opt_stale:
            self._phi_opt_host = np.array(self._backend.get_phi_opt(), copy=True)
            self._phi_opt_stale = False
        return self._phi_opt_host

    def _optical_lowering(self):
        """CG2 space + per field (direction, inflow dofs, inflow flux): the lowered
        form of ``optical.spatial.add_BC('<field>/Phi', facets, flux)``
        (``optical1.py1:57-61``)."""
        if getattr(self, '_opt_low', None) is None:
            from ..fem.cg2 import CG2Space
            space = CG2Space(self.mesh)
            osp = self.pdd.problem_data.optical.spatial
            fields = []
            for fld in self.fields:
                dofs, vals = [], []
                for bc, facets in osp.bc_facets(fld.key + '/Phi'):
                    if len(facets) == 0:
                        continue
                    d = space.facet_dofs(facets)
                    v = bc.value
                    v = v.m_as('1/mesh_unit^2/s') if hasattr(v, 'm_as') else v
                    dofs.append(d)
                    vals.append(np.full(len(d), float(v)))
                if dofs:
                    d = np.concatenate(dofs)
                    ud, first = np.unique(d, return_index=True)
                    fields.append((tuple(float(x) for x in fld.direction[:2]),
                                   ud.astype(np.int32), np.concatenate(vals)[first]))
                else:
                    fields.append((tuple(float(x) for x in fld.direction[:2]),
                                   np.zeros(0, np.int32), np.zeros(0)))
            self._opt_low = (space, fields)
        return self._opt_low

    def optical_update(self):
        """``run_optical_subsolvers`` (``physics/steppers.py:120-136``): for every
        field update_input -> solve -> update_output, on the backend's device state."""
        if not self.fields:
            return
        space, fields = self._optical_lowering()
        scale = float(self.pdd.problem_data.optical.Phi_scale.magnitude)
        specs = [(f, d, dofs, vals * scale) for f, (d, dofs, vals) in enumerate(fields)]
        self.backend.optical_update(space, specs)
        self._phi_opt_stale = True       # the backend's copy is newer; pulled on demand

    def photon_flux_host(self):
        """[n_fields, Nc, 6] photon flux on the PDD mesh (``Phi_pddproj``)."""
        return self.phi_opt

    def set_photon_flux(self, field_index, values):
        """P2 nodal photon flux [Nc, 6] of one optical field, 1/(mesh_unit^2 s)."""
        self.phi_opt[field_index] = values
        if self._backend is not None:
            self._backend.set_phi_opt(self.phi_opt)

    # -- hot-path forwards ------------------------------------------------------------
    def assemble(self):
        be = self.backend
        if hasattr(be, 'set_natbc'):
            be.set_natbc(self.natbc_spec())          # exterior-facet integrals on the backend
        else:
            be.set_natbc_values(self.natbc_values())
        be.assemble()

    def rebase_w(self, band):
        be = self.backend
        o = self.ohmic[band] if self.ohmic else None
        bobj = self.bands[band]
        thr = tuple(getattr(bobj, 'mixedqfl_debug_fill_thresholds', (0.0, 0.0)))
        if thr != (0.0, 0.0) or getattr(bobj, 'mixedqfl_debug_fill_with_zero_except_bc', False):
            return self._rebase_w_flood_fill(band, o, thr, bobj)
        if o is None:
            be.rebase(band, None, None)
            return
        if 'thm' not in o:
            th = self.thmeq_phi[o['c']]
            mid = 3 + o['l']                      # P2 node on the midpoint of edge l
            avg_f = (th[np.arange(len(o['c'])), o['va']] + th[np.arange(len(o['c'])), o['vb']]
                     + 4.0 * th[np.arange(len(o['c'])), mid]) / 6.0
            o['thm'] = np.bincount(o['inv'], weights=o['w'] * avg_f)
        be.rebase(band, o, o['thm'])

    def _ohmic_bc_cell_values(self, band, o):
        """(cells, values) next to the Ohmic contacts of ``band``: facet average of
        u_to_w(u_bc) = e (phi_thmeq - phi) weighted per cell (``poisson_drift_diffusion1.py1:
        362-367``, ``fem/spatial.py:185-235``) -- the host restatement of what ``sb_rebase_w``
        receives as (bc_cells, bc_avg)."""
        if o is None:
            return (np.zeros(0, np.int64), np.zeros(0))
        if 'thm' not in o:
            th = self.thmeq_phi[o['c']]
            mid = 3 + o['l']
            ar = np.arange(len(o['c']))
            avg_f = (th[ar, o['va']] + th[ar, o['vb']] + 4.0 * th[ar, mid]) / 6.0
            o['thm'] = np.bincount(o['inv'], weights=o['w'] * avg_f)
        phi = self.phi_nodal()[o['c']]
        ar = np.arange(len(o['c']))
        phi_f = 0.5 * (phi[ar, o['va']] + phi[ar, o['vb']])
        phi_avg = np.bincount(o['inv'], weights=o['w'] * phi_f)
        return (np.asarray(o['cells'], dtype=np.int64), o['thm'] - phi_avg)

    def _rebase_w_flood_fill(self, band, o, thresholds, bobj):
        """``mixedqfl_do_rebase_w`` with non-zero ``mixedqfl_debug_fill_thresholds``
        (``poisson_drift_diffusion1.py1:339-383`` -> ``PartitionOffsets``,
        ``fem/offset_partitioning.py:108-169``): offsets by flood fill on the host, delta_w
        re-expressed so that base_w + delta_w is unchanged."""
        from ..fem.offset_partitioning import partition_offsets
        lay = self.pa.layout()
        dg = self.pa.cell_dofs[:, lay['dg'][1 + band]]
        u = np.array(self.u_host(), dtype=np.float64)
        base = np.array(self.base_w_host(), dtype=np.float64)
        w_avg = base[band] + u[dg].mean(axis=1)
        bc = ((np.zeros(0, np.int64), np.zeros(0)))
        if getattr(bobj, 'mixedqfl_debug_fill_from_boundary', True):
            bc = self._ohmic_bc_cell_values(band, o)
        new = partition_offsets(
            w_avg, self.mesh.cell_to_cells_adjacency_via_facet(), thresholds, bc,
            getattr(bobj, 'mixedqfl_debug_fill_with_zero_except_bc', False))
        u[dg] += (base[band] - new)[:, None]
        base[band] = new
        self._base_w0 = base
        self.set_u_host(u)
        if self._backend is not None:
            self._backend.set_state(self._u_host, self._base_w0 if self.full else None,
                                    self.thmeq_phi if self.full else None, self.phi_opt)

    def terminal_current(self, band, facets):
        """(integral of j.n ds, total length) over exterior facets, form units."""
        p = 1 + band
        c = self.mesh.facet_cells[facets, 0]
        l = self.mesh.facet_local[facets, 0]
        return self.backend.surface_flux(p, c, l)

    # -- save / restore for the adaptive stepper (adaptive_stepper.py:295-307) ---------
    def save(self):
        d = dict(u=self.u_host().copy(), base_w=np.array(self.base_w_host(), copy=True))
        if self.fields:                  # optical/<field>/mixed_solution (optical.py:291-297)
            d['phi_opt'] = np.array(self.phi_opt, copy=True)
        return d

    def restore(self, saved):
        if 'phi_opt' in saved:
            self._phi_opt_host = saved['phi_opt'].copy()
            self._phi_opt_stale = False
        self._u_host = saved['u'].copy()
        self._base_w0 = saved['base_w'].copy()
        if self._backend is not None:
            self._backend.set_state(self._u_host, self._base_w0 if self.full else None,
                                    self.thmeq_phi if self.full else None, self.phi_opt)


class LCNSystem(_SystemBase):
    """goal 'local charge neutrality': DG1 potential only, residual
    <w rho(phi)/eps> with UFL's estimated degree-4 (6-point) rule
    (``poisson_drift_diffusion1.py1:89-125``).  Cells decouple (block-diagonal
    3x3 Jacobian); one Newton iteration is a single kernel launch
    (``sb_lcn_iteration``, ``csrc/lcn.inl``) driven through the backend."""

    def __init__(self, pdd):
        super().__init__(pdd)
        tp = M.new_tag_params(len(self.cvs))
        self._band_rows(tp)
        self.tag_params = tp
        self.model = M.make_model(self.band_model_list(), [], n_tags=len(self.cvs),
                                  bands_have_dofs=False)
        nc = self.mesh.num_cells
        self._phi_host = np.zeros((nc, 3))
        self.cell_vertices = np.ascontiguousarray(
            self.mesh.coordinates[self.mesh.cells].reshape(nc, 6))
        self.N = 3 * nc
        self.full = False
        self._backend = None

    @property
    def backend(self):
        if self._backend is None:
            factory = self.pdd.problem_data.backend_factory
            if factory is None:
                from ..fem.backend import CudaLCNBackend
                cls = CudaLCNBackend
            else:
                cls = factory.lcn_backend
            self._backend = cls(self)
            self._backend.set_phi(self._phi_host)
        return self._backend

    @property
    def phi(self):
        if self._backend is not None:
            return self._backend.get_phi()
        return self._phi_host

    @phi.setter
    def phi(self, value):
        self._phi_host = np.array(value, dtype=np.float64, copy=True).reshape(-1, 3)
        if self._backend is not None:
            self._backend.set_phi(self._phi_host)

    def phi_nodal(self):
        return self.phi

    def u_host(self):
        return self.phi.ravel()

    def tolerance_vector(self):
        return np.full(self.N, 1e-6)

    def initialize_from(self, other):
        self.phi = np.array(other.phi_nodal(), copy=True)

    def iterate(self, tol, omega, max_du, rel_tol, abs_tol):
        """One damped Newton iteration; returns the metric dict of ``newton_update``."""
        return self.backend.iterate(tol, omega, max_du, rel_tol, abs_tol)

    def save(self):
        return dict(phi=self.phi.copy())

    def restore(self, saved):
        self.phi = saved['phi'].copy()
