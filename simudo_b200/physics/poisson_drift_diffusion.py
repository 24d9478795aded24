"""Poisson / drift-diffusion problem definition (host API).

Keeps the reference's user-facing surface
(``simudo/physics/poisson_drift_diffusion.py``): ``PoissonDriftDiffusion`` with
``bands``, ``poisson``, ``electro_optical_processes``, ``spatial``,
``easy_add_band`` (:138-204), ``easy_add_electro_optical_process`` (:242-245),
``easy_auto_pre_solve`` (:282-294), ``initialize_from`` (:296-306),
``get_to_save`` (:308-316); band classes ``MixedQflNondegenerateBand`` /
``MixedQflIntermediateBand`` and their statistics-only bases.

The weak forms themselves (``poisson_drift_diffusion1.py1``) are not built
symbolically: ``simudo_b200.physics.lowering.PDDSystem`` turns the declared
bands / processes / rules / BCs into the tables the CUDA assembly consumes.
"""
import logging

import numpy as np

from ..fem.expr import CellFunction, Constant, ThermalEquilibriumU, UnitVec, ZeroVec, p2_from_p1
from ..fem.spatial import Spatial
from ..mesh import CellRegions
from ..util import NameDict, SetattrInitMixin
from ..fem import model as M

__all__ = ['SaveItem', 'GetToSaveMixin', 'PoissonDriftDiffusion', 'Poisson',
           'MixedPoisson', 'LocalChargeNeutralityPoisson', 'Band', 'NondegenerateBand',
           'IntermediateBand', 'MixedQflNondegenerateBand', 'MixedQflIntermediateBand',
           'MixedQflBandMixin', 'MeshUtil']


class SaveItem(SetattrInitMixin):
    name = None
    value = None
    file_save = True
    solver_save = False


class GetToSaveMixin(object):
    def get_to_save(self):
        return NameDict()


class MeshUtil(SetattrInitMixin):
    """The few ``mesh_util`` members user scripts touch (``fem/mesh_util1.py1``):
    ``zerovec``, ``unitvec_x``, ``unitvec_y``, ``zero``, ``Constant``, ``unit_registry``."""
    mesh_data = None
    unit_registry = None

    @property
    def zerovec(self):
        return ZeroVec()

    @property
    def unitvec_x(self):
        return UnitVec(0)

    @property
    def unitvec_y(self):
        return UnitVec(1)

    @property
    def zero(self):
        return Constant(0.0)

    def Constant(self, q):
        return Constant(q.magnitude) * q.units if hasattr(q, 'units') else Constant(q)

    @property
    def mesh(self):
        return self.mesh_data.mesh


class _Sub(SetattrInitMixin):
    pdd = None

    @property
    def spatial(self):
        return self.pdd.spatial

    @property
    def unit_registry(self):
        return self.pdd.unit_registry

    @property
    def mesh_util(self):
        return self.pdd.mesh_util


class Poisson(_Sub, GetToSaveMixin):
    """Electrostatics sub-problem; ``phi`` / ``thermal_equilibrium_phi`` are
    live views (unit-carrying ``CellFunction``) on the solution."""
    key = 'poisson'
    mixed_debug_quad_degree_rho = 20

    @property
    def name(self):
        return self.key

    @property
    def at_thermal_equilibrium(self):
        return self.pdd.problem_data.goal == 'thermal equilibrium'

    @property
    def phi(self):
        U = self.unit_registry
        return CellFunction(getter=lambda: self.pdd.system.phi_nodal(), name='phi') * U.V

    @property
    def thermal_equilibrium_phi(self):
        U = self.unit_registry
        goal = self.pdd.problem_data.goal
        if goal == 'local charge neutrality':
            return None
        if goal == 'thermal equilibrium':       # we *are* the thermal equilibrium
            return CellFunction(getter=lambda: p2_from_p1(self.pdd.system.phi_nodal()),
                                name='thmeq_phi') * U.V
        return CellFunction(getter=lambda: self.pdd.system.thmeq_phi,
                            name='thmeq_phi') * U.V


class MixedPoisson(Poisson):
    pass


class LocalChargeNeutralityPoisson(Poisson):
    pass


class Band(_Sub, GetToSaveMixin):
    """A band whose carriers share one quasi-Fermi level.  ``key``/``name``,
    ``sign`` (-1 electrons, +1 holes), ``subdomain`` (CellRegion)."""
    key = None
    sign = None
    subdomain = None
    subdomain_inv = None
    kind = None
    has_dofs = False
    use_constant_mobility = False

    @property
    def name(self):
        return self.key

    @property
    def spatial_prefix(self):
        return self.name + '/'

    @property
    def thermal_equilibrium_u(self):
        """Carrier density at qfl = 0 in the thermal-equilibrium potential; only
        meaningful as an Ohmic-contact value (``poisson_drift_diffusion1.py1:145``)."""
        U = self.unit_registry
        return ThermalEquilibriumU(self) * U('1/mesh_unit^3')

    @property
    def index(self):
        return list(self.pdd.bands.keys()).index(self.key)


class NondegenerateBand(Band):
    """Boltzmann statistics: u = N0 exp(s (E0 - (w + q phi)) / kT)
    (``poisson_drift_diffusion1.py1:170-197``); keys ``<band>/energy_level``,
    ``<band>/effective_density_of_states``, ``<band>/mobility``."""
    kind = M.BAND_NONDEGENERATE
    density_key = 'effective_density_of_states'
    mobility_key = 'mobility'


class IntermediateBand(Band):
    """Sharp level with Fermi-Dirac filling: u = N_I f((s (w + q phi - E_I))/kT)
    (``:199-245``); keys ``<band>/number_of_states``, ``<band>/energy_level``,
    ``<band>/mobility0`` (or deprecated ``empty_band_mobility``),
    ``<band>/degeneracy``."""
    kind = M.BAND_INTERMEDIATE
    density_key = 'number_of_states'
    mobility_key = 'mobility0'


class MixedQflBandMixin(object):
    """Mixed method with (delta_w, j) unknowns (``:259-430``); debug attributes
    keep the reference names (``poisson_drift_diffusion.py:694-700``)."""
    has_dofs = True
    mixedqfl_debug_fill_thresholds = (0., 0.)
    mixedqfl_debug_fill_from_boundary = True
    mixedqfl_debug_fill_with_zero_except_bc = False
    mixedqfl_debug_use_bchack = False
    mixedqfl_debug_quad_degree_super = 20
    mixedqfl_debug_quad_degree_g = 8

    def mixedqfl_do_rebase_w(self):
        self.pdd.system.rebase_w(self.index)

    @property
    def mixedqfl_base_w(self):
        U = self.unit_registry
        return CellFunction(getter=lambda: self.pdd.system.base_w_host()[self.index][:, None],
                            name=self.name + '/base_w') * U.eV


class MixedQflNondegenerateBand(MixedQflBandMixin, NondegenerateBand):
    _base_band_class = NondegenerateBand


class MixedQflIntermediateBand(MixedQflBandMixin, IntermediateBand):
    _base_band_class = IntermediateBand


class PoissonDriftDiffusion(GetToSaveMixin, SetattrInitMixin):
    problem_data = None

    def __init__(self, **kwargs):
        super().__init__(**kwargs)
        self.bands = NameDict()
        self.electro_optical_processes = NameDict()
        self.spatial = Spatial(parent=self)
        self._system = None
        self._poisson = None

    # -- plumbing ----------------------------------------------------------------
    @property
    def mesh_data(self):
        return self.problem_data.mesh_data

    @property
    def unit_registry(self):
        return self.problem_data.unit_registry

    @property
    def mesh_util(self):
        return MeshUtil(mesh_data=self.mesh_data, unit_registry=self.unit_registry)

    @property
    def temperature(self):
        return self.spatial.get_table('temperature', 'K', required=True)

    @property
    def kT(self):
        """{cell value: k T in eV} (``poisson_drift_diffusion.py:125-131``)."""
        U = self.unit_registry
        k = (1.0 * U.K * U.boltzmann_constant).m_as('eV')
        return {cv: T * k for cv, T in self.temperature.items()}

    @property
    def poisson(self):
        if self._poisson is None:
            goal = self.problem_data.goal
            if goal == 'local charge neutrality':
                cls = LocalChargeNeutralityPoisson
            elif goal in ('thermal equilibrium', 'full'):
                cls = MixedPoisson
            else:
                raise AssertionError('bad goal')
            self._poisson = cls(pdd=self)
        return self._poisson

    def get_subproblem_children(self):
        return (self.poisson,) + tuple(self.bands)

    # -- declaration ---------------------------------------------------------------
    def easy_add_band(self, name, band_type=None, sign='auto', subdomain=None):
        cls = band_type
        if cls is None:
            if name in ('CB', 'VB'):
                cls = 'nondegenerate'
            elif name == 'IB':
                cls = 'intermediate'
        if sign == 'auto':
            if name in ('CB', 'IB'):
                sign = -1
            elif name == 'VB':
                sign = 1
            else:
                raise ValueError(
                    "Don't know how to automatically set sign for this band, either set "
                    "it manually or pass sign=None")
        if cls in ('nondegenerate', 'boltzmann'):
            cls = MixedQflNondegenerateBand
        elif cls in ('degenerate', 'parabolic'):
            raise NotImplementedError('degenerate parabolic bands not implemented yet')
        elif cls in ('intermediate', 'sharp'):
            cls = MixedQflIntermediateBand
        if not (isinstance(cls, type) and issubclass(cls, Band)):
            raise TypeError()
        kwargs = dict(key=name, subdomain=subdomain)
        if sign is not None:
            kwargs['sign'] = sign
        return self.add_band(cls, kwargs)

    def add_band(self, cls, kwargs):
        kwargs.setdefault('pdd', self)
        if self.problem_data.goal != 'full':
            cls = cls._base_band_class           # statistics only, qfl == 0
        band = cls(**kwargs)
        R = CellRegions()
        if band.subdomain is None:
            band.subdomain = R.domain
        band.subdomain_inv = R.domain - band.subdomain
        self.spatial.add_rule('{}/extent'.format(band.name), band.subdomain, 1)
        self.spatial.add_rule('{}/extent'.format(band.name), band.subdomain_inv, float('nan'))
        self.bands.add(band)
        self._system = None
        return band

    def easy_add_electro_optical_process(self, cls, **kwargs):
        inst = cls(pdd=self, optical=self.problem_data.optical, **kwargs)
        self.electro_optical_processes.add(inst)
        self._system = None
        return inst

    # -- lowered system -------------------------------------------------------------
    @property
    def system(self):
        if self._system is None:
            from .lowering import PDDSystem, LCNSystem
            goal = self.problem_data.goal
            self._system = (LCNSystem(self) if goal == 'local charge neutrality'
                            else PDDSystem(self))
        return self._system

    def get_solution_function(self):
        return self.system

    # -- stages --------------------------------------------------------------------
    def easy_create_newton_solver(self):
        from ..fem.newton_solver import (LocalChargeNeutralitySolver,
                                         ThermalEquilibriumSolver)
        goal = self.problem_data.goal
        if goal == 'local charge neutrality':
            cls = LocalChargeNeutralitySolver
        elif goal == 'thermal equilibrium':
            cls = ThermalEquilibriumSolver
        else:
            raise AssertionError('bad goal')
        return cls.from_nice_obj(self)

    def easy_auto_pre_solve(self, parameters=None):
        solver = self.easy_create_newton_solver()
        if parameters is not None:
            solver.parameters.update(parameters)
        solver.logger = logging.getLogger('newton.' + self.problem_data.goal_abbreviated)
        solver.parameters.update(
            absolute_tolerance_vector=self.system.tolerance_vector())
        solver.solve()
        return solver

    def initialize_from(self, other_pdd):
        self.system.initialize_from(other_pdd.system)

    def get_to_save(self):
        r = NameDict()
        r.add(SaveItem(name='pdd/mixed_solution', value=self.system, solver_save=True))
        return r
