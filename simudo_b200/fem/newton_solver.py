"""Newton solver with the reference's control surface.

Mirror of ``simudo/fem/newton_solver.py``: ``NewtonSolver.from_nice_obj(obj)``
(:266-271), ``solver.parameters`` keys (:285-291, :354-361, :535), the three
hook lists (:261-263), ``solve()`` / ``do_iteration()`` (:319-326, :363-369),
``has_converged`` (:371-374), the ``extra_iterations`` endgame (:376-399),
damping variants ``NewtonSolverMaxDu`` (:533-541) and
``NewtonSolverLogDamping`` (:542-547), and ``NewtonSolution`` with cached
``b_norm``, ``du_norm``, ``rel_du_norm``, ``combined_du_norm``, ``has_nans``
(:79-122).  Assembly, linear solve, update and metrics run on the device
through the problem's backend; nothing is copied to the host per iteration
except the five scalars of the convergence test.
"""
import logging

import numpy as np

__all__ = ['NewtonSolution', 'NewtonSolver', 'NewtonSolverMaxDu', 'NewtonSolverLogDamping',
           'LocalChargeNeutralitySolver', 'ThermalEquilibriumSolver',
           'ExitOptimizerException']


class ExitOptimizerException(Exception):
    pass


class NewtonSolution(object):
    """Per-iteration quantities; ``u_`` / ``du`` / ``b`` are host copies made on
    demand (the device owns the data)."""

    def __init__(self, solver):
        self.solver = solver
        self.b_norm = -1.0
        self.du_norm = float('inf')
        self.rel_du_norm = -1.0
        self.combined_du_norm = float('inf')
        self.has_nans = False

    @property
    def system(self):
        return self.solver.system

    @property
    def u_(self):
        return self.system.u_host()

    @property
    def du(self):
        be = getattr(self.system, '_backend', None)
        return be.get_du() if be is not None else self.solver._du_host

    @property
    def b(self):
        be = getattr(self.system, '_backend', None)
        return be.get_b() if be is not None else self.solver._b_host

    @property
    def size(self):
        return self.system.N

    def str_error(self):
        return '|b| {:.6e}, |du/u| {:.6e}, |du/(a*u+c)| {:.6e}'.format(
            self.b_norm, self.rel_du_norm, self.combined_du_norm)


class NewtonSolver(object):
    iteration = 0
    solution_class = NewtonSolution
    extra_iterations = None

    def __init__(self, system, parameters=None):
        self.system = system
        self.do_init_parameters()
        if parameters is not None:
            self.parameters.update(parameters)
        self.solution = self.solution_class(self)
        self.user_before_first_iteration_hooks = []
        self.user_pre_iteration_hooks = []
        self.user_post_iteration_hooks = []
        self.logger = logging.getLogger('newton')

    @classmethod
    def from_nice_obj(cls, obj):
        """``obj`` exposes ``get_solution_function()`` (the lowered system)."""
        return cls(obj.get_solution_function())

    def do_init_parameters(self):
        self.parameters = dict(maximum_iterations=25, minimum_iterations=-1,
                               relaxation_parameter=1.0, absolute_tolerance=1e-5,
                               relative_tolerance=1e-6, extra_iterations=3)

    # -- iteration ---------------------------------------------------------------
    def get_omega(self):
        # only 'omega_callback' is read (newton_solver.py:354-361); the 'omega_cb' entries the
        # stage solvers and steppers put into their parameters are dead in the reference too,
        # so the damping that applies is relaxation_parameter for the first
        # num_damped_iterations (5) iterations, then a full step
        cb = self.parameters.get('omega_callback', None)
        if cb is not None:
            return cb(self)
        if self.iteration <= self.parameters.get('num_damped_iterations', 5):
            return self.parameters['relaxation_parameter']
        return 1.0

    def get_max_du(self):
        return None

    def get_log_alpha(self):
        return None

    def _tolerance_vector(self):
        tol = self.parameters.get('absolute_tolerance_vector', None)
        if tol is None:
            tol = np.ones(self.system.N)
            self.parameters['absolute_tolerance_vector'] = tol
        return tol

    def do_iteration(self):
        for hook in self.user_pre_iteration_hooks:
            hook(self)
        sysm = self.system
        sysm.assemble()
        be = sysm.backend
        be.solve()
        if getattr(self, '_tol_dev', None) is None:
            self._tol_dev = be.T(self._tolerance_vector()) if hasattr(be, 'T') \
                else self._tolerance_vector()
        extra = {}
        if self.get_log_alpha() is not None:
            extra['log_alpha'] = self.get_log_alpha()
        m = be.newton_update(self._tol_dev, self.get_omega(), self.get_max_du(),
                             self.parameters['relative_tolerance'],
                             self.parameters['absolute_tolerance'], **extra)
        s = self.solution
        s.b_norm, s.du_norm = m['b_norm'], m['du_norm']
        s.rel_du_norm, s.combined_du_norm = m['rel_du_norm'], m['combined_du_norm']
        s.has_nans = m['has_nans']
        self.logger.info('{}* iteration {:>3d}; {}'.format(
            '\n' if self.iteration == 0 else '', self.iteration, s.str_error()))
        for hook in self.user_post_iteration_hooks:
            hook(self)

    def solve(self):
        for hook in self.user_before_first_iteration_hooks:
            hook(self)
        while True:
            self.iteration += 1
            self.do_iteration()
            if self.should_stop_real():
                break

    def has_converged(self):
        s = self.solution
        return (not s.has_nans) and (s.combined_du_norm <= 1.)

    def should_stop_real(self):
        extra = self.extra_iterations
        should_stop = self.should_stop()
        if self.solution.has_nans:
            return True
        if extra is not None:
            if should_stop:
                self.extra_iterations = extra = extra - 1
                return extra <= 0
            self.extra_iterations = None
        if should_stop:
            extra = self.parameters.get('extra_iterations')
            if extra is None:
                return True
            self.extra_iterations = extra
            return extra <= 0
        return False

    def should_stop(self):
        if self.iteration < self.parameters['minimum_iterations']:
            return False
        return (self.iteration >= self.parameters['maximum_iterations']
                or self.has_converged() or np.isnan(self.solution.du_norm))


class NewtonSolverMaxDu(NewtonSolver):
    def get_max_du(self):
        return self.parameters.get('maximum_du', None)


class NewtonSolverLogDamping(NewtonSolver):
    """``newton_solver.py:542-547`` (ref. Gaury 2018): the applied step is
    ``sign(du) log1p(alpha |du|) / alpha``; done inside the update kernel
    (``sb_newton_update_damped``)."""

    def get_log_alpha(self):
        return self.parameters.get('logdamping_alpha', 1.72)


class _BlockDiagonalMixin(object):
    """The local-charge-neutrality stage is a block-diagonal 3x3 problem per
    cell: assembly, solve, update and metrics are one launch (``LCNSystem.iterate``
    -> ``sb_lcn_iteration``)."""

    def do_iteration(self):
        sysm = self.system
        if not hasattr(sysm, 'iterate'):
            return super().do_iteration()
        for hook in self.user_pre_iteration_hooks:
            hook(self)
        m = sysm.iterate(self._tolerance_vector(), self.get_omega(), self.get_max_du(),
                         self.parameters['relative_tolerance'],
                         self.parameters['absolute_tolerance'])
        s = self.solution
        s.b_norm, s.du_norm = m['b_norm'], m['du_norm']
        s.rel_du_norm, s.combined_du_norm = m['rel_du_norm'], m['combined_du_norm']
        s.has_nans = m['has_nans']
        self.logger.info('* iteration {:>3d}; {}'.format(self.iteration, s.str_error()))
        for hook in self.user_post_iteration_hooks:
            hook(self)


class LocalChargeNeutralitySolver(_BlockDiagonalMixin, NewtonSolverMaxDu):
    """``poisson_drift_diffusion.py:318-331``."""

    def do_init_parameters(self):
        super().do_init_parameters()
        self.parameters.update(
            relaxation_parameter=0.5, maximum_iterations=2000, extra_iterations=5,
            relative_tolerance=1e-10, absolute_tolerance=0.1, maximum_du=0.02,
            omega_cb=lambda self: 0.2 if self.iteration < 100 else 0.5)


class ThermalEquilibriumSolver(NewtonSolverMaxDu):
    """``poisson_drift_diffusion.py:333-345``."""

    def do_init_parameters(self):
        super().do_init_parameters()
        self.parameters.update(
            relaxation_parameter=0.5, maximum_iterations=500, relative_tolerance=1e-5,
            absolute_tolerance=0.1, extra_iterations=15,
            omega_cb=lambda self: 0.4 if self.iteration < 30 else 0.9)
