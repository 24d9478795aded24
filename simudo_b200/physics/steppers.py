"""Bias / illumination continuation steppers.

Mirror of ``simudo/physics/steppers.py``: ``NonequilibriumCoupledStepper``
builds one Newton solver per step with the reference's parameters (:82-89:
``maximum_iterations`` 200 on the first step else 25, ``minimum_iterations`` 3,
``relative_tolerance`` 1e-4, ``absolute_tolerance`` 1, ``extra_iterations`` 3
then 0) and hooks (:138-173: ``rebase_w`` before every iteration, bail-out when
``|du| > 1e40`` after 20 iterations, optional optical sub-solves);
``VoltageStepper`` (:180-189) and ``OpticalIntensityAdaptiveStepper`` (:192-210).
"""
import logging

from ..fem.adaptive_stepper import AdaptiveStepper, ConstantStepperMixin
from ..fem.newton_solver import NewtonSolver

__all__ = ['NewtonBailoutException', 'NonequilibriumCoupledStepper',
           'NonequilibriumCoupledConstantStepper', 'VoltageStepper',
           'OpticalIntensityAdaptiveStepper']


class NewtonBailoutException(Exception):
    pass


class NonequilibriumCoupledStepper(AdaptiveStepper):
    selfconsistent_optics = False

    def __init__(self, **kwargs):
        ow = kwargs.get('output_writer', None)
        if ow is not None:
            if isinstance(ow, str):
                from ..io.output_writer import OutputWriter
                ow = OutputWriter(filename_prefix=ow, plot_iv=True)
            ow.stepper = self
            kwargs['output_writer'] = ow
        super().__init__(**kwargs)

    @property
    def to_save_objects(self):
        return {'pdd/mixed_solution': self.solution.pdd.system}

    def user_make_solver(self, solution):
        pdd = solution.pdd
        be_extra_careful = not self.get_last_successful(1)
        # the reference hard-wires NewtonSolver (steppers.py:87); `newton_solver_class` lets a
        # sweep choose one of its other solvers, e.g. NewtonSolverLogDamping
        solver = (getattr(self, 'newton_solver_class', None) or NewtonSolver).from_nice_obj(pdd)
        solver.parameters.update(
            maximum_iterations=200 if be_extra_careful else 25,
            omega_cb=lambda s: 1.0,
            extra_iterations=3 if be_extra_careful else 0,
            minimum_iterations=3, relative_tolerance=1e-4, absolute_tolerance=1,
            absolute_tolerance_vector=pdd.system.tolerance_vector())

        def bailout_if_error_too_large(slv):
            if slv.solution.du_norm > 1e40 and slv.iteration > 20:
                raise NewtonBailoutException()

        def rebase_w(slv):
            for b in pdd.bands:
                if hasattr(b, 'mixedqfl_do_rebase_w'):
                    b.mixedqfl_do_rebase_w()

        if self.selfconsistent_optics:
            def run_optical(slv):
                solution.optical_update(pdd)
            if hasattr(solution, 'optical_update'):
                solver.user_before_first_iteration_hooks.append(run_optical)
                solver.user_post_iteration_hooks.append(run_optical)
        for proc in pdd.electro_optical_processes:
            for lst, hook in ((solver.user_pre_iteration_hooks, proc.pre_iteration_hook),
                              (solver.user_before_first_iteration_hooks,
                               proc.pre_first_iteration_hook),
                              (solver.user_post_iteration_hooks, proc.post_iteration_hook)):
                if hook is not None:
                    lst.append(hook)
        solver.user_post_iteration_hooks.append(bailout_if_error_too_large)
        solver.user_pre_iteration_hooks.append(rebase_w)
        return solver

    def user_solver(self, solution, parameter):
        logging.info('%%%%%%%%%% adaptive solve parameter={} step_size={}'.format(
            parameter, self.step_size))
        try:
            return super().user_solver(solution, parameter)
        except NewtonBailoutException:
            return False


class NonequilibriumCoupledConstantStepper(ConstantStepperMixin, NonequilibriumCoupledStepper):
    pass


class VoltageStepper(NonequilibriumCoupledConstantStepper):
    """Ramp the bias at one or more contacts."""
    update_parameter_success_factor = 1.5
    update_parameter_failure_factor = 0.5
    step_size = 0.1


class OpticalIntensityAdaptiveStepper(NonequilibriumCoupledConstantStepper):
    """Ramp ``optical.Phi_scale`` from 0 to 1."""
    update_parameter_success_factor = 3
    update_parameter_failure_factor = 0.2
    step_size = 1e-30
    parameter_target_values = [1.0]

    @property
    def constants(self):
        return [self.solution.optical.Phi_scale]

    @constants.setter
    def constants(self, v):
        pass

    @property
    def parameter_unit(self):
        return self.unit_registry.dimensionless

    @parameter_unit.setter
    def parameter_unit(self, v):
        pass
