"""Adaptive continuation in a scalar parameter (bias, light intensity).

Same algorithm and attribute names as ``simudo/fem/adaptive_stepper.py``:
grow the step by ``update_parameter_success_factor`` after a converged solve,
restore the last saved solution and shrink by
``update_parameter_failure_factor`` after a failure (:138-173), stop with
``StepperError`` when the step falls below ``stepper_rel_tol`` (:221-257),
keep the last ``keep_saved_solutions_count`` saved solutions (:82,105-110).
``RuntimeError`` from the device path counts as a failed step (:212-217).
"""
import logging
from collections import namedtuple

from ..util import SetattrInitMixin

__all__ = ['BaseAdaptiveStepper', 'AdaptiveStepper', 'ConstantStepperMixin', 'StepperError']

AdaptiveLogEntry = namedtuple('AdaptiveLogEntry', ['parameter', 'success', 'saved_solution'])


class StepperError(Exception):
    pass


class BaseAdaptiveStepper(SetattrInitMixin):
    solution = None
    parameter_start_value = 0.0
    parameter_target_values = [1.0]
    step_size = 1e-2
    stepper_rel_tol = 1e-5
    update_parameter_success_factor = 2.
    update_parameter_failure_factor = 0.5
    keep_saved_solutions_count = 2
    output_writer = None
    break_criteria_cb = None

    def __init__(self, **kwargs):
        super().__init__(**kwargs)
        self.parameter = self.parameter_start_value
        self.prior_solutions = []
        self._idx_ok = []
        self._idx_saved = []

    def add_prior_solution(self, parameter_value, solver_succeeded, saved_solution):
        if saved_solution is not None and not solver_succeeded:
            raise AssertionError('why save a failed solution?')
        self.prior_solutions.append(
            AdaptiveLogEntry(parameter_value, solver_succeeded, saved_solution))
        i = len(self.prior_solutions) - 1
        if saved_solution is not None:
            self._idx_saved.append(i)
        if solver_succeeded:
            self._idx_ok.append(i)
        while len(self._idx_saved) > self.keep_saved_solutions_count:
            k = self._idx_saved.pop(0)
            self.prior_solutions[k] = self.prior_solutions[k]._replace(saved_solution=None)

    def get_last_successful(self, count):
        return [self.prior_solutions[i] for i in self._idx_ok[-count:]]

    def reset_parameter_to_last_successful(self):
        last_ok = self.get_last_successful(1)
        self.parameter = last_ok[0].parameter if last_ok else self.parameter_start_value

    def update_parameter(self):
        last_ok = self.get_last_successful(1)
        if not last_ok:
            self.reset_parameter_to_last_successful()
            return
        last_ok = last_ok[0]
        last = self.prior_solutions[-1]
        if not last.success:
            logging.getLogger('stepper').info('** Newton solver failed, trying smaller step size')
            self.step_size *= self.update_parameter_failure_factor
        self.actual_step = min(self.step_size,
                               abs(self.parameter_target_value - last_ok.parameter))
        if (self.actual_step == self.step_size) and last.success:
            self.step_size *= self.update_parameter_success_factor
        if (self.actual_step < self.step_size) and not last.success:
            self.step_size = self.actual_step * self.update_parameter_failure_factor
        if not last.success:
            self.user_solution_add(self.solution, 0.0, last_ok.saved_solution, 1.0)
        self.parameter = last_ok.parameter + self.step_sign * self.actual_step

    def user_solver(self, solution, parameter):
        raise NotImplementedError()

    def user_solution_save(self, solution):
        raise NotImplementedError()

    def user_solution_add(self, solution, solution_scale, saved_solution, saved_scale):
        raise NotImplementedError()

    def do_iteration(self):
        self.update_parameter()
        try:
            success = self.user_solver(self.solution, self.parameter)
        except RuntimeError:
            success = False
            if len(self.prior_solutions) == 0:
                raise
        saved = self.user_solution_save(self.solution) if success else None
        self.add_prior_solution(self.parameter, success, saved)

    def do_loop(self):
        sufficient_progress = True
        initial_step_size = self.step_size
        for val in self.parameter_target_values:
            self.reset_parameter_to_last_successful()
            self.step_sign = -1. if self.parameter > val else 1.
            self.parameter_target_value = val
            while True:
                last_ok = self.get_last_successful(1)
                failed_prev = bool(self.prior_solutions) and not self.prior_solutions[-1].success
                if last_ok and last_ok[0].parameter == val:
                    break
                if (failed_prev and abs(self.step_size) / max(abs(self.parameter),
                                                               initial_step_size)
                        < self.stepper_rel_tol):
                    sufficient_progress = False
                    break
                self.do_iteration()
            if self.output_writer is not None and self.get_last_successful(1):
                logging.getLogger('stepper').info('Writing output')
                self.output_writer.write_output(self.solution, self.parameter)
            if not sufficient_progress:
                logging.getLogger('stepper').error('Insufficient progress, stopping.')
                raise StepperError(
                    'Solver unable to progress. Among many possible causes of this error '
                    'are: poorly formed problem, incorrect boundary conditions, '
                    'inappropriate material parameters, or insufficient meshing.')
            if self.break_criteria_cb is not None and self.break_criteria_cb(self.solution):
                logging.getLogger('stepper').info('*** Break criteria met, stopping.')
                break


class AdaptiveStepper(BaseAdaptiveStepper):
    """``to_save_objects``: {name: object with save()/restore()}."""
    to_save_objects = None

    def user_solver(self, solution, parameter):
        self.user_apply_parameter_to_solution(solution, parameter)
        solver = self.user_make_solver(solution)
        solver.solve()
        self.last_solver = solver
        return solver.has_converged()

    def user_apply_parameter_to_solution(self, solution, parameter):
        raise NotImplementedError('override me')

    def user_make_solver(self, solution):
        raise NotImplementedError('override me')

    def user_solution_save(self, solution):
        return {k: obj.save() for k, obj in self.to_save_objects.items()}

    def user_solution_add(self, solution, solution_scale, saved_solution, saved_scale):
        if solution_scale != 0 or saved_scale != 1.0:
            raise NotImplementedError('only plain restore is used by the steppers')
        for k, obj in self.to_save_objects.items():
            obj.restore(saved_solution[k])


class ConstantStepperMixin(object):
    """``constants``: quantities wrapping a ``Constant``; each iteration assigns
    the current parameter value (``adaptive_stepper.py:309-334``)."""
    parameter_unit = None
    constants = ()

    def get_mapped_parameter(self):
        x = self.parameter
        if self.parameter_unit is not None:
            x = self.parameter_unit * x
        return x

    @property
    def unit_registry(self):
        return self.solution.unit_registry

    def user_apply_parameter_to_solution(self, solution, parameter_value):
        x = self.get_mapped_parameter()
        for c in self.constants:
            v = x.m_as(c.units) if hasattr(x, 'm_as') else x
            c.magnitude.assign(v)
