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
    """Continuation toward a list of target values of one scalar parameter.

    Written from the behaviour of the reference's stepper (``adaptive_stepper.py:82-257``),
    with its public attribute and hook names, as an explicit state machine:

    * state: the trail of attempts ``(parameter, converged?)``, a bounded set of
      checkpoints (the ``keep_saved_solutions_count`` most recent converged states) and
      the step size;
    * the next attempt starts from the last converged parameter; after a failure the step
      is multiplied by ``update_parameter_failure_factor`` and the last checkpoint is put
      back (``user_solution_add(solution, 0, checkpoint, 1)``); after a success that used
      the full step it is multiplied by ``update_parameter_success_factor``; the step never
      overshoots the current target, and a failed clipped step shrinks from the clipped
      length;
    * a ``RuntimeError`` of the solver (dolfin / PETSc in the reference, the device library
      here) is a failed attempt, except on the very first one;
    * the march to a target ends when it has been reached, or -- ``StepperError`` -- when
      an attempt failed and step / max(abs(parameter), initial step) < ``stepper_rel_tol``;
    * after every target ``output_writer.write_output(solution, parameter)``; a true
      ``break_criteria_cb(solution)`` ends the sweep.
    """
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
        self._trail = []            # [(parameter, converged)]
        self._checkpoints = []      # [(trail index, saved solution)], newest last
        self.actual_step = None

    # -- the trail ---------------------------------------------------------------------
    @property
    def prior_solutions(self):
        """Every attempt so far as ``AdaptiveLogEntry`` (saved solutions only while kept)."""
        kept = dict(self._checkpoints)
        return [AdaptiveLogEntry(p, ok, kept.get(i)) for i, (p, ok) in enumerate(self._trail)]

    def add_prior_solution(self, parameter_value, solver_succeeded, saved_solution):
        if saved_solution is not None and not solver_succeeded:
            raise AssertionError('why save a failed solution?')
        self._trail.append((parameter_value, bool(solver_succeeded)))
        if saved_solution is not None:
            self._checkpoints.append((len(self._trail) - 1, saved_solution))
            del self._checkpoints[:-self.keep_saved_solutions_count or None]

    def get_last_successful(self, count):
        kept = dict(self._checkpoints)
        good = [i for i, (_, ok) in enumerate(self._trail) if ok][-count:]
        return [AdaptiveLogEntry(self._trail[i][0], True, kept.get(i)) for i in good]

    def _last_converged(self):
        for i in range(len(self._trail) - 1, -1, -1):
            if self._trail[i][1]:
                return i
        return None

    def reset_parameter_to_last_successful(self):
        i = self._last_converged()
        self.parameter = self.parameter_start_value if i is None else self._trail[i][0]

    # -- one attempt ---------------------------------------------------------------------
    def update_parameter(self):
        i = self._last_converged()
        if i is None:                       # nothing converged yet: (re)try the start value
            self.parameter = self.parameter_start_value
            return
        base = self._trail[i][0]
        failed = not self._trail[-1][1]
        if failed:
            logging.getLogger('stepper').info('** Newton solver failed, trying smaller step size')
            self.step_size *= self.update_parameter_failure_factor
        room = abs(self.parameter_target_value - base)
        clipped = room < self.step_size
        self.actual_step = room if clipped else self.step_size
        if failed:
            if clipped:
                self.step_size = self.actual_step * self.update_parameter_failure_factor
            self.user_solution_add(self.solution, 0.0, dict(self._checkpoints).get(i), 1.0)
        elif not clipped:
            self.step_size *= self.update_parameter_success_factor
        self.parameter = base + self.step_sign * self.actual_step

    def user_solver(self, solution, parameter):
        raise NotImplementedError()

    def user_solution_save(self, solution):
        raise NotImplementedError()

    def user_solution_add(self, solution, solution_scale, saved_solution, saved_scale):
        raise NotImplementedError()

    def do_iteration(self):
        self.update_parameter()
        try:
            ok = bool(self.user_solver(self.solution, self.parameter))
        except RuntimeError:
            if not self._trail:
                raise
            ok = False
        self.add_prior_solution(self.parameter, ok,
                                self.user_solution_save(self.solution) if ok else None)

    # -- the sweep -----------------------------------------------------------------------
    def _stalled(self, initial_step):
        if not self._trail or self._trail[-1][1]:
            return False
        scale = max(abs(self.parameter), initial_step)
        return abs(self.step_size) / scale < self.stepper_rel_tol

    def do_loop(self):
        log = logging.getLogger('stepper')
        initial_step = self.step_size
        for target in self.parameter_target_values:
            self.reset_parameter_to_last_successful()
            self.parameter_target_value = target
            self.step_sign = -1. if self.parameter > target else 1.
            gave_up = False
            while True:
                i = self._last_converged()
                if i is not None and self._trail[i][0] == target:
                    break
                if self._stalled(initial_step):
                    gave_up = True
                    break
                self.do_iteration()
            if self.output_writer is not None and self._last_converged() is not None:
                log.info('Writing output')
                self.output_writer.write_output(self.solution, self.parameter)
            if gave_up:
                log.error('Insufficient progress, stopping.')
                raise StepperError(
                    'Solver unable to progress. Among many possible causes of this error '
                    'are: poorly formed problem, incorrect boundary conditions, '
                    'inappropriate material parameters, or insufficient meshing.')
            if self.break_criteria_cb is not None and self.break_criteria_cb(self.solution):
                log.info('*** Break criteria met, stopping.')
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
