"""Behaviour of the adaptive continuation stepper (contract of
``simudo/fem/adaptive_stepper.py:82-257``) with a scripted solver."""
import pytest

from simudo_b200.fem.adaptive_stepper import BaseAdaptiveStepper, StepperError


class Scripted(BaseAdaptiveStepper):
    """Solver that fails whenever the step from the last converged point exceeds ``limit``
    (or raises RuntimeError like the device library would)."""
    limit = 0.3
    raise_instead = False

    def __init__(self, **kw):
        super().__init__(**kw)
        self.state = 'init'
        self.calls, self.restores = [], []
        self.last_ok = None

    def user_solver(self, solution, parameter):
        self.calls.append(parameter)
        ok = self.last_ok is None or abs(parameter - self.last_ok) <= self.limit + 1e-15
        self.state = 'dirty@%r' % parameter
        if not ok and self.raise_instead:
            raise RuntimeError('device failure')
        if ok:
            self.last_ok = parameter
            self.state = 'ok@%r' % parameter
        return ok

    def user_solution_save(self, solution):
        return self.state

    def user_solution_add(self, solution, solution_scale, saved_solution, saved_scale):
        assert (solution_scale, saved_scale) == (0.0, 1.0) and saved_solution is not None
        self.restores.append(saved_solution)
        self.state = saved_solution


class Writer(object):
    def __init__(self):
        self.rows = []

    def write_output(self, solution, parameter):
        self.rows.append(parameter)


@pytest.mark.parametrize('raise_instead', [False, True])
def test_step_growth_shrink_restore_and_targets(raise_instead):
    w = Writer()
    s = Scripted(parameter_target_values=[0.0, 1.0, 0.5], step_size=0.1, output_writer=w,
                 raise_instead=raise_instead)
    s.do_loop()
    assert w.rows == [0.0, 1.0, 0.5]                    # one output per target, in order
    ok = [e.parameter for e in s.prior_solutions if e.success]
    assert ok[0] == 0.0 and 1.0 in ok and ok[-1] == 0.5
    # steps grew (x2 after full successful steps) until a step above the limit failed ...
    assert s.calls[:4] == [0.0, 0.1, pytest.approx(0.3), pytest.approx(0.7)]
    failed = [e.parameter for e in s.prior_solutions if not e.success]
    assert failed and all(p > 0.3 for p in failed)
    # ... after which the last converged state was put back before the retry
    assert s.restores and s.restores[0] == 'ok@%r' % s.calls[2]
    # never overshoots a target, and the sweep turned around for the last one
    assert max(s.calls) <= 1.0 + 1e-15 and s.step_sign == -1.0
    # only the newest checkpoints are kept
    assert sum(e.saved_solution is not None for e in s.prior_solutions) <= s.keep_saved_solutions_count


def test_gives_up_when_the_step_underflows():
    s = Scripted(parameter_target_values=[0.0, 1.0], step_size=0.1, limit=-1.0, stepper_rel_tol=1e-3)
    s.last_ok = None
    with pytest.raises(StepperError):
        s.do_loop()
    assert s.calls[0] == 0.0 and len(s.calls) > 5
    assert abs(s.step_size) / 0.1 < 1e-3


def test_first_failure_by_exception_propagates():
    class Boom(Scripted):
        def user_solver(self, solution, parameter):
            raise RuntimeError('no device')
    with pytest.raises(RuntimeError):
        Boom(parameter_target_values=[1.0]).do_loop()


def test_thermionic_heterojunction_fails_by_name():
    # reference physics/heterojunction.py:11-135: same constructor; the facet term is not
    # built on the device path and there is no CPU fallback, so register() must raise
    from simudo_b200.physics import ThermionicHeterojunction
    h = ThermionicHeterojunction(band=object(), boundary=object())
    with pytest.raises(NotImplementedError, match='heterojunction'):
        h.register()
