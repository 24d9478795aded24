from .newton_solver import (NewtonSolver, NewtonSolution, NewtonSolverMaxDu,
                            NewtonSolverLogDamping)
from .adaptive_stepper import AdaptiveStepper, ConstantStepperMixin, StepperError
from .spatial import Spatial
from .expr import Constant


def setup_dolfin_parameters():
    """No-op: there is no FFC JIT to configure (``fem/dolfin_parameters.py``)."""
