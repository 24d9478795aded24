"""``ProblemData``: root object of one solver stage.

Same constructor and attributes as ``simudo/physics/problem_data.py:12-60``:
``ProblemData(goal=..., mesh_data=..., unit_registry=...)`` with ``goal`` in
``'local charge neutrality' | 'thermal equilibrium' | 'full'``; ``.pdd`` and
``.optical`` are created on first access.
"""
from ..util import SetattrInitMixin

__all__ = ['ProblemData']


class ProblemData(SetattrInitMixin):
    goal = 'full'
    mesh_data = None
    unit_registry = None
    backend_factory = None     # None -> CUDA backend (simudo_b200.fem.backend)

    @property
    def pdd(self):
        if getattr(self, '_pdd', None) is None:
            from .poisson_drift_diffusion import PoissonDriftDiffusion
            self._pdd = PoissonDriftDiffusion(problem_data=self)
        return self._pdd

    @property
    def optical(self):
        if getattr(self, '_optical', None) is None:
            from .optical import Optical
            self._optical = Optical(problem_data=self)
        return self._optical

    def optical_update(self, pdd=None):
        """Solve every optical field for the current PDD state (the body of
        ``run_optical_subsolvers``, ``physics/steppers.py:120-136``)."""
        (pdd or self.pdd).system.optical_update()

    @property
    def goal_abbreviated(self):
        return {'local charge neutrality': 'ntrl', 'thermal equilibrium': 'thmq',
                'full': 'full'}[self.goal]
