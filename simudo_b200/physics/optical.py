"""Optical fields feeding the generation term.

``Optical`` / ``OpticalField`` keep the reference's surface
(``simudo/physics/optical.py:29-86``: ``easy_add_field``, ``Phi_scale``,
``spatial``) and hold each field's photon flux on the PDD mesh as P2 nodal
values per cell (``Phi_pddproj``, ``:205-208``), which is what the assembly
kernel reads.  ROUND-1 SCOPE: the flux is the Beer-Lambert solution along the
propagation direction for the *current* absorption profile on Xs x Ys product
meshes with direction (+-1, 0) (every shipped example); the CG2 MSORTE
discretisation (``optical1.py1:27-45``) is SURVEY 8(a10) and not built yet.
"""
import numpy as np

from ..fem.expr import Constant
from ..fem.spatial import Spatial
from ..util import NameDict, SetattrInitMixin

__all__ = ['Optical', 'OpticalField', 'AbsorptionRangesHelper']


class OpticalField(SetattrInitMixin):
    optical = None
    key = None
    photon_energy = None
    direction = None
    solid_angle = 4 * np.pi
    Phi_cells = None           # [Nc, 6] photon flux, 1/(mesh_unit^2 s)

    def __init__(self, **kwargs):
        d = kwargs.get('direction')
        if d is not None:
            kwargs['direction'] = np.array(d, dtype=float)
        super().__init__(**kwargs)

    @property
    def name(self):
        return self.key

    def __str__(self):
        return self.key


class Optical(SetattrInitMixin):
    problem_data = None

    def __init__(self, **kwargs):
        super().__init__(**kwargs)
        self.fields = NameDict()
        self.spatial = Spatial(parent=self)
        U = self.unit_registry
        self.Phi_scale = Constant(1.0) * U.dimensionless

    @property
    def mesh_data(self):
        return self.problem_data.mesh_data

    @property
    def unit_registry(self):
        return self.problem_data.unit_registry

    def easy_add_field(self, key, photon_energy, direction, solid_angle=None):
        if solid_angle is None:
            solid_angle = 4 * np.pi
        self.fields.add(OpticalField(optical=self, key=key, photon_energy=photon_energy,
                                     direction=direction, solid_angle=solid_angle))

    def get_to_save(self):
        return NameDict()


class AbsorptionRangesHelper(SetattrInitMixin):
    """Non-overlapping absorption windows (``optical.py:300-374``): transition
    a<->b absorbs photons in ``[|E_a - E_b|, next higher transition energy)``;
    upper bound 100 eV when there is none."""
    inf_upper_bound_eV = 100.0

    @staticmethod
    def transition_bounds(levels):
        """levels: {band key: energy [eV]} -> {frozenset(a, b): (lower, upper)}."""
        keys = list(levels)
        lowers = {}
        for i in range(len(keys)):
            for j in range(i + 1, len(keys)):
                lowers[frozenset((keys[i], keys[j]))] = abs(levels[keys[i]] - levels[keys[j]])
        out = {}
        for k, lo in lowers.items():
            cands = [(100.0 if lo2 < lo else lo2) for k2, lo2 in lowers.items() if k2 != k]
            out[k] = (lo, min(cands) if cands else 100.0)
        return out
