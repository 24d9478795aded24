"""Blackbody photon-flux helper (host side).

Same quantities as ``simudo/util/blackbody.py:10-142``: Planck spectral
radiance integrated over a photon-energy window, scaled by the sun/earth
geometric factors, used for the optical Dirichlet data
(``example/fourlayer/fourlayer.py:519-535``).
"""
import numpy as np
from scipy.integrate import quad

__all__ = ['Blackbody']


class Blackbody(object):
    earth_sun_distance = 149.6e6
    sun_radius = 695508

    def __init__(self, temperature):
        self.unit_registry = temperature._REGISTRY
        self.temperature = temperature

    @property
    def u(self):
        return self.unit_registry

    @property
    def geometric_factor(self):
        # half-space emission (2 pi) times the cosine-weighted factor 1/2
        return 2 * np.pi * 0.5

    @property
    def distance_factor(self):
        return (self.sun_radius / self.earth_sun_distance) ** 2

    def _kT_eV(self):
        return (self.temperature * self.u.boltzmann_constant).m_as('eV')

    def _prefactor(self):
        u = self.u
        return (2 * u.eV ** 3 / u.h ** 3 / u.c ** 2).to('1/m**2/s')

    def photon_flux_integral_on_earth(self, E0, E1):
        """Photon flux in [E0, E1] from a black sun seen from Earth, normal incidence."""
        T = self._kT_eV()
        r = quad(lambda E: E ** 2 / np.expm1(E / T), E0.m_as('eV'), E1.m_as('eV'))[0]
        fac = self._prefactor() * self.distance_factor * self.geometric_factor
        return (r * fac).to('1/cm^2/s')

    def total_intensity_on_earth(self):
        T = self._kT_eV()
        r = quad(lambda E: E ** 3 / np.expm1(E / T), 0, 20.)[0]
        fac = self._prefactor() * self.distance_factor * self.geometric_factor
        return (r * fac * self.u.eV).to('W/m^2')

    def non_overlapping_energy_ranges(self, energies, inf=None):
        """{key: lower} -> {key: (lower, next_higher_lower_or_inf)}."""
        if inf is None:
            inf = 20 * self.u.eV
        Es = sorted(energies.items(), key=lambda kv: kv[1].m_as('eV'), reverse=True)
        result = {}
        prev = inf
        for k, E in Es:
            result[k] = (E, prev)
            prev = E
        return result
