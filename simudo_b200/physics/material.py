"""Materials: parameter dictionaries registered as spatial rules.

Same API as ``simudo/physics/material.py:11-106`` (``Material.get_dict``,
``.register``); ``SimpleSiliconMaterial`` carries the reference's silicon
numbers (``:76-89``).
"""
from ..util import SetattrInitMixin

__all__ = ['Material', 'SimpleSiliconMaterial', 'BadSimpleSiliconMaterial']


class Material(SetattrInitMixin):
    name = None
    problem_data = None

    def get_dict(self):
        return {}

    @property
    def dict(self):
        if getattr(self, '_dict', None) is None:
            self._dict = self.get_dict()
        return self._dict

    @property
    def pdd(self):
        return self.problem_data.pdd

    @property
    def temperature(self):
        return self.pdd.temperature

    @property
    def unit_registry(self):
        return self.problem_data.unit_registry

    def register(self, name=None):
        if name is None:
            name = self.name
        for k, v in self.dict.items():
            self.pdd.spatial.add_material_data(key=k, material_name=name, value=v)


class SimpleSiliconMaterial(Material):
    """Very simple silicon, no temperature dependence."""
    name = 'Si'

    def get_dict(self):
        d = super().get_dict()
        U = self.unit_registry
        d.update({
            'CB/energy_level': U('1.12 eV'),
            'VB/energy_level': U('0 eV'),
            'CB/effective_density_of_states': U('3.2e19 1/cm^3'),
            'VB/effective_density_of_states': U('1.8e19 1/cm^3'),
            'CB/mobility': U('1400 cm^2/V/s'),
            'VB/mobility': U(' 450 cm^2/V/s'),
            'poisson/permittivity': U('11.7 vacuum_permittivity'),
        })
        return d


class BadSimpleSiliconMaterial(SimpleSiliconMaterial):
    name = 'badSi'

    def get_dict(self):
        d = super().get_dict()
        U = self.unit_registry
        d['CB/mobility'] = d['CB/mobility'] / 3
        d['VB/mobility'] = U('233 cm^2/V/s')
        return d
