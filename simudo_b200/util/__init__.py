from .units import (Unit, Quantity, UnitRegistry, make_unit_registry,
                    DimensionalityError)
from .name_dict import NameDict
from .setattr_init_mixin import SetattrInitMixin
from .blackbody import Blackbody

__all__ = ['Unit', 'Quantity', 'UnitRegistry', 'make_unit_registry',
           'DimensionalityError', 'NameDict', 'SetattrInitMixin', 'Blackbody']
