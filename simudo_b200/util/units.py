"""Minimal fixed-dimension unit system replacing the reference's pint registry.

The reference builds a ``pint.UnitRegistry`` from bundled definition files
(``simudo/util/pint.py:24-30``, ``simudo/util/pint/default_en.txt``,
``constants_en.txt``).  pint is not available here, so this module provides
the small subset of the pint surface that Simudo scripts and the physics
layer actually touch ("API leaks"):

* ``make_unit_registry(("mesh_unit = 1 micrometer",))``
* ``U('300 K')``, ``U('1400 cm^2/V/s')``, ``U.V``, ``U.eV``, ``U.mesh_unit``
* ``Quantity`` arithmetic, ``.to()``, ``.m_as()``, ``.magnitude/.m``,
  ``.units/.u``, ``._REGISTRY``, ``U.Quantity(x, 'unit')``

Numerical constants are the reference's own (``default_en.txt:88``:
e = 1.602176487e-19 C; ``constants_en.txt:8,10,11,15,34``).

Magnitudes may be floats, numpy arrays or lazy ``Expr`` objects
(``simudo_b200.fem.expr``): every conversion is a plain multiplication.
"""
import ast
import math
import re
from fractions import Fraction

__all__ = ['Unit', 'Quantity', 'UnitRegistry', 'make_unit_registry',
           'DimensionalityError']

# base dimensions: length, mass, time, current, temperature
_NDIM = 5
_ZERO = (Fraction(0),) * _NDIM


class DimensionalityError(ValueError):
    pass


def _dim(**kw):
    names = ('length', 'mass', 'time', 'current', 'temperature')
    return tuple(Fraction(kw.get(n, 0)) for n in names)


class Unit(object):
    """A unit = SI scale factor x dimension vector (plus a display name)."""
    __slots__ = ('factor', 'dims', 'name', '_REGISTRY')
    __array_priority__ = 1000

    def __init__(self, factor, dims, name=None, registry=None):
        self.factor = float(factor)
        self.dims = tuple(dims)
        self.name = name
        self._REGISTRY = registry

    # -- algebra ---------------------------------------------------------
    def _new(self, factor, dims, name):
        return Unit(factor, dims, name, self._REGISTRY)

    def __mul__(self, other):
        if isinstance(other, Unit):
            return self._new(self.factor * other.factor,
                             tuple(a + b for a, b in zip(self.dims, other.dims)),
                             '({})*({})'.format(self, other))
        if isinstance(other, Quantity):
            return Quantity(other.magnitude, self * other.units)
        return Quantity(other, self)

    __rmul__ = __mul__

    def __truediv__(self, other):
        if isinstance(other, Unit):
            return self._new(self.factor / other.factor,
                             tuple(a - b for a, b in zip(self.dims, other.dims)),
                             '({})/({})'.format(self, other))
        if isinstance(other, Quantity):
            return Quantity(1.0 / other.magnitude, self / other.units)
        return Quantity(1.0 / other, self)

    def __rtruediv__(self, other):
        inv = self ** -1
        if isinstance(other, Quantity):
            return Quantity(other.magnitude, other.units * inv)
        return Quantity(other, inv)

    def __pow__(self, p):
        p = Fraction(p).limit_denominator(1000)
        return self._new(self.factor ** float(p),
                         tuple(a * p for a in self.dims),
                         '({})**{}'.format(self, p))

    def __eq__(self, other):
        return (isinstance(other, Unit) and self.dims == other.dims
                and math.isclose(self.factor, other.factor, rel_tol=1e-14))

    def __hash__(self):
        return hash(self.dims)

    def __repr__(self):
        return self.name if self.name is not None else \
            '<Unit {:g} {}>'.format(self.factor, self.dims)

    @property
    def dimensionless(self):
        return self.dims == _ZERO

    # pint compatibility: ``unit.units`` / ``unit.magnitude``
    @property
    def units(self):
        return self

    u = units

    @property
    def magnitude(self):
        return 1.0

    m = magnitude

    def conversion_factor_to(self, other):
        if self.dims != other.dims:
            raise DimensionalityError(
                'cannot convert {!r} to {!r}'.format(self, other))
        return self.factor / other.factor


class Quantity(object):
    __slots__ = ('magnitude', 'units')
    __array_priority__ = 1000

    def __init__(self, magnitude, units):
        if isinstance(magnitude, Quantity):
            units = magnitude.units * units
            magnitude = magnitude.magnitude
        self.magnitude = magnitude
        self.units = units

    # -- pint-like accessors ------------------------------------------------
    @property
    def m(self):
        return self.magnitude

    @property
    def u(self):
        return self.units

    @property
    def _REGISTRY(self):
        return self.units._REGISTRY

    def _resolve(self, unit):
        if isinstance(unit, Quantity):
            if unit.magnitude != 1:
                return Unit(unit.units.factor * unit.magnitude, unit.units.dims,
                            None, unit.units._REGISTRY)
            unit = unit.units
        if isinstance(unit, str):
            q = self._REGISTRY(unit)
            unit = q.units if isinstance(q, Quantity) else q
        return unit

    def m_as(self, unit):
        unit = self._resolve(unit)
        f = self.units.conversion_factor_to(unit)
        return self.magnitude if f == 1.0 else self.magnitude * f

    def to(self, unit):
        unit = self._resolve(unit)
        return Quantity(self.m_as(unit), unit)

    # -- arithmetic -----------------------------------------------------------
    def __mul__(self, other):
        if isinstance(other, Quantity):
            return Quantity(self.magnitude * other.magnitude,
                            self.units * other.units)
        if isinstance(other, Unit):
            return Quantity(self.magnitude, self.units * other)
        return Quantity(self.magnitude * other, self.units)

    def __rmul__(self, other):
        if isinstance(other, Unit):
            return Quantity(self.magnitude, other * self.units)
        return Quantity(other * self.magnitude, self.units)

    def __truediv__(self, other):
        if isinstance(other, Quantity):
            return Quantity(self.magnitude / other.magnitude,
                            self.units / other.units)
        if isinstance(other, Unit):
            return Quantity(self.magnitude, self.units / other)
        return Quantity(self.magnitude / other, self.units)

    def __rtruediv__(self, other):
        return Quantity(other / self.magnitude, self.units ** -1)

    def __pow__(self, p):
        return Quantity(self.magnitude ** p, self.units ** p)

    def __neg__(self):
        return Quantity(-self.magnitude, self.units)

    def __pos__(self):
        return self

    def _other_m(self, other):
        if isinstance(other, (Quantity, Unit)):
            return other.units.conversion_factor_to(self.units), other.magnitude
        if self.units.dimensionless:
            return 1.0 / self.units.factor, other
        if _is_zero(other):
            return 1.0, other
        raise DimensionalityError(
            'cannot add dimensionless value to {!r}'.format(self.units))

    def __add__(self, other):
        f, m = self._other_m(other)
        return Quantity(self.magnitude + (m if f == 1.0 else m * f), self.units)

    def __radd__(self, other):
        f, m = self._other_m(other)
        return Quantity((m if f == 1.0 else m * f) + self.magnitude, self.units)

    def __sub__(self, other):
        f, m = self._other_m(other)
        return Quantity(self.magnitude - (m if f == 1.0 else m * f), self.units)

    def __rsub__(self, other):
        f, m = self._other_m(other)
        return Quantity((m if f == 1.0 else m * f) - self.magnitude, self.units)

    def __float__(self):
        return float(self.m_as(Unit(1.0, _ZERO, 'dimensionless',
                                    self.units._REGISTRY)))

    def __repr__(self):
        return '<Quantity({!r}, {!r})>'.format(self.magnitude, self.units)

    __str__ = __repr__

    def __format__(self, spec):
        return '{} {}'.format(format(self.magnitude, spec), self.units)

    # comparisons (plain numbers only)
    def _cmp_m(self, other):
        f, m = self._other_m(other)
        return m * f

    def __lt__(self, o): return self.magnitude < self._cmp_m(o)
    def __le__(self, o): return self.magnitude <= self._cmp_m(o)
    def __gt__(self, o): return self.magnitude > self._cmp_m(o)
    def __ge__(self, o): return self.magnitude >= self._cmp_m(o)


def _is_zero(x):
    try:
        return x == 0
    except Exception:
        return False


_PREFIXES = [
    ('yocto', 1e-24), ('zepto', 1e-21), ('atto', 1e-18), ('femto', 1e-15),
    ('pico', 1e-12), ('nano', 1e-9), ('micro', 1e-6), ('milli', 1e-3),
    ('centi', 1e-2), ('deci', 1e-1), ('deca', 1e1), ('hecto', 1e2),
    ('kilo', 1e3), ('mega', 1e6), ('giga', 1e9), ('tera', 1e12),
    ('da', 1e1), ('y', 1e-24), ('z', 1e-21), ('a', 1e-18), ('f', 1e-15),
    ('p', 1e-12), ('n', 1e-9), ('u', 1e-6), ('µ', 1e-6), ('m', 1e-3),
    ('c', 1e-2), ('d', 1e-1), ('h', 1e2), ('k', 1e3), ('M', 1e6),
    ('G', 1e9), ('T', 1e12),
]

_NUM_RE = re.compile(r'(?<![A-Za-z_])(\d+\.?\d*(?:[eE][+-]?\d+)?|\.\d+(?:[eE][+-]?\d+)?)')


class UnitRegistry(object):
    """Callable registry: ``U('1.12 eV')`` parses a quantity string."""

    def __init__(self):
        self._units = {}
        self._cache = {}
        d = self._define_raw
        # SI base (``default_en.txt:45-51``)
        d('meter', 1.0, _dim(length=1), 'm', 'metre')
        d('gram', 1e-3, _dim(mass=1), 'g')
        d('second', 1.0, _dim(time=1), 's', 'sec')
        d('ampere', 1.0, _dim(current=1), 'A', 'amp')
        d('kelvin', 1.0, _dim(temperature=1), 'K', 'degK')
        d('dimensionless', 1.0, _ZERO)
        d('radian', 1.0, _ZERO, 'rad')
        d('count', 1.0, _ZERO)
        # derived
        class _L(object):
            __getitem__ = staticmethod(self._lookup)
        U = _L()
        self.define_unit('newton', U['kilogram'] * U['meter'] / U['second'] ** 2, 'N')
        self.define_unit('joule', U['newton'] * U['meter'], 'J')
        self.define_unit('watt', U['joule'] / U['second'], 'W')
        self.define_unit('coulomb', U['ampere'] * U['second'], 'C')
        self.define_unit('volt', U['joule'] / U['coulomb'], 'V')
        self.define_unit('farad', U['coulomb'] / U['volt'], 'F')
        self.define_unit('ohm', U['volt'] / U['ampere'])
        self.define_unit('hertz', U['second'] ** -1, 'Hz')
        self.define_unit('minute', 60 * U['second'], 'min')
        self.define_unit('hour', 3600 * U['second'], 'hr')
        # constants (``constants_en.txt:7-15,34``; ``default_en.txt:88,108``)
        pi = 3.1415926535897932384626433832795028841971693993751
        self.define_unit('speed_of_light', 299792458 * U['meter'] / U['second'], 'c')
        self.define_unit('vacuum_permeability',
                         4 * pi * 1e-7 * U['newton'] / U['ampere'] ** 2,
                         'mu_0', 'magnetic_constant')
        self.define_unit('vacuum_permittivity',
                         1 / (U['mu_0'] * U['c'] ** 2),
                         'epsilon_0', 'electric_constant')
        self.define_unit('planck_constant', 6.62606957e-34 * U['joule'] * U['second'], 'h')
        self.define_unit('hbar', U['planck_constant'] / (2 * pi))
        self.define_unit('elementary_charge', 1.602176487e-19 * U['coulomb'], 'e')
        self.define_unit('electron_volt', U['elementary_charge'] * U['volt'], 'eV')
        self.define_unit('boltzmann_constant',
                         1.3806488e-23 * U['joule'] / U['kelvin'], 'k')
        self.define_unit('stefan_boltzmann_constant',
                         5.670373e-8 * U['watt'] / U['meter'] ** 2 / U['kelvin'] ** 4)
        self.define_unit('angstrom', 1e-10 * U['meter'])
        self.define_unit('micron', 1e-6 * U['meter'])

    # -- definitions ---------------------------------------------------------
    def _define_raw(self, name, factor, dims, *aliases):
        u = Unit(factor, dims, name, self)
        self._units[name] = u
        for a in aliases:
            self._units[a] = Unit(factor, dims, a, self)
        self._cache.clear()
        return u

    def define_unit(self, name, value, *aliases):
        if isinstance(value, Quantity):
            factor = value.units.factor * value.magnitude
            dims = value.units.dims
        else:
            factor, dims = value.factor, value.dims
        return self._define_raw(name, factor, dims, *aliases)

    def define(self, definition):
        """pint-style ``'mesh_unit = 1 micrometer'`` (``pn_diode.py:47``)."""
        parts = [p.strip() for p in definition.split('=')]
        name, value = parts[0], self(parts[1])
        self.define_unit(name, value, *parts[2:])

    # -- lookup ------------------------------------------------------------------
    def _lookup(self, name):
        u = self._units.get(name)
        if u is not None:
            return u
        if name.endswith('s') and name[:-1] in self._units and len(name) > 2:
            return self._units[name[:-1]]
        for prefix, scale in _PREFIXES:
            if name.startswith(prefix) and len(name) > len(prefix):
                rest = name[len(prefix):]
                base = self._units.get(rest)
                if base is None and rest.endswith('s') and len(rest) > 2:
                    base = self._units.get(rest[:-1])
                if base is not None:
                    long_prefix = len(prefix) > 2 or prefix == 'da'
                    long_base = len(base.name or rest) > 2 and rest == base.name
                    # pint pairs long prefixes with long names, short with symbols;
                    # we accept any unambiguous combination.
                    return Unit(base.factor * scale, base.dims, name, self)
        raise AttributeError('undefined unit {!r}'.format(name))

    def __getattr__(self, name):
        if name.startswith('_'):
            raise AttributeError(name)
        return self._lookup(name)

    def __getitem__(self, name):
        return self._lookup(name)

    def Quantity(self, magnitude, units=None):
        if units is None:
            units = self._units['dimensionless']
        elif isinstance(units, str):
            units = self(units)
        if isinstance(units, Quantity):
            return Quantity(magnitude * units.magnitude, units.units)
        return Quantity(magnitude, units)

    # -- parser --------------------------------------------------------------
    def __call__(self, text):
        return self.parse_expression(text)

    def parse_expression(self, text):
        if not isinstance(text, str):
            return self.Quantity(text)
        cached = self._cache.get(text)
        if cached is not None:
            return cached
        src = text.strip().replace('^', '**')
        # implicit multiplication: "300 K", "1e4 cm", ") cm", "3.2e19 1/cm^3"
        src = re.sub(r'(?<=[0-9A-Za-z_µ.)])\s+(?=[0-9A-Za-z_µ(.])', '*', src)
        try:
            tree = ast.parse(src, mode='eval')
        except SyntaxError as e:
            raise ValueError('cannot parse unit expression {!r}'.format(text)) from e
        val = self._eval(tree.body)
        if not isinstance(val, (Quantity, Unit)):
            val = self.Quantity(val)
        elif isinstance(val, Unit):
            pass
        self._cache[text] = val
        return val

    def _eval(self, node):
        if isinstance(node, ast.BinOp):
            l, r = self._eval(node.left), self._eval(node.right)
            if isinstance(node.op, ast.Mult):
                return l * r
            if isinstance(node.op, ast.Div):
                return l / r
            if isinstance(node.op, ast.Pow):
                return l ** r
            if isinstance(node.op, ast.Add):
                return l + r
            if isinstance(node.op, ast.Sub):
                return l - r
        elif isinstance(node, ast.UnaryOp):
            v = self._eval(node.operand)
            if isinstance(node.op, ast.USub):
                return -v
            if isinstance(node.op, ast.UAdd):
                return v
        elif isinstance(node, ast.Constant):
            return node.value
        elif isinstance(node, ast.Name):
            return self._lookup(node.id)
        elif isinstance(node, ast.Attribute):
            # tolerate the reference's stray "u.second" (optical1.py1:52)
            return self._lookup(node.attr)
        raise ValueError('unsupported unit expression node {!r}'.format(node))


def make_unit_registry(extra_definitions=()):
    """Mirror of ``simudo.util.pint.make_unit_registry`` (``util/pint.py:24``)."""
    registry = UnitRegistry()
    for definition in extra_definitions:
        registry.define(definition)
    return registry
