"""Region algebra over cell values and signed facet values.

API mirror of ``simudo/mesh/topology.py``: ``CellRegions`` / ``FacetRegions``
are auto-vivifying name -> region containers; regions compose with
``& | ^ -`` and cell regions produce facet regions through
``.boundary(other)``, ``.internal_facets()``,
``.subdomain_internal_facets()``; facet regions support ``.flip()``,
``.both()``, ``.unsigned()``.  ``region.evaluate(context)`` returns a set of
cell values, or a set of ``(facet_value, sign)`` pairs.  The expected set
identities are those checked by the reference's ``test/test_topology.py``.
"""
import operator
from functools import reduce

__all__ = ['CellRegion', 'FacetRegion', 'CellRegions', 'FacetRegions']


class _Regions(object):
    _factory = None

    def __init__(self, mapping=None):
        object.__setattr__(self, '_mapping_', {} if mapping is None else mapping)

    def __getitem__(self, name):
        if isinstance(name, (set, frozenset)):
            return {k: self[k] for k in name}
        m = self._mapping_
        if name not in m:
            m[name] = type(self)._factory(name)
        return m[name]

    def __getattr__(self, name):
        return self[name]

    def __setitem__(self, name, value):
        self._mapping_[name] = value

    __setattr__ = __setitem__

    def __repr__(self):
        return '{}({!r})'.format(type(self).__name__, self._mapping_)


class _Region(object):
    """Expression node; equality/hash are structural."""
    _fields = ()

    def _key(self):
        return (type(self),) + tuple(getattr(self, f) for f in self._fields)

    def __hash__(self):
        return hash(self._key())

    def __eq__(self, other):
        return isinstance(other, _Region) and self._key() == other._key()

    def evaluate(self, context):
        raise NotImplementedError()


class CellRegion(_Region):
    def __new__(cls, *args, **kwargs):
        if cls is CellRegion:
            return object.__new__(_NamedCellRegion)
        return object.__new__(cls)

    def __and__(self, x): return _CellOp(operator.and_, '&', (self, x))
    def __or__(self, x): return _CellOp(operator.or_, '|', (self, x))
    def __xor__(self, x): return _CellOp(operator.xor, '^', (self, x))
    def __sub__(self, x): return _CellOp(operator.sub, '-', (self, x))

    def boundary(self, x):
        """Signed boundary between ``self`` and ``x``."""
        return _CellsToFacets('boundary', (self, x))

    def internal_facets(self):
        return _CellsToFacets('internal', (self,))

    def subdomain_internal_facets(self):
        return _CellsToFacets('cell_value_internal', (self,))

    @staticmethod
    def null():
        return _NullCellRegion()


class _NamedCellRegion(CellRegion):
    _fields = ('name',)

    def __init__(self, name):
        self.name = name

    def evaluate(self, context):
        return set(context['region_name_to_cvs'][self.name])

    def __repr__(self):
        return 'c{!r}'.format(self.name)


class _NullCellRegion(CellRegion):
    def evaluate(self, context):
        return set()

    def __repr__(self):
        return 'null'


def _check(operands, base):
    if not all(isinstance(o, base) for o in operands):
        raise TypeError('operands must be {} instances'.format(base.__name__))


class _CellOp(CellRegion):
    _fields = ('symbol', 'operands')

    def __init__(self, func, symbol, operands):
        _check(operands, CellRegion)
        self.func, self.symbol, self.operands = func, symbol, tuple(operands)

    def evaluate(self, context):
        return reduce(self.func, (o.evaluate(context) for o in self.operands))

    def __repr__(self):
        return '({} {})'.format(self.symbol, ' '.join(map(repr, self.operands)))


class FacetRegion(_Region):
    def __new__(cls, *args, **kwargs):
        if cls is FacetRegion:
            return object.__new__(_NamedFacetRegion)
        return object.__new__(cls)

    def __and__(self, x): return _FacetOp(operator.and_, '&', (self, x))
    def __or__(self, x): return _FacetOp(operator.or_, '|', (self, x))
    def __xor__(self, x): return _FacetOp(operator.xor, '^', (self, x))
    def __sub__(self, x): return _FacetOp(operator.sub, '-', (self, x))

    def flip(self):
        """Invert signs: ``X.boundary(Y) == Y.boundary(X).flip()``."""
        return _FacetMap('flip', (self,))

    def both(self):
        """Both orientations: ``f.both() == f | f.flip()``."""
        return _FacetMap('both', (self,))

    def unsigned(self):
        return _FacetMap('unsigned', (self,))

    @staticmethod
    def null():
        return _NullFacetRegion()


class _NamedFacetRegion(FacetRegion):
    _fields = ('name',)

    def __init__(self, name):
        self.name = name

    def evaluate(self, context):
        return set(context['facets_name_to_fvs'][self.name])

    def __repr__(self):
        return 'f{!r}'.format(self.name)


class _NullFacetRegion(FacetRegion):
    def evaluate(self, context):
        return set()

    def __repr__(self):
        return 'null'


class _FacetOp(FacetRegion):
    _fields = ('symbol', 'operands')

    def __init__(self, func, symbol, operands):
        _check(operands, FacetRegion)
        self.func, self.symbol, self.operands = func, symbol, tuple(operands)

    def evaluate(self, context):
        return reduce(self.func, (o.evaluate(context) for o in self.operands))

    def __repr__(self):
        return '({} {})'.format(self.symbol, ' '.join(map(repr, self.operands)))


class _FacetMap(FacetRegion):
    _fields = ('kind', 'operands')

    def __init__(self, kind, operands):
        _check(operands, FacetRegion)
        self.kind, self.operands = kind, tuple(operands)

    def evaluate(self, context):
        fvs = self.operands[0].evaluate(context)
        if self.kind == 'flip':
            return set((fv, -s) for fv, s in fvs)
        if self.kind == 'both':
            return set((fv, s) for fv, _ in fvs for s in (-1, 1))
        return set((fv, 1) for fv, _ in fvs)

    def __repr__(self):
        return '({} {!r})'.format(self.kind, self.operands[0])


class _CellsToFacets(FacetRegion):
    _fields = ('method', 'operands')

    def __init__(self, method, operands):
        _check(operands, CellRegion)
        self.method, self.operands = method, tuple(operands)

    def evaluate(self, context):
        args = [o.evaluate(context) for o in self.operands]
        return getattr(context['facets_manager'], self.method)(*args)

    def __repr__(self):
        return '({} {})'.format(self.method, ' '.join(map(repr, self.operands)))


class CellRegions(_Regions):
    _factory = CellRegion


class FacetRegions(_Regions):
    _factory = FacetRegion
