"""Tiny lazy-expression layer for boundary-condition and rule values.

The reference lets users write BC values as UFL expressions over solver
objects -- ``phi0 - V_ext``, ``U('A/cm^2') * mu.zerovec``,
``VB.thermal_equilibrium_u``, ``problem0.pdd.poisson.phi``
(``example/pn_diode/pn_diode.py:127-159,172-177``).  Here those lower to a
linear combination of three kinds of leaves that can be evaluated at points of
mesh facets on the host (a few facets per contact): ``Constant`` (mutable
scalar, stands in for ``dolfin.Constant``), ``CellFunction`` (cell-wise
polynomial, DG0/DG1/DG2) and plain floats.  Nothing here is UFL.
"""
import numpy as np

__all__ = ['Expr', 'Constant', 'CellFunction', 'ZeroVec', 'UnitVec', 'ThermalEquilibriumU',
           'as_expr', 'p2_from_p1']


class Expr(object):
    """Linear combination  sum_k coef_k * leaf_k  (+ float offset)."""
    __array_priority__ = 2000

    def __init__(self, terms=(), offset=0.0):
        self.terms = list(terms)          # [(coef, leaf)]
        self.offset = float(offset)

    # -- algebra (linear only) ---------------------------------------------------
    def _lin(self):
        return self.terms, self.offset

    def __add__(self, other):
        o = as_expr(other)
        t1, c1 = self._lin()
        t2, c2 = o._lin()
        return Expr(list(t1) + list(t2), c1 + c2)

    __radd__ = __add__

    def __neg__(self):
        t, c = self._lin()
        return Expr([(-a, l) for a, l in t], -c)

    def __sub__(self, other):
        return self + (-as_expr(other))

    def __rsub__(self, other):
        return as_expr(other) + (-self)

    def __mul__(self, k):
        if isinstance(k, Expr):
            raise TypeError('only linear expressions are supported')
        if hasattr(k, 'units'):
            return NotImplemented
        t, c = self._lin()
        return Expr([(a * float(k), l) for a, l in t], c * float(k))

    __rmul__ = __mul__

    def __truediv__(self, k):
        return self * (1.0 / float(k))

    # -- evaluation ----------------------------------------------------------------
    def eval_points(self, cells, ref_points):
        """Values at reference points ``ref_points[n, q, 2]`` of ``cells[n]``."""
        out = np.full(ref_points.shape[:2], self.offset, dtype=np.float64)
        for a, leaf in self.terms:
            out += a * leaf._eval_leaf(cells, ref_points)
        return out

    def is_zero(self):
        return self.offset == 0.0 and all(a == 0.0 or leaf.is_zero() for a, leaf in self.terms)

    def constants(self):
        return [l for _, l in self.terms if isinstance(l, Constant)]


class _Leaf(Expr):
    offset = 0.0

    def __init__(self):
        pass

    @property
    def terms(self):
        return [(1.0, self)]

    def _lin(self):
        return [(1.0, self)], 0.0

    def eval_points(self, cells, ref_points):
        return self._eval_leaf(cells, ref_points)

    def is_zero(self):
        return False


class Constant(_Leaf):
    """Mutable scalar; ``assign`` mirrors ``dolfin.Constant.assign`` as used by
    the steppers (``fem/adaptive_stepper.py:331-334``)."""

    def __init__(self, value=0.0):
        self._value = float(value)

    def assign(self, value):
        self._value = float(value)

    def __float__(self):
        return self._value

    def values(self):
        return np.array([self._value])

    def _eval_leaf(self, cells, ref_points):
        return np.full(ref_points.shape[:2], self._value)

    def __repr__(self):
        return 'Constant({!r})'.format(self._value)


class ZeroVec(_Leaf):
    """``mesh_util.zerovec``: the zero vector field (essential flux BCs)."""
    is_vector = True

    def _eval_leaf(self, cells, ref_points):
        return np.zeros(ref_points.shape[:2])

    def is_zero(self):
        return True


class UnitVec(_Leaf):
    """``mesh_util.unitvec_x`` / ``unitvec_y`` (``fem/mesh_util1.py1:160-161``): constant unit
    vector fields, for non-zero essential flux values.  ``vector_components`` turns a linear
    combination of them into (vx, vy)."""
    is_vector = True

    def __init__(self, axis):
        self.axis = int(axis)

    def _eval_leaf(self, cells, ref_points):
        raise TypeError('a vector field has no scalar value; use vector_components')


def vector_components(expr):
    """(vx, vy) of a linear combination of ZeroVec / UnitVec leaves (constant vector field)."""
    v = [0.0, 0.0]
    e = as_expr(expr)
    if e.offset != 0.0:
        raise TypeError('scalar offset in a vector-valued boundary value')
    for a, leaf in e.terms:
        if isinstance(leaf, ZeroVec):
            continue
        if isinstance(leaf, UnitVec):
            v[leaf.axis] += a
        elif isinstance(leaf, Constant):
            raise TypeError('scalar Constant in a vector-valued boundary value')
        else:
            raise NotImplementedError('essential flux values must be constant vectors '
                                      '(zerovec, unitvec_x, unitvec_y)')
    return v[0], v[1]


def _p1_basis(X):
    return np.stack([1.0 - X[..., 0] - X[..., 1], X[..., 0], X[..., 1]], axis=-1)


def _p2_basis(X):
    l = _p1_basis(X)
    return np.stack([l[..., 0] * (2 * l[..., 0] - 1), l[..., 1] * (2 * l[..., 1] - 1),
                     l[..., 2] * (2 * l[..., 2] - 1), 4 * l[..., 1] * l[..., 2],
                     4 * l[..., 0] * l[..., 2], 4 * l[..., 0] * l[..., 1]], axis=-1)


def p2_from_p1(v):
    """P2 nodal values (3 vertices + 3 edge midpoints) of a P1 function."""
    mid = 0.5 * (v[:, [1, 0, 0]] + v[:, [2, 2, 1]])
    return np.concatenate([v, mid], axis=1)


class CellFunction(_Leaf):
    """Cell-wise polynomial scalar field: ``values[Nc, k]`` with k = 1 (DG0),
    3 (DG1, vertex values) or 6 (DG2, vertices + edge midpoints).  ``getter``
    lets the field track a live solution vector."""

    def __init__(self, values=None, getter=None, name=None):
        self._values = values
        self._getter = getter
        self.name = name

    @property
    def values(self):
        return self._getter() if self._getter is not None else self._values

    def _eval_leaf(self, cells, ref_points):
        v = np.asarray(self.values)[cells]
        k = v.shape[1]
        if k == 1:
            return np.broadcast_to(v, ref_points.shape[:2]).copy()
        basis = _p1_basis(ref_points) if k == 3 else _p2_basis(ref_points)
        return np.einsum('nk,nqk->nq', v, basis)


class ThermalEquilibriumU(_Leaf):
    """``band.thermal_equilibrium_u`` used as an Ohmic-contact value
    (``poisson_drift_diffusion1.py1:145-146``).  It never needs to be evaluated
    as a density: the mixed-qfl method only uses ``u_to_w`` of it, which is
    ``e (phi_thmeq - phi)`` (``:362-367``)."""

    def __init__(self, band):
        self.band = band

    def _eval_leaf(self, cells, ref_points):
        raise TypeError('thermal_equilibrium_u is only meaningful as a contact value')


def as_expr(x):
    if isinstance(x, Expr):
        return x
    return Expr([], float(x))
