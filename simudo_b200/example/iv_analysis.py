"""J-V curve figures of merit (mirror of ``simudo/example/fourlayer/IV_Analysis.py:15-73``).

``IV_params(i, v, powersign)``: ``isc`` and ``voc`` by linear interpolation, the maximum
power point from a quartic interpolating spline of ``powersign * i * v`` (roots of its
derivative), ``ff = pmax / (voc * isc)``.  ``efficiency(jv, X, Ts)`` divides ``pmax`` by the
incident power ``X * Blackbody(Ts).total_intensity_on_earth()`` like
``fourlayer_example.py:329-363``.
"""
import numpy as np
from scipy.interpolate import UnivariateSpline, interp1d

__all__ = ['IV_params', 'efficiency']


class IV_params(object):
    def __init__(self, i, v, powersign=1.0):
        i, v = np.asarray(i, dtype=float), np.asarray(v, dtype=float)
        if len(i) != len(v):
            raise ValueError('unequal length arrays')
        if v[0] > v[-1]:
            i, v = i[::-1], v[::-1]
        if not (np.diff(v) > 0).all():
            raise ValueError('voltage array must be monotonic')
        if not (i.max() >= 0.0 >= i.min() and i.max() > i.min()):
            raise ValueError('current array does not cross zero')
        self.i, self.v = i, v
        self.isc = float(interp1d(v, i, kind='linear')(0.0))
        order = np.argsort(i)
        self.voc = float(interp1d(i[order], v[order], kind='linear', bounds_error=False)(0.0))
        spline = UnivariateSpline(v, powersign * i * v, k=4, s=0)
        best = None
        for r in spline.derivative().roots():
            p = float(spline(r))
            if best is None or p > best[0]:
                best = (p, float(r))
        if best is None:
            raise ValueError('no maximum power point inside the voltage range')
        self.pmax, self.mppv = best
        self.mppi = float(interp1d(v, i, kind='linear')(self.mppv))
        self.ff = self.pmax / (self.voc * self.isc)


def efficiency(jv, X=100.0, Ts=6000.0, powersign=-1.0):
    """``jv``: [(V [V], J [mA/cm^2])]; power-producing current is J < 0 (``powersign = -1``)."""
    from ..util import Blackbody, make_unit_registry
    U = make_unit_registry()
    p_in = (X * Blackbody(temperature=Ts * U.K).total_intensity_on_earth()).to('mW/cm^2').magnitude
    v = [a for a, _ in jv]
    j = [b for _, b in jv]
    return IV_params(j, v, powersign).pmax / p_in
