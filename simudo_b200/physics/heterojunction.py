"""Thermionic-emission heterojunction boundary condition: NOT BUILT on the device path.

``ThermionicHeterojunction(band, boundary)`` keeps the reference's constructor
(``simudo/physics/heterojunction.py:11-47``) so that a problem definition that uses it
fails by name instead of by ``ImportError``.  ``register()`` (reference ``:49-135``) would
(1) take ``boundary`` out of the band's ``base_w`` jump term and (2) add the nonlinear
interior-facet term ``-dS (Delta_w - Delta_w_BC(j.n, u_bar)) xi.n``; the assembly kernels
have no facet-quadrature stage for (2), and there is no CPU
fallback, so it raises.
"""

__all__ = ['ThermionicHeterojunction']


class ThermionicHeterojunction:
    def __init__(self, band, boundary):
        self.band = band
        self.boundary = boundary

    @property
    def vth(self):
        return self.band.spatial.get(self.band.spatial_prefix + 'vth')

    def register(self):
        raise NotImplementedError(
            'ThermionicHeterojunction: the thermionic interior-facet term '
            '(reference physics/heterojunction.py:49-135) is not built in the CUDA '
            'assembly path, see DESIGN.md section 7')
