"""Generation / recombination processes (declarative).

Class names, constructor keywords and spatial keys follow
``simudo/physics/electro_optical_process.py``; instead of returning UFL
generation expressions each class states which kernel process kind it lowers
to (``lowering.py`` fills the per-tag parameter table, the CUDA kernels
evaluate g and its derivatives, ``csrc/pdd_device.cuh: sb_generation``):

=================================  =====================  ====================
class                              reference lines        kernel kind
=================================  =====================  ====================
SRHRecombination                   :495-536               SB_PROC_SRH
NonRadiativeTrap                   :539-562               SB_PROC_SR_TRAP
NonOverlappingTopHatBeerLambert    :398-431               SB_PROC_RAD_BB
NonOverlappingTopHatBeerLambertIB  :434-492               SB_PROC_RAD_IB
=================================  =====================  ====================
"""
from ..fem import model as M
from ..util import SetattrInitMixin

__all__ = ['ElectroOpticalProcess', 'SRHRecombination', 'NonRadiativeTrap',
           'NonOverlappingTopHatBeerLambert', 'NonOverlappingTopHatBeerLambertIB',
           'TwoBandEOPMixin', 'TrapEOPMixin']


class ElectroOpticalProcess(SetattrInitMixin):
    pdd = None
    optical = None
    name = None
    kind = None
    pre_iteration_hook = None
    pre_first_iteration_hook = None
    post_iteration_hook = None

    @property
    def unit_registry(self):
        return self.pdd.unit_registry

    def get_quantum_yield(self, photon_energy):
        return 1.0


class TwoBandEOPMixin(object):
    """Destination band gains carriers as the process intensifies
    (``:193-205``)."""
    src_band = None
    dst_band = None

    def get_band_generation_sign(self, band):
        if band is self.dst_band:
            return +1
        if band is self.src_band:
            return -(self.dst_band.sign * self.src_band.sign)
        return None


class TrapEOPMixin(object):
    trap_band = None

    @classmethod
    def easy_add_two_traps_to_pdd(cls, pdd, name_prefix, top_band, bottom_band,
                                  trap_band, **kwargs):
        """``name_top``: trap <-> top band, ``name_bottom``: bottom band <-> trap
        (``:214-228``)."""
        pdd.easy_add_electro_optical_process(
            cls, dst_band=top_band, src_band=trap_band, trap_band=trap_band,
            name=name_prefix + '_top', **kwargs)
        pdd.easy_add_electro_optical_process(
            cls, dst_band=trap_band, src_band=bottom_band, trap_band=trap_band,
            name=name_prefix + '_bottom', **kwargs)

    @property
    def reg_band(self):
        if self.dst_band is self.trap_band:
            return self.src_band
        if self.src_band is self.trap_band:
            return self.dst_band
        return None


class SRHRecombination(TrapEOPMixin, TwoBandEOPMixin, ElectroOpticalProcess):
    """dst_band (typically CB) loses carriers with src_band (VB) through a
    mid-gap level: keys ``SRH/<band>/tau``, ``SRH/energy_level``."""
    name = 'SRH'
    kind = M.PROC_SRH
    tau_from_spatial = True


class NonRadiativeTrap(TrapEOPMixin, TwoBandEOPMixin, ElectroOpticalProcess):
    """Shockley-Read trapping into an explicit trap band: keys
    ``<name>/sigma_th``, ``<name>/v_th``."""
    name = 'nonradiative'
    kind = M.PROC_SR_TRAP


class NonOverlappingTopHatBeerLambert(TwoBandEOPMixin, ElectroOpticalProcess):
    """Band-to-band absorption + radiative recombination: key ``<name>/alpha``."""
    name = 'beer_lambert'
    kind = M.PROC_RAD_BB
    enable_radiative_recombination = True


class NonOverlappingTopHatBeerLambertIB(TrapEOPMixin, TwoBandEOPMixin,
                                        ElectroOpticalProcess):
    """IB <-> band absorption + radiative trapping: key ``<name>/sigma_opt``."""
    name = 'beer_lambert_IB'
    kind = M.PROC_RAD_IB
    enable_radiative_recombination = True
