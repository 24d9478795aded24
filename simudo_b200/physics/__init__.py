from .problem_data import ProblemData
from .poisson_drift_diffusion import (
    SaveItem, GetToSaveMixin, PoissonDriftDiffusion, Poisson, MixedPoisson,
    LocalChargeNeutralityPoisson, Band, NondegenerateBand, IntermediateBand,
    MixedQflNondegenerateBand, MixedQflIntermediateBand, MixedQflBandMixin)
from .electro_optical_process import (
    ElectroOpticalProcess, SRHRecombination, NonRadiativeTrap,
    NonOverlappingTopHatBeerLambert, NonOverlappingTopHatBeerLambertIB)
from .material import Material, SimpleSiliconMaterial, BadSimpleSiliconMaterial
from .optical import Optical, OpticalField, AbsorptionRangesHelper
from .steppers import (VoltageStepper, OpticalIntensityAdaptiveStepper,
                       NonequilibriumCoupledStepper, NewtonBailoutException)
from .heterojunction import ThermionicHeterojunction
