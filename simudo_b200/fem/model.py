"""ctypes mirror of the C-ABI input structs in ``include/simudo_b200.h``.

``SbModel`` is what ``PoissonDriftDiffusion`` + its bands and electro-optical
processes lower to (band statistics ``poisson_drift_diffusion1.py1:170-245``,
process kinds ``electro_optical_process.py``); the ``TP_*`` offsets index one
row of the per-cell-tag parameter table that replaces ``Spatial.get``'s
conditional tree (``fem/spatial.py:350-376``).
"""
import ctypes as C

import numpy as np

MAX_BANDS = 4
MAX_PROCS = 8
MAX_FIELDS = 4
MAX_QUAD_M = 11

BAND_NONDEGENERATE = 0
BAND_INTERMEDIATE = 1

PROC_SRH = 0
PROC_SR_TRAP = 1
PROC_RAD_BB = 2
PROC_RAD_IB = 3

TP_KT = 0
TP_EPS = 1
TP_RHO_STATIC = 2
TP_BAND0 = 3
TPB_IN_SUBDOMAIN = 0
TPB_N = 1
TPB_E = 2
TPB_MOBILITY = 3
TP_PROC0 = TP_BAND0 + 5 * MAX_BANDS
TPP_P0, TPP_P1, TPP_P2, TPP_P3 = 0, 1, 2, 3
TPP_FIELD0 = 4
TAG_STRIDE = TP_PROC0 + 8 * MAX_PROCS

# reference's unit constants in form units (util/pint/default_en.txt:88)
E_CHARGE = 1.602176487e-19


class SbModel(C.Structure):
    _fields_ = [
        ('n_bands', C.c_int32), ('bands_have_dofs', C.c_int32),
        ('n_procs', C.c_int32), ('n_fields', C.c_int32), ('n_tags', C.c_int32),
        ('quad_m_rho', C.c_int32), ('quad_m_super', C.c_int32),
        ('quad_m_g', C.c_int32),
        ('band_kind', C.c_int32 * MAX_BANDS),
        ('band_sign', C.c_int32 * MAX_BANDS),
        ('band_const_mobility', C.c_int32 * MAX_BANDS),
        ('proc_kind', C.c_int32 * MAX_PROCS),
        ('proc_dst', C.c_int32 * MAX_PROCS),
        ('proc_src', C.c_int32 * MAX_PROCS),
        ('proc_trap', C.c_int32 * MAX_PROCS),
        ('proc_radiative', C.c_int32 * MAX_PROCS),
        ('e_charge', C.c_double), ('dd_scale', C.c_double),
        ('ext_w_coef', C.c_double), ('ext_j_coef', C.c_double),
    ]


def make_model(bands, procs=(), n_fields=0, n_tags=1, bands_have_dofs=True,
               quad_degree_rho=20, quad_degree_super=20, quad_degree_g=8):
    """bands: list of dict(kind, sign, const_mobility); procs: list of
    dict(kind, dst, src, trap=-1, radiative=1)."""
    if len(bands) > MAX_BANDS or len(procs) > MAX_PROCS or n_fields > MAX_FIELDS:
        raise ValueError('model exceeds SB_MAX_* limits')
    m = SbModel()
    m.n_bands = len(bands)
    m.bands_have_dofs = 1 if bands_have_dofs else 0
    m.n_procs = len(procs)
    m.n_fields = n_fields
    m.n_tags = n_tags
    for name, deg in (('quad_m_rho', quad_degree_rho),
                      ('quad_m_super', quad_degree_super),
                      ('quad_m_g', quad_degree_g)):
        if deg <= 6:
            raise ValueError('quadrature degree must be > 6 (collapsed rule)')
        mm = (deg + 2) // 2
        if mm > MAX_QUAD_M:
            raise ValueError('quadrature degree too high')
        setattr(m, name, mm)
    for i, b in enumerate(bands):
        m.band_kind[i] = b['kind']
        m.band_sign[i] = b['sign']
        m.band_const_mobility[i] = 1 if b.get('const_mobility', False) else 0
    for i, p in enumerate(procs):
        m.proc_kind[i] = p['kind']
        m.proc_dst[i] = p['dst']
        m.proc_src[i] = p['src']
        m.proc_trap[i] = p.get('trap', -1)
        m.proc_radiative[i] = 1 if p.get('radiative', True) else 0
    m.e_charge = E_CHARGE
    m.dd_scale = 1.0 / E_CHARGE
    m.ext_w_coef = 1.0               # 1 A um^-3 / eV  in form units
    m.ext_j_coef = 1.0e4 / E_CHARGE  # 1 V s cm        in form units
    return m


def n_subproblems(model):
    return 1 + (model.n_bands if model.bands_have_dofs else 0)


def new_tag_params(n_tags):
    return np.zeros((n_tags, TAG_STRIDE), dtype=np.float64)
