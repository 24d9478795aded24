"""Known answers the reference tree itself holds for this path, pinned on
the *lowering* (unit folding, Strandberg integral, SRH levels) and on the mesh layout --
the parts oracle and kernels share as input and therefore cannot check on each other.

Sources (paths under /root/reference/simudo):
* misc/strandberg_comparison_parameters.py:45-66 (closed form by sympy), :88-90 and
  :101-102 (comment values "1308237.46130036 / second", "3750.0 / centimeter")
* misc/comparison_parameters.py:18-29 (n_i_eff = 9389112518.281794 cm^-3 defines the SRH
  trap level 0.5525628437189991 eV used by example/pn_diode/pn_diode.py:163-164; comment
  values p1 = 1.87786e10, n1 = 4.69466e9 cm^-3)
* mesh/product2d.py:116-122 (cells_at_ix)
The numbers were produced with stock pint constants (CODATA 2010 e), the reference's own
registry uses e = 1.602176487e-19 (util/pint/default_en.txt:88): agreement is at the 1e-7
level, asserted at 2e-6.
"""
import numpy as np
import pytest

from simudo_b200.fem import model as M
from simudo_b200.util import make_unit_registry


class _Stop(Exception):
    pass


def _capture_full_problem(run, **kw):
    """Run an example's set-up (local charge neutrality and thermal equilibrium through the
    oracle backend) up to the lowering of the full problem and return its ProblemArrays."""
    from oracle_backend import OracleBackend
    cap = {}

    class Capture(OracleBackend):
        def __init__(self, pa, n_threads=1):
            if pa.model.n_procs > 0:
                cap['pa'] = pa
                raise _Stop()
            super().__init__(pa, n_threads)

    try:
        run(backend_factory=Capture, **kw)
    except _Stop:
        pass
    return cap['pa']


@pytest.fixture(scope='module')
def U():
    return make_unit_registry(('mesh_unit = 1 micrometer',))


def test_strandberg_integral_closed_form_vs_quadrature(U):
    """strandberg_I = K int_{E_L}^{E_U} E^2 exp(-E / kT) dE, K = 8 pi n^2 / (h^3 c^2)
    (electro_optical_process.py:324-353; strandberg_comparison_parameters.py:41-56)."""
    from scipy import integrate
    from simudo_b200.physics.lowering import strandberg_I
    kT = (U('1 boltzmann_constant') * U('300 K')).m_as('eV')
    h = 6.62606957e-34 / 1.602176487e-19            # eV s   (constants_en.txt:15, default_en.txt:88)
    c = 299792458.0 * 1e6                           # um / s
    K = 8 * np.pi / (h ** 3 * c ** 2)
    for lo, hi in ((0.8, 1.2), (1.2, 2.0), (2.0, 1e3)):
        q, _ = integrate.quad(lambda E: E * E * np.exp(-E / kT), lo, min(hi, lo + 60 * kT * 40),
                              epsabs=0, epsrel=1e-13)
        assert abs(strandberg_I(lo, hi, kT, U) / (K * q) - 1) < 1e-10


def test_strandberg_lowering_matches_reference_comment_values(U):
    """Four-layer cell at the reference's parameters (fourlayer_example.py:63-110 =
    strandberg_comparison_parameters.py:11-75): the lowered pseudo-capture coefficient
    sigma_opt I_ci / u1 of the IB -> CB process times the empty states at half filling is the
    reference's 1 / tau_ci = 1308237.46 / s; alpha at half filling is 3750 / cm."""
    from simudo_b200.example import fourlayer
    small = dict(p_thickness=0.2, n_thickness=0.2, IB_thickness=0.4, mesh_start=0.01,
                 mesh_factor=1.5, mesh_max=0.05)
    pa = _capture_full_problem(lambda **kw: fourlayer.run(P=small, **kw))
    model, tp = pa.model, pa.tag_params
    kinds = [model.proc_kind[p] for p in range(model.n_procs)]
    assert kinds == [M.PROC_RAD_IB, M.PROC_RAD_IB, M.PROC_RAD_BB, M.PROC_SR_TRAP, M.PROC_SR_TRAP]
    NI = tp[:, M.TP_BAND0 + 5 * 1 + M.TPB_N]          # IB states per um^3, per tag
    tag_I = int(np.argmax(NI))
    assert abs(NI[tag_I] * 1e12 / 2.5e17 - 1) < 1e-12 and (np.delete(NI, tag_I) == 0).all()
    coef_ci = tp[tag_I, M.TP_PROC0 + 8 * 0 + M.TPP_P0]            # um^3 / s
    inv_tau_ci = coef_ci * NI[tag_I] / 2
    assert abs(inv_tau_ci / 1308237.46130036 - 1) < 2e-6
    # absorption cross section of the same process in field 'forward_ci' only, 3750 / cm at f = 1/2
    sig = tp[tag_I, M.TP_PROC0 + 8 * 0 + M.TPP_FIELD0:M.TP_PROC0 + 8 * 0 + M.TPP_FIELD0 + 3]
    assert (sig > 0).sum() == 1
    alpha_half = sig.max() * NI[tag_I] / 2                         # 1 / um
    assert abs(alpha_half * 1e4 / 3750.0 - 1) < 1e-12
    assert abs((U('2.5e17 1/cm^3') / 2 * U('3e-14 cm^2')).m_as('1/cm') / 3750.0 - 1) < 1e-12
    # processes that have no rule outside the IB layer are exactly off there
    for t in range(tp.shape[0]):
        if t != tag_I:
            for p in (0, 1, 3, 4):
                assert (tp[t, M.TP_PROC0 + 8 * p:M.TP_PROC0 + 8 * p + 8] == 0).all()


def test_srh_levels_of_the_tutorial_diode():
    """pn_diode.py:161-164 with material.py:76-89: SRH n1, p1 as lowered for the kernels.
    The trap level 0.5525628437189991 eV is defined by p1 = n_i_eff = 9389112518.281794
    cm^-3 (comparison_parameters.py:18-29), hence also n1 = n_i^2 / p1 = n_i_eff."""
    from simudo_b200.example import pn_diode
    pa = _capture_full_problem(pn_diode.run)
    model, tp = pa.model, pa.tag_params
    assert model.n_procs == 1 and model.proc_kind[0] == M.PROC_SRH
    dst, src = model.proc_dst[0], model.proc_src[0]
    assert (model.band_sign[dst], model.band_sign[src]) == (-1, +1)       # CB gains, VB
    o = M.TP_PROC0
    for t in range(tp.shape[0]):
        tau_n, tau_p, n1, p1 = tp[t, o:o + 4]
        assert (tau_n, tau_p) == (1e-9, 1e-6)
        assert abs(p1 * 1e12 / 9389112518.281794 - 1) < 2e-6
        assert abs(n1 * 1e12 / 9389112518.281794 - 1) < 2e-6
    # the comment values of the same file (a trap level ln 2 kT off mid-gap): mass action
    ni2 = 9389112518.281794 ** 2
    assert abs(1.87786e10 * 4.69466e9 / ni2 - 1) < 1e-4
    # n_i_eff itself from the lowered band parameters
    kT = tp[0, M.TP_KT]
    Nc, Ec = tp[0, M.TP_BAND0 + 5 * dst + M.TPB_N], tp[0, M.TP_BAND0 + 5 * dst + M.TPB_E]
    Nv, Ev = tp[0, M.TP_BAND0 + 5 * src + M.TPB_N], tp[0, M.TP_BAND0 + 5 * src + M.TPB_E]
    ni = np.sqrt(Nc * Nv) * np.exp(-(Ec - Ev) / (2 * kT)) * 1e12
    assert abs(ni / 9389112518.281794 - 1) < 2e-6


def test_product2d_cell_layout():
    """mesh/product2d.py:116-122 and the geometry behind it: the cells of x-interval ix are
    contiguous and lie inside that interval."""
    from simudo_b200.mesh import Product2DMesh
    o = Product2DMesh([0, 1, 2, 3], [0, 1, 2])
    assert list(o.cells_at_ix(1)) == [4, 5, 6, 7]
    m = o.mesh
    for ix in range(3):
        x = m.coordinates[m.cells[list(o.cells_at_ix(ix))]][:, :, 0]
        assert x.min() == ix and x.max() == ix + 1
    assert list(o.vertices_at_ix(2)) == [6, 7, 8] and list(o.vertices_at_iy(1)) == [1, 4, 7, 10]
