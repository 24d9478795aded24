"""End-to-end: the tutorial pn-diode J-V sweep (BASELINE config 1).

CPU part: host logic (lowering, three-stage initialisation, Newton control
flow, adaptive stepper, terminal-current extraction) driven through the oracle
backend, pinned against the reference's own published J-V curve
(tests/golden/tutorial_pn_diode_JV_svg.json, decoded from the reference's
doc/source/_static/tutorial-pn-diode-JV.svg).
GPU part: the same sweep through the CUDA backend against the oracle fixture at
the north star's 1e-6 relative J-V tolerance.
"""
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def _golden(name):
    with open(os.path.join(HERE, 'golden', name)) as f:
        return json.load(f)['jv']


@pytest.fixture(scope='module')
def oracle_sweep():
    from oracle_backend import OracleBackend
    from simudo_b200.example import pn_diode
    jv, problem, stepper = pn_diode.run(backend_factory=OracleBackend, return_problem=True)
    return jv, problem, stepper


def test_mesh_and_dof_counts_match_survey(oracle_sweep):
    jv, problem, stepper = oracle_sweep
    pa = problem.pdd.system.pa
    assert pa.n_cells == 132 and pa.P == 3          # 67 x-points
    assert pa.N == 3 * (6 * 132 + 3 * pa.mesh.num_facets)


def test_thermal_equilibrium_is_current_free(oracle_sweep):
    jv, _, _ = oracle_sweep
    assert jv[0][0] == 0.0
    assert abs(jv[0][1]) < 1e-9 * abs(jv[1][1])


def test_jv_matches_reference_published_curve(oracle_sweep):
    """The only end-to-end numbers of this path in the reference tree.  The
    figure's Simudo/FEniCS versions are not recorded, so
    this is a band, not a 1e-6 golden: measured agreement 7e-6 .. 1.1e-3 up to
    0.9 V; the high-injection tail (> 0.9 V) deviates up to 5 %.  That tail is the one part of
    the curve that is not mesh-converged on the tutorial mesh: J(1.0 V) moves from -12 % (4x
    coarser) over +3.9 % (2x coarser), +5.5 % (tutorial mesh) to +6.4 % (4x finer) relative to
    the figure while J(<= 0.9 V) stays within 0.2 % on all of them -- the figure's last point
    is consistent with a slightly coarser effective mesh, not with different physics."""
    jv, _, _ = oracle_sweep
    ref = _golden('tutorial_pn_diode_JV_svg.json')
    ours = {round(v, 2): j for v, j in jv}
    for V, J in ref:
        rel = abs(ours[round(V, 2)] - J) / J
        if V <= 0.9 + 1e-9:
            assert rel < 2e-3, (V, rel)
        else:
            assert rel < 8e-2, (V, rel)
    # ideality checks the survey lists: ~60 mV/decade region around 0.6-0.7 V
    slope = np.log10(ours[0.7] / ours[0.6]) / 2
    assert 0.70 < slope < 0.80


def test_jv_regression_against_oracle_fixture(oracle_sweep):
    jv, _, _ = oracle_sweep
    ref = _golden('pn_diode_JV_oracle.json')
    for (V, J), (V0, J0) in zip(jv[1:], ref[1:]):
        assert V == pytest.approx(V0)
        assert J == pytest.approx(J0, rel=1e-8)


def test_newton_iteration_budget(oracle_sweep):
    _, problem, stepper = oracle_sweep
    # every bias point converged on the first try (no step halving) ...
    assert all(e.success for e in stepper.prior_solutions)
    # ... and the rebasing keeps |delta_w| small: base_w carries the qfl
    sysm = problem.pdd.system
    lay = sysm.pa.layout()
    dw = sysm.u_host()[sysm.pa.cell_dofs[:, lay['dg'][1]]]
    assert np.abs(dw).max() < 0.2
    assert np.abs(sysm.base_w_host()).max() > 0.5


@pytest.mark.gpu
def test_jv_sweep_on_gpu_matches_oracle_fixture():
    from simudo_b200 import _lib
    _lib._reset_for_tests()
    from simudo_b200.example import pn_diode
    jv = pn_diode.run()                       # default backend: CUDA
    ref = _golden('pn_diode_JV_oracle.json')
    assert len(jv) == len(ref)
    for (V, J), (V0, J0) in zip(jv[1:], ref[1:]):
        assert V == pytest.approx(V0)
        assert J == pytest.approx(J0, rel=1e-6), (V, J, J0)


@pytest.mark.gpu
def test_2d_pn_diode_newton_solve_on_gpu_multifrontal():
    """BASELINE configs[2] at test size: the tutorial diode on a genuinely two-dimensional
    product mesh (67 x 9 points) solved through the CUDA backend -- assembly by the native
    kernels, Newton systems by the multifrontal LU (the band of this mesh is too wide for
    the banded solver).  Data and contacts do not depend on y, so the discrete solution is
    the 1D one: J(V) must reproduce the oracle fixture of the 1D sweep to 1e-6."""
    from simudo_b200.example import pn_diode
    from simudo_b200.fem.multifrontal import MultifrontalSolver
    from simudo_b200.fem import linear_solver
    old = linear_solver.BAND_LIMIT_BYTES
    linear_solver.BAND_LIMIT_BYTES = 4 << 20        # force the 2D path at this small size
    try:
        jv, problem, _ = pn_diode.run(voltages=[0.0, 0.3, 0.6], return_problem=True,
                                      product2d_Ys=np.linspace(0.0, 1.0, 9))
    finally:
        linear_solver.BAND_LIMIT_BYTES = old
    be = problem.pdd.system.backend
    assert isinstance(be._solver, MultifrontalSolver)
    ref = {round(v, 2): j for v, j in _golden('pn_diode_JV_oracle.json')}
    for v, j in jv:
        j0 = ref[round(v, 2)]
        if v == 0.0:
            assert abs(j) < 1e-9 * abs(ref[0.3])
        else:
            assert abs(j - j0) <= 1e-6 * abs(j0), (v, j, j0)


def test_nonzero_essential_flux_value_is_lowered_and_imposed():
    """General essential BC values (fem/spatial.py:275-306; DirichletBC on the BDM sub-space):
    E = 0.5 V/um * unitvec_y on the non-conductive sides of the diode's thermal-equilibrium
    stage.  The lowered dof values follow FIAT's scaled-normal functionals, the solved field
    carries them, and its boundary flux is the prescribed E . n times the length."""
    from oracle_backend import OracleBackend
    from oracle import oracle as orc_mod
    from simudo_b200.example import pn_diode
    from simudo_b200.example.pn_diode import (CellRegions, FacetRegions, ProblemData,
                                              SimpleSiliconMaterial, topology_standard_contacts)
    U, mesh_data = pn_diode.build(product2d_Ys=np.linspace(0.0, 0.2, 4))
    R, F = CellRegions(), FacetRegions()
    topology_standard_contacts(R, F)
    R.Si = R.pSi | R.nSi
    F.p_contact, F.n_contact = F.left_contact, F.right_contact

    def create(goal, phi_cn=None):
        root = ProblemData(goal=goal, mesh_data=mesh_data, unit_registry=U,
                           backend_factory=OracleBackend)
        pdd = root.pdd
        pdd.easy_add_band(name='CB')
        pdd.easy_add_band(name='VB')
        sp = pdd.spatial
        sp.add_rule('temperature', R.domain, U('300 K'))
        sp.add_rule('poisson/static_rho', R.pSi, -1e18 * U('elementary_charge/cm^3'))
        sp.add_rule('poisson/static_rho', R.nSi, +1e18 * U('elementary_charge/cm^3'))
        SimpleSiliconMaterial(problem_data=root).register()
        mu = pdd.mesh_util
        if goal == 'thermal equilibrium':
            sp.add_BC('poisson/phi', F.p_contact | F.n_contact, phi_cn)
        zeroE = F.exterior - (F.p_contact | F.n_contact).both()
        sp.add_BC('poisson/E', zeroE, U('V/um') * (0.5 * mu.unitvec_y))
        return root
    p0 = create('local charge neutrality')
    p0.pdd.easy_auto_pre_solve()
    p1 = create('thermal equilibrium', phi_cn=p0.pdd.poisson.phi)
    p1.pdd.initialize_from(p0.pdd)
    sysm = p1.pdd.system
    pa, mesh = sysm.pa, sysm.mesh
    # lowered values: horizontal facets run from the lower to the higher vertex id, i.e. in
    # +x: n_scaled = (t_y, -t_x) = (0, -dx), so every dof is -0.5 dx
    xy = mesh.coordinates[mesh.facet_vertices]
    dx = {}
    for f in mesh.boundary_facets():
        if abs(xy[f, 0, 1] - xy[f, 1, 1]) < 1e-14:
            dx[f] = abs(xy[f, 1, 0] - xy[f, 0, 0])
    assert len(dx) == 2 * (len(np.unique(mesh.coordinates[:, 0])) - 1)
    want = {}
    for f, d in dx.items():
        for dof in pa.facet_dofs(0, np.array([f]))[0]:
            want[int(dof)] = -0.5 * d
    got = dict(zip(pa.bc_dofs.tolist(), sysm.bc_values.tolist()))
    assert set(want) == set(got)
    for k in want:
        assert abs(got[k] - want[k]) <= 1e-15
    p1.pdd.easy_auto_pre_solve()
    u = sysm.u_host()
    for k in want:                                  # Newton's x0 lifting lands on the value
        assert abs(u[k] - want[k]) <= 1e-12
    # boundary flux of the solved E through the top side: E . n = +0.5 V/um, bottom: -0.5
    ymid = xy[list(dx), :, 1].mean(axis=1)
    fl = np.array(list(dx))
    top, bottom = fl[ymid > 0.1], fl[ymid < 0.1]
    Lx = mesh.coordinates[:, 0].max()
    ft, _ = orc_mod.surface_flux(pa, u, 0, top)
    fb, _ = orc_mod.surface_flux(pa, u, 0, bottom)
    assert abs(ft - 0.5 * Lx) < 1e-12 and abs(fb + 0.5 * Lx) < 1e-12


def test_line_cut_csv_of_the_converged_solution(oracle_sweep, tmp_path):
    """``OutputWriter.plot_1d`` (simudo/io/output_writer.py:94-101, io/csv.py:21-77): the line
    cut of the unknowns at the last bias point; phi drops by the applied bias + built-in
    potential across the diode, j_tot is constant along x (steady state, 1D) and equals the
    terminal current of the J-V table."""
    from simudo_b200.io.csv import LineCutCsvPlot, solution_line_cut
    jv, problem, stepper = oracle_sweep
    prefix = str(tmp_path / 'cut.csv')
    with LineCutCsvPlot(prefix, None, resolution=401) as plotter:
        solution_line_cut(plotter, problem, 0)
    with open(prefix + '.0') as f:
        names = f.readline().strip().split(',')
        data = np.array([[float(x) for x in line.split(',')] for line in f])
    col = {n: data[:, i] for i, n in enumerate(names)}
    assert data.shape[0] == 401 and names[:2] == ['coord_x', 'coord_y']
    for k in ('phi', 'E_x', 'E_y', 'j_CB_x', 'j_VB_x', 'w_CB_delta', 'w_VB_base', 'qfl_CB', 'j_tot_x'):
        assert k in col and np.isfinite(col[k]).all(), k
    assert np.allclose(col['coord_y'], col['coord_y'][0])
    # quasi-1D: no transverse field or current
    assert np.abs(col["E_y"]).max() < 1e-4 * np.abs(col["E_x"]).max()
    # steady state: total current constant along the device, equal to the terminal current
    J = jv[-1][1]
    assert np.abs(np.abs(col['j_tot_x']) - abs(J)).max() < 2e-3 * abs(J)


def test_rebasing_with_flood_fill_thresholds_preserves_the_quasi_fermi_level(oracle_sweep):
    """``mixedqfl_debug_fill_thresholds`` != (0, 0) (``poisson_drift_diffusion1.py1:339-383``,
    ``fem/offset_partitioning.py:108-169``): base_w becomes piecewise constant over regions,
    delta_w absorbs the difference, w = base_w + delta_w is unchanged at every DG1 node."""
    _, problem, _ = oracle_sweep
    pdd = problem.pdd
    sysm = pdd.system
    band = pdd.bands['CB']
    bi = band.index
    lay = sysm.pa.layout()
    dg = sysm.pa.cell_dofs[:, lay['dg'][1 + bi]]
    saved = sysm.save()
    try:
        w_before = sysm.base_w_host()[bi][:, None] + sysm.u_host()[dg]
        n_before = len(np.unique(np.round(sysm.base_w_host()[bi], 12)))
        band.mixedqfl_debug_fill_thresholds = (5e-2, 0.0)
        band.mixedqfl_do_rebase_w()
        base = sysm.base_w_host()[bi]
        w_after = base[:, None] + sysm.u_host()[dg]
        assert np.abs(w_after - w_before).max() < 1e-12
        assert len(np.unique(base)) < n_before / 4          # regions, not cells
        # ... and the residual of the converged state is unchanged by the re-split
        sysm.assemble()
        assert np.abs(sysm.backend.get_b()).max() < 1e-6
    finally:
        del band.mixedqfl_debug_fill_thresholds
        sysm.restore(saved)


def test_thermal_equilibrium_potential_against_the_depletion_approximation(oracle_sweep):
    """The reference's depletion-approximation helper (``simudo/trash/deplapprox.py:6-78``):
    ``V_bi = kT/q ln(n_n0 p_p0 / n_i^2)``, ``w_n,p = sqrt(2 eps V_bi / (q (N_D,A / N_A,D + 1)
    N_D,A))`` and the piecewise-parabolic ``depl_V(x)``.  The thermal-equilibrium stage of the
    tutorial diode (N_A = N_D = 1e18 cm^-3, silicon) must (1) drop exactly V_bi between the
    neutral contacts -- an identity for Boltzmann statistics -- and (2) follow depl_V to the
    accuracy of that approximation (a few kT: it neglects the carrier tails)."""
    _, problem, _ = oracle_sweep
    sysm = problem.pdd.system
    U = problem.pdd.unit_registry
    pa = sysm.pa
    from simudo_b200.fem import model as M
    tp = pa.tag_params
    kT = tp[0, M.TP_KT]                                          # eV
    eps = tp[0, M.TP_EPS]                                        # elementary charges / (V um)
    NC = tp[0, M.TP_BAND0 + 0 * 5 + M.TPB_N]                     # 1 / um^3
    NV = tp[0, M.TP_BAND0 + 1 * 5 + M.TPB_N]
    Eg = tp[0, M.TP_BAND0 + 0 * 5 + M.TPB_E] - tp[0, M.TP_BAND0 + 1 * 5 + M.TPB_E]
    ni2 = NC * NV * np.exp(-Eg / kT)
    assert abs(eps / (11.7 * 55.26349406) - 1.0) < 1e-6          # 11.7 vacuum permittivities
    q = 1.0                                                      # charge is counted in e
    N_A = N_D = 1e18 * 1e-12                                     # 1 / um^3
    n_n0 = 0.5 * (N_D + np.sqrt(N_D ** 2 + 4 * ni2))
    p_p0 = 0.5 * (N_A + np.sqrt(N_A ** 2 + 4 * ni2))
    V_bi = kT * np.log(n_n0 * p_p0 / ni2)                        # V (kT in eV)
    w_n = np.sqrt(2 * V_bi * eps / (q * (N_D / N_A + 1) * N_D))
    w_p = np.sqrt(2 * V_bi * eps / (q * (N_A / N_D + 1) * N_A))
    # thermal-equilibrium potential at the cell vertices (P2 nodal values 0..2)
    x = pa.cell_vertices[:, :, 0].ravel()
    phi = sysm.thmeq_phi[:, :3].ravel()
    order = np.argsort(x, kind='stable')
    x, phi = x[order], phi[order]
    x_pn = 0.5 * (x[0] + x[-1])
    drop = abs(phi[-1] - phi[0])
    assert abs(drop - V_bi) < 1e-6 * V_bi, (drop, V_bi)
    assert 0.9 < V_bi < 1.0 and 0.02 < w_n < 0.03               # 0.956 V, 25 nm for 1e18 / 1e18

    def depl_V(xx):
        z_p, z_n = x_pn - w_p, x_pn + w_n
        if xx < z_p:
            return 0.0
        V = q / eps * N_A * 0.5 * (min(x_pn, xx) - z_p) ** 2
        if xx >= x_pn:
            V += q / eps * N_D * 0.5 * ((x_pn - z_n) ** 2 - (z_n - min(xx, z_n)) ** 2)
        return V
    ref = np.array([depl_V(v) for v in x])
    s = np.sign(phi[-1] - phi[0])
    dev = np.abs(s * (phi - phi[0]) - ref).max()
    assert dev < 3.0 * kT, (dev, kT)                             # measured ~1.3 kT
    # the depletion region is where the potential varies: 10 % .. 90 % of V_bi within w_p + w_n
    inside = (s * (phi - phi[0]) > 0.1 * V_bi) & (s * (phi - phi[0]) < 0.9 * V_bi)
    assert x[inside].max() - x[inside].min() < w_n + w_p
