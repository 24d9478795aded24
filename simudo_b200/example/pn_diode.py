"""Tutorial silicon pn-diode, dark J-V sweep (BASELINE config 1).

The same script as the reference's ``simudo/example/pn_diode/pn_diode.py``
with ``simudo`` replaced by ``simudo_b200`` (and ``dolfin`` by the shim): two
0.25 um Si layers doped +-1e18 cm^-3, geometric mesh 1e-3 um x 1.2^n, SRH
recombination, contacts and sweep ``linspace(0, 1, 21)`` V exactly as
:49-65, :110-164, :191-202.  ``run()`` returns the list of (V, J) with
J = avg:current_CB:p_contact + avg:current_VB:p_contact in mA/cm^2
(``pn_diode_plot.py:6-13``).
"""
import logging

import numpy as np

from simudo_b200 import dolfin
from simudo_b200.io.output_writer import (MetaExtractorBandInfo, MetaExtractorIntegrals,
                                          OutputWriter)
from simudo_b200.mesh import CellRegions, ConstructionHelperLayeredStructure, FacetRegions
from simudo_b200.physics import (ProblemData, SimpleSiliconMaterial, SRHRecombination,
                                 VoltageStepper)
from simudo_b200.util import make_unit_registry


def topology_standard_contacts(cell_regions, facet_regions):
    R, F = cell_regions, facet_regions
    F.left_contact = lc = R.exterior_left.boundary(R.domain)
    F.right_contact = rc = R.domain.boundary(R.exterior_right)
    F.exterior = R.domain.boundary(R.exterior)
    F.contacts = lc | rc
    F.nonconductive = F.exterior - F.contacts.both()


def build(length=0.5, product2d_Ys=None, mesh_start=0.001, mesh_factor=1.2, edge_length=None):
    """``mesh_start`` / ``mesh_factor`` / ``edge_length``: the tutorial's geometric x-mesh
    (pn_diode.py:49-65: 1e-3 um, 1.2, length / 20 -> 67 points); finer values give the
    2D configurations of BASELINE.json (configs[2]: ~224 x 224 points)."""
    U = make_unit_registry(("mesh_unit = 1 micrometer",))
    layers = [
        dict(name='pSi', material='Si', thickness=length / 2,
             mesh=dict(type="geometric", start=mesh_start, factor=mesh_factor)),
        dict(name='nSi', material='Si', thickness=length / 2,
             mesh=dict(type="geometric", start=mesh_start, factor=mesh_factor)),
    ]
    ls = ConstructionHelperLayeredStructure()
    ls.params = dict(edge_length=edge_length or length / 20, layers=layers, extra_regions=[],
                     mesh_unit=U.mesh_unit)
    if product2d_Ys is not None:
        ls.params['product2d_Ys'] = product2d_Ys
    ls.run()
    return U, ls.mesh_data


def run(output_prefix=None, voltages=None, backend_factory=None, product2d_Ys=None,
        return_problem=False, mesh=None):
    U, mesh_data = build(product2d_Ys=product2d_Ys, **(mesh or {}))
    R = CellRegions()
    F = FacetRegions()
    topology_standard_contacts(R, F)
    R.Si = R.pSi | R.nSi
    F.p_contact = F.left_contact
    F.pn_junction = R.pSi.boundary(R.nSi)
    F.n_contact = F.right_contact

    def create_problemdata(goal='full', V_ext=None, phi_cn=None):
        root = ProblemData(goal=goal, mesh_data=mesh_data, unit_registry=U,
                           backend_factory=backend_factory)
        pdd = root.pdd
        CB = pdd.easy_add_band(name='CB')
        VB = pdd.easy_add_band(name='VB')
        spatial = pdd.spatial
        spatial.add_rule('temperature', R.domain, U('300 K'))
        doping = 1e18
        spatial.add_rule('poisson/static_rho', R.pSi,
                         -float(doping) * U('elementary_charge/cm^3'))
        spatial.add_rule('poisson/static_rho', R.nSi,
                         +float(doping) * U('elementary_charge/cm^3'))
        SimpleSiliconMaterial(problem_data=root).register()
        mu = pdd.mesh_util
        if goal == 'full':
            pdd.easy_add_electro_optical_process(SRHRecombination, dst_band=CB, src_band=VB)
            spatial.add_BC('CB/j', F.nonconductive, U('A/cm^2') * mu.zerovec)
            spatial.add_BC('VB/j', F.nonconductive, U('A/cm^2') * mu.zerovec)
            spatial.add_BC('CB/j', F.p_contact, U('A/cm^2') * mu.zerovec)
            spatial.add_BC('VB/j', F.n_contact, U('A/cm^2') * mu.zerovec)
            spatial.add_BC('VB/u', F.p_contact, VB.thermal_equilibrium_u)
            spatial.add_BC('CB/u', F.n_contact, CB.thermal_equilibrium_u)
            phi0 = pdd.poisson.thermal_equilibrium_phi
            spatial.add_BC('poisson/phi', F.p_contact, phi0)
            spatial.add_BC('poisson/phi', F.n_contact, phi0 - V_ext)
        elif goal == 'thermal equilibrium':
            spatial.add_BC('poisson/phi', F.p_contact | F.n_contact, phi_cn)
        zeroE = F.exterior - (F.p_contact | F.n_contact).both()
        spatial.add_BC('poisson/E', zeroE, U('V/m') * mu.zerovec)
        spatial.add_rule('SRH/CB/tau', R.domain, U('1e-9 s'))
        spatial.add_rule('SRH/VB/tau', R.domain, U('1e-6 s'))
        spatial.add_rule('SRH/energy_level', R.domain, U('0.5525628437189991 eV'))
        return root

    problem = create_problemdata(goal='local charge neutrality')
    problem.pdd.easy_auto_pre_solve()

    problem0 = problem
    problem = create_problemdata(goal='thermal equilibrium', phi_cn=problem0.pdd.poisson.phi)
    problem.pdd.initialize_from(problem0.pdd)
    problem.pdd.easy_auto_pre_solve()

    V_ext = U.V * dolfin.Constant(0.0)
    problem0 = problem
    problem = create_problemdata(goal='full', V_ext=V_ext)
    problem.pdd.initialize_from(problem0.pdd)

    meta_extractors = (
        MetaExtractorBandInfo,
        lambda **kw: MetaExtractorIntegrals(
            facets=F[{'p_contact', 'n_contact'}], cells=R[{'pSi', 'nSi'}], **kw))
    writer = OutputWriter(filename_prefix=output_prefix or 'output V_stepper',
                          plot_iv=output_prefix is not None,
                          meta_extractors=meta_extractors)
    if voltages is None:
        voltages = np.linspace(0, 1, 21)
    stepper = VoltageStepper(solution=problem, constants=[V_ext],
                             parameter_target_values=voltages, parameter_unit=U.V,
                             output_writer=writer, selfconsistent_optics=False)
    stepper.do_loop()
    jv = [(row['sweep_parameter:parameter']['value'],
           row['avg:current_CB:p_contact']['value'] + row['avg:current_VB:p_contact']['value'])
          for row in writer.rows]
    if return_problem:
        return jv, problem, stepper
    return jv


if __name__ == '__main__':
    logging.basicConfig(level=logging.INFO)
    for V, J in run(output_prefix='output V_stepper'):
        print('{:.5f} V   {: .6e} mA/cm^2'.format(V, J))
