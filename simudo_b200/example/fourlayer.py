"""Four-layer intermediate-band solar cell (BASELINE config 2).

Same device, parameters and solve sequence as the reference's
``simudo/example/fourlayer/fourlayer.py:306-691`` driven by
``fourlayer_example.py:63-110``: front-surface field / p / IB / n layers, three
bands (IB only inside the I layer), three forward optical fields fed by a
6000 K blackbody times the concentration X through inflow Dirichlet data,
band-to-band and IB absorption with radiative recombination, Shockley-Read
trapping, then

  local charge neutrality -> thermal equilibrium -> full problem,
  light ramp  (OpticalIntensityAdaptiveStepper 0 -> 1, self-consistent optics),
  bias sweep  (VoltageStepper over V_exts, self-consistent optics).

``run(P=..., backend_factory=...)`` returns the list of (V, J [mA/cm^2]).
"""
import logging

import numpy as np

from simudo_b200 import dolfin
from simudo_b200.io.output_writer import (MetaExtractorBandInfo, MetaExtractorIntegrals,
                                          OutputWriter)
from simudo_b200.mesh import CellRegions, ConstructionHelperLayeredStructure, FacetRegions
from simudo_b200.physics import (Material, NonOverlappingTopHatBeerLambert,
                                 NonOverlappingTopHatBeerLambertIB, NonRadiativeTrap,
                                 OpticalIntensityAdaptiveStepper, ProblemData, VoltageStepper)
from simudo_b200.util import Blackbody, make_unit_registry

# fourlayer_example.py:63-110
DEFAULT_PARAMS = dict(
    simple_mobility_model=True, mu_I=500, X=100, sigma_opt_ci=3e-14, sigma_opt_iv=3e-14,
    sigma_th_ci=5e-18, sigma_th_iv=5e-18, vth_CB=2e7, vth_VB=2e7, NI=2.5e17, IB_E=1.2,
    CB_E=2.0, VB_E=0.0, NC=2e19, NV=2e19, mu_n=500, mu_p=500, alpha_cv=1e4, epsilon_r=13,
    T=300, Ts=6000, NA=-1e17, ND=1e17, NFSF=-1e19, f_0=0.5,
    FSF_thickness=0.05, p_thickness=1.0, n_thickness=1.0, IB_thickness=3,
    mesh_start=0.005, mesh_factor=1.2, mesh_max=0.02,
    V_exts=list(np.linspace(0, 1.7, 26)))


class Semiconductor(Material):
    name = 'semiconductor'
    params = None

    def get_dict(self):
        d = super().get_dict()
        U, P = self.unit_registry, self.params
        d.update({
            'CB/energy_level': float(P['CB_E']) * U.eV,
            'IB/energy_level': float(P['IB_E']) * U.eV,
            'VB/energy_level': float(P['VB_E']) * U.eV,
            'IB/degeneracy': 1,
            'CB/effective_density_of_states': float(P['NC']) * U('1/cm^3'),
            'VB/effective_density_of_states': float(P['NV']) * U('1/cm^3'),
            'CB/mobility': float(P['mu_n']) * U('cm^2/V/s'),
            'VB/mobility': float(P['mu_p']) * U('cm^2/V/s'),
            'opt_cv/alpha': float(P['alpha_cv']) * U('cm^-1'),
            'poisson/permittivity': float(P['epsilon_r']) * U('vacuum_permittivity'),
        })
        return d


class IBSemiconductor(Semiconductor):
    name = 'IB'

    def get_dict(self):
        d = super().get_dict()
        U, P = self.unit_registry, self.params
        d.update({
            'opt_ci/sigma_opt': float(P['sigma_opt_ci']) * U('cm^2'),
            'opt_iv/sigma_opt': float(P['sigma_opt_iv']) * U('cm^2'),
            'IB/number_of_states': float(P['NI']) * U('cm^-3'),
            'IB/mobility0': float(P['mu_I']) * U('cm^2/V/s'),
        })
        return d


class FrontSurfaceField(Semiconductor):
    name = 'fsf'

    def get_dict(self):
        d = super().get_dict()
        d['opt_cv/alpha'] = self.unit_registry('0 cm^-1')
        return d


def topology_standard_contacts(cell_regions, facet_regions):
    R, F = cell_regions, facet_regions
    F.left_contact = lc = R.exterior_left.boundary(R.domain)
    F.right_contact = rc = R.domain.boundary(R.exterior_right)
    F.exterior = R.domain.boundary(R.exterior)
    F.contacts = lc | rc
    F.nonconductive = F.exterior - F.contacts.both()


def run(P=None, backend_factory=None, output_prefix=None, return_problem=False,
        product2d_Ys=None):
    P = dict(DEFAULT_PARAMS, **(P or {}))
    U = make_unit_registry(("mesh_unit = 1 micrometer",))
    s = dict(type='geometric', start=float(P['mesh_start']), factor=float(P['mesh_factor']))
    layers = [
        dict(name='fsf', material='fsf', thickness=float(P['FSF_thickness']), mesh=s),
        dict(name='p', material='semiconductor', thickness=float(P['p_thickness']), mesh=s),
        dict(name='I', material='IB', thickness=float(P['IB_thickness']), mesh=s),
        dict(name='n', material='semiconductor', thickness=float(P['n_thickness']), mesh=s),
    ]
    ls = ConstructionHelperLayeredStructure()
    ls.params = dict(edge_length=float(P['mesh_max']), layers=layers, extra_regions=[],
                     mesh_unit=U.um)
    if product2d_Ys is not None:
        ls.params['product2d_Ys'] = product2d_Ys
    ls.run()
    mesh_data = ls.mesh_data

    R, F = CellRegions(), FacetRegions()
    topology_standard_contacts(R, F)
    R.cell = R.p | R.n | R.I | R.fsf
    F.p_contact = F.left_contact
    F.n_contact = F.right_contact
    F.IB_bounds = R.IB.boundary(R.domain - R.IB) | R.IB.boundary(R.exterior)

    def create_problemdata(goal='full', V_ext=None, phi_cn=None):
        root = ProblemData(goal=goal, mesh_data=mesh_data, unit_registry=U,
                           backend_factory=backend_factory)
        pdd = root.pdd
        CB = pdd.easy_add_band('CB')
        IB = pdd.easy_add_band('IB', subdomain=R.I)
        VB = pdd.easy_add_band('VB')
        spatial = pdd.spatial
        q = U('elementary_charge/cm^3')
        spatial.add_rule('temperature', R.domain, float(P['T']) * U.K)
        spatial.add_rule('poisson/static_rho', R.p, float(P['NA']) * q)
        spatial.add_rule('poisson/static_rho', R.n, float(P['ND']) * q)
        spatial.add_rule('poisson/static_rho', R.fsf, float(P['NFSF']) * q)
        spatial.add_rule('poisson/static_rho', R.I, float(P['f_0']) * float(P['NI']) * q)
        spatial.add_rule('nr_top/sigma_th', R.I, float(P['sigma_th_ci']) * U('cm^2'))
        spatial.add_rule('nr_bottom/sigma_th', R.I, float(P['sigma_th_iv']) * U('cm^2'))
        spatial.add_rule('nr_top/v_th', R.I, float(P['vth_CB']) * U('cm/s'))
        spatial.add_rule('nr_bottom/v_th', R.I, float(P['vth_VB']) * U('cm/s'))
        FrontSurfaceField(problem_data=root, params=P).register()
        IBSemiconductor(problem_data=root, params=P).register()
        Semiconductor(problem_data=root, params=P).register()
        if P['simple_mobility_model']:
            IB.use_constant_mobility = True
        mu = pdd.mesh_util
        zeroE = F.exterior
        if goal == 'full':
            optical = root.optical
            ospatial = optical.spatial

            def ediff(lower, upper):
                return (float(P[upper + '_E']) - float(P[lower + '_E']) + 1e-10) * U.eV

            optical.easy_add_field('forward_ci', photon_energy=ediff('IB', 'CB'),
                                   direction=(1.0, 0.0))
            optical.easy_add_field('forward_iv', photon_energy=ediff('VB', 'IB'),
                                   direction=(1.0, 0.0))
            optical.easy_add_field('forward_cv', photon_energy=ediff('VB', 'CB'),
                                   direction=(1.0, 0.0))
            bb = Blackbody(temperature=float(P['Ts']) * U.K)
            keys = ('forward_cv', 'forward_ci', 'forward_iv')
            for name, (lower, upper) in bb.non_overlapping_energy_ranges(
                    {k: optical.fields[k].photon_energy for k in keys}).items():
                flux = (bb.photon_flux_integral_on_earth(lower, upper) * float(P['X'])).to(
                    '1/cm^2/s')
                ospatial.add_BC(name + '/Phi', F.left_contact, flux)
                logging.getLogger('main').info(
                    'Input photon flux for field {!r}: {}'.format(name, flux))
            pdd.easy_add_electro_optical_process(
                NonOverlappingTopHatBeerLambertIB, name='opt_ci', dst_band=CB, src_band=IB,
                trap_band=IB)
            pdd.easy_add_electro_optical_process(
                NonOverlappingTopHatBeerLambertIB, name='opt_iv', dst_band=IB, src_band=VB,
                trap_band=IB)
            pdd.easy_add_electro_optical_process(
                NonOverlappingTopHatBeerLambert, name='opt_cv', dst_band=CB, src_band=VB)
            NonRadiativeTrap.easy_add_two_traps_to_pdd(pdd, 'nr', CB, VB, IB)
            zj = U('A/cm^2') * mu.zerovec
            spatial.add_BC('CB/j', F.nonconductive, zj)
            spatial.add_BC('VB/j', F.nonconductive, zj)
            spatial.add_BC('IB/j', F.nonconductive | F.IB_bounds, zj)
            spatial.add_BC('VB/u', F.p_contact, VB.thermal_equilibrium_u)
            spatial.add_BC('CB/u', F.n_contact, CB.thermal_equilibrium_u)
            spatial.add_BC('VB/j', F.n_contact, zj)
            spatial.add_BC('CB/j', F.p_contact, zj)
            phi0 = pdd.poisson.thermal_equilibrium_phi
            spatial.add_BC('poisson/phi', F.p_contact, phi0 + V_ext)
            spatial.add_BC('poisson/phi', F.n_contact, phi0)
            zeroE = zeroE - (F.p_contact | F.n_contact).both()
        elif goal == 'thermal equilibrium':
            spatial.add_BC('poisson/phi', F.p_contact | F.n_contact, phi_cn)
            zeroE = zeroE - (F.p_contact | F.n_contact).both()
        spatial.add_BC('poisson/E', zeroE, U('V/m') * mu.zerovec)
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
            facets=F[{'p_contact', 'n_contact'}], cells=R[{'fsf', 'p', 'n', 'I'}], **kw))
    optics = float(P['X']) != 0
    if optics:
        stepper = OpticalIntensityAdaptiveStepper(
            solution=problem, parameter_target_values=[0, 1], step_size=1e-25,
            stepper_rel_tol=1e-6, selfconsistent_optics=True,
            output_writer=OutputWriter(filename_prefix=(output_prefix or 'fourlayer') + '/I',
                                       plot_iv=False, meta_extractors=meta_extractors))
        stepper.do_loop()
    writer = OutputWriter(filename_prefix=(output_prefix or 'fourlayer') + '/V',
                          plot_iv=output_prefix is not None, meta_extractors=meta_extractors)
    stepper = VoltageStepper(
        solution=problem, constants=[V_ext], step_size=1e-4, stepper_rel_tol=1e-6,
        parameter_target_values=list(P['V_exts']), parameter_start_value=P['V_exts'][0],
        parameter_unit=U.V, output_writer=writer, selfconsistent_optics=optics)
    stepper.do_loop()
    jv = [(row['sweep_parameter:parameter']['value'],
           row['avg:current_CB:p_contact']['value'] + row['avg:current_VB:p_contact']['value'])
          for row in writer.rows]
    if return_problem:
        return jv, problem, stepper
    return jv


if __name__ == '__main__':
    logging.basicConfig(level=logging.INFO)
    for V, J in run(output_prefix='out_fourlayer'):
        print('{:.5f} V   {: .6e} mA/cm^2'.format(V, J))
