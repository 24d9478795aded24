"""Sweep output: terminal currents and the sweep-summary CSV.

The part of ``simudo/io/output_writer.py`` on the J-V path:
``OutputWriter.write_output`` (:68-111) -> meta extractors ->
``MetaCSVWriter`` (:113-135); ``MetaExtractorBandInfo`` (:144-152) and
``MetaExtractorIntegrals.add_surface_flux`` (:174-231) giving
``avg:current_<band>:<facet region>`` in mA/cm^2 and ``tot:current_...`` in
mA.  ``plot_1d`` writes the line cut of the unknowns (``io/csv.py``, :94-101); XDMF /
YAML+HDF5 are file formats outside the hot path.
"""
import csv
import os

import numpy as np

from ..util import SetattrInitMixin

__all__ = ['OutputWriter', 'MetaCSVWriter', 'MetaExtractorBandInfo',
           'MetaExtractorIntegrals']


class MetaCSVWriter(object):
    def __init__(self, filename, meta):
        self.columns = list(meta.keys())
        self.file = open(filename, 'wt')
        self.writer = csv.writer(self.file)
        self.writer.writerow(self.columns)
        units = [meta[c].get('units', '') for c in self.columns]
        units[0] = '# ' + units[0]
        self.writer.writerow(units)

    def add_row(self, meta):
        self.writer.writerow([meta[c]['value'] for c in self.columns])
        self.file.flush()


class OutputWriter(SetattrInitMixin):
    stepper = None
    plot_iv = True
    plot_1d = False
    plot_mesh = False
    parameter_name = 'parameter'
    filename_prefix = 'output'
    meta_extractors = ()
    line_cut_resolution = 5001
    _meta_csv = None

    def __init__(self, **kwargs):
        super().__init__(**kwargs)
        self.rows = []          # every written meta dict, for programmatic use

    def format_parameter(self, solution, parameter_value):
        return '{:.14g}'.format(parameter_value)

    def write_output(self, solution, parameter_value):
        d = os.path.dirname(self.filename_prefix)
        if d:
            os.makedirs(d, exist_ok=True)
        meta = {'sweep_parameter:{}'.format(self.parameter_name): dict(value=parameter_value)}
        for extractor in self.meta_extractors:
            extractor(solution=solution, parameter_value=parameter_value,
                      output_writer=self, meta=meta).call()
        self.rows.append(meta)
        if self.plot_1d:
            from .csv import LineCutCsvPlot, solution_line_cut
            prefix = self.filename_prefix + '_{}={}'.format(
                self.parameter_name, self.format_parameter(solution, parameter_value))
            with LineCutCsvPlot(prefix + '.csv', None,
                                resolution=self.line_cut_resolution) as plotter:
                solution_line_cut(plotter, solution, 0)
        if self.plot_iv:
            if self._meta_csv is None:
                self._meta_csv = MetaCSVWriter(
                    self.filename_prefix + '_{}.csv'.format(self.parameter_name), meta)
            self._meta_csv.add_row(meta)


class MetaExtractorBandInfo(SetattrInitMixin):
    prefix = 'band_info:'

    def call(self):
        for k, band in self.solution.pdd.bands.items():
            self.meta['{}{}:sign'.format(self.prefix, band.name)] = dict(value=band.sign)


class MetaExtractorIntegrals(SetattrInitMixin):
    """``facets``: {name: FacetRegion}; ``cells``: {name: CellRegion} (volume
    integrals of g are not on the J-V path and are skipped)."""
    prefix = ''
    facets = {}
    cells = {}

    def call(self):
        pdd = self.solution.pdd
        U = pdd.unit_registry
        md = pdd.mesh_data
        mesh = md.mesh
        ff = np.asarray(md.facet_function)
        for bi, (k, band) in enumerate(pdd.bands.items()):
            if not getattr(band, 'has_dofs', False):
                continue
            for reg_name, reg in self.facets.items():
                flux, length = 0.0, 0.0
                for fv, sign in md.evaluate_topology(reg):
                    facets = np.nonzero(ff == fv)[0]
                    ext = facets[mesh.facet_cells[facets, 1] < 0]
                    if len(ext):
                        f, l = pdd.system.terminal_current(bi, ext)
                        # ds(fv_pos) - ds(fv_neg): the facet sign orients the region
                        flux += sign * f
                        length += l
                tot = (flux * U('A/mesh_unit')).m_as('mA/mesh_unit')
                avg = ((flux / length) * U('A/mesh_unit^2')).m_as('mA/cm^2') if length else np.nan
                self.meta['{}avg:current_{}:{}'.format(self.prefix, k, reg_name)] = dict(
                    value=avg, units='mA/cm^2')
                self.meta['{}tot:current_{}:{}'.format(self.prefix, k, reg_name)] = dict(
                    value=tot, units='mA')
