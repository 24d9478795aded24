"""Line cuts of the solution into CSV files.

Mirror of ``simudo/io/csv.py:21-77`` (``LineCutCsvPlot``) and of the part of
``solution_plot`` (``simudo/io/output_writer.py:242-330``) that plots the unknowns: the mid
line ``y = mean(y range)`` of the mesh sampled at ``resolution`` (default 5001) points,
one CSV column per quantity, one file per ``new(timestamp)``.  The reference renders each
function through ``Function_eval_many`` (``fem/extra_bindings.py:27-75``); here the point
location and evaluation run on the device (``sb_eval_points``), the only thing that leaves
the GPU is the sampled line.  Vector quantities give two columns ``name_x``, ``name_y``.
Derived fields the reference obtains by projection (``u``, ``g``, mobility, optical
fields) are not plotted -- they are not unknowns of the path.
"""
import os

import numpy as np

__all__ = ['LineCutCsvPlot', 'solution_line_cut']


class LineCutCsvPlot(object):
    timestamp = 0.0

    def __init__(self, filename, system=None, resolution=5001):
        self.base_filename = filename
        self.system = system
        self.resolution = resolution
        self._columns = {}
        self._coords = None

    def new(self, timestamp):
        self.close()
        self.timestamp = timestamp

    def coordinates(self):
        """``resolution`` points from (xmin, ymid) to (xmax, ymid) (csv.py:30-41)."""
        if self._coords is None:
            v = np.asarray(self.system.pa.cell_vertices, dtype=np.float64).reshape(-1, 2)
            (x0, y0), (x1, y1) = v.min(axis=0), v.max(axis=0)
            t = np.linspace(0, 1, self.resolution)
            ym = 0.5 * (y0 + y1)
            self._coords = np.outer(1 - t, (x0, ym)) + np.outer(t, (x1, ym))
        if not self._columns:
            self._columns['coord_x'] = self._coords[:, 0]
            self._columns['coord_y'] = self._coords[:, 1]
        return self._coords

    def add(self, name, sub_problem, vector=False, scale=1.0):
        """Sample sub-function ``sub_problem`` of the system's current solution."""
        pts = self.coordinates()
        vals = self.system.backend.eval_points(sub_problem, pts, vector=vector) * scale
        if vector:
            self._columns[name + '_x'] = vals[:, 0]
            self._columns[name + '_y'] = vals[:, 1]
        else:
            self._columns[name] = vals

    def add_values(self, name, values):
        self.coordinates()
        self._columns[name] = np.asarray(values, dtype=np.float64)

    def close(self):
        if len(self._columns) <= 2:
            self._columns = {}
            return
        d = os.path.dirname(self.base_filename)
        if d:
            os.makedirs(d, exist_ok=True)
        names = list(self._columns)
        data = np.column_stack([self._columns[k] for k in names])
        with open(self.base_filename + '.{}'.format(self.timestamp), 'wt') as f:
            f.write(','.join(names) + '\n')
            for row in data:
                f.write(','.join(repr(float(x)) for x in row) + '\n')
        self._columns = {}

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def solution_line_cut(plotter, solution, timestep):
    """The unknowns of ``solution_plot`` (output_writer.py:242-330) along the cut: ``phi`` [V],
    ``E`` [V/mesh_unit], and per band with unknowns ``j_<k>`` [mA/cm^2], ``w_<k>_delta`` [eV],
    ``w_<k>_base`` [eV], ``qfl_<k>`` = base + delta [eV]; ``j_tot``."""
    pdd = solution.pdd
    U = pdd.unit_registry
    system = pdd.system
    plotter.system = system
    plotter.new(timestep)
    plotter.add('phi', 0)
    plotter.add('E', 0, vector=True)
    jscale = (1.0 * U('A/mesh_unit^2')).m_as('mA/cm^2')
    jtot = None
    if getattr(system, 'full', False):
        base_w = system.base_w_host()
        cells = None
        for bi, (k, band) in enumerate(pdd.bands.items()):
            if not getattr(band, 'has_dofs', False):
                continue
            p = 1 + bi
            plotter.add('j_' + k, p, vector=True, scale=jscale)
            jk = np.column_stack([plotter._columns['j_%s_x' % k], plotter._columns['j_%s_y' % k]])
            jtot = jk if jtot is None else jtot + jk
            plotter.add('w_%s_delta' % k, p)
            if cells is None:
                cells = system.backend.locate_points(plotter.coordinates())
            wb = np.where(cells >= 0, np.asarray(base_w[bi])[np.maximum(cells, 0)], np.nan)
            plotter.add_values('w_%s_base' % k, wb)
            plotter.add_values('qfl_' + k, wb + plotter._columns['w_%s_delta' % k])
        if jtot is not None:
            plotter.add_values('j_tot_x', jtot[:, 0])
            plotter.add_values('j_tot_y', jtot[:, 1])
