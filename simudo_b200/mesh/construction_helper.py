"""Mesh construction helpers (layered 1D structures on a product mesh).

Host-side restatement of the parts of ``simudo/mesh/construction_helper.py``
used by every shipped example: ``BaseConstructionHelper.run`` (``:108-115``),
exterior boundary pseudo cell values right/top/left/bottom (``:143-169``),
``FacetsManager`` construction (``:171-176``),
``ConstructionHelperIntervalProduct2DMesh`` (``:405-440``) and
``ConstructionHelperLayeredStructure`` (``:442-578``).  gmsh / mshr / Plaza
refinement back-ends are out of scope.
"""
import warnings

import numpy as np

from .facet import FacetsManager, facet2d_angle_array
from .interval1dtag import (CInterval, GeometricallyExpandingMeshInterval,
                            Interval, Interval1DTag, clip_intervals)
from .mesh_data import MeshData
from .topology import CellRegion

__all__ = ['BaseConstructionHelper',
           'ConstructionHelperIntervalProduct2DMesh',
           'ConstructionHelperLayeredStructure']


class _AttrDict(object):
    def __init__(self, d):
        object.__setattr__(self, '_d', d)

    def __getattr__(self, k):
        try:
            return self._d[k]
        except KeyError:
            raise AttributeError(k)


class BaseConstructionHelper(object):
    params = None
    mesh = None
    cf = None
    ff = None
    product2d = None
    material_to_region = None

    @property
    def p(self):
        return _AttrDict(self.params)

    @property
    def mesh_data(self):
        if getattr(self, '_mesh_data', None) is None:
            self._mesh_data = MeshData(
                mesh=self.mesh,
                cell_function=self.cf,
                facet_function=self.ff,
                region_name_to_cvs=dict(self.cell_regions),
                facets_name_to_fvs=dict(self.facet_regions),
                material_to_region_map=self.material_to_region,
                facets_manager=self.facets,
                mesh_unit=self.params.get('mesh_unit', None),
                product2d=self.product2d)
        return self._mesh_data

    def run(self):
        self.generate_mesh()
        self.compute_facets()
        self.user_extra_definitions()
        self.util_define_internal()
        return self

    def generate_mesh(self):
        """override me"""

    def user_extra_definitions(self):
        """override me"""

    def util_define_internal(self):
        for k, cvs in self.cell_regions.items():
            self.facet_regions['_internal_' + k] = self.facets.internal(cvs)

    def allocate_subdomain_range(self, length):
        start = int(max(self.used_cell_values)) + 1
        self.used_cell_values.update(range(start, start + length))
        return start

    def user_mark_external_boundary_facets(self, mesh, boundary_facet_function):
        """Tag exterior facets with pseudo cell values right/top/left/bottom
        (offset+0..3) and define ``exterior*`` cell regions."""
        cell_regions = self.cell_regions
        exterior = cell_regions['exterior'] = set()
        offset = self.allocate_subdomain_range(4)
        bf = mesh.boundary_facets()
        angle = facet2d_angle_array(mesh.boundary_outward_normals(bf))
        boundary_facet_function[bf] = angle + offset
        for k, v in dict(right=0, top=1, left=2, bottom=3).items():
            cell_regions['exterior_' + k] = {v + offset}
            exterior.add(v + offset)

    def compute_facets(self):
        mesh = self.mesh.init()
        bff = np.full(mesh.num_facets, 1000, dtype=np.int64)
        self.user_mark_external_boundary_facets(mesh, bff)
        self.facets = FacetsManager(mesh, self.cf, bff)
        self.ff = self.facets.facet_mesh_function
        self.facet_regions = {}

    @classmethod
    def from_existing_mesh_cf(cls, mesh, cf, cell_regions, run=True,
                              params=None, mesh_unit=None):
        self = cls()
        self.mesh = mesh
        self.cf = np.asarray(cf, dtype=np.int64)
        self.cell_regions = {k: set(v) for k, v in cell_regions.items()}
        self.used_cell_values = set(np.unique(self.cf).tolist())
        self.params = {} if params is None else params
        if mesh_unit is not None:
            self.params['mesh_unit'] = mesh_unit
        if run:
            self.run()
        return self


class ConstructionHelperIntervalProduct2DMesh(BaseConstructionHelper):
    Interval1DTag = Interval1DTag
    product2d_Ys = (0.0, 1.0)

    def user_define_interval_regions(self):
        return (self.params['domain'], self.params['intervals'])

    @property
    def interval_1d_tag(self):
        if getattr(self, '_i1t', None) is None:
            (x0, x1), intervals = self.user_define_interval_regions()
            clip_intervals(x0, x1, intervals)
            self._i1t = self.Interval1DTag(intervals)
        return self._i1t

    def generate_mesh(self):
        o = self.interval_1d_tag
        o.product2d_Ys = self.params.get('product2d_Ys', self.product2d_Ys)
        pm = o.product2d_mesh
        self.product2d = pm
        self.mesh = pm.mesh
        self.cf = pm.cell_function
        self.cell_regions = {k: set(v) for k, v in
                             o.subdomains['tag_to_cell_values'].items()}
        self.used_cell_values = set(o.subdomains['used_cell_values'])


class ConstructionHelperLayeredStructure(ConstructionHelperIntervalProduct2DMesh):
    """Stack of named layers along x, each with its own meshing rule.

    ``params``: ``edge_length`` (global cap), ``layers`` (dicts with ``name``,
    ``material``, ``thickness``, optional ``mesh`` = ``{type: geometric,
    start, factor}`` or ``{type: constant, edge_length}``),
    ``extra_regions``, ``mesh_unit``; extension: ``product2d_Ys`` for a true
    2D tensor mesh (BASELINE configs 3 and 4).
    """

    @property
    def layers(self):
        return self.params['layers']

    @property
    def extra_regions(self):
        if 'extra_regions' in self.params:
            return self.params['extra_regions']
        warnings.warn('`simple_overmesh_regions` has been renamed to '
                      '`extra_regions`', DeprecationWarning)
        return self.params['simple_overmesh_regions']

    @property
    def material_to_region(self):
        d = {}
        for layer in self.layers:
            region = CellRegion(layer['name'])
            m = layer['material']
            d[m] = region if m not in d else (d[m] | region)
        return d

    @material_to_region.setter
    def material_to_region(self, value):
        pass

    def user_layer_extra_intervals(self, intervals, name_interval):
        """override if necessary"""

    @staticmethod
    def _mkinterval(meshing):
        meshing = dict(meshing)
        kind = meshing.pop('type', None)
        kw = {k: meshing[k] for k in ('x0', 'x1', 'tags')}
        if kind == 'geometric':
            return GeometricallyExpandingMeshInterval(
                edge_length_start=meshing['start'],
                edge_length_expansion_factor=meshing['factor'], **kw)
        if kind == 'constant':
            return CInterval(edge_length=meshing['edge_length'], **kw)
        if kind is None:
            return Interval(**kw)
        raise ValueError('bad meshing strategy {!r}'.format(kind))

    def user_define_interval_regions(self):
        intervals = [CInterval(-np.inf, np.inf,
                               edge_length=self.params['edge_length'])]
        name_interval = {}
        x00 = 0
        x1 = x00
        for layer in self.layers:
            x0, x1 = x1, x1 + layer['thickness']
            if 'mesh' in layer:
                meshing = dict(layer['mesh'])
            elif 'edge_length' in layer:
                warnings.warn('`edge_length` should live under the layer key '
                              '`mesh`', DeprecationWarning)
                meshing = dict(type='constant', edge_length=layer['edge_length'])
            else:
                meshing = dict(type=None)
            meshing.update(x0=x0, x1=x1, tags=(layer['name'], layer['material']))
            iv = self._mkinterval(meshing)
            intervals.append(iv)
            name_interval[layer['name']] = iv
        domain_extent = (x00, x1)

        def locate(label):
            name, offset = label
            if name[0] == '-':
                return name_interval[name[1:]].x0 + offset
            if name[0] == '+':
                return name_interval[name[1:]].x1 + offset
            raise ValueError('must start with +/-')

        for meshing in self.extra_regions:
            meshing = dict(meshing)
            for k in ('x0', 'x1'):
                meshing[k] = locate(meshing[k])
            meshing.setdefault('tags', ())
            if 'edge_length' in meshing and 'type' not in meshing:
                meshing['type'] = 'constant'
            intervals.append(self._mkinterval(meshing))

        self.user_layer_extra_intervals(intervals, name_interval)
        intervals.append(Interval(domain_extent[0], domain_extent[1],
                                  tags=('domain',)))
        return (domain_extent, intervals)
