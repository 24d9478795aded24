"""Spatially dependent parameters and boundary conditions.

API mirror of ``simudo/fem/spatial.py`` (``Spatial.add_rule`` :395,
``add_BC`` :378, ``add_material_data`` :428, ``get`` :350).  Where the
reference compiles rules into a UFL conditional tree on the DG0 tag function
(``fem/expr.py:106-153``), here ``get_table`` resolves them to one value per
cell value (missing rule => 0.0, the reference's default) -- the rows of the
per-tag parameter table the kernels read.
"""
from collections import defaultdict

import numpy as np

from ..util import SetattrInitMixin
from .expr import Expr, as_expr

__all__ = ['BoundaryCondition', 'ValueRule', 'Spatial']


class BoundaryCondition(SetattrInitMixin):
    """``region`` (FacetRegion), ``value``, ``priority`` (lower wins)."""
    require_natural = False
    internal = False
    external = True
    should_extract = True
    priority = 0


class ValueRule(SetattrInitMixin):
    priority = 0


class Spatial(SetattrInitMixin):
    def __init__(self, **kwargs):
        super().__init__(**kwargs)
        self.value_rules = defaultdict(list)
        self.bcs = defaultdict(list)

    @property
    def mesh_data(self):
        return self.parent.mesh_data

    @property
    def material_to_region_map(self):
        return self.mesh_data.material_to_region_map

    # -- rules -----------------------------------------------------------------
    def add_rule(self, key, region, value, priority=0):
        if isinstance(region, str):
            raise AssertionError('`region` must not be a string; arguments swapped?')
        self.value_rules[key].append(ValueRule(region=region, value=value,
                                               priority=priority))
        self.value_rules[key].sort(key=lambda r: r.priority)

    def add_value_rule(self, region, key, value, priority=0):
        return self.add_rule(key=key, region=region, value=value, priority=priority)

    def add_material_data(self, key, material_name, value, priority=0):
        region = self.material_to_region_map.get(material_name, None)
        if region is None:
            return
        self.add_rule(key=key, region=region, value=value, priority=priority)

    def has_rule(self, key):
        return len(self.value_rules.get(key, ())) > 0

    def get_table(self, key, unit=None, default=0.0, required=False):
        """{cell value: float magnitude in ``unit``} for every cell value of the
        mesh; first matching rule in priority order wins (``spatial.py:353-358``)."""
        rules = self.value_rules.get(key, ())
        if not rules and required:
            raise AssertionError('no value rules for key {!r}'.format(key))
        cvs = sorted(set(np.unique(self.mesh_data.cell_function).tolist()))
        table = {cv: float(default) for cv in cvs}
        seen = set()
        for rule in rules:
            for cv in self.mesh_data.evaluate_topology(rule.region):
                if cv in seen or cv not in table:
                    continue
                seen.add(cv)
                v = rule.value
                if hasattr(v, 'm_as'):
                    v = v.m_as(unit) if unit is not None else v.magnitude
                table[cv] = float(v)
        return table

    def get(self, key):
        """Region-wise constant as {cell value: value} (reference returns a UFL
        expression; the tables are what the kernels consume)."""
        return self.get_table(key, required=True)

    # -- boundary conditions ------------------------------------------------------
    def add_BC(self, key, region, value, priority=0, require_natural=False,
               internal=False, external=True, **kwargs):
        if isinstance(region, str):
            raise AssertionError('`region` must not be a string; arguments swapped?')
        bc = BoundaryCondition(region=region, value=value, priority=priority,
                               require_natural=require_natural, internal=internal,
                               external=external, **kwargs)
        self.bcs[key].append(bc)
        self.bcs[key].sort(key=lambda b: b.priority)
        return bc

    def facet_value_map(self, key):
        """{facet value: BoundaryCondition}; a facet value claimed in one
        orientation is not re-claimed in the other (``spatial.py:107-126``)."""
        mapping = {}
        for bc in self.bcs.get(key, ()):
            for fv, sign in self.mesh_data.evaluate_topology(bc.region):
                if fv not in mapping:
                    mapping[fv] = bc
        return mapping

    def bc_facets(self, key, exterior_only=False):
        """[(BoundaryCondition, facet indices)] for the BCs registered under key."""
        ff = np.asarray(self.mesh_data.facet_function)
        fmap = self.facet_value_map(key)
        groups = defaultdict(list)
        for fv, bc in fmap.items():
            groups[id(bc)].append((bc, fv))
        out = []
        mesh = self.mesh_data.mesh.init()
        for items in groups.values():
            bc = items[0][0]
            fvs = [fv for _, fv in items]
            facets = np.nonzero(np.isin(ff, fvs))[0]
            if exterior_only:
                facets = facets[mesh.facet_cells[facets, 1] < 0]
            out.append((bc, facets))
        return out

    def check_no_overlap(self, essential_key, natural_key):
        e = set(self.facet_value_map(essential_key))
        n = set(self.facet_value_map(natural_key))
        if e & n:
            raise AssertionError(
                'Essential and natural boundary conditions overlap. Keys {!r} and {!r}.'
                .format(essential_key, natural_key))
