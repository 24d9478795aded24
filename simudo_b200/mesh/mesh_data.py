"""``MeshData``: mesh + cell/facet markers + region dictionaries.

Same attributes and ``evaluate_topology`` contract as
``simudo/fem/mesh_data.py:9-58``; ``cell_function`` / ``facet_function`` are
plain integer numpy arrays indexed by cell / facet number.
"""
from ..util import SetattrInitMixin

__all__ = ['MeshData']


class MeshData(SetattrInitMixin):
    mesh = None
    cell_function = None
    facet_function = None
    region_name_to_cvs = None
    facets_name_to_fvs = None
    facets_manager = None
    material_to_region_map = None
    mesh_unit = None
    product2d = None     # set when the mesh is an Xs x Ys product mesh

    def evaluate_topology(self, region):
        """CellRegion -> set of cell values; FacetRegion -> set of (fv, sign)."""
        return region.evaluate(dict(
            region_name_to_cvs=self.region_name_to_cvs,
            facets_name_to_fvs=self.facets_name_to_fvs,
            facets_manager=self.facets_manager))

    @property
    def cell_to_cells_adjacency_via_facet(self):
        return self.mesh.cell_to_cells_adjacency_via_facet()
