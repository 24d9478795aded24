from .mesh import Mesh
from .product2d import Product2DMesh
from .interval1dtag import (clip_intervals, Interval, CInterval,
                            GeometricallyExpandingMeshInterval,
                            BaseInterval1DTag, Interval1DTag)
from .facet import FacetsManager, facet2d_angle_array
from .topology import CellRegion, FacetRegion, CellRegions, FacetRegions
from .mesh_data import MeshData
from .construction_helper import (BaseConstructionHelper,
                                  ConstructionHelperIntervalProduct2DMesh,
                                  ConstructionHelperLayeredStructure)

__all__ = ['Mesh', 'Product2DMesh', 'clip_intervals', 'Interval', 'CInterval',
           'GeometricallyExpandingMeshInterval', 'BaseInterval1DTag',
           'Interval1DTag', 'FacetsManager', 'facet2d_angle_array',
           'CellRegion', 'FacetRegion', 'CellRegions', 'FacetRegions',
           'MeshData', 'BaseConstructionHelper',
           'ConstructionHelperIntervalProduct2DMesh',
           'ConstructionHelperLayeredStructure']
