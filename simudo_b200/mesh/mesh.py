"""Triangle mesh container (numpy) with dolfin-style entity numbering.

Stands in for ``dolfin.Mesh`` as used by the reference
(``simudo/mesh/product2d.py:25-65``).  Conventions reproduced
(upstream dolfin, not verifiable here):

* vertices / cells keep exactly the indices handed in by the builder;
* ``order()``: every cell lists its vertices in increasing global index;
* ``init()``: facets (edges) are numbered by lexicographic order of their
  sorted vertex pair (dolfin >= 2018.1 key-matching entity computation);
* local facet ``i`` of a cell is the edge opposite local vertex ``i``;
* ``facet_cells[f]`` lists the incident cells in increasing cell index, so
  the '+' side of an interior facet is the lower-numbered cell (Appendix A8).

Everything is vectorised: a 1M-cell mesh builds in about a second.
"""
import numpy as np

__all__ = ['Mesh']


class Mesh(object):
    def __init__(self, coordinates, cells):
        self.coordinates = np.ascontiguousarray(coordinates, dtype=np.float64)
        cells = np.asarray(cells, dtype=np.int64)
        # mesh.order(): vertex-sorted cells (detJ may become negative)
        self.cells = np.ascontiguousarray(np.sort(cells, axis=1).astype(np.int32))
        self._init_done = False

    # -- sizes -----------------------------------------------------------------
    @property
    def num_vertices(self):
        return self.coordinates.shape[0]

    @property
    def num_cells(self):
        return self.cells.shape[0]

    @property
    def num_facets(self):
        self.init()
        return self.facet_vertices.shape[0]

    # -- connectivity ------------------------------------------------------------
    def init(self):
        if self._init_done:
            return self
        c = self.cells.astype(np.int64)
        nv = self.num_vertices
        # local facet i is opposite vertex i: (v1,v2), (v0,v2), (v0,v1)
        lo = np.stack([c[:, 1], c[:, 0], c[:, 0]], axis=1)
        hi = np.stack([c[:, 2], c[:, 2], c[:, 1]], axis=1)
        key = (lo * nv + hi).ravel()
        ukey, inv = np.unique(key, return_inverse=True)
        nf = ukey.shape[0]
        self.cell_facets = np.ascontiguousarray(
            inv.reshape(-1, 3).astype(np.int32))
        self.facet_vertices = np.stack([ukey // nv, ukey % nv], axis=1).astype(np.int32)
        # facet -> cells (sorted by cell index), -1 pads boundary facets
        nc = self.num_cells
        cell_of = np.repeat(np.arange(nc, dtype=np.int64), 3)
        local_of = np.tile(np.arange(3, dtype=np.int64), nc)
        order = np.lexsort((cell_of, inv))
        f_sorted = inv[order]
        first = np.ones(f_sorted.shape[0], dtype=bool)
        first[1:] = f_sorted[1:] != f_sorted[:-1]
        fc = np.full((nf, 2), -1, dtype=np.int32)
        fl = np.full((nf, 2), -1, dtype=np.int8)
        fc[f_sorted[first], 0] = cell_of[order][first]
        fl[f_sorted[first], 0] = local_of[order][first]
        fc[f_sorted[~first], 1] = cell_of[order][~first]
        fl[f_sorted[~first], 1] = local_of[order][~first]
        self.facet_cells = fc
        self.facet_local = fl          # local facet number inside each incident cell
        self._init_done = True
        return self

    # -- geometry ----------------------------------------------------------------
    def cell_jacobians(self):
        """J[c] = [x1-x0, x2-x0] (columns), detJ signed (Appendix A1)."""
        x = self.coordinates[self.cells]          # (Nc, 3, 2)
        J = np.empty((self.num_cells, 2, 2))
        J[:, :, 0] = x[:, 1] - x[:, 0]
        J[:, :, 1] = x[:, 2] - x[:, 0]
        det = J[:, 0, 0] * J[:, 1, 1] - J[:, 0, 1] * J[:, 1, 0]
        return J, det

    def facet_midpoints(self):
        self.init()
        return self.coordinates[self.facet_vertices].mean(axis=1)

    def facet_lengths(self):
        self.init()
        x = self.coordinates[self.facet_vertices]
        return np.linalg.norm(x[:, 1] - x[:, 0], axis=1)

    def boundary_facets(self):
        self.init()
        return np.nonzero(self.facet_cells[:, 1] < 0)[0]

    def boundary_outward_normals(self, facets):
        """Unit outward normals of boundary facets (first incident cell)."""
        self.init()
        x = self.coordinates[self.facet_vertices[facets]]
        t = x[:, 1] - x[:, 0]
        n = np.stack([t[:, 1], -t[:, 0]], axis=1)
        n /= np.linalg.norm(n, axis=1)[:, None]
        cells = self.facet_cells[facets, 0]
        centroid = self.coordinates[self.cells[cells]].mean(axis=1)
        mid = x.mean(axis=1)
        flip = np.einsum('ij,ij->i', n, mid - centroid) < 0
        n[flip] *= -1
        return n

    def cell_to_cells_adjacency_via_facet(self):
        """Same content as ``MeshData.cell_to_cells_adjacency_via_facet``
        (``simudo/fem/mesh_data.py:62-77``) as a padded (Nc, 3) array."""
        self.init()
        nb = np.full((self.num_cells, 3), -1, dtype=np.int32)
        fc = self.facet_cells
        interior = fc[:, 1] >= 0
        for side in (0, 1):
            c = fc[interior, side]
            o = fc[interior, 1 - side]
            l = self.facet_local[interior, side]
            nb[c, l] = o
        return nb
