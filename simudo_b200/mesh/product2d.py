"""Tensor-product triangle mesh ``Xs x Ys`` (two triangles per rectangle).

Same vertex / cell layout as ``simudo/mesh/product2d.py:25-65``: vertex
``ix*nY + iy``; rectangle ``(ix, iy)`` is split along its rising diagonal into
cells ``(b, c, c+1)`` and ``(b, c+1, b+1)`` with ``b = ix*nY + iy``,
``c = b + nY``; cells of one x-interval are contiguous
(``cells_at_ix``, pinned by the reference's own test ``product2d.py:116-122``).
"""
import numpy as np

from .mesh import Mesh

__all__ = ['Product2DMesh']


class Product2DMesh(object):
    def __init__(self, Xs, Ys=None):
        if Ys is None:
            Ys = (0.0, 1.0)
        self.Xs = np.asarray(Xs, dtype='double')
        self.Ys = np.asarray(Ys, dtype='double')
        self._mesh = None

    @property
    def nX(self):
        return len(self.Xs)

    @property
    def nY(self):
        return len(self.Ys)

    @property
    def mesh(self):
        if self._mesh is None:
            nX, nY = self.nX, self.nY
            xx, yy = np.meshgrid(self.Xs, self.Ys, indexing='ij')
            coords = np.stack([xx.ravel(), yy.ravel()], axis=1)
            ix, iy = np.meshgrid(np.arange(nX - 1), np.arange(nY - 1), indexing='ij')
            b = (ix * nY + iy).ravel()
            c = b + nY
            cells = np.empty((b.shape[0], 2, 3), dtype=np.int64)
            cells[:, 0] = np.stack([b, c, c + 1], axis=1)
            cells[:, 1] = np.stack([b, c + 1, b + 1], axis=1)
            self._mesh = Mesh(coords, cells.reshape(-1, 3)).init()
        return self._mesh

    def cells_at_ix(self, ix):
        nc = 2 * (self.nY - 1)
        return range(nc * ix, nc * ix + nc)

    def vertices_at_ix(self, ix):
        return range(self.nY * ix, self.nY * ix + self.nY)

    def vertices_at_iy(self, iy):
        return range(iy, self.nY * self.nX, self.nY)

    def ix_from_vertex(self, v):
        return v // self.nY

    def iy_from_vertex(self, v):
        return v % self.nY

    def ix_from_cell(self, c):
        return c // (2 * (self.nY - 1))
