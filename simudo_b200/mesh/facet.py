"""Facet markers between cell-value subdomains (``FacetsManager``).

Restates ``simudo/mesh/facet.py:19-91,188-294`` with numpy: every facet is
keyed by ``(cv_lo, cv_hi, sign)`` where the two cell values are those of the
incident cells (for a boundary facet the "outside" value is the pseudo cell
value of the exterior side), ``sign = +1`` if the first incident cell has the
smaller-or-equal value.  Facet values ``1..n_cv`` are reserved for
"internal to cell value v"; the other keys are numbered in order of first
appearance over facets in index order (``facet.py:55-90``).
"""
import itertools

import numpy as np

__all__ = ['facet2d_angle_array', 'FacetsManager']


def facet2d_angle_array(normals):
    """right=0, up=1, left=2, down=3 (``facet.py:120-127``), vectorised."""
    n = np.asarray(normals)
    a = np.abs(n)
    xsig = a[:, 0] > a[:, 1]
    return np.where(xsig, np.where(n[:, 0] > 0, 0, 2),
                    np.where(n[:, 1] > 0, 1, 3))


class FacetsManager(object):
    def __init__(self, mesh, cell_function, boundary_facet_function):
        mesh.init()
        cfa = np.asarray(cell_function, dtype=np.int64)
        bffa = np.asarray(boundary_facet_function, dtype=np.int64)
        fc = mesh.facet_cells
        x = cfa[fc[:, 0]]
        interior = fc[:, 1] >= 0
        y = np.where(interior, cfa[np.where(interior, fc[:, 1], 0)], bffa)
        sign = np.where(x <= y, 1, -1)
        lo = np.minimum(x, y)
        hi = np.maximum(x, y)

        cell_values = sorted(set(np.unique(cfa).tolist()))
        d = {}
        fv = 1
        for v in cell_values:
            d[(int(v), int(v), 1)] = fv
            fv += 1

        base = int(max(hi.max(), 1)) + 1
        code = (lo * base + hi) * 2 + (sign > 0)
        ucode, first_idx, inv = np.unique(code, return_index=True,
                                          return_inverse=True)
        code_to_fv = np.zeros(ucode.shape[0], dtype=np.int64)
        for k in np.argsort(first_idx, kind='stable'):
            cd = int(ucode[k])
            s = 1 if (cd & 1) else -1
            pair = cd >> 1
            key = (pair // base, pair % base, s)
            v = d.get(key)
            if v is None:
                d[key] = v = fv
                fv += 1
            code_to_fv[k] = v
        self.facet_mesh_function = code_to_fv[inv]
        self.facet_value_dict = d

        self.internal_facets_dict = ifvs = {}
        self.fvs = fvs = {}
        for (cv1, cv2, s), v in d.items():
            if cv1 == cv2:
                ifvs[cv1] = v
            else:
                fvs.setdefault((cv1, cv2), set()).add((v, s))
                fvs.setdefault((cv2, cv1), set()).add((v, -s))

    def boundary(self, X, Y, intersection=None):
        """Signed facet values between cell-value sets X and Y."""
        X, Y = set(X), set(Y)
        if intersection is not None:
            if intersection == 0:
                Y = Y - X
            elif intersection == 1:
                X = X - Y
            else:
                raise ValueError(repr(intersection))
        else:
            assert not (X & Y)
        r = set()
        for x, y in itertools.product(X, Y):
            r.update(self.fvs.get((x, y), ()))
        return r

    def cell_value_internal(self, X):
        ifvs = self.internal_facets_dict
        return set((ifvs[x], 1) for x in X if x in ifvs)

    def internal(self, X):
        r = self.cell_value_internal(X)
        for x, y in itertools.combinations(X, 2):
            r.update(self.fvs.get((x, y), ()))
        return r
