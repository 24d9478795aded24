"""DOF maps and the fixed CSR structure of the mixed DG1 x BDM2 space.

Space layout (``poisson_drift_diffusion.py:115-123``,
``poisson_drift_diffusion1.py1:59-72,412-425``): ``P`` sub-problems
(Poisson first, then the bands in insertion order), each a ``[DG1, BDM2]``
pair, so a cell has ``L = 15 P`` local dofs ordered
``15 p + (0..2 DG1 | 3 + 3 f + k facet f | 12 + k interior)``.

Two global numberings are offered (hard part 1):

``'ufc'``   dolfin's *un-reordered* UFC numbering: sub-element by
            sub-element; DG1 -> ``3 cell + i``; BDM2 -> facet dofs
            ``3 facet + k`` then interior ``3 Nf + 3 cell + k``.
            dolfin additionally applies a graph reordering that cannot be
            reproduced here; comparing against a real dolfin dump needs the
            permutation from ``tools/dump_dolfin_reference.py``.
``'b200'``  cell-blocked: the ``6 P`` cell-interior dofs of a cell are
            contiguous, the ``3 P`` dofs of a facet are contiguous after all
            cell blocks.  Same space, same matrix up to that permutation;
            rows then consist of a few long contiguous column runs, which is
            what the scatter stage of the CUDA kernel wants.

The sparsity pattern is dolfin's (Appendix A11): for every cell the full
``L x L`` product of its dofs.  Because DG1 and BDM2 have no vertex dofs, a
row is shared by at most two cells, so the position of cell-local column
``j`` inside a row is a small per-cell rank table instead of an ``L x L``
position map.
"""
import numpy as np

__all__ = ['local_layout', 'build_cell_dofs', 'CsrStructure']


def local_layout(P):
    """Index helpers for the local dof vector of a cell."""
    L = 15 * P
    dg = np.array([[15 * p + i for i in range(3)] for p in range(P)])
    bdm = np.array([[15 * p + 3 + i for i in range(12)] for p in range(P)])
    fac = np.array([[[15 * p + 3 + 3 * f + k for k in range(3)]
                     for f in range(3)] for p in range(P)])
    interior = np.array([[15 * p + 12 + k for k in range(3)] for p in range(P)])
    return dict(L=L, dg=dg, bdm=bdm, facet=fac, interior=interior)


def build_cell_dofs(mesh, P, numbering='b200'):
    """Return ``(cell_dofs[Nc, 15P] int32, N)``."""
    mesh.init()
    nc, nf = mesh.num_cells, mesh.num_facets
    L = 15 * P
    if 6 * P * nc + 3 * P * nf >= 2 ** 31:
        raise ValueError('dof count exceeds int32')
    cd = np.empty((nc, L), dtype=np.int64)
    c = np.arange(nc, dtype=np.int64)
    cf = mesh.cell_facets.astype(np.int64)
    if numbering == 'ufc':
        base = 0
        for p in range(P):
            for i in range(3):
                cd[:, 15 * p + i] = base + 3 * c + i
            base += 3 * nc
            for f in range(3):
                for k in range(3):
                    cd[:, 15 * p + 3 + 3 * f + k] = base + 3 * cf[:, f] + k
            for k in range(3):
                cd[:, 15 * p + 12 + k] = base + 3 * nf + 3 * c + k
            base += 3 * nf + 3 * nc
    elif numbering == 'b200':
        fbase = 6 * P * nc
        for p in range(P):
            for i in range(3):
                cd[:, 15 * p + i] = 6 * P * c + 6 * p + i
            for k in range(3):
                cd[:, 15 * p + 12 + k] = 6 * P * c + 6 * p + 3 + k
            for f in range(3):
                for k in range(3):
                    cd[:, 15 * p + 3 + 3 * f + k] = fbase + 3 * P * cf[:, f] + 3 * p + k
    else:
        raise ValueError(numbering)
    N = 6 * P * nc + 3 * P * nf
    return np.ascontiguousarray(cd.astype(np.int32)), N


class CsrStructure(object):
    """Row pointers + per-cell rank tables of the dolfin-pattern CSR.

    ``rowptr[N+1]`` int64; ``rank_own[Nc, L]`` uint8 = position of local
    column j in a row owned by this cell alone; ``rank_fac[Nc, 3, L]`` uint8 =
    position of local column j in the rows of the dofs on local facet f
    (shared with the neighbour across f, if any).  Position in ``vals`` of
    local entry (i, j) of cell c is ``rowptr[cell_dofs[c, i]] + rank``.
    ``patterns[npat, 4, L]`` / ``cell_pattern[Nc]``: the same tables
    de-duplicated (class 0 = own rows, 1..3 = rows on local facet 0..2).
    """

    def __init__(self, mesh, cell_dofs, P, N):
        mesh.init()
        self.P, self.N = P, N
        L = 15 * P
        self.L = L
        nc = mesh.num_cells
        cd = np.asarray(cell_dofs)
        lay = local_layout(P)
        is_facet_local = np.zeros(L, dtype=bool)
        is_facet_local[lay['facet'].ravel()] = True

        # rows owned by a single cell: rank within the cell's sorted dofs
        order = np.argsort(cd, axis=1, kind='stable')
        rank_own = np.empty((nc, L), dtype=np.uint8)
        np.put_along_axis(rank_own, order,
                          np.broadcast_to(np.arange(L, dtype=np.uint8), (nc, L)), axis=1)

        rank_fac = np.repeat(rank_own[:, None, :], 3, axis=1)
        fc = mesh.facet_cells
        fl = mesh.facet_local
        interior = np.nonzero(fc[:, 1] >= 0)[0]
        rowlen = np.full(N, L, dtype=np.int64)
        if interior.size:
            c0, c1 = fc[interior, 0], fc[interior, 1]
            both = np.concatenate([cd[c0], cd[c1]], axis=1)         # [nfi, 2L]
            o = np.argsort(both, axis=1, kind='stable')
            vs = np.take_along_axis(both, o, axis=1)
            newv = np.ones(vs.shape, dtype=np.int64)
            newv[:, 0] = 0
            newv[:, 1:] = vs[:, 1:] != vs[:, :-1]
            rs = np.cumsum(newv, axis=1)
            r = np.empty_like(rs)
            np.put_along_axis(r, o, rs, axis=1)
            ncols = rs[:, -1] + 1
            rank_fac[c0, fl[interior, 0].astype(np.int64), :] = r[:, :L].astype(np.uint8)
            rank_fac[c1, fl[interior, 1].astype(np.int64), :] = r[:, L:].astype(np.uint8)
            # row lengths of the shared facet dofs
            for p in range(P):
                for k in range(3):
                    rows = cd[c0, lay['facet'][p, :, k][fl[interior, 0].astype(np.int64)]]
                    rowlen[rows] = ncols
        self.rank_own = rank_own
        self.rank_fac = np.ascontiguousarray(rank_fac)
        # cells with identical rank tables share one pattern (a product mesh has
        # a handful): the kernels look ranks up in this small table
        tab = np.concatenate([rank_own[:, None, :], self.rank_fac], axis=1).reshape(nc, 4 * L)
        pats, inv = np.unique(tab, axis=0, return_inverse=True)
        self.patterns = np.ascontiguousarray(pats.reshape(-1, 4, L))
        self.cell_pattern = np.ascontiguousarray(inv.reshape(-1).astype(np.int32))
        self.rowptr = np.zeros(N + 1, dtype=np.int64)
        np.cumsum(rowlen, out=self.rowptr[1:])
        self.nnz = int(self.rowptr[-1])
        self._mesh = mesh
        self._cd = cd

    def column_indices(self):
        """Materialise ``colidx[nnz]`` (int32) -- for export / solvers / tests."""
        cd, L = self._cd, self.L
        P = self.P
        lay = local_layout(P)
        col = np.full(self.nnz, -1, dtype=np.int32)
        nc = cd.shape[0]
        is_fac = np.zeros(L, dtype=np.int64) - 1
        for f in range(3):
            is_fac[lay['facet'][:, f, :].ravel()] = f
        for i in range(L):
            rows = cd[:, i].astype(np.int64)
            f = is_fac[i]
            ranks = self.rank_own if f < 0 else self.rank_fac[:, f, :]
            pos = self.rowptr[rows][:, None] + ranks.astype(np.int64)
            col[pos.ravel()] = cd.ravel()
        assert (col >= 0).all()
        return col

    def entry_positions(self, cells=None):
        """pos[c, i, j] into vals (small meshes / tests only)."""
        cd, L, P = self._cd, self.L, self.P
        lay = local_layout(P)
        if cells is None:
            cells = np.arange(cd.shape[0])
        pos = np.empty((len(cells), L, L), dtype=np.int64)
        is_fac = np.zeros(L, dtype=np.int64) - 1
        for f in range(3):
            is_fac[lay['facet'][:, f, :].ravel()] = f
        for i in range(L):
            f = is_fac[i]
            ranks = self.rank_own[cells] if f < 0 else self.rank_fac[cells, f, :]
            pos[:, i, :] = self.rowptr[cd[cells, i].astype(np.int64)][:, None] + ranks
        return pos
