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
``'b200'``  sub-problem major, entity blocked: for every sub-problem ``p`` the 6
            cell-interior dofs of cell ``c`` are ``6 Nc p + 6 c + (0..2 DG1 |
            3..5 interior BDM2)``; after all cell dofs the 3 dofs of facet ``f``
            are ``6 P Nc + 3 Nf p + 3 f + k``.  Same space, same matrix up to
            that permutation.  With CSR values stored in row order this makes
            everything ONE sub-problem's rows hold a contiguous range of
            ``vals``: the row kernel of sub-problem ``p`` (one launch) then
            writes a dense stream instead of every fourth ~1 KB piece of a
            cell-major array (measured on B200: 4.3 -> 3.4 ms per 1M cells for
            the row kernels; partially written DRAM pages cost bandwidth).

The sparsity pattern is dolfin's (Appendix A11): for every cell the full
``L x L`` product of its dofs.  Because DG1 and BDM2 have no vertex dofs, a
row is shared by at most two cells, so the position of cell-local column
``j`` inside a row is a small per-cell rank table instead of an ``L x L``
position map.
"""
import numpy as np

__all__ = ['local_layout', 'build_cell_dofs', 'structural_mask', 'CsrStructure']


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
                cd[:, 15 * p + i] = 6 * nc * p + 6 * c + i
            for k in range(3):
                cd[:, 15 * p + 12 + k] = 6 * nc * p + 6 * c + 3 + k
            for f in range(3):
                for k in range(3):
                    cd[:, 15 * p + 3 + 3 * f + k] = fbase + 3 * nf * p + 3 * cf[:, f] + k
    else:
        raise ValueError(numbering)
    N = 6 * P * nc + 3 * P * nf
    return np.ascontiguousarray(cd.astype(np.int32)), N


def structural_mask(P):
    """S[i, j] = True iff local entry (i, j) of the element tensor can be
    non-zero for the Poisson / drift-diffusion form:

    * w   rows (Poisson DG1):  E, phi, every delta_w_k
    * psi rows (Poisson BDM2): E, phi
    * v^k rows (band DG1):     j_k, phi, every delta_w_l
    * xi^k rows (band BDM2):   j_k, delta_w_k, phi
    """
    L = 15 * P
    S = np.zeros((L, L), dtype=bool)
    dg = lambda p: slice(15 * p, 15 * p + 3)
    bdm = lambda p: slice(15 * p + 3, 15 * p + 15)
    for p in range(P):
        S[dg(p), bdm(p)] = True
        S[dg(p), dg(0)] = True
        for q in range(1, P):
            S[dg(p), dg(q)] = True
        S[bdm(p), bdm(p)] = True
        S[bdm(p), dg(p)] = True
        S[bdm(p), dg(0)] = True
    return S


class CsrStructure(object):
    """Row pointers + rank patterns of the fixed CSR.

    ``pattern='dolfin'``     dolfin's sparsity (Appendix A11): the full L x L
                             product of every cell's dofs.
    ``pattern='structural'`` the same matrix with the structural zeros of the
                             form dropped (``structural_mask``): what the
                             device path stores natively -- rows become
                             contiguous runs of values that are all written.

    ``rowptr[N+1]`` int64.  The position in ``vals`` of local entry (i, j) of
    cell c is ``rowptr[cell_dofs[c, i]] + patterns[cell_pattern[c], i, j]``
    (0xFF = absent).  Because DG1 and BDM2 have no vertex dofs a row is shared
    by at most two cells, and the rank tables of all cells fall into a
    handful of distinct patterns (8 on a product mesh).
    """

    def __init__(self, mesh, cell_dofs, P, N, pattern='dolfin'):
        mesh.init()
        self.P, self.N, self.pattern = P, N, pattern
        L = 15 * P
        self.L = L
        nc = mesh.num_cells
        cd = np.asarray(cell_dofs).astype(np.int64)
        self._cd = cd
        self._mesh = mesh
        if pattern == 'dolfin':
            S = np.ones((L, L), dtype=bool)
        elif pattern in ('structural', 'native', 'bsr3'):
            S = structural_mask(P)
        else:
            raise ValueError(pattern)
        self.mask = S
        nbr = mesh.cell_to_cells_adjacency_via_facet().astype(np.int64)   # [Nc, 3]
        # local facet number of the shared facet inside the neighbour
        fl_of = np.full((nc, 3), -1, dtype=np.int64)
        fc, fl = mesh.facet_cells, mesh.facet_local.astype(np.int64)
        interior = fc[:, 1] >= 0
        fl_of[fc[interior, 0], fl[interior, 0]] = fl[interior, 1]
        fl_of[fc[interior, 1], fl[interior, 1]] = fl[interior, 0]
        self._nbr, self._fl_of = nbr, fl_of

        rowlen = np.zeros(N, dtype=np.int64)
        sig = np.zeros(nc, dtype=np.uint64)
        mult = np.uint64(0x9E3779B97F4A7C15)
        allc = np.arange(nc)
        with np.errstate(over='ignore'):
            for i in range(L):
                rep = self._row_type_representative(i)
                if rep != i:
                    # the three dofs of one (sub-problem, entity) share one column set
                    rowlen[cd[:, i]] = rowlen[cd[:, rep]]
                    continue
                ranks, nrow = self._ranks_of_row(i, allc)
                rowlen[cd[:, i]] = nrow
                h = np.zeros(nc, dtype=np.uint64)
                for col in range(ranks.shape[1]):
                    h = h * mult + ranks[:, col].astype(np.uint64) + np.uint64(1)
                sig = sig * np.uint64(0xC2B2AE3D27D4EB4F) + h
        usig, first, inv = np.unique(sig, return_index=True, return_inverse=True)
        self.cell_pattern = np.ascontiguousarray(inv.reshape(-1).astype(np.int32))
        pats = np.full((len(usig), L, L), 0xFF, dtype=np.uint8)
        reps = first.astype(np.int64)
        for i in range(L):
            ranks, _ = self._ranks_of_row(i, reps)
            pats[:, i, S[i]] = ranks.astype(np.uint8)
        self.rowptr = np.zeros(N + 1, dtype=np.int64)
        np.cumsum(rowlen, out=self.rowptr[1:])
        self.nnz = int(self.rowptr[-1])
        self.csr_rowptr = self.rowptr            # true CSR row pointers (row lengths)
        if pattern == 'bsr3':
            # Block CSR with 3 x 3 blocks (every dof triple of a mesh entity is a block row /
            # block column, every structural block is full): block rows in dof order, block
            # columns in the 'native' emission order, each block row-major.  Same entries and
            # nnz as 'native'; the value of local entry (i, j) sits at
            #   rowptr[dof_i] + patterns[..][i][j]
            # with rowptr[3 T + k] = (first value of block row T) + 3 k and
            # patterns = 9 * (block rank) + (j within its triple) -- the form the C ABI takes.
            if self.rowptr[-1] >= 2 ** 62 or N % 3:
                raise ValueError('bsr3 needs dof triples')
            base = self.rowptr[0:N:3]
            rp = np.empty(N + 1, dtype=np.int64)
            for k in range(3):
                rp[k:N:3] = base + 3 * k
            rp[N] = self.nnz
            self.rowptr = rp
            ok = pats != 0xFF
            pb = (pats.astype(np.int64) // 3) * 9 + pats.astype(np.int64) % 3
            assert pb[ok].max() < 0xFF
            pats = np.where(ok, pb, 0xFF).astype(np.uint8)
        self.patterns = np.ascontiguousarray(pats)
        if self.patterns.shape[0] > 4096:
            raise ValueError('too many distinct rank patterns for this mesh/numbering')

    @staticmethod
    def _row_type_representative(i):
        """First local dof of the (sub-problem, entity) group local dof i belongs to."""
        p, r = divmod(i, 15)
        return 15 * p + (0 if r < 3 else 3 + 3 * ((r - 3) // 3))

    def _row_facet(self, i):
        r = i % 15
        return (r - 3) // 3 if 3 <= r < 12 else -1

    def _native_ranks_of_row(self, i, cells):
        """``pattern='native'``: same entries as 'structural', columns of a row in the
        *emission order* of the device kernels instead of ascending, so that every
        cell's contribution to a row is one contiguous run (csrc/native.inl):

        * DG1 row of sub-problem p:       [BDM(p) m = 0..11 | DG(q) l = 0..2, q = 0..P-1]
        * interior BDM row:               [BDM(p) m = 0..11 | DG(p) | DG(0) if p > 0]
        * facet BDM row (facet f, dof k): [the 3 dofs of f | run A | run B], run X =
          [the 9 other BDM(p) dofs of cell X, ascending m | DG(p) of X | DG(0) of X if p > 0];
          A = ``facet_cells[f, 0]`` (the owner, lower cell index), B the other cell
          (absent on the boundary)."""
        S = self.mask
        cols = np.nonzero(S[i])[0]
        p, r = divmod(i, 15)
        f = self._row_facet(i)
        n = len(cells)
        rk = np.empty(len(cols), dtype=np.int64)
        nrun = 12 if p == 0 else 15                 # length of run A / B of a facet row
        for a, j in enumerate(cols):
            q, rj = divmod(j, 15)
            if rj >= 3:                             # BDM column (q == p)
                m = rj - 3
                if f < 0:
                    rk[a] = m
                elif m // 3 == f:
                    rk[a] = m - 3 * f
                else:
                    rk[a] = 3 + (m if m < 3 * f else m - 3)
            elif r < 3:                             # DG row: DG(q) for every q
                rk[a] = 12 + 3 * q + rj
            else:                                   # BDM row: DG(p), then DG(0)
                off = 12 if f < 0 else 3 + 9
                rk[a] = off + rj + (3 if (q == 0 and p > 0) else 0)
        ranks = np.broadcast_to(rk, (n, len(cols))).copy()
        if f < 0:
            return ranks, np.full(n, len(cols), dtype=np.int64)
        nb = self._nbr[cells, f]
        has = nb >= 0
        side_b = has & (nb < cells)                 # this cell is the second cell of the facet
        shift = np.where(rk >= 3, nrun, 0)
        ranks[side_b] += shift[None, :]
        return ranks, np.where(has, 3 + 2 * nrun, 3 + nrun).astype(np.int64)

    def _ranks_of_row(self, i, cells):
        """Rank of each structural column of local row i inside its global row,
        for the given cells: ``(ranks[len(cells), n_i], row_length[len(cells)])``."""
        if self.pattern in ('native', 'bsr3'):
            return self._native_ranks_of_row(i, np.asarray(cells))
        cd, L, S = self._cd, self.L, self.mask
        cols = np.nonzero(S[i])[0]
        own = cd[cells][:, cols]                              # [n, n_i]
        f = self._row_facet(i)
        n = len(cells)
        if f < 0:
            order = np.argsort(own, axis=1, kind='stable')
            ranks = np.empty_like(order)
            np.put_along_axis(ranks, order,
                              np.broadcast_to(np.arange(len(cols)), own.shape), axis=1)
            return ranks, np.full(n, len(cols), dtype=np.int64)
        nb = self._nbr[cells, f]
        has = nb >= 0
        # the same global row seen from the neighbour: same sub-problem and facet
        # dof number, on the neighbour's local facet fl_of
        p, k = i // 15, (i % 15 - 3) % 3
        other = np.full_like(own, np.iinfo(np.int64).max)
        if has.any():
            f2 = self._fl_of[cells[has], f]
            i2 = 15 * p + 3 + 3 * f2 + k
            # structural columns of a BDM row depend on the sub-problem only,
            # up to which facet block is "its own": S rows of one sub-problem's
            # BDM dofs are identical
            assert (S[i2] == S[i]).all()
            other[has] = cd[nb[has]][:, cols]
        both = np.concatenate([own, other], axis=1)            # [n, 2 n_i]
        o = np.argsort(both, axis=1, kind='stable')
        vs = np.take_along_axis(both, o, axis=1)
        newv = np.ones(vs.shape, dtype=np.int64)
        newv[:, 0] = 0
        newv[:, 1:] = vs[:, 1:] != vs[:, :-1]
        rs = np.cumsum(newv, axis=1)
        r = np.empty_like(rs)
        np.put_along_axis(r, o, rs, axis=1)
        valid = vs != np.iinfo(np.int64).max
        nrow = (newv * valid).sum(axis=1) + 1
        return r[:, :len(cols)], nrow

    # -- views used by tests, the oracle driver and exports ----------------------
    def entry_positions(self, cells=None):
        """pos[c, i, j] into vals, -1 where absent (small meshes / tests only)."""
        cd, L = self._cd, self.L
        if cells is None:
            cells = np.arange(cd.shape[0])
        pat = self.patterns[self.cell_pattern[cells]].astype(np.int64)     # [n, L, L]
        pos = self.rowptr[cd[cells]][:, :, None] + pat
        pos[pat == 0xFF] = -1
        return pos

    def column_indices(self, want_rows=False):
        """Materialise ``colidx[nnz]`` (int32), the column of every stored value -- for
        export / solvers / tests.  ``want_rows``: also the row of every stored value."""
        cd, L = self._cd, self.L
        col = np.full(self.nnz, -1, dtype=np.int32)
        row = np.full(self.nnz, -1, dtype=np.int32) if want_rows else None
        nc = cd.shape[0]
        step = max(1, (1 << 22) // (L * L))
        for c0 in range(0, nc, step):
            cells = np.arange(c0, min(nc, c0 + step))
            pos = self.entry_positions(cells)
            cols = np.broadcast_to(cd[cells][:, None, :], pos.shape)
            ok = pos >= 0
            col[pos[ok]] = cols[ok]
            if want_rows:
                rows = np.broadcast_to(cd[cells][:, :, None], pos.shape)
                row[pos[ok]] = rows[ok]
        assert (col >= 0).all()
        return (col, row) if want_rows else col

    def value_rows(self):
        """Row of every stored value (any pattern)."""
        if self.pattern != 'bsr3':
            return np.repeat(np.arange(self.N, dtype=np.int64), np.diff(self.rowptr))
        return self.column_indices(want_rows=True)[1].astype(np.int64)

    def sorted_csr(self):
        """``(rowptr, colidx, perm)``: the same matrix as a CSR with ascending columns;
        ``vals_csr = vals[perm]``.  Cached."""
        if getattr(self, '_sorted', None) is None:
            col = self.column_indices()
            perm = np.lexsort((col, self.value_rows()))
            self._sorted = (self.csr_rowptr, np.ascontiguousarray(col[perm]), perm)
        return self._sorted
