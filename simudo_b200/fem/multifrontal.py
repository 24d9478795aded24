"""On-device sparse direct solve of the Newton systems on genuine 2D meshes.

Replaces ``NewtonSolution.do_linear_solve`` -> ``PETScLUSolver("mumps")``
(``simudo/fem/newton_solver.py:124-181``; MUMPS options ``:140-146``) where the banded
solver of ``csrc/band_solver.cu`` does not apply (BASELINE configs 3 and 4: the band of a
224 x 224 mesh would need a terabyte).  cuDSS is not installed in this image, so this is a
multifrontal LU written for the structure of the mixed DG1 x BDM2 space:

* **Symbolic phase, once per sparsity pattern** (numpy, host; reused by every Newton
  step like MUMPS' analysis, ``newton_solver.py:140-146`` ICNTL(28)): nested dissection of
  the *cells* by recursive coordinate bisection.  Cells couple only through facet dofs, so
  the facets between the two halves of a region are a vertex separator of the dof graph:
  a cell's interior dofs are eliminated in its leaf, a facet's dofs in the lowest tree
  node containing both its cells.  A region eliminated with its boundary fluxes held is a
  pure-Neumann problem -- singular in the constant mode of every DG1 unknown whenever the
  zeroth-order coupling vanishes -- so one DG1 dof per sub-problem and region is delayed
  to the parent front (the static analogue of MUMPS' delayed pivots).
* **Numeric phase, on the device**: the fronts of one tree level that have the same shape
  are one batched dense problem: ``F11 = P L U`` with partial pivoting
  (``torch.linalg.lu_factor`` = cuSOLVER/cuBLAS batched getrf), ``X = F11^-1 F12``,
  Schur complement ``F22 - F21 X`` (batched GEMM), extend-add into the parent fronts with
  one ``index_add_``.  Row/column equilibration as in the banded solver; refinement steps
  with the compensated residual ``sb_csr_residual_dd``.

Dense library kernels (cuBLAS / cuSOLVER through torch) do the arithmetic -- the solve is
not the hot path of this repo (that is the assembly); it is timed separately.
"""
import ctypes as C

import numpy as np
import torch

from .. import _lib

__all__ = ['MultifrontalSolver']


def _dof_entities(pa):
    """For every global dof: (cell, -1) for cell-interior dofs, (-1, facet) for facet dofs,
    and its sub-problem; plus the list of DG1 dofs [Nc, P, 3]."""
    mesh, lay, P = pa.mesh, pa.layout(), pa.P
    N = pa.N
    dcell = np.full(N, -1, dtype=np.int64)
    dfacet = np.full(N, -1, dtype=np.int64)
    cd = pa.cell_dofs.astype(np.int64)
    cells = np.arange(pa.n_cells, dtype=np.int64)
    for p in range(P):
        for i in range(3):
            dcell[cd[:, lay['dg'][p][i]]] = cells
            dcell[cd[:, lay['interior'][p][i]]] = cells
        for f in range(3):
            for k in range(3):
                dfacet[cd[:, lay['facet'][p][f][k]]] = mesh.cell_facets[:, f]
    return dcell, dfacet


class _Tree(object):
    """Nested dissection of the cells: heap-numbered binary tree (root 1)."""

    def __init__(self, pa, leaf_cells):
        mesh = pa.mesh
        xv = pa.cell_vertices
        # key = lower-left corner of the cell's bounding box: both triangles of a quad of a
        # product mesh fall on the same side of every split
        key = xv.min(axis=1)
        nc = pa.n_cells
        self.cell_leaf = np.zeros(nc, dtype=np.int64)
        self.rep = {}                  # node -> representative cell (first cell of its set)
        self.children = {}
        self.leaves = []
        stack = [(1, np.arange(nc, dtype=np.int64))]
        while stack:
            node, cs = stack.pop()
            self.rep[node] = int(cs.min())
            if len(cs) <= leaf_cells:
                self.cell_leaf[cs] = node
                self.leaves.append(node)
                continue
            k = key[cs]
            ext = k.max(axis=0) - k.min(axis=0)
            ax = 0 if ext[0] >= ext[1] else 1
            vals = np.unique(k[:, ax])
            if len(vals) == 1:                      # degenerate: split the other way / by index
                ax = 1 - ax
                vals = np.unique(k[:, ax])
            if len(vals) == 1:
                left = np.zeros(len(cs), dtype=bool)
                left[:len(cs) // 2] = True
            else:
                split = vals[len(vals) // 2]
                left = k[:, ax] < split
            self.children[node] = (2 * node, 2 * node + 1)
            stack.append((2 * node, cs[left]))
            stack.append((2 * node + 1, cs[~left]))
        self.max_depth = max(int(np.log2(n)) for n in self.leaves)

    @staticmethod
    def depth(h):
        return np.floor(np.log2(h)).astype(np.int64)

    def lca(self, a, b):
        a, b = a.copy(), b.copy()
        da, db = self.depth(a), self.depth(b)
        m = da > db
        a[m] >>= (da - db)[m]
        m = db > da
        b[m] >>= (db - da)[m]
        for _ in range(self.max_depth + 1):
            m = a != b
            if not m.any():
                break
            a[m] >>= 1
            b[m] >>= 1
        return a


class MultifrontalSolver(object):
    """``solve(vals, b) -> x`` for the CSR Jacobian of ``pa`` (patterns with contiguous rows:
    'structural', 'native', 'dolfin')."""

    def __init__(self, pa, device, leaf_cells=16, n_refine=3):
        if pa.csr.pattern == 'bsr3':
            raise ValueError('the multifrontal solver takes a CSR value layout '
                             "('structural', 'native' or 'dolfin')")
        self.pa = pa
        self.device = torch.device(device)
        self.n_refine = n_refine
        self.n_ruiz = 5          # sweeps over all non-zeros: the cost of the scaling grows with this
        self.N = pa.N
        self._symbolic(pa, leaf_cells)

    # ------------------------------------------------------------------ symbolic phase
    def _symbolic(self, pa, leaf_cells):
        N, P = pa.N, pa.P
        mesh = pa.mesh
        tree = _Tree(pa, leaf_cells)
        dcell, dfacet = _dof_entities(pa)
        # elimination node of every dof
        fc = mesh.facet_cells.astype(np.int64)
        leaf_of = tree.cell_leaf
        f_node = leaf_of[fc[:, 0]].copy()
        two = fc[:, 1] >= 0
        f_node[two] = tree.lca(leaf_of[fc[two, 0]], leaf_of[fc[two, 1]])
        enode = np.where(dcell >= 0, leaf_of[np.maximum(dcell, 0)], f_node[np.maximum(dfacet, 0)])
        # delayed DG1 dofs: dof 0 of every sub-problem of the representative cell of a node
        # leaves the node (and every ancestor it also represents) for the first ancestor
        # with another representative
        lay = pa.layout()
        rep_nodes = {}
        for node, c in tree.rep.items():
            rep_nodes.setdefault(c, []).append(node)
        for c, nodes in rep_nodes.items():
            top = min(nodes)                          # highest node this cell represents
            target = top >> 1 if top > 1 else 1
            # only if the cell really is the representative of its own leaf
            if leaf_of[c] in nodes:
                for p in range(P):
                    enode[pa.cell_dofs[c, lay['dg'][p][0]]] = target
        self.enode = enode
        # structure in storage order
        col, row = pa.csr.column_indices(want_rows=True)
        row = row.astype(np.int64)
        col = col.astype(np.int64)
        er, ec = enode[row], enode[col]
        # the entry is assembled in the front of the deeper of the two nodes (ancestors of
        # one another by construction of the separators)
        anode = np.maximum(er, ec)
        dr, dc = tree.depth(er), tree.depth(ec)
        shallow = np.where(dr <= dc, er, ec)
        deep_d = np.maximum(dr, dc)
        if not ((anode >> (deep_d - np.minimum(dr, dc))) == shallow).all():
            raise RuntimeError('nested dissection produced a non-separating split')
        # nodes, bottom-up
        nodes = sorted(set(enode.tolist()) | set(tree.rep.keys()), reverse=True)
        node_index = {n: i for i, n in enumerate(nodes)}
        # E sets
        order = np.argsort(enode, kind='stable')
        en_sorted = enode[order]
        bounds = np.searchsorted(en_sorted, np.array(nodes), side='left')
        bounds_r = np.searchsorted(en_sorted, np.array(nodes), side='right')
        E = {n: order[a:b] for n, a, b in zip(nodes, bounds, bounds_r)}
        # adjacency of E(v) reaching ancestors
        oa = np.argsort(anode, kind='stable')
        an_sorted = anode[oa]
        a0 = np.searchsorted(an_sorted, np.array(nodes), side='left')
        a1 = np.searchsorted(an_sorted, np.array(nodes), side='right')
        B = {}
        kids = {}
        for n in nodes:
            if n > 1:
                kids.setdefault(n >> 1, []).append(n)
        # a node may have been skipped in `nodes` if it has no dofs and no children; handle
        # chains: every non-root node passes its boundary to n >> 1
        for n in nodes:
            sl = oa[a0[node_index[n]]:a1[node_index[n]]]
            cand = [row[sl][er[sl] != n], col[sl][ec[sl] != n]]
            for k in kids.get(n, []):
                cand.append(B[k])
            b = np.unique(np.concatenate(cand)) if cand else np.zeros(0, np.int64)
            b = b[enode[b] != n]
            B[n] = b
        if len(B[1]):
            raise RuntimeError('root front has a boundary')
        # local numbering: position of global dof in its elimination node (E part) ...
        self.nodes = nodes
        nE = {n: len(E[n]) for n in nodes}
        nB = {n: len(B[n]) for n in nodes}
        # group by (depth, nE, nB)
        groups = {}
        for n in nodes:
            if nE[n] == 0 and nB[n] == 0:
                continue
            groups.setdefault((int(np.log2(n)), nE[n], nB[n]), []).append(n)
        self.groups = []
        front_off, rhs_off = {}, {}
        fo, ro = 0, 0
        for key in sorted(groups, key=lambda k: -k[0]):
            d, ne, nb = key
            ns = groups[key]
            n = ne + nb
            self.groups.append(dict(depth=d, nE=ne, nB=nb, nodes=ns, foff=fo, roff=ro, cnt=len(ns)))
            for i, v in enumerate(ns):
                front_off[v] = fo + i * n * n
                rhs_off[v] = ro + i * n
            fo += len(ns) * n * n
            ro += len(ns) * n
        self.front_size, self.rhs_size = fo, ro
        # global dof -> local index in a given node: build per node dictionaries lazily via
        # searchsorted on sorted arrays
        local_sorted = {}
        for n in nodes:
            ids = np.concatenate([E[n], B[n]])
            o = np.argsort(ids, kind='stable')
            local_sorted[n] = (ids[o], o)

        def local(n, dofs):
            ids, o = local_sorted[n]
            pos = np.searchsorted(ids, dofs)
            assert (ids[pos] == dofs).all()
            return o[pos]
        # scatter of A: destination of every stored value
        dest = np.empty(len(row), dtype=np.int64)
        for n in nodes:
            sl = oa[a0[node_index[n]]:a1[node_index[n]]]
            if len(sl) == 0:
                continue
            nn = nE[n] + nB[n]
            dest[sl] = front_off[n] + local(n, row[sl]) * nn + local(n, col[sl])
        self._dest = dest
        self._row, self._col = row, col
        # extend-add maps: child boundary -> parent local index
        for g in self.groups:
            nb, cnt = g['nB'], g['cnt']
            if nb == 0:
                g['pmap'] = None
                continue
            pmap = np.empty((cnt, nb), dtype=np.int64)
            pfo = np.empty(cnt, dtype=np.int64)
            pro = np.empty(cnt, dtype=np.int64)
            pn = np.empty(cnt, dtype=np.int64)
            for i, v in enumerate(g['nodes']):
                par = v >> 1
                while par not in front_off:           # skipped empty ancestors
                    par >>= 1
                pmap[i] = local(par, B[v])
                pfo[i], pro[i], pn[i] = front_off[par], rhs_off[par], nE[par] + nB[par]
            g['pmap'], g['pfo'], g['pro'], g['pn'] = pmap, pfo, pro, pn
        # rhs scatter: dof -> slot in its elimination node
        rslot = np.empty(N, dtype=np.int64)
        for n in nodes:
            if nE[n]:
                rslot[E[n]] = rhs_off[n] + np.arange(nE[n])
        self._rslot = rslot
        self._to_device()

    def _to_device(self):
        dev = self.device
        T = lambda a: torch.as_tensor(a, device=dev)
        self.d_dest = T(self._dest)
        self.d_row = T(self._row)
        self.d_col = T(self._col)
        self.d_rslot = T(self._rslot)
        for g in self.groups:
            if g['pmap'] is not None:
                pm = T(g['pmap'])
                pn = T(g['pn'])[:, None, None]
                g['d_fidx'] = (T(g['pfo'])[:, None, None] + pm[:, :, None] * pn + pm[:, None, :]).reshape(-1)
                g['d_ridx'] = (T(g['pro'])[:, None] + pm).reshape(-1)
        self.fronts = torch.zeros(self.front_size, dtype=torch.float64, device=dev)
        self.rhs = torch.zeros(self.rhs_size, dtype=torch.float64, device=dev)
        pa = self.pa
        self.d_rowptr = T(np.ascontiguousarray(pa.csr.rowptr, dtype=np.int64))
        self.d_colidx = T(np.ascontiguousarray(self._col, dtype=np.int32))
        self._factored = False

    def info(self):
        return dict(fronts=len(self.nodes), groups=len(self.groups),
                    front_doubles=int(self.front_size),
                    largest_front=max(g['nE'] + g['nB'] for g in self.groups),
                    root_size=[g['nE'] for g in self.groups if g['depth'] == 0])

    # ------------------------------------------------------------------ numeric phase
    def factor(self, vals):
        """Numeric factorisation of the matrix with the stored values ``vals`` [nnz]."""
        N = self.N
        a = vals.abs()
        # Ruiz equilibration, a few sweeps (see csrc/band_solver.cu: k_band_ruiz_max): rows
        # and columns balanced together; one row pass + one column pass leaves the minority
        # quasi-Fermi modes of wide-gap devices unresolved
        self.rs = torch.ones(N, dtype=torch.float64, device=self.device)
        self.cs = torch.ones(N, dtype=torch.float64, device=self.device)
        for _ in range(self.n_ruiz):
            s = a * self.rs[self.d_row] * self.cs[self.d_col]
            rmax = torch.zeros(N, dtype=torch.float64, device=self.device)
            rmax.scatter_reduce_(0, self.d_row, s, reduce='amax', include_self=True)
            cmax = torch.zeros(N, dtype=torch.float64, device=self.device)
            cmax.scatter_reduce_(0, self.d_col, s, reduce='amax', include_self=True)
            self.rs = self.rs / torch.sqrt(torch.where(rmax > 0, rmax, torch.ones_like(rmax)))
            self.cs = self.cs / torch.sqrt(torch.where(cmax > 0, cmax, torch.ones_like(cmax)))
        self.fronts.zero_()
        self.fronts.index_add_(0, self.d_dest, vals * self.rs[self.d_row] * self.cs[self.d_col])
        for g in self.groups:
            ne, nb, cnt = g['nE'], g['nB'], g['cnt']
            n = ne + nb
            F = self.fronts[g['foff']:g['foff'] + cnt * n * n].view(cnt, n, n)
            if ne == 0:
                S = F
            else:
                # _ex: no host synchronisation for the error check (a singular front shows up
                # as a non-finite solution, checked once per solve)
                LU, piv, _ = torch.linalg.lu_factor_ex(F[:, :ne, :ne], check_errors=False)
                g['LU'], g['piv'] = LU, piv
                if nb:
                    X = torch.linalg.lu_solve(LU, piv, F[:, :ne, ne:])
                    g['X'] = X
                    g['F21'] = F[:, ne:, :ne]
                    S = F[:, ne:, ne:] - torch.bmm(g['F21'], X)
                else:
                    S = None
            if nb:
                self.fronts.index_add_(0, g['d_fidx'], S.reshape(-1))
        self._factored = True

    def _solve_scaled(self, b):
        """x with (R A C) x = b, fronts factored."""
        rhs = self.rhs
        rhs.zero_()
        rhs.index_add_(0, self.d_rslot, b)
        for g in self.groups:                      # forward, bottom-up
            ne, nb, cnt = g['nE'], g['nB'], g['cnt']
            n = ne + nb
            v = rhs[g['roff']:g['roff'] + cnt * n].view(cnt, n)
            if ne:
                y = torch.linalg.lu_solve(g['LU'], g['piv'], v[:, :ne].unsqueeze(-1)).squeeze(-1)
                v[:, :ne] = y
                if nb:
                    v[:, ne:] -= torch.bmm(g['F21'], y.unsqueeze(-1)).squeeze(-1)
            if nb:
                rhs.index_add_(0, g['d_ridx'], v[:, ne:].reshape(-1).clone())
        for g in reversed(self.groups):            # backward, top-down
            ne, nb, cnt = g['nE'], g['nB'], g['cnt']
            n = ne + nb
            v = rhs[g['roff']:g['roff'] + cnt * n].view(cnt, n)
            if nb:
                xb = rhs[g['d_ridx']].view(cnt, nb)
                v[:, ne:] = xb
                if ne:
                    v[:, :ne] -= torch.bmm(g['X'], xb.unsqueeze(-1)).squeeze(-1)
        return rhs[self.d_rslot]

    def _residual(self, vals, x, b):
        r = torch.empty_like(b)
        if self.device.type == 'cuda' or _lib.is_emulated():
            p = lambda t: C.c_void_p(t.data_ptr())
            st = (C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
                  if self.device.type == 'cuda' else None)
            _lib.check(_lib.get().sb_csr_residual_dd(self.N, p(self.d_rowptr), p(self.d_colidx),
                                                     p(vals), p(x), p(b), p(r), st))
            return r
        # host tensors without the library (unit tests of the algorithm): extended precision
        v = vals.numpy().astype(np.longdouble) * x.numpy().astype(np.longdouble)[self._col]
        acc = np.zeros(self.N, dtype=np.longdouble)
        np.add.at(acc, self._row, v)
        return torch.as_tensor((b.numpy().astype(np.longdouble) - acc).astype(np.float64))

    def solve(self, vals, b, x=None, refactor=True):
        """``x = A^-1 b``.  ``refactor=False`` reuses the numeric factors of the last call
        (modified Newton / refinement with the same matrix)."""
        if refactor or not self._factored:
            self.factor(vals)
        xs = self.cs * self._solve_scaled(self.rs * b)
        for _ in range(self.n_refine):
            r = self._residual(vals, xs, b)
            xs = xs + self.cs * self._solve_scaled(self.rs * r)
        if not bool(torch.isfinite(xs).all()):
            raise RuntimeError('simudo_b200: singular Jacobian in the multifrontal LU')
        if x is None:
            return xs
        x.copy_(xs)
        return x
