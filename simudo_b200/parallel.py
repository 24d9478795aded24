"""Multi-GPU partitioning of the hot path (one process per GPU).

Two natural shardings, neither needs a collective on the
assembly itself:

* **y-strips of one large product mesh** (BASELINE config 4): every rank owns a
  strip; the only data a Newton step must exchange between neighbouring strips
  are the ghost BDM facet dofs on the strip interface (3 P doubles per interface
  facet) and the ghost ``base_w`` of the cells across it (``n_bands`` doubles
  per facet).  ``StripHalo`` gathers / scatters exactly those values and
  exchanges them with grouped point-to-point sends (NCCL on GPUs, gloo in the
  CPU tests).  The lower rank owns an interface: after ``exchange`` the upper
  rank's bottom-boundary dofs equal the lower rank's top-boundary dofs.
* **independent sweep members** (bias points / device parameters, config 5):
  ``shard_round_robin`` assigns members to ranks; results are gathered on the
  host, no data-path communication.
"""
import numpy as np
import torch

__all__ = ['StripHalo', 'StripDecomposition', 'shard_round_robin']


def shard_round_robin(n_items, rank, world):
    """Indices of the sweep members owned by ``rank``."""
    return list(range(rank, n_items, world))


class StripHalo(object):
    def __init__(self, pa, rank, world, device):
        self.rank, self.world = rank, world
        mesh = pa.mesh
        bf = mesh.boundary_facets()
        n = mesh.boundary_outward_normals(bf)
        mid = mesh.facet_midpoints()[bf]
        bottom = bf[n[:, 1] < -0.5]
        top = bf[n[:, 1] > 0.5]
        bottom = bottom[np.argsort(mesh.facet_midpoints()[bottom][:, 0], kind='stable')]
        top = top[np.argsort(mesh.facet_midpoints()[top][:, 0], kind='stable')]
        P = pa.P
        dev = torch.device(device)

        def dofs(facets):
            d = np.stack([pa.facet_dofs(p, facets) for p in range(P)], axis=1)   # [n, P, 3]
            return torch.as_tensor(d.reshape(-1).astype(np.int64), device=dev)

        def cells(facets):
            return torch.as_tensor(mesh.facet_cells[facets, 0].astype(np.int64), device=dev)

        self.dofs_bottom, self.dofs_top = dofs(bottom), dofs(top)
        self.cells_bottom, self.cells_top = cells(bottom), cells(top)
        self.n_bands = pa.model.n_bands if pa.model.bands_have_dofs else 0
        nd = self.dofs_top.numel()
        nw = self.cells_top.numel() * self.n_bands
        self.n_dof_values, self.n_w_values = nd, nw
        mk = lambda: torch.zeros(nd + nw, dtype=torch.float64, device=dev)
        self.send_up, self.send_down = mk(), mk()
        self.recv_from_below, self.recv_from_above = mk(), mk()
        self.ghost_w_below = torch.zeros(self.n_bands, self.cells_bottom.numel(),
                                         dtype=torch.float64, device=dev)
        self.ghost_w_above = torch.zeros_like(self.ghost_w_below)

    @property
    def bytes_per_exchange(self):
        up = self.rank < self.world - 1
        down = self.rank > 0
        return 8 * (self.n_dof_values + self.n_w_values) * (int(up) + int(down))

    def pack(self, u, base_w):
        nd = self.n_dof_values
        self.send_up[:nd] = u[self.dofs_top]
        self.send_down[:nd] = u[self.dofs_bottom]
        if self.n_bands:
            self.send_up[nd:] = base_w[:, self.cells_top].reshape(-1)
            self.send_down[nd:] = base_w[:, self.cells_bottom].reshape(-1)

    def exchange(self, dist):
        """Grouped isend/irecv with the (at most two) neighbours."""
        if self.world == 1:
            return
        ops = []
        if self.rank > 0:
            ops.append(dist.P2POp(dist.isend, self.send_down, self.rank - 1))
            ops.append(dist.P2POp(dist.irecv, self.recv_from_below, self.rank - 1))
        if self.rank < self.world - 1:
            ops.append(dist.P2POp(dist.isend, self.send_up, self.rank + 1))
            ops.append(dist.P2POp(dist.irecv, self.recv_from_above, self.rank + 1))
        for r in dist.batch_isend_irecv(ops):
            r.wait()

    def unpack(self, u):
        """Owner -> ghost: the lower strip owns the interface dofs."""
        nd = self.n_dof_values
        if self.rank > 0:
            u[self.dofs_bottom] = self.recv_from_below[:nd]
            if self.n_bands:
                self.ghost_w_below.copy_(self.recv_from_below[nd:].view(self.n_bands, -1))
        if self.rank < self.world - 1 and self.n_bands:
            self.ghost_w_above.copy_(self.recv_from_above[nd:].view(self.n_bands, -1))

    def step(self, dist, u, base_w):
        self.pack(u, base_w)
        self.exchange(dist)
        self.unpack(u)


class StripDecomposition(object):
    """Domain decomposition of ONE product mesh into y-strips (BASELINE configs[3]: the 1M-cell
    mesh on 1/2/4/8 GPUs; SURVEY.md section 8e row 3 -- the reference has none: its MUMPS
    runs sequentially, ``newton_solver.py:144-146``).

    Rank r assembles the cells of its rows [j0, j1) plus one ghost row of quads on every
    interior side (``synthetic.*_problem(strip=(r, world))`` builds that local problem).
    *Owner computes*: a cell belongs to the rank of its row, a facet to the lowest rank
    among its cells' ranks; every row of the Jacobian / residual that belongs to an owned
    dof is complete after the local assembly (the ghost cells are recomputed, no reduction),
    rows of ghost dofs are discarded.  Per Newton step the only communication is the ghost
    update of ``u`` and ``base_w``: every rank sends the values it owns of the dofs in its
    neighbours' ghost rows (grouped point-to-point sends: NCCL on the GPUs, gloo in the CPU
    tests).  Both sides order the exchanged dofs by a global entity key computed from the
    mesh geometry, so no set-up communication is needed.

    ``gkey[N]``: global key of every local dof (identical on every rank holding the dof);
    ``owned[N]``: this rank owns the dof (its row of the system is complete here).
    """

    def __init__(self, pa, device='cpu'):
        s = pa.strip
        if s is None:
            raise ValueError('pa was not built as a strip (synthetic.*_problem(strip=...))')
        self.pa, self.info = pa, s
        self.rank, self.world = s['rank'], s['world']
        mesh = pa.mesh
        nXg = len(np.unique(mesh.coordinates[:, 0]))
        nYl = len(np.unique(mesh.coordinates[:, 1]))
        nYg = s['n_rows_global'] + 1
        row0 = s['row0']
        # local vertex v = ix * nYl + iy  (mesh/product2d.py)
        def vij(v):
            return v // nYl, v % nYl + row0
        # cells: row and global index 2 (ix (nYg - 1) + row) + t, t = position inside the quad
        cv = mesh.cells.astype(np.int64)
        cix, ciy = vij(cv.min(axis=1))
        t = (np.arange(mesh.num_cells) % 2).astype(np.int64)
        cell_key = 2 * (cix * (nYg - 1) + ciy) + t
        cell_row = ciy
        # facets: horizontal (on line j), vertical / diagonal (inside row j)
        fv = mesh.facet_vertices.astype(np.int64)
        ax, ay = vij(fv[:, 0])
        bx, by = vij(fv[:, 1])
        horiz = ay == by
        vert = ax == bx
        ftype = np.where(horiz, 0, np.where(vert, 1, 2)).astype(np.int64)
        fix, fj = np.minimum(ax, bx), np.minimum(ay, by)
        facet_key = (ftype * nXg + fix) * nYg + fj
        # owner ranks: row j -> rank; horizontal line j -> rank of row j - 1 (row 0 for j = 0)
        rows = np.arange(nYg - 1)
        rank_of_row = np.searchsorted(
            np.array([(s['n_rows_global'] * (r + 1)) // self.world for r in range(self.world)]),
            rows, side='right')
        frow = np.where(horiz, np.maximum(fj - 1, 0), fj)
        cell_owner = rank_of_row[cell_row]
        facet_owner = rank_of_row[frow]
        # per dof
        N, P = pa.N, pa.P
        lay = pa.layout()
        cd = pa.cell_dofs.astype(np.int64)
        nkey = 2 * (nXg - 1) * (nYg - 1) + 3 * nXg * nYg        # cells, then facets
        gkey = np.full(N, -1, dtype=np.int64)
        owner = np.full(N, -1, dtype=np.int64)
        for p in range(P):
            for i in range(3):
                for slot, loc in ((i, lay['dg'][p][i]), (3 + i, lay['interior'][p][i])):
                    d = cd[:, loc]
                    gkey[d] = (cell_key * P + p) * 6 + slot
                    owner[d] = cell_owner
            for f in range(3):
                fa = mesh.cell_facets[:, f].astype(np.int64)
                for k in range(3):
                    d = cd[:, lay['facet'][p][f][k]]
                    gkey[d] = nkey * P * 6 + (facet_key[fa] * P + p) * 3 + k
                    owner[d] = facet_owner[fa]
        assert (gkey >= 0).all()
        self.gkey, self.owner = gkey, owner
        self.owned = owner == self.rank
        self.cell_key, self.cell_owner = cell_key, cell_owner
        self.owned_cells = cell_owner == self.rank
        # exchange lists with the neighbours, ordered by global key on both sides.
        # A dof owned by rank q exists on rank q +- 1 iff its entity lies in that rank's
        # ghost row: send what I own and the neighbour holds, receive what the neighbour owns.
        dev = torch.device(device)
        self.send, self.recv, self.send_cells, self.recv_cells = {}, {}, {}, {}
        n_rows = s['n_rows_global']
        for q in (self.rank - 1, self.rank + 1):
            if q < 0 or q >= self.world:
                continue
            qj0, qj1 = (n_rows * q) // self.world, (n_rows * (q + 1)) // self.world
            qlo, qhi = qj0 - int(q > 0), qj1 + int(q < self.world - 1)      # rows held by q
            in_q_cell = (cell_row >= qlo) & (cell_row < qhi)
            in_q_facet = np.where(horiz, (fj >= qlo) & (fj <= qhi), (fj >= qlo) & (fj < qhi))
            dof_in_q = np.zeros(N, dtype=bool)
            for p in range(P):
                for i in range(3):
                    dof_in_q[cd[in_q_cell][:, lay['dg'][p][i]]] = True
                    dof_in_q[cd[in_q_cell][:, lay['interior'][p][i]]] = True
            fmask = in_q_facet[mesh.cell_facets]
            for p in range(P):
                for f in range(3):
                    for k in range(3):
                        d = cd[:, lay['facet'][p][f][k]]
                        dof_in_q[d[fmask[:, f]]] = True
            snd = np.nonzero(self.owned & dof_in_q)[0]
            rcv = np.nonzero(owner == q)[0]
            snd = snd[np.argsort(gkey[snd], kind='stable')]
            rcv = rcv[np.argsort(gkey[rcv], kind='stable')]
            self.send[q] = torch.as_tensor(snd, device=dev)
            self.recv[q] = torch.as_tensor(rcv, device=dev)
            sc = np.nonzero(self.owned_cells & in_q_cell)[0]
            rc = np.nonzero(cell_owner == q)[0]
            self.send_cells[q] = torch.as_tensor(sc[np.argsort(cell_key[sc], kind='stable')], device=dev)
            self.recv_cells[q] = torch.as_tensor(rc[np.argsort(cell_key[rc], kind='stable')], device=dev)
        nb = pa.model.n_bands if pa.model.bands_have_dofs else 0
        self.n_bands = nb
        mk = lambda n: torch.zeros(n, dtype=torch.float64, device=dev)
        self._sbuf = {q: mk(len(self.send[q]) + nb * len(self.send_cells[q])) for q in self.send}
        self._rbuf = {q: mk(len(self.recv[q]) + nb * len(self.recv_cells[q])) for q in self.recv}

    @property
    def bytes_per_exchange(self):
        return 8 * sum(b.numel() for b in self._sbuf.values())

    def exchange(self, dist, u, base_w=None):
        """Ghost update of ``u`` (and ``base_w``) from their owners, in place."""
        ops = []
        for q, idx in self.send.items():
            buf = self._sbuf[q]
            buf[:len(idx)] = u[idx]
            if self.n_bands:
                buf[len(idx):] = base_w[:, self.send_cells[q]].reshape(-1)
            ops.append(dist.P2POp(dist.isend, buf, q))
            ops.append(dist.P2POp(dist.irecv, self._rbuf[q], q))
        if ops:
            for r in dist.batch_isend_irecv(ops):
                r.wait()
        for q, idx in self.recv.items():
            buf = self._rbuf[q]
            u[idx] = buf[:len(idx)]
            if self.n_bands:
                base_w[:, self.recv_cells[q]] = buf[len(idx):].view(self.n_bands, -1)
