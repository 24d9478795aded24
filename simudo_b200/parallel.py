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

__all__ = ['StripHalo', 'shard_round_robin']


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
