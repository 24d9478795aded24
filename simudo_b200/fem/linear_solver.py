"""On-device linear solve of the Newton systems.

Replaces ``NewtonSolution.do_linear_solve`` -> ``PETScLUSolver("mumps")``
(``simudo/fem/newton_solver.py:124-181``).  cuDSS is not installed in this
image, so the solver is the hand-written banded LU of
``csrc/band_solver.cu``: dofs are ordered by the x-position of their mesh
entity, which makes the Jacobian of a (thin) product mesh banded.
"""
import ctypes as C

import numpy as np
import torch

from .. import _lib

__all__ = ['x_major_permutation', 'BandSolver', 'make_solver']

# band storage above which a mesh is treated as genuinely two-dimensional
BAND_LIMIT_BYTES = 256 << 20


def make_solver(pa, device, band_limit_bytes=None):
    """The linear solver of a lowered problem: banded LU for product meshes that are thin in
    y (the 1D examples, BASELINE configs 1, 2, 5), multifrontal LU on nested dissection
    otherwise (configs 3, 4).  Both keep their symbolic phase for all later solves."""
    if band_limit_bytes is None:
        band_limit_bytes = BAND_LIMIT_BYTES
    try:
        return BandSolver(pa, device, max_band_bytes=band_limit_bytes)
    except RuntimeError as e:
        if 'status -5' not in str(e):            # SB_ERR_UNSUPPORTED: band too wide
            raise
    from .multifrontal import MultifrontalSolver
    return MultifrontalSolver(pa, device)


def x_major_permutation(pa):
    """New index of each dof when dofs are sorted by the x-coordinate of the
    mesh entity (cell centroid / facet midpoint) carrying them."""
    mesh = pa.mesh
    lay = pa.layout()
    xkey = np.zeros(pa.N)
    cx = pa.cell_vertices[:, :, 0].mean(axis=1)
    fx = mesh.facet_midpoints()[:, 0]
    for p in range(pa.P):
        for i in range(3):
            xkey[pa.cell_dofs[:, lay['dg'][p][i]]] = cx
            xkey[pa.cell_dofs[:, lay['interior'][p][i]]] = cx
        for f in range(3):
            for k in range(3):
                xkey[pa.cell_dofs[:, lay['facet'][p][f][k]]] = fx[mesh.cell_facets[:, f]]
    order = np.argsort(xkey, kind='stable')
    perm = np.empty(pa.N, dtype=np.int32)
    perm[order] = np.arange(pa.N, dtype=np.int32)
    return perm


class BandSolver(object):
    """Banded LU with partial pivoting + equilibration, one CTA per system."""

    def __init__(self, pa, device, batch=1, max_band_bytes=8 << 30, csr=None):
        """``pa``: ProblemArrays of the PDD system, or ``None`` with
        ``csr=(n, rowptr, colidx, perm)`` for any other banded system (the CG2
        optical matrices)."""
        self.lib = _lib.get()
        self.pa = pa
        self.device = torch.device(device)
        self.batch = batch
        if csr is None:
            csr = (pa.N, pa.csr.rowptr, pa.csr.column_indices(), x_major_permutation(pa))
        n, rowptr, colidx, perm = csr
        self.n = int(n)
        rowptr = np.ascontiguousarray(rowptr, dtype=np.int64)
        colidx = np.ascontiguousarray(colidx, dtype=np.int32)
        perm = np.ascontiguousarray(perm, dtype=np.int32)
        h = C.c_void_p()
        _lib.check(self.lib.sb_band_create(
            self.n, rowptr.ctypes.data_as(C.c_void_p), colidx.ctypes.data_as(C.c_void_p),
            perm.ctypes.data_as(C.c_void_p), batch, max_band_bytes, C.byref(h)))
        self._h = h
        kl, ku, n = C.c_int32(), C.c_int32(), C.c_int64()
        self.lib.sb_band_info(h, C.byref(n), C.byref(kl), C.byref(ku))
        self.kl, self.ku = kl.value, ku.value

    def __del__(self):
        h = getattr(self, '_h', None)
        if h:
            try:
                self.lib.sb_band_destroy(h)
            except Exception:
                pass
            self._h = None

    def solve(self, vals, b, x=None, check=True, stream=None):
        """x = A^{-1} b; ``vals`` [nsys, nnz] or [nnz], ``b`` [nsys, n] or [n].
        Raises RuntimeError on a singular matrix (as PETSc/MUMPS would)."""
        nsys = 1 if vals.dim() == 1 else vals.shape[0]
        if x is None:
            x = torch.empty_like(b)
        info = (C.c_int32 * nsys)()
        st = None
        if stream is not None:
            st = C.c_void_p(stream.cuda_stream)
        elif self.device.type == 'cuda':
            st = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        _lib.check(self.lib.sb_band_solve(
            self._h, nsys, C.c_void_p(vals.data_ptr()), vals.shape[-1],
            C.c_void_p(b.data_ptr()), C.c_void_p(x.data_ptr()),
            C.cast(info, C.c_void_p) if check else None, st))
        if check:
            bad = [i for i in range(nsys) if info[i] != 0]
            if bad:
                raise RuntimeError('simudo_b200: singular Jacobian in banded LU '
                                   '(zero pivot at column %d)' % (info[bad[0]] - 1))
        return x

    def solve_factored(self, b, x=None, stream=None):
        """x = A^{-1} b with the factors of the last ``solve`` (a matrix that does not change:
        the CG2 mass matrix of the optical projection)."""
        nsys = 1 if b.dim() == 1 else b.shape[0]
        if x is None:
            x = torch.empty_like(b)
        st = None
        if stream is not None:
            st = C.c_void_p(stream.cuda_stream)
        elif self.device.type == 'cuda':
            st = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        _lib.check(self.lib.sb_band_solve_factored(
            self._h, nsys, C.c_void_p(b.data_ptr()), C.c_void_p(x.data_ptr()), st))
        return x
