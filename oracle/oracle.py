"""ORACLE driver (test infrastructure, not product code).  PARITY UNPINNED.

ctypes front-end of ``oracle/pdd_oracle.c`` plus numpy restatements of the
small non-assembly pieces of the path (rebasing, Newton metrics, terminal
current).  Input is a ``ProblemArrays`` (plain lowered data) and a state;
the element tables come from ``oracle/fiat_like.py`` -- nothing here touches
the CUDA extension or ``simudo_b200.fem.reference_element``.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from . import fiat_like as fl

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    so = os.path.join(_HERE, 'libpdd_oracle.so')
    src = os.path.join(_HERE, 'pdd_oracle.c')
    hdr = os.path.join(_HERE, '..', 'include', 'simudo_b200.h')
    stale = (not os.path.exists(so)
             or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)))
    if force or stale:
        subprocess.check_call(['make', '-C', _HERE, '-s', '-B', 'libpdd_oracle.so'])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.pdd_oracle_assemble.restype = C.c_int
        _LIB.pdd_oracle_cell_tensors.restype = C.c_int
        _LIB.pdd_oracle_max_threads.restype = C.c_int
    return _LIB


def _sb_model_type():
    # the oracle shares only the *input struct layout* with the product
    from simudo_b200.fem.model import SbModel
    return SbModel


class _Rule(C.Structure):
    _fields_ = [('n', C.c_int32), ('w', C.c_void_p), ('p1', C.c_void_p),
                ('p2', C.c_void_p), ('bdm', C.c_void_p), ('bdmdiv', C.c_void_p)]


def _problem_struct():
    SbModel = _sb_model_type()

    class _Problem(C.Structure):
        _fields_ = [
            ('model', SbModel), ('n_cells', C.c_int64), ('n_dofs', C.c_int64),
            ('P', C.c_int32),
            ('cell_vertices', C.c_void_p), ('cell_tag', C.c_void_p),
            ('cell_dofs', C.c_void_p), ('cell_neighbors', C.c_void_p),
            ('cell_jump_mask', C.c_void_p), ('tag_params', C.c_void_p),
            ('rowptr', C.c_void_p), ('colidx', C.c_void_p),
            ('n_bc', C.c_int64), ('bc_dofs', C.c_void_p), ('bc_values', C.c_void_p),
            ('n_natbc', C.c_int64), ('natbc_cell', C.c_void_p),
            ('natbc_row', C.c_void_p), ('natbc_values', C.c_void_p),
            ('rule_default', _Rule), ('rule_rho', _Rule), ('rule_super', _Rule),
            ('rule_g', _Rule),
            ('n_facet_q', C.c_int32), ('facet_w', C.c_void_p),
            ('facet_trace', C.c_void_p),
            ('u', C.c_void_p), ('base_w', C.c_void_p), ('thmeq_phi', C.c_void_p),
            ('phi_opt', C.c_void_p), ('n_threads', C.c_int32), ('cols_sorted', C.c_int32)]
    return _Problem


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None and a.size else None


class Oracle(object):
    def __init__(self, pa, n_threads=1, colidx=None):
        """``colidx``: the column of every stored value if the caller already has it (large
        meshes: materialising it on the host takes minutes); default ``pa.csr``'s."""
        self.pa = pa
        self.n_threads = n_threads
        self._keep = []
        # The C restatement inserts by searching the column in the row (like PETSc's
        # MatSetValues).  CSR layouts are assembled in place (linear search when the columns
        # of a row are not ascending: 'native'); the block layout 'bsr3' is assembled as a
        # sorted CSR and permuted back into the caller's layout.
        if pa.csr.pattern == 'bsr3':
            self._rowptr_csr, self._colidx_sorted, self._perm = pa.csr.sorted_csr()
            self._sorted, self._inplace_sorted = False, True
            self.colidx = None
        else:
            self.colidx = pa.csr.column_indices() if colidx is None else np.ascontiguousarray(colidx, dtype=np.int32)
            self._rowptr_csr, self._colidx_sorted, self._perm = pa.csr.rowptr, self.colidx, None
            self._sorted = True
            self._inplace_sorted = pa.csr.pattern != 'native'
        self._rules = {}
        m = pa.model
        self.rule_default = self._make_rule(*fl.default_triangle_scheme(4))
        self.rule_rho = self._make_rule(*fl.collapsed_triangle_rule(m.quad_m_rho))
        self.rule_super = self._make_rule(*fl.collapsed_triangle_rule(m.quad_m_super))
        self.rule_g = self._make_rule(*fl.collapsed_triangle_rule(m.quad_m_g))
        _, fw, ftr, _, _ = fl.facet_tabulation(2)
        self.facet_w = np.ascontiguousarray(fw)
        self.facet_trace = np.ascontiguousarray(ftr)

    def _make_rule(self, pts, w):
        val, div = fl.tabulate_bdm2(pts)
        arrs = dict(w=np.ascontiguousarray(w), p1=np.ascontiguousarray(fl.tabulate_p1(pts)),
                    p2=np.ascontiguousarray(fl.tabulate_p2(pts)),
                    bdm=np.ascontiguousarray(val), bdmdiv=np.ascontiguousarray(div))
        self._keep.append(arrs)
        r = _Rule()
        r.n = len(w)
        for k, v in arrs.items():
            setattr(r, k, _ptr(v))
        return r

    def _struct(self, u, base_w, thmeq_phi, phi_opt, bc_values, natbc_values):
        pa = self.pa
        Problem = _problem_struct()
        q = Problem()
        C.memmove(C.byref(q.model), C.byref(pa.model), C.sizeof(pa.model))
        q.n_cells, q.n_dofs, q.P = pa.n_cells, pa.N, pa.P
        keep = dict(
            cell_vertices=pa.cell_vertices, cell_tag=pa.cell_tag,
            cell_dofs=pa.cell_dofs, cell_neighbors=pa.cell_neighbors,
            cell_jump_mask=pa.cell_jump_mask, tag_params=pa.tag_params,
            rowptr=self._rowptr_csr, colidx=self._colidx_sorted, bc_dofs=pa.bc_dofs,
            natbc_cell=pa.natbc_cell, natbc_row=pa.natbc_row,
            facet_w=self.facet_w, facet_trace=self.facet_trace)
        nb, nc, nf = pa.model.n_bands, pa.n_cells, pa.model.n_fields
        keep['u'] = np.ascontiguousarray(u, dtype=np.float64)
        assert keep['u'].shape == (pa.N,)
        keep['base_w'] = (None if base_w is None else
                          np.ascontiguousarray(base_w, dtype=np.float64).reshape(nb, nc))
        keep['thmeq_phi'] = (None if thmeq_phi is None else
                             np.ascontiguousarray(thmeq_phi, dtype=np.float64).reshape(nc, 6))
        keep['phi_opt'] = (None if phi_opt is None or nf == 0 else
                           np.ascontiguousarray(phi_opt, dtype=np.float64).reshape(nf, nc, 6))
        if nf > 0:
            assert keep['phi_opt'] is not None
        bcv = np.zeros(len(pa.bc_dofs)) if bc_values is None else \
            np.ascontiguousarray(bc_values, dtype=np.float64)
        assert bcv.shape == pa.bc_dofs.shape
        keep['bc_values'] = bcv
        nv = np.zeros(len(pa.natbc_cell)) if natbc_values is None else \
            np.ascontiguousarray(np.asarray(natbc_values, dtype=np.float64)[pa.natbc_order])
        keep['natbc_values'] = nv
        for k, v in keep.items():
            setattr(q, k, _ptr(v))
        q.n_bc = len(pa.bc_dofs)
        q.n_natbc = len(pa.natbc_cell)
        q.rule_default, q.rule_rho = self.rule_default, self.rule_rho
        q.rule_super, q.rule_g = self.rule_super, self.rule_g
        q.n_facet_q = len(self.facet_w)
        q.n_threads = self.n_threads
        q.cols_sorted = 1 if self._inplace_sorted else 0
        return q, keep

    def assemble(self, u, base_w=None, thmeq_phi=None, phi_opt=None,
                 bc_values=None, natbc_values=None, want_A=True, want_b=True):
        """Return ``(vals[nnz], b[N])`` in the dolfin-pattern CSR of ``pa.csr``."""
        q, keep = self._struct(u, base_w, thmeq_phi, phi_opt, bc_values, natbc_values)
        vals = np.zeros(self.pa.csr.nnz) if want_A else None
        b = np.zeros(self.pa.N) if want_b else None
        rc = lib().pdd_oracle_assemble(C.byref(q), _ptr(vals), _ptr(b),
                                       C.c_int((1 if want_A else 0) | (2 if want_b else 0)))
        if rc != 0:
            raise RuntimeError('oracle assembly failed')
        if want_A and not self._sorted:
            out = np.empty_like(vals)
            out[self._perm] = vals
            vals = out
        return vals, b

    def cell_tensors(self, cells, u, base_w=None, thmeq_phi=None, phi_opt=None):
        q, keep = self._struct(u, base_w, thmeq_phi, phi_opt, None, None)
        cells = np.ascontiguousarray(cells, dtype=np.int64)
        L = self.pa.L
        A = np.zeros((len(cells), L, L))
        b = np.zeros((len(cells), L))
        lib().pdd_oracle_cell_tensors(C.byref(q), C.c_int64(len(cells)), _ptr(cells),
                                      _ptr(A), _ptr(b))
        return A, b

    def to_scipy(self, vals):
        import scipy.sparse as sp
        # copies: scipy sorts the indices of a CSR matrix in place, which would silently
        # permute the caller's value array of a layout with unsorted columns ('native')
        v = np.array(vals, copy=True) if self._perm is None else np.asarray(vals)[self._perm]
        return sp.csr_matrix((v, self._colidx_sorted.copy(), self._rowptr_csr.copy()),
                             shape=(self.pa.N, self.pa.N))


# ---------------------------------------------------------------------------
# small numpy restatements (non-assembly rows of SURVEY 8a)
# ---------------------------------------------------------------------------
def natural_bc_integrals(pa, cells, local_facets, c0, p2):
    """``oint g psi_k . n ds`` for the three BDM2 dofs k of exterior facets (cell, local
    facet), ``g = c0 + P2 function`` given by nodal values on the cell -- the `ds` term of
    ``poisson_drift_diffusion1.py1:47`` built by ``fem/spatial.py:237-273``.  Restated in
    physical space, independent of the product's reference-trace tables: Piola-mapped BDM2
    values (``fiat_like.tabulate_bdm2``) at the Gauss-Legendre points of the edge, the
    geometric outward normal, ds = |edge| dt.  Returns [n, 3]."""
    cells = np.asarray(cells, dtype=np.int64)
    lf = np.asarray(local_facets, dtype=np.int64)
    t, w = np.polynomial.legendre.leggauss(3)
    t, w = 0.5 * (t + 1.0), 0.5 * w
    ref_v = np.array([[0.0, 0.0], [1.0, 0.0], [0.0, 1.0]])
    edge = [(1, 2), (0, 2), (0, 1)]                    # local facet i is opposite vertex i
    out = np.zeros((len(cells), 3))
    xv = pa.cell_vertices
    for i, (c, l) in enumerate(zip(cells, lf)):
        a, b = edge[l]
        X = ref_v[a][None, :] + t[:, None] * (ref_v[b] - ref_v[a])[None, :]      # [q, 2]
        val, _ = fl.tabulate_bdm2(X)                    # [q, 12, 2] reference values
        val = np.asarray(val)
        if val.shape[0] == 12:                          # accept [12, q, 2] too
            val = np.transpose(val, (1, 0, 2))
        x = xv[c]
        J = np.stack([x[1] - x[0], x[2] - x[0]], axis=1)
        det = np.linalg.det(J)
        phys = np.einsum('ab,qkb->qka', J, val) / det   # contravariant Piola
        e = x[b] - x[a]
        n = np.array([e[1], -e[0]])
        mid, opp = 0.5 * (x[a] + x[b]), x[l]
        if np.dot(n, mid - opp) < 0:
            n = -n                                      # outward, |n| = |edge|
        g = c0[i] + np.asarray(fl.tabulate_p2(X)) @ p2[i]                         # [q]
        for k in range(3):
            out[i, k] = np.sum(w * g * (phys[:, 3 * l + k, :] @ n))
    return out


def rebase_w(pa, u, base_w, band, bc_cells=(), bc_avg=()):
    """``mixedqfl_do_rebase_w`` with thresholds (0, 0)
    (poisson_drift_diffusion1.py1:339-383, offset_partitioning.py:95-107):
    DG0 L2-projection of a P1 function = mean of its three vertex values."""
    lay = pa.layout()
    dofs = pa.cell_dofs[:, lay['dg'][1 + band]]
    dw = u[dofs]
    old = base_w[band].copy()
    new = old + dw.mean(axis=1)
    if len(bc_cells):
        new[np.asarray(bc_cells)] = bc_avg
    u = u.copy()
    u[dofs] = dw + (old - new)[:, None]
    base_w = base_w.copy()
    base_w[band] = new
    return u, base_w


def newton_metrics(u, du, b, tol, rel_tol, abs_tol):
    """newton_solver.py:79-122; ``u`` is the solution *after* the update, which is
    what the reference's lazily evaluated norms see (:306-317, :346-352)."""
    with np.errstate(divide='ignore', invalid='ignore'):
        r = np.abs(du / u)
        ok = ~np.isnan(r)
        rel = np.linalg.norm(r[ok], ord=2) if ok.any() else np.nan
        comb = np.abs(du) / (np.abs(u) * rel_tol + abs_tol * tol)
    return dict(b_norm=np.linalg.norm(b, ord=np.inf),
                du_norm=np.linalg.norm(du, ord=np.inf),
                rel_du_norm=rel,
                combined_du_norm=np.linalg.norm(comb, ord=np.inf),
                has_nans=bool(np.isnan(u).any() or np.isnan(b).any()))


def surface_flux(pa, u, p, facets):
    """integral of j.n over exterior facets and their total length
    (output_writer.py:174-231): BDM2 normal trace is quadratic along the
    edge with nodes at 1/4, 1/2, 3/4 -> weights (2/3, -1/3, 2/3)."""
    mesh = pa.mesh
    d = pa.facet_dofs(p, facets)
    c = mesh.facet_cells[facets, 0]
    l = mesh.facet_local[facets, 0].astype(np.int64)
    _, det = mesh.cell_jacobians()
    ref_out = np.array([1.0, -1.0, 1.0])
    sgn = ref_out[l] * np.sign(det[c])
    flux = (u[d] @ np.array([2 / 3, -1 / 3, 2 / 3])) * sgn
    return flux.sum(), mesh.facet_lengths()[facets].sum()


def eval_points(pa, u, p, points, vector=False):
    """Point evaluation of sub-function ``p`` (Function_eval_many,
    simudo/fem/extra_bindings.py:27-75): brute-force search for the containing cell with the
    lowest index, then DG1 barycentric interpolation, or the BDM2 basis tabulated from its
    symbolic definition (fiat_like.tabulate_bdm2) pushed forward by the contravariant Piola
    map.  Returns (values [n] or [n, 2], cells [n]); points outside the mesh: NaN, -1."""
    from . import fiat_like
    pts = np.asarray(points, dtype=np.float64).reshape(-1, 2)
    xv = np.asarray(pa.cell_vertices, dtype=np.float64).reshape(-1, 3, 2)
    v0 = xv[:, 0, :]
    J = np.stack([xv[:, 1, :] - v0, xv[:, 2, :] - v0], axis=2)          # [c, xy, XY]
    det = J[:, 0, 0] * J[:, 1, 1] - J[:, 0, 1] * J[:, 1, 0]
    vals = np.full((len(pts), 2) if vector else (len(pts),), np.nan)
    cells = np.full(len(pts), -1, dtype=np.int64)
    cd = np.asarray(pa.cell_dofs, dtype=np.int64)
    tol = 1e-12
    for i, (x, y) in enumerate(pts):
        dx, dy = x - v0[:, 0], y - v0[:, 1]
        X = (J[:, 1, 1] * dx - J[:, 0, 1] * dy) / det
        Y = (J[:, 0, 0] * dy - J[:, 1, 0] * dx) / det
        hit = np.nonzero((X >= -tol) & (Y >= -tol) & (X + Y <= 1 + tol))[0]
        if not len(hit):
            continue
        c = hit[0]
        cells[i] = c
        ref = np.array([[X[c], Y[c]]])
        if not vector:
            vals[i] = fiat_like.tabulate_p1(ref)[0] @ u[cd[c, 15 * p:15 * p + 3]]
        else:
            bv, _ = fiat_like.tabulate_bdm2(ref)                          # [1, 12, 2]
            hat = u[cd[c, 15 * p + 3:15 * p + 15]] @ bv[0]
            vals[i] = J[c] @ hat / det[c]
    return vals, cells
