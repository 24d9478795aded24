"""Synthetic problems of the BASELINE config shapes (no solve needed).

Builds ``ProblemArrays`` + a physically scaled state directly, so assembly
parity and throughput can be exercised without running the three-stage
initialisation.  Parameter values follow the reference examples:
``example/pn_diode/pn_diode.py:110-164`` + ``physics/material.py:76-89`` (Si
pn junction with SRH) and ``example/fourlayer/fourlayer_example.py:63-110``
(four-layer IB cell with Shockley-Read trapping, radiative processes and
three optical fields).  States are smooth analytic profiles (depletion-like
potential, qfl splitting, small currents) with a y-perturbation, seed-free
(SURVEY.md section 8d, config 4).
"""
import numpy as np

from .fem import model as M
from .fem.problem_arrays import ProblemArrays
from .mesh import Product2DMesh

__all__ = ['pn_problem', 'ib_problem', 'synthetic_state']

_KT300 = 1.3806488e-23 * 300.0 / 1.602176487e-19
_EPS0 = 1.0 / (4 * np.pi * 1e-7 * 299792458.0 ** 2) / 1.602176487e-19 / 1e6  # e/(V um)


def _graded(n, length, junctions):
    """n points on [0, length], clustered around the junction planes."""
    s = np.linspace(0.0, 1.0, n)
    x = s.copy()
    for xj in junctions:
        t = xj / length
        x = x - 0.12 / np.pi / len(junctions) * np.sin(2 * np.pi * (s - t)) * np.exp(-((s - t) / 0.35) ** 2)
    x = np.sort((x - x[0]) / (x[-1] - x[0])) * length
    return x


def strip_rows(n_rows, rank, world):
    """Quad rows [j0, j1) of y-strip ``rank`` of ``world`` (balanced)."""
    j0 = (n_rows * rank) // world
    j1 = (n_rows * (rank + 1)) // world
    return j0, j1


def _strip_Ys(Ys, strip):
    """Local y-lines of a strip with one ghost row of quads on every interior side
    (simudo_b200/parallel.py: StripDecomposition) and the bookkeeping that goes with it."""
    if strip is None:
        return Ys, None
    rank, world = strip
    n_rows = len(Ys) - 1
    if n_rows < 2 * world:
        raise ValueError('every strip needs at least two rows of quads')
    j0, j1 = strip_rows(n_rows, rank, world)
    g_lo, g_hi = int(rank > 0), int(rank < world - 1)
    info = dict(rank=rank, world=world, j0=j0, j1=j1, g_lo=g_lo, g_hi=g_hi, row0=j0 - g_lo,
                n_rows_global=n_rows, Ys_global=np.asarray(Ys, dtype=float))
    return Ys[j0 - g_lo:j1 + g_hi + 1], info


def _boundary_setup(mesh, pa_P, cell_dofs_fn=None):
    bf = mesh.boundary_facets()
    n = mesh.boundary_outward_normals(bf)
    left = bf[n[:, 0] < -0.5]
    right = bf[n[:, 0] > 0.5]
    horiz = bf[np.abs(n[:, 1]) > 0.5]
    return left, right, horiz


def pn_problem(nx=9, ny=2, length=0.5, height=1.0, numbering='b200', with_bcs=True,
               pattern='structural', strip=None):
    """2D Si pn junction: CB + VB Boltzmann bands, SRH (BASELINE config 1/3).
    ``strip=(rank, world)``: the y-strip of that rank with its ghost rows."""
    Xs = _graded(nx, length, [length / 2])
    Ys, sinfo = _strip_Ys(np.linspace(0.0, height, ny), strip)
    mesh = Product2DMesh(Xs, Ys).mesh
    xc = mesh.coordinates[mesh.cells].mean(axis=1)[:, 0]
    cell_tag = (xc > length / 2).astype(np.int32)          # 0 = p side, 1 = n side
    bands = [dict(kind=M.BAND_NONDEGENERATE, sign=-1),
             dict(kind=M.BAND_NONDEGENERATE, sign=+1)]
    procs = [dict(kind=M.PROC_SRH, dst=0, src=1)]
    model = M.make_model(bands, procs, n_fields=0, n_tags=2)
    tp = M.new_tag_params(2)
    kT = _KT300
    Et = 0.5525628437189991
    for t, dop in ((0, -1e18), (1, +1e18)):
        tp[t, M.TP_KT] = kT
        tp[t, M.TP_EPS] = 11.7 * _EPS0
        tp[t, M.TP_RHO_STATIC] = dop * 1e-12
        for b, (N, E, mob) in enumerate(((3.2e19, 1.12, 1400.0), (1.8e19, 0.0, 450.0))):
            o = M.TP_BAND0 + 5 * b
            tp[t, o + M.TPB_IN_SUBDOMAIN] = 1.0
            tp[t, o + M.TPB_N] = N * 1e-12
            tp[t, o + M.TPB_E] = E
            tp[t, o + M.TPB_MOBILITY] = mob * 1e8
        o = M.TP_PROC0
        tp[t, o + M.TPP_P0] = 1e-9        # tau CB
        tp[t, o + M.TPP_P1] = 1e-6        # tau VB
        tp[t, o + M.TPP_P2] = 3.2e19 * 1e-12 * np.exp(-(1.12 - Et) / kT)   # n1
        tp[t, o + M.TPP_P3] = 1.8e19 * 1e-12 * np.exp(-Et / kT)            # p1
    nf = mesh.num_facets
    jmask = np.full(nf, 0b11, dtype=np.uint8)
    pa = ProblemArrays(mesh, cell_tag, model, tp, numbering=numbering,
                       jump_facet_mask=jmask, pattern=pattern)
    pa.strip = sinfo
    pa.extent = (length, height)
    if with_bcs:
        _attach_bcs(pa, mesh, n_dof_bands=2, minority_left=[0], minority_right=[1])
    return pa


def ib_problem(nx=9, ny=2, height=1.0, numbering='b200', with_bcs=True,
               const_mobility=True, pattern='structural', strip=None, ib_bounds_bc=False):
    """Four-layer FSF/p/IB/n intermediate-band cell (BASELINE config 2/4):
    CB, IB, VB; two Shockley-Read traps, radiative CV, radiative CI/IV with
    absorption, three optical fields."""
    th = [0.05, 1.0, 3.0, 1.0]
    edges = np.cumsum([0.0] + th)
    length = edges[-1]
    Xs = _graded(nx, length, list(edges[1:-1]))
    # make sure every layer owns at least one x-interval
    Xs = np.unique(np.concatenate([Xs, edges]))
    Ys, sinfo = _strip_Ys(np.linspace(0.0, height, ny), strip)
    mesh = Product2DMesh(Xs, Ys).mesh
    xc = mesh.coordinates[mesh.cells].mean(axis=1)[:, 0]
    cell_tag = (np.searchsorted(edges, xc) - 1).astype(np.int32)   # 0 fsf,1 p,2 I,3 n
    bands = [dict(kind=M.BAND_NONDEGENERATE, sign=-1),
             dict(kind=M.BAND_INTERMEDIATE, sign=-1, const_mobility=const_mobility),
             dict(kind=M.BAND_NONDEGENERATE, sign=+1)]
    CB, IB, VB = 0, 1, 2
    procs = [dict(kind=M.PROC_RAD_IB, dst=CB, src=IB, trap=IB),     # opt_ci
             dict(kind=M.PROC_RAD_IB, dst=IB, src=VB, trap=IB),     # opt_iv
             dict(kind=M.PROC_RAD_BB, dst=CB, src=VB),              # opt_cv
             dict(kind=M.PROC_SR_TRAP, dst=CB, src=IB, trap=IB),    # nr_top
             dict(kind=M.PROC_SR_TRAP, dst=IB, src=VB, trap=IB)]    # nr_bottom
    model = M.make_model(bands, procs, n_fields=3, n_tags=4)
    tp = M.new_tag_params(4)
    kT = _KT300
    NC = NV = 2e19 * 1e-12
    NI = 2.5e17 * 1e-12
    ECB, EIB = 2.0, 1.2
    dop = [-1e19, -1e17, 0.5 * 2.5e17, 1e17]
    for t in range(4):
        tp[t, M.TP_KT] = kT
        tp[t, M.TP_EPS] = 13.0 * _EPS0
        tp[t, M.TP_RHO_STATIC] = dop[t] * 1e-12
        for b, (N, E, mob, inside) in enumerate((
                (NC, ECB, 500.0, 1.0),
                (NI if t == 2 else 0.0, EIB, 500.0 if t == 2 else 0.0, 1.0 if t == 2 else 0.0),
                (NV, 0.0, 500.0, 1.0))):
            o = M.TP_BAND0 + 5 * b
            tp[t, o + M.TPB_IN_SUBDOMAIN] = inside
            tp[t, o + M.TPB_N] = N
            tp[t, o + M.TPB_E] = E
            tp[t, o + M.TPB_MOBILITY] = mob * 1e8
        if t == 2:
            sig_opt = 3e-14 * 1e8                 # um^2
            strb = 1.0e9                          # representative I [1/(um^2 s)]
            u1_c = NC * np.exp(-(ECB - EIB) / kT)
            u1_v = NV * np.exp(-EIB / kT)
            o = M.TP_PROC0 + 8 * 0
            tp[t, o + M.TPP_P0] = sig_opt * strb / u1_c
            tp[t, o + M.TPP_FIELD0 + 1] = sig_opt
            o = M.TP_PROC0 + 8 * 1
            tp[t, o + M.TPP_P0] = sig_opt * strb * 1e-3 / u1_v
            tp[t, o + M.TPP_FIELD0 + 2] = sig_opt
            cth = 5e-18 * 1e8 * 2e7 * 1e4         # sigma_th v_th [um^3/s]
            tp[t, M.TP_PROC0 + 8 * 3 + M.TPP_P0] = cth
            tp[t, M.TP_PROC0 + 8 * 4 + M.TPP_P0] = cth
        if t in (1, 2, 3):
            alpha = 1e4 * 1e-4                    # 1/um
            o = M.TP_PROC0 + 8 * 2
            tp[t, o + M.TPP_P0] = alpha * 1.0e3
            tp[t, o + M.TPP_FIELD0 + 0] = alpha
    # jump term: CB, VB on all interior facets; IB only inside the IB layer
    mesh.init()
    fc = mesh.facet_cells
    jmask = np.zeros(mesh.num_facets, dtype=np.uint8)
    interior = fc[:, 1] >= 0
    jmask[interior] = 0b101
    both_I = interior & (cell_tag[fc[:, 0]] == 2) & (cell_tag[np.maximum(fc[:, 1], 0)] == 2)
    jmask[both_I] |= 0b010
    pa = ProblemArrays(mesh, cell_tag, model, tp, numbering=numbering,
                       jump_facet_mask=jmask, pattern=pattern)
    pa.strip = sinfo
    pa.extent = (length, height)
    if with_bcs:
        _attach_bcs(pa, mesh, n_dof_bands=3, minority_left=[0], minority_right=[2],
                    extra_all=[1])
        if ib_bounds_bc:
            # IB/j = 0 on the bounds of the IB layer (fourlayer.py: spatial.add_BC('IB/j',
            # F.nonconductive | F.IB_bounds, zj)): Dirichlet dofs on INTERIOR facets
            fm = mesh.facet_midpoints()
            fv = mesh.facet_vertices
            vx = mesh.coordinates[fv][:, :, 0]
            vertical = np.abs(vx[:, 0] - vx[:, 1]) < 1e-14
            on_edge = vertical & ((np.abs(fm[:, 0] - edges[2]) < 1e-12) |
                                  (np.abs(fm[:, 0] - edges[3]) < 1e-12))
            extra = pa.facet_dofs(1 + IB, np.nonzero(on_edge)[0]).ravel()
            pa.set_bc_dofs(np.unique(np.concatenate([pa.bc_dofs, extra])))
    pa.layer_edges = edges
    return pa


def _attach_bcs(pa, mesh, n_dof_bands, minority_left, minority_right, extra_all=()):
    """Typical contact set (``pn_diode.py:127-159``): zero normal flux of E and
    all j on the horizontal (non-conductive) sides, zero minority current at
    each contact; natural phi terms on both contacts."""
    left, right, horiz = _boundary_setup(mesh, pa.P)
    sinfo = getattr(pa, 'strip', None)
    if sinfo is not None:
        # the artificial bottom / top of a strip (outer edge of its ghost rows) is no boundary
        ymid = mesh.facet_midpoints()[horiz][:, 1]
        Yg = sinfo['Ys_global']
        horiz = horiz[(np.abs(ymid - Yg[0]) < 1e-12 * max(1.0, abs(Yg[-1]))) |
                      (np.abs(ymid - Yg[-1]) < 1e-12 * max(1.0, abs(Yg[-1])))]
    bc = [pa.facet_dofs(0, horiz).ravel()]
    for k in range(n_dof_bands):
        bc.append(pa.facet_dofs(1 + k, horiz).ravel())
    for k in minority_left:
        bc.append(pa.facet_dofs(1 + k, left).ravel())
    for k in minority_right:
        bc.append(pa.facet_dofs(1 + k, right).ravel())
    for k in extra_all:
        bc.append(pa.facet_dofs(1 + k, left).ravel())
        bc.append(pa.facet_dofs(1 + k, right).ravel())
    pa.set_bc_dofs(np.concatenate(bc))
    lay = pa.layout()
    cells, rows = [], []
    for facets in (left, right):
        c = mesh.facet_cells[facets, 0]
        l = mesh.facet_local[facets, 0].astype(np.int64)
        loc = lay['facet'][0][l]                       # poisson psi rows
        cells.append(np.repeat(c, 3))
        rows.append(loc.ravel())
    cells = np.concatenate(cells).astype(np.int32)
    rows = np.concatenate(rows).astype(np.int32)
    o = np.argsort(cells, kind='stable')
    pa.natbc_order = np.arange(len(cells))
    pa.natbc_cell = np.ascontiguousarray(cells[o])
    pa.natbc_row = np.ascontiguousarray(rows[o])
    pa.contacts = dict(left=left, right=right, horiz=horiz)


def dof_coordinates(pa):
    """(x, y) of the mesh entity carrying every global dof (cell centroid / facet midpoint)
    and the dof's index inside its entity triple: a numbering-independent label."""
    mesh, lay = pa.mesh, pa.layout()
    xy = np.zeros((pa.N, 2))
    kk = np.zeros(pa.N)
    cc = pa.cell_vertices.mean(axis=1)
    fm = mesh.facet_midpoints()
    cd = pa.cell_dofs
    for p in range(pa.P):
        for i in range(3):
            xy[cd[:, lay['dg'][p][i]]] = cc
            xy[cd[:, lay['interior'][p][i]]] = cc
            kk[cd[:, lay['dg'][p][i]]] = i + 10 * p
            kk[cd[:, lay['interior'][p][i]]] = i + 3 + 10 * p
        for f in range(3):
            for k in range(3):
                d = cd[:, lay['facet'][p][f][k]]
                xy[d] = fm[mesh.cell_facets[:, f]]
                kk[d] = k + 6 + 10 * p
    return xy, kk


def synthetic_state(pa, bias=0.3, y_amp=1e-3):
    """Smooth, physically scaled state: returns dict(u, base_w, thmeq_phi,
    phi_opt, bc_values, natbc_values)."""
    mesh = pa.mesh
    model = pa.model
    nb = model.n_bands
    lay = pa.layout()
    X = mesh.coordinates
    Lx = X[:, 0].max()
    Ly = max(X[:, 1].max(), 1e-30)
    if getattr(pa, 'extent', None) is not None:
        Lx, Ly = pa.extent[0], max(pa.extent[1], 1e-30)     # the global domain, also on a strip
    tp = pa.tag_params
    kT = tp[0, M.TP_KT]

    def phi_fun(x, y):
        return 0.45 * np.tanh((x - 0.5 * Lx) / (0.15 * Lx)) * (1 + y_amp * np.sin(2 * np.pi * y / Ly))

    xv = pa.cell_vertices                       # [Nc,3,2]
    u = np.zeros(pa.N)
    phi_v = phi_fun(xv[:, :, 0], xv[:, :, 1])
    u[pa.cell_dofs[:, lay['dg'][0]]] = phi_v
    # thermal-equilibrium potential as P2 nodal values (vertices + edge midpoints)
    mid = 0.5 * (xv[:, [1, 0, 0], :] + xv[:, [2, 2, 1], :])
    nodes = np.concatenate([xv, mid], axis=1)   # [Nc,6,2]
    thmeq = phi_fun(nodes[:, :, 0], nodes[:, :, 1]) * 0.98
    base_w = np.zeros((nb, pa.n_cells))
    out = dict(thmeq_phi=thmeq)
    if model.bands_have_dofs:
        xc = xv.mean(axis=1)
        for k in range(nb):
            s = model.band_sign[k]
            split = 0.5 * bias * s * (1.0 if model.band_kind[k] == 0 else 0.3)
            wfun = lambda x, y: split * (0.5 + 0.5 * np.tanh((x - 0.5 * Lx) / (0.3 * Lx))) \
                * (1 + y_amp * np.sin(2 * np.pi * y / Ly))
            base_w[k] = wfun(xc[:, 0], xc[:, 1])
            wv = wfun(xv[:, :, 0], xv[:, :, 1])
            u[pa.cell_dofs[:, lay['dg'][1 + k]]] = wv - base_w[k][:, None]
        # fluxes: small values scaled like 1 V/um fields, 1e-9 A/um^2 currents; a function of
        # the dof's mesh entity, not of its number (same state on any numbering / strip)
        dxy, dk = dof_coordinates(pa)
        rng_phase = 37.1 * dxy[:, 0] + 19.7 * dxy[:, 1] + 1.3 * dk
        for p in range(pa.P):
            d = np.unique(pa.cell_dofs[:, lay['bdm'][p]])
            scale = 0.5 if p == 0 else 1e-9
            u[d] = scale * np.sin(rng_phase[d])
    else:
        dxy, dk = dof_coordinates(pa)
        d = np.unique(pa.cell_dofs[:, lay['bdm'][0]])
        u[d] = 0.5 * np.sin(37.1 * dxy[d, 0] + 19.7 * dxy[d, 1] + 1.3 * dk[d])
    out['u'] = u
    out['base_w'] = base_w
    nfld = model.n_fields
    if nfld:
        phi_opt = np.zeros((nfld, pa.n_cells, 6))
        for f in range(nfld):
            flux = 1e9 * (f + 1)                # photons / (um^2 s)
            phi_opt[f] = flux * np.exp(-nodes[:, :, 0] * (0.5 + 0.2 * f)) - 1e6
        out['phi_opt'] = phi_opt
    else:
        out['phi_opt'] = None
    dxy, dk = dof_coordinates(pa)
    bd = pa.bc_dofs
    out['bc_values'] = 1e-12 * np.cos(31.0 * dxy[bd, 0] + 17.0 * dxy[bd, 1] + dk[bd])
    # natural-BC values in the caller's order (before the plan's sort by cell)
    ncell = np.empty(len(pa.natbc_cell), dtype=np.int64)
    nrow = np.empty(len(pa.natbc_cell), dtype=np.int64)
    ncell[pa.natbc_order] = pa.natbc_cell
    nrow[pa.natbc_order] = pa.natbc_row
    cc = pa.cell_vertices.mean(axis=1)
    out['natbc_values'] = 0.3 * np.sin(23.0 * cc[ncell, 0] + 11.0 * cc[ncell, 1] + 1.3 * nrow)
    return out
