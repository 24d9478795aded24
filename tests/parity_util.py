"""Shared comparison helpers for the parity tests."""
import numpy as np


def row_scaled_error(pa, vals, vals_ref):
    """max |a - a_ref| / max_row |a_ref| over all CSR entries."""
    if pa.csr.pattern == 'bsr3':                 # rows are not contiguous in the block layout
        rowid = pa.csr.value_rows()
        rowmax = np.zeros(pa.N)
        np.maximum.at(rowmax, rowid, np.abs(vals_ref))
        return float((np.abs(vals - vals_ref) / np.maximum(rowmax[rowid], 1e-300)).max())
    rp = pa.csr.rowptr
    worst = 0.0
    step = 1 << 22                               # rows per chunk: bounded temporaries at 1e9 nnz
    for r0 in range(0, pa.N, step):
        r1 = min(pa.N, r0 + step)
        a, b = rp[r0], rp[r1]
        ref = np.abs(vals_ref[a:b])
        rowmax = np.maximum.reduceat(ref, rp[r0:r1] - a)
        scale = np.repeat(np.maximum(rowmax, 1e-300), np.diff(rp[r0:r1 + 1]))
        worst = max(worst, float((np.abs(vals[a:b] - vals_ref[a:b]) / scale).max()))
    return worst


def entry_relative_error(vals, vals_ref, floor_rel=1e-10):
    """Worst |a - a_ref| / |a_ref| over the entries with |a_ref| above ``floor_rel`` times
    the largest entry of the matrix (the north star's wording is per assembled entry;
    entries that are round-off of cancelled sums have no meaningful relative error)."""
    top = float(np.abs(vals_ref).max())
    worst = 0.0
    step = 1 << 26
    for a in range(0, len(vals_ref), step):
        ref = np.abs(vals_ref[a:a + step])
        big = ref > floor_rel * top
        if big.any():
            worst = max(worst, float((np.abs(vals[a:a + step] - vals_ref[a:a + step])[big] / ref[big]).max()))
    return worst


def vector_error(b, b_ref):
    return float(np.abs(b - b_ref).max() / max(np.abs(b_ref).max(), 1e-300))


def block_scaled_error(pa, vals, vals_ref):
    """Entry error relative to the max of its (row sub-problem, col sub-problem)
    block class -- stricter than row scaling for small couplings."""
    L = pa.L
    pos = pa.csr.entry_positions()
    worst = 0.0
    for pi in range(pa.P):
        for pj in range(pa.P):
            sl = pos[:, 15 * pi:15 * pi + 15, 15 * pj:15 * pj + 15].ravel()
            ref = vals_ref[sl]
            scale = np.abs(ref).max()
            if scale == 0.0:
                worst = max(worst, float(np.abs(vals[sl]).max()))
                continue
            # entries of one cell block vary over decades: scale per cell
            refc = np.abs(vals_ref[pos[:, 15 * pi:15 * pi + 15, 15 * pj:15 * pj + 15]]).reshape(pos.shape[0], -1).max(axis=1)
            d = np.abs(vals[pos[:, 15 * pi:15 * pi + 15, 15 * pj:15 * pj + 15]]
                       - vals_ref[pos[:, 15 * pi:15 * pi + 15, 15 * pj:15 * pj + 15]]).reshape(pos.shape[0], -1).max(axis=1)
            ok = refc > 0
            if ok.any():
                worst = max(worst, float((d[ok] / refc[ok]).max()))
    return worst


def run_plan(plan, st):
    T = plan.tensor
    vals = plan.new_values()
    b = plan.new_vector()
    nat = st.get('natbc_values')
    plan.assemble(T(st['u']), T(st['base_w']), T(st['thmeq_phi']), T(st['phi_opt']),
                  T(st['bc_values']) if plan.n_bc else None,
                  T(plan.natbc_device_order(nat)) if plan.n_natbc else None, vals, b)
    return vals, b
