"""Flood-fill partitioning of the DG0 quasi-Fermi offsets (``base_w``).

Mirror of ``PartitionOffsets`` (``simudo/fem/offset_partitioning.py:14-169``), written from
the algorithm its docstring states (steps 0-9, ``:40-69``).  With the default thresholds
``(0, 0)`` rebasing is the per-cell average (+ the boundary-condition override) and runs on
the device (``sb_rebase_w``); with non-zero ``mixedqfl_debug_fill_thresholds`` neighbouring
cells whose values are "close" receive one common offset, so that ``base_w`` is piecewise
constant over whole regions instead of per cell.  The fill is inherently sequential (a
breadth-first walk whose result depends on the visiting order) and a debug option in the
reference, so it stays on the host: O(cells), once per Newton iteration when enabled.
"""
import collections

import numpy as np

__all__ = ['partition_offsets']


def partition_offsets(values, adjacency, thresholds=(0.0, 0.0), bc_cells=((), ()),
                      fill_with_zero_except_bc=False):
    """``values`` [Nc]: cell values (base_w + cell average of delta_w); ``adjacency`` [Nc, 3]:
    neighbour across each facet or -1; ``thresholds`` = (absolute, relative); ``bc_cells`` =
    (cell indices, values) of the cells next to boundary conditions.  Returns the new offsets.
    """
    vec = np.asarray(values, dtype=np.float64)
    n = len(vec)
    abs_t, rel_t = float(thresholds[0]), float(thresholds[1])
    bc_idx = np.asarray(bc_cells[0], dtype=np.int64)
    bc_val = np.asarray(bc_cells[1], dtype=np.float64)
    if abs_t == 0.0 and rel_t == 0.0:
        out = vec.copy()
        out[bc_idx] = bc_val
        return out
    marker = np.full(n, np.nan)
    queue = collections.deque()
    for c, v in zip(bc_idx, bc_val):                       # step 1
        marker[c] = v
        queue.appendleft(int(c))
    if fill_with_zero_except_bc:
        marker[np.isnan(marker)] = 0.0
        return marker
    seen = np.zeros(n, dtype=bool)
    next_unseen = 0
    adjacency = np.asarray(adjacency)
    while True:                                            # step 2
        if queue:
            cell = queue.pop()                             # step 4
        else:                                              # step 3
            while next_unseen < n and seen[next_unseen]:
                next_unseen += 1
            if next_unseen >= n:
                break
            cell = next_unseen
        seen[cell] = True                                  # step 5
        this = marker[cell]                                # step 6
        unmarked = np.isnan(this)
        if unmarked:
            this = vec[cell]
        marked_any = False
        for adj in adjacency[cell]:
            if adj < 0 or seen[adj]:
                continue
            a = vec[adj]
            if abs(this - a) <= abs_t + rel_t * max(abs(this), abs(a)):    # step 7
                marker[adj] = this
                seen[adj] = True                           # step 8
                queue.appendleft(int(adj))
                marked_any = True
        if marked_any and unmarked:                        # step 9
            marker[cell] = this
    return np.where(np.isnan(marker), vec, marker)
