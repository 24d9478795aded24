#!/usr/bin/env python
"""Time the two on-device Newton linear solvers on the quasi-1D IB cell (BASELINE configs
2 / 5: 279 x-points x {0, 1}): banded LU (csrc/band_solver.cu) vs multifrontal LU
(fem/multifrontal.py).  GPU only.  Usage: tools/time_linear_solvers.py [nx] [reps]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from simudo_b200 import synthetic as syn
from simudo_b200.fem.assembly import AssemblyPlan
from simudo_b200.fem.linear_solver import BandSolver
from simudo_b200.fem.multifrontal import MultifrontalSolver


def main(nx=279, reps=5):
    dev = 'cuda:0'
    pa = syn.ib_problem(nx=nx, ny=2, pattern='native')
    st = syn.synthetic_state(pa)
    plan = AssemblyPlan(pa, device=dev)
    T = plan.tensor
    vals, b = plan.new_values(), plan.new_vector()
    plan.assemble(T(st['u']), T(st['base_w']), T(st['thmeq_phi']), T(st['phi_opt']),
                  T(st['bc_values']), T(plan.natbc_device_order(st['natbc_values'])), vals, b)
    torch.cuda.synchronize()
    out = dict(cells=pa.n_cells, dofs=pa.N, nnz=int(pa.csr.nnz))
    xs = {}
    which = os.environ.get('SOLVERS', 'band,multifrontal').split(',')
    for name, mk in (('band', lambda: BandSolver(pa, dev)),
                     ('multifrontal', lambda: MultifrontalSolver(pa, dev))):
        if name not in which:
            continue
        t0 = time.perf_counter()
        s = mk()
        out[name + '_setup_s'] = time.perf_counter() - t0
        x = s.solve(vals, b)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(reps):
            x = s.solve(vals, b)
        e1.record()
        torch.cuda.synchronize()
        out[name + '_ms_per_solve'] = e0.elapsed_time(e1) / reps
        out[name + '_wall_ms_per_solve'] = (time.perf_counter() - t0) * 1e3 / reps
        xs[name] = x.cpu().numpy()
        if name == 'band':
            out['band_kl_ku'] = [s.kl, s.ku]
    if len(xs) == 2:
        d = np.abs(xs['band'] - xs['multifrontal']).max() / np.abs(xs['band']).max()
        out['max_rel_difference'] = float(d)
    print(json.dumps(out))


if __name__ == '__main__':
    main(*(int(a) for a in sys.argv[1:]))
