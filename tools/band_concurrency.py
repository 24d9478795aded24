#!/usr/bin/env python
"""Do K banded solves on K streams overlap on the device?  (GPU only.)  One thread enqueues a
solve per stream without host synchronisation; wall time for K solves vs one."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from simudo_b200 import synthetic as syn
from simudo_b200.fem.assembly import AssemblyPlan
from simudo_b200.fem.linear_solver import BandSolver


def main(nx=100, kmax=8):
    dev = 'cuda:0'
    pa = syn.ib_problem(nx=nx, ny=2, pattern='native')
    st = syn.synthetic_state(pa)
    plan = AssemblyPlan(pa, device=dev)
    T = plan.tensor
    vals, b = plan.new_values(), plan.new_vector()
    plan.assemble(T(st['u']), T(st['base_w']), T(st['thmeq_phi']), T(st['phi_opt']),
                  T(st['bc_values']), T(plan.natbc_device_order(st['natbc_values'])), vals, b)
    torch.cuda.synchronize()
    solvers = [BandSolver(pa, dev) for _ in range(kmax)]
    xs = [torch.empty_like(b) for _ in range(kmax)]
    streams = [torch.cuda.Stream() for _ in range(kmax)]
    for s, x, stq in zip(solvers, xs, streams):          # warm-up (and graph capture)
        for _ in range(3):
            s.solve(vals, b, x=x, check=False, stream=stq)
    torch.cuda.synchronize()
    for k in (1, 2, 4, kmax):
        t0 = time.perf_counter()
        for rep in range(3):
            for s, x, stq in list(zip(solvers, xs, streams))[:k]:
                s.solve(vals, b, x=x, check=False, stream=stq)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 3
        print('n=%d  %d solves on %d streams: %.1f ms (%.1f ms per solve if serial)' % (
            pa.N, k, k, dt * 1e3, dt * 1e3 / k))


if __name__ == '__main__':
    main(*(int(a) for a in sys.argv[1:]))
