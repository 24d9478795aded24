#!/usr/bin/env python
"""Does an sb_assemble call block on the host while another thread's band solve runs on
another stream?  (GPU only.)"""
import os
import sys
import threading
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from simudo_b200 import synthetic as syn
from simudo_b200.fem.assembly import AssemblyPlan
from simudo_b200.fem.linear_solver import BandSolver


def setup(nx, dev):
    pa = syn.ib_problem(nx=nx, ny=2, pattern='native')
    st = syn.synthetic_state(pa)
    plan = AssemblyPlan(pa, device=dev)
    T = plan.tensor
    args = (T(st['u']), T(st['base_w']), T(st['thmeq_phi']), T(st['phi_opt']), T(st['bc_values']),
            T(plan.natbc_device_order(st['natbc_values'])), plan.new_values(), plan.new_vector())
    plan.assemble(*args)
    torch.cuda.synchronize()
    return pa, plan, args


def main():
    dev = 'cuda:0'
    pa, plan, args = setup(100, dev)
    pa2, plan2, args2 = setup(100, dev)
    bs = BandSolver(pa2, dev)
    x = torch.empty_like(args2[7])
    stop = threading.Event()

    def solver_thread():
        torch.cuda.set_device(0)
        with torch.cuda.stream(torch.cuda.Stream()):
            while not stop.is_set():
                bs.solve(args2[6], args2[7], x=x)

    def time_assembles(label):
        with torch.cuda.stream(torch.cuda.Stream()):
            ts = []
            for _ in range(60):
                t = time.perf_counter()
                plan.assemble(*args)
                ts.append(time.perf_counter() - t)
                torch.cuda.current_stream().synchronize()
            print('%-28s sb_assemble host time: median %.3f ms, max %.3f ms' % (
                label, 1e3 * np.median(ts), 1e3 * np.max(ts)))

    time_assembles('alone')
    th = threading.Thread(target=solver_thread)
    th.start()
    time.sleep(0.5)
    time_assembles('while a band solve runs')
    stop.set()
    th.join()


if __name__ == '__main__':
    main()
