#!/usr/bin/env python
"""Where does a chain's wall time go when K chains are in flight in one process?  Wraps the
device entry points with per-thread timers.  Usage: tools/prof_ensemble.py K  (GPU only)"""
import collections
import os
import sys
import threading
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from simudo_b200.example import ibsc_ensemble
from simudo_b200.fem import assembly, backend, linear_solver, optics_solver

acc = collections.defaultdict(lambda: collections.defaultdict(float))
cnt = collections.defaultdict(lambda: collections.defaultdict(int))


def wrap(cls, name, label):
    f = getattr(cls, name)

    def g(*a, **k):
        t = time.perf_counter()
        try:
            return f(*a, **k)
        finally:
            tid = threading.get_ident()
            acc[tid][label] += time.perf_counter() - t
            cnt[tid][label] += 1
    setattr(cls, name, g)


wrap(linear_solver.BandSolver, 'solve', 'band.solve')
wrap(backend.CudaBackend, 'assemble', 'assemble')
wrap(backend.CudaBackend, 'newton_update', 'newton_update(+sync)')
wrap(backend.CudaBackend, 'rebase', 'rebase')
wrap(backend.CudaBackend, 'optical_update', 'optical_update(total)')
wrap(backend.CudaBackend, 'surface_flux', 'surface_flux')
wrap(assembly.AssemblyPlan, '__init__', 'plan create')

from simudo_b200 import _lib as _lib0
_cd = _lib0.get()
for _name in ('sb_assemble', 'sb_band_solve', 'sb_newton_update_damped', 'sb_newton_update'):
    _orig = getattr(_cd, _name)

    def _timed(*a, _orig=_orig, _name=_name):
        t = time.perf_counter()
        r = _orig(*a)
        tid = threading.get_ident()
        acc[tid]['C:' + _name] += time.perf_counter() - t
        cnt[tid]['C:' + _name] += 1
        return r
    setattr(_cd, _name, _timed)

K = int(sys.argv[1]) if len(sys.argv) > 1 else 8
t0 = time.perf_counter()
out = ibsc_ensemble.run(K, 6, in_flight=K)
wall = time.perf_counter() - t0
print('K=%d wall %.1f s' % (K, wall))
labels = sorted({l for d in acc.values() for l in d})
for l in labels:
    v = [d[l] for d in acc.values() if l in d]
    c = [d[l] for d in cnt.values() if l in d]
    print('%-24s mean %.2f s per chain (%d calls, %.2f ms per call)' % (
        l, np.mean(v), np.mean(c), 1e3 * np.sum(v) / max(np.sum(c), 1)))

