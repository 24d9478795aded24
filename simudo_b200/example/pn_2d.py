"""BASELINE configs[2]: single-bias Newton solve of the Si pn diode on a 2D product mesh.

The tutorial diode (``simudo/example/pn_diode/pn_diode.py``) on ``Xs x Ys`` with ~224 x 224
points (~100 k cells, P = 3, N ~ 3.2 M): thermal equilibrium, then one bias point.  Every
Newton step = native assembly kernels + multifrontal LU (``fem/multifrontal.py``); the
script reports where the time goes and checks the terminal current against the same
x-mesh solved as a thin strip (banded solver) -- data and contacts do not depend on y, so
both must give the same J.

    python -m simudo_b200.example.pn_2d [--ny 224] [--bias 0.6]
"""
import argparse
import json
import logging
import time

import numpy as np
import torch

from simudo_b200.example import pn_diode

MESH = dict(mesh_start=2.5e-4, mesh_factor=1.06, edge_length=0.5 / 80)     # 233 x-points


def run(ny=224, bias=0.6, height=0.5, check=True):
    out = {}
    t0 = time.perf_counter()
    jv, problem, _ = pn_diode.run(voltages=[0.0, bias], return_problem=True, mesh=MESH,
                                  product2d_Ys=np.linspace(0.0, height, ny))
    torch.cuda.synchronize()
    out['wall_s'] = time.perf_counter() - t0
    sysm = problem.pdd.system
    be = sysm.backend
    pa = sysm.pa
    out.update(cells=pa.n_cells, dofs=pa.N, nnz=pa.csr.nnz, native_kernels=be.plan.is_native(),
               solver=type(be._solver).__name__, J_mA_cm2=jv[-1][1], bias_V=bias)
    if hasattr(be._solver, 'info'):
        out['fronts'] = be._solver.info()
    # time one assembly / factorisation / solve of the converged state
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    be.assemble(); be.solve()                      # warm
    ev[0].record(); be.assemble(); ev[1].record()
    be._solver.factor(be.vals); ev[2].record()
    be._solver.solve(be.vals, be.b, be.du, refactor=False); ev[3].record()
    torch.cuda.synchronize()
    out['assemble_ms'] = ev[0].elapsed_time(ev[1])
    out['factor_ms'] = ev[1].elapsed_time(ev[2])
    out['solve_ms'] = ev[2].elapsed_time(ev[3])
    if check:
        jv1 = pn_diode.run(voltages=[0.0, bias], mesh=MESH)        # thin strip, banded LU
        out['J_strip_mA_cm2'] = jv1[-1][1]
        out['rel_diff_vs_strip'] = abs(jv[-1][1] - jv1[-1][1]) / abs(jv1[-1][1])
    return out


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--ny', type=int, default=224)
    ap.add_argument('--bias', type=float, default=0.6)
    ap.add_argument('--verbose', action='store_true')
    a = ap.parse_args()
    if a.verbose:
        logging.basicConfig(level=logging.INFO)
    print(json.dumps(run(ny=a.ny, bias=a.bias)))
