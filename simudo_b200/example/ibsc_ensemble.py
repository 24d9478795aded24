"""BASELINE configs[4]: ensemble of intermediate-band cells x bias points over the GPUs.

The reference multiplexes device parameters of the four-layer IB cell over OS processes
(``example/fourlayer/fourlayer.py:233-284``, ``fourlayer_example.py:208-285``); here the
members of the Cartesian product NI x mu_I x sigma_opt_ci x IB_thickness are dealt to the
ranks of a torch.distributed job (one process per GPU), every rank keeps `--in-flight`
continuation chains going at once (threads + CUDA streams), no data-path collective:

  python -m simudo_b200.example.ibsc_ensemble --members 64 --bias-points 40 --in-flight 8
  python -m torch.distributed.run --nproc-per-node 8 -m simudo_b200.example.ibsc_ensemble ...

Band edges follow the fixture of tests/golden (CB 1.1 eV, IB 0.65 eV): at the reference's
2.0 eV gap the light ramp of this implementation still stalls.
"""
import argparse
import itertools
import json
import os
import time

import numpy as np

SMALL = dict(p_thickness=0.2, n_thickness=0.2, IB_thickness=0.4, mesh_start=0.01,
             mesh_factor=1.5, mesh_max=0.05, CB_E=1.1, IB_E=0.65, X=100)


def members(n=64, small=True):
    """4 x 4 x 2 x 2 device variants (fourlayer.py:233-284 style multiplexing)."""
    NI = [1e17, 2.5e17, 5e17, 1e18]
    mu_I = [50, 100, 200, 500]
    sig = [1e-14, 3e-14]
    thick = [0.4, 0.6] if small else [2.0, 3.0]
    out = []
    for a, b, c, d in itertools.product(NI, mu_I, sig, thick):
        out.append(dict(SMALL if small else {}, NI=a, mu_I=b, sigma_opt_ci=c, IB_thickness=d))
    return out[:n]


def run_member(params, backend_factory=None):
    from simudo_b200.example import fourlayer
    return fourlayer.run(P=params, backend_factory=backend_factory)


def run(n_members=64, bias_points=40, v_max=0.9, in_flight=8, dist=None, small=True):
    from simudo_b200.sweep import run_ensemble
    V = list(np.linspace(0.0, v_max, bias_points))
    ms = [dict(m, V_exts=V) for m in members(n_members, small)]
    t0 = time.perf_counter()
    res, times = run_ensemble(ms, run_member, dist, in_flight=in_flight)
    wall = time.perf_counter() - t0
    world = dist.get_world_size() if dist is not None and dist.is_initialized() else 1
    jsc = [float(-r[0][1]) for r in res]
    return dict(members=len(ms), bias_points=bias_points, in_flight=in_flight, n_gpus=world,
                wall_s=max(times) if times else wall, wall_s_per_rank=times,
                bias_points_per_s=len(ms) * bias_points / max(max(times), 1e-9),
                Jsc_mA_cm2_min_max=[min(jsc), max(jsc)], finite=bool(np.isfinite(jsc).all()))


def main():
    import torch
    import torch.distributed as dist
    ap = argparse.ArgumentParser()
    ap.add_argument('--members', type=int, default=64)
    ap.add_argument('--bias-points', type=int, default=40)
    ap.add_argument('--in-flight', type=int, default=8)
    a = ap.parse_args()
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if world > 1:
        torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', '0')))
        dist.init_process_group('nccl')
    out = run(a.members, a.bias_points, in_flight=a.in_flight, dist=dist if world > 1 else None)
    if int(os.environ.get('RANK', '0')) == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
