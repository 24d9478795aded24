"""Ensemble of pn-diode devices x bias points sharded across GPUs.

  python -m torch.distributed.run --nproc-per-node N -m simudo_b200.example.sweep_ensemble

Each member is the tutorial diode with its own doping / SRH lifetime (the
analogue of the reference's ``fourlayer`` parameter multiplexing,
``example/fourlayer/fourlayer.py:233-284``); members are independent J-V
chains, one per GPU at a time, no data-path collective.
"""
import itertools
import json
import os

import numpy as np


def members(n=8):
    dopings = [5e17, 1e18, 2e18, 4e18]
    taus = [1e-9, 1e-8]
    out = [dict(doping=d, tau_n=t) for d, t in itertools.product(dopings, taus)]
    return out[:n]


def run_member(params, backend_factory=None, voltages=None):
    from simudo_b200.example import pn_diode_param
    return pn_diode_param.run(backend_factory=backend_factory, voltages=voltages, **params)


def main():
    import torch
    import torch.distributed as dist
    from simudo_b200.sweep import run_ensemble
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if world > 1:
        torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', '0')))
        dist.init_process_group('nccl')
    res, times = run_ensemble(members(), run_member, dist if world > 1 else None)
    if int(os.environ.get('RANK', '0')) == 0:
        print(json.dumps(dict(n_members=len(res), n_gpus=world, wall_s_per_rank=times,
                              wall_s=max(times),
                              J_at_0p6V=[float(dict((round(v, 2), j) for v, j in r)[0.6])
                                         for r in res])))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
