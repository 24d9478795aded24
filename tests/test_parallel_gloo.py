"""N > 1 path on CPU: world_size-2 gloo run of the strip halo exchange and the
sweep sharding (host-side logic of bench.py --gpus N)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from simudo_b200 import synthetic as syn
    from simudo_b200.parallel import StripHalo, shard_round_robin
    pa = syn.ib_problem(nx=9, ny=4)
    halo = StripHalo(pa, rank, world, 'cpu')
    u = torch.full((pa.N,), float(rank + 1), dtype=torch.float64)
    u += torch.arange(pa.N, dtype=torch.float64) * 1e-6
    base_w = torch.full((3, pa.n_cells), 10.0 * (rank + 1), dtype=torch.float64)
    top_before = u[halo.dofs_top].clone()
    halo.step(dist, u, base_w)
    res = dict(rank=rank, bytes=halo.bytes_per_exchange,
               n_iface=int(halo.cells_top.numel()),
               top=top_before.numpy(), bottom_after=u[halo.dofs_bottom].numpy(),
               ghost_below=halo.ghost_w_below.numpy().copy(),
               ghost_above=halo.ghost_w_above.numpy().copy(),
               shard=shard_round_robin(7, rank, world))
    # max-over-ranks timing reduction used by bench.py
    t = torch.tensor([float(rank + 3)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    res['tmax'] = float(t.item())
    q.put(res)
    dist.barrier()
    dist.destroy_process_group()


def test_strip_halo_exchange_and_sharding_world2():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted([q.get(timeout=120) for _ in procs], key=lambda d: d['rank'])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    r0, r1 = out
    # interface: 8 facets x (3 P = 12 dofs + 3 base_w) doubles, one neighbour each
    assert r0['n_iface'] == 8 + 3            # nx = 9 -> 8 intervals (+3 layer-edge points)
    assert r0['bytes'] == r1['bytes'] == 8 * r0['n_iface'] * (12 + 3)
    # owner -> ghost: rank 1's bottom interface dofs now equal rank 0's top dofs
    assert np.array_equal(r1['bottom_after'], r0['top'])
    assert np.all(r1['ghost_below'] == 10.0) and np.all(r0['ghost_above'] == 20.0)
    assert np.all(r0['ghost_below'] == 0.0)   # rank 0 has nobody below
    assert r0['shard'] == [0, 2, 4, 6] and r1['shard'] == [1, 3, 5]
    assert r0['tmax'] == r1['tmax'] == 4.0


def _ensemble_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from functools import partial
    from oracle_backend import OracleBackend
    from simudo_b200.example import sweep_ensemble as se
    from simudo_b200.sweep import run_ensemble
    mem = se.members(3)
    run = partial(se.run_member, backend_factory=OracleBackend, voltages=[0.0, 0.3, 0.6])
    res, times = run_ensemble(mem, run, dist)
    q.put(dict(rank=rank, res=res, n_times=len(times)))
    dist.barrier()
    dist.destroy_process_group()


def test_sweep_ensemble_sharded_equals_serial_world2():
    """config 5 host logic: members dealt to 2 ranks, gathered tables identical
    to a serial run; no data-path collective involved."""
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_ensemble_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted([q.get(timeout=600) for _ in procs], key=lambda d: d['rank'])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert out[0]['res'] == out[1]['res'] and out[0]['n_times'] == 2
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    from functools import partial
    from oracle_backend import OracleBackend
    from simudo_b200.example import sweep_ensemble as se
    from simudo_b200.sweep import run_ensemble
    serial, _ = run_ensemble(se.members(3), partial(se.run_member, backend_factory=OracleBackend,
                                                    voltages=[0.0, 0.3, 0.6]))
    assert serial == out[0]['res']
    # heavier doping -> larger built-in potential -> smaller forward current at 0.6 V
    j06 = [dict((round(v, 2), j) for v, j in r)[0.6] for r in serial]
    assert j06[0] > j06[2] > 0
