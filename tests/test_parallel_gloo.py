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


def _dd_worker(rank, world, port, q, pattern):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    import conftest
    from simudo_b200 import _lib
    _lib._use_library_for_tests(conftest.build_emulation_library())
    from simudo_b200 import synthetic as syn
    from simudo_b200.fem.assembly import AssemblyPlan
    from simudo_b200.parallel import StripDecomposition
    from parity_util import run_plan
    pa = syn.ib_problem(nx=7, ny=9, pattern=pattern, strip=(rank, world))
    dd = StripDecomposition(pa)
    st = syn.synthetic_state(pa)
    # ghosts start as garbage: the exchange must restore them from their owners
    u = torch.as_tensor(st['u']).clone()
    bw = torch.as_tensor(st['base_w']).clone()
    u[torch.as_tensor(~dd.owned)] = 1e30
    bw[:, torch.as_tensor(~dd.owned_cells)] = -1e30
    dd.exchange(dist, u, bw)
    assert np.array_equal(u.numpy(), st['u']) and np.array_equal(bw.numpy(), st['base_w'])
    st['u'], st['base_w'] = u.numpy(), bw.numpy()
    plan = AssemblyPlan(pa)
    assert plan.is_native()
    vals, b = run_plan(plan, st)
    col, row = pa.csr.column_indices(want_rows=True)
    own = dd.owned[row]
    q.put(dict(rank=rank, bytes=dd.bytes_per_exchange, rk=dd.gkey[row][own], ck=dd.gkey[col][own],
               v=vals.numpy()[own], bk=dd.gkey[dd.owned], b=b.numpy()[dd.owned],
               n_owned_cells=int(dd.owned_cells.sum())))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize('pattern,world', [('bsr3', 2), ('native', 2), ('bsr3', 4)])
def test_strip_domain_decomposition_reproduces_single_domain_world2(emulated_lib, pattern, world):
    """One IB mesh split into y-strips (one ghost row of quads per interior side, owner
    computes, ghost update of u / base_w over gloo): the Jacobian and residual rows every rank
    owns are, bit for bit, the rows of the single-domain assembly; together they are all rows.
    world = 4 has middle ranks with two neighbours (what 6 of the 8 GPUs of a box are)."""
    from simudo_b200 import synthetic as syn
    from simudo_b200.fem.assembly import AssemblyPlan
    from simudo_b200.parallel import StripDecomposition
    from parity_util import run_plan
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + (7 if pattern == 'native' else 0) + 13 * world
    procs = [ctx.Process(target=_dd_worker, args=(r, world, port, q, pattern)) for r in range(world)]
    for p in procs:
        p.start()
    out = sorted([q.get(timeout=300) for _ in procs], key=lambda d: d['rank'])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single domain, same kernels
    g = syn.ib_problem(nx=7, ny=9, pattern=pattern)
    g.strip = dict(rank=0, world=1, j0=0, j1=8, g_lo=0, g_hi=0, row0=0, n_rows_global=8,
                   Ys_global=np.linspace(0.0, 1.0, 9))
    G = StripDecomposition(g)
    vals, b = run_plan(AssemblyPlan(g), syn.synthetic_state(g))
    col, row = g.csr.column_indices(want_rows=True)
    ref = {(int(r), int(c)): v for r, c, v in zip(G.gkey[row], G.gkey[col], vals.numpy())}
    bref = dict(zip(G.gkey.tolist(), b.numpy().tolist()))
    seen = 0
    for o in out:
        for r, c, v in zip(o['rk'], o['ck'], o['v']):
            assert ref[(int(r), int(c))] == v          # bitwise
        for k, v in zip(o['bk'], o['b']):
            assert bref[int(k)] == v
        seen += len(o['v'])
    assert seen == g.csr.nnz                            # the owned rows are all rows, once
    assert sum(o['n_owned_cells'] for o in out) == g.n_cells
    assert all(o['bytes'] > 0 for o in out)
    if world > 2:                                       # a middle rank talks to two neighbours
        assert out[1]['bytes'] > out[0]['bytes']
