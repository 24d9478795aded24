"""Parameter / bias-sweep ensembles sharded across GPUs (BASELINE config 5).

The reference runs independent device-parameter instances as OS processes
(``multiprocessing.Pool(...).apply_async(run, ...)``,
``simudo/example/fourlayer/fourlayer_example.py:213,272``), one J-V
continuation chain each, with no communication.  Here the members of an
ensemble are dealt round-robin to the ranks of a ``torch.distributed`` job
(one process per GPU); every rank runs its chains on its own device and the
J-V tables are gathered on rank 0 with ``all_gather_object`` -- the only
collective, outside the data path.
"""
import time

from .parallel import shard_round_robin

__all__ = ['run_ensemble']


def _run_in_flight(items, fn, in_flight):
    """Run ``fn(item)`` for every item with up to ``in_flight`` of them at a time, each in its
    own thread and CUDA stream: one continuation chain is a sequence of small launches with a
    host decision after each Newton iteration, so a single chain leaves the GPU mostly idle;
    several chains interleave (ctypes and torch release the GIL while the device works)."""
    if in_flight <= 1 or len(items) <= 1:
        return [fn(x) for x in items]
    import concurrent.futures as cf
    import torch

    # the current CUDA device is per thread and a new thread starts on device 0: hand the
    # rank's device to its workers (otherwise every rank > 0 ran its chains on GPU 0)
    dev = torch.cuda.current_device() if torch.cuda.is_available() else None

    def worker(x):
        if dev is not None:
            torch.cuda.set_device(dev)
            with torch.cuda.stream(torch.cuda.Stream(device=dev)):
                r = fn(x)
                torch.cuda.current_stream().synchronize()
                return r
        return fn(x)
    with cf.ThreadPoolExecutor(max_workers=in_flight) as ex:
        return list(ex.map(worker, items))


def run_ensemble(members, run_member, dist=None, in_flight=1):
    """``members``: list of parameter dicts; ``run_member(params) -> result``
    (e.g. a list of (V, J)).  Members are dealt round-robin to the ranks (one process per
    GPU); a rank keeps ``in_flight`` of its members going at once (BASELINE config 5: 8
    devices per GPU in flight).  Returns ``(results, timings)`` on every rank:
    ``results[i]`` is member i's result, ``timings[r]`` the wall time of rank r."""
    rank = dist.get_rank() if dist is not None and dist.is_initialized() else 0
    world = dist.get_world_size() if dist is not None and dist.is_initialized() else 1
    mine = shard_round_robin(len(members), rank, world)
    t0 = time.perf_counter()
    out = _run_in_flight([members[i] for i in mine], run_member, in_flight)
    local = dict(zip(mine, out))
    dt = time.perf_counter() - t0
    if world == 1:
        return [local[i] for i in range(len(members))], [dt]
    gathered = [None] * world
    dist.all_gather_object(gathered, (local, dt))
    merged = {}
    for part, _ in gathered:
        merged.update(part)
    return [merged[i] for i in range(len(members))], [t for _, t in gathered]
