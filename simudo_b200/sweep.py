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


def run_ensemble(members, run_member, dist=None):
    """``members``: list of parameter dicts; ``run_member(params) -> result``
    (e.g. a list of (V, J)).  Returns ``(results, timings)`` on every rank:
    ``results[i]`` is member i's result, ``timings[r]`` the wall time of rank r."""
    rank = dist.get_rank() if dist is not None and dist.is_initialized() else 0
    world = dist.get_world_size() if dist is not None and dist.is_initialized() else 1
    mine = shard_round_robin(len(members), rank, world)
    t0 = time.perf_counter()
    local = {i: run_member(members[i]) for i in mine}
    dt = time.perf_counter() - t0
    if world == 1:
        return [local[i] for i in range(len(members))], [dt]
    gathered = [None] * world
    dist.all_gather_object(gathered, (local, dt))
    merged = {}
    for part, _ in gathered:
        merged.update(part)
    return [merged[i] for i in range(len(members))], [t for _, t in gathered]
