#!/usr/bin/env python
"""bench.py -- Jacobian+residual assembly throughput (DOF/s, FP64).

One "step" = one pass of the hot path (sb_assemble: residual + Jacobian of the
coupled Poisson / drift-diffusion system) over the mesh.  Workload at N = 1 is
BASELINE.json configs[3]: the 2D intermediate-band cell on a synthetic ~1M-cell
product mesh (4 sub-problems, 60 local dofs/cell, 3 optical fields, ~42 M dofs).
At N > 1 every rank owns one y-strip of that size (weak scaling); the only
exchange a Newton step needs between strips -- ghost BDM facet dofs and ghost
base_w across the strip interfaces -- is done with NCCL point-to-point inside the
timed step.

  python bench.py --gpus 1 --steps 5 --warmup 3
  python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
  python bench.py --impl reference            # CPU arm (oracle port, host cores)

Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = 'jacobian_residual_assembly_dof_per_s'
UNIT = 'DOF/s'
BYTES_PER_CELL = {'ib': 10396, 'pn': 7424}
# dram__bytes_read.sum + dram__bytes_write.sum summed over the kernels of one assembly pass,
# per cell, from the committed `ncu --set full` capture (profiles/r1_s2_kernels_summary.csv,
# 320 396 IB cells: 3.662 GB read + 3.970 GB written over the 7 launches of one pass)
NCU_DRAM_BYTES_PER_CELL = {('ib', 'structural'): 7.632e9 / 320396.0}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--nx', type=int, default=708, help='x points of the product mesh')
    ap.add_argument('--ny', type=int, default=708, help='y points per rank strip')
    ap.add_argument('--kind', default='ib', choices=['ib', 'pn'])
    ap.add_argument('--pattern', default='native', choices=['bsr3', 'native', 'structural', 'dolfin'],
                    help="CSR layout: 'native' = structural non-zeros, columns in kernel emission "
                         "order (closed-form kernels); 'structural' = same entries, ascending "
                         "columns; 'dolfin' = dense cell blocks")
    ap.add_argument('--no-e2e', action='store_true', help='kernel experiments: skip the host-buffer arm')
    ap.add_argument('--outputs', default='Ab', choices=['Ab', 'A', 'b'],
                    help='kernel experiments: assemble Jacobian and residual, or only one of them')
    ap.add_argument('--cpu-cells', type=int, default=40000,
                    help='cells of the bounded CPU-baseline sample')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)['hbm_gbs']), 'measured'
    return 6650.0, 'fallback'


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index = index
        self.samples = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                 '--format=csv,noheader,nounits', '-lms', '100'],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=['nvidia-smi unavailable'])
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for s in self.samples:
            parts = [x.strip() for x in s.split(',')]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx = float(parts[1])
            except ValueError:
                continue
            for n, v in zip(names, parts[2:6]):
                if v.lower().startswith('active'):
                    reasons.add(n)
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=mx,
                    reasons=sorted(reasons), samples=len(sm))


def make_problem(kind, nx, ny, pattern='structural'):
    from simudo_b200 import synthetic as syn
    mk = syn.ib_problem if kind == 'ib' else syn.pn_problem
    pa = mk(nx=nx, ny=ny, pattern=pattern)
    st = syn.synthetic_state(pa)
    return pa, st


def cpu_baseline(kind, cells, steps=2, threads=None):
    """Oracle (CPU port of the FEniCS-equivalent assembly) timed on the host
    cores on a bounded sample of the same workload."""
    from oracle.oracle import Oracle, lib
    n = max(int(np.sqrt(cells / 2.0)) + 1, 4)
    pa, st = make_problem(kind, n, n)
    threads = threads or lib().pdd_oracle_max_threads()
    orc = Oracle(pa, n_threads=threads)
    args = (st['u'], st['base_w'], st['thmeq_phi'], st['phi_opt'], st['bc_values'],
            st['natbc_values'])
    orc.assemble(*args)                      # warm-up (page faults, caches)
    t0 = time.perf_counter()
    for _ in range(steps):
        orc.assemble(*args)
    dt = (time.perf_counter() - t0) / steps
    return dict(value=pa.N / dt, unit=UNIT, cores=int(threads), kind='port',
                sample='%s mesh %dx%d points, %d cells, %d dofs, %d passes of A+b, %.2f s/pass'
                       % (kind, n, n, pa.n_cells, pa.N, steps, dt)), dt, pa


def workload_name(kind, nx, ny):
    """Same string on both arms (the driver matches them)."""
    return ('2D %s cell, synthetic product mesh nx=%d ny=%d per GPU (BASELINE configs[3])'
            % (kind, nx, ny))


def run_reference(args):
    """--impl reference: the reference's CPU path.  Real FEniCS cannot be
    installed here, so this is the oracle port on all host cores."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    steps = max(args.steps, 1)
    from oracle.oracle import Oracle, lib
    n = max(int(np.sqrt(args.cpu_cells / 2.0)) + 1, 4)
    pa, st = make_problem(args.kind, n, n)
    threads = lib().pdd_oracle_max_threads()
    orc = Oracle(pa, n_threads=threads)
    a = (st['u'], st['base_w'], st['thmeq_phi'], st['phi_opt'], st['bc_values'],
         st['natbc_values'])
    for _ in range(max(args.warmup, 1)):
        orc.assemble(*a)
    t0 = time.perf_counter()
    for _ in range(steps):
        orc.assemble(*a)
    dt = (time.perf_counter() - t0) / steps
    value = pa.N / dt
    sample = ('%s mesh %dx%d points, %d cells, %d dofs per step (bounded sample of the '
              '~1M-cell workload)' % (args.kind, n, n, pa.n_cells, pa.N))
    line = dict(metric=METRIC, value=value, unit=UNIT, impl='reference', n_gpus=args.gpus,
                steps=steps, warmup=args.warmup, ms_per_step=dt * 1e3,
                higher_is_better=True, scaling='weak', vs_baseline=None, dtype='f64',
                data='synthetic',
                config=dict(workload=workload_name(args.kind, args.nx, args.ny),
                            sample_cells=pa.n_cells, sample_dofs=pa.N,
                            note='each step = one assembly pass over a bounded sample of the '
                                 'workload (same mesh family, state and physics)'),
                cpu_baseline=dict(value=value, unit=UNIT, cores=int(threads), kind='port',
                                  sample=sample),
                e2e=dict(value=value, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                gpu_launches=0)
    print(json.dumps(line))


def run_b200(args):
    import torch
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise RuntimeError('bench.py --impl b200 needs a CUDA device (no CPU fallback)')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=dev)

    from simudo_b200.fem.assembly import AssemblyPlan
    pa, st = make_problem(args.kind, args.nx, args.ny, args.pattern)
    plan = AssemblyPlan(pa, device=dev)
    plan.set_profiling(True)
    T = plan.tensor

    # host (pinned) copies of every per-step input, and the device state
    host = {}
    for k in ('u', 'base_w', 'thmeq_phi', 'phi_opt'):
        if st[k] is not None:
            host[k] = torch.as_tensor(np.ascontiguousarray(st[k])).pin_memory()
    host_b = torch.empty(pa.N, dtype=torch.float64).pin_memory()
    d = {k: v.to(dev) for k, v in host.items()}
    bcv = T(st['bc_values'])
    nat = T(plan.natbc_device_order(st['natbc_values']))
    vals = plan.new_values()
    b = plan.new_vector()
    # ghost exchange across the strip interfaces (simudo_b200/parallel.py)
    from simudo_b200.parallel import StripHalo
    halo = StripHalo(pa, rank, world, dev)

    def step_device():
        if world > 1:
            halo.step(dist, d['u'], d['base_w'])
        plan.assemble(d['u'], d.get('base_w'), d.get('thmeq_phi'), d.get('phi_opt'),
                      bcv, nat, vals if 'A' in args.outputs else None,
                      b if 'b' in args.outputs else None)

    # end-to-end arm: host buffers in, residual out, through the public pipelined API
    # (copies of neighbouring steps overlap the kernels; every step still moves all of
    # its inputs host->device and its residual device->host)
    from simudo_b200.fem.pipeline import PipelinedAssembler
    host_b2 = [host_b, torch.empty(pa.N, dtype=torch.float64).pin_memory()]
    pipe = None

    def make_pipe():
        pre = (lambda dd: halo.step(dist, dd['u'], dd['base_w'])) if world > 1 else None
        return PipelinedAssembler(plan, host, bcv, nat, depth=2, pre_assemble=pre)

    def step_e2e():
        pipe.submit(host, host_b2[pipe.step % 2])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        kms = []
        e0.record()
        for _ in range(steps):
            fn()
            kms.append(None)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    launches0 = plan.launch_count()
    for _ in range(max(args.warmup, 3)):
        step_device()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    # kernel-only timing of the dominant kernel: per-launch CUDA events (in the library)
    kernel_ms = []
    barrier()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    launches1 = plan.launch_count()
    e0.record()
    for _ in range(args.steps):
        step_device()
    e1.record()
    barrier()
    total_ms = e0.elapsed_time(e1)
    kernel_ms.append(plan.last_kernel_ms())
    if world > 1:
        t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    launches = plan.launch_count() - launches1
    # a few extra launches to average the dominant kernel's own duration
    for _ in range(min(args.steps, 5)):
        step_device()
        kernel_ms.append(plan.last_kernel_ms())
    stage_ms = plan.last_stage_ms()
    k_ms = float(np.mean(kernel_ms))
    # end-to-end through the public API with host buffers
    if args.no_e2e:
        if rank == 0:
            sampler.stop()
            print(json.dumps(dict(ms_per_step=total_ms / args.steps, kernel_ms=k_ms, stage_ms=stage_ms,
                                  native=plan.is_native(), cells=pa.n_cells, launches=int(launches))))
        return
    pipe = make_pipe()
    for _ in range(3):
        step_e2e()
    pipe.wait()

    def e2e_run(steps):
        barrier()
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for sname in ('s_h2d', 's_cmp', 's_d2h'):
            getattr(pipe, sname).wait_event(e0)
        for _ in range(steps):
            step_e2e()
        pipe.join()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms
    e2e_ms = e2e_run(args.steps)
    assert torch.equal(host_b2[0], host_b2[1]) and bool(torch.isfinite(host_b2[0]).all())
    clocks = sampler.stop() if rank == 0 else None

    ms_per_step = total_ms / args.steps
    dofs_total = pa.N * world
    value = dofs_total / (ms_per_step * 1e-3)
    e2e_value = dofs_total / (e2e_ms / args.steps * 1e-3)
    h2d = sum(v.numel() * 8 for v in host.values())
    d2h = host_b.numel() * 8
    peak, peak_kind = peaks()
    alg_bytes = BYTES_PER_CELL[args.kind] * pa.n_cells
    achieved = alg_bytes / (k_ms * 1e-3) / 1e9
    finite = bool(torch.isfinite(b).all().item())

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu, _, _ = cpu_baseline(args.kind, args.cpu_cells)
    if rank == 0:
        line = dict(
            metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps,
            warmup=max(args.warmup, 3), ms_per_step=ms_per_step, higher_is_better=True,
            scaling='weak', vs_baseline=None, dtype='f64', data='synthetic',
            config=dict(
                workload=workload_name(args.kind, args.nx, args.ny),
                x_points=int(len(np.unique(pa.mesh.coordinates[:, 0]))),
                cells_per_gpu=pa.n_cells, dofs_per_gpu=pa.N, nnz_per_gpu=pa.csr.nnz,
                sub_problems=pa.P, local_dofs=pa.L, optical_fields=pa.model.n_fields,
                pattern={'structural': 'structural non-zeros of the dolfin cell blocks, ascending columns',
                         'native': 'structural non-zeros of the dolfin cell blocks, columns in emission order',
                         'bsr3': 'structural non-zeros of the dolfin cell blocks as block CSR with 3 x 3 blocks',
                         'dolfin': 'dolfin dense cell blocks'}[pa.pattern],
                numbering=pa.numbering, native_kernels=plan.is_native(),
                parallelism=('y-strips x%d, NCCL ghost exchange %d B/rank/step' % (world, halo.bytes_per_exchange)) if world > 1 else 'single GPU',
                l2='inputs (%.0f MB) and outputs (%.1f GB) far exceed the 126 MB L2'
                   % (h2d / 1e6, (alg_bytes) / 1e9),
                finite=finite),
            roofline=dict(bound='hbm', achieved=achieved, peak=peak, unit='GB/s',
                          frac=achieved / peak,
                          traffic=(NCU_DRAM_BYTES_PER_CELL[(args.kind, args.pattern)] * pa.n_cells
                                   if (args.kind, args.pattern) in NCU_DRAM_BYTES_PER_CELL else None),
                          kernel='assembly pass = k_moments + k_generation + k_rows_staged x P + '
                                 'k_rows_special (CUDA events around the pass inside sb_assemble)',
                          kernel_ms=k_ms, stage_ms=stage_ms,
                          algorithmic_bytes=alg_bytes, peak_source=peak_kind),
            e2e=dict(value=e2e_value, unit=UNIT, h2d_bytes_per_step=h2d,
                     d2h_bytes_per_step=d2h, ms_per_step=e2e_ms / args.steps,
                     pipelined='depth 2: H2D, kernels and D2H of neighbouring steps overlap '
                               '(simudo_b200.fem.pipeline.PipelinedAssembler)'),
            gpu_launches=int(launches),
            clocks=clocks)
        if cpu is not None:
            line['cpu_baseline'] = cpu
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_b200(args)


if __name__ == '__main__':
    main()
