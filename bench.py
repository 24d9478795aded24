#!/usr/bin/env python
"""bench.py -- Jacobian+residual assembly throughput (DOF/s, FP64).

One "step" = one pass of the hot path (sb_assemble: residual + Jacobian of the coupled
Poisson / drift-diffusion system) over the mesh.  Workload: BASELINE.json configs[3], the 2D
intermediate-band cell on a synthetic product mesh of 708 x 708 points (1 003 940 cells,
4 sub-problems, 60 local dofs/cell, 3 optical fields, ~42 M dofs, 1.07e9 non-zeros).

N = 1: the whole mesh on one GPU.  N > 1: the SAME mesh decomposed into N y-strips
(simudo_b200/parallel.py: StripDecomposition -- one ghost row of quads per interior side,
owner computes, per step a ghost update of u / base_w between neighbouring ranks over NCCL
point-to-point, no collective on the data path): strong scaling, `value` = global dofs /
max-over-ranks time.  The weak-scaling number (every rank a full 1M-cell mesh) is reported
beside it under "weak".

  python bench.py --gpus 1 --steps 5 --warmup 3
  python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
  python bench.py --impl reference            # CPU arm (oracle port, all host cores)

Prints ONE JSON line (rank 0).
"""
import argparse
import hashlib
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


def build_hash():
    """Hash of the kernel sources: ties profiles/*_per_cell.json to the build it measured."""
    h = hashlib.sha256()
    d = os.path.join(ROOT, 'simudo_b200', 'csrc')
    for name in ('pdd_kernels.cu', 'pdd_device.cuh', 'native.inl', 'pdd_tables.h'):
        with open(os.path.join(d, name), 'rb') as f:
            h.update(f.read())
    return h.hexdigest()[:16]


def ncu_per_cell(kind, pattern):
    """DRAM bytes and FP64 instructions per cell of one assembly pass, from the committed
    `ncu --set full` capture of this build (profiles/r2_per_cell.json, written by
    tools/ncu_per_cell.py from the raw page of the .ncu-rep); None when the kernels changed
    since the capture."""
    p = os.path.join(ROOT, 'profiles', 'r2_per_cell.json')
    if not os.path.exists(p):
        return None
    with open(p) as f:
        d = json.load(f)
    e = d.get('%s/%s' % (kind, pattern))
    if e is None or e.get('build_hash') != build_hash():
        return None
    return e


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--nx', type=int, default=708, help='x points of the product mesh')
    ap.add_argument('--ny', type=int, default=708, help='y points per rank strip')
    ap.add_argument('--kind', default='ib', choices=['ib', 'pn'])
    ap.add_argument('--pattern', default='bsr3', choices=['bsr3', 'native', 'structural', 'dolfin'],
                    help="CSR layout: 'native' = structural non-zeros, columns in kernel emission "
                         "order (closed-form kernels); 'structural' = same entries, ascending "
                         "columns; 'dolfin' = dense cell blocks")
    ap.add_argument('--no-e2e', action='store_true', help='kernel experiments: skip the host-buffer arm')
    ap.add_argument('--outputs', default='Ab', choices=['Ab', 'A', 'b'],
                    help='kernel experiments: assemble Jacobian and residual, or only one of them')
    ap.add_argument('--cpu-cells', type=int, default=40000,
                    help='cells of the bounded CPU-baseline sample')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-solve', action='store_true',
                    help='skip the separately timed Newton linear solve (configs[2], 2D pn diode)')
    ap.add_argument('--no-weak', action='store_true', help='N > 1: skip the weak-scaling measurement')
    ap.add_argument('--no-sweep', action='store_true',
                    help='skip the J-V sweep ensemble (second half of the BASELINE metric)')
    ap.add_argument('--sweep-members', type=int, default=8, help='sweep members per GPU')
    ap.add_argument('--sweep-bias-points', type=int, default=6)
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


def make_problem(kind, nx, ny, pattern='structural', strip=None):
    from simudo_b200 import synthetic as syn
    mk = syn.ib_problem if kind == 'ib' else syn.pn_problem
    pa = mk(nx=nx, ny=ny, pattern=pattern, strip=strip)
    st = syn.synthetic_state(pa)
    return pa, st


def host_threads():
    """Host cores this process may use.  torchrun exports OMP_NUM_THREADS=1, which would
    silently run the CPU arm on one core: the thread count is always passed explicitly."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def cpu_sample(kind, cells, steps, warmup, threads):
    """Oracle (CPU port of the FEniCS-equivalent assembly) timed on the host cores on a
    bounded sample of the same workload: (DOF/s, seconds per pass, pa)."""
    from oracle.oracle import Oracle
    n = max(int(np.sqrt(cells / 2.0)) + 1, 4)
    pa, st = make_problem(kind, n, n)
    orc = Oracle(pa, n_threads=threads)
    args = (st['u'], st['base_w'], st['thmeq_phi'], st['phi_opt'], st['bc_values'],
            st['natbc_values'])
    for _ in range(max(warmup, 1)):
        orc.assemble(*args)                  # warm-up (page faults, caches)
    t0 = time.perf_counter()
    for _ in range(steps):
        orc.assemble(*args)
    dt = (time.perf_counter() - t0) / steps
    return pa.N / dt, dt, pa, n


def cpu_baseline(kind, cells, steps=2, threads=None):
    threads = threads or host_threads()
    value, dt, pa, n = cpu_sample(kind, cells, steps, 1, threads)
    return dict(value=value, unit=UNIT, cores=int(threads), kind='port',
                sample='%s mesh %dx%d points, %d cells, %d dofs, %d passes of A+b, %.2f s/pass'
                       % (kind, n, n, pa.n_cells, pa.N, steps, dt))


def workload_name(kind, nx, ny):
    """Same string on both arms (the driver matches them)."""
    return ('2D %s cell, synthetic product mesh nx=%d ny=%d, decomposed into y-strips over the '
            'GPUs (BASELINE configs[3])' % (kind, nx, ny))


def run_reference(args):
    """--impl reference: the reference's CPU path.  Real FEniCS cannot be installed here
    (dolfin / ffc / PETSc absent, no network), so this is the oracle port on all
    host cores, each step a bounded sample of the workload (DOF/s is size-normalised)."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    steps = max(args.steps, 1)
    threads = host_threads()
    value, dt, pa, n = cpu_sample(args.kind, args.cpu_cells, steps, args.warmup, threads)
    sample = ('%s mesh %dx%d points, %d cells, %d dofs per step (bounded sample of the '
              '~1M-cell workload), %d threads' % (args.kind, n, n, pa.n_cells, pa.N, threads))
    line = dict(metric=METRIC, value=value, unit=UNIT, impl='reference', n_gpus=args.gpus,
                steps=steps, warmup=args.warmup, ms_per_step=dt * 1e3,
                higher_is_better=True, scaling='strong', vs_baseline=None, dtype='f64',
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


def solve_timing():
    """The Newton linear solve, timed separately (north star): BASELINE configs[2], the Si pn
    diode on a 233 x 224-point mesh (~103 k cells, 3.3 M dofs), one bias point from thermal
    equilibrium on the GPU: native assembly + multifrontal LU (fem/multifrontal.py)."""
    from simudo_b200.example import pn_2d
    r = pn_2d.run(ny=224, bias=0.6, check=False)
    return dict(kind=r['solver'] + ': nested-dissection multifrontal LU, batched dense fronts '
                     '(cuSOLVER/cuBLAS via torch), symbolic phase reused, 2 refinement steps with '
                     'double-double residual',
                config='BASELINE configs[2]: 2D Si pn diode, %d cells, %d dofs, %d nnz, bias %.1f V'
                       % (r['cells'], r['dofs'], r['nnz'], r['bias_V']),
                factor_ms=r['factor_ms'], solve_ms=r['solve_ms'], ms=r['factor_ms'] + r['solve_ms'],
                assemble_ms=r['assemble_ms'], newton_wall_s=r['wall_s'],
                J_mA_cm2=r['J_mA_cm2'], fronts=r.get('fronts'))


def sweep_timing(args, dist):
    """J-V sweep wall time (BASELINE metric, second half; configs[4] at reduced size): an
    ensemble of four-layer IB cells (device-parameter variants, fourlayer.py:233-284), each a
    light ramp + bias continuation with self-consistent optics, dealt to the ranks, up to 8
    chains in flight per GPU (threads + streams), no data-path collective.  Weak scaling, as
    BASELINE configs[4] is specified (64 devices on 8 GPUs = 8 per GPU): `--sweep-members`
    members PER GPU, so N GPUs run 8 N members in about the wall time one GPU needs for 8 and
    `bias_points_per_s` is the scaling number (one continuation chain is latency-bound: a fixed
    set of 8 chains would not get faster on more GPUs)."""
    from simudo_b200.example import ibsc_ensemble
    world = dist.get_world_size() if dist is not None and dist.is_initialized() else 1
    try:
        r = ibsc_ensemble.run(args.sweep_members * world, args.sweep_bias_points, in_flight=8,
                              dist=dist)
        r['scaling'] = 'weak'
        r['members_per_gpu'] = args.sweep_members
        return r
    except Exception as e:                      # a failed chain must not lose the assembly numbers
        return dict(error='%s: %s' % (type(e).__name__, e))


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

    from simudo_b200 import _lib
    from simudo_b200.fem.assembly import AssemblyPlan
    from simudo_b200.fem.pipeline import PipelinedAssembler
    from simudo_b200.parallel import StripDecomposition
    import ctypes as C

    # ---- the workload: the global mesh (N = 1) or this rank's strip of it (N > 1) ----------
    strip = (rank, world) if world > 1 else None
    pa, st = make_problem(args.kind, args.nx, args.ny, args.pattern, strip=strip)
    dd = StripDecomposition(pa, dev) if world > 1 else None
    plan = AssemblyPlan(pa, device=dev)
    plan.set_profiling(True)
    T = plan.tensor
    n_owned_dofs = int(dd.owned.sum()) if dd else pa.N
    n_owned_cells = int(dd.owned_cells.sum()) if dd else pa.n_cells
    if world > 1:
        t = torch.tensor([n_owned_dofs, n_owned_cells], dtype=torch.int64, device=dev)
        dist.all_reduce(t)
        dofs_global, cells_global = int(t[0].item()), int(t[1].item())
    else:
        dofs_global, cells_global = pa.N, pa.n_cells

    # host (pinned) copies of what a caller changes per Newton step, the rest stays resident
    host = {k: torch.as_tensor(np.ascontiguousarray(st[k])).pin_memory() for k in ('u', 'base_w')}
    d = {k: v.to(dev) for k, v in host.items()}
    resident = {k: T(st[k]) for k in ('thmeq_phi', 'phi_opt') if st[k] is not None}
    bcv = T(st['bc_values'])
    nat = T(plan.natbc_device_order(st['natbc_values']))
    vals = plan.new_values()
    b = plan.new_vector()

    def step_device():
        if dd is not None:
            dd.exchange(dist, d['u'], d['base_w'])
        plan.assemble(d['u'], d.get('base_w'), resident.get('thmeq_phi'), resident.get('phi_opt'),
                      bcv, nat, vals if 'A' in args.outputs else None,
                      b if 'b' in args.outputs else None)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        """K steps bracketed by barrier + synchronize, CUDA events, max over ranks (ms)."""
        barrier()
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    warmup = max(args.warmup, 3)
    for _ in range(warmup):
        step_device()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches1 = plan.launch_count()
    total_ms = timed(step_device, args.steps)
    launches = plan.launch_count() - launches1
    # the pass itself: CUDA events recorded on the launch stream inside sb_assemble
    kernel_ms = []
    for _ in range(min(args.steps, 5)):
        step_device()
        kernel_ms.append(plan.last_kernel_ms())
    stage_ms = plan.last_stage_ms()
    k_ms = float(np.mean(kernel_ms))
    if world > 1:
        t = torch.tensor([k_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        k_ms = float(t.item())
    if args.no_e2e:
        if rank == 0:
            sampler.stop()
            print(json.dumps(dict(ms_per_step=total_ms / args.steps, kernel_ms=k_ms, stage_ms=stage_ms,
                                  native=plan.is_native(), cells=pa.n_cells, launches=int(launches))))
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- end to end through the public host-buffer API --------------------------------------
    # every step: H2D of the step's inputs (u, base_w) from pinned host memory, the ghost
    # exchange (N > 1), the assembly, D2H of the residual; thermal_equilibrium_phi and the
    # optical flux do not change per Newton step and stay resident; the Jacobian stays on the
    # device for its consumer, the on-device linear solve (sb_band_solve / MultifrontalSolver)
    host_b2 = [torch.empty(pa.N, dtype=torch.float64).pin_memory() for _ in range(2)]
    pre = (lambda dd_in: dd.exchange(dist, dd_in['u'], dd_in['base_w'])) if dd is not None else None
    pipe = PipelinedAssembler(plan, host, bcv, nat, depth=2, pre_assemble=pre, resident=resident)

    def step_e2e():
        pipe.submit(host, host_b2[pipe.step % 2])

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

    # ---- weak scaling beside it (N > 1): every rank one full mesh, no exchange --------------
    weak = None
    if world > 1 and not args.no_weak:
        del pipe
        pa_w, st_w = make_problem(args.kind, args.nx, args.ny, args.pattern)
        plan_w = AssemblyPlan(pa_w, device=dev)
        dw = {k: T(st_w[k]) for k in ('u', 'base_w', 'thmeq_phi', 'phi_opt') if st_w[k] is not None}
        bcw, natw = T(st_w['bc_values']), T(plan_w.natbc_device_order(st_w['natbc_values']))
        vw, bw = plan_w.new_values(), plan_w.new_vector()

        def step_w():
            plan_w.assemble(dw['u'], dw.get('base_w'), dw.get('thmeq_phi'), dw.get('phi_opt'),
                            bcw, natw, vw, bw)
        for _ in range(3):
            step_w()
        wms = timed(step_w, args.steps) / args.steps
        weak = dict(value=pa_w.N * world / (wms * 1e-3), unit=UNIT, ms_per_step=wms,
                    cells_per_gpu=pa_w.n_cells, note='every rank assembles one full mesh')

    ms_per_step = total_ms / args.steps
    value = dofs_global / (ms_per_step * 1e-3)
    e2e_value = dofs_global / (e2e_ms / args.steps * 1e-3)
    h2d = sum(v.numel() * 8 for v in host.values())
    d2h = host_b2[0].numel() * 8
    peak, peak_kind = peaks()
    alg_bytes = BYTES_PER_CELL[args.kind] * n_owned_cells      # this rank's share of the output
    achieved = alg_bytes / (k_ms * 1e-3) / 1e9
    finite = bool(torch.isfinite(b).all().item())
    fp64_peak = C.c_double(0.0)
    _lib.check(_lib.get().sb_measure_fp64_peak(C.byref(fp64_peak)))
    per_cell = ncu_per_cell(args.kind, args.pattern)

    cpu, solve = None, None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline(args.kind, args.cpu_cells)
    if rank == 0 and world == 1 and not args.no_solve:
        del vals
        torch.cuda.empty_cache()
        solve = solve_timing()
    sweep = None
    if not args.no_sweep:
        sweep = sweep_timing(args, dist if world > 1 else None)
    if rank == 0:
        pattern_txt = {'structural': 'structural non-zeros of the dolfin cell blocks, CSR, ascending columns',
                       'native': 'structural non-zeros of the dolfin cell blocks, CSR, columns in emission order',
                       'bsr3': 'structural non-zeros of the dolfin cell blocks, block CSR with 3 x 3 blocks',
                       'dolfin': 'dolfin dense cell blocks, CSR'}[pa.pattern]
        roof = dict(bound='hbm', achieved=achieved, peak=peak, unit='GB/s', frac=achieved / peak,
                    traffic=(per_cell['dram_bytes_per_cell'] * pa.n_cells if per_cell else None),
                    kernel='assembly pass = k_moments_nat + k_generation_nat + k_rows_nat x P + '
                           'k_rows_special (CUDA events around the pass on the launch stream, '
                           'inside sb_assemble)',
                    kernel_ms=k_ms, stage_ms=stage_ms, algorithmic_bytes=alg_bytes,
                    peak_source=peak_kind)
        if per_cell:
            roof['traffic_source'] = per_cell['source']
            fl = 2.0 * per_cell['fp64_instr_per_cell'] * pa.n_cells
            roof['fp64'] = dict(achieved=fl / (k_ms * 1e-3) / 1e12, peak=fp64_peak.value,
                                unit='TFLOP/s', frac=fl / (k_ms * 1e-3) / 1e12 / fp64_peak.value,
                                note='FP64 pipe instructions per pass (ncu, x 2 flop) / pass time, '
                                     'against the DFMA throughput measured now (sb_measure_fp64_peak)')
        else:
            roof['fp64'] = dict(achieved=None, peak=fp64_peak.value, unit='TFLOP/s', frac=None,
                                note='no ncu capture of this build under profiles/')
        line = dict(
            metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps,
            warmup=warmup, ms_per_step=ms_per_step, higher_is_better=True,
            scaling='strong', vs_baseline=None, dtype='f64', data='synthetic',
            config=dict(
                workload=workload_name(args.kind, args.nx, args.ny),
                x_points=int(len(np.unique(pa.mesh.coordinates[:, 0]))),
                cells=cells_global, dofs=dofs_global,
                cells_this_rank=pa.n_cells, owned_cells_this_rank=n_owned_cells,
                nnz_this_rank=pa.csr.nnz, sub_problems=pa.P, local_dofs=pa.L,
                optical_fields=pa.model.n_fields, pattern=pattern_txt, numbering=pa.numbering,
                native_kernels=plan.is_native(),
                parallelism=('domain decomposition into %d y-strips, one ghost row of quads per '
                             'interior side, owner computes; per step NCCL send/recv of the ghost '
                             'u / base_w values with the neighbouring ranks, %d B sent by rank 0'
                             % (world, dd.bytes_per_exchange)) if world > 1 else 'single GPU',
                l2='inputs (%.0f MB) and outputs (%.1f GB) far exceed the 126 MB L2'
                   % ((h2d + sum(v.numel() * 8 for v in resident.values())) / 1e6, alg_bytes / 1e9),
                finite=finite),
            roofline=roof,
            e2e=dict(value=e2e_value, unit=UNIT, h2d_bytes_per_step=h2d,
                     d2h_bytes_per_step=d2h, ms_per_step=e2e_ms / args.steps,
                     what='per step: H2D of u and base_w from pinned host memory, ghost exchange, '
                          'assembly, D2H of the residual; thermal_equilibrium_phi and the optical '
                          'flux stay resident (they do not change per Newton step); the Jacobian '
                          'stays on the device for the on-device linear solve',
                     pipelined='depth 2: H2D, kernels and D2H of neighbouring steps overlap '
                               '(simudo_b200.fem.pipeline.PipelinedAssembler)'),
            gpu_launches=int(launches),
            clocks=clocks)
        if weak is not None:
            line['weak'] = weak
        if cpu is not None:
            line['cpu_baseline'] = cpu
        if solve is not None:
            line['solve'] = solve
        if sweep is not None:
            line['sweep'] = sweep
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
