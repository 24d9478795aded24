#!/usr/bin/env python
"""Per-cell DRAM traffic and FP64 instruction count of one assembly pass from an
`ncu --set full` report that captured exactly the launches of one pass.

    tools/ncu_per_cell.py rep.ncu-rep <cells> <kind/pattern> [out.json]

Writes / updates profiles/r2_per_cell.json: bench.py reads `traffic` and the FP64 roofline
numerator from it (only while the kernel sources still hash to `build_hash`).
"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main(rep, cells, key, out=None):
    from bench import build_hash
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True,
                         text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, launches = rows[0], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    units = rows[1]

    def val(r, name):
        v = float(r[col[name]].replace(',', ''))
        u = units[col[name]].lower()
        scale = {'byte': 1.0, 'kbyte': 1e3, 'mbyte': 1e6, 'gbyte': 1e9}.get(u, 1.0)
        return v * scale
    dram = fp64 = time_us = 0.0
    names = []
    for r in launches:
        names.append(r[col['Kernel Name']].split('(')[0])
        dram += val(r, 'dram__bytes_read.sum') + val(r, 'dram__bytes_write.sum')
        cyc = val(r, 'smsp__cycles_elapsed.avg')
        for op in ('dfma', 'dmul', 'dadd'):
            fp64 += val(r, 'smsp__sass_thread_inst_executed_op_%s_pred_on.sum.per_cycle_elapsed' % op) * cyc
        time_us += val(r, 'gpu__time_duration.sum')
    cells = int(cells)
    entry = dict(build_hash=build_hash(), cells=cells, launches=names,
                 dram_bytes_per_cell=dram / cells, fp64_instr_per_cell=fp64 / cells,
                 pass_us_under_ncu=time_us,
                 source='ncu --set full of one pass at %d cells (%s), %s' % (
                     cells, key, os.path.basename(rep)))
    out = out or os.path.join(ROOT, 'profiles', 'r2_per_cell.json')
    d = {}
    if os.path.exists(out):
        with open(out) as f:
            d = json.load(f)
    d[key] = entry
    with open(out, 'w') as f:
        json.dump(d, f, indent=1)
    print(json.dumps(entry, indent=1))


if __name__ == '__main__':
    main(*sys.argv[1:])
