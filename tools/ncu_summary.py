#!/usr/bin/env python
"""Turn an ``ncu --set full`` report into the small CSVs kept under profiles/:
``<name>_summary.csv`` (key metrics of the first captured launch) and
``<name>_opcodes.csv`` (executed warp-instructions and stall samples per SASS
opcode from the source page).  Usage: tools/ncu_summary.py rep.ncu-rep profiles/NAME
"""
import collections
import csv
import subprocess
import sys

KEEP = ['Kernel Name', 'Block Size', 'Grid Size', 'gpu__time_duration.sum',
        'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'launch__registers_per_thread', 'launch__occupancy_limit_registers',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum',
        'l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum',
        'l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum',
        'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct']


def main(rep, out):
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True,
                         text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    with open(out + '_summary.csv', 'w') as f:
        w = csv.writer(f)
        w.writerow(['metric', 'unit', 'value'])
        for h, u, v in zip(hdr, units, vals):
            if h in KEEP or (h.startswith('smsp__pcsamp_warps_issue_stalled')
                             and not h.endswith('not_issued')):
                w.writerow([h, u, v])
    src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True,
                         text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    hdr = rows[1]
    i_src, i_smp, i_ex = hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed')
    ex, smp = collections.Counter(), collections.Counter()
    for r in rows[2:]:
        try:
            s, e = int(r[i_smp]), int(r[i_ex])
        except (ValueError, IndexError):
            continue
        toks = r[i_src].split()
        op = toks[0] if toks else '?'
        if op.startswith('@') and len(toks) > 1:
            op = toks[1]
        op = op.split('.')[0]
        ex[op] += e
        smp[op] += s
    with open(out + '_opcodes.csv', 'w') as f:
        w = csv.writer(f)
        w.writerow(['opcode', 'warp_instructions_executed', 'stall_samples'])
        for op, e in ex.most_common(40):
            w.writerow([op, e, smp[op]])


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2])
