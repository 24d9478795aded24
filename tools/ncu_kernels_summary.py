#!/usr/bin/env python
"""One column per captured launch of an ``ncu --set full`` report: the key metrics and
the warp-stall sample counts kept under profiles/ (``*_kernels_summary.csv``).
Usage: tools/ncu_kernels_summary.py rep.ncu-rep out.csv"""
import csv
import subprocess
import sys

KEEP = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'launch__registers_per_thread',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum',
        'l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum',
        'l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum',
        'lts__t_sector_hit_rate.pct',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed']
STALL = 'smsp__pcsamp_warps_issue_stalled_'


def main(rep, out):
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True,
                         text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, launches = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    names = [r[col['Kernel Name']][:30] for r in launches]
    stalls = sorted(h for h in hdr if h.startswith(STALL) and not h.endswith('not_issued'))
    with open(out, 'w') as f:
        w = csv.writer(f)
        w.writerow(['metric', 'unit'] + names)
        for h in KEEP:
            if h in col:
                w.writerow([h, units[col[h]]] + [r[col[h]] for r in launches])
        for h in stalls:
            w.writerow(['stall_' + h[len(STALL):], units[col[h]]] + [r[col[h]] for r in launches])


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2])
