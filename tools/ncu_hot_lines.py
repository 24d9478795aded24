#!/usr/bin/env python
"""Hot CUDA source lines per kernel from an ncu report captured with --import-source on.
Usage: tools/ncu_hot_lines.py rep.ncu-rep [kernel-substring] [top-n]"""
import csv
import subprocess
import sys


def main(rep, filt='', top=25):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source',
                          'cuda,sass'], capture_output=True, text=True).stdout
    secs, cur = [], None
    for x in csv.reader(out.splitlines()):
        if not x:
            continue
        if x[0] == 'File Path':
            cur = dict(file=x[1].split('/')[-1], rows=[])
            secs.append(cur)
        elif x[0] == 'Function Name':
            cur['fn'] = x[1]
        elif x[0] == 'Line No':
            cur['hdr'] = x
        elif cur is not None and x[0].isdigit() and x[2] == '-':
            cur['rows'].append(x)
    by_fn = {}
    for s in secs:
        by_fn.setdefault(s['fn'], []).append(s)
    for fn, ss in by_fn.items():
        if filt not in fn:
            continue
        h = ss[0]['hdr']
        i_s, i_i = h.index('# Samples'), h.index('Instructions Executed')
        i_l = h.index('stall_long_sb')
        lines = [(int(x[i_s]), int(x[i_i]), int(x[i_l]), s['file'], x[0], x[1].strip())
                 for s in ss for x in s['rows']]
        tot = sum(l[0] for l in lines)
        print('==== %s: %d samples, %d warp-instructions' % (fn[:60], tot, sum(l[1] for l in lines)))
        for sm, ins, lsb, f, ln, src in sorted(lines, reverse=True)[:top]:
            print('%5.1f%% lsb=%5.1f%% inst=%9d %s:%s  %s' % (100. * sm / tot, 100. * lsb / tot,
                                                               ins, f[:12], ln, src[:100]))


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else '',
         int(sys.argv[3]) if len(sys.argv) > 3 else 25)
