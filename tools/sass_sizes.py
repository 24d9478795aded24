#!/usr/bin/env python
"""SASS instruction count and opcode histogram per kernel of a shared object.
Usage: tools/sass_sizes.py lib.so [name-substring]"""
import collections
import re
import subprocess
import sys


def main(so, filt=''):
    out = subprocess.run(['cuobjdump', '-sass', so], capture_output=True, text=True).stdout
    name, n, ops = None, collections.Counter(), collections.defaultdict(collections.Counter)
    for line in out.splitlines():
        m = re.search(r'Function : (\S+)', line)
        if m:
            name = m.group(1)
            continue
        m = re.match(r'\s+/\*[0-9a-f]+\*/\s+(.*?);', line)
        if m and name:
            ins = m.group(1).split()
            op = ins[1] if ins[0].startswith('@') else ins[0]
            n[name] += 1
            ops[name][op.split('.')[0]] += 1
    for k in sorted(n):
        if filt in k:
            top = ', '.join('%s %d' % kv for kv in ops[k].most_common(8))
            print('%6d  %s  [%s]' % (n[k], k, top))


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else '')
