#!/usr/bin/env python3
"""Attribute sampled warp stalls to SASS opcodes: `ncu -i rep --page source --csv --print-source sass`.
Prints, per stall reason, which opcode classes the stalled samples sit on (an instruction is sampled while
it WAITS to issue, so e.g. long_sb samples on a DFMA mean "DFMA waiting for a load result")."""
import csv
import subprocess
import sys
from collections import defaultdict


def main(path, top=8):
    raw = subprocess.run(['ncu', '-i', path, '--page', 'source', '--csv', '--print-source', 'sass'],
                         stdout=subprocess.PIPE, text=True).stdout.splitlines()
    rows = list(csv.reader(raw[1:]))
    hdr = rows[0]
    col = {n: i for i, n in enumerate(hdr)}
    reasons = [n for n in hdr if n.startswith('stall_') and '(Not Issued)' not in n]
    tot = defaultdict(float)
    by = {r: defaultdict(float) for r in reasons}
    prev_ops = []
    waits_on = defaultdict(lambda: defaultdict(float))
    for r in rows[1:]:
        if len(r) < len(hdr):
            continue
        src = r[col['Source']].strip()
        parts = src.split()
        op = parts[1] if parts and parts[0].startswith('@') else (parts[0] if parts else '?')
        op = op.split('.')[0]
        for reason in reasons:
            try:
                v = float(r[col[reason]])
            except ValueError:
                v = 0
            if v:
                by[reason][op] += v
                tot[reason] += v
    grand = sum(tot.values())
    for reason in sorted(reasons, key=lambda x: -tot[x]):
        if tot[reason] < 0.01 * grand:
            continue
        ops = sorted(by[reason].items(), key=lambda x: -x[1])[:top]
        print(f'{reason:22s} {100 * tot[reason] / grand:5.1f}%  ' + ' '.join(f'{o}:{100 * v / tot[reason]:.0f}%' for o, v in ops))


if __name__ == '__main__':
    main(sys.argv[1])
