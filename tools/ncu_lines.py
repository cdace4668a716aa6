#!/usr/bin/env python3
"""Attribute the sampled warp stalls of an .ncu-rep to CUDA source lines (needs -lineinfo + --import-source on):
`ncu -i rep --page source --csv --print-source cuda,sass` lists every source line with its aggregated sample
count.  Usage: ncu_lines.py rep [file-substring] [start:name ...]  -- with start:name pairs the samples are
summed per phase (a phase runs from its start line to the next one)."""
import csv
import subprocess
import sys
from collections import defaultdict


def main(path, fname, phases):
    raw = subprocess.run(['ncu', '-i', path, '--page', 'source', '--csv', '--print-source', 'cuda,sass'],
                         stdout=subprocess.PIPE, text=True).stdout.splitlines()
    cur = None
    by = defaultdict(lambda: defaultdict(float))
    col = None
    stall_cols = []
    for r in csv.reader(raw):
        if len(r) == 2 and r[0] == 'File Path':
            cur = r[1]
            continue
        if r and r[0] == 'Line No':
            col = {n: i for i, n in enumerate(r)}
            stall_cols = [n for n in r if n.startswith('stall_') and 'Not Issued' not in n]
            continue
        if col is None or not r or not r[0].strip().isdigit() or len(r) < len(col):
            continue
        if fname and fname not in (cur or ''):
            continue
        line = int(r[0])
        try:
            by[line]['all'] += float(r[col['# Samples']])
        except ValueError:
            continue
        for s in stall_cols:
            try:
                by[line][s] += float(r[col[s]])
            except ValueError:
                pass
    tot = sum(v['all'] for v in by.values()) or 1.0
    if phases:
        ph = sorted((int(a.split(':')[0]), a.split(':')[1]) for a in phases)
        agg = defaultdict(lambda: defaultdict(float))
        for line, v in by.items():
            name = '(before)'
            for start, n in ph:
                if line >= start:
                    name = n
            for k, x in v.items():
                agg[name][k] += x
        for n, v in sorted(agg.items(), key=lambda x: -x[1]['all']):
            top = sorted(((x, k) for k, x in v.items() if k != 'all'), reverse=True)[:4]
            print(f'{n:22s} {100 * v["all"] / tot:5.1f}%   ' + ' '.join(f'{k[6:]}:{100 * x / v["all"]:.0f}%' for x, k in top))
    else:
        for line, v in sorted(by.items(), key=lambda x: -x[1]['all'])[:30]:
            top = sorted(((x, k) for k, x in v.items() if k != 'all'), reverse=True)[:3]
            print(f'line {line:4d} {100 * v["all"] / tot:5.1f}%   ' + ' '.join(f'{k[6:]}:{100 * x / v["all"]:.0f}%' for x, k in top))


if __name__ == '__main__':
    args = sys.argv[2:]
    fname = args[0] if args and ':' not in args[0] else ''
    main(sys.argv[1], fname, [a for a in args if ':' in a])
