#!/usr/bin/env python3
"""Kernel share of device time from an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import csv
import sys
from collections import defaultdict


def main(path, note=''):
    rows = [r for r in csv.reader(open(path, errors='replace')) if r]
    i0 = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
    hdr = rows[i0]
    kn, mv, mu = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
    tot, cnt = defaultdict(float), defaultdict(int)
    for r in rows[i0 + 1:]:
        if len(r) <= mv:
            continue
        try:
            v = float(r[mv].replace(',', ''))
        except ValueError:
            continue
        scale = {'ns': 1e-6, 'nsecond': 1e-6, 'us': 1e-3, 'usecond': 1e-3, 'ms': 1.0, 'msecond': 1.0, 's': 1e3, 'second': 1e3}.get(r[mu].strip(), 1e-6)
        tot[r[kn]] += v * scale
        cnt[r[kn]] += 1
    total = sum(tot.values())
    print(f'kernel share of device time ({note}; ms, cold-cache serialised launches under ncu)')
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:10]:
        print(f'  {v:9.3f} ms  {100 * v / total:5.1f}%  x{cnt[k]:<4d} {k[:70]}')


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else '')
