#!/usr/bin/env python3
"""Per-phase FP64-pipe utilisation of a kernel from an .ncu-rep: for every source-line range (start:name pairs)
FP64 warp instructions executed x 2 issue cycles / (share of warp-stall samples x active SMSP cycles).
Assumes the sampled share of a phase equals its share of the run time (warps are resident throughout)."""
import csv
import io
import subprocess
import sys
from collections import defaultdict

FP64 = ('DFMA', 'DMUL', 'DADD', 'DSETP')


def main(path, phases):
    raw = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    ci = rows[0].index('smsp__cycles_active.sum')
    cycles = sum(float(r[ci]) for r in rows[2:] if len(r) > ci and r[ci])   # the source page sums all launches in the report
    src = subprocess.run(['ncu', '-i', path, '--page', 'source', '--csv', '--print-source', 'cuda,sass'],
                         stdout=subprocess.PIPE, text=True).stdout.splitlines()
    ph = sorted((int(a.split(':')[0]), a.split(':')[1]) for a in phases)
    col = None
    line = None
    fname = ''
    samp = defaultdict(float)
    f64 = defaultdict(float)
    lds = defaultdict(float)
    allinst = defaultdict(float)
    for r in csv.reader(src):
        if len(r) == 2 and r[0] == 'File Path':
            fname = r[1]
            continue
        if r and r[0] == 'Line No':
            col = {n: i for i, n in enumerate(r)}
            continue
        if col is None or len(r) < len(col):
            continue
        if r[0].strip().isdigit():
            line = int(r[0])
            continue
        name = 'helpers'
        if 'bk2' in fname:
            name = '(before)'
            for start, n in ph:
                if line >= start:
                    name = n
        try:
            s = float(r[col['# Samples']])
            n = float(r[col['Instructions Executed']])
        except ValueError:
            continue
        op = r[3].split()
        op = (op[1] if op and op[0].startswith('@') else (op[0] if op else '?')).split('.')[0]
        samp[name] += s
        allinst[name] += n
        if op in FP64:
            f64[name] += n
        if op in ('LDS', 'STS', 'LDG', 'STG'):
            lds[name] += n
    tot = sum(samp.values())
    # normalise to the measured pipe utilisation (reports with several launches aggregate differently per page)
    pi = rows[0].index('sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active')
    measured = [float(r[pi]) for r in rows[2:] if len(r) > pi and r[pi]]
    cycles *= (100 * 2 * sum(f64.values()) / cycles) / (sum(measured) / len(measured))
    print(f'{"phase":14s} {"time%":>6s} {"fp64 util%":>10s} {"fp64 instr%":>11s} {"instr/fp64":>10s} {"ld/st per fp64":>14s}')
    for n in sorted(samp, key=lambda x: -samp[x]):
        t = samp[n] / tot * cycles
        print(f'{n:14s} {100 * samp[n] / tot:6.1f} {100 * 2 * f64[n] / t if t else 0:10.1f} {100 * f64[n] / sum(f64.values()):11.1f} '
              f'{allinst[n] / f64[n] if f64[n] else 0:10.2f} {lds[n] / f64[n] if f64[n] else 0:14.3f}')
    print(f'total fp64 warp instr {sum(f64.values()):.3e}, all {sum(allinst.values()):.3e}, overall util {100 * 2 * sum(f64.values()) / cycles:.1f}%')


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2:])
