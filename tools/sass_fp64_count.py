#!/usr/bin/env python3
"""FP64-pipe instructions per state of the BK1 kernel of a compiled module, counted in its SASS (the kernel is
straight-line code: one pass = one state per thread, so the static count IS the executed count)."""
import collections
import re
import subprocess
import sys


def count(lib, kernel='kx_bk1_f64ILb0'):
    txt = subprocess.run(['cuobjdump', '-sass', lib], stdout=subprocess.PIPE, text=True).stdout
    for part in re.split(r'\n\s*Function : ', txt)[1:]:
        name = part.split('\n', 1)[0].strip()
        if kernel not in name:
            continue
        ops = collections.Counter()
        for line in part.split('\n'):
            m = re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_]+)', line)
            if m:
                ops[m.group(1)] += 1
        fp64 = sum(v for k, v in ops.items() if k in ('DFMA', 'DMUL', 'DADD', 'DSETP', 'DMNMX'))
        return fp64, sum(ops.values()), ops
    raise SystemExit(f'{kernel} not found in {lib}')


if __name__ == '__main__':
    for lib in sys.argv[1:]:
        fp64, total, ops = count(lib)
        print(f'{lib}: {fp64} FP64-pipe of {total} instructions ({total * 16 // 1024} KB); MUFU {ops["MUFU"]}, LDS {ops["LDS"]}, '
              f'LDG {ops["LDG"]}, LDL {ops["LDL"]}, STL {ops["STL"]}, LDTM {ops["LDTM"]}')
