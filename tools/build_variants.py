#!/usr/bin/env python3
"""Development aid: build several emitter variants of one mechanism into build/variants/<name>/ (each a
cache dir usable with `tools/quick_time.py --cache build/variants/<name>`)."""
import json
import os
import sys
from concurrent.futures import ProcessPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def one(args):
    name, mech, opts = args
    from kinetix_b200 import jit
    opts = dict(opts)
    sp = bool(opts.pop('_sp', False))          # "_sp": true builds the --single-precision module (<mech>-sp)
    out = os.path.join(ROOT, 'build', 'variants', name, mech + ('-sp' if sp else ''))
    jit.ensure_module(os.path.join(ROOT, 'kinetix_b200', 'mechanisms', mech + '.yaml'), out, emit_options=opts,
                      single_precision=sp)
    log = open(os.path.join(out, 'ptxas.log')).read()
    return name, [l.strip() for l in log.splitlines() if 'registers' in l or 'spill' in l]


if __name__ == '__main__':
    mech = sys.argv[1]
    variants = json.loads(sys.argv[2])          # {"name": {emit options}, ...}
    with ProcessPoolExecutor(max_workers=8) as ex:
        for name, info in ex.map(one, [(n, mech, o) for n, o in variants.items()]):
            print(name, info)
