"""Static instruction census of a compiled mechanism module (cuobjdump -sass).

The BK1 kernel is straight-line code -- one pass = one state per thread -- so the number of FP64-pipe instructions
in its SASS IS the number it executes per state (exception: P-log mechanisms, whose warp-uniform pressure branches
are all in the text; the static count is then an upper bound).  jit.ensure_module writes the census next to the
library (`counts.json`), bench.py reads it from the directory of the module it actually loaded, so the roofline
line can never quote the work of a different build.  Kernels with loops (BK2) need a profiler for the executed
count: tools/ncu_counts.py -> profiles/counts_r02.json, keyed by the hash of the module source.
"""
import collections
import hashlib
import json
import os
import re
import shutil
import subprocess

FP64_OPS = ('DFMA', 'DMUL', 'DADD', 'DSETP', 'DMNMX')


def census(lib):
    """{kernel name fragment: {fp64, total, ops}} for every kernel in the library"""
    exe = shutil.which('cuobjdump') or '/usr/local/cuda/bin/cuobjdump'
    txt = subprocess.run([exe, '-sass', lib], stdout=subprocess.PIPE, text=True).stdout
    out = {}
    for part in re.split(r'\n\s*Function : ', txt)[1:]:
        name = part.split('\n', 1)[0].strip()
        ops = collections.Counter()
        for line in part.split('\n'):
            m = re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_]+)', line)
            if m:
                ops[m.group(1)] += 1
        out[name] = dict(fp64=sum(v for k, v in ops.items() if k in FP64_OPS), total=sum(ops.values()),
                         mufu=ops.get('MUFU', 0), ldl=ops.get('LDL', 0), stl=ops.get('STL', 0), lds=ops.get('LDS', 0),
                         ldg=ops.get('LDG', 0), ldtm=ops.get('LDTM', 0), ffma=ops.get('FFMA', 0) + ops.get('FMUL', 0) + ops.get('FADD', 0))
    return out


def source_hash(module_dir):
    with open(os.path.join(module_dir, 'kx_mech.cu'), 'rb') as fh:
        return hashlib.sha256(fh.read()).hexdigest()


def write_counts(module_dir):
    lib = os.path.join(module_dir, 'libkx_mech.so')
    c = census(lib)

    def pick(fragment):
        # several instantiations may match (BK2: full-batch and small-launch flavours): the largest is the main one
        best = None
        for name, v in c.items():
            if fragment in name and (best is None or v['total'] > best['total']):
                best = dict(v, kernel=name)
        return best
    counts = dict(source_sha256=source_hash(module_dir),
                  bk1=pick('kx_bk1_f64ILb0') or pick('kx_bk1_f32IdLb0'),
                  bk1_small=pick('kx_bk1_f64sILb0'),      # the classic layout carried for small launches: the minimal form
                  bk1_f32=pick('kx_bk1_f32IfLb0'),
                  bk2=pick('kx_bk2Id'), bk2_f32=pick('kx_bk2If'))
    with open(os.path.join(module_dir, 'counts.json'), 'w') as fh:
        json.dump(counts, fh, indent=1)
    return counts


def read_counts(module_dir):
    try:
        with open(os.path.join(module_dir, 'counts.json')) as fh:
            c = json.load(fh)
        return c if c.get('source_sha256') == source_hash(module_dir) else None
    except Exception:
        return None
