#!/usr/bin/env python3
"""TEST INFRASTRUCTURE ONLY -- build oracle/_ref: the reference's own CPU implementation of the path.

For each requested (mechanism, variant) this
  1. runs the UNMODIFIED reference generator from /root/reference (python, via the ruamel shim in
     oracle/shims) into oracle/_ref/gen/<mech>.<variant>/     -- reference kinetix/__main__.py
  2. compiles oracle/ref_harness.cpp with those generated files force-included into
     oracle/_ref/libref_<mech>.<variant>.so                   -- g++ only, no OCCA/MPI/cmake

variants
  parity   --unroll-loops, g++ -O2, no fast-math     : the PARITY oracle
  serial   rolled (SERIAL default), reference flags  : -O3 -march=native -mtune=native -ffast-math
                                                       (benchmark/src/kinetix.cpp:567-578) -> CPU baseline timing
  fpmix    --single-precision --unroll-loops, -O2    : what `kinetix_bk --single-precision` really runs
                                                       (double storage, float math)
  rcpdiff  --unroll-loops --fit-rcpdiffcoeffs, -O2   : reciprocal diffusion-coefficient fits (changes BK2 results)

oracle/_ref/ is git-ignored (never committed) but travels to the GPU box.  /root/reference is only
needed for step 1; if it is absent and the .so already exists the step is skipped.

`-march=native` would bake this container's ISA into a library that must run on the GPU host, so the
serial variant is compiled at run time by ensure_serial_native() instead when the host has gcc; the
prebuilt one uses -march=x86-64-v3.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = '/root/reference'
OUT = os.path.join(HERE, '_ref')

# every shipped mechanism has its parity oracle (the unrolled -O2 build of EtOHKonnov takes ~5 min of g++, once)
DEFAULT = [('gri30', 'parity'), ('gri30', 'serial'), ('gri30', 'fpmix'),
           ('LiDryer', 'parity'), ('LiDryer', 'serial'),
           ('NH3Konnov_edit', 'parity'), ('chempolimi_edit', 'parity'), ('LiDryer', 'rcpdiff'),
           ('H2_Konnov', 'parity'), ('H2_new_mech', 'parity'), ('gri30-20', 'parity'), ('gri30-27', 'parity'),
           ('gri30-35', 'parity'), ('heptaneLu88', 'parity'), ('EtOHKonnov', 'parity'), ('EtOHKonnov', 'serial')]

COMMON_DEFS = [
    '-D__KINETIX_DEVICE__=', '-D__KINETIX_CONST__=const', "-D__KINETIX_INLINE__=static inline",
    '-D__KINETIX_MAX=fmax', '-include', 'cmath', '-include', 'cstdio',
]
FP64_DEFS = ['-Ddfloat=double', '-Dcfloat=double', '-DCFLOAT_MAX=1e300', '-DCFLOAT_MIN=1e-300',
             '-D__KINETIX_EXP__=exp', '-D__KINETIX_LOG10__=log10', '-D__KINETIX_LOG__=log',
             '-D__KINETIX_POW__=pow', '-D__KINETIX_MIN_CFLOAT=fmin']
FPMIX_DEFS = ['-Ddfloat=double', '-Dcfloat=float', '-DCFLOAT_MAX=1e37f', '-DCFLOAT_MIN=1e-37f',
              '-D__KINETIX_EXP__=expf', '-D__KINETIX_LOG10__=log10f', '-D__KINETIX_LOG__=logf',
              '-D__KINETIX_POW__=powf', '-D__KINETIX_MIN_CFLOAT=fminf']


def gen_dir(mech, variant):
    return os.path.join(OUT, 'gen', f'{mech}.{variant}')


def lib_path(mech, variant):
    return os.path.join(OUT, f'libref_{mech}.{variant}.so')


def run_generator(mech, variant):
    """Step 1: reference generator, unmodified."""
    out = gen_dir(mech, variant)
    os.makedirs(out, exist_ok=True)
    cmd = [sys.executable, os.path.join(REF, 'kinetix', '__main__.py'),
           '--mechanism', os.path.join(REF, 'kinetix', 'mechanisms', mech + '.yaml'),
           '--output', out, '--align-width', '64', '--target', 'c++17']
    if variant in ('parity', 'fpmix', 'unroll_fast'):
        cmd.append('--unroll-loops')
    if variant == 'fpmix':
        cmd.append('--single-precision')
    if variant == 'rcpdiff':
        cmd += ['--unroll-loops', '--fit-rcpdiffcoeffs']
    env = dict(os.environ, PYTHONPATH=os.path.join(HERE, 'shims'))
    subprocess.run(cmd, check=True, env=env, stdout=subprocess.DEVNULL)


def compile_lib(mech, variant, march='x86-64-v3', out_path=None):
    """Step 2: g++ on the harness + generated files."""
    g = gen_dir(mech, variant)
    pre = 'f' if variant == 'fpmix' else ''
    inc = ['-include', os.path.join(g, 'mech.h')]
    for f in ('rates', 'enthalpy_RT', 'heat_capacity_R', 'conductivity', 'viscosity', 'diffusivity'):
        inc += ['-include', os.path.join(g, pre + f + '.cpp')]
    if variant in ('serial', 'unroll_fast'):
        opt = ['-O3', f'-march={march}', f'-mtune={"native" if march == "native" else "generic"}', '-ffast-math']
    else:
        opt = ['-O2']
    defs = COMMON_DEFS + (FPMIX_DEFS if variant == 'fpmix' else FP64_DEFS)
    out_path = out_path or lib_path(mech, variant)
    cmd = ['g++', '-std=c++17', '-shared', '-fPIC', '-w'] + opt + defs + inc + \
          [os.path.join(HERE, 'ref_harness.cpp'), '-o', out_path]
    subprocess.run(cmd, check=True)
    return out_path


def ensure_serial_native(mech, variant='serial'):
    """Re-compile the timing variant with the reference's -march=native on the machine it will be timed
    on (the GPU host).  Falls back to the prebuilt x86-64-v3 library if g++ is unavailable."""
    native = os.path.join(OUT, f'libref_{mech}.{variant}.native.so')
    if os.path.exists(native):
        return native
    try:
        return compile_lib(mech, variant, march='native', out_path=native)
    except Exception:
        return lib_path(mech, variant)


def build(targets=DEFAULT, force=False):
    os.makedirs(OUT, exist_ok=True)
    built = []
    for mech, variant in targets:
        lib = lib_path(mech, variant)
        have_ref = os.path.isdir(REF)
        if os.path.exists(lib) and not force:
            built.append(lib)
            continue
        if not have_ref and not os.path.isdir(gen_dir(mech, variant)):
            print(f'[build_ref] {mech}.{variant}: no /root/reference and nothing prebuilt -- skipped')
            continue
        if have_ref:
            run_generator(mech, variant)
        compile_lib(mech, variant)
        built.append(lib)
        print('[build_ref] built', os.path.relpath(lib, os.path.dirname(HERE)))
    return built


if __name__ == '__main__':
    args = [a for a in sys.argv[1:] if a != '--force']
    tg = [tuple(a.split('.', 1)) if '.' in a else (a, 'parity') for a in args] or DEFAULT
    build(tg, force='--force' in sys.argv)
