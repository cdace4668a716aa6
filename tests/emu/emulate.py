"""TEST INFRASTRUCTURE: run the GENERATED BK1 kernel on the CPU, thread by thread.

The BK1 kernel is text produced by kinetix_b200/core/emit_bk1.py; its correctness lives in the emitter's decisions
(reaction order, activation / retirement of species, slot recycling in shared and tensor memory, live-range
splitting through the output rows, folded constants).  This module takes the BK1 part of the module source exactly
as emit_module() wrote it, puts host stand-ins under it (tests/emu/cuda_emu.h for the CUDA keywords and intrinsics,
tests/emu/kx_tm_emu.h for tensor memory, and a copy of csrc/kx_math.cuh whose few inline-PTX statements are replaced
by their C meaning -- the exp / log / reciprocal code itself is the product's), compiles it with g++ and runs one
"thread" per state.  Only tests use it; the product has no CPU path (tests/test_abi.py::test_no_cpu_fallback).
"""
import ctypes
import hashlib
import os
import re
import subprocess
import tempfile

import numpy as np

from kinetix_b200.core import constants as const
from kinetix_b200.core.emit_module import emit_module
from kinetix_b200.core.mechanism import load_mechanism

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, 'kinetix_b200', 'csrc')

# inline-PTX statement (identified by a substring) -> its meaning in C
_ASM_MEANING = [
    ('rcp.approx.ftz.f64', 'x = kx_emu_rcp_seed(a);'),
    ('mad.lo.s64 a, %2, %3, %1;\\n\\tld.global.nc', 'v = base[(long long)(K / 8) * offset];'),
    ('mad.lo.s64 a, %1, %2, %0;\\n\\tst.global', 'base[(long long)(K / 8) * offset] = v;'),
    ('add.f64 t, t, %3', 'base[(long long)(K / 8) * offset] += v;'),
    ('ld.global.nc.L1::no_allocate.f64', 'v = *p;'),
    ('ld.global.nc.L1::evict_last.f64', 'v = *p;'),
    ('st.global.L1::no_allocate.f64', '*p = v;'),
    ('ex2.approx.ftz.f32', 'y = exp2f(x);'),
    ('lg2.approx.ftz.f32', 'y = log2f(x);'),
    ('rcp.approx.ftz.f32', 'y = 1.0f / x;'),
    ('ld.global.nc.L1::no_allocate.f32', 'v = *p;'),
    ('st.global.L1::no_allocate.f32', '*p = v;'),
    ('cp.async.ca.shared.global', 'abort();'),
    ('cp.async.wait_group', 'abort();'),
    ('ld.shared.f64', 'abort(); v = 0;'),
]


def _replace_asm(text):
    """replace every `asm [volatile] ( ... );` statement by its C meaning"""
    out, i = [], 0
    pat = re.compile(r'\basm\s*(?:volatile\s*)?\(')
    while True:
        m = pat.search(text, i)
        if not m:
            out.append(text[i:])
            break
        # skip matches inside // comments
        line_start = text.rfind('\n', 0, m.start()) + 1
        if '//' in text[line_start:m.start()]:
            out.append(text[i:m.end()])
            i = m.end()
            continue
        j, depth, in_str = m.end(), 1, False
        while depth:
            c = text[j]
            if in_str:
                if c == '\\':
                    j += 1
                elif c == '"':
                    in_str = False
            elif c == '"':
                in_str = True
            elif c == '(':
                depth += 1
            elif c == ')':
                depth -= 1
            j += 1
        assert text[j] == ';', text[m.start():j + 20]
        stmt = text[m.start():j + 1]
        for key, meaning in _ASM_MEANING:
            if key in stmt:
                out.append(text[i:m.start()] + meaning)
                break
        else:
            raise RuntimeError('emulate.py: no C meaning for ' + stmt)
        i = j + 1
    return ''.join(out)


def _math_header():
    text = open(os.path.join(CSRC, 'kx_math.cuh')).read()
    text = text.replace('#include <cuda_runtime.h>', '#include "cuda_emu.h"')
    # kx_ld_row<K> etc. take K*8 as an immediate: the template parameter is the row index, the C meaning wants K
    text = _replace_asm(text).replace('(long long)(K / 8) * offset', '(long long)K * offset')
    return text


def bk1_source(mech_name, options=None, single_precision=False):
    """BK1 part of the module text (constants, NASA table, kernel) + launch shape, as emit_module plans it"""
    path = mech_name if os.path.exists(mech_name) else os.path.join(ROOT, 'kinetix_b200', 'mechanisms', mech_name + '.yaml')
    mech = load_mechanism(path)
    opts = dict(options or {})
    opts.setdefault('bk1_small', False)      # the small-launch instantiation IS the classic layout (emulated on request)
    src, stats = emit_module(mech, None, opts, single_precision=single_precision)
    bk1 = src[:src.index('kx_rcpM[')]
    bk1 = bk1[:bk1.rindex('\n')]                    # drop the started table line
    m = re.search(r'__launch_bounds__\((\d+), (\d+)\)', bk1)
    return mech, bk1, int(m.group(1)), stats


class BK1Emulator:
    def __init__(self, mech_name, options=None, single_precision=False):
        """single_precision: the FP32-math kernel of the --single-precision module, run on FP64 buffers ("fpmix");
        ex2 / lg2 / rcp.approx become exp2f / log2f / 1/x, i.e. the hardware approximations' 2^-22 errors are NOT
        modelled -- this checks the emitter's log2-space algebra and its immediates, not the MUFU accuracy."""
        self.sp = bool(single_precision)
        self.mech, src, self.block, self.stats = bk1_source(mech_name, options, self.sp)
        slots = self.stats['bk1_schedule']['smem_slots']
        src = src.replace('#include <cuda_runtime.h>', '#include "cuda_emu.h"').replace('#include <math_constants.h>', '')
        src = src.replace('#include "kx_math.cuh"', '#include "kx_math_emu.h"').replace('#include "kx_tm.cuh"', '#include "kx_tm_emu.h"')
        if 'kx_tm_emu.h' not in src:
            src = src.replace('#include "kx_math_emu.h"', '#include "kx_math_emu.h"\n#include "kx_tm_emu.h"')
        assert src.count('extern __shared__ double kx_sm[];') == (0 if self.sp else 1)
        src = src.replace('extern __shared__ double kx_sm[];', 'double* const kx_sm = emu_smem;')
        src = src.replace('__shared__ unsigned kx_tm_slot;', 'static unsigned kx_tm_slot;')
        src = src.replace('#include "kx_math_emu.h"', f'static double emu_smem[{max(slots, 1)} * {self.block}];\n#include "kx_math_emu.h"', 1)
        has_pool = 'KxParamPool' in src
        if self.sp:
            call = ('if (pfield) kx_bk1_f32<double, true>(n, offsetT, offset, (float)pressure_R, (float)P, (float)log(P), '
                    'state, rates, Tref, pfield);\n    else kx_bk1_f32<double, false>(n, offsetT, offset, (float)pressure_R, '
                    '(float)P, (float)log(P), state, rates, Tref, nullptr);')
        else:
            pool = ', kx_param_pool' if has_pool else ''
            call = (f'if (pfield) kx_bk1_f64<true>(n, offsetT, offset, pressure_R, P, log(P), state, rates, Tref, pfield{pool});\n'
                    f'    else kx_bk1_f64<false>(n, offsetT, offset, pressure_R, P, log(P), state, rates, Tref, nullptr{pool});')
        harness = f'''
extern "C" unsigned emu_tm_columns() {{ return emu_tm_max_col; }}
extern "C" int emu_bk1(long long n, long long offsetT, long long offset, double pressure_R, double P,
                       const double* state, double* rates, double Tref, const double* pfield)
{{
  const unsigned block = {self.block};
  const long long n_threads = (n + block - 1) / block * block;      // tail threads run too (they store nothing)
  uint64_t poison = 0x7ff8dead00000000ull;
  double nanv;
  memcpy(&nanv, &poison, 8);
  blockDim.x = block;
  gridDim.x = (unsigned)(n_threads / block);
  for (long long t = 0; t < n_threads; t++) {{
    threadIdx.x = (unsigned)(t % block);
    blockIdx.x = (unsigned)(t / block);
    // a slot read before this thread wrote it must not find a plausible value
    for (int s = 0; s < {max(slots, 1)}; s++) emu_smem[s * block + threadIdx.x] = nanv;
    const unsigned lane = (((threadIdx.x >> 5) & 3u) << 5) + (threadIdx.x & 31u);
    for (int c = 0; c < 512; c++) emu_tmem[lane][c] = (c & 1) ? 0x7ff8deadu : 0u;
    {call}
  }}
  return 0;
}}
'''
        full = src + harness
        math = _math_header()
        key = hashlib.sha256((full + math + open(os.path.join(HERE, 'cuda_emu.h')).read() +
                              open(os.path.join(HERE, 'kx_tm_emu.h')).read()).encode()).hexdigest()[:16]
        work = os.path.join(tempfile.gettempdir(), f'kx_emu_{os.getuid()}')
        os.makedirs(work, exist_ok=True)
        lib = os.path.join(work, f'emu_{os.path.basename(mech_name)}{"_sp" if self.sp else ""}_{key}.so')
        if not os.path.exists(lib):
            d = tempfile.mkdtemp(dir=work)
            with open(os.path.join(d, 'kx_math_emu.h'), 'w') as fh:
                fh.write(math)
            with open(os.path.join(d, 'bk1_emu.cpp'), 'w') as fh:
                fh.write(full)
            cmd = ['g++', '-std=c++17', '-O0', '-ffp-contract=off', '-shared', '-fPIC', '-w', '-I', d, '-I', HERE,
                   '-o', lib + '.tmp', os.path.join(d, 'bk1_emu.cpp')]
            r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
            if r.returncode != 0:
                raise RuntimeError('g++ failed on the emulated kernel:\n' + r.stdout[-4000:])
            os.replace(lib + '.tmp', lib)
        self.lib = ctypes.CDLL(lib)
        self.lib.emu_bk1.argtypes = [ctypes.c_longlong] * 3 + [ctypes.c_double] * 2 + [ctypes.c_void_p] * 2 + \
                                    [ctypes.c_double, ctypes.c_void_p]

    def production_rates(self, state, pressure, p_field=None, Tref=1.0):
        """state: (N+1, S) species-major like the device slab; returns rates of the same shape.  `pressure` in Pa
        (or, with p_field, the reference pressure the per-state factors multiply)."""
        state = np.ascontiguousarray(state, dtype=np.float64)
        S = state.shape[1]
        rates = np.full_like(state, np.nan)
        pf = None if p_field is None else np.ascontiguousarray(p_field, dtype=np.float64)
        self.lib.emu_bk1(S, S, S, pressure / const.R_GAS, pressure, state.ctypes.data, rates.ctypes.data, Tref,
                         None if pf is None else pf.ctypes.data)
        return rates

    def tmem_columns_touched(self):
        return int(self.lib.emu_tm_columns())
