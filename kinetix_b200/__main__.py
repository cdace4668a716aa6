"""Generator command line: `python3 -m kinetix_b200 --mechanism gri30.yaml --output DIR [--compile]`.

Keeps the reference generator's flags (reference kinetix/utils/general_utils.py:20-87) so that the
host library and scripts can drive it the same way (benchmark/src/kinetix.cpp:655-675), and adds
`--target sm_100a` (the only target), `--compile` (run nvcc, produce libkx_mech.so) and `--block-size`.
Flags that select between the reference's alternative CPU/GPU code shapes (--unroll-loops,
--loop-gibbsexp, --group-rxnunroll, --group-vis, --nonsymDij, --align-width) are accepted for
compatibility; the sm_100a emitter always produces its one specialised form.  `--fit-rcpdiffcoeffs`
changes the fitted quantity exactly as in the reference (mix_transport.py:198-206).
"""
import argparse
import sys

from . import jit


def main(argv=None):
    p = argparse.ArgumentParser(prog='kinetix_b200', description='Generate B200 (sm_100a) kernels for the '
                                'production-rate (BK1), transport (BK2) and thermo routines of a mechanism')
    p.add_argument('--mechanism', required=True, help='Path to yaml mechanism file.')
    p.add_argument('--output', required=True, help='Output directory.')
    p.add_argument('--single-precision', action='store_true')
    p.add_argument('--header-only', action='store_true', help='Only write mech.json (no kernels).')
    p.add_argument('--unroll-loops', action='store_true')
    p.add_argument('--align-width', default=64)
    p.add_argument('--target', default='sm_100a')
    p.add_argument('--loop-gibbsexp', action='store_true')
    p.add_argument('--group-rxnunroll', action='store_true')
    p.add_argument('--transport', default=True)
    p.add_argument('--group-vis', action='store_true')
    p.add_argument('--nonsymDij', action='store_true')
    p.add_argument('--fit-rcpdiffcoeffs', action='store_true')
    p.add_argument('--compile', action='store_true', help='Compile the module with nvcc (libkx_mech.so).')
    p.add_argument('--block-size', type=int, default=0)
    p.add_argument('--emit-routines', action='store_true',
                   help='Write the reference-signature device routines (mech.h, rates, enthalpy_RT, heat_capacity_R, '
                        'conductivity, viscosity, diffusivity) into <output>/routines instead of / beside the kernels.')
    p.add_argument('--routine-ext', default='cuh', choices=['cuh', 'cpp'],
                   help="File extension of the routine files; 'cpp' gives the reference's literal names.")
    p.add_argument('--force', action='store_true')
    p.add_argument('--verbose', action='store_true')
    a = p.parse_args(argv)
    if a.target not in ('sm_100a', 'CUDA'):
        sys.exit(f"Error: unsupported --target '{a.target}': kinetix_b200 only emits sm_100a CUDA")
    transport = str(a.transport).lower() not in ('0', 'false', 'no')
    if a.emit_routines:
        if a.single_precision:
            sys.exit('Error: --emit-routines is FP64 only (the FP32-math kernels have no per-routine form)')
        import os
        jit.ensure_routines(a.mechanism, os.path.join(a.output, 'routines'), fit_rcp_diff=a.fit_rcpdiffcoeffs,
                            transport=transport, ext=a.routine_ext, force=a.force, compile_kernels=a.compile,
                            block_size=a.block_size or 128, verbose=a.verbose)
        if a.header_only or not a.compile:
            return 0
    jit.ensure_module(a.mechanism, a.output, fit_rcp_diff=a.fit_rcpdiffcoeffs,
                      single_precision=a.single_precision, block_size=a.block_size,
                      transport=transport and not a.header_only, force=a.force, verbose=a.verbose,
                      compile_module=a.compile and not a.header_only)
    return 0


if __name__ == '__main__':
    sys.exit(main())
