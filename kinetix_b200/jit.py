"""Build driver: nvcc for the per-mechanism modules and the C-ABI host library (in-tree, sm_100a only).

Two artefacts, both shared objects that stay inside the package directory (they travel with the repo
snapshot to the GPU box; nothing is installed into site-packages):

  kinetix_b200/libkinetix_b200.so                 host library, csrc/kx_host.cpp      (include/kinetix_b200.h)
  kinetix_b200/_cache/<tag>/libkx_mech.so         generated kernels for one (mechanism, option set)

The cache mirrors the reference's `.cache/KINETIX/<yaml-stem>/` + `.hash` scheme
(reference benchmark/src/kinetix.cpp:319-321,677-699): a module is regenerated when the hash over
its inputs (options, mechanism file, emitter + csrc sources) changes.
"""
import hashlib
import json
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, 'csrc')
INCLUDE = os.path.join(os.path.dirname(PKG), 'include')
HOST_LIB = os.path.join(PKG, 'libkinetix_b200.so')
CUDA_HOME = os.environ.get('CUDA_HOME', '/usr/local/cuda')
NVCC = shutil.which('nvcc') or os.path.join(CUDA_HOME, 'bin', 'nvcc')
ARCH_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a']


def default_cache():
    return os.environ.get('KINETIX_B200_CACHE', os.path.join(PKG, '_cache'))


def module_tag(mechanism_path, fit_rcp_diff=False, single_precision=False, block_size=0):
    """Directory name of a module inside the cache -- must match kx_init() in csrc/kx_host.cpp."""
    tag = os.path.splitext(os.path.basename(mechanism_path))[0]
    if fit_rcp_diff:
        tag += '-rcpdiff'
    if single_precision:
        tag += '-sp'
    if block_size and block_size > 0:
        tag += f'-b{block_size}'
    return tag


def _hash_files(paths, extra=''):
    h = hashlib.sha256(extra.encode())
    for p in sorted(paths):
        with open(p, 'rb') as fh:
            h.update(p.encode())
            h.update(fh.read())
    return h.hexdigest()


def _emitter_sources():
    core = os.path.join(PKG, 'core')
    out = [os.path.join(core, f) for f in os.listdir(core) if f.endswith('.py')]
    out += [os.path.join(core, 'data', f) for f in os.listdir(os.path.join(core, 'data'))]
    out += [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith('.cuh') or f.endswith('.cu')]
    return out


def _run(cmd, what):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError(f'{what} failed ({" ".join(cmd)}):\n{r.stdout}')
    return r.stdout


def build_host_library(force=False):
    """Compile csrc/kx_host.cpp -> libkinetix_b200.so (g++ + CUDA runtime)."""
    src = os.path.join(CSRC, 'kx_host.cpp')
    hdr = os.path.join(INCLUDE, 'kinetix_b200.h')
    stamp = HOST_LIB + '.hash'
    digest = _hash_files([src, hdr])
    if not force and os.path.exists(HOST_LIB) and os.path.exists(stamp) and open(stamp).read() == digest:
        return HOST_LIB
    cmd = [NVCC, '-O2', '-std=c++17', '-shared', '-Xcompiler', '-fPIC', '-cudart', 'shared',
           '-I', INCLUDE, '-o', HOST_LIB, src, '-ldl',
           '-Xlinker', f'-rpath={os.path.join(CUDA_HOME, "lib64")}']
    _run(cmd, 'host library build')
    with open(stamp, 'w') as fh:
        fh.write(digest)
    return HOST_LIB


def ensure_module(mechanism_path, output_dir=None, fit_rcp_diff=False, single_precision=False, block_size=0,
                  transport=True, force=False, verbose=False, compile_module=True, extra_nvcc=(),
                  emit_options=None):
    """Generate (and compile) the sm_100a module for one mechanism; returns the module directory."""
    from .core.emit_module import emit_module
    from .core.mechanism import load_mechanism, mechanism_to_dict
    from .core.transport_fit import fit_transport

    mechanism_path = os.path.abspath(mechanism_path)
    out = output_dir or os.path.join(default_cache(), module_tag(mechanism_path, fit_rcp_diff, single_precision,
                                                                 block_size))
    lib = os.path.join(out, 'libkx_mech.so')
    opts = dict(fit_rcp_diff=bool(fit_rcp_diff), single_precision=bool(single_precision),
                block_size=int(block_size or 0), transport=bool(transport), extra_nvcc=list(extra_nvcc),
                emit_options=dict(emit_options or {}))
    digest = _hash_files(_emitter_sources() + [mechanism_path], json.dumps(opts, sort_keys=True))
    stamp = os.path.join(out, '.hash')
    fresh = os.path.exists(stamp) and open(stamp).read() == digest
    if fresh and not force and (os.path.exists(lib) or not compile_module):
        return out

    os.makedirs(out, exist_ok=True)
    mech = load_mechanism(mechanism_path)
    fits = fit_transport(mech, reciprocal_diffusivity=fit_rcp_diff) if transport else None
    options = {}
    if block_size:
        options.update(block_bk1=int(block_size), block_bk2=int(block_size))
    options.update(emit_options or {})
    src, stats = emit_module(mech, fits, options, single_precision=single_precision)
    cu = os.path.join(out, 'kx_mech.cu')
    with open(cu, 'w') as fh:
        fh.write(src)
    with open(os.path.join(out, 'mech.json'), 'w') as fh:
        json.dump(dict(mechanism=mechanism_to_dict(mech), stats=stats, options=opts), fh)
    if compile_module:
        cmd = [NVCC] + ARCH_FLAGS + ['-O3', '-lineinfo', '-std=c++17', '-shared', '-Xcompiler', '-fPIC',
                                     '-cudart', 'shared', '-I', CSRC, '-Xptxas', '-v', '-o', lib, cu,
                                     '-Xlinker', f'-rpath={os.path.join(CUDA_HOME, "lib64")}'] + list(extra_nvcc)
        log = _run(cmd, f'nvcc ({mech.name})')
        with open(os.path.join(out, 'ptxas.log'), 'w') as fh:
            fh.write(log)
        if verbose:
            sys.stderr.write(log)
    with open(stamp, 'w') as fh:
        fh.write(digest)
    return out


def ensure_routines(mechanism_path, output_dir=None, fit_rcp_diff=False, transport=True, ext='cuh', force=False,
                    compile_kernels=False, block_size=128, verbose=False):
    """Write the reference-signature device routines of one mechanism (core/emit_routines.py: mech.h, rates,
    enthalpy_RT, heat_capacity_R, conductivity, viscosity, diffusivity + umbrella header) into `output_dir`
    (default: <cache>/<tag>/routines); with compile_kernels also build libkx_routines.so = the reference's three
    OKL kernels restated in CUDA around those routines (csrc/kx_routine_kernels.cu).  Returns the directory."""
    from .core.emit_routines import write_routines
    from .core.mechanism import load_mechanism
    from .core.transport_fit import fit_transport

    mechanism_path = os.path.abspath(mechanism_path)
    out = output_dir or os.path.join(default_cache(), module_tag(mechanism_path, fit_rcp_diff), 'routines')
    lib = os.path.join(out, 'libkx_routines.so')
    opts = dict(fit_rcp_diff=bool(fit_rcp_diff), transport=bool(transport), ext=ext, block_size=int(block_size),
                kind='routines')
    digest = _hash_files(_emitter_sources() + [mechanism_path], json.dumps(opts, sort_keys=True))
    stamp = os.path.join(out, '.hash')
    fresh = os.path.exists(stamp) and open(stamp).read() == digest
    if fresh and not force and (os.path.exists(lib) or not compile_kernels):
        return out
    mech = load_mechanism(mechanism_path)
    fits = fit_transport(mech, reciprocal_diffusivity=fit_rcp_diff) if transport else None
    _, stats = write_routines(mech, fits, out, ext=ext)
    if compile_kernels:
        if not transport:
            raise RuntimeError('ensure_routines: the wrapper kernels need the transport routines')
        cmd = [NVCC] + ARCH_FLAGS + ['-O3', '-lineinfo', '-std=c++17', '-shared', '-Xcompiler', '-fPIC', '-cudart', 'shared',
                                     '-I', out, f'-Dp_BLOCKSIZE={int(block_size)}', '-Xptxas', '-v', '-o', lib,
                                     os.path.join(CSRC, 'kx_routine_kernels.cu'),
                                     '-Xlinker', f'-rpath={os.path.join(CUDA_HOME, "lib64")}']
        log = _run(cmd, f'nvcc routines ({mech.name})')
        with open(os.path.join(out, 'ptxas.log'), 'w') as fh:
            fh.write(log)
        if verbose:
            sys.stderr.write(log)
    with open(stamp, 'w') as fh:
        fh.write(digest)
    return out
