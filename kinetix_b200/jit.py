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


def _input_name(path, mechanism_path=None):
    """location-independent name of an input file: package files relative to the package, the mechanism by its
    base name -- so that a cache built in one checkout is still fresh when the tree is copied elsewhere (gpurun
    snapshots), while /a/mech.yaml and /b/mech.yaml with different contents hash differently"""
    if mechanism_path is not None and os.path.abspath(path) == os.path.abspath(mechanism_path):
        return '@mechanism/' + os.path.basename(path)
    ap = os.path.abspath(path)
    root = os.path.dirname(PKG)
    return os.path.relpath(ap, root) if ap.startswith(root + os.sep) else os.path.basename(ap)


def _hash_files(paths, extra='', mechanism_path=None):
    h = hashlib.sha256(extra.encode())
    for name, p in sorted((_input_name(p, mechanism_path), p) for p in paths):
        with open(p, 'rb') as fh:
            h.update(name.encode())
            h.update(fh.read())
    return h.hexdigest()


def _write_inputs_stamp(out, paths, mechanism_path, pinned):
    """`.inputs`: what the C host (csrc/kx_host.cpp, prepare_module) re-checks before it trusts a cached module --
    an FNV-1a-64 over (name, bytes) of every input file, then the names; `pinned` marks modules built with options
    the C path cannot reproduce (explicit emit options): a stale pinned module is an error there, not a rebuild."""
    names = sorted((_input_name(p, mechanism_path), p) for p in paths)
    lib = _fnv_library()
    h = 0xcbf29ce484222325
    for name, p in names:
        with open(p, 'rb') as fh:
            data = name.encode() + fh.read()
        h = lib(h, data)
    with open(os.path.join(out, '.inputs'), 'w') as fh:
        fh.write(f'fnv {h:016x}\npinned {1 if pinned else 0}\n' + ''.join(n + '\n' for n, _ in names))


def _fnv_library():
    """FNV-1a-64 step over a byte string: the host library exports it (kx_fnv1a64) so that Python and C agree by
    construction; pure-Python fallback when the library is not built yet"""
    try:
        import ctypes
        L = ctypes.CDLL(HOST_LIB)
        f = L.kx_fnv1a64
        f.restype = ctypes.c_uint64
        f.argtypes = [ctypes.c_uint64, ctypes.c_char_p, ctypes.c_size_t]
        return lambda h, data: f(h, data, len(data))
    except Exception:
        def slow(h, data):
            for b in data:
                h = ((h ^ b) * 0x100000001b3) & 0xffffffffffffffff
            return h
        return slow


class _CacheLock:
    """exclusive advisory lock per module directory: several ranks calling kx_init / ensure_module on an uncached
    module serialise here, the first one builds, the others find it fresh"""

    def __init__(self, out):
        self.path = os.path.abspath(out).rstrip(os.sep) + '.lock'

    def __enter__(self):
        import fcntl
        os.makedirs(os.path.dirname(self.path), exist_ok=True)
        self.fh = open(self.path, 'w')
        fcntl.flock(self.fh, fcntl.LOCK_EX)
        return self

    def __exit__(self, *exc):
        import fcntl
        fcntl.flock(self.fh, fcntl.LOCK_UN)
        self.fh.close()


ROUTINE_ONLY = ('emit_routines.py', 'kx_routine_kernels.cu')     # inputs of ensure_routines, not of the kernel modules


def _emitter_sources(routines=False):
    core = os.path.join(PKG, 'core')
    out = [os.path.join(core, f) for f in os.listdir(core) if f.endswith('.py')]
    out += [os.path.join(core, 'data', f) for f in os.listdir(os.path.join(core, 'data'))]
    out += [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith('.cuh') or f.endswith('.cu')]
    return [p for p in out if routines or os.path.basename(p) not in ROUTINE_ONLY]


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
    inputs = _emitter_sources() + [mechanism_path]
    digest = _hash_files(inputs, json.dumps(opts, sort_keys=True), mechanism_path)
    stamp = os.path.join(out, '.hash')

    def is_fresh():
        return (os.path.exists(stamp) and open(stamp).read() == digest and not force
                and (os.path.exists(lib) or not compile_module))

    if is_fresh():
        if not os.path.exists(os.path.join(out, '.inputs')):
            _write_inputs_stamp(out, inputs, mechanism_path, pinned=bool(emit_options))
        if compile_module and not os.path.exists(os.path.join(out, 'counts.json')):
            try:
                from . import sass
                sass.write_counts(out)
            except Exception:
                pass
        return out
    with _CacheLock(out):
        if is_fresh():                      # another process built it while we waited for the lock
            return out
        os.makedirs(out, exist_ok=True)
        for f in (stamp, os.path.join(out, '.inputs')):      # never leave a stamp that vouches for a half-written module
            if os.path.exists(f):
                os.remove(f)
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
            tmp = lib + f'.tmp{os.getpid()}'
            cmd = [NVCC] + ARCH_FLAGS + ['-O3', '-lineinfo', '-std=c++17', '-shared', '-Xcompiler', '-fPIC',
                                         '-cudart', 'shared', '-I', CSRC, '-Xptxas', '-v', '-o', tmp, cu,
                                         '-Xlinker', f'-rpath={os.path.join(CUDA_HOME, "lib64")}'] + list(extra_nvcc)
            log = _run(cmd, f'nvcc ({mech.name})')
            os.replace(tmp, lib)            # readers see the old module or the complete new one, never a partial file
            with open(os.path.join(out, 'ptxas.log'), 'w') as fh:
                fh.write(log)
            if verbose:
                sys.stderr.write(log)
            try:                                # static FP64 census of the kernels just built (bench.py's roofline)
                from . import sass
                sass.write_counts(out)
            except Exception as e:              # cuobjdump missing: the bench then reports the count as unknown
                sys.stderr.write(f'[kinetix_b200] no SASS census for {out}: {e}\n')
        _write_inputs_stamp(out, inputs, mechanism_path, pinned=bool(emit_options))
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
    digest = _hash_files(_emitter_sources(routines=True) + [mechanism_path], json.dumps(opts, sort_keys=True), mechanism_path)
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
