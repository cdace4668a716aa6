"""CPU tests of the C-ABI boundary: the library loads, exports every symbol include/kinetix_b200.h declares,
and fails loudly -- not silently on a CPU path -- when no CUDA device is present."""
import ctypes
import os
import re

import pytest

from kinetix_b200 import jit
from tests.common import ROOT, mech_path


@pytest.fixture(scope='module')
def lib():
    return ctypes.CDLL(jit.build_host_library())


def declared_symbols():
    text = open(os.path.join(ROOT, 'include', 'kinetix_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(kx_[a-z0-9_]+)\s*\(', text)))


def test_header_symbols_are_exported(lib):
    syms = declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), f'{s} declared in include/kinetix_b200.h but not exported'


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    import kinetix_b200.host as kinetix
    with pytest.raises(kinetix.KinetixError) as e:
        kinetix.init(mech_path('LiDryer'))
    assert 'no CUDA device' in str(e.value)
    with pytest.raises(kinetix.KinetixError):
        kinetix.productionRates(1, 1, 1, 1.0, 0, 0)


def test_generated_module_exports_interface():
    """the prebuilt GRI-3.0 module (built by __graft_entry__.build()) exports the kxm_* interface with the
    mechanism's metadata -- host-side calls only, no kernel launch."""
    d = jit.ensure_module(mech_path('LiDryer'))
    mod = ctypes.CDLL(os.path.join(d, 'libkx_mech.so'))
    from kinetix_b200.core.emit_module import ABI_VERSION
    assert mod.kxm_abi_version() == ABI_VERSION
    assert mod.kxm_n_species() == 9 and mod.kxm_n_active_species() == 8 and mod.kxm_n_reactions() == 21
    mod.kxm_species_names.restype = ctypes.c_char_p
    names = mod.kxm_species_names().decode().split()
    assert names[-1] == 'AR' or len(names) == 9
    M = (ctypes.c_double * 9)()
    mod.kxm_molar_masses(M)
    assert abs(M[names.index('H2')] - 2.016e-3) < 1e-12


def test_module_builder_hook(tmp_path):
    """kx_set_module_builder = the reference's kinetixBuildKernel_t hook (kinetix.hpp:11-13): a host-supplied builder
    replaces the built-in generator run when the module is not cached; its failures surface as errors; NULL restores
    the default.  kx_prepare only (no CUDA)."""
    import shutil
    import kinetix_b200.host as kinetix
    calls = []
    prebuilt = jit.ensure_module(mech_path('LiDryer'))

    def failing(yaml_path, options, output_dir):
        calls.append((yaml_path, options['single_precision'], output_dir))
        return 7

    def copying(yaml_path, options, output_dir):        # an application shipping ahead-of-time compiled modules
        calls.append((yaml_path, options['single_precision'], output_dir))
        os.makedirs(output_dir, exist_ok=True)
        shutil.copy(os.path.join(prebuilt, 'libkx_mech.so'), os.path.join(output_dir, 'libkx_mech.so'))
        return 0

    def lazy(yaml_path, options, output_dir):
        return 0                                        # claims success without producing the module

    try:
        kinetix.setBuildKernel(failing)
        with pytest.raises(kinetix.KinetixError) as e:
            kinetix.prepare(mech_path('LiDryer'), cache_dir=str(tmp_path))
        assert 'module builder hook returned 7' in str(e.value)
        assert calls and calls[-1][0].endswith('LiDryer.yaml') and calls[-1][2] == str(tmp_path / 'LiDryer')
        kinetix.setBuildKernel(lazy)
        with pytest.raises(kinetix.KinetixError) as e:
            kinetix.prepare(mech_path('LiDryer'), cache_dir=str(tmp_path))
        assert 'did not produce' in str(e.value)
        kinetix.setBuildKernel(copying)
        kinetix.prepare(mech_path('LiDryer'), cache_dir=str(tmp_path))
        assert os.path.exists(tmp_path / 'LiDryer' / 'libkx_mech.so')
        n = len(calls)
        kinetix.prepare(mech_path('LiDryer'), cache_dir=str(tmp_path))      # cached now: the hook is not called again
        assert len(calls) == n
    finally:
        kinetix.setBuildKernel(None)


def test_product_does_not_import_the_oracle():
    """the oracle is test infrastructure: nothing under kinetix_b200/ may reference it"""
    for dirpath, _, files in os.walk(os.path.join(ROOT, 'kinetix_b200')):
        if '_cache' in dirpath:
            continue
        for f in files:
            if f.endswith(('.py', '.cpp', '.cuh', '.h')):
                text = open(os.path.join(dirpath, f)).read()
                assert 'import oracle' not in text and 'from oracle' not in text and 'oracle/' not in text, f


def test_cached_module_is_rechecked_against_its_inputs(tmp_path):
    """kx_prepare / kx_init trust a cached module only while the mechanism file and the emitter sources it was
    generated from are unchanged (the `.inputs` stamp; the reference hashes its generator command line and
    regenerates on mismatch, kinetix.cpp:677-699): an edited mechanism, and a DIFFERENT mechanism under the same file
    name in another directory, both regenerate; an identical copy elsewhere does not.  Two processes preparing the
    same uncached module at once serialise on the cache lock and both succeed."""
    import json
    import shutil
    import subprocess
    import sys
    import kinetix_b200.host as kinetix
    cache = str(tmp_path / 'cache')
    a = tmp_path / 'a'
    b = tmp_path / 'b'
    a.mkdir()
    b.mkdir()
    shutil.copyfile(mech_path('LiDryer'), a / 'mech.yaml')
    lib = os.path.join(cache, 'mech', 'libkx_mech.so')
    # two concurrent first builds
    code = (f"import sys; sys.path.insert(0, {ROOT!r}); import kinetix_b200.host as k; "
            f"k.prepare({str(a / 'mech.yaml')!r}, cache_dir={cache!r})")
    procs = [subprocess.Popen([sys.executable, '-c', code]) for _ in range(2)]
    assert [p.wait(timeout=600) for p in procs] == [0, 0]
    assert os.path.exists(lib) and os.path.exists(os.path.join(cache, 'mech', '.inputs'))
    assert not [f for f in os.listdir(os.path.join(cache, 'mech')) if '.tmp' in f]
    t0 = os.stat(lib).st_mtime_ns
    kinetix.prepare(str(a / 'mech.yaml'), cache_dir=cache)
    assert os.stat(lib).st_mtime_ns == t0                       # fresh: nothing ran
    shutil.copyfile(a / 'mech.yaml', b / 'mech.yaml')
    kinetix.prepare(str(b / 'mech.yaml'), cache_dir=cache)
    assert os.stat(lib).st_mtime_ns == t0                       # same contents elsewhere: still fresh
    # a different mechanism under the same name
    shutil.copyfile(mech_path('H2_Konnov'), b / 'mech.yaml')
    kinetix.prepare(str(b / 'mech.yaml'), cache_dir=cache)
    t1 = os.stat(lib).st_mtime_ns
    assert t1 != t0
    meta = json.load(open(os.path.join(cache, 'mech', 'mech.json')))
    assert len(meta['mechanism']['species']) == 13              # H2_Konnov, not the 9-species LiDryer
    # an edited mechanism (one pre-exponential factor)
    text = open(a / 'mech.yaml').read()
    assert '3.547e+15' in text
    open(a / 'mech.yaml', 'w').write(text.replace('3.547e+15', '3.600e+15', 1))
    kinetix.prepare(str(a / 'mech.yaml'), cache_dir=cache)
    assert os.stat(lib).st_mtime_ns != t1
    assert '3.6e+15' in open(os.path.join(cache, 'mech', 'kx_mech.cu')).read() or \
        len(json.load(open(os.path.join(cache, 'mech', 'mech.json')))['mechanism']['species']) == 9


def test_select_device_needs_an_initialised_context(lib):
    """kx_select_device switches the calling thread between per-device contexts; a device nothing was initialised
    on is an error with a message, not a silent no-op"""
    lib.kx_last_error.restype = ctypes.c_char_p
    assert lib.kx_select_device(5) != 0
    assert b'no mechanism has been initialised on device 5' in lib.kx_last_error()
    assert lib.kx_current_device() in (-1, 0, 1, 2, 3, 4, 5, 6, 7)
