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
