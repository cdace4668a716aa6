"""CPU tests of the product's host logic: mechanism front-end and transport fits against the golden
fixtures dumped from the reference's own Python front-end (oracle/gen_golden.py), the emitter, sharding."""
import json
import os

import numpy as np
import pytest

from kinetix_b200.core.emit_module import choose_tile, emit_module
from kinetix_b200.core.mechanism import load_mechanism, mechanism_to_dict
from kinetix_b200.core.transport_fit import fit_transport
from kinetix_b200.sharding import shard_bounds
from tests.common import mech_path

GOLDEN = os.path.join(os.path.dirname(__file__), 'golden')
ALL = ['LiDryer', 'H2_Konnov', 'H2_new_mech', 'gri30-20', 'gri30-27', 'gri30-35', 'chempolimi_edit',
       'NH3Konnov_edit', 'gri30', 'heptaneLu88', 'EtOHKonnov']
_cache = {}


def mech(name):
    if name not in _cache:
        _cache[name] = load_mechanism(mech_path(name))
    return _cache[name]


@pytest.mark.parametrize('name', ALL)
def test_ir_matches_reference_parse(name):
    """species order (inert last), molar masses, NASA-7 pieces and the full reaction table are what the
    reference's get_reaction_from_model / get_species_from_model produce (bit-for-bit floats)."""
    ours = json.loads(json.dumps(mechanism_to_dict(mech(name))))
    with open(os.path.join(GOLDEN, name + '.mech.json')) as fh:
        gold = json.load(fh)
    for key in ('n_species', 'n_active', 'n_reactions'):
        assert ours[key] == gold[key]
    assert ours['species'] == gold['species']
    for i, (a, b) in enumerate(zip(ours['reactions'], gold['reactions'])):
        assert a == b, (name, i, a, b)


@pytest.mark.parametrize('name', ['LiDryer', 'H2_Konnov', 'gri30-20', 'chempolimi_edit', 'NH3Konnov_edit', 'gri30'])
def test_transport_fits_match_reference(name):
    m = mech(name)
    fits = fit_transport(m)
    gold = np.load(os.path.join(GOLDEN, name + '.transport.npz'))
    N = m.n_species
    tri = np.array([fits.diffusivity[k][j] for k in range(N) for j in range(k)]).reshape(-1, 5)
    lnT = np.log(np.linspace(300.0, 3000.0, 40))
    V = np.vander(lnT, 5, increasing=True)
    for ours, ref in ((fits.conductivity, gold['conductivity']), (fits.viscosity, gold['viscosity']),
                      (tri, gold['diffusivity_lower'])):
        a, b = ours @ V.T, ref @ V.T
        assert np.max(np.abs(a - b) / np.abs(b)) < 1e-12


def test_species_order_moves_inert_species_last():
    m = mech('gri30')
    assert m.species_names[-1] == 'AR' and m.n_active == 52 and m.n_species == 53
    m = mech('LiDryer')
    assert m.n_species == 9 and m.n_active == 8


@pytest.mark.parametrize('name', ['LiDryer', 'chempolimi_edit', 'EtOHKonnov'])
def test_emitter_produces_a_module(name):
    """text-level checks only (no nvcc): every reaction is emitted once, the kernel entry points and
    the kxm_* host interface are present, P-log / SRI reactions are handled."""
    m = mech(name)
    src, stats = emit_module(m, None)
    for sym in ('kx_bk1_f64', 'kx_thermo', 'kxm_production_rates', 'kxm_transport', 'kxm_thermo',
                'kxm_n_species', 'kxm_species_names', 'kxm_molar_masses', 'kxm_abi_version'):
        assert sym in src
    for i in range(m.n_reactions):
        assert f'  // {i + 1}: ' in src
    assert stats['bk1_schedule']['peak_live'] <= m.n_species
    if any(r.kind == 'P-log' for r in m.reactions):
        assert 'lnPv' in src and 'Pv >' in src and 'lnPv = kx_log(Pv)' in src
    if any(r.kind == 'SRI' for r in m.reactions):
        assert 'kx_pow' in src
    # the --single-precision flavour of the same mechanism: FP32 kernel template + both storage types
    src32, _ = emit_module(m, None, single_precision=True)
    assert 'kx_bk1_f32' in src32 and 'typedef float real;' in src32 and 'launch_bk1<float>' in src32
    for i in range(m.n_reactions):
        assert f'  // {i + 1}: ' in src32


def test_bk1_scratch_slot_plan():
    """BK1 scratch slots (exp(+-g_k) of live species, third-body sums, concentrations).  Mid-size mechanisms (10-36
    live species: GRI-3.0 ...) run two 256-thread CTAs per SM at 128 registers -- four warps per scheduler -- with
    the concentrations C_k of the live species in shared-memory slots and exp(+-g_k) / third-body sums in the CTA's 256
    tensor-memory columns first; the smallest (LiDryer) keeps four 128-thread CTAs with everything in shared memory,
    heptaneLu88 three with C_k in slots; EtOHKonnov (210 slots) gets one 256-thread CTA, <= 110 shared-memory and
    <= 128 TMEM doubles per thread, live set capped.  Every TMEM value a reaction reads is loaded (tcgen05.ld) inside
    that reaction's scope, behind a wait, before its first use."""
    import re
    src, stats = emit_module(mech('gri30'), None)
    sch = stats['bk1_schedule']
    assert '__launch_bounds__(256, 2)' in src and 'kx_tm_alloc_all<256>' in src and 'kx_tm_free_all<256>' in src
    assert 0 < sch['tmem_slots'] <= 64 and sch['smem_slots'] <= 53 and sch['cold_activations'] == 53
    assert sch['smem_slots'] * 8 * 256 * 2 <= 220 * 1024 and '__syncthreads();\n  // ---- unit' not in src[:src.index('kx_bk1_f64s(')]
    # ... and the module carries the classic kernel for launches of at most one wave of it (same NASA table, one pool)
    assert 'kx_bk1_f64s(' in src and 'kx_bk1_f64s<false><<<g, 128,' in src and src.count('kx_nasa_tab[') > 2
    assert src.count('double kx_nasa_tab[') == 1 and '__launch_bounds__(128, 3)' in src
    assert 'kx_bk1_f64s' not in emit_module(mech('gri30'), None, options={'bk1_small': False})[0]
    # the classic layout on request: three 128-thread CTAs at 168 registers, every slot in shared memory
    src, stats = emit_module(mech('gri30'), None, options={'bk1_layout': 'classic'})
    assert 'kx_tm_alloc_all' not in src and '__launch_bounds__(128, 3)' in src and 'kx_bk1_f64s' not in src
    assert stats['bk1_schedule']['tmem_slots'] == 0
    for name, bounds in (('LiDryer', '(128, 4)'), ('heptaneLu88', '(128, 3)'), ('NH3Konnov_edit', '(256, 2)'),
                         ('H2_Konnov', '(256, 2)')):
        assert f'__launch_bounds__{bounds}' in emit_module(mech(name), None)[0], name

    src, stats = emit_module(mech('EtOHKonnov'), None)
    sch = stats['bk1_schedule']
    assert '__launch_bounds__(256, 1)' in src and 'kx_tm_alloc_all<512>' in src and 'kx_tm_free_all<512>' in src
    assert 0 < sch['tmem_slots'] <= 128 and sch['smem_slots'] <= 110 and sch['peak_live'] <= 60
    assert sch['smem_slots'] * 8 * 256 <= 227 * 1024
    body = src[src.index('kx_bk1_f64('):src.index('kx_tm_free_all<512>')]
    lines = body.split('\n')
    for i, ln in enumerate(lines):
        for t in re.findall(r'\btv(\d+)\b', ln):
            if f'const double tv{t} = kx_tm_pin' in ln:
                continue
            # walk back to the definition: it must come before leaving the reaction's scope, after a wait
            j = i
            while f'const double tv{t} = kx_tm_pin(tl{t}, th{t});' not in lines[j]:
                j -= 1
                assert j > 0 and lines[j].strip() != '}', f'tv{t} used outside its fetch scope (line {i})'
            assert any('kx_tm_wait_ld();' in lines[k] for k in range(max(0, j - 12), j))
    # a TMEM store is followed by a wait::st before the next TMEM load
    dirty = False
    for ln in lines:
        if 'kx_tm_st_d(' in ln:
            dirty = True
        elif 'kx_tm_wait_st();' in ln:
            dirty = False
        elif 'kx_tm_ld_d(' in ln:
            assert not dirty, 'tcgen05.ld issued after a tcgen05.st without wait::st'
    # forcing the layout with two CTAs per SM halves the columns each may allocate
    src2, _ = emit_module(mech('gri30'), None, options={'bk1_tmem': True, 'bk1_tmem_block': 128, 'bk1_tmem_ctas': 2})
    assert 'kx_tm_alloc_all<256>' in src2 and '__launch_bounds__(128, 2)' in src2
    # a shape whose shared + tensor memory cannot hold the slots falls back to the shared-memory layout
    src3, _ = emit_module(mech('gri30'), None, options={'bk1_tmem': True, 'bk1_smem_cap': 0, 'bk1_tmem_block': 512,
                                                        'bk1_tmem_ctas': 2})
    assert 'kx_tm_alloc_all' not in src3


@pytest.mark.parametrize('name', ALL)
def test_wilke_mass_factor_matrix_is_low_rank(name):
    """c_kj = (8 (1 + M_k/M_j))^-1/2 (reference mix_transport.py:311-317, 534-553) is a smooth kernel in
    ln M_k - ln M_j: the SVD factors the BK2 kernels use reproduce every entry to 1e-13 with rank <= 16."""
    from kinetix_b200.core.emit_module import wilke_low_rank
    M = np.array(mech(name).molar_masses)
    U, V, rank = wilke_low_rank(M)
    C = 1.0 / np.sqrt(8.0 * (1.0 + M[:, None] / M[None, :]))
    assert U.shape == (len(M), rank) and V.shape == (len(M), rank)
    assert np.max(np.abs(U @ V.T - C) / C) <= 1e-13
    assert rank <= 16 and (rank % 2 == 0 or rank == len(M))


def _macros(src):
    import re
    return {m.group(1): int(m.group(2)) for m in re.finditer(r'#define (KX_\w+) (-?\d+)\n', src)}


def test_bk2_kernel_plan():
    """which BK2 kernel a mechanism gets, and that the plan respects the hardware limits it was made for:
    227 KB shared memory per CTA, 512 tensor-memory columns per lane, 16-byte bulk-copy granularity."""
    import re
    from kinetix_b200.core.transport_fit import fit_transport
    # GRI-3.0: persistent tensor-memory kernel, two states per thread
    m = mech('gri30')
    src, _ = emit_module(m, fit_transport(m))
    d = _macros(src)
    assert '#include "kx_bk2_tmem.cuh"' in src and d['KX_P'] == 2 and d['KX_BK2_BLOCK'] == 256 and d['KX_WR'] == 12
    # tensor memory: ceil(warps / 4) x P states per lane, each KX_NS sums + 2 parked scalars; the top row block never
    # goes to tensor memory (descending block order), so KX_NS = KX_NP - KX_TB
    assert d['KX_NS'] == d['KX_NP'] - d['KX_TB']
    assert d['KX_L'] == 1 and -(-(d['KX_BK2_BLOCK'] // 32) // 4) * d['KX_P'] * 2 * (d['KX_NS'] + 2) <= 512
    offs = [int(x) for x in re.search(r'kx_chunk_off\[(\d+)\] = \{([^}]*)\}', src).group(2).split(',')]
    assert len(offs) == d['KX_N_CHUNKS'] + 1 and offs[0] == 0
    sizes = np.diff(offs)
    assert (sizes > 0).all() and (sizes % 2 == 0).all() and sizes.max() <= d['KX_CHUNK_MAX']
    n_tiles = (d['KX_NP'] // d['KX_TB']) * (d['KX_NP'] // d['KX_TB'] + 1) // 2
    assert d['KX_NVC'] + d['KX_NUC'] + n_tiles == d['KX_N_CHUNKS']
    smem = int(re.search(r'const int block = KX_BK2_BLOCK;\n  const size_t smem = (\d+);', src).group(1))
    assert smem <= 227 * 1024
    assert smem >= (d['KX_STAGES'] * d['KX_CHUNK_MAX'] + 53 * 512) * 8
    # small launches run the one-state-per-thread instantiation of the same kernel
    assert 'kx_bk2<S, 1, KX_L><<<' in src and 'kx_bk2<S, KX_P, KX_L><<<' in src
    # the 129-species mechanism: X_k leaves room for 128 states per SM; 256 threads work as two halves sharing them
    # (KX_L = 2), each half takes half of a tile's columns, and the exchange areas fit the 512 tensor-memory columns
    m = mech('EtOHKonnov')
    d = _macros(emit_module(m, fit_transport(m))[0])
    assert d['KX_P'] == 1 and d['KX_L'] == 2 and d['KX_BK2_BLOCK'] == 256 and d['KX_TB'] % 2 == 0
    assert d['KX_COL_UNROLL'] == d['KX_TB'] // 2
    nx = max(3 * d['KX_WR'] + 2, d['KX_TB'])
    assert 2 * (d['KX_NS'] + 2 + 2 * nx) <= 512
    assert (129 * 128 + d['KX_STAGES'] * d['KX_CHUNK_MAX']) * 8 <= 227 * 1024
    # a mechanism that holds 256 one-state threads keeps one warp per state group
    m = mech('heptaneLu88')
    d = _macros(emit_module(m, fit_transport(m))[0])
    assert d['KX_P'] == 1 and d['KX_L'] == 1 and d['KX_BK2_BLOCK'] == 256
    # small mechanisms keep the one-state-per-thread kernel with the dense Wilke matrix
    m = mech('LiDryer')
    src, _ = emit_module(m, fit_transport(m))
    assert '#include "kx_bk2.cuh"' in src and 'KX_WR' not in _macros(src)
    # the --single-precision module keeps the one-state-per-thread kernel too (tensor-memory kernel is FP64)
    m = mech('gri30')
    src, _ = emit_module(m, fit_transport(m), single_precision=True)
    assert '#include "kx_bk2.cuh"' in src and _macros(src)['KX_WR'] == 12


def test_tile_choice():
    assert choose_tile(53) == (9, 54)
    assert choose_tile(9) == (9, 9)
    tb, NP = choose_tile(129)
    assert NP % tb == 0 and NP >= 129


def test_shard_bounds_cover_the_batch():
    for n in (0, 1, 7, 16, 1000003):
        for w in (1, 2, 3, 8):
            b = shard_bounds(n, w)
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            sizes = [e - s for s, e in b]
            assert max(sizes) - min(sizes) <= 1
