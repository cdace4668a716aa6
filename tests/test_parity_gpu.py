"""GPU parity tests: the CUDA path (through the C ABI / host mirror) against the oracle.

Bound: FP64 per-state scaled error <= 1e-10 (BASELINE.json north_star; norm of SURVEY.md 8c), plus the
reference's own Cantera known-answer files at the reference's tolerances (bk.cpp:198-199,258,148).
"""
import os

import numpy as np
import pytest

from tests.common import Oracle, R, bk1_errors, elementwise_errors, mech_path, rel_err
from oracle.port import load_cantera_ci, synthetic_states

torch = pytest.importorskip('torch')
pytestmark = pytest.mark.gpu

TOL = 1e-10
P_ATM = 101325.0
GOLDEN = os.path.join(os.path.dirname(__file__), 'golden', 'ci_data')


@pytest.fixture(scope='module')
def kinetix():
    import kinetix_b200.host as kx
    yield kx
    kx.finalize()


def _setup(kx, mech, p_ref=P_ATM, T_ref=1.0):
    kx.init(mech_path(mech))
    N = kx.nSpecies()
    kx.build(p_ref, T_ref, [1.0 / N] * N, True)
    return N


def _run_bk1(kx, st, pressure_nd):
    S = st.shape[1]
    d_state = torch.from_numpy(st).cuda()
    d_rates = torch.full_like(d_state, float('nan'))
    kx.productionRates(S, S, S, pressure_nd, d_state, d_rates)
    torch.cuda.synchronize()
    return d_rates.cpu().numpy()


def _run_bk2(kx, st, pressure_nd):
    N, S = st.shape[0] - 1, st.shape[1]
    d_state = torch.from_numpy(st).cuda()
    visc = torch.full((S,), float('nan'), dtype=torch.float64, device='cuda')
    cond = torch.full_like(visc, float('nan'))
    rhoD = torch.full((N, S), float('nan'), dtype=torch.float64, device='cuda')
    kx.mixtureAvgTransportProps(S, S, S, pressure_nd, d_state, visc, cond, rhoD)
    torch.cuda.synchronize()
    return cond.cpu().numpy(), visc.cpu().numpy(), rhoD.cpu().numpy()


@pytest.mark.parametrize('mech', ['gri30', 'LiDryer'])
def test_bk1_matches_oracle_on_random_states(kinetix, mech):
    N = _setup(kinetix, mech)
    assert kinetix.modulePath().endswith('libkx_mech.so')
    orc = Oracle(mech)
    st = synthetic_states(N, 20000, seed=1234)
    new = _run_bk1(kinetix, st, 1.0)
    ref = orc.production_rates(st, P_ATM)
    assert np.isfinite(new).all()
    rate_err, hrr_err = bk1_errors(new, ref)
    e_all, e_sig = elementwise_errors(new, ref)
    print(f'{mech} BK1 vs {orc.kind}: rates {rate_err:.3e} hrr {hrr_err:.3e}; element-wise: all elements {e_all:.3e}, '
          f'elements >= 1e-3 of their state\'s largest rate {e_sig:.3e}')
    assert orc.kind == 'reference'
    assert rate_err <= TOL and hrr_err <= TOL and e_sig <= TOL


@pytest.mark.parametrize('mech', ['gri30', 'NH3Konnov_edit', 'H2_Konnov'])
def test_bk1_small_and_wide_kernels_at_the_switch(kinetix, mech):
    """Mechanisms with 10-36 live species carry two BK1 kernels: the four-warps-per-scheduler layout (two 256-thread CTAs
    at 128 registers, concentrations in shared-memory slots, exp(+-g) in tensor memory) and the classic layout for
    launches of at most one wave of it (SMs x CTAs per SM x 128 states).  Both sides of the switch against the oracle on
    the same states, ragged sizes included; the two kernels agree with each other far inside the bound."""
    import torch
    N = _setup(kinetix, mech)
    orc = Oracle(mech)
    n_sm = torch.cuda.get_device_properties(0).multi_processor_count
    p = P_ATM * (1.0 if mech != 'NH3Konnov_edit' else 3.7)       # inside the P-log tables
    for ctas in (3, 4):                                           # the classic kernel runs 3 or 4 CTAs per SM
        wave = n_sm * ctas * 128
        st = synthetic_states(N, wave + 1, seed=77)
        ref = orc.production_rates(st, p)
        small = _run_bk1(kinetix, np.ascontiguousarray(st[:, :wave]), p / P_ATM)
        wide = _run_bk1(kinetix, st, p / P_ATM)
        for name, new, r in (('<= one wave', small, ref[:, :wave]), ('one wave + 1', wide, ref)):
            rate_err, hrr_err = bk1_errors(new, r)
            print(f'{mech} {new.shape[1]} states ({name}) BK1 vs {orc.kind}: rates {rate_err:.3e} hrr {hrr_err:.3e}')
            assert np.isfinite(new).all() and rate_err <= TOL and hrr_err <= TOL
        assert bk1_errors(small, wide[:, :wave])[0] <= 1e-12
    tiny = synthetic_states(N, 77, seed=5)
    assert bk1_errors(_run_bk1(kinetix, tiny, p / P_ATM), orc.production_rates(tiny, p))[0] <= TOL


def test_wide_bk1_kernel_pitched_rows_and_pressure_field(kinetix):
    """The four-warps-per-scheduler BK1 kernel (more than one wave of states) with what the small-size tests exercise on the
    classic kernel only: species rows at a padded pitch behind a gap (offsetT / offset), T_ref != 1, p != p_ref, a ragged
    last CTA -- nothing outside the addressed rows may be written -- and its per-state-pressure instantiation."""
    mech = 'gri30'
    kinetix.init(mech_path(mech))
    N = kinetix.nSpecies()
    p_ref, T_ref = 2.0e5, 1000.0
    kinetix.build(p_ref, T_ref, [1.0 / N] * N, True)
    orc = Oracle(mech)
    n_sm = torch.cuda.get_device_properties(0).multi_processor_count
    S = n_sm * 3 * 128 + 2 * 256 + 77             # beyond the small-launch switch, not a multiple of the CTA size
    pitch = S + 13
    st = synthetic_states(N, S, seed=4321, Tref=T_ref)
    slab = np.full((N + 2) * pitch + 7, np.nan)
    slab[0:S] = st[0]
    offsetT = 7 + pitch
    for k in range(N):
        slab[offsetT + k * pitch: offsetT + k * pitch + S] = st[k + 1]
    d_state = torch.from_numpy(slab).cuda()
    d_rates = torch.full_like(d_state, float('nan'))
    p = 3.0e5
    kinetix.productionRates(S, offsetT, pitch, p / p_ref, d_state, d_rates)
    torch.cuda.synchronize()
    out = d_rates.cpu().numpy()
    new = np.empty_like(st)
    new[0] = out[0:S]
    for k in range(N):
        new[k + 1] = out[offsetT + k * pitch: offsetT + k * pitch + S]
    ref = orc.production_rates(st, p, Tref=T_ref)
    rate_err, hrr_err = bk1_errors(new, ref)
    print(f'{mech} wide kernel, pitched rows: rates {rate_err:.3e} hrr {hrr_err:.3e}')
    assert np.isfinite(new).all() and rate_err <= TOL and hrr_err <= TOL
    mask = np.ones(out.shape, bool)
    mask[0:S] = False
    for k in range(N):
        mask[offsetT + k * pitch: offsetT + k * pitch + S] = False
    assert np.isnan(out[mask]).all(), 'the kernel wrote outside the addressed rows'
    # one pressure per state through the same (wide) kernel: two groups against the scalar calls
    field = np.where(np.arange(S) % 2 == 0, 0.5, 4.0)
    d_p = torch.from_numpy(np.ascontiguousarray(field)).cuda()
    d_rates.fill_(float('nan'))
    kinetix.productionRatesPressureField(S, offsetT, pitch, d_p, d_state, d_rates)
    torch.cuda.synchronize()
    out = d_rates.cpu().numpy()
    for g, pn in enumerate((0.5, 4.0)):
        sel = np.arange(S) % 2 == g
        got = np.empty((N + 1, int(sel.sum())))
        got[0] = out[0:S][sel]
        for k in range(N):
            got[k + 1] = out[offsetT + k * pitch: offsetT + k * pitch + S][sel]
        e = bk1_errors(got, orc.production_rates(np.ascontiguousarray(st[:, sel]), pn * p_ref, Tref=T_ref))
        assert max(e) <= TOL, (pn, e)


@pytest.mark.parametrize('mech', ['NH3Konnov_edit', 'chempolimi_edit'])
def test_bk1_plog_mechanisms_across_pressures(kinetix, mech):
    """pressure-dependent-Arrhenius (P-log) reactions: below, inside and above the tabulated pressures
    (reference reaction_rates.py:358-388); p != p_ref exercises the non-dimensional pressure argument."""
    N = _setup(kinetix, mech)
    orc = Oracle(mech)
    st = synthetic_states(N, 4000, seed=21)
    for p in (P_ATM, 1013.25, 5.0e5, 2.0265e6, 2.0e7):
        new = _run_bk1(kinetix, st, p / P_ATM)
        ref = orc.production_rates(st, p)
        rate_err, hrr_err = bk1_errors(new, ref)
        print(f'{mech} p={p:g} BK1 vs {orc.kind}: rates {rate_err:.3e} hrr {hrr_err:.3e}')
        assert np.isfinite(new).all() and rate_err <= TOL and hrr_err <= TOL
    cond, visc, rhoD = _run_bk2(kinetix, st, 1.0)
    rc, rv, rrd = orc.transport(st, 1.0)
    assert max(rel_err(cond, rc), rel_err(visc, rv), rel_err(rhoD, rrd)) <= TOL


@pytest.mark.parametrize('mech', ['H2_Konnov', 'H2_new_mech', 'gri30-20', 'gri30-27', 'gri30-35', 'heptaneLu88'])
def test_remaining_shipped_mechanisms(kinetix, mech):
    """every mechanism shipped in kinetix/mechanisms goes through the emitter and matches the reference's own
    generated code (oracle/_ref; the reference's stock kinetix_bk cannot even run the ones without a Pele
    directory)."""
    N = _setup(kinetix, mech)
    orc = Oracle(mech)
    assert orc.kind == 'reference'
    st = synthetic_states(N, 3000, seed=17)
    new = _run_bk1(kinetix, st, 1.0)
    ref = orc.production_rates(st, P_ATM)
    rate_err, hrr_err = bk1_errors(new, ref)
    cond, visc, rhoD = _run_bk2(kinetix, st, 1.0)
    rc, rv, rrd = orc.transport(st, 1.0)
    e2 = max(rel_err(cond, rc), rel_err(visc, rv), rel_err(rhoD, rrd))
    print(f'{mech} ({N} sp): BK1 {rate_err:.2e} hrr {hrr_err:.2e} BK2 {e2:.2e}')
    assert np.isfinite(new).all() and rate_err <= TOL and hrr_err <= TOL and e2 <= TOL


def test_largest_mechanism_etoh(kinetix):
    """EtOHKonnov: 129 species / 1231 reactions incl. the only SRI falloff reactions of the shipped mechanisms
    (reaction_rates.py:346-357; BASELINE config 4); oracle = the reference's own unrolled code (oracle/_ref)."""
    mech = 'EtOHKonnov'
    N = _setup(kinetix, mech)
    assert N == 129 and kinetix.nReactions() == 1231
    orc = Oracle(mech)
    assert orc.kind == 'reference'
    st = synthetic_states(N, 6000, seed=8)
    new = _run_bk1(kinetix, st, 1.0)
    ref = orc.production_rates(st, P_ATM)
    rate_err, hrr_err = bk1_errors(new, ref)
    e_all, e_sig = elementwise_errors(new, ref)
    print(f'{mech} BK1 vs {orc.kind}: rates {rate_err:.3e} hrr {hrr_err:.3e}; element-wise all {e_all:.3e} significant {e_sig:.3e}')
    assert rate_err <= TOL and hrr_err <= TOL and e_sig <= TOL
    # BK2: the 129-species kernel runs 256 threads as two halves sharing 128 states (warps w and w + 4 own the same
    # tensor-memory lanes and split every loop over species); 2.4 batches per persistent CTA plus a ragged tail
    st2 = synthetic_states(N, 148 * 128 * 2 + 57 * 128 + 77, seed=9)
    cond, visc, rhoD = _run_bk2(kinetix, st2, 1.0)
    rc, rv, rrd = orc.transport(st2, 1.0)
    errs = rel_err(cond, rc), rel_err(visc, rv), rel_err(rhoD, rrd)
    print(f'{mech} BK2 {st2.shape[1]} states vs {orc.kind}: {errs}')
    assert np.isfinite(rhoD).all() and max(errs) <= TOL


@pytest.mark.parametrize('mech', ['gri30', 'LiDryer'])
def test_bk2_matches_oracle_on_random_states(kinetix, mech):
    N = _setup(kinetix, mech)
    orc = Oracle(mech)
    st = synthetic_states(N, 20000, seed=4321)
    cond, visc, rhoD = _run_bk2(kinetix, st, 1.0)
    rc, rv, rrd = orc.transport(st, 1.0)
    errs = rel_err(cond, rc), rel_err(visc, rv), rel_err(rhoD, rrd)
    print(f'{mech} BK2 vs {orc.kind}: cond {errs[0]:.3e} visc {errs[1]:.3e} rhoD {errs[2]:.3e}')
    assert max(errs) <= TOL


def test_bk2_last_round_goes_to_a_second_launch(kinetix):
    """Persistent BK2 CTAs take batches round-robin; when the last round would occupy at most half the SMs, its states
    are handled by a second launch of the one-state-per-thread instantiation (pointer-shifted buffers).  Every element
    on both sides of that split, ragged end included, against the oracle; a pitched layout as well."""
    mech = 'gri30'
    N = _setup(kinetix, mech)
    orc = Oracle(mech)
    n_sm = torch.cuda.get_device_properties(0).multi_processor_count
    S = n_sm * 512 * 2 + 10 * 512 + 77                      # two full rounds + 11 batches (the last one ragged)
    st = synthetic_states(N, S, seed=2025)
    cond, visc, rhoD = _run_bk2(kinetix, st, 1.0)
    rc, rv, rrd = orc.transport(st, 1.0)
    split = n_sm * 512 * 2
    for name, sl in (('main launch', slice(0, split)), ('tail launch', slice(split, S))):
        errs = rel_err(cond[sl], rc[sl]), rel_err(visc[sl], rv[sl]), rel_err(rhoD[:, sl], rrd[:, sl])
        print(f'{mech} BK2 {name}: {errs}')
        assert np.isfinite(rhoD[:, sl]).all() and max(errs) <= TOL
    # pitched rows behind a gap: the shifted pointers of the tail launch must respect offsetT / offset
    pitch = S + 11
    slab = np.full((N + 2) * pitch + 5, np.nan)
    slab[0:S] = st[0]
    offsetT = 5 + pitch
    for k in range(N):
        slab[offsetT + k * pitch: offsetT + k * pitch + S] = st[k + 1]
    d_state = torch.from_numpy(slab).cuda()
    d_visc = torch.full((S,), float('nan'), dtype=torch.float64, device='cuda')
    d_cond = torch.full_like(d_visc, float('nan'))
    d_rhoD = torch.full((N * pitch,), float('nan'), dtype=torch.float64, device='cuda')
    kinetix.mixtureAvgTransportProps(S, offsetT, pitch, 1.0, d_state, d_visc, d_cond, d_rhoD)
    torch.cuda.synchronize()
    out = d_rhoD.cpu().numpy().reshape(N, pitch)
    assert np.array_equal(out[:, :S], rhoD) and np.isnan(out[:, S:]).all()
    assert np.array_equal(d_visc.cpu().numpy(), visc) and np.array_equal(d_cond.cpu().numpy(), cond)


def test_gri30_one_million_states_full_compare(kinetix):
    """BASELINE configs 1 and 2 at their stated size: 1 Mi seeded GRI-3.0 states, EVERY output element of BK1 and
    BK2 compared with the reference's own generated code (oracle/_ref on all host cores, ~20 s).  1 Mi states =
    13.8 batches per persistent CTA of the default tensor-memory BK2 kernel (ring wrap, mbarrier phase flips, TMEM
    reuse across batches) and 55 waves of BK1 CTAs."""
    mech = 'gri30'
    N = _setup(kinetix, mech)
    orc = Oracle(mech)
    assert orc.kind == 'reference'
    S = 1 << 20
    st = synthetic_states(N, S, seed=20261017)
    new = _run_bk1(kinetix, st, 1.0)
    ref = orc.production_rates(st, P_ATM)
    assert np.isfinite(new).all()
    rate_err, hrr_err = bk1_errors(new, ref)
    e_all, e_sig = elementwise_errors(new, ref)
    print(f'{mech} BK1 {S} states vs {orc.kind}: rates {rate_err:.3e} hrr {hrr_err:.3e}; element-wise all {e_all:.3e} significant {e_sig:.3e}')
    assert rate_err <= TOL and hrr_err <= TOL and e_sig <= TOL
    del new, ref
    cond, visc, rhoD = _run_bk2(kinetix, st, 1.0)
    rc, rv, rrd = orc.transport(st, 1.0)
    errs = rel_err(cond, rc), rel_err(visc, rv), rel_err(rhoD, rrd)
    print(f'{mech} BK2 {S} states vs {orc.kind}: cond {errs[0]:.3e} visc {errs[1]:.3e} rhoD {errs[2]:.3e} (element-wise)')
    assert max(errs) <= TOL


@pytest.mark.parametrize('variant', ['single', 'single_dense', 'tmem_p1'])
def test_bk2_kernel_variants_match_oracle(kinetix, variant):
    """the opt-in BK2 kernels (prebuilt by __graft_entry__.build() from the same emitter with different options:
    one state per thread with / without the low-rank Wilke factorisation, the tensor-memory kernel with one state per
    thread) compute the same transport properties as the default kernel's
    oracle; ragged size: 2.4 batches per persistent CTA of the tensor-memory kernels (148 SMs x 512 states x 2 + a
    partial round + a partial batch), many waves of the one-state-per-thread kernels."""
    import __graft_entry__ as entry
    lib = os.path.join(entry.variant_dir(variant), 'libkx_mech.so')
    if not os.path.exists(lib):
        pytest.skip(f'variant module {variant} not prebuilt')
    kinetix.init(mech_path('gri30'), cache_dir=os.path.dirname(entry.variant_dir(variant)))
    assert os.path.samefile(kinetix.modulePath(), lib)
    N = kinetix.nSpecies()
    kinetix.build(P_ATM, 1.0, [1.0 / N] * N, True)
    orc = Oracle('gri30')
    st = synthetic_states(N, 148 * 512 * 2 + 57 * 512 + 333, seed=99)
    cond, visc, rhoD = _run_bk2(kinetix, st, 1.0)
    rc, rv, rrd = orc.transport(st, 1.0)
    errs = rel_err(cond, rc), rel_err(visc, rv), rel_err(rhoD, rrd)
    print(f'{variant}: cond {errs[0]:.3e} visc {errs[1]:.3e} rhoD {errs[2]:.3e}')
    assert max(errs) <= TOL


def test_etoh_large_batch_is_replication_invariant(kinetix):
    """size-independent property at a many-wave size for the large-mechanism BK1 layout (scratch slots in shared +
    tensor memory, live set capped with partial rates flushed to / accumulated in the output rows): a batch made of
    320 copies of 1024 states (> 8 full waves of one 256-thread CTA per SM) must give 320 bit-identical copies of the
    1024 results, every row written, and the first copy must match the oracle."""
    mech = 'EtOHKonnov'
    N = _setup(kinetix, mech)
    base = synthetic_states(N, 1024, seed=5)
    reps = 320
    st = np.ascontiguousarray(np.tile(base, (1, reps)))
    new = _run_bk1(kinetix, st, 1.0)
    assert np.isfinite(new).all()
    blocks = new.reshape(N + 1, reps, 1024)
    assert (blocks == blocks[:, :1, :]).all()
    orc = Oracle(mech)
    assert orc.kind == 'reference'
    ref = orc.production_rates(base, P_ATM)
    rate_err, hrr_err = bk1_errors(np.ascontiguousarray(blocks[:, 0, :]), ref)
    print(f'{mech} {st.shape[1]} states: replicas identical; first copy vs {orc.kind}: rates {rate_err:.3e} hrr {hrr_err:.3e}')
    assert rate_err <= TOL and hrr_err <= TOL


@pytest.mark.parametrize('variant', ['bk1_tm', 'bk1_tm_all', 'bk1_tm_2cta'])
def test_bk1_tensor_memory_slots_match_oracle(kinetix, variant):
    """BK1 with its per-thread scratch slots (exp(+-g_k), third-body sums) partly / entirely in tensor memory --
    the layout large mechanisms get automatically (EtOHKonnov, test_largest_mechanism_etoh) -- forced on GRI-3.0:
    one 256-thread CTA per SM with 512 TMEM columns, or two 128-thread CTAs with 256 columns each.  Ragged size:
    the tail warps execute the warp-collective tcgen05.ld / st with clamped state indices."""
    import __graft_entry__ as entry
    lib = os.path.join(entry.variant_dir(variant), 'libkx_mech.so')
    if not os.path.exists(lib):
        pytest.skip(f'variant module {variant} not prebuilt')
    kinetix.init(mech_path('gri30'), cache_dir=os.path.dirname(entry.variant_dir(variant)))
    assert os.path.samefile(kinetix.modulePath(), lib)
    N = kinetix.nSpecies()
    kinetix.build(P_ATM, 1.0, [1.0 / N] * N, True)
    orc = Oracle('gri30')
    for S in (40 * 256 + 77, 5):
        st = synthetic_states(N, S, seed=7)
        new = _run_bk1(kinetix, st, 1.0)
        ref = orc.production_rates(st, P_ATM)
        rate_err, hrr_err = bk1_errors(new, ref)
        print(f'{variant} S={S}: rates {rate_err:.3e} hrr {hrr_err:.3e}')
        assert np.isfinite(new).all() and rate_err <= TOL and hrr_err <= TOL


@pytest.mark.parametrize('mech', ['gri30', 'LiDryer'])
def test_thermo_matches_oracle(kinetix, mech):
    N = _setup(kinetix, mech)
    orc = Oracle(mech)
    st = synthetic_states(N, 5000, seed=99)
    S = st.shape[1]
    d_state = torch.from_numpy(st).cuda()
    rho = torch.empty(S, dtype=torch.float64, device='cuda')
    cp = torch.empty((N, S), dtype=torch.float64, device='cuda')
    rcp = torch.empty(S, dtype=torch.float64, device='cuda')
    kinetix.thermodynamicProps(S, S, S, 1.0, d_state, rho, cp, rcp)
    torch.cuda.synchronize()
    a, b, c = orc.thermo(st, P_ATM)
    errs = rel_err(rho.cpu().numpy(), a), rel_err(cp.cpu().numpy(), b), rel_err(rcp.cpu().numpy(), c)
    print(f'{mech} thermo: {errs}')
    assert max(errs) <= 1e-13


@pytest.mark.parametrize('mech,states', [('gri30', ['initial', 'ignition', 'final']),
                                          ('LiDryer', ['initial', 'ignition'])])
def test_cantera_known_answers(kinetix, mech, states):
    """The reference's --cimode 1 check (bk.cpp:503-520,629-642,182-200,258): state 0 is the reference
    state, rates converted back to molar and compared with Cantera."""
    kinetix.init(mech_path(mech))
    N = kinetix.nSpecies()
    cis = [load_cantera_ci(os.path.join(GOLDEN, f'{mech}.{s}.cantera'), N) for s in states]
    M = np.array(kinetix.molarMasses())
    assert np.max(np.abs(M - cis[0]['M']) / M) < 1e-7          # species-order guard (bk.cpp:291-305)
    p_ref, T_ref = cis[0]['p'], cis[0]['T']
    kinetix.build(p_ref, T_ref, list(cis[0]['Y']), True)
    S = len(cis)
    st = np.empty((N + 1, S))
    for i, c in enumerate(cis):
        st[0, i] = c['T'] / T_ref
        st[1:, i] = c['Y']
    rates = _run_bk1(kinetix, st, cis[0]['p'] / p_ref)
    cond, visc, rhoD = _run_bk2(kinetix, st, cis[0]['p'] / p_ref)
    mw = np.array(kinetix.molecularWeights()) * kinetix.refMeanMolecularWeight()
    for i, c in enumerate(cis):
        molar = rates[1:, i] / mw
        e = [abs(rates[0, i] - c['hrr']) / abs(c['hrr'])]
        for k in range(kinetix.nActiveSpecies()):
            e.append(abs(molar[k] - c['wdot'][k]) / abs(c['wdot'][k]) if abs(c['wdot'][k]) > 1e-50 else abs(molar[k]))
        rtol = 5e-5 if i >= 2 else 2e-8
        print(f'{mech}.{states[i]} rates error_inf {max(e):.3e} < {rtol}')
        assert max(e) < rtol
        et = max(abs(cond[i] - c['conductivity']) / c['conductivity'], abs(visc[i] - c['viscosity']) / c['viscosity'],
                 np.max(np.abs(rhoD[:, i] - c['rhoD']) / c['rhoD']))
        print(f'{mech}.{states[i]} transport error_inf {et:.3e} < 1e-3')
        assert et < 1e-3


def test_offsets_ref_state_and_ragged_sizes(kinetix):
    """offsetT/offset addressing with a padded pitch, T_ref != 1, p != p_ref, and state counts that are
    not multiples of the block size (including 1 and 0)."""
    mech = 'LiDryer'
    kinetix.init(mech_path(mech))
    N = kinetix.nSpecies()
    p_ref, T_ref = 2.0e5, 1000.0
    kinetix.build(p_ref, T_ref, [1.0 / N] * N, True)
    orc = Oracle(mech)
    for S in (1, 31, 129, 1000):
        pitch = S + 13
        st = synthetic_states(N, S, seed=S, Tref=T_ref)
        slab = np.full((N + 2) * pitch + 7, np.nan)
        slab[0:S] = st[0]
        offsetT = 7 + pitch                      # species rows start after a gap
        for k in range(N):
            slab[offsetT + k * pitch: offsetT + k * pitch + S] = st[k + 1]
        d_state = torch.from_numpy(slab).cuda()
        d_rates = torch.full_like(d_state, float('nan'))
        p = 3.0e5
        kinetix.productionRates(S, offsetT, pitch, p / p_ref, d_state, d_rates)
        torch.cuda.synchronize()
        out = d_rates.cpu().numpy()
        new = np.empty_like(st)
        new[0] = out[0:S]
        for k in range(N):
            new[k + 1] = out[offsetT + k * pitch: offsetT + k * pitch + S]
        ref = orc.production_rates(st, p, Tref=T_ref)
        rate_err, hrr_err = bk1_errors(new, ref)
        assert rate_err <= TOL and hrr_err <= TOL, (S, rate_err, hrr_err)
        # nothing outside the addressed rows may be written
        mask = np.ones_like(out, dtype=bool)
        mask[0:S] = False
        for k in range(N):
            mask[offsetT + k * pitch: offsetT + k * pitch + S] = False
        assert np.isnan(out[mask]).all()
    # zero states: no launch, no error
    kinetix.productionRates(0, 0, 0, 1.0, None, None)


def test_negative_mass_fractions_are_clamped(kinetix):
    """Y_k < 0 is clamped to 0 before use (productionRates.okl:27, transportProps.okl:25)."""
    mech = 'LiDryer'
    N = _setup(kinetix, mech)
    orc = Oracle(mech)
    st = synthetic_states(N, 256, seed=5)
    st[1 + (np.arange(256) % N), np.arange(256)] *= -1.0
    new = _run_bk1(kinetix, st, 1.0)
    ref = orc.production_rates(st, P_ATM)
    rate_err, hrr_err = bk1_errors(new, ref)
    assert rate_err <= TOL and hrr_err <= TOL
    cond, visc, rhoD = _run_bk2(kinetix, st, 1.0)
    rc, rv, rrd = orc.transport(st, 1.0)
    assert max(rel_err(cond, rc), rel_err(visc, rv), rel_err(rhoD, rrd)) <= TOL


def test_host_buffer_entry_point_matches_device_path(kinetix):
    mech = 'gri30'
    N = _setup(kinetix, mech)
    S = 1200000          # > 2 pipeline chunks, ragged tail
    st = synthetic_states(N, S, seed=77)
    dev = _run_bk1(kinetix, st, 1.0)
    h_state = torch.from_numpy(st).pin_memory()
    h_rates = torch.empty_like(h_state).pin_memory()
    kinetix.productionRatesHost(S, S, S, 1.0, h_state, h_rates)
    assert np.array_equal(h_rates.numpy(), dev)
    c, v, rd = _run_bk2(kinetix, st, 1.0)
    hv = torch.empty(S, dtype=torch.float64).pin_memory()
    hc = torch.empty(S, dtype=torch.float64).pin_memory()
    hrd = torch.empty((N, S), dtype=torch.float64).pin_memory()
    kinetix.mixtureAvgTransportPropsHost(S, S, S, 1.0, h_state, hv, hc, hrd)
    assert np.array_equal(hv.numpy(), v) and np.array_equal(hc.numpy(), c) and np.array_equal(hrd.numpy(), rd)
    # fused call: one upload, both kernels
    for t in (h_rates, hv, hc, hrd):
        t.fill_(float('nan'))
    kinetix.ratesAndTransportHost(S, S, S, 1.0, h_state, h_rates, hv, hc, hrd)
    assert np.array_equal(h_rates.numpy(), dev) and np.array_equal(hv.numpy(), v)
    assert np.array_equal(hc.numpy(), c) and np.array_equal(hrd.numpy(), rd)


@pytest.mark.parametrize('mech', ['gri30', 'LiDryer'])
def test_single_precision_modes(kinetix, mech):
    """--single-precision: FP32 math with FP64 buffers ("fpmix", what the reference CLI runs) and with FP32
    buffers.  Stated bound: per-state scaled error <= 1e-4 (BK1 rates), 5e-3 (heat release: a cancelling sum
    over species, measured 7e-6 .. 1.4e-3), 5e-5 (BK2), 1e-5 (thermo) against the FP64 oracle over the full T in [300, 2500] K range -- the reference's FP32 code is
    NaN/Inf below ~615 K and its own tolerance is 2e-2 (bk.cpp:198)."""
    kinetix.init(mech_path(mech), single_precision=True)
    N = kinetix.nSpecies()
    kinetix.build(P_ATM, 1.0, [1.0 / N] * N, True)
    orc = Oracle(mech)
    st = synthetic_states(N, 20000, seed=31)
    S = st.shape[1]
    ref = orc.production_rates(st, P_ATM)
    rc, rv, rrd = orc.transport(st, 1.0)
    rho_r, cp_r, rcp_r = orc.thermo(st, P_ATM)
    for dtype, tdt in ((0, torch.float64), (1, torch.float32)):
        d_state = torch.from_numpy(st).to(tdt).cuda()
        d_rates = torch.full_like(d_state, float('nan'))
        kinetix.productionRates(S, S, S, 1.0, d_state, d_rates, dtype=dtype)
        visc = torch.empty(S, dtype=tdt, device='cuda')
        cond = torch.empty_like(visc)
        rhoD = torch.empty((N, S), dtype=tdt, device='cuda')
        kinetix.mixtureAvgTransportProps(S, S, S, 1.0, d_state, visc, cond, rhoD, dtype=dtype)
        rho = torch.empty(S, dtype=tdt, device='cuda')
        cp = torch.empty((N, S), dtype=tdt, device='cuda')
        rcp = torch.empty(S, dtype=tdt, device='cuda')
        kinetix.thermodynamicProps(S, S, S, 1.0, d_state, rho, cp, rcp, dtype=dtype)
        torch.cuda.synchronize()
        new = d_rates.double().cpu().numpy()
        assert np.isfinite(new).all(), 'FP32 path must stay finite down to 300 K'
        rate_err, hrr_err = bk1_errors(new, ref)
        e2 = max(rel_err(cond.double().cpu().numpy(), rc), rel_err(visc.double().cpu().numpy(), rv),
                 rel_err(rhoD.double().cpu().numpy(), rrd))
        e3 = max(rel_err(rho.double().cpu().numpy(), rho_r), rel_err(cp.double().cpu().numpy(), cp_r),
                 rel_err(rcp.double().cpu().numpy(), rcp_r))
        print(f'{mech} single precision dtype={dtype}: BK1 {rate_err:.2e} hrr {hrr_err:.2e} BK2 {e2:.2e} thermo {e3:.2e}')
        assert rate_err <= 1e-4 and hrr_err <= 5e-3 and e2 <= 5e-5 and e3 <= 1e-5
    # FP32 buffers without single_precision are rejected, not silently converted
    kinetix.init(mech_path(mech))
    kinetix.build(P_ATM, 1.0, [1.0 / N] * N, True)
    with pytest.raises(kinetix.KinetixError):
        kinetix.productionRates(S, S, S, 1.0, d_state, d_rates, dtype=1)


def test_fit_rcp_diff_coeffs_option(kinetix):
    """--fit-rcpDiffCoeffs fits 1/D_kj instead of D_kj (no division at run time); results differ from the
    default at the 1e-4 level, so parity is checked against the reference generated WITH the flag."""
    mech = 'LiDryer'
    orc = Oracle(mech, variant='rcpdiff')
    if orc.kind != 'reference':
        pytest.skip('oracle/_ref rcpdiff variant not built')
    kinetix.init(mech_path(mech), fit_rcpDiffCoeffs=True)
    N = kinetix.nSpecies()
    kinetix.build(P_ATM, 1.0, [1.0 / N] * N, True)
    assert '-rcpdiff' in kinetix.modulePath()
    st = synthetic_states(N, 5000, seed=12)
    cond, visc, rhoD = _run_bk2(kinetix, st, 1.0)
    rc, rv, rrd = orc.transport(st, 1.0)
    errs = rel_err(cond, rc), rel_err(visc, rv), rel_err(rhoD, rrd)
    print(f'{mech} BK2 --fit-rcpDiffCoeffs vs reference: {errs}')
    assert max(errs) <= TOL
    # and it really is a different fit than the default
    d_rc, d_rv, d_rrd = Oracle(mech).transport(st, 1.0)
    assert rel_err(rhoD, d_rrd) > 1e-8


def test_batch_beyond_32bit_element_indices(kinetix):
    """(N+1) * n_states > 2^31: the reference's `int` indexing (productionRates.okl:27,49; bk.cpp:29) overflows at
    39.7 M GRI states; here offsets are 64-bit.  The last states of a 41 M batch must match the oracle."""
    mech = 'gri30'
    N = _setup(kinetix, mech)
    S = 41_000_000
    assert (N + 1) * S > 2 ** 31
    free, _ = torch.cuda.mem_get_info()
    if free < 2.2 * (N + 1) * S * 8:
        pytest.skip('not enough device memory for the 41 M state batch')
    n_tail = 2048
    tail = synthetic_states(N, n_tail, seed=404)
    d_state = torch.empty((N + 1, S), dtype=torch.float64, device='cuda')
    d_state[0].fill_(1500.0)
    d_state[1:].fill_(1.0 / N)
    d_state[:, S - n_tail:] = torch.from_numpy(tail).cuda()
    d_rates = torch.empty_like(d_state)
    kinetix.productionRates(S, S, S, 1.0, d_state, d_rates)
    torch.cuda.synchronize()
    new = d_rates[:, S - n_tail:].cpu().numpy()
    ref = Oracle(mech).production_rates(tail, P_ATM)
    rate_err, hrr_err = bk1_errors(new, ref)
    assert rate_err <= TOL and hrr_err <= TOL
    # BK2 on the same slab (rho*D rows reuse the rates buffer)
    visc = torch.empty(S, dtype=torch.float64, device='cuda')
    cond = torch.empty_like(visc)
    rhoD = d_rates[1:]
    kinetix.mixtureAvgTransportProps(S, S, S, 1.0, d_state, visc, cond, rhoD)
    torch.cuda.synchronize()
    rc, rv, rrd = Oracle(mech).transport(tail, 1.0)
    assert rel_err(cond[S - n_tail:].cpu().numpy(), rc) <= TOL
    assert rel_err(rhoD[:, S - n_tail:].cpu().numpy(), rrd) <= TOL
    del d_state, d_rates, visc, cond
    torch.cuda.empty_cache()


@pytest.mark.parametrize('mech', ['gri30', 'NH3Konnov_edit'])
def test_per_state_pressure_field(kinetix, mech):
    """extension: one pressure per state.  States are dealt round-robin to four pressures; every
    group must match the oracle run at that group's (single, reference-style) pressure -- for NH3Konnov_edit the
    pressures straddle its P-log tables, so ln P is evaluated per state -- and a constant field must reproduce the
    scalar call (to rounding: p_ref/R * p instead of (p_ref * p)/R)."""
    N = _setup(kinetix, mech)
    orc = Oracle(mech)
    assert orc.kind == 'reference'
    S = 4000
    st = synthetic_states(N, S, seed=77)
    p_nd = np.array([0.3, 1.0, 7.0, 40.0])
    field = np.ascontiguousarray(p_nd[np.arange(S) % 4])
    d_state = torch.from_numpy(st).cuda()
    d_p = torch.from_numpy(field).cuda()
    d_rates = torch.full_like(d_state, float('nan'))
    kinetix.productionRatesPressureField(S, S, S, d_p, d_state, d_rates)
    rho = torch.full((S,), float('nan'), dtype=torch.float64, device='cuda')
    rhocp = torch.full_like(rho, float('nan'))
    cpi = torch.full((N, S), float('nan'), dtype=torch.float64, device='cuda')
    kinetix.thermodynamicPropsPressureField(S, S, S, d_p, d_state, rho, cpi, rhocp)
    torch.cuda.synchronize()
    new, rho, rhocp, cpi = d_rates.cpu().numpy(), rho.cpu().numpy(), rhocp.cpu().numpy(), cpi.cpu().numpy()
    for g, pn in enumerate(p_nd):
        sel = np.arange(S) % 4 == g
        sub = np.ascontiguousarray(st[:, sel])
        ref = orc.production_rates(sub, pn * P_ATM)
        rate_err, hrr_err = bk1_errors(np.ascontiguousarray(new[:, sel]), ref)
        r_rho, r_cp, r_rhocp = orc.thermo(sub, pn * P_ATM)
        e3 = max(rel_err(rho[sel], r_rho), rel_err(cpi[:, sel], r_cp), rel_err(rhocp[sel], r_rhocp))
        print(f'{mech} p = {pn} atm: BK1 {rate_err:.2e} hrr {hrr_err:.2e} thermo {e3:.2e}')
        assert rate_err <= TOL and hrr_err <= TOL and e3 <= TOL
    # constant field == scalar pressure
    d_p.fill_(7.0)
    kinetix.productionRatesPressureField(S, S, S, d_p, d_state, d_rates)
    d_ref = torch.empty_like(d_rates)
    kinetix.productionRates(S, S, S, 7.0, d_state, d_ref)
    torch.cuda.synchronize()
    e_rate, e_hrr = bk1_errors(d_rates.cpu().numpy(), d_ref.cpu().numpy())
    assert e_rate <= 1e-13 and e_hrr <= 1e-13


def test_errors_are_loud(kinetix):
    kinetix.finalize()
    with pytest.raises(kinetix.KinetixError):
        kinetix.productionRates(1, 1, 1, 1.0, 0, 0)            # not initialised
    with pytest.raises(kinetix.KinetixError):
        kinetix.init('/nonexistent/mech.yaml')
    with pytest.raises(kinetix.KinetixError):
        kinetix.init(mech_path('gri30'), tool='Pele')


def test_two_devices_from_one_process(kinetix):
    """per-device contexts (kx_select_device): one process initialises DIFFERENT mechanisms on two GPUs and
    alternates between them; each launch lands on its context's device whatever device torch has current"""
    if torch.cuda.device_count() < 2:
        pytest.skip('needs two GPUs')
    mechs = {0: 'gri30', 1: 'LiDryer'}
    for dev, mech in mechs.items():
        kinetix.init(mech_path(mech), device_id=dev)
        N = kinetix.nSpecies()
        kinetix.build(P_ATM, 1.0, [1.0 / N] * N, True)
    torch.cuda.set_device(0)
    for dev in (1, 0, 1):
        kinetix.selectDevice(dev)
        assert kinetix.currentDevice() == dev
        N = kinetix.nSpecies()
        st = synthetic_states(N, 3000, seed=dev)
        with torch.cuda.device(dev):
            d_state = torch.from_numpy(st).cuda()
            d_rates = torch.full_like(d_state, float('nan'))
        kinetix.productionRates(3000, 3000, 3000, 1.0, d_state, d_rates, stream=0)
        torch.cuda.synchronize(dev)
        ref = Oracle(mechs[dev]).production_rates(st, P_ATM)
        rate_err, hrr_err = bk1_errors(d_rates.cpu().numpy(), ref)
        assert rate_err <= TOL and hrr_err <= TOL
    for dev in (0, 1):
        kinetix.selectDevice(dev)
        kinetix.finalize()
