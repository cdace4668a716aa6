"""CPU parity of the GENERATED BK1 kernel text: tests/emu executes the emitter's output thread by thread on the host
(CUDA keywords, tensor memory and the few inline-PTX statements replaced by their C meaning; exp / log / reciprocal
are the product's own code) and compares with the oracle at the 1e-10 bound.  This checks, without a GPU, what the
emitter decides: reaction schedule, activation / retirement, slot recycling in shared and tensor memory, live-range
splitting through the output rows, P-log branches, per-state pressure.  It is test infrastructure, not a CPU path
of the product (the GPU tests in test_parity_gpu.py remain the parity tests proper)."""
import numpy as np
import pytest

from oracle.port import synthetic_states
from tests.common import Oracle, bk1_errors
from tests.emu.emulate import BK1Emulator

TOL = 1e-10
P_ATM = 101325.0


def _check(emu, orc, st, p, label):
    new = emu.production_rates(st, p)
    ref = orc.production_rates(st, p)
    assert np.isfinite(new).all(), f'{label}: a scratch slot was read before it was written, or a row was not stored'
    rate_err, hrr_err = bk1_errors(new, ref)
    print(f'{label}: emulated kernel vs {orc.kind}: rates {rate_err:.2e} hrr {hrr_err:.2e}')
    assert rate_err <= TOL and hrr_err <= TOL
    return new


@pytest.mark.parametrize('mech,prefer_ref', [('LiDryer', True), ('gri30', True), ('NH3Konnov_edit', True),
                                             ('H2_Konnov', True), ('H2_new_mech', True), ('gri30-20', True),
                                             ('gri30-27', True), ('gri30-35', True), ('heptaneLu88', True)])
def test_generated_bk1_on_cpu(mech, prefer_ref):
    """every shipped mechanism (EtOHKonnov and chempolimi_edit have their own tests below), at 1 atm and at 30 bar"""
    emu = BK1Emulator(mech)
    orc = Oracle(mech, prefer_ref=prefer_ref)
    st = synthetic_states(emu.mech.n_species, 300, seed=11)       # ragged: tail threads re-run the last state
    _check(emu, orc, st, P_ATM, mech)
    _check(emu, orc, st[:, :64], 3.0e6, mech + ' 30 bar')


def test_generated_bk1_plog_and_pressure_field_on_cpu():
    """chempolimi_edit: P-log reactions below / inside / above their pressure tables, then the same three pressures
    as ONE launch with a per-state pressure field (the PF = true kernel)."""
    mech = 'chempolimi_edit'
    emu = BK1Emulator(mech)
    orc = Oracle(mech)
    N = emu.mech.n_species
    st = synthetic_states(N, 90, seed=5)
    pressures = (1013.25, 5.0e5, 2.0e7)
    per_p = [_check(emu, orc, st, p, f'{mech} p={p:g}') for p in pressures]
    field = np.repeat(np.array(pressures) / P_ATM, 90)
    got = emu.production_rates(np.tile(st, (1, 3)), P_ATM, p_field=field)
    for i in range(3):
        blk = got[:, 90 * i:90 * (i + 1)]
        assert np.array_equal(blk, per_p[i]) or bk1_errors(blk, per_p[i])[0] <= 1e-13


@pytest.mark.parametrize('options', [
    {'bk1_tmem': True},                                                              # third-body sums in TMEM
    {'bk1_tmem': True, 'bk1_smem_cap': 0},                                           # every slot in TMEM
    {'bk1_tmem': True, 'bk1_tmem_block': 128, 'bk1_tmem_ctas': 2, 'bk1_smem_cap': 20},   # two CTAs x 256 columns
    {'live_cap': 12},                                                                # 86 suspensions / re-activations
    {'bk1_tmem': True, 'live_cap': 10, 'bk1_smem_cap': 8},
    {'cold_uses': 15, 'cold_slot_cap': 90},                  # rarely used species: C_k / wdot_k in shared-memory slots
    {'cold_uses': 1000, 'cold_slot_cap': 120, 'live_cap': 14},
    {'bk1_layout': 'classic'},       # three 128-thread CTAs, all slots in shared memory: the small-launch kernel's text
    {'cold_uses': 1000, 'cold_slot_cap': 90, 'cold_conc_only': True},    # C_k alone in a slot (heptaneLu88's layout)
    {'bk1_tmem': True, 'bk1_smem_cap': 30, 'cold_uses': 1000, 'cold_conc_only': True, 'gibbs_prefer_tm': True, 'live_cap': 16},
], ids=['tm', 'tm_all', 'tm_2cta', 'live_cap', 'tm_live_cap', 'cold', 'cold_live_cap', 'classic', 'cold_conc', 'tm_cold_conc'])
def test_generated_bk1_slot_layouts_on_cpu(options):
    emu = BK1Emulator('gri30', options)
    st = synthetic_states(53, 300, seed=7)
    _check(emu, Oracle('gri30'), st, P_ATM, f'gri30 {options}')
    sch = emu.stats['bk1_schedule']
    if options.get('bk1_tmem'):
        # the emulated tensor memory saw accesses, inside the CTA's column budget and the thread's own lane
        assert sch['tmem_slots'] > 0 and 0 < emu.tmem_columns_touched() <= 512
    if options.get('live_cap'):
        assert sch['peak_live'] <= options['live_cap'] and sch['suspensions'] > 0
    if options.get('cold_uses'):
        assert sch['cold_activations'] > 0 and sch['smem_slots'] <= options.get('cold_slot_cap', 1 << 20)


def test_generated_bk1_largest_mechanism_on_cpu():
    """EtOHKonnov as shipped: 256-thread CTA, 110 shared + 49 tensor-memory slots, live set capped at 60 (28
    suspensions), SRI falloff."""
    emu = BK1Emulator('EtOHKonnov')
    sch = emu.stats['bk1_schedule']
    assert emu.block == 256 and sch['tmem_slots'] > 0 and sch['peak_live'] <= 60
    st = synthetic_states(129, 260, seed=2)
    _check(emu, Oracle('EtOHKonnov', prefer_ref=False), st, P_ATM, 'EtOHKonnov')


@pytest.mark.parametrize('mech', ['LiDryer', 'gri30'])
def test_generated_fp32_math_bk1_on_cpu(mech):
    """the --single-precision kernel (log2-space FP32 math, constants as immediates) on FP64 buffers, against the
    FP64 oracle at the stated single-precision bounds (rates 1e-4, heat release 5e-3) over the full
    T in [300, 2500] K range -- the reference's own FP32 code is NaN / Inf below ~615 K."""
    emu = BK1Emulator(mech, single_precision=True)
    st = synthetic_states(emu.mech.n_species, 2000, seed=31)
    new = emu.production_rates(st, P_ATM)
    ref = Oracle(mech).production_rates(st, P_ATM)
    assert np.isfinite(new).all()
    rate_err, hrr_err = bk1_errors(new, ref)
    print(f'{mech} FP32-math kernel (emulated) vs FP64 oracle: rates {rate_err:.2e} hrr {hrr_err:.2e}')
    assert rate_err <= 1e-4 and hrr_err <= 5e-3


def test_plog_with_a_single_tabulated_pressure(tmp_path):
    """a P-log reaction with ONE rate entry has nothing to interpolate: the emitter writes its plain Arrhenius rate
    (the text used to be invalid C for this case).  Checked against the same mechanism with that reaction turned into
    an ordinary elementary reaction: identical results at any pressure."""
    import os
    from tests.common import mech_path
    text = open(mech_path('chempolimi_edit')).read()
    old = """  type: pressure-dependent-Arrhenius
  rate-constants:
  - {P: 0.1 atm, A: 9.2e+11, b: -0.01, Ea: 5039.4924}
  - {P: 1.0 atm, A: 1.2e+12, b: -0.03, Ea: 5074.4662}
  - {P: 10.0 atm, A: 4.7e+12, b: -0.2, Ea: 5344.4435}
"""
    assert old in text
    one = tmp_path / 'plog1.yaml'
    one.write_text(text.replace(old, """  type: pressure-dependent-Arrhenius
  rate-constants:
  - {P: 1.0 atm, A: 1.2e+12, b: -0.03, Ea: 5074.4662}
"""))
    plain = tmp_path / 'plain.yaml'
    plain.write_text(text.replace(old, "  rate-constant: {A: 1.2e+12, b: -0.03, Ea: 5074.4662}\n"))
    a, b = BK1Emulator(str(one)), BK1Emulator(str(plain))
    st = synthetic_states(a.mech.n_species, 64, seed=2)
    for p in (2.0e4, P_ATM, 5.0e6):
        ra, rb = a.production_rates(st, p), b.production_rates(st, p)
        assert np.isfinite(ra).all()
        e_rate, e_hrr = bk1_errors(ra, rb)
        assert e_rate <= 1e-13 and e_hrr <= 1e-12, (p, e_rate, e_hrr)


def test_non_positive_pre_exponential_factor_is_rejected_with_a_message(tmp_path):
    from kinetix_b200.core.emit_module import emit_module
    from kinetix_b200.core.mechanism import load_mechanism
    from tests.common import mech_path
    text = open(mech_path('LiDryer')).read()
    bad = tmp_path / 'negA.yaml'
    bad.write_text(text.replace('A: 3.547e+15', 'A: -3.547e+15', 1))
    with pytest.raises(SystemExit) as e:
        emit_module(load_mechanism(str(bad)), None, {})
    assert 'not positive' in str(e.value)
