"""CPU tests of the generated reference-signature routines (kinetix_b200/core/emit_routines.py):
file set, signatures, mech.h, and parity of the routines -- compiled for the host by tests/emu/routines_host.py --
with the oracle at the FP64 bound 1e-10.  The GPU counterpart is tests/test_routines_gpu.py."""
import os
import re

import numpy as np
import pytest

from oracle.port import synthetic_states
from tests.common import Oracle, bk1_errors, rel_err
from tests.emu.routines_host import RoutineHost

GOLDEN = os.path.join(os.path.dirname(__file__), 'golden', 'mech_h')
P_ATM = 101325.0
TOL = 1e-10

# the signatures the reference prints (reaction_rates.py:560-562, thermodynamics.py:70-72,88-89,
# mix_transport.py:480-481,504-505,607-609), whitespace-normalised
SIGNATURES = {
    'rates': '__KINETIX_DEVICE__ __KINETIX_INLINE__ void kinetix_species_rates(const cfloat lnT, const cfloat T, '
             'const cfloat T2, const cfloat T3, const cfloat T4, const cfloat rcpT, const cfloat P, const cfloat lnP, '
             'const cfloat* Ci, cfloat* wdot)',
    'enthalpy_RT': '__KINETIX_DEVICE__ __KINETIX_INLINE__ void kinetix_enthalpy_RT(const cfloat T, const cfloat T2, '
                   'const cfloat T3, const cfloat T4, const cfloat rcpT, cfloat* h_RT)',
    'heat_capacity_R': '__KINETIX_DEVICE__ __KINETIX_INLINE__ void kinetix_molar_heat_capacity_R(const cfloat T, '
                       'const cfloat T2, const cfloat T3, const cfloat T4, cfloat* cp_R)',
    'conductivity': '__KINETIX_DEVICE__ __KINETIX_INLINE__ cfloat kinetix_conductivity(cfloat rcpMbar, cfloat lnT, '
                    'cfloat lnT2, cfloat lnT3, cfloat lnT4, cfloat Xi[])',
    'viscosity': '__KINETIX_DEVICE__ __KINETIX_INLINE__ cfloat kinetix_viscosity(cfloat lnT, cfloat lnT2, cfloat lnT3, '
                 'cfloat lnT4, cfloat Xi[])',
    'diffusivity': '__KINETIX_DEVICE__ __KINETIX_INLINE__ void kinetix_diffusivity(cfloat Mbar, cfloat p, cfloat TsqrT, '
                   'cfloat lnT, cfloat lnT2, cfloat lnT3, cfloat lnT4, cfloat Xi[], cfloat* Dkm)',
}


@pytest.fixture(scope='module')
def lidryer():
    return RoutineHost('LiDryer')


def test_file_set_and_signatures(lidryer):
    """generate.py:62-77 file set; every routine carries the reference's signature verbatim"""
    names = set(os.path.basename(p) for p in lidryer.paths.values())
    assert names == {'mech.h', 'kinetix_b200_routines.cuh'} | {f + '.cuh' for f in SIGNATURES}
    for f, sig in SIGNATURES.items():
        text = re.sub(r'\s+', ' ', open(lidryer.paths[f + '.cuh']).read())
        assert sig in text, f
    # the reference's literal file names on request
    cpp = RoutineHost('LiDryer', ext='cpp')
    assert {'rates.cpp', 'diffusivity.cpp', 'enthalpy_RT.cpp'} <= set(os.path.basename(p) for p in cpp.paths.values())


@pytest.mark.parametrize('mech', ['LiDryer', 'gri30'])
def test_mech_h_is_the_references(mech):
    """mech.h equals the file the reference generator writes (mechanism.py:29-48), byte for byte"""
    rh = RoutineHost(mech, transport=False)
    assert open(rh.paths['mech.h']).read().strip() == open(os.path.join(GOLDEN, mech + '.mech.h')).read().strip()


@pytest.mark.parametrize('mech,n', [('LiDryer', 2000), ('gri30', 600)])
def test_routines_match_oracle(mech, n):
    rh = RoutineHost(mech)
    orc = Oracle(mech)
    st = synthetic_states(rh.N, n, seed=3)
    e_rate, e_hrr = bk1_errors(rh.production_rates(st, P_ATM), orc.production_rates(st, P_ATM))
    c, v, rd = rh.transport(st)
    rc, rv, rrd = orc.transport(st)
    e2 = max(rel_err(c, rc), rel_err(v, rv), rel_err(rd, rrd))
    e3 = max(rel_err(a, b) for a, b in zip(rh.thermo(st, P_ATM), orc.thermo(st, P_ATM)))
    print(f'{mech} routines vs {orc.kind}: rates {e_rate:.2e} hrr {e_hrr:.2e} transport {e2:.2e} thermo {e3:.2e}')
    assert e_rate <= TOL and e_hrr <= TOL and e2 <= TOL and e3 <= 1e-13


def test_routines_plog_pressures_and_accumulation():
    """P-log mechanism across its pressure tables (P and lnP are ARGUMENTS of the routine); wdot is ADDED to"""
    mech = 'chempolimi_edit'
    rh = RoutineHost(mech, transport=False)
    orc = Oracle(mech)
    st = synthetic_states(rh.N, 300, seed=21)
    for p in (P_ATM, 1013.25, 2.0265e6, 2.0e7):
        e_rate, e_hrr = bk1_errors(rh.production_rates(st, p), orc.production_rates(st, p))
        assert e_rate <= TOL and e_hrr <= TOL, (p, e_rate, e_hrr)


def test_rcp_diff_variant_changes_only_the_fit(lidryer):
    st = synthetic_states(lidryer.N, 200, seed=5)
    a = RoutineHost('LiDryer', fit_rcp_diff=True).transport(st)
    b = lidryer.transport(st)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    assert 1e-8 < rel_err(a[2], b[2]) < 1e-2
    orc = Oracle('LiDryer', variant='rcpdiff')
    if orc.kind == 'reference':
        assert rel_err(a[2], orc.transport(st)[2]) <= TOL
