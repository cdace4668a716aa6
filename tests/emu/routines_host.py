"""TEST INFRASTRUCTURE: the GENERATED reference-signature routines (kinetix_b200/core/emit_routines.py) compiled for
the HOST and driven state by state, so that the `-m "not gpu"` suite can compare them with the oracle.

The routine files are taken exactly as write_routines() wrote them; the includer-supplied macros get host values
(the reference does the same for its SERIAL backend, benchmark/src/kinetix.cpp:215-252), and kx_math.cuh is the
product's own text with its inline-PTX statements replaced by their C meaning (tests/emu/emulate.py).  The driver
loops restate what csrc/kx_routine_kernels.cu does per thread.  Only tests use this; the product has no CPU path.
"""
import ctypes
import hashlib
import os
import shutil
import subprocess
import tempfile

import numpy as np

from kinetix_b200.core import constants as const
from kinetix_b200.core.emit_routines import write_routines
from kinetix_b200.core.mechanism import load_mechanism
from kinetix_b200.core.transport_fit import fit_transport
from tests.emu.emulate import HERE, ROOT, _math_header

_DRIVER = r'''
#include "cuda_emu.h"
#define __KINETIX_DEVICE__
#define __KINETIX_INLINE__ static inline
#define __KINETIX_CONST__ static const
#define cfloat double
#define dfloat double
#include "kinetix_b200_routines.cuh"
static const double R_GAS = 1.380649e-23 * 6.02214076e23;

static double composition(const double* state, long id, long offsetT, long offset, double* w)
{
  double s = 0;
  for (int k = 0; k < n_species; k++) { const double y = state[id + offsetT + k * offset]; w[k] = (y > 0 ? y : 0) * kinetix_rcp_molar_mass[k]; s += w[k]; }
  return s;
}

extern "C" int rt_n_species() { return n_species; }
extern "C" void rt_production_rates(long n, long offsetT, long offset, double pressure_R, double pressure,
                                    const double* state, double* rates, double Tref)
{
  for (long id = 0; id < n; id++) {
    const double T = Tref * state[id], rcpT = 1 / T, lnT = log(T);
    double c[n_species], wdot[n_species];
    const double rho = pressure_R * rcpT / composition(state, id, offsetT, offset, c);
    for (int k = 0; k < n_species; k++) { c[k] *= rho; wdot[k] = 0; }
    kinetix_species_rates(lnT, T, T * T, T * T * T, T * T * T * T, rcpT, pressure, log(pressure), c, wdot);
    for (int k = 0; k < n_species; k++) rates[id + offsetT + k * offset] = kinetix_molar_mass[k] * wdot[k];
    kinetix_enthalpy_RT(T, T * T, T * T * T, T * T * T * T, rcpT, c);
    double h = 0;
    for (int k = 0; k < n_species; k++) h += wdot[k] * c[k];
    rates[id] = -R_GAS * T * h;
  }
}
#ifndef RT_NO_TRANSPORT
extern "C" void rt_transport(long n, long offsetT, long offset, double pressure, const double* state, double* cond,
                             double* visc, double* rhoD, double Tref)
{
  for (long id = 0; id < n; id++) {
    const double T = Tref * state[id], lnT = log(T), l2 = lnT * lnT, sq = sqrt(T);
    double X[n_species], D[n_species];
    const double rcpMbar = composition(state, id, offsetT, offset, X), Mbar = 1 / rcpMbar;
    for (int k = 0; k < n_species; k++) X[k] *= Mbar;
    cond[id] = sq * kinetix_conductivity(rcpMbar, lnT, l2, l2 * lnT, l2 * l2, X);
    visc[id] = sq * kinetix_viscosity(lnT, l2, l2 * lnT, l2 * l2, X);
    kinetix_diffusivity(Mbar, pressure, T * sq, lnT, l2, l2 * lnT, l2 * l2, X, D);
    const double rho = pressure / R_GAS / T * Mbar;
    for (int k = 0; k < n_species; k++) rhoD[k * offset + id] = rho * D[k];
  }
}
#endif
extern "C" void rt_thermo(long n, long offsetT, long offset, double pressure_R, const double* state, double* rho,
                          double* cp, double* rhoCp, double Tref)
{
  for (long id = 0; id < n; id++) {
    const double T = Tref * state[id];
    double w[n_species], cpR[n_species];
    const double rcpMbar = composition(state, id, offsetT, offset, w), Mbar = 1 / rcpMbar;
    const double d = pressure_R / T * Mbar;
    rho[id] = d;
    kinetix_molar_heat_capacity_R(T, T * T, T * T * T, T * T * T * T, cpR);
    double m = 0;
    for (int k = 0; k < n_species; k++) { cp[k * offset + id] = cpR[k] * R_GAS * kinetix_rcp_molar_mass[k]; m += cpR[k] * w[k] * Mbar; }
    rhoCp[id] = d * (m * R_GAS * rcpMbar);
  }
}
'''


class RoutineHost:
    def __init__(self, mech_name, fit_rcp_diff=False, transport=True, ext='cuh'):
        mech = load_mechanism(os.path.join(ROOT, 'kinetix_b200', 'mechanisms', mech_name + '.yaml'))
        fits = fit_transport(mech, reciprocal_diffusivity=fit_rcp_diff) if transport else None
        work = os.path.join(tempfile.gettempdir(), f'kx_emu_{os.getuid()}')
        os.makedirs(work, exist_ok=True)
        d = tempfile.mkdtemp(dir=work)
        self.paths, self.stats = write_routines(mech, fits, d, ext=ext)
        self.dir = d
        math = _math_header()
        with open(os.path.join(d, 'kx_math.cuh'), 'w') as fh:     # host meaning of the product's math header
            fh.write(math)
        shutil.copyfile(os.path.join(HERE, 'cuda_emu.h'), os.path.join(d, 'cuda_emu.h'))
        with open(os.path.join(d, 'driver.cpp'), 'w') as fh:
            fh.write(_DRIVER)
        h = hashlib.sha256()
        for f in sorted(os.listdir(d)):
            h.update(open(os.path.join(d, f), 'rb').read())
        lib = os.path.join(work, f'rt_{mech_name}_{h.hexdigest()[:16]}.so')
        if not os.path.exists(lib):
            cmd = ['g++', '-std=c++17', '-O1', '-ffp-contract=off', '-shared', '-fPIC', '-w', '-I', d,
                   '-o', lib + '.tmp', os.path.join(d, 'driver.cpp')] + ([] if transport else ['-DRT_NO_TRANSPORT'])
            r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
            if r.returncode != 0:
                raise RuntimeError('g++ failed on the generated routines:\n' + r.stdout[-4000:])
            os.replace(lib + '.tmp', lib)
        self.lib = ctypes.CDLL(lib)
        self.N = self.lib.rt_n_species()

    @staticmethod
    def _p(a):
        return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))

    def production_rates(self, st, p, Tref=1.0):
        st = np.ascontiguousarray(st)
        S = st.shape[1]
        out = np.full_like(st, np.nan)
        L = ctypes.c_long
        self.lib.rt_production_rates(L(S), L(S), L(S), ctypes.c_double(p / const.R_GAS), ctypes.c_double(p),
                                     self._p(st), self._p(out), ctypes.c_double(Tref))
        return out

    def transport(self, st, pressure_nd=1.0, Tref=1.0):
        st = np.ascontiguousarray(st)
        S = st.shape[1]
        c, v, rd = np.empty(S), np.empty(S), np.empty((self.N, S))
        L = ctypes.c_long
        self.lib.rt_transport(L(S), L(S), L(S), ctypes.c_double(pressure_nd), self._p(st), self._p(c), self._p(v),
                              self._p(rd), ctypes.c_double(Tref))
        return c, v, rd

    def thermo(self, st, p, Tref=1.0):
        st = np.ascontiguousarray(st)
        S = st.shape[1]
        rho, cp, rcp = np.empty(S), np.empty((self.N, S)), np.empty(S)
        L = ctypes.c_long
        self.lib.rt_thermo(L(S), L(S), L(S), ctypes.c_double(p / const.R_GAS), self._p(st), self._p(rho), self._p(cp),
                           self._p(rcp), ctypes.c_double(Tref))
        return rho, cp, rcp
