"""Shared helpers for the test-suite: oracle access (tests are allowed to use oracle/) and error norms."""
import ctypes
import os

import numpy as np

from oracle import port as oport

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, 'oracle', '_ref')
MECH_DIR = os.path.join(ROOT, 'kinetix_b200', 'mechanisms')
R = oport.R
_dp = ctypes.POINTER(ctypes.c_double)


def mech_path(name):
    return os.path.join(MECH_DIR, name + '.yaml')


def ref_library(mech, variant='parity'):
    """The reference's generated code compiled by oracle/build_ref.py, or None if not built."""
    path = os.path.join(REF_DIR, f'libref_{mech}.{variant}.so')
    return ctypes.CDLL(path) if os.path.exists(path) else None


def _p(a):
    return a.ctypes.data_as(_dp)


class Oracle:
    """BK1/BK2/thermo answers for a (N+1, S) state slab: oracle/_ref when built, else the numpy port."""

    def __init__(self, mech, prefer_ref=True, variant='parity'):
        self.port = oport.Port(mech)
        self.lib = ref_library(mech, variant) if prefer_ref else None
        self.kind = 'reference' if self.lib is not None else 'port'
        self.N = self.port.N

    def production_rates(self, st, p, Tref=1.0):
        if self.lib is None:
            return self.port.production_rates(st, p / R, p, Tref)
        st = np.ascontiguousarray(st)
        S = st.shape[1]
        out = np.empty_like(st)
        self.lib.ref_production_rates(ctypes.c_long(S), ctypes.c_long(S), ctypes.c_long(S), ctypes.c_double(p / R),
                                      ctypes.c_double(p), _p(st), _p(out), ctypes.c_double(Tref))
        return out

    def transport(self, st, pressure_nd=1.0, Tref=1.0):
        if self.lib is None:
            return self.port.transport(st, pressure_nd, Tref)
        st = np.ascontiguousarray(st)
        S = st.shape[1]
        c, v, rd = np.empty(S), np.empty(S), np.empty((self.N, S))
        self.lib.ref_transport(ctypes.c_long(S), ctypes.c_long(S), ctypes.c_long(S), ctypes.c_double(pressure_nd),
                               _p(st), _p(c), _p(v), _p(rd), ctypes.c_double(Tref))
        return c, v, rd

    def thermo(self, st, p, Tref=1.0):
        if self.lib is None:
            return self.port.thermo(st, p / R, Tref)
        st = np.ascontiguousarray(st)
        S = st.shape[1]
        rho, cp, rcp = np.empty(S), np.empty((self.N, S)), np.empty(S)
        self.lib.ref_thermo(ctypes.c_long(S), ctypes.c_long(S), ctypes.c_long(S), ctypes.c_double(p / R), _p(st),
                            _p(rho), _p(cp), _p(rcp), ctypes.c_double(Tref))
        return rho, cp, rcp


def bk1_errors(new, ref):
    """SURVEY.md 8c norm: per state, max_k |new-ref| / max_k |ref| over the N mass-rate rows; the heat
    release row relative to its own magnitude (floored by the per-state rate scale)."""
    scale = np.abs(ref[1:]).max(axis=0)
    scale = np.where(scale > 0, scale, 1.0)
    rate_err = (np.abs(new[1:] - ref[1:]).max(axis=0) / scale).max()
    hrr_floor = 1e-6 * np.abs(ref[0]).max()
    hrr_err = (np.abs(new[0] - ref[0]) / np.maximum(np.abs(ref[0]), hrr_floor)).max()
    return float(rate_err), float(hrr_err)


def rel_err(new, ref):
    return float(np.max(np.abs(new - ref) / np.abs(ref)))
