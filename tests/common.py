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


def _parallel(fn, st, min_chunk=4096):
    """run fn(sub_slab) over column chunks of `st` on all host cores (the reference library releases the GIL under
    ctypes; states are independent) and concatenate the results along the state axis"""
    from concurrent.futures import ThreadPoolExecutor
    S = st.shape[1]
    workers = max(1, min(len(os.sched_getaffinity(0)) if hasattr(os, 'sched_getaffinity') else (os.cpu_count() or 1),
                         S // min_chunk))
    if workers == 1:
        return fn(st)
    bounds = np.linspace(0, S, workers + 1).astype(int)
    subs = [np.ascontiguousarray(st[:, a:b]) for a, b in zip(bounds[:-1], bounds[1:])]
    with ThreadPoolExecutor(workers) as ex:
        parts = list(ex.map(fn, subs))
    if isinstance(parts[0], tuple):
        return tuple(np.concatenate([p[i] for p in parts], axis=-1) for i in range(len(parts[0])))
    return np.concatenate(parts, axis=-1)


class Oracle:
    """BK1/BK2/thermo answers for a (N+1, S) state slab: oracle/_ref when built, else the numpy port.
    Large slabs are cut into per-core chunks (results are identical: one state never sees another)."""

    def __init__(self, mech, prefer_ref=True, variant='parity'):
        self.port = oport.Port(mech)
        self.lib = ref_library(mech, variant) if prefer_ref else None
        self.kind = 'reference' if self.lib is not None else 'port'
        self.N = self.port.N

    def production_rates(self, st, p, Tref=1.0):
        if st.shape[1] >= 16384 and self.lib is not None:
            return _parallel(lambda sub: self.production_rates(sub, p, Tref), st)
        if self.lib is None:
            return self.port.production_rates(st, p / R, p, Tref)
        st = np.ascontiguousarray(st)
        S = st.shape[1]
        out = np.empty_like(st)
        self.lib.ref_production_rates(ctypes.c_long(S), ctypes.c_long(S), ctypes.c_long(S), ctypes.c_double(p / R),
                                      ctypes.c_double(p), _p(st), _p(out), ctypes.c_double(Tref))
        return out

    def transport(self, st, pressure_nd=1.0, Tref=1.0):
        if st.shape[1] >= 16384 and self.lib is not None:
            return _parallel(lambda sub: self.transport(sub, pressure_nd, Tref), st)
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


def elementwise_errors(new, ref, significant=1e-3):
    """Element-wise relative errors of the N mass-rate rows, next to the per-state scaled norm of bk1_errors
    ('report element-wise too'): (max over ALL non-zero elements, max over the elements that carry at
    least `significant` of their state's largest rate).  Rates are sums over reactions with cancellation, so the
    first number is dominated by elements ~1e-12 of the state's scale and is reported, not bounded: the reference's
    own rolled / unrolled / fast-math builds differ by up to 9e-10 there; the second is held to the
    1e-10 contract."""
    n, r = new[1:], ref[1:]
    scale = np.abs(r).max(axis=0, keepdims=True)
    nz = r != 0
    e = np.zeros_like(r)
    e[nz] = np.abs(n[nz] - r[nz]) / np.abs(r[nz])
    sig = np.abs(r) >= significant * scale
    return float(e.max()), float(e[sig & nz].max()) if (sig & nz).any() else 0.0
