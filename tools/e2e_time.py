#!/usr/bin/env python3
"""Time the host-buffer entry points (pinned host memory) for one mechanism: states/s through BK1+BK2 with the
reference's two calls and with the fused call.  KX_HOST_CHUNK=<states> selects the pipeline chunk."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import kinetix_b200.host as kinetix  # noqa: E402
from oracle.port import synthetic_states  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 21
kinetix.init(os.path.join(ROOT, 'kinetix_b200', 'mechanisms', 'gri30.yaml'))
N = kinetix.nSpecies()
kinetix.build(101325.0, 1.0, [1.0 / N] * N, True)
base = torch.from_numpy(synthetic_states(N, 1 << 16, seed=1))
st = base.repeat(1, n // (1 << 16)).contiguous().pin_memory()
rates = torch.empty_like(st).pin_memory()
visc = torch.empty(n, dtype=torch.float64).pin_memory()
cond = torch.empty(n, dtype=torch.float64).pin_memory()
rhoD = torch.empty((N, n), dtype=torch.float64).pin_memory()


def sep():
    kinetix.productionRatesHost(n, n, n, 1.0, st, rates)
    kinetix.mixtureAvgTransportPropsHost(n, n, n, 1.0, st, visc, cond, rhoD)


def fused():
    kinetix.ratesAndTransportHost(n, n, n, 1.0, st, rates, visc, cond, rhoD)


for name, fn in (('separate', sep), ('fused', fused)):
    fn()
    t0 = time.perf_counter()
    for _ in range(4):
        fn()
    dt = (time.perf_counter() - t0) / 4
    print(f'chunk {os.environ.get("KX_HOST_CHUNK", "default")} n {n} {name}: {n / dt / 1e6:.1f} M states/s')
