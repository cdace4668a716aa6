#!/usr/bin/env python3
"""Replication invariance of BK2 (and BK1) at scale: R copies of the same B states must give R bit-identical result
blocks -- a data race on the coefficient ring, the tensor-memory exchange areas or the X rows would show as differences."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import kinetix_b200.host as kx  # noqa: E402
from oracle.port import synthetic_states  # noqa: E402

mech = sys.argv[1] if len(sys.argv) > 1 else 'EtOHKonnov'
B, R = 1 << 16, int(sys.argv[2]) if len(sys.argv) > 2 else 64
kx.init(os.path.join(ROOT, 'kinetix_b200', 'mechanisms', mech + '.yaml'))
N = kx.nSpecies()
kx.build(101325.0, 1.0, [1.0 / N] * N, True)
S = B * R
st = torch.from_numpy(synthetic_states(N, B, seed=11)).cuda().repeat(1, R).contiguous()
rates = torch.empty_like(st)
visc = torch.empty(S, dtype=torch.float64, device='cuda')
cond = torch.empty_like(visc)
rhoD = torch.empty((N, S), dtype=torch.float64, device='cuda')
bad = 0
for it in range(3):
    kx.productionRates(S, S, S, 1.0, st, rates)
    kx.mixtureAvgTransportProps(S, S, S, 1.0, st, visc, cond, rhoD)
    torch.cuda.synchronize()
    for name, t in (('rates', rates), ('rhoD', rhoD), ('visc', visc[None]), ('cond', cond[None])):
        v = t.view(t.shape[0], R, B)
        diff = int((v != v[:, :1]).sum())
        bad += diff
        if diff:
            print(f'{mech} pass {it}: {name}: {diff} elements differ between replicas')
print(f'{mech}: {S} states = {R} replicas x {B}, 3 passes: {"bit-identical replicas" if not bad else str(bad) + " DIFFERENCES"}')
sys.exit(1 if bad else 0)
