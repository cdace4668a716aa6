#!/usr/bin/env python3
"""Quick device-side timing of BK1/BK2/thermo for one mechanism (development aid; bench.py is the contract)."""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import kinetix_b200.host as kinetix  # noqa: E402
from oracle.port import synthetic_states  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--mech', default='gri30')
ap.add_argument('--n', type=int, default=1 << 22)
ap.add_argument('--reps', type=int, default=5)
ap.add_argument('--cache', default=None)
ap.add_argument('--check', action='store_true')
ap.add_argument('--tag', default='')
ap.add_argument('--sp', type=int, default=-1, help='-1: FP64; 0: fpmix (FP64 buffers); 1: FP32 buffers')
a = ap.parse_args()
kinetix.init(os.path.join(ROOT, 'kinetix_b200', 'mechanisms', a.mech + '.yaml'), cache_dir=a.cache,
             single_precision=a.sp >= 0)
tdt = torch.float32 if a.sp == 1 else torch.float64
dt = max(a.sp, 0)
N = kinetix.nSpecies()
kinetix.build(101325.0, 1.0, [1.0 / N] * N, True)
S = a.n
base = torch.from_numpy(synthetic_states(N, 1 << 16, seed=1)).cuda()
st = base.repeat(1, -(-S // (1 << 16)))[:, :S].contiguous().to(tdt)      # (any S, not only multiples of 64 Ki)
rates = torch.empty_like(st)
visc = torch.empty(S, dtype=tdt, device='cuda')
cond = torch.empty_like(visc)
rhoD = torch.empty((N, S), dtype=tdt, device='cuda')


def timeit(fn):
    fn(); fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / a.reps


t1 = timeit(lambda: kinetix.productionRates(S, S, S, 1.0, st, rates, dtype=dt))
t2 = timeit(lambda: kinetix.mixtureAvgTransportProps(S, S, S, 1.0, st, visc, cond, rhoD, dtype=dt))
t3 = timeit(lambda: kinetix.thermodynamicProps(S, S, S, 1.0, st, visc, rhoD, cond, dtype=dt))
kinetix.mixtureAvgTransportProps(S, S, S, 1.0, st, visc, cond, rhoD, dtype=dt)
torch.cuda.synchronize()
msg = ''
if a.check:
    from tests.common import Oracle, bk1_errors, rel_err
    n = 4096
    orc = Oracle(a.mech)
    ref = orc.production_rates(base[:, :n].cpu().numpy(), 101325.0)
    e1 = bk1_errors(rates[:, :n].double().cpu().numpy(), ref)
    rc, rv, rrd = orc.transport(base[:, :n].cpu().numpy(), 1.0)
    e2 = max(rel_err(cond[:n].double().cpu().numpy(), rc), rel_err(visc[:n].double().cpu().numpy(), rv), rel_err(rhoD[:, :n].double().cpu().numpy(), rrd))
    msg = f' | err bk1 {e1[0]:.1e}/{e1[1]:.1e} bk2 {e2:.1e}'
print(f'{a.tag} {a.mech} S={S}: BK1 {t1:.3f} ms = {S / t1 / 1e3:.1f} Mstates/s | BK2 {t2:.3f} ms = {S / t2 / 1e3:.1f} Mstates/s | '
      f'thermo {t3:.3f} ms = {S / t3 / 1e3:.1f} Mstates/s ({(2 * N + 3) * 8 * S / t3 / 1e6:.0f} GB/s){msg}')
