#!/usr/bin/env python3
"""Small invocation of every kernel for `compute-sanitizer --tool memcheck|racecheck python tools/sanitize_smoke.py`."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import kinetix_b200.host as kinetix  # noqa: E402
from oracle.port import synthetic_states  # noqa: E402

# gri30 / heptaneLu88: tensor-memory BK2 kernel (256 / 128 threads); LiDryer: one-state-per-thread kernel; sp: FP32;
# EtOHKonnov: BK1 with scratch slots in shared + tensor memory (256-thread CTA)
MECHS = (('gri30', False), ('heptaneLu88', False), ('LiDryer', False), ('LiDryer', True), ('EtOHKonnov', False))
if len(sys.argv) > 1:
    MECHS = tuple((m, False) for m in sys.argv[1:])
for mech, sp in MECHS:
    kinetix.init(os.path.join(ROOT, 'kinetix_b200', 'mechanisms', mech + '.yaml'), single_precision=sp)
    N = kinetix.nSpecies()
    kinetix.build(101325.0, 1.0, [1.0 / N] * N, True)
    # more than one persistent-CTA round of 512 states, ragged tail; KX_SMOKE_STATES=60001 reaches the wide BK1 kernel
    S = int(os.environ.get('KX_SMOKE_STATES', '1555'))
    st = torch.from_numpy(synthetic_states(N, S)).cuda()
    r = torch.empty_like(st)
    v = torch.empty(S, dtype=torch.float64, device='cuda')
    c = torch.empty_like(v)
    d = torch.empty((N, S), dtype=torch.float64, device='cuda')
    kinetix.productionRates(S, S, S, 1.0, st, r)
    kinetix.mixtureAvgTransportProps(S, S, S, 1.0, st, v, c, d)
    kinetix.thermodynamicProps(S, S, S, 1.0, st, v, d, c)
    torch.cuda.synchronize()
    kinetix.finalize()
print('sanitize_smoke: done')
