#!/usr/bin/env python3
"""One-off data extraction (run in the dev container, where /root/reference exists).

Writes kinetix_b200/core/data/mm_collision_integrals.json: the Monchick & Mason (1961)
reduced collision-integral tables Omega*(2,2) and A* versus (T*, delta*) that both Cantera
(MMCollisionInt.cpp) and the reference (kinetix/core/constants.py:155-330) tabulate.  These are
published physical data, stored here as data so that our own transport-fit code
(kinetix_b200/core/transport_fit.py) reproduces the reference's polynomial fits.
"""
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..', 'oracle', 'shims'))
sys.path.insert(0, '/root/reference')
from kinetix.core import constants as rc  # noqa: E402

out = {
    'source': 'Monchick & Mason, J. Chem. Phys. 35, 1676 (1961); as tabulated in Cantera MMCollisionInt.cpp',
    'delta_star': list(rc.header_delta_star),
    'T_star': list(rc.header_T_star[1:-1]),  # 37 tabulated reduced temperatures 0.1 .. 100
    'omega22': rc.collision_integrals_Omega_star_22,   # 37 rows  (T* = 0.1 .. 100)
    'a_star': rc.collision_integrals_A_star,           # 39 rows  (T* = eps, 0.1 .. 100, 500)
}
assert len(out['omega22']) == 37 and len(out['a_star']) == 39 and len(out['T_star']) == 37
path = os.path.join(os.path.dirname(__file__), '..', 'kinetix_b200', 'core', 'data', 'mm_collision_integrals.json')
with open(path, 'w') as fh:
    json.dump(out, fh, indent=0)
print('wrote', os.path.normpath(path))
