"""Physical constants and element data used by the mechanism front-end.

Values follow Cantera's ct_defs.h (2019 SI redefinition), which is what the reference uses
(reference kinetix/core/constants.py:16-30): results are compared bit-for-bit with the reference's
generated tables, so the *same* derived values (R = kB*NA, eps0 from alpha, h, e, c) are needed.
"""
import json
import os

K_BOLTZMANN = 1.380649e-23        # J/K
N_AVOGADRO = 6.02214076e23        # 1/mol   (Cantera uses 1/kmol)
SPEED_OF_LIGHT = 299792458.0      # m/s
FINE_STRUCTURE = 7.2973525693e-3
PLANCK = 6.62607015e-34           # J s
ELECTRON_CHARGE = 1.602176634e-19  # C
DEBYE = 3.33564e-30               # C m
CAL = 4.184                       # J
ONE_ATM = 1.01325e5               # Pa

MU0 = 2. * FINE_STRUCTURE * PLANCK / (ELECTRON_CHARGE * ELECTRON_CHARGE * SPEED_OF_LIGHT)
EPSILON0 = 1. / (SPEED_OF_LIGHT * SPEED_OF_LIGHT * MU0)
R_GAS = K_BOLTZMANN * N_AVOGADRO  # J/mol/K = 8.31446261815324

# floating-point guards the reference folds into generated code (general_utils.py:235-236)
FLOAT_MIN = 1e-300
FLOAT_MAX = 1e300

# IUPAC standard atomic weights [g/mol] as abridged in Cantera's Elements.cpp
# (reference kinetix/core/constants.py:33-152 holds the same table).
ATOMIC_WEIGHT = dict(
    H=1.008, He=4.002602, Li=6.94, Be=9.0121831, B=10.81, C=12.011, N=14.007, O=15.999,
    F=18.998403163, Ne=20.1797, Na=22.98976928, Mg=24.305, Al=26.9815384, Si=28.085,
    P=30.973761998, S=32.06, Cl=35.45, Ar=39.95, K=39.0983, Ca=40.078, Sc=44.955908,
    Ti=47.867, V=50.9415, Cr=51.9961, Mn=54.938043, Fe=55.845, Co=58.933194, Ni=58.6934,
    Cu=63.546, Zn=65.38, Ga=69.723, Ge=72.630, As=74.921595, Se=78.971, Br=79.904,
    Kr=83.798, Rb=85.4678, Sr=87.62, Y=88.90584, Zr=91.224, Nb=92.90637, Mo=95.95,
    Ru=101.07, Rh=102.90549, Pd=106.42, Ag=107.8682, Cd=112.414, In=114.818, Sn=118.710,
    Sb=121.760, Te=127.60, I=126.90447, Xe=131.293, Cs=132.90545196, Ba=137.327,
    W=183.84, Pt=195.084, Au=196.966570, Hg=200.592, Pb=207.2, U=238.02891,
    D=2.0141017781, E=0.000545,
)


def atomic_weight(symbol):
    """Standard atomic weight in g/mol; element symbols are matched case-insensitively
    (mechanism files write ``AR``/``Ar``, ``HE``/``He``)."""
    key = symbol[0].upper() + symbol[1:].lower()
    return ATOMIC_WEIGHT[key]


def load_collision_tables():
    """Monchick-Mason Omega*(2,2) and A* tables (data file, see tools/extract_collision_tables.py)."""
    path = os.path.join(os.path.dirname(__file__), 'data', 'mm_collision_integrals.json')
    with open(path) as fh:
        return json.load(fh)
