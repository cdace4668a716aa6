"""Mechanism front-end: Cantera YAML -> intermediate representation (IR).

This is the host-side "model" the sm_100a emitter consumes.  It reproduces the *semantics* of the
reference front-end so the emitted kernels compute what the reference's generated routines compute:

  * species order = YAML definition order with species that take part in no reaction moved to the
    end (reference kinetix/core/generate.py:37-47);
  * molar masses in kg/mol from standard atomic weights (mix_transport.py:218-220);
  * NASA-7 two-range polynomials, single-range species get a duplicated piece with T_mid = 1000 K
    (mix_transport.py:222-231), more than two ranges are rejected (:235-239);
  * reactions: integer stoichiometry parsed from the equation string, A converted from
    cm-mol-s to m-mol-s with (1e-6)^(order-1), activation energy to a temperature, reaction
    classification, falloff / Troe / SRI / P-log parameters, third-body efficiencies
    (reaction_rates.py:30-190).

Everything is plain data (lists / floats) so it can be dumped to JSON for the C++ host and for
parity tests against the reference's own parse (tests/golden/*.mech.json).
"""
import math
import re
from dataclasses import dataclass, field

from . import constants as const
from .yaml12 import load_file

_THIRD_BODY = re.compile(r'(\s*\(\+[^\)]+\))|(\s*\+\s*M\s*)')
_IRREV_ARROW = re.compile(r'[^<]=>')
_PRESSURE_NUMBER = re.compile(r'[-+]?\d*\.?\d+([eE][-+]?\d+)?')


@dataclass
class Arrhenius:
    A: float        # pre-exponential, SI (m^3/mol)^(order-1)/s
    b: float        # temperature exponent
    Ta: float       # activation temperature Ea/R [K]

    def as_list(self):
        return [self.A, self.b, self.Ta]


@dataclass
class Reaction:
    equation: str
    kind: str                   # elementary | irreversible | three-body | pressure-modification | Troe | SRI | P-log
    reversible: bool
    nu_reac: list               # per-species reactant stoichiometric coefficients (ints)
    nu_prod: list
    rate: Arrhenius             # k_inf for falloff, first entry for P-log
    k0: Arrhenius = None
    troe: dict = None           # {A, T3, T1, T2 (inf when absent)}
    sri: dict = None            # {A, B, C, D, E}
    plog: list = None           # [(P_pa, [Arrhenius, ...]), ...] ascending as listed
    efficiencies: list = None   # per-species efficiency (default filled in) or None
    third_body_index: int = -1  # explicit collider "(+X)" that is a species, else -1

    @property
    def nu_net(self):
        return [p - r for r, p in zip(self.nu_reac, self.nu_prod)]

    @property
    def sum_net(self):
        return sum(self.nu_net)


@dataclass
class Species:
    name: str
    M: float                    # kg/mol
    T_mid: float
    nasa_lo: list               # 7 coefficients, T <= T_mid
    nasa_hi: list
    T_ranges: list              # as listed in the YAML (2 or 3 entries)
    transport: dict = field(default_factory=dict)


@dataclass
class Mechanism:
    name: str
    species: list
    n_active: int
    reactions: list
    units: dict

    @property
    def n_species(self):
        return len(self.species)

    @property
    def n_reactions(self):
        return len(self.reactions)

    @property
    def species_names(self):
        return [s.name for s in self.species]

    @property
    def molar_masses(self):
        return [s.M for s in self.species]


def _parse_side(side, names):
    """'2 OH (+M)' style side -> per-species integer coefficients.  Tokens that are not species
    (e.g. a bare 'M') contribute nothing, exactly like the reference's list comprehension."""
    counts = [0] * len(names)
    index = {n: i for i, n in enumerate(names)}
    for token in side.split(' + '):
        token = _THIRD_BODY.sub('', token.strip())
        if ' ' in token:
            parts = token.split(' ')
            coeff, name = int(parts[0]), parts[1]
        else:
            coeff, name = 1, token
        if name in index:
            counts[index[name]] += coeff
    return counts


def _arrhenius(entry, units, cm3_exponent):
    ea_units = units['activation-energy']
    if ea_units == 'K':
        Ta = entry['Ea']
    elif ea_units == 'cal/mol':
        Ta = entry['Ea'] * const.CAL / const.R_GAS
    else:
        raise SystemExit('Error: Unknown units for activation-energy')
    return Arrhenius(entry['A'] * pow(1e-6, cm3_exponent), entry['b'], Ta)


def parse_reaction(names, units, r):
    """One YAML reaction entry -> Reaction (reference reaction_rates.py:30-190)."""
    eq = r['equation']
    sides = [s.strip() for s in re.split('<?=>', eq)]
    nu_reac, nu_prod = (_parse_side(s, names) for s in sides[:2])

    third_body_index = -1
    m = _THIRD_BODY.search(eq)
    if m:
        tb = (m.group(1) or m.group(2))
        for ch in '+() ':
            tb = tb.replace(ch, '')
        if tb != 'M' and tb in names:
            third_body_index = names.index(tb)

    rtype = r.get('type')
    order = sum(nu_reac) + (1 if rtype == 'three-body' else 0)
    if r.get('rate-constant') or r.get('high-P-rate-constant'):
        rate = _arrhenius(r.get('rate-constant', r.get('high-P-rate-constant')), units, order - 1)
    else:
        rate = _arrhenius(r['rate-constants'][0], units, order - 1)

    irreversible_arrow = bool(_IRREV_ARROW.search(eq))
    rx = Reaction(eq, 'elementary', True, nu_reac, nu_prod, rate, third_body_index=third_body_index)

    if rtype is None or rtype == 'elementary':
        if irreversible_arrow:
            rx.kind, rx.reversible = 'irreversible', False
        elif '<=>' in eq or '= ' in eq:
            assert r.get('reversible', True)
        else:
            raise SystemExit(f"Error: unknown reaction: '{r}'")
    elif rtype == 'three-body':
        rx.kind, rx.reversible = 'three-body', not irreversible_arrow
    elif rtype == 'falloff':
        rx.reversible = not irreversible_arrow
        rx.k0 = _arrhenius(r['low-P-rate-constant'], units, order)
        if r.get('Troe'):
            t = r['Troe']
            rx.kind = 'Troe'
            rx.troe = dict(A=t['A'], T3=t['T3'], T1=t['T1'], T2=t.get('T2', float('inf')))
        elif r.get('SRI'):
            s = r['SRI']
            rx.kind = 'SRI'
            rx.sri = dict(A=s['A'], B=s['B'], C=s['C'], D=s.get('D', 1), E=s.get('E', 0))
        else:
            rx.kind = 'pressure-modification'
    elif rtype == 'pressure-dependent-Arrhenius':
        rx.kind = 'P-log'
        rx.plog = []
        last_p = None
        for idx, entry in enumerate(r['rate-constants']):
            k = _arrhenius(entry, units, order - 1)
            p = float(_PRESSURE_NUMBER.search(entry['P']).group()) * const.ONE_ATM
            if idx != 0 and p == last_p:
                rx.plog[-1][1].append(k)       # several Arrhenius terms at one pressure: summed
            elif any(p == q for q, _ in rx.plog):
                # the reference keys a dict by pressure: a non-adjacent repeat overwrites
                for item in rx.plog:
                    if item[0] == p:
                        item[1][:] = [k]
            else:
                rx.plog.append((p, [k]))
            last_p = p
        # reference quirk (reaction_rates.py:177-179): "'=' in equation" is always true, so P-log
        # reactions are always treated as reversible, whatever arrow they use.
        rx.reversible = True
    else:
        raise SystemExit(f"Error: unknown reaction: '{r}'")

    if r.get('efficiencies'):
        default = r.get('default-efficiency', 1)
        rx.efficiencies = [r['efficiencies'].get(n, default) for n in names]
    return rx


def _species_from_yaml(s):
    M = sum(count * const.atomic_weight(el) / 1e3 for el, count in s['composition'].items())
    th = s['thermo']
    pieces = [list(map(float, p)) for p in th['data']]
    T_ranges = list(th['temperature-ranges'])
    if len(pieces) > 2:
        raise SystemExit(f"Specie {s['name']} has more than 2 sets of thermodynamic coefficients. "
                         f"This is not currently supported!")
    if len(pieces) < 2:
        T_mid = 1000
        pieces.append(list(pieces[0]))
    else:
        T_mid = T_ranges[1]
    tr = s.get('transport', {})
    transport = {}
    if tr:
        transport = dict(
            dof={'atom': 0, 'linear': 1, 'nonlinear': 3 / 2}[tr['geometry']],
            well_depth=tr['well-depth'] * const.K_BOLTZMANN,       # J
            diameter=tr['diameter'] * 1e-10,                        # m
            dipole=tr.get('dipole', 0) * const.DEBYE,               # C m
            polarizability=tr.get('polarizability', 0) * 1e-30,     # m^3
            rot_relax=float(tr.get('rotational-relaxation', 0)),
        )
    return Species(s['name'], M, T_mid, pieces[0], pieces[1], T_ranges, transport)


def load_mechanism(path):
    """Read a Cantera YAML file and build the IR with the reference's species ordering."""
    model = load_file(path)
    u = model['units']
    # the reference only supports cm / s / mol input (mechanism.py:22-25)
    if not (u.get('length', 'm') == 'cm' and u.get('time', 's') == 's' and u.get('quantity', 'mol') == 'mol'):
        raise SystemExit('Error: mechanism units must be length: cm, time: s, quantity: mol')

    names0 = [s['name'] for s in model['species']]
    first_pass = [parse_reaction(names0, u, r) for r in model['reactions']]
    takes_part = [any(rx.nu_net[i] != 0 for rx in first_pass) for i in range(len(names0))]
    active = [s for s, t in zip(model['species'], takes_part) if t]
    inert = [s for s, t in zip(model['species'], takes_part) if not t]

    species = [_species_from_yaml(s) for s in active + inert]
    names = [s.name for s in species]
    reactions = [parse_reaction(names, u, r) for r in model['reactions']]
    import os
    stem = os.path.splitext(os.path.basename(path))[0]
    return Mechanism(stem, species, len(active), reactions, dict(u))


# ---------------------------------------------------------------------------------------------
# serialisation (host getters, parity tests)
# ---------------------------------------------------------------------------------------------

def _num(x):
    if isinstance(x, float) and math.isinf(x):
        return 'inf' if x > 0 else '-inf'
    return x


def reaction_to_dict(rx):
    d = dict(equation=rx.equation, kind=rx.kind, reversible=rx.reversible,
             reactants={str(i): c for i, c in enumerate(rx.nu_reac) if c},
             products={str(i): c for i, c in enumerate(rx.nu_prod) if c},
             rate=rx.rate.as_list(), third_body_index=rx.third_body_index)
    if rx.k0 is not None:
        d['k0'] = rx.k0.as_list()
    if rx.troe is not None:
        d['troe'] = {k: _num(v) for k, v in rx.troe.items()}
    if rx.sri is not None:
        d['sri'] = dict(rx.sri)
    if rx.plog is not None:
        d['plog'] = [[p, [k.as_list() for k in ks]] for p, ks in rx.plog]
    if rx.efficiencies is not None:
        d['efficiencies'] = list(rx.efficiencies)
    return d


def mechanism_to_dict(mech):
    return dict(
        name=mech.name, n_species=mech.n_species, n_active=mech.n_active, n_reactions=mech.n_reactions,
        species=[dict(name=s.name, M=s.M, T_mid=s.T_mid, nasa_lo=s.nasa_lo, nasa_hi=s.nasa_hi,
                      T_ranges=s.T_ranges) for s in mech.species],
        reactions=[reaction_to_dict(r) for r in mech.reactions],
    )
