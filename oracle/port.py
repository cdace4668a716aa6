"""TEST INFRASTRUCTURE ONLY -- numpy restatement ("port") of the reference's BK1 / BK2 / thermo path.

Kind: "port".  Pinned (see tests/test_oracle.py) against
  * the reference's Cantera known-answer files tests/golden/ci_data/*.cantera at the reference's own
    tolerances (benchmark/src/bk.cpp:148,198-199,258), and
  * oracle/_ref (the reference's generated code compiled with g++) on seeded random states to ~1e-12,
    whenever that library has been built.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  The product (kinetix_b200/) never does; it fails loudly without its CUDA library.

The mechanism description comes from tests/golden/<mech>.mech.json and <mech>.transport.npz, which
were dumped from the reference's own Python front-end by oracle/gen_golden.py -- no parsing or
fitting code is shared with the product.

What is restated (all states at once, one numpy array per scalar of the reference's code):
  productionRates.okl:10-64        -> Port.production_rates
  kinetix_species_rates (unrolled) -> Port._species_rates   reaction_rates.py:291-416,549-613
  kinetix_enthalpy_RT              -> Port._h_RT            thermodynamics.py:65-80
  transportProps.okl:11-49         -> Port.transport
  kinetix_conductivity             -> mix_transport.py:474-495
  kinetix_viscosity                -> mix_transport.py:498-557 (non-grouped form)
  kinetix_diffusivity              -> mix_transport.py:595-626 (symmetric form)
  thermoCoeffs.okl:10-40           -> Port.thermo           thermodynamics.py:83-97
"""
import json
import math
import os

import numpy as np

KB = 1.380649e-23
NA = 6.02214076e23
R = KB * NA                      # constants.py:30 / kinetix.cpp:37
ONE_ATM = 1.01325e5
FLOAT_MIN = 1e-300               # general_utils.py:236 (generation-time guard)
CFLOAT_MIN = 1e-300              # kinetix.cpp:244 (run-time guard, FP64 build)

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests', 'golden')


def _f(x):
    return float(x) if not isinstance(x, str) else float(x)


class Port:
    def __init__(self, mech, golden_dir=GOLDEN, transport=True):
        with open(os.path.join(golden_dir, mech + '.mech.json')) as fh:
            m = json.load(fh)
        self.name = mech
        self.N = m['n_species']
        self.n_active = m['n_active']
        self.R_count = m['n_reactions']
        self.names = [s['name'] for s in m['species']]
        self.M = np.array([s['M'] for s in m['species']])
        self.rcpM = np.array([1. / s['M'] for s in m['species']])       # mechanism.py:43-44
        self.T_mid = [s['T_mid'] for s in m['species']]
        self.lo = [s['nasa_lo'] for s in m['species']]
        self.hi = [s['nasa_hi'] for s in m['species']]
        self.reactions = m['reactions']
        self.tp = None
        tpath = os.path.join(golden_dir, mech + '.transport.npz')
        if transport and os.path.exists(tpath):
            self.tp = np.load(tpath)

    # ---- state decode shared by the three OKL kernels ------------------------------------------
    def _decode(self, state):
        Y = np.maximum(0., state[1:])
        w = Y * self.rcpM[:, None]
        rcpMbar = np.zeros(state.shape[1])
        for k in range(self.N):
            rcpMbar = rcpMbar + w[k]
        return w, rcpMbar, 1. / rcpMbar

    def _piecewise(self, T, expr):
        """write_energy(): `if (T <= T_mid) low-piece else high-piece` per species."""
        out = []
        for k in range(len(self._sel)):
            a, b = expr(self.lo[self._sel[k]]), expr(self.hi[self._sel[k]])
            out.append(np.where(T <= self.T_mid[self._sel[k]], a, b))
        return out

    # ---- BK1 ----------------------------------------------------------------------------------
    @staticmethod
    def _arrhenius(rc, lnT, T, rcpT):
        A, beta, E = rc
        if beta == 0 and E != 0:
            return np.exp(math.log(A) + (-E) * rcpT)
        if beta == 0 and E == 0:
            return np.full_like(T, A)
        if E == 0:
            if beta == -2:
                return A * rcpT * rcpT
            if beta == -1:
                return A * rcpT
            if beta == 1:
                return A * T
            if beta == 2:
                return A * T * T
            return np.exp(math.log(A) + beta * lnT)
        return np.exp(math.log(A) + beta * lnT + (-E) * rcpT)

    @staticmethod
    def _arrhenius_ratio(r, lnT, T, rcpT):
        """k0/k_inf as one exponential (reaction_rates.py:244-259)."""
        A_inf, b_inf, E_inf = r['rate']
        A0, b0, E0 = r['k0']
        if (A0 - A_inf) != 0 and ((b0 - b_inf) != 0 or (E0 - E_inf) != 0):
            arg = np.zeros_like(T)
            if (E0 - E_inf) != 0:
                arg = arg + (-E0 + E_inf) * rcpT
            if (b0 - b_inf) != 0:
                arg = arg + (b0 - b_inf) * lnT
            return np.exp(arg + (math.log(A0) - math.log(A_inf)))
        return np.full_like(T, A0 / A_inf)

    def _plog(self, r, P, lnP, lnT, T, rcpT):
        """Unrolled P-log branch chain (reaction_rates.py:358-388); P is uniform over the batch."""
        plog = r['plog']

        def ksum(ks):
            tot = None
            for rc in ks:
                v = self._arrhenius(rc, lnT, T, rcpT)
                tot = v if tot is None else tot + v
            return tot

        n = len(plog)
        for i in range(n - 1):
            p1, p2 = plog[i][0], plog[i + 1][0]
            if p1 < P < p2:
                kf1, kf2 = ksum(plog[i][1]), ksum(plog[i + 1][1])
                lnp1, lnp2 = math.log(p1), math.log(p2)
                return np.exp(np.log(kf1) + (np.log(kf2) - np.log(kf1)) * (lnP - lnp1) * (1 / (lnp2 - lnp1)))
            if i == 0 and P <= p1:
                return ksum(plog[i][1])
            if i > 0 and P == p1:
                return ksum(plog[i][1])
            if i == n - 2 and P >= p2:
                return ksum(plog[i + 1][1])
        raise RuntimeError('P-log: pressure falls through the reference branch chain (kf undefined there)')

    def _species_rates(self, lnT, T, T2, T3, T4, rcpT, P, lnP, C):
        N = self.N
        self._sel = list(range(self.n_active))
        g = self._piecewise(T, lambda a: (a[5] * rcpT + (a[0] - a[6]) + (-a[0]) * lnT + (-a[1] / 2) * T +
                                          ((1. / 3. - 1. / 2.) * a[2]) * T2 + ((1. / 4. - 1. / 3.) * a[3]) * T3 +
                                          ((1. / 5. - 1. / 4.) * a[4]) * T4))
        Cm = np.zeros_like(T)
        for k in range(N):
            Cm = Cm + C[k]
        C0 = (ONE_ATM / R) * rcpT
        rcpC0 = (R / ONE_ATM) * T
        wdot = [np.zeros_like(T) for _ in range(N)]

        def prod(side):
            out = None
            for k in sorted(int(i) for i in side):
                for _ in range(side[str(k)]):
                    out = C[k] if out is None else out * C[k]
            return out

        for r in self.reactions:
            kind = r['kind']
            eff = None
            if 'efficiencies' in r:
                eff = Cm
                for k, e in enumerate(r['efficiencies']):
                    if e != 1:
                        eff = eff + (C[k] if e == 2 else (e - 1) * C[k])

            def collider():
                if eff is not None:
                    return eff
                if r['third_body_index'] >= 0:
                    return C[r['third_body_index']]
                return Cm

            if kind in ('elementary', 'irreversible'):
                kf = self._arrhenius(r['rate'], lnT, T, rcpT)
            elif kind == 'three-body':
                kf = self._arrhenius(r['rate'], lnT, T, rcpT) * collider()
            elif kind == 'pressure-modification':
                kf = self._arrhenius(r['rate'], lnT, T, rcpT)
                Pr = self._arrhenius_ratio(r, lnT, T, rcpT) * collider()
                kf = kf * (Pr / (1 + Pr))
            elif kind == 'Troe':
                kf = self._arrhenius(r['rate'], lnT, T, rcpT)
                Pr = self._arrhenius_ratio(r, lnT, T, rcpT) * collider()
                logPr = np.log10(Pr + CFLOAT_MIN)
                tr = {k: _f(v) for k, v in r['troe'].items()}
                t2 = np.exp(-tr['T2'] * rcpT) if tr['T2'] < float('inf') else 0.
                if tr['A'] == 0:
                    Fc = np.exp(-1. / (tr['T3'] + FLOAT_MIN) * T) + t2
                elif tr['A'] == 1:
                    Fc = np.exp(-1. / (tr['T1'] + FLOAT_MIN) * T) + t2
                else:
                    Fc = ((1 - tr['A']) * np.exp(-1. / (tr['T3'] + FLOAT_MIN) * T) +
                          tr['A'] * np.exp(-1. / (tr['T1'] + FLOAT_MIN) * T) + t2)
                logFc = np.log10(Fc)
                c = -.4 - .67 * logFc
                n = .75 - 1.27 * logFc
                f1 = (c + logPr) / (n - .14 * (c + logPr))
                F = np.power(10., logFc / (1.0 + f1 * f1))
                kf = kf * (Pr / (1 + Pr) * F)
            elif kind == 'SRI':
                kf = self._arrhenius(r['rate'], lnT, T, rcpT)
                Pr = self._arrhenius_ratio(r, lnT, T, rcpT) * collider()
                logPr = np.log10(Pr)
                s = r['sri']
                F = (s['D'] * np.power(s['A'] * np.exp(-s['B'] * rcpT) + np.exp(-1. / (s['C'] + FLOAT_MIN) * T),
                                       1. / (1. + logPr * logPr)) * np.power(T, s['E']))
                kf = kf * (Pr / (1 + Pr) * F)
            elif kind == 'P-log':
                kf = self._plog(r, P, lnP, lnT, T, rcpT)
            else:
                raise ValueError(kind)

            Rf = prod(r['reactants'])
            net = {}
            for k, c in r['reactants'].items():
                net[int(k)] = net.get(int(k), 0) - c
            for k, c in r['products'].items():
                net[int(k)] = net.get(int(k), 0) + c
            net = {k: v for k, v in sorted(net.items()) if v != 0}
            if not r['reversible']:
                cR = kf * Rf
            else:
                arg = None
                for k, v in net.items():
                    term = g[k] if v == 1 else (-g[k] if v == -1 else v * g[k])
                    arg = term if arg is None else arg + term
                kr = np.exp(arg)
                s_net = sum(net.values())
                for _ in range(abs(s_net)):
                    kr = kr * (C0 if s_net < 0 else rcpC0)
                cR = kf * (Rf - kr * prod(r['products']))
            for k, v in net.items():
                wdot[k] = wdot[k] + (cR if v == 1 else (-cR if v == -1 else v * cR))
        return wdot

    def _h_RT(self, T, T2, T3, T4, rcpT):
        self._sel = list(range(self.N))
        return self._piecewise(T, lambda a: (a[0] + (a[1] / 2) * T + (a[2] / 3) * T2 + (a[3] / 4) * T3 +
                                             (a[4] / 5) * T4 + a[5] * rcpT))

    def production_rates(self, state, pressure_R, pressure_, Tref=1.0):
        """state: (N+1, S) species-major, row 0 = T/Tref, rows 1.. = Y_k.  Returns rates (N+1, S):
        row 0 heat release rate [W/m^3], rows 1.. mass production rates [kg/m^3/s]."""
        state = np.asarray(state, dtype=np.float64)
        T = Tref * state[0]
        rcpT = 1 / T
        lnT = np.log(T)
        T2, T3, T4 = T * T, T * T * T, T * T * T * T
        P, lnP = pressure_, math.log(pressure_)
        w, _, Mbar = self._decode(state)
        rho = pressure_R * rcpT * Mbar
        C = [w[k] * rho for k in range(self.N)]
        wdot = self._species_rates(lnT, T, T2, T3, T4, rcpT, P, lnP, C)
        out = np.empty_like(state)
        for k in range(self.N):
            out[k + 1] = self.M[k] * wdot[k]
        h = self._h_RT(T, T2, T3, T4, rcpT)
        acc = np.zeros_like(T)
        for k in range(self.N):
            acc = acc + wdot[k] * h[k]
        out[0] = (-R * T) * acc
        return out

    # ---- BK2 ----------------------------------------------------------------------------------
    def transport(self, state, pressure, Tref=1.0):
        """Returns (conductivity[S], viscosity[S], rhoD[N, S])."""
        assert self.tp is not None, 'no transport fixture for ' + self.name
        state = np.asarray(state, dtype=np.float64)
        N = self.N
        T = Tref * state[0]
        lnT = np.log(T)
        rcpT = 1 / T
        sqrT = np.sqrt(T)
        lnT2, lnT3, lnT4 = lnT * lnT, lnT * lnT * lnT, lnT * lnT * lnT * lnT
        w, rcpMbar, Mbar = self._decode(state)
        X = [w[k] * Mbar for k in range(N)]

        def quartic(P):
            return P[0] + P[1] * lnT + P[2] * lnT2 + P[3] * lnT3 + P[4] * lnT4

        # conductivity
        s1 = np.zeros_like(T)
        s2 = np.zeros_like(T)
        for k in range(N):
            lam = quartic(self.tp['conductivity'][k])
            s1 = s1 + X[k] * lam
            s2 = s2 + X[k] / lam
        cond = sqrT * (0.5 * (s1 + 1. / s2))

        # viscosity (Wilke)
        v = [quartic(self.tp['viscosity'][k]) for k in range(N)]
        sums = [np.zeros_like(T) for _ in range(N)]
        M = self.M
        for j in range(N):
            rj = 1. / v[j]
            for k in range(N):
                Va = np.sqrt(1 / np.sqrt(8) * 1 / np.sqrt(1. + M[k] / M[j]))
                Vb = Va * np.sqrt(np.sqrt(M[j] / M[k]))
                t = Va + Vb * v[k] * rj
                sums[k] = sums[k] + X[j] * (t * t)
        vis = np.zeros_like(T)
        for k in range(N):
            vis = vis + X[k] * (v[k] * v[k]) * (1. / sums[k])
        visc = sqrT * vis

        # mixture-averaged diffusion
        S = [None] * N
        tri = self.tp['diffusivity_lower']
        idx = 0
        for k in range(N):
            for j in range(k):
                D = 1 / quartic(tri[idx])
                idx += 1
                a = X[j] * D
                S[k] = a if S[k] is None else S[k] + a
                b = X[k] * D
                S[j] = b if S[j] is None else S[j] + b
        TsqrT = T * sqrT
        rho = pressure / R * rcpT * Mbar
        rhoD = np.empty((N, state.shape[1]))
        for k in range(N):
            Dkm = TsqrT * (Mbar - M[k] * X[k]) / (pressure * Mbar * S[k])
            rhoD[k] = rho * Dkm
        return cond, visc, rhoD

    # ---- thermo -------------------------------------------------------------------------------
    def thermo(self, state, pressure_R, Tref=1.0):
        """Returns (rho[S], cp[N, S] J/kg/K, rhoCp[S])."""
        state = np.asarray(state, dtype=np.float64)
        N = self.N
        T = Tref * state[0]
        rcpT = 1 / T
        T2, T3, T4 = T * T, T * T * T, T * T * T * T
        w, rcpMbar, Mbar = self._decode(state)
        rho = pressure_R * rcpT * Mbar
        self._sel = list(range(N))
        cpR = self._piecewise(T, lambda a: a[0] + a[1] * T + a[2] * T2 + a[3] * T3 + a[4] * T4)
        cp = np.empty((N, state.shape[1]))
        mean = np.zeros_like(T)
        for k in range(N):
            cp[k] = cpR[k] * R * self.rcpM[k]
            mean = mean + cpR[k] * w[k] * Mbar
        return rho, cp, rho * (mean * R * rcpMbar)


# ---- synthetic inputs shared by tests and bench (north star: T ~ U[300,2500] K, normalised Y) ----
def synthetic_states(n_species, n_states, seed=1234, T_lo=300., T_hi=2500., Tref=1.0):
    rng = np.random.default_rng(seed)
    st = np.empty((n_species + 1, n_states))
    st[0] = rng.uniform(T_lo, T_hi, n_states) / Tref
    Y = rng.uniform(0., 1., (n_species, n_states))
    st[1:] = Y / Y.sum(axis=0, keepdims=True)
    return st


def load_cantera_ci(path, n_species):
    """Parse a 13-line reference CI file (written by ci_data/generateCIdata.ipynb cells 9-11; read by
    benchmark/src/bk.cpp:271-374)."""
    lines = open(path).read().split('\n')
    vec = lambda s: np.array([float(x) for x in s.split()])[:n_species]
    d = dict(names=lines[0].split(), M=vec(lines[1]), T=float(lines[2]), p=float(lines[3]), X=vec(lines[4]),
             rho=float(lines[5]), cp_mole_e3=float(lines[6]), cp_k=vec(lines[7]), wdot=vec(lines[8]),
             hrr=float(lines[9]), conductivity=float(lines[10]), viscosity=float(lines[11]), rhoD=vec(lines[12]))
    d['X'] = d['X'] / d['X'].sum()
    Mbar = float((d['X'] * d['M']).sum())
    d['Y'] = d['X'] * d['M'] / Mbar
    d['Mbar'] = Mbar
    return d
