"""Pure-species / binary transport properties from kinetic theory and their ln(T) polynomial fits.

Per mechanism (host side, once) this produces the coefficient tables the BK2 kernel evaluates:

    conductivity[k]   quartic in ln T of   lambda_k(T) / sqrt(T)
    viscosity[k]      quartic in ln T of   sqrt( mu_k(T) / sqrt(T) )
    diffusivity[j][k] quartic in ln T of   p*D_jk(T) / T^1.5      (or its reciprocal)

Physics and numerical recipe follow the reference so that the tables agree with the reference's to
rounding (reference kinetix/core/mix_transport.py:48-207; Cantera's GasTransport fits):
Lennard-Jones / Stockmayer reduced collision integrals Omega*(2,2) and A* from the Monchick-Mason
tables, quadratic Lagrange interpolation in ln T* across three table rows whose delta* dependence is
a degree-6 polynomial fit *evaluated with its first six coefficients only* (a Cantera quirk the
reference reproduces deliberately, mix_transport.py:133-137), 50 temperatures spanning the common
validity range of all species, weighted least squares with weights 1/|y| (general_utils.py:224-231).

Unlike the reference (which walks Python scalars: ~40 s for a 129-species mechanism) everything here is
vectorised over the 50 temperatures and over species pairs.
"""
import numpy as np
from numpy.polynomial import polynomial as npoly

from . import constants as const

_N_T = 50
_FIT_DEGREE = 4


class _CollisionIntegrals:
    """Omega*(2,2)(T*, delta*) and A*(T*, delta*)."""

    def __init__(self):
        data = const.load_collision_tables()
        self.ln_T_star = np.log(np.array(data['T_star']))          # 37 nodes, T* = 0.1 .. 100
        delta = np.array(data['delta_star'])
        self.omega22_rows = np.array(data['omega22'])               # 37 x 8
        self.a_star_rows = np.array(data['a_star'])                 # 39 x 8 (row 0 is T* -> 0)
        # degree-6 fit in delta* per table row, all weights -1 (constants.py:332-344)
        w = np.full(len(delta), -1.0)
        self.omega22_fit = np.array([npoly.polyfit(delta, row, 6, w=w) for row in self.omega22_rows])
        self.a_star_fit = np.array([npoly.polyfit(delta, row, 6, w=w) for row in self.a_star_rows])

    def _interp(self, row0, table, fit, ln_t_star, delta_star):
        """ln_t_star: array over temperatures; delta_star: scalar.  Returns array over T."""
        x = self.ln_T_star
        n = len(x)
        # first node strictly greater than ln T*; if none -> 0 (the reference's loop leaves start=0)
        gt = ln_t_star[:, None] < x[None, :]
        start = np.where(gt.any(axis=1), gt.argmax(axis=1), 0)
        i0 = np.maximum(start - 1, 0)
        i0 = np.where(i0 + 3 > n - 1, n - 4, i0)
        out = np.empty_like(ln_t_star)
        for t in range(len(ln_t_star)):
            i = int(i0[t])
            xs = x[i:i + 3]
            if delta_star == 0.0:
                ys = table[i + row0:i + row0 + 3, 0]
            else:
                powers = np.array([pow(delta_star, k) for k in range(6)])
                ys = [np.dot(P[:6], powers) for P in fit[i + row0:i + row0 + 3]]
            x0 = ln_t_star[t]
            L0 = ((x0 - xs[1]) * (x0 - xs[2])) / ((xs[0] - xs[1]) * (xs[0] - xs[2]))
            L1 = ((x0 - xs[0]) * (x0 - xs[2])) / ((xs[1] - xs[0]) * (xs[1] - xs[2]))
            L2 = ((x0 - xs[0]) * (x0 - xs[1])) / ((xs[2] - xs[0]) * (xs[2] - xs[1]))
            out[t] = L0 * ys[0] + L1 * ys[1] + L2 * ys[2]
        return out

    def omega22(self, ln_t_star, delta_star):
        return self._interp(0, self.omega22_rows, self.omega22_fit, ln_t_star, delta_star)

    def a_star(self, ln_t_star, delta_star):
        return self._interp(1, self.a_star_rows, self.a_star_fit, ln_t_star, delta_star)


class TransportFits:
    """Container: conductivity[N][5], viscosity[N][5], diffusivity[N][N][5] (symmetric)."""

    def __init__(self, conductivity, viscosity, diffusivity, reciprocal_diffusivity, T_min, T_max):
        self.conductivity = conductivity
        self.viscosity = viscosity
        self.diffusivity = diffusivity
        self.reciprocal_diffusivity = reciprocal_diffusivity
        self.T_min = T_min
        self.T_max = T_max


def _weighted_fit(ln_T, y):
    return npoly.polyfit(ln_T, y, deg=_FIT_DEGREE, w=1.0 / np.abs(y))


def fit_transport(mech, reciprocal_diffusivity=False):
    """Fit all pure-species and binary transport polynomials for ``mech`` (a Mechanism)."""
    sp = mech.species
    N = len(sp)
    kB, NA, pi = const.K_BOLTZMANN, const.N_AVOGADRO, np.pi
    eps = np.array([s.transport['well_depth'] for s in sp])
    sigma = np.array([s.transport['diameter'] for s in sp])
    mu = np.array([s.transport['dipole'] for s in sp])
    alpha = np.array([s.transport['polarizability'] for s in sp])
    rot = np.array([s.transport['rot_relax'] for s in sp])
    dof = np.array([s.transport['dof'] for s in sp])
    M = np.array([s.M for s in sp])

    T_min = max(s.T_ranges[0] for s in sp)
    T_max = min(s.T_ranges[-1] for s in sp)
    T = np.linspace(T_min, T_max, _N_T)
    ln_T = np.log(T)
    ci = _CollisionIntegrals()

    def xi(j, k):
        # induction correction for a polar / non-polar pair
        if (mu[j] > 0.) == (mu[k] > 0.):
            return 1.
        p, n = (j, k) if mu[j] != 0. else (k, j)
        return (1. + 1. / 4. * alpha[n] / (sigma[n] * sigma[n] * sigma[n]) *
                np.square(mu[p] / np.sqrt(4. * pi * const.EPSILON0 * eps[p] * (sigma[p] * sigma[p] * sigma[p]))) *
                np.sqrt(eps[p] / eps[n]))

    def pair(j, k):
        """Return (omega22(T), omega11(T), reduced mass, sigma_jk) for the pair."""
        x = xi(j, k)
        eps_jk = np.sqrt(eps[j] * eps[k]) * np.square(x)
        ln_t_star = np.log(T * kB / eps_jk)
        s_mean = (sigma[j] + sigma[k]) / 2.
        delta = 0.5 * mu[j] * mu[k] / (4. * pi * const.EPSILON0 * np.sqrt(eps[j] * eps[k]) * (s_mean * s_mean * s_mean))
        om22 = ci.omega22(ln_t_star, delta)
        om11 = om22 / ci.a_star(ln_t_star, delta)
        red_mass = M[j] / NA * M[k] / NA / (M[j] / NA + M[k] / NA)
        sigma_jk = s_mean * pow(x, -1. / 6.)
        return om22, om11, red_mass, sigma_jk

    # (kB T)^1.5 element by element: numpy's vectorised pow differs from the scalar pow the
    # reference calls by an ulp, and the fits are compared with the reference's to rounding.
    kT_32 = np.array([pow(kB * t, 3. / 2.) for t in T])

    def binary_pD(j, k):
        _, om11, red_mass, sigma_jk = pair(j, k)
        return (3. / 16. * np.sqrt(2. * pi / red_mass) * kT_32 /
                (pi * np.square(sigma_jk) * om11))

    visc_T = []     # pure-species viscosity mu_k(T)
    self_pD = []    # p*D_kk(T)
    for k in range(N):
        om22, om11, red_mass, sigma_kk = pair(k, k)
        visc_T.append(5. / 16. * np.sqrt(pi * M[k] / NA * kB * T) / (pi * np.square(sigma[k]) * om22))
        self_pD.append(3. / 16. * np.sqrt(2. * pi / red_mass) * kT_32 /
                       (pi * np.square(sigma_kk) * om11))

    def cp_R(k):
        s = sp[k]
        out = np.empty_like(T)
        for i, t in enumerate(T):
            a = s.nasa_lo if t < s.T_mid else s.nasa_hi
            out[i] = a[0] + a[1] * t + a[2] * t * t + a[3] * t * t * t + a[4] * t * t * t * t
        return out

    def F(t_star):
        return (1. + pow(pi, 3. / 2.) / np.sqrt(t_star) * (1. / 2. + 1. / t_star) +
                (1. / 4. * np.square(pi) + 2.) / t_star)

    conductivity = []
    viscosity = []
    for k in range(N):
        # Mason-Monchick: translational / rotational / vibrational contributions
        f_vib = M[k] / NA / (kB * T) * self_pD[k] / visc_T[k]
        t_star = T * kB / eps[k]
        A = 5. / 2. - f_vib
        B = rot[k] * F(298. * kB / eps[k]) / F(t_star) + 2. / pi * (5. / 3. * dof[k] + f_vib)
        f_rot = f_vib * (1. + 2. / pi * A / B)
        f_trans = 5. / 2. * (1. - 2. / pi * A / B * dof[k] / (3. / 2.))
        Cv = cp_R(k) - 5. / 2. - dof[k]
        lam = (visc_T[k] / (M[k] / NA)) * kB * (f_trans * 3. / 2. + f_rot * dof[k] + f_vib * Cv)
        conductivity.append(_weighted_fit(ln_T, lam / np.sqrt(T)))
        viscosity.append(_weighted_fit(ln_T, np.sqrt(visc_T[k] / np.sqrt(T))))

    diffusivity = np.zeros((N, N, _FIT_DEGREE + 1))
    TsqrtT = T * np.sqrt(T)
    for j in range(N):
        for k in range(j + 1):
            pD = self_pD[j] if j == k else binary_pD(j, k)
            y = TsqrtT / pD if reciprocal_diffusivity else pD / TsqrtT
            c = _weighted_fit(ln_T, y)
            diffusivity[j, k] = c
            diffusivity[k, j] = c
    return TransportFits(np.array(conductivity), np.array(viscosity), diffusivity,
                         reciprocal_diffusivity, T_min, T_max)
