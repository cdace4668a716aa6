#!/usr/bin/env python3
"""Derive the polynomial coefficients used by kinetix_b200/csrc/kx_math.cuh (run once, results pasted
into the header).  mpmath Chebyshev interpolation (near-minimax) at 60 digits, then rounded to double.

  exp:  e^r           on |r| <= ln2/2            (no table)          degree 10 / 11
  expt: e^r           on |r| <= ln2/64           (32-entry table)    degree 5 / 6
  log:  atanh-series  log(m) = 2 s (1 + s^2 q(s^2)),  s=(m-1)/(m+1), m in [sqrt(.5), sqrt(2))
"""
import mpmath as mp

mp.mp.dps = 60


def cheb_poly(f, a, b, deg):
    """power-basis coefficients (in x) of the Chebyshev interpolant of f on [a,b]"""
    n = deg + 1
    nodes = [mp.cos(mp.pi * (k + mp.mpf(1) / 2) / n) for k in range(n)]
    xs = [(a + b) / 2 + (b - a) / 2 * t for t in nodes]
    A = mp.matrix(n, n)
    y = mp.matrix(n, 1)
    for i, x in enumerate(xs):
        for j in range(n):
            A[i, j] = x ** j
        y[i] = f(x)
    c = mp.lu_solve(A, y)
    return [c[i] for i in range(n)]


def max_rel_err(f, coefs, a, b, n=4001):
    worst = 0
    dc = [float(c) for c in coefs]
    for i in range(n):
        x = float(a + (b - a) * i / (n - 1))
        v = 0.0
        for c in dc[::-1]:
            v = v * x + c
        ex = f(mp.mpf(x))
        if ex != 0:
            worst = max(worst, abs((mp.mpf(v) - ex) / ex))
    return float(worst)


if __name__ == '__main__':
    ln2 = mp.log(2)
    for deg in (10, 11, 12):
        c = cheb_poly(mp.exp, -ln2 / 2, ln2 / 2, deg)
        print('exp deg', deg, 'err', max_rel_err(mp.exp, c, -ln2 / 2, ln2 / 2))
        print('  ', ', '.join(repr(float(x)) for x in c))
    for deg in (5, 6):
        h = ln2 / 64
        c = cheb_poly(mp.exp, -h, h, deg)
        print('expt32 deg', deg, 'err', max_rel_err(mp.exp, c, -h, h))
        print('  ', ', '.join(repr(float(x)) for x in c))
    for deg in (4, 5):
        h = ln2 / 128
        c = cheb_poly(mp.exp, -h, h, deg)
        print('expt64 deg', deg, 'err', max_rel_err(mp.exp, c, -h, h))
        print('  ', ', '.join(repr(float(x)) for x in c))
    # log: q(z), z = s^2 in [0, smax^2], with atanh(s)/s = 1 + z/3 + z^2/5 + ... = 1 + z*q(z)
    smax = (mp.sqrt(2) - 1) / (mp.sqrt(2) + 1)
    zmax = smax ** 2

    def q(z):
        if z == 0:
            return mp.mpf(1) / 3
        s = mp.sqrt(z)
        return (mp.atanh(s) / s - 1) / z
    for deg in (6, 7, 8):
        c = cheb_poly(q, mp.mpf(0), zmax, deg)
        # error of full log relative: 2 s (1 + z q) vs atanh
        worst = 0
        dc = [float(x) for x in c]
        for i in range(1, 2001):
            s = float(smax) * i / 2000
            z = s * s
            v = 0.0
            for cc in dc[::-1]:
                v = v * z + cc
            approx = 2 * s * (1 + z * v)
            ex = 2 * mp.atanh(mp.mpf(s))
            worst = max(worst, abs((mp.mpf(approx) - ex) / ex))
        print('log q deg', deg, 'err', float(worst))
        print('  ', ', '.join(repr(float(x)) for x in c))
    print('table 2^(j/32):')
    print(', '.join(repr(float(mp.mpf(2) ** (mp.mpf(j) / 32))) for j in range(32)))
