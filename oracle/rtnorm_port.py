"""
oracle/rtnorm_port.py -- TEST INFRASTRUCTURE (see oracle/__init__.py).

CPU restatement of the reference's truncated-normal sampler (Chopin 2011,
Mazet's table variant) with the random source injected.

Follows /root/reference/lib/rtnorm.py:
  * ``rtnorm``      -> lib/rtnorm.py:21-92  (standardise, draw, un-standardise)
  * ``rtstdnorm``   -> lib/rtnorm.py:95-223 (the four branches)
  * ``build_tables``-> regenerates the arrays stored at lib/rtnorm.py:227
    (``x``), :1230 (``yu``), :2233 (``ncell``) from Chopin's construction
    instead of copying 13k literals (the reference file is GPL-2, the tables
    are printed to 12 significant digits): N=4000 equal-area strips of the
    standard normal pdf, x[1954]=0, strip k has height
    yu[k] = max(pdf(x[k]), pdf(x[k+1])) and width A/yu[k]; the strip area A is
    fixed by x[4001] = xmax = 3.48672170399 (lib/rtnorm.py:102); ``ncell[i]`` is
    the index of the strip that contains (i - I0)/INVH.  Verified against the
    reference arrays in the build container: max|dx| = 5e-12, max rel dyu =
    5e-12 (= the printing precision), ncell identical
    (tests/test_rtnorm_tables.py; golden samples in tests/golden/).

The ``rng`` argument is any object with
    rng.rand(low=0.0) -> uniform in [low, 1)     (numpy ``uniform(low, 1.0)``)
    rng.randn()       -> standard normal
    rng.randi(lo, hi) -> integer in [lo, hi)     (numpy ``randint(lo, hi)``)
mirroring the three aliases at lib/rtnorm.py:17.
"""

import math
import numpy as np

# lib/rtnorm.py:101-102, 135-141
XMIN = -2.00443204036
XMAX = 3.48672170399
KMIN = 5
INVH = 1631.73284006
I0 = 3271
ALPHA = 1.837877066409345
N = 4000
YL0 = 0.053513975472
YLN = 0.000914116389555
K_ZERO = 1954            # x[K_ZERO] == 0.0 ; "elif k <= 1954" at lib/rtnorm.py:196,213
NCELL_LEN = 8961         # len(ncell) at lib/rtnorm.py:2233

_SQRT_2PI = math.sqrt(2.0 * math.pi)


def _pdf(t):
    return math.exp(-0.5 * t * t) / _SQRT_2PI


def _grid(area):
    x = [0.0] * (N + 2)
    for k in range(K_ZERO, N + 1):
        x[k + 1] = x[k] + area / _pdf(x[k])
    for k in range(K_ZERO - 1, -1, -1):
        x[k] = x[k + 1] - area / _pdf(x[k + 1])
    return x


_TABLES = None


def build_tables():
    """Returns (x[4002] f64, yu[4001] f64, ncell[8961] int64)."""
    global _TABLES
    if _TABLES is not None:
        return _TABLES
    lo, hi = 2.4448e-4, 2.4450e-4
    for _ in range(100):
        mid = 0.5 * (lo + hi)
        if _grid(mid)[N + 1] > XMAX:
            hi = mid
        else:
            lo = mid
    x = np.array(_grid(0.5 * (lo + hi)))
    x[K_ZERO] = 0.0
    yu = np.array([_pdf(x[k + 1]) if k < K_ZERO else _pdf(x[k])
                   for k in range(N + 1)])
    h = 1.0 / INVH
    pos = (np.arange(NCELL_LEN) - I0) * h
    # strip containing pos; the half-ulp guard resolves the one abscissa
    # (i = I0+1, pos == x[1955] to 1e-15) the way the reference table does.
    ncell = np.searchsorted(x, pos + 1e-9 * h, side='right') - 1
    ncell = np.clip(ncell, 0, N).astype(np.int64)
    _TABLES = (x, yu, ncell)
    return _TABLES


def rtnorm(a, b, mu=0., sigma=1., rng=None, tables=None):
    """lib/rtnorm.py:21-92 with size=1, probabilities=False. Returns a float."""
    mu = float(mu)
    sigma = float(sigma)
    a = float(a)
    b = float(b)
    if not mu == 0. or not sigma == 1.:
        a = (a - mu) / sigma
        b = (b - mu) / sigma
    r = rtstdnorm(a, b, rng, tables if tables is not None else build_tables())
    if not mu == 0. or not sigma == 1.:
        r = r * sigma + mu
    return r


def rtstdnorm(a, b, rng, tables):
    """lib/rtnorm.py:95-223."""
    x, yu, ncell = tables
    log = np.log          # the reference calls numpy's log/exp/floor (lib/rtnorm.py:18)
    exp = np.exp
    if a >= b:
        raise Exception('Truncated ndst in [a,b]: b MUST be greater than a.')
    elif abs(a) > abs(b):
        return -rtstdnorm(-b, -a, rng, tables)                 # :108-109
    elif a > XMAX:                                             # :112-124
        twoasq = 2 * a ** 2
        expab = exp(-a * (b - a)) - 1
        while True:
            z = log(1 + rng.rand(low=1E-15) * expab)
            e = -log(rng.rand(low=1E-15))
            if twoasq * e > z ** 2:
                break
        return a - z / a
    elif a < XMIN:                                             # :127-131
        while True:
            r = rng.randn()
            if (r >= a) and (r <= b):
                return r
    else:                                                      # :133-222
        i = int(I0 + np.floor(a * INVH))
        ka = int(ncell[i])
        if b >= XMAX:
            kb = N
        else:
            i = int(I0 + np.floor(b * INVH))
            kb = int(ncell[i])
        if abs(kb - ka) < KMIN:                                # :154-163
            twoasq = 2 * a ** 2
            expab = exp(-a * (b - a)) - 1
            while True:
                z = log(1 + rng.rand() * expab)
                e = -log(rng.rand())
                if twoasq * e > z ** 2:
                    break
            return a - z / a
        while True:                                            # :164-222
            k = rng.randi(ka, kb + 1)
            if k == N:
                lbound = x[-1]
                z = -log(rng.rand())
                e = -log(rng.rand())
                z = z / lbound
                if (z ** 2 <= 2 * e) and (z < b - lbound):
                    return lbound + z
            elif (k <= ka + 2) or (k >= kb and b < XMAX):
                sim = x[k] + (x[k + 1] - x[k]) * rng.rand()
                if (sim >= a) and (sim <= b):
                    simy = yu[k] * rng.rand()
                    if k == 0:
                        ylk = YL0
                    elif k == N:
                        ylk = YLN
                    elif k <= K_ZERO:
                        ylk = yu[k - 1]
                    else:
                        ylk = yu[k + 1]
                    if (simy < ylk) or (sim ** 2 + 2 * log(simy) + ALPHA < 0):
                        return sim
            else:
                u = rng.rand()
                simy = yu[k] * u
                d = x[k + 1] - x[k]
                if k == 1:
                    ylk = YL0
                elif k == N:
                    ylk = YLN
                elif k <= K_ZERO:
                    ylk = yu[k - 1]
                else:
                    ylk = yu[k + 1]
                if simy < ylk:
                    return x[k] + u * d * yu[k] / ylk
                sim = x[k] + d * rng.rand()
                if sim ** 2 + 2 * log(simy) + ALPHA < 0:
                    return sim
