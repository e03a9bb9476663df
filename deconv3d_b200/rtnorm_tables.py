"""
Host-side construction of the truncated-normal tables used by the device
sampler (csrc/d3d_rng.cuh ``rtstdnorm``).

The reference stores them as 13k literals (lib/rtnorm.py:227 ``x``, :1230 ``yu``,
:2233 ``ncell``) inside a GPL-2 file; they are NOT copied here but rebuilt from
Chopin's construction (N. Chopin, Stat Comput 21, 2011), which reproduces the
reference's arrays to their 12-digit printing precision and ``ncell`` exactly
(tests/test_oracle_golden.py, tests/test_host_api.py):

  * 4000 vertical strips of equal area A under the standard normal pdf,
    with x[1954] = 0, strip k spanning [x[k], x[k+1]] with upper bound
    yu[k] = max(pdf(x[k]), pdf(x[k+1])) and width A / yu[k];
  * A is the value for which the last edge x[4001] equals the reference's
    xmax = 3.48672170399 (lib/rtnorm.py:102);
  * ncell[i] = index of the strip containing (i - 3271) / 1631.73284006
    (I0, INVH at lib/rtnorm.py:136-137).
"""
import math

import numpy as np

N_STRIPS = 4000
ZERO_EDGE = 1954
X_RIGHT = 3.48672170399
INV_H = 1631.73284006
I_ZERO = 3271
N_CELL = 8961
_NORM = 1.0 / math.sqrt(2.0 * math.pi)

_cache = None


def _edges(area):
    x = [0.0] * (N_STRIPS + 2)
    for k in range(ZERO_EDGE, N_STRIPS + 1):           # rightwards: height at the left edge
        x[k + 1] = x[k] + area / (_NORM * math.exp(-0.5 * x[k] * x[k]))
    for k in range(ZERO_EDGE, 0, -1):                  # leftwards: height at the right edge
        x[k - 1] = x[k] - area / (_NORM * math.exp(-0.5 * x[k] * x[k]))
    return x


def tables():
    """(x float64[4002], yu float64[4001], ncell int32[8961])"""
    global _cache
    if _cache is None:
        lo, hi = 2.4448e-4, 2.4450e-4
        for _ in range(100):                           # bisection on the strip area
            mid = 0.5 * (lo + hi)
            if _edges(mid)[-1] > X_RIGHT:
                hi = mid
            else:
                lo = mid
        x = np.array(_edges(0.5 * (lo + hi)))
        x[ZERO_EDGE] = 0.0
        pdf = _NORM * np.exp(-0.5 * x * x)
        yu = np.maximum(pdf[:-1], pdf[1:])
        grid = (np.arange(N_CELL) - I_ZERO) / INV_H
        ncell = np.searchsorted(x, grid + 1e-9 / INV_H, side='right') - 1
        _cache = (x, yu, np.clip(ncell, 0, N_STRIPS).astype(np.int32))
    return _cache
