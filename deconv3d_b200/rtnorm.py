"""
Truncated-normal sampler, mirror of the reference's lib/rtnorm.py ``rtnorm``
(Chopin 2011 / Mazet tables), executed on the GPU with the counter-based
stream of csrc/d3d_rng.cuh.

    rtnorm(a, b, mu=0., sigma=1., size=1, probabilities=False, seed=None)

returns an ndarray of ``size`` variates of N(mu, sigma^2) truncated to [a, b]
(lib/rtnorm.py:21-92); with ``probabilities=True`` also their densities.
"""
import os

import numpy as np

from . import _native, rtnorm_tables
from .convolution import default_context

__all__ = ['rtnorm']


def rtnorm(a, b, mu=0., sigma=1., size=1, probabilities=False, seed=None):
    a, b, mu, sigma = float(a), float(b), float(mu), float(sigma)
    if a >= b:
        raise Exception('Truncated ndst in [a,b]: b MUST be greater than a.')
    if seed is None:
        seed = int.from_bytes(os.urandom(8), 'little')
    ctx = default_context()
    r, _ = ctx.rtnorm_batch(np.full(size, a), np.full(size, b), np.full(size, mu),
                            np.full(size, sigma), seed=seed)
    if probabilities:
        from math import erf, sqrt, pi
        if not mu == 0. or not sigma == 1.:        # the bounds are standardised first (:74-76)
            a, b = (a - mu) / sigma, (b - mu) / sigma
        z = sqrt(pi / 2) * sigma * (erf(b / sqrt(2)) - erf(a / sqrt(2)))
        z = max(z, 1e-15)
        return r, np.exp(-(r - mu) ** 2 / 2 / sigma ** 2) / z
    return r
