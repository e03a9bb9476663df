"""
oracle/streams.py -- TEST INFRASTRUCTURE (see oracle/__init__.py).

Random sources that can be injected into ``reference_port.run_chain`` and
``rtnorm_port.rtnorm``.  Every call site of randomness in the reference
(lib/run.py:313 initial parameters, :578 Cauchy jump, :435 acceptance;
lib/rtnorm.py:17 ``rand``/``randn``/``randi``) goes through one of these
methods, so the same port can reproduce

  * the reference itself          -> ``NumpyGlobalStream`` (the very same
    ``numpy.random`` calls in the same order => bit-identical chains for a
    given ``numpy.random.seed``), and
  * the device sampler            -> ``PhiloxStream`` ("d3d stream v1",
    oracle/philox.py), or ``ReplayStream`` fed with literal numbers.
"""

import math
import numpy as np

from . import philox

CIRCLE_4TH = np.pi / 2.      # lib/run.py:35


class NumpyGlobalStream(object):
    """The reference's own random calls on the global numpy state."""

    def begin_site(self, sweep, site):
        pass

    def init_uniforms(self, n):                       # lib/run.py:313
        return np.random.rand(n)

    def jump_uniforms(self, n):                       # lib/run.py:578
        return np.random.uniform(-CIRCLE_4TH, CIRCLE_4TH, size=n)

    def accept_uniform(self):                         # lib/run.py:435
        return np.random.rand()

    def rand(self, low=0.0):                          # lib/rtnorm.py:17 (uniform)
        return np.random.uniform(low=low)

    def randn(self):                                  # lib/rtnorm.py:17 (normal)
        return np.random.normal()

    def randi(self, lo, hi):                          # lib/rtnorm.py:17 (randint)
        return np.random.randint(low=lo, high=hi)


class _CountedStream(object):
    """Shared mapping from raw [0,1) draws to the reference's distributions."""

    def _next(self):
        raise NotImplementedError()

    def init_uniforms(self, n):
        return np.array([self._next() for _ in range(n)])

    def jump_uniforms(self, n):
        # numpy's uniform(low, high) is low + (high-low)*random_sample()
        lo, hi = -CIRCLE_4TH, CIRCLE_4TH
        return np.array([lo + (hi - lo) * self._next() for _ in range(n)])

    def accept_uniform(self):
        return self._next()

    def rand(self, low=0.0):
        return low + (1.0 - low) * self._next()

    def randn(self):
        # Box-Muller on two consecutive draws; 1-u keeps the log finite.
        u1 = self._next()
        u2 = self._next()
        return math.sqrt(-2.0 * math.log(1.0 - u1)) * math.cos(2.0 * math.pi * u2)

    def randi(self, lo, hi):
        return lo + int(math.floor(self._next() * (hi - lo)))


class PhiloxStream(_CountedStream):
    """d3d stream v1: draws addressed by (seed, chain, sweep, site, k)."""

    def __init__(self, seed, chain=0):
        self.seed = int(seed)
        self.chain = int(chain)
        self.sweep = 0
        self.site = 0
        self.k = 0
        self.max_k = 0

    def begin_site(self, sweep, site):
        self.sweep = int(sweep)
        self.site = int(site)
        self.k = 0

    def _next(self):
        u = philox.draw(self.seed, self.chain, self.sweep, self.site, self.k)
        self.k += 1
        self.max_k = max(self.max_k, self.k)
        return u


class ReplayStream(_CountedStream):
    """Raw [0,1) draws replayed from a per-site table: draws[sweep][site][k]
    (dict or nested array).  Raises when a site asks for more than supplied."""

    def __init__(self, draws):
        self.draws = draws
        self.cur = None
        self.k = 0

    def begin_site(self, sweep, site):
        self.cur = self.draws[sweep][site]
        self.k = 0

    def _next(self):
        u = float(self.cur[self.k])
        self.k += 1
        return u
