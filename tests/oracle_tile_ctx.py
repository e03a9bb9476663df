"""
Test helper: a CPU stand-in for ``deconv3d_b200._native.Context`` in the tiled coloured sweep
(``deconv3d_b200.dist.TiledSweeper``), built on the oracle.  One colour phase = the oracle's
loop (oracle/reference_port.run_chain, lib/run.py:344-537) over the owned sites of the class,
started from the current parameters with the Philox stream positioned at that sweep.  Test
infrastructure only.
"""
import numpy as np

from oracle import reference_port as port
from oracle import streams

RECORD_DOUBLES = 8


class _SweepShifted(streams.PhiloxStream):
    def __init__(self, seed, chain, shift):
        streams.PhiloxStream.__init__(self, seed, chain)
        self.shift = int(shift)

    def begin_site(self, sweep, site):
        streams.PhiloxStream.begin_site(self, sweep + self.shift, site)


class OracleTileCtx(object):
    device_records = False

    def __init__(self, data, var, fsf, lsf, mask, init, seed, tables=None):
        self.data, self.var, self.fsf, self.lsf = data, var, fsf, lsf
        self.mask = np.array(mask, dtype=float)
        self.params = np.array(init, dtype=float)
        self.seed, self.tables = seed, tables
        D, self.H, self.W = data.shape
        self.fh, self.fw = fsf.shape
        self.tile = (0, self.H, 0, self.W)
        self.lik = np.zeros((self.H, self.W))
        self.accepted = int(self.mask.sum())                     # lib/run.py:341
        self.iters = 1
        self.nly = -(-self.H // self.fh)
        self.nlx = -(-self.W // self.fw)

    def set_tile(self, y0, y1, x0, x1):
        self.tile = (y0, y1, x0, x1)

    def _mine(self, y, x):
        y0, y1, x0, x1 = self.tile
        return y0 <= y < y1 and x0 <= x < x1

    def record_slots(self):
        return self.nly * self.nlx

    def colour_begin(self, it, min_rate=0.0):
        self.iters = it + 1

    def colour_phase(self, it, cy, cx, rec=None):
        sites, slots = [], []
        for iy in range(self.nly):
            for ix in range(self.nlx):
                y, x = cy + iy * self.fh, cx + ix * self.fw
                if y < self.H and x < self.W and self._mine(y, x) and self.mask[y, x] == 1:
                    sites.append((y, x))
                    slots.append(iy * self.nlx + ix)
        if rec is not None:
            rec[:] = 0.0
            rec[:, 0] = -1.0
        if not sites:
            return
        trace = {}
        res = port.run_chain(self.data, self.fsf, self.lsf, _SweepShifted(self.seed, 0, it - 1),
                             mask=self.mask.copy(), variance_cube=self.var,
                             initial_parameters=self.params.copy(), max_iterations=2,
                             min_acceptance_rate=0.0, refresh_every=0, trace=trace,
                             rtnorm_tables=self.tables, site_order=sites)
        self.params = res['last_parameters']
        for (y, x), slot in zip(sites, slots):
            t = trace[(1, y, x)]
            self.lik[y, x] = t[0]
            self.accepted += int(bool(t[3]))
            if rec is not None:
                rec[slot] = [y * self.W + x, 0, self.params[y, x, 0], self.params[y, x, 1],
                             self.params[y, x, 2], t[0], float(bool(t[3])), 0.0]

    def apply_records(self, rec, n=None):
        rec = np.asarray(rec).reshape(-1, RECORD_DOUBLES)
        for r in rec:
            if r[0] < 0:
                continue
            y, x = divmod(int(r[0]), self.W)
            if self._mine(y, x):
                continue
            self.params[y, x] = r[2:5]
            self.lik[y, x] = r[5]
            self.accepted += int(r[6] != 0)

    def get_params(self):
        return self.params[None].copy()

    def get_likelihoods(self):
        return self.lik[None].copy()

    def chain_control(self):
        return (np.array([self.accepted], dtype=np.int64), np.array([self.iters], dtype=np.int64),
                np.array([1], dtype=np.int32))

    def forward(self):
        pass
