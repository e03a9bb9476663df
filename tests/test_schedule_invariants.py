"""
CPU checks of the two scheduling invariants the CUDA sweeps rely on (no GPU needed).

1. COLOURED mode (d3d_kernels.cuh / d3d_tile.cuh): the sites of one colour class
   (y mod fh, x mod fw) have pairwise DISJOINT FSF windows, for odd and even field sizes, masks
   and non-square FSFs, and the classes cover every masked site exactly once.  This is what makes
   "all sites of a class at once" equal to "one after the other" (lib/run.py:367-519 per site).

2. PIPELINED sequential mode (d3d_pipe.cuh): the window sums of site k taken from a residual that
   lacks the updates of the last L sites, plus the rank-one corrections through the static cross
   tables X_d, equal the sums taken from the up-to-date residual.  The identity is restated here
   in numpy with the oracle's own window arithmetic (lib/run.py:400-424, 456-519).
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import reference_port as port      # noqa: E402


def _window(y, x, fh, fw, H, W):
    fhh, fhw = (fh - 1) // 2, (fw - 1) // 2
    return max(y - fhh, 0), min(y + fhh + 1, H), max(x - fhw, 0), min(x + fhw + 1, W)


@pytest.mark.parametrize('H,W,fh,fw', [(9, 10, 7, 7), (13, 13, 13, 13), (14, 15, 23, 23), (40, 40, 13, 13),
                                       (5, 6, 13, 13), (17, 19, 5, 9), (30, 31, 3, 3), (1, 12, 7, 7)])
@pytest.mark.parametrize('masked', [False, True])
def test_colour_classes_have_disjoint_windows(H, W, fh, fw, masked):
    rs = np.random.RandomState(H * 100 + W)
    mask = (rs.rand(H, W) > 0.3).astype(float) if masked else np.ones((H, W))
    order = port.colour_class_order(mask, fh, fw)
    # every masked site exactly once
    assert sorted(order) == sorted((int(y), int(x)) for y, x in zip(*np.nonzero(mask == 1)))
    classes = {}
    for (y, x) in order:
        classes.setdefault((y % fh, x % fw), []).append((y, x))
    assert len(classes) <= min(fh, H) * min(fw, W)
    for sites in classes.values():
        cover = np.zeros((H, W), dtype=np.int32)
        for (y, x) in sites:
            y0, y1, x0, x1 = _window(y, x, fh, fw, H, W)
            cover[y0:y1, x0:x1] += 1
        assert cover.max() <= 1, 'two windows of one colour class overlap'


def _paste(F, y, x, H, W):
    """FSF image pasted at spaxel (y, x), clipped at the borders (lib/run.py:697-706)."""
    fh, fw = F.shape
    fhh, fhw = (fh - 1) // 2, (fw - 1) // 2
    out = np.zeros((H, W))
    y0, y1, x0, x1 = _window(y, x, fh, fw, H, W)
    out[y0:y1, x0:x1] = F[y0 - (y - fhh):y1 - (y - fhh), x0 - (x - fhw):x1 - (x - fhw)]
    return out


@pytest.mark.parametrize('L', [1, 2, 3])
@pytest.mark.parametrize('fh,fw', [(7, 7), (5, 9)])
def test_stale_sums_plus_cross_terms_equal_fresh_sums(L, fh, fw):
    rs = np.random.RandomState(7 + L)
    D, H, W = 11, 12, 16
    F = rs.rand(fh, fw)
    F /= F.sum()
    iv = 1.0 / (0.05 ** 2 * (1 + rs.rand(D, H, W)))
    e = rs.randn(D, H, W)
    y = 4
    row = [(y, x) for x in range(W)]                    # one run: consecutive sites of one row
    foot = [_paste(F, yy, xx, H, W) for (yy, xx) in row]
    # static cross tables X_d[k][z] = sum_v F_{k-d}(v) F_k(v) iv[z, v]   (xtab_kernel)
    X = np.zeros((L + 1, W, D))
    for k in range(W):
        for d in range(1, L + 1):
            if k - d >= 0:
                X[d, k] = np.einsum('yx,zyx->z', foot[k - d] * foot[k], iv)

    def h_of(res, k):                                    # h[z] = sum_v F_k(v) iv e   (header of d3d_kernels.cuh)
        return np.einsum('yx,zyx->z', foot[k], iv * res)

    coefs = []
    stale = e.copy()                                     # the window warps' residual: lags L sites
    fresh = e.copy()
    for k in range(W):
        # the window warps have applied the updates of sites <= k-L-1 when they form h0_k
        if k - L - 1 >= 0:
            stale += coefs[k - L - 1][:, None, None] * foot[k - L - 1][None]
        h0 = h_of(stale, k)
        corr = np.zeros(D)
        for d in range(1, L + 1):
            if k - d >= 0:
                corr += coefs[k - d] * X[d, k]
        np.testing.assert_allclose(h0 + corr, h_of(fresh, k), rtol=1e-11, atol=1e-9)
        # the two scalar sums of the decision, through the brackets warp X forms ahead of time
        T = rs.randn(D)
        direct = np.dot(T, h_of(fresh, k))
        br = np.dot(T, h0)
        for d in range(1, L + 1):
            if k - d >= 0:
                br += np.dot(T * X[d, k], coefs[k - d])
        assert abs(br - direct) <= 1e-10 * max(1.0, abs(direct))
        c = rs.randn(D)                                  # a*Lu_old - r*L_end of this site (lib/run.py:508-515)
        coefs.append(c)
        fresh += c[:, None, None] * foot[k][None]
