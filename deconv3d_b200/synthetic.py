"""
Synthetic MUSE-shaped workloads (BASELINE.json configs 2-5).

Pure numpy on the host: builds the truth parameter maps, the FSF/LSF and the
noise; the noiseless cube itself is produced by the caller with the product's
own forward model (``Context.simulate``) so that no CPU convolution is needed.
"""
import numpy as np

from .instruments import MUSE
from .spread_functions import MoffatFieldSpreadFunction, GaussianFieldSpreadFunction


def halpha_truth(D=40, H=40, W=40):
    """cfg2 truth maps: a = 10 exp(-r^2 / (2*6^2)), c = D/2 + 6 tanh((x-(W-1)/2)/8), w = 2."""
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float64)
    r2 = (yy - (H - 1) / 2.) ** 2 + (xx - (W - 1) / 2.) ** 2
    a = 10.0 * np.exp(-r2 / (2. * 6. ** 2))
    c = D / 2. + 6.0 * np.tanh((xx - (W - 1) / 2.) / 8.)
    w = np.full((H, W), 2.0)
    return np.dstack([a, c, w])


def narrow_field_truth(D=64, H=256, W=256):
    """cfg4 truth maps (one oversized cube): the cfg2 galaxy stretched over the larger field."""
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float64)
    r2 = (yy - (H - 1) / 2.) ** 2 + (xx - (W - 1) / 2.) ** 2
    a = 10.0 * np.exp(-r2 / (2. * (0.15 * H) ** 2))
    c = D / 2. + 0.15 * D * np.tanh((xx - (W - 1) / 2.) / (0.2 * W))
    w = np.full((H, W), 2.5)
    return np.dstack([a, c, w])


def muse_wfm_instrument(fsf='moffat', fsf_size=13, fsf_fwhm=0.8, beta=2.5):
    """MUSE WFM: 0.2"/px, 1.25 A/channel; Moffat FWHM 0.8" beta 2.5 truncated to an odd stamp
    (the reference's Moffat image is cube-sized, lib/spread_functions.py:167) or the default
    Gaussian FSF; Gaussian LSF of FWHM 2.675 A."""
    if fsf == 'moffat':
        f = MoffatFieldSpreadFunction(fwhm=fsf_fwhm, beta=beta, size=fsf_size)
    else:
        f = GaussianFieldSpreadFunction(fwhm=fsf_fwhm)
    return MUSE(fsf=f)


def noise(shape, sigma=0.05, seed=1234):
    return sigma * np.random.default_rng(seed).standard_normal(shape)
