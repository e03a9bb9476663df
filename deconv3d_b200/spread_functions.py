"""
Field (spatial) and line (spectral) spread functions, mirror of the reference's
lib/spread_functions.py.  These run once per ``Run`` on the host and produce the
constants of the CUDA kernels: the FSF image (shared-memory stencil weights)
and the LSF vector (turned into the circular spectral kernel by the library).

Interfaces kept from the reference: ``FieldSpreadFunction.as_image(for_cube)``
(lib/spread_functions.py:23-36) and ``LineSpreadFunction.as_vector(for_cube)``
(:195-209).
"""
import math

import numpy as np

__all__ = ['FieldSpreadFunction', 'NoFieldSpreadFunction', 'ImageFieldSpreadFunction',
           'GaussianFieldSpreadFunction', 'MoffatFieldSpreadFunction',
           'LineSpreadFunction', 'VectorLineSpreadFunction', 'GaussianLineSpreadFunction',
           'MUSELineSpreadFunction']

_FWHM_PER_SIGMA = 2 * math.sqrt(2 * math.log(2))


def _centre(n):
    """Centre index used throughout the reference: (n-1)//2 - (n%2 - 1)
    (lib/spread_functions.py:107-109, 170-172, 251)."""
    return (n - 1) // 2 - (n % 2 - 1)


def _pixel_scale_arcsec(cube):
    return cube.get_step(1).to('arcsec').value


# ---- field spread functions ---------------------------------------------------

class FieldSpreadFunction(object):
    """Interface: ``as_image(for_cube)`` returns the 2-D FSF image."""

    def as_image(self, for_cube):
        raise NotImplementedError()


class NoFieldSpreadFunction(FieldSpreadFunction):
    """Image of ones with the spatial shape of the cube (lib/spread_functions.py:39-52)."""

    def as_image(self, for_cube):
        return np.ones(for_cube.shape[1:])


class ImageFieldSpreadFunction(FieldSpreadFunction):
    """A user-supplied FSF image, used as is (lib/spread_functions.py:55-68)."""

    def __init__(self, image_2d):
        self.image_2d = image_2d

    def as_image(self, for_cube):
        return self.image_2d

    def __str__(self):
        return "Custom Image PSF"


class GaussianFieldSpreadFunction(FieldSpreadFunction):
    """Elliptical Gaussian FSF (lib/spread_functions.py:71-131).

    fwhm [arcsec]; pa: clockwise angle from Y in degrees; ba: axis ratio b/a."""

    def __init__(self, fwhm=None, pa=0, ba=1.0):
        self.fwhm = fwhm
        self.pa = pa
        self.ba = ba

    def __str__(self):
        return 'Gaussian PSF :\n    fwhm = %s "\n    pa   = %s °\n    ba   = %s' \
            % (self.fwhm, self.pa, self.ba)

    def _elliptical_radius(self, xo, yo, x, y):
        """Radius in the rotated, squashed frame (lib/spread_functions.py:118-131)."""
        dx, dy = xo - x, yo - y
        t = np.radians(self.pa)
        u = dx * np.cos(t) - dy * np.sin(t)
        v = dx * np.sin(t) + dy * np.cos(t)
        return np.sqrt(u ** 2 + v ** 2 / self.ba ** 2)

    # the reference's name for the helper; kept for subclasses written against it
    _radius = _elliptical_radius

    def _grid(self, shape, xo, yo):
        if xo is None:
            xo = _centre(shape[1])
        if yo is None:
            yo = _centre(shape[0])
        y, x = np.indices(shape)
        return self._elliptical_radius(xo, yo, x, y)

    def as_image(self, for_cube, xo=None, yo=None):
        sigma_px = self.fwhm / _pixel_scale_arcsec(for_cube) / _FWHM_PER_SIGMA
        side = int(math.ceil(6. * sigma_px))          # +-3 sigma, forced odd (:101-104)
        if side % 2 == 0:
            side += 1
        r = self._grid((side, side), xo, yo)
        image = np.exp(-0.5 * (r / sigma_px) ** 2)
        return image / image.sum()


class MoffatFieldSpreadFunction(GaussianFieldSpreadFunction):
    """Moffat FSF (lib/spread_functions.py:134-189): (1 + (r/alpha)^2)^(-beta), given by
    ``fwhm`` or ``alpha`` [arcsec] and ``beta``.  Like the reference, the image has
    the spatial shape of the cube unless ``size`` (odd int or (h, w), an extension
    of this package) asks for a truncated, renormalised stamp -- the reference's
    cube-sized image must be odd-shaped to be accepted by ``Run`` (lib/run.py:210)."""

    def __init__(self, fwhm=None, alpha=None, beta=None, pa=None, ba=None, size=None):
        self.alpha = alpha
        self.beta = beta
        self.size = size
        GaussianFieldSpreadFunction.__init__(self, fwhm, 0. if pa is None else pa,
                                             1.0 if ba is None else ba)

    def __str__(self):
        return 'Moffat PSF :\n  fwhm = %s "\n  alpha = %s "\n  beta = %s\n  pa = %s °\n  ba = %s' \
            % (self.fwhm, self.alpha, self.beta, self.pa, self.ba)

    def as_image(self, for_cube, xo=None, yo=None):
        if self.size is None:
            shape = tuple(for_cube.shape[1:])
        elif np.isscalar(self.size):
            shape = (int(self.size), int(self.size))
        else:
            shape = (int(self.size[0]), int(self.size[1]))
        r = self._grid(shape, xo, yo)
        scale = _pixel_scale_arcsec(for_cube)
        if self.alpha is None:
            alpha_px = self.fwhm / scale / (2. * np.sqrt(2. ** (1. / self.beta) - 1))
        else:
            alpha_px = self.alpha / scale
        image = (1. + (r / alpha_px) ** 2) ** (-self.beta)
        return image / image.sum()


# ---- line spread functions ------------------------------------------------------

class LineSpreadFunction(object):
    """Interface: ``as_vector(for_cube)`` returns the LSF, one value per channel."""

    def as_vector(self, for_cube):
        raise NotImplementedError()


class VectorLineSpreadFunction(LineSpreadFunction):
    """A user-supplied LSF vector of the cube's spectral length, centred on
    ``(n-1)//2 - (n%2-1)`` (lib/spread_functions.py:212-228)."""

    def __init__(self, vector):
        self.vector = vector

    def as_vector(self, for_cube):
        return self.vector

    def __str__(self):
        return "Custom Vector LSF"


class GaussianLineSpreadFunction(LineSpreadFunction):
    """Gaussian LSF of given FWHM in micrometres (lib/spread_functions.py:231-277)."""

    def __init__(self, fwhm):
        self.fwhm = fwhm

    def __str__(self):
        return "Gaussian LSF : fwhm = %s µm \n" % self.fwhm

    def as_vector(self, for_cube):
        sigma_ch = self.fwhm / 2.35482 / for_cube.get_step(0).to('um').value
        depth = for_cube.shape[0]
        mid = _centre(depth)
        offsets = np.arange(depth) - mid
        if sigma_ch == 0:
            vector = np.zeros(depth)
            vector[mid] = 1.
        else:
            vector = self.gaussian(offsets, 0, sigma_ch)
        return vector / vector.sum()

    @staticmethod
    def gaussian(x, mu, sigma):
        """Un-normalised Gaussian (lib/spread_functions.py:263-277)."""
        return np.exp((x - mu) ** 2 / (-2. * sigma ** 2))


class MUSELineSpreadFunction(LineSpreadFunction):
    """LSF from MPDAF's MUSE model (lib/spread_functions.py:280-315).  Needs the
    ``mpdaf`` package; out of scope of the hot path, kept for API parity."""

    def __init__(self, model="qsim_v1"):
        self.model = model
        try:
            from mpdaf.MUSE import LSF
        except ImportError:
            raise ImportError("You need to install the mpdaf module "
                              "to use MUSELineSpreadFunction.")
        self.lsf = LSF(type=self.model)

    def __str__(self):
        return "MUSE LSF : model = '%s'" % self.model

    def as_vector(self, cube):
        depth = cube.shape[0]
        odd = depth if depth % 2 == 1 else depth + 1
        vector = self.lsf.get_LSF(lbda=cube.z_central * 1e4, step=cube.z_step * 1e4, size=odd)
        if depth % 2 == 0:
            vector = vector[:-1]
        return vector / vector.sum()
