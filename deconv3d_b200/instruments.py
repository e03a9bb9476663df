"""Instrument = one FSF + one LSF, mirror of the reference's lib/instruments.py."""
from .cube import Cube
from .spread_functions import (LineSpreadFunction, FieldSpreadFunction,
                               GaussianLineSpreadFunction, GaussianFieldSpreadFunction)

__all__ = ['Instrument', 'MUSE']


class Instrument(object):
    """Holder of the spread functions (lib/instruments.py:11-34)."""

    def __init__(self, lsf, fsf):
        if not isinstance(lsf, LineSpreadFunction):
            raise ValueError("lsf= MUST be an instance of LineSpreadFunction")
        if not isinstance(fsf, FieldSpreadFunction):
            raise ValueError("fsf= MUST be an instance of FieldSpreadFunction")
        self.lsf = lsf
        self.fsf = fsf

    def __str__(self):
        return "\nfsf = %s\nlsf = %s\n" % (self.fsf, self.lsf)


class MUSE(Instrument):
    """MUSE defaults: Gaussian LSF of FWHM 2.675e-4 um, Gaussian FSF of FWHM 1"
    (lib/instruments.py:103-119)."""

    def __init__(self, lsf=None, fsf=None, lsf_fwhm=0.0002675,
                 fsf_fwhm=1.0, fsf_pa=0., fsf_ba=1.0):
        if lsf is None:
            lsf = GaussianLineSpreadFunction(fwhm=lsf_fwhm)
        if fsf is None:
            fsf = GaussianFieldSpreadFunction(fwhm=fsf_fwhm, pa=fsf_pa, ba=fsf_ba)
        Instrument.__init__(self, lsf=lsf, fsf=fsf)

    def build_cube(self, data):
        """Wraps a bare array with MUSE WFM axis metadata: 0.2"/px, 1.25 A/channel
        (lib/instruments.py:121-151)."""
        header = {
            'CDELT1': 5.5555555555555e-05, 'CDELT2': 5.5555555555555e-05, 'CDELT3': 1.25,
            'CRVAL1': 1.0, 'CRVAL2': 1.0, 'CRVAL3': 6564.0,
            'CRPIX1': 1.0, 'CRPIX2': 1.0, 'CRPIX3': 15.0,
            'CUNIT1': 'deg', 'CUNIT2': 'deg', 'CUNIT3': 'Angstrom',
            'CTYPE1': 'RA---TAN', 'CTYPE2': 'DEC--TAN',
        }
        return Cube(data=data, meta={'fits': header})
