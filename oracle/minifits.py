"""
oracle/minifits.py -- TEST INFRASTRUCTURE (see oracle/__init__.py).

Minimal FITS primary-HDU reader and the duck-typed stand-in for the
``hyperspectral.HyperspectralCube`` the reference imports (lib/run.py:10; the
PyPI package and astropy are absent from this image).  Only what the
reference's hot path touches is provided: ``data``, ``shape``, ``meta``,
``is_empty()``, ``get_step(axis)`` (0 -> spectral step, 1 -> spatial step, as
used at lib/spread_functions.py:96,247) and ``from_fits``.
"""

import numpy as np

_BITPIX = {8: 'u1', 16: '>i2', 32: '>i4', 64: '>i8', -32: '>f4', -64: '>f8'}


def read_fits(path):
    """Returns (data ndarray in native byte order, header dict) of HDU 0."""
    raw = open(path, 'rb').read()
    header = {}
    pos = 0
    done = False
    while not done:
        block = raw[pos:pos + 2880]
        pos += 2880
        for i in range(0, 2880, 80):
            card = block[i:i + 80].decode('ascii', 'replace')
            key = card[:8].strip()
            if key == 'END':
                done = True
                break
            if card[8:10] != '= ':
                continue
            val = card[10:].split('/')[0].strip() if "'" not in card[10:] \
                else card[10:].strip()
            if val.startswith("'"):
                header[key] = val[1:val.index("'", 1)].strip()
            elif val in ('T', 'F'):
                header[key] = (val == 'T')
            else:
                try:
                    header[key] = int(val)
                except ValueError:
                    header[key] = float(val)
    naxis = header['NAXIS']
    shape = tuple(header['NAXIS%d' % (naxis - i)] for i in range(naxis))
    dt = np.dtype(_BITPIX[header['BITPIX']])
    n = int(np.prod(shape)) if naxis else 0
    data = np.frombuffer(raw, dtype=dt, count=n, offset=pos).reshape(shape)
    data = data.astype(dt.newbyteorder('='))
    if 'BSCALE' in header or 'BZERO' in header:
        data = data * header.get('BSCALE', 1.0) + header.get('BZERO', 0.0)
    return data, header


_TO = {('deg', 'arcsec'): 3600.0, ('arcsec', 'arcsec'): 1.0,
       ('Angstrom', 'um'): 1e-4, ('um', 'um'): 1.0, ('micron', 'um'): 1.0}


class Quantity(object):
    def __init__(self, value, unit):
        self.value = value
        self.unit = unit

    def to(self, unit):
        unit = getattr(unit, 'name', unit)
        return Quantity(self.value * _TO[(self.unit, unit)], unit)


class Cube(object):
    def __init__(self, data=None, meta=None, **_):
        self.data = data
        self.meta = meta if meta is not None else {}

    @property
    def shape(self):
        return self.data.shape

    def is_empty(self):
        return self.data is None

    def __array__(self, dtype=None, copy=None):
        return np.asarray(self.data, dtype=dtype)

    def get_step(self, axis):
        h = self.meta['fits']
        if axis == 0:
            return Quantity(h['CDELT3'], h.get('CUNIT3', 'Angstrom').strip())
        return Quantity(abs(h['CDELT2']), h.get('CUNIT2', 'deg').strip())

    @staticmethod
    def from_fits(path):
        data, header = read_fits(path)
        return Cube(data=data, meta={'fits': header})


MUSE_META = {                        # lib/instruments.py:126-141 (build_cube)
    'CDELT1': 5.5555555555555e-05, 'CDELT2': 5.5555555555555e-05, 'CDELT3': 1.25,
    'CRVAL1': 1.0, 'CRVAL2': 1.0, 'CRVAL3': 6564.0,
    'CRPIX1': 1.0, 'CRPIX2': 1.0, 'CRPIX3': 15.0,
    'CUNIT1': 'deg', 'CUNIT2': 'deg', 'CUNIT3': 'Angstrom',
    'CTYPE1': 'RA---TAN', 'CTYPE2': 'DEC--TAN',
}


def muse_cube(data):
    return Cube(data=data, meta={'fits': dict(MUSE_META)})
