"""
Minimal hyperspectral cube container + FITS primary-HDU I/O.

The reference delegates this to the external ``hyperspectral`` package
(``HyperspectralCube``; call sites lib/run.py:10,121,135,542-549,778,
lib/spread_functions.py:96,180,247, lib/instruments.py:142-151) which, like
astropy, carries no hot-path arithmetic.  This module provides the same duck
type with no third-party dependency: ``data``, ``shape``, ``meta``,
``is_empty()``, ``get_step(axis)``, ``from_fits``, ``to_fits``.
"""
import numpy as np

__all__ = ['Cube', 'HyperspectralCube', 'Quantity', 'read_fits', 'write_fits']

_DTYPES = {8: 'u1', 16: '>i2', 32: '>i4', 64: '>i8', -32: '>f4', -64: '>f8'}
_BITPIX = {'u1': 8, 'i2': 16, 'i4': 32, 'i8': 64, 'f4': -32, 'f8': -64}

# conversion factors to the units the spread functions ask for
_ANGLE_TO_ARCSEC = {'deg': 3600.0, 'degree': 3600.0, 'arcmin': 60.0, 'arcsec': 1.0, 'mas': 1e-3}
_LENGTH_TO_UM = {'angstrom': 1e-4, 'aa': 1e-4, 'nm': 1e-3, 'um': 1.0, 'micron': 1.0, 'mm': 1e3,
                 'm': 1e6}


class Quantity(object):
    """A number with a unit name; ``.to(unit).value`` like astropy's."""

    def __init__(self, value, unit):
        self.value = value
        self.unit = str(unit).strip()

    def to(self, unit):
        target = str(getattr(unit, 'name', unit)).strip()
        src, dst = self.unit.lower(), target.lower()
        for table in (_ANGLE_TO_ARCSEC, _LENGTH_TO_UM):
            if src in table and dst in table:
                return Quantity(self.value * table[src] / table[dst], target)
        raise ValueError("cannot convert '%s' to '%s'" % (self.unit, target))

    def __repr__(self):
        return '%r %s' % (self.value, self.unit)


def _parse_card_value(text):
    text = text.strip()
    if text.startswith("'"):
        end = text.find("'", 1)
        return text[1:end].rstrip()
    text = text.split('/')[0].strip()
    if text in ('T', 'F'):
        return text == 'T'
    try:
        return int(text)
    except ValueError:
        try:
            return float(text.replace('D', 'E'))
        except ValueError:
            return text


def read_fits(path):
    """(data, header) of the primary HDU of a FITS file; data in native byte order."""
    with open(path, 'rb') as f:
        raw = f.read()
    header, pos, ended = {}, 0, False
    while not ended:
        block = raw[pos:pos + 2880]
        if len(block) < 2880:
            raise ValueError('%s: truncated FITS header' % path)
        pos += 2880
        for i in range(0, 2880, 80):
            card = block[i:i + 80].decode('ascii', 'replace')
            key = card[:8].strip()
            if key == 'END':
                ended = True
                break
            if card[8:10] == '= ':
                header[key] = _parse_card_value(card[10:])
    naxis = int(header.get('NAXIS', 0))
    if naxis == 0:
        return None, header
    shape = tuple(int(header['NAXIS%d' % k]) for k in range(naxis, 0, -1))
    dt = np.dtype(_DTYPES[int(header['BITPIX'])])
    data = np.frombuffer(raw, dtype=dt, count=int(np.prod(shape)), offset=pos).reshape(shape)
    data = data.astype(dt.newbyteorder('='))
    if header.get('BSCALE', 1) != 1 or header.get('BZERO', 0) != 0:
        data = data * header.get('BSCALE', 1) + header.get('BZERO', 0)
    return data, header


def _card(key, value):
    if isinstance(value, bool):
        v = '%20s' % ('T' if value else 'F')
    elif isinstance(value, (int, np.integer)):
        v = '%20d' % value
    elif isinstance(value, (float, np.floating)):
        v = '%20s' % repr(float(value)).upper()
    else:
        v = "'%-8s'" % str(value)
    return ('%-8s= %s' % (key[:8], v)).ljust(80)[:80]


def write_fits(path, data, header=None, clobber=False):
    import os
    if os.path.exists(path) and not clobber:
        raise IOError("File '%s' already exists." % path)
    data = np.asarray(data)
    kind = data.dtype.kind + str(data.dtype.itemsize)
    if kind not in _BITPIX:
        data = data.astype('f8')
        kind = 'f8'
    cards = [_card('SIMPLE', True), _card('BITPIX', _BITPIX[kind]), _card('NAXIS', data.ndim)]
    for k in range(data.ndim):
        cards.append(_card('NAXIS%d' % (k + 1), data.shape[data.ndim - 1 - k]))
    skip = {'SIMPLE', 'BITPIX', 'NAXIS', 'EXTEND', 'BSCALE', 'BZERO'}
    for key, value in (header or {}).items():
        if key in skip or key.startswith('NAXIS'):
            continue
        cards.append(_card(key, value))
    cards.append('END'.ljust(80))
    head = ''.join(cards)
    head += ' ' * (-len(head) % 2880)
    body = data.astype(data.dtype.newbyteorder('>')).tobytes()
    body += b'\0' * (-len(body) % 2880)
    with open(path, 'wb') as f:
        f.write(head.encode('ascii'))
        f.write(body)


class Cube(object):
    """Duck type of ``hyperspectral.HyperspectralCube`` (z, y, x indexed data)."""

    # what MUSE.build_cube assumes when a bare array is given (lib/instruments.py:126-141)
    DEFAULT_STEPS = {'CDELT1': 5.5555555555555e-05, 'CDELT2': 5.5555555555555e-05, 'CDELT3': 1.25,
                     'CUNIT1': 'deg', 'CUNIT2': 'deg', 'CUNIT3': 'Angstrom'}

    def __init__(self, data=None, meta=None, **_ignored):
        self.data = None if data is None else np.asarray(data)
        self.meta = meta if meta is not None else {}

    @property
    def shape(self):
        return () if self.data is None else self.data.shape

    def is_empty(self):
        return self.data is None or self.data.size == 0

    def __array__(self, dtype=None, copy=None):
        return np.asarray(self.data, dtype=dtype)

    def _header(self):
        h = self.meta.get('fits') if isinstance(self.meta, dict) else None
        return h if h is not None else self.DEFAULT_STEPS

    def get_step(self, axis):
        """Step of an axis as a Quantity: 0 = spectral (z), 1 = y, 2 = x."""
        h = self._header()
        n = {0: 3, 1: 2, 2: 1}[axis]
        step = h.get('CDELT%d' % n, self.DEFAULT_STEPS['CDELT%d' % n])
        unit = h.get('CUNIT%d' % n, self.DEFAULT_STEPS['CUNIT%d' % n])
        return Quantity(abs(step) if axis else step, unit)

    @classmethod
    def from_fits(cls, path):
        data, header = read_fits(path)
        return cls(data=data, meta={'fits': header})

    def to_fits(self, path, clobber=False):
        write_fits(path, self.data, self._header() if 'fits' in self.meta else None, clobber)

    def __str__(self):
        return 'Cube%s' % (self.shape,)


HyperspectralCube = Cube
