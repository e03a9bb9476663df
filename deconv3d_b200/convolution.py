"""
Spectral convolution routines, mirror of the reference's lib/convolution.py,
executed on the GPU.

``convolve_1d(data, psf, compute_fourier=True, axis=0) -> (conv, fftpsf)`` keeps
the reference signature and semantics (lib/convolution.py:89-120): circular
convolution on the power-of-two padded length with both operands placed at
offset ``half`` followed by an fftshift and a crop -- which in direct space is
    out[j] = sum_i data[i] * psf[((j - i + P/2) mod P) - half]     (0 <= index < n)
and wraps flux between the two ends of the spectrum whenever P - n is small
(n = 30, 32, 64...).  The CUDA kernel (csrc/d3d_kernels.cuh ``conv1d_kernel``)
evaluates exactly that sum; no FFT is involved on the device.

The second return value is, as in the reference, the real FFT of the padded PSF;
passing it back with ``compute_fourier=False`` is supported (the PSF is recovered
from it), so reference-style memoisation (lib/run.py:679-682) keeps working.
"""
import numpy as np

from . import _native

__all__ = ['convolve_1d', 'padding']

_default_ctx = None


def default_context():
    """Lazily created context on GPU 0 for the stateless helpers."""
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = _native.Context(device=0, dtype=_native.F64)
    return _default_ctx


def _padded_length(n):
    """2**len(binary_repr(n-1)) (lib/convolution.py:141-144)."""
    return 2 ** len(np.binary_repr(n - 1))


def _half(n):
    diff = _padded_length(n) - n
    return diff // 2 + 1 if diff & 1 else diff // 2


def padding(cube, axes=None):
    """Zero-pads ``cube`` to power-of-two lengths along ``axes`` (default [0, 1]);
    returns (padded, slices) where ``slices`` locate the data inside the padded
    array (lib/convolution.py:123-160)."""
    cube = np.asarray(cube)
    if axes is None:
        axes = [0, 1]
    axes = [int(a) for a in np.atleast_1d(axes)]
    shape = list(cube.shape)
    where = [slice(0, n) for n in cube.shape]
    for ax in axes:
        n = cube.shape[ax]
        shape[ax] = _padded_length(n)
        where[ax] = slice(_half(n), n + _half(n))
    padded = np.zeros(shape)
    padded[tuple(where)] = cube
    return padded, where


def convolve_1d(data, psf, compute_fourier=True, axis=0):
    """Convolves ``data`` with ``psf`` along ``axis`` on the GPU; see module docstring."""
    data = np.asarray(data, dtype=np.float64)
    axis = int(np.atleast_1d(axis)[0])
    n = data.shape[axis]
    P = _padded_length(n)
    h = _half(n)
    if compute_fourier:
        psf = np.asarray(psf, dtype=np.float64)
        if psf.shape != (n,):
            raise ValueError("psf must be a vector of the length of the convolved axis")
        padded = np.zeros(P)
        padded[h:h + n] = psf
        fftpsf = np.fft.rfft(padded)
        kernel = psf
    else:
        fftpsf = np.asarray(psf)
        kernel = np.fft.irfft(fftpsf.reshape(-1), n=P)[h:h + n]
    moved = np.moveaxis(data, axis, -1)
    out = default_context().conv1d(np.ascontiguousarray(moved), kernel)
    return np.moveaxis(out, -1, axis), fftpsf
