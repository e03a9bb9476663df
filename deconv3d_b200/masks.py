"""Spatial masks, mirror of the reference's lib/masks.py."""
import numpy as np

from .cube import Cube

__all__ = ['above_percentile', 'read_hyperspectral_cube']


def read_hyperspectral_cube(cube):
    """Path or Cube -> Cube, with the reference's errors (lib/masks.py:6-14)."""
    if isinstance(cube, str):
        cube = Cube.from_fits(cube)
    if not isinstance(cube, Cube):
        raise TypeError("Provided cube is not a HyperspectralCube")
    if cube.is_empty():
        raise ValueError("Provided cube is empty")
    return cube


def above_percentile(cube, percentile=30):
    """0/1 image selecting the spaxels whose spectrally summed flux is at or above
    the given percentile (lib/masks.py:17-29)."""
    cube = read_hyperspectral_cube(cube)
    flux = np.nansum(cube.data, axis=0)
    threshold = np.nanpercentile(flux, percentile)
    return np.where(flux >= threshold, 1.0, np.where(flux < threshold, 0.0, flux))
