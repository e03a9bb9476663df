"""Host-side one-off helpers, mirror of the reference's lib/math_utils.py."""
import numpy as np

__all__ = ['median_clip', 'merge_where_nan']


def merge_where_nan(target, filler):
    """In-place: NaNs of ``target`` take the values of ``filler`` (lib/math_utils.py:4-13)."""
    np.copyto(target, filler, where=np.isnan(target))


def median_clip(data, clip_sigma=3., limit_ratio=1e-3, max_iterations=5):
    """Iteratively sigma-clipped median (lib/math_utils.py:16-57).
    Returns (median, sigma, number_of_iterations)."""
    values = np.asarray(data)
    values = values[np.isfinite(values)]
    centre = np.median(values)
    rounds = 0
    while True:
        rounds += 1
        previous = centre
        centre = np.median(values)
        spread = np.std(values)
        keep = np.nonzero(np.abs(values - centre) < clip_sigma * spread)
        if np.size(keep) > 0:
            values = values[keep]
        converged = abs(centre - previous) / abs(previous) < limit_ratio
        if converged or rounds >= max_iterations:
            break
    return np.median(values), np.std(values), rounds
