"""
Line-model plug-in interface, mirror of the reference's lib/line_models.py.

``LineModel`` keeps the reference's hook names and meaning
(lib/line_models.py:4-61): parameters(), gibbs_parameter_index(),
min_boundaries(runner), max_boundaries(runner), post_jump(runner, old, new),
modelize(runner, x, parameters).  ``SingleGaussianLineModel``
(lib/line_models.py:64-109) is the model the CUDA sweep evaluates natively:
its Gaussian is computed inside the kernels (csrc/d3d_kernels.cuh
``unit_gaussian``); the Python ``modelize`` below is only the user-facing
evaluation of a single line.
"""
import numpy as np

__all__ = ['LineModel', 'SingleGaussianLineModel', 'TiedGaussiansLineModel']


class LineModel(object):
    """Interface of a spectral-line model (lib/line_models.py:4-61)."""

    def parameters(self):
        """Unique names of the model parameters."""
        raise NotImplementedError()

    def gibbs_parameter_index(self):
        """Index of the parameter drawn by the Gibbs step, or None."""
        return None

    def min_boundaries(self, runner):
        raise NotImplementedError()

    def max_boundaries(self, runner):
        raise NotImplementedError()

    def post_jump(self, runner, old_parameters, new_parameters):
        """Optional hook run after the Cauchy jump; may mutate ``new_parameters``."""
        pass

    def modelize(self, runner, x, parameters):
        """Values of the line at abscissae ``x``."""
        raise NotImplementedError()


class SingleGaussianLineModel(LineModel):
    """One Gaussian line: amplitude a, centre c, RMS width w (channel units)."""

    #: the CUDA kernels implement exactly this model
    native = True

    def parameters(self):
        return ['a', 'c', 'w']

    def gibbs_parameter_index(self):
        return 0

    def min_boundaries(self, runner):
        return [0, 0, 0]

    def max_boundaries(self, runner):
        # lib/line_models.py:79-90: the FSF is normalised, so the amplitude of a line can
        # exceed the cube maximum by 1/max(fsf); centre within the cube, width up to its depth
        data = runner.cube.data
        peak = np.amax(runner.fsf)
        top = np.amax(data)
        if peak > 0:
            top = top / peak
        depth = data.shape[0]
        return [top, depth - 1, depth]

    def modelize(self, runner, x, parameters):
        return self.gaussian(x, parameters[0], parameters[1], parameters[2])

    @staticmethod
    def gaussian(x, a, c, w):
        """a * exp(-(x-c)^2 / (2 w^2))   (lib/line_models.py:98-109)."""
        x = np.asarray(x)
        return a * np.exp(-1. * (x - c) ** 2 / (2. * w ** 2))


class TiedGaussiansLineModel(SingleGaussianLineModel):
    """A multiplet of Gaussian lines tied to one another: component k sits ``offsets[k]`` channels
    from the first one and carries ``ratios[k]`` times its amplitude; all share the centre shift
    and the width.  Parameters stay (a, c, w) -- amplitude and centre of the FIRST component -- and
    the model stays linear in ``a``, which remains the Gibbs parameter (lib/run.py:456-519).
    Examples: [NII]6548 - Halpha - [NII]6583, the [OII]3726,3729 doublet.

    This is the family of user models (lib/line_models.py:4-61) the CUDA sweep evaluates natively
    (``d3d_set_line_model``); at most 4 components."""

    native = True

    def __init__(self, offsets, ratios):
        offsets = np.asarray(offsets, dtype=np.float64).ravel()
        ratios = np.asarray(ratios, dtype=np.float64).ravel()
        if offsets.shape != ratios.shape or not 1 <= offsets.size <= 4:
            raise ValueError("offsets and ratios must hold the same number (1..4) of components.")
        if ratios[0] == 0:
            raise ValueError("The first component carries the amplitude: its ratio must not be 0.")
        self.offsets = offsets - offsets[0]
        self.ratios = ratios / ratios[0]

    def native_components(self):
        """(offsets, ratios) handed to the device."""
        return self.offsets, self.ratios

    def modelize(self, runner, x, parameters):
        x = np.asarray(x)
        a, c, w = parameters[0], parameters[1], parameters[2]
        unit = np.exp(-1. * (x - c) ** 2 / (2. * w ** 2))
        for off, ratio in zip(self.offsets[1:], self.ratios[1:]):
            unit = unit + ratio * np.exp(-1. * (x - c - off) ** 2 / (2. * w ** 2))
        return a * unit
