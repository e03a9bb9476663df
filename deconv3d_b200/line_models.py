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

__all__ = ['LineModel', 'SingleGaussianLineModel']


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
