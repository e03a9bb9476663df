"""
deconv3d_b200 -- B200-native (sm_100a) implementation of the per-iteration
likelihood hot path of irap-omp/deconv3d, behind the reference's Python API.

Facade identical to the reference's package ``__init__`` (its star-imports of
lib.instruments, lib.run, lib.spread_functions, lib.line_models, lib.masks,
__init__.py:10-14), so ``from deconv3d_b200 import Run, MUSE`` replaces
``from deconv3d import Run, MUSE``.
"""
__version__ = '0.1.0'

from .cube import Cube, HyperspectralCube                       # noqa: F401
from .instruments import *                                      # noqa: F401,F403
from .run import *                                              # noqa: F401,F403
from .spread_functions import *                                 # noqa: F401,F403
from .line_models import *                                      # noqa: F401,F403
from .masks import *                                            # noqa: F401,F403
from .convolution import convolve_1d, padding                   # noqa: F401
from .rtnorm import rtnorm                                      # noqa: F401
