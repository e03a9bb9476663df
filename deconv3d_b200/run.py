# coding=utf-8
"""
``Run`` -- drop-in for the reference's MCMC runner (lib/run.py ``class Run``),
with the Metropolis-Hastings-within-Gibbs sweep executed on a B200 by
libdeconv3d_b200.so.

Same constructor signature, same attributes and methods, same exceptions as
lib/run.py:95-109 / 553-926.  What changes underneath:

  * the residual cube, inverse variance and parameter maps live on the GPU for
    the whole run; only the saved chain rows come back;
  * the H*W full-cube ``contributions`` array (lib/run.py:285-288) does not
    exist: every contribution is rank-1, FSF (x) (LSF * line), and is recomputed
    from the three parameters inside the kernels;
  * random numbers come from a counter-based Philox stream (``seed=``) instead of
    the global ``numpy.random`` state, so a run is reproducible;
  * extensions, all keyword-only and defaulting to the reference behaviour:
    ``mode`` ('sequential' = the reference's row-major update order,
    'coloured' = colour classes of non-overlapping windows updated concurrently),
    ``dtype`` ('float64' | 'float32' storage), ``seed``, ``n_chains``, ``device``,
    ``first_chain_id``.

There is no CPU fallback: without the CUDA library / a GPU, ``Run`` raises.
"""
import logging
import math
import os
from os.path import splitext

import numpy as np

from . import _native, rtnorm_tables
from .cube import Cube
from .instruments import Instrument
from .line_models import LineModel, SingleGaussianLineModel
from .math_utils import median_clip
from .convolution import convolve_1d

__all__ = ['Run', 'logger']

logging.basicConfig(level=logging.INFO)
logger = logging.getLogger('deconv3d')

REFRESH_EVERY = 1000          # lib/run.py:525: residual recomputed every 1000 iterations
_CHUNK = 1000                 # iterations per native call (log line + early-exit check)


def _host_rows(shape, zeroed):
    """Host array for chain rows.  Page-locked (through torch's caching host allocator, so that a
    second Run re-uses the block): the rows of a sweep travel device -> host by DMA at PCIe speed
    while the next sweeps compute, instead of through the driver's bounce buffer into freshly
    faulted pages (272 MB per 20 sweeps of 256 chains at the reference's keep_one_in=1).  The
    numpy array keeps the tensor alive.  Falls back to ordinary memory when pinning fails."""
    n = int(np.prod(shape))
    if n * 8 >= (1 << 20) and not os.environ.get('D3D_NO_PINNED_ROWS'):
        try:
            import torch
            t = (torch.zeros if zeroed else torch.empty)(shape, dtype=torch.float64, pin_memory=True)
            return t.numpy()
        except (RuntimeError, ImportError):
            pass
    return np.zeros(shape)


class Run(object):
    """
    Deconvolves the emission-line kinematics of a hyperspectral cube.

    cube: str | Cube
        Path to a FITS file, or a ``Cube`` (duck type of HyperspectralCube).
    instrument: Instrument
        e.g. ``MUSE()``; provides the FSF image and the LSF vector.
    mask: ndarray | str
        (H, W) image of 0/1; only spaxels at 1 are deconvolved.  Mutated in place
        like in the reference (NaN spaxels are set to 0).
    variance: Cube | ndarray | str | None
        Variance cube of the shape of the data; None = scalar guessed by sigma
        clipping a corner of the cube (lib/run.py:186-192).
    model: LineModel class or instance (default SingleGaussianLineModel)
    initial_parameters: ndarray | str | None
        (H, W, P) array or a ``.npy`` path; None = uniform in the boundaries.
    jump_amplitude: float | ndarray
        Cauchy jump scale(s); the Gibbs parameter's entry is forced to 0.
    gibbs_apriori_variance: float | None   (default: a_max**2)
    max_iterations, keep_one_in, write_every, min_acceptance_rate:
        as in the reference (lib/run.py:82-92, 344-359).
    """

    def __init__(self, cube, instrument=None, mask=None, variance=None,
                 model=SingleGaussianLineModel, initial_parameters=None,
                 jump_amplitude=0.1, gibbs_apriori_variance=None,
                 max_iterations=100000, keep_one_in=1, write_every=10000,
                 min_acceptance_rate=0.01, **extensions):
        mode = extensions.pop('mode', 'sequential')
        dtype = extensions.pop('dtype', 'float64')
        seed = extensions.pop('seed', None)
        n_chains = int(extensions.pop('n_chains', 1))
        device = int(extensions.pop('device', 0))
        first_chain_id = int(extensions.pop('first_chain_id', 0))
        chain_on_device = bool(extensions.pop('chain_on_device', False))
        if extensions:
            raise TypeError("Unknown arguments: %s" % ', '.join(sorted(extensions)))
        if mode not in ('sequential', 'coloured'):
            raise ValueError("mode= must be 'sequential' or 'coloured'")
        if dtype not in ('float64', 'float32'):
            raise ValueError("dtype= must be 'float64' or 'float32'")

        assert keep_one_in > 0, "keep_one_in= MUST be a positive integer"         # :112-114
        assert write_every > 0, "write_every= MUST be a positive integer"
        assert max_iterations > 0, "max_iterations= MUST be a positive integer"
        assert n_chains > 0, "n_chains= MUST be a positive integer"

        self.logger = logger
        self.mode = mode
        self.seed = int.from_bytes(os.urandom(8), 'little') if seed is None else int(seed)

        # -- inputs (lib/run.py:119-211) --------------------------------------------
        self.cube = self._load_cube(cube)
        data = self.cube.data
        signal_max = np.max(data)                                                  # :140-143
        assert signal_max > 1e-10, \
            "The input cube has data that is too small and will cause " \
            "numerical instability, infinite loops, or worse : bad science."
        shape = data.shape
        depth, height, width = shape

        self.mask = self._load_mask(mask, data)
        spaxels_count = int(np.sum(self.mask == 1))
        n_saved = int(math.ceil(max_iterations / float(keep_one_in)))              # :168

        self.variance_cube, variance_scalar = self._load_variance(variance, data)
        self.error_cube = np.sqrt(self.variance_cube)                              # :200

        if not isinstance(instrument, Instrument):                                 # :203-204
            raise TypeError("Provided instrument is not an Instrument")
        self.instrument = instrument
        self.lsf = self.instrument.lsf.as_vector(self.cube)                        # :208
        self.fsf = np.asarray(self.instrument.fsf.as_image(self.cube), dtype=np.float64)
        if self.fsf.shape[0] % 2 == 0 or self.fsf.shape[1] % 2 == 0:               # :210-211
            raise ValueError("FSF *must* be of odd dimensions")

        # -- model (lib/run.py:226-265) ---------------------------------------------
        if isinstance(model, LineModel):
            self.model = model
        else:
            self.model = model()
            if not isinstance(self.model, LineModel):
                raise TypeError("Provided model is not a LineModel")
        self._require_native_model()

        min_boundaries = np.array(self.model.min_boundaries(self), dtype=np.float64)
        max_boundaries = np.array(self.model.max_boundaries(self), dtype=np.float64)
        names = self.model.parameters()
        self.logger.info("Min boundaries : %s" % dict(zip(names, min_boundaries)))
        self.logger.info("Max boundaries : %s" % dict(zip(names, max_boundaries)))
        if (min_boundaries > max_boundaries).any():
            raise ValueError("Boundaries are inconsistent: min > max.")
        n_params = len(names)

        jumping_amplitude = np.ones(n_params) * np.array(jump_amplitude)           # :251-252
        gpi = self.model.gibbs_parameter_index()
        self.logger.info("MH within Gibbs enabled for parameter `%s`." % names[gpi])
        jumping_amplitude[gpi] = 0                                                 # :262
        if gibbs_apriori_variance is None:
            gibbs_apriori_variance = float(max_boundaries[gpi] ** 2)               # :265
        self.min_boundaries = min_boundaries
        self.max_boundaries = max_boundaries
        self.jumping_amplitude = jumping_amplitude
        self.gibbs_apriori_variance = float(gibbs_apriori_variance)

        # -- chain storage (lib/run.py:267-281) ---------------------------------------
        try:
            if chain_on_device:
                # the chain stays in HBM; only the posterior summary and the final
                # cubes come back (``run.chain`` copies it to the host on first access)
                import torch
                tdev = torch.device('cuda', device)
                chains = torch.zeros((n_chains, n_saved, height, width, n_params),
                                     dtype=torch.float64, device=tdev)
                likelihoods = torch.zeros((n_chains, n_saved, height, width),
                                          dtype=torch.float64, device=tdev)
            else:
                # every kept row of every chain is written by the device unless a chain stops on
                # min_acceptance_rate: only then do the arrays need the zeros of lib/run.py:270-271
                may_stop = min_acceptance_rate > 0
                chains = _host_rows((n_chains, n_saved, height, width, n_params), may_stop)
                likelihoods = _host_rows((n_chains, n_saved, height, width), may_stop)
        except (MemoryError, RuntimeError):
            self.logger.error("Not enough RAM available for that many iterations. "
                              "Use a higher value in the keep_one_in= parameter.")
            return

        # -- device problem -----------------------------------------------------------
        self._ctx = ctx = _native.Context(
            device=device, dtype=_native.F64 if dtype == 'float64' else _native.F32)
        ctx.set_rtnorm_tables(*rtnorm_tables.tables())
        ctx.set_rng(self.seed, first_chain_id)
        ctx.set_problem(
            data, variance_scalar if variance_scalar is not None else self.variance_cube,
            self.fsf, self.lsf, min_boundaries, max_boundaries, jumping_amplitude,
            self.gibbs_apriori_variance, mask=self.mask, chains_per_cube=n_chains)
        if self._native_components is not None:
            ctx.set_line_model(*self._native_components)

        # -- initial parameters (lib/run.py:293-314) ------------------------------------
        if initial_parameters is not None:
            if isinstance(initial_parameters, str):
                initial_parameters = np.load(initial_parameters)
            initial_parameters = np.array(initial_parameters)
            ip_shape = initial_parameters.shape
            if len(ip_shape) < 2 or ip_shape[0] != height or ip_shape[1] != width:
                raise ValueError(
                    "Initial params MUST have (%d, %d) shape, got %s."
                    % (height, width, str(tuple(ip_shape[:2]))))
            first_row = np.ascontiguousarray(
                np.broadcast_to(initial_parameters, (n_chains, height, width, n_params)), dtype=np.float64)
            ctx.set_params(first_row)
        else:
            ctx.init_params_uniform()
            first_row = ctx.get_params()
        if chain_on_device:
            chains[:, 0] = torch.from_numpy(first_row).to(tdev)
            # torch works on its own (legacy default) stream, the library on a non-blocking one:
            # order the two explicitly before the library touches the tensors
            torch.cuda.synchronize(tdev)
        else:
            chains[:, 0] = first_row

        # -- sweeps (lib/run.py:316-537) --------------------------------------------------
        self.logger.info("Iteration #1")
        ctx.forward(write_err=True)                     # err_old = data - sim, :317-334
        native_mode = _native.SEQ_EXACT if mode == 'sequential' else _native.COLOURED
        accepted = np.full(n_chains, spaxels_count, dtype=np.int64)
        reached = np.ones(n_chains, dtype=np.int64)
        cur = 1
        self.elapsed_ms = 0.0
        while cur < max_iterations:
            n = min(_CHUNK - (cur % _CHUNK) if cur % _CHUNK else _CHUNK, max_iterations - cur)
            accepted, reached, ms = ctx.sweep(
                cur, n, mode=native_mode, keep_one_in=keep_one_in,
                refresh_every=REFRESH_EVERY, min_acceptance_rate=min_acceptance_rate,
                chain_out=chains, lik_out=likelihoods)
            self.elapsed_ms += ms
            cur += n
            rate = float(accepted[0]) / float(max(1, spaxels_count) * max(1, reached[0] - 1))
            self.logger.info("Iteration #%d / %d, %2.0f%%" % (reached[0], max_iterations, 100 * rate))
            if (reached < cur).all():                   # every chain hit min_acceptance_rate
                break
        self.accepted_count = accepted
        self.iterations_done = reached

        # rows of spaxels outside the mask are never written by the sweep (the reference
        # leaves them uninitialised, lib/run.py:270); keep the initial values there
        off = self.mask != 1
        if off.any() and n_saved > 1:
            if chain_on_device:
                off_t = torch.from_numpy(off).to(tdev)
                chains[:, 1:, off_t] = chains[:, :1, off_t]
                torch.cuda.synchronize(tdev)            # before d3d_chain_mean reads the rows
            else:
                chains[:, 1:, off] = chains[:, :1, off]

        # -- outputs (lib/run.py:539-549) ----------------------------------------------------
        if chain_on_device:
            self._chains_device = chains
            self._likelihoods_device = likelihoods
        else:
            self.chains = chains
            self.chain = chains[0]
            self.all_likelihoods = likelihoods
            self.likelihoods = likelihoods[0]
        self.parameters = self.extract_parameters()
        self.convolved_cube = Cube(data=self.simulate_convolved(shape, self.parameters),
                                   meta=self.cube.meta)
        self.clean_cube = Cube(data=self.simulate_clean(shape, self.parameters),
                               meta=self.cube.meta)

    # -- input plumbing ---------------------------------------------------------------------

    @staticmethod
    def _load_cube(cube):
        if isinstance(cube, str):                                                  # :120-121
            cube = Cube.from_fits(cube)
        if not isinstance(cube, Cube):
            raise TypeError("Provided cube is not a HyperspectralCube")            # :134
        if cube.is_empty():
            raise ValueError("Provided cube is empty")                             # :136
        return cube

    @staticmethod
    def _load_mask(mask, data):
        if mask is None:                                                           # :153-157
            mask = np.ones(data.shape[1:])
        if isinstance(mask, str):
            from .cube import read_fits
            mask = read_fits(mask)[0]
        mask[np.isnan(np.sum(data, 0))] = 0                                        # :162
        return mask

    def _load_variance(self, variance, data):
        """Returns (variance_cube, scalar_or_None); lib/run.py:171-198."""
        scalar = None
        if variance is not None:
            if isinstance(variance, str):
                variance = Cube.from_fits(variance)
            if isinstance(variance, Cube):
                if variance.data is None:
                    self.logger.warning("Provided variance cube is empty")
                self.logger.info("Using provided variance : %s" % variance)
                self.logger.info("Replacing zeros in the variance cube by 1e12")
                cube = np.where(variance.data == 0.0, 1e12, variance.data)
            elif isinstance(variance, np.ndarray):
                cube = variance
            else:
                raise TypeError("Provided variance is not a Cube")
        else:
            corner = np.copy(data[2:-2, 2:-4, 2:4])
            _, sigma, _ = median_clip(corner, 2.5)
            if sigma == 0:
                sigma = 1e-20
            scalar = float(sigma ** 2)
            cube = np.ones(data.shape) * sigma ** 2
        if cube.shape != data.shape:
            raise ValueError("Provided variance has not the correct shape."
                             "Expected %s, got %s" % (str(data.shape), str(cube.shape)))
        if scalar is not None and np.isnan(data).any():
            # NaN voxels must drop out of the chi2 sums (nansum, lib/run.py:24-27, 420-425): that
            # needs per-voxel weights, which the scalar fast path does not carry
            scalar = None
        return cube, scalar

    def _require_native_model(self):
        """The kernels evaluate the tied-multiplet family of line models (one Gaussian =
        SingleGaussianLineModel, several = TiedGaussiansLineModel; ``d3d_set_line_model``) with
        the Cauchy jump and the row-major masked iteration.  Anything else would need a CPU path,
        which this package does not have -- fail loudly instead (SURVEY.md section 7 'Hard
        parts')."""
        from .line_models import TiedGaussiansLineModel
        m = type(self.model)
        single = (m.modelize is SingleGaussianLineModel.modelize
                  and getattr(m, 'gaussian') is SingleGaussianLineModel.gaussian)
        tied = isinstance(self.model, TiedGaussiansLineModel) and \
            m.modelize is TiedGaussiansLineModel.modelize and \
            m.native_components is TiedGaussiansLineModel.native_components
        native = ((single or tied)
                  and m.post_jump is LineModel.post_jump
                  and self.model.gibbs_parameter_index() == 0
                  and list(self.model.parameters()) == ['a', 'c', 'w'])
        if not native:
            raise NotImplementedError(
                "deconv3d_b200 evaluates SingleGaussianLineModel and TiedGaussiansLineModel on the "
                "GPU; custom modelize/post_jump/Gibbs parameters are not supported (no CPU fallback).")
        self._native_components = self.model.native_components() if tied else None
        if type(self).jump_from is not Run.jump_from or \
                type(self).spaxel_iterator is not Run.spaxel_iterator:
            raise NotImplementedError(
                "Overriding jump_from / spaxel_iterator is not supported by the GPU sweep; "
                "use mode='coloured' for another update order.")

    # -- iterators / MCMC hooks (documentation of what the kernels do) ---------------------------

    def spaxel_iterator(self):
        """Yields (y, x) row by row over the spaxels whose mask is 1 (lib/run.py:553-566);
        the order of the 'sequential' sweep."""
        h, w = self.cube.data.shape[1], self.cube.data.shape[2]
        for y in range(h):
            for x in range(w):
                if self.mask[y, x] == 1:
                    yield (y, x)

    def jump_from(self, parameters, amplitude):
        """Cauchy jump parameters + amplitude * tan(U(-pi/2, pi/2)) (lib/run.py:570-579).
        Host-side illustration on numpy's generator; the sweep draws on the device."""
        u = np.random.uniform(-np.pi / 2., np.pi / 2., size=len(parameters))
        return parameters + amplitude * np.tan(u)

    def __getattr__(self, name):
        # chain_on_device=True: the host copies of the chain are made on first access
        if name in ('chains', 'chain', 'all_likelihoods', 'likelihoods') and \
                '_chains_device' in self.__dict__:
            self.chains = self._chains_device.cpu().numpy()
            self.chain = self.chains[0]
            self.all_likelihoods = self._likelihoods_device.cpu().numpy()
            self.likelihoods = self.all_likelihoods[0]
            return self.__dict__[name]
        raise AttributeError(name)

    def extract_parameters(self, percentage=20.):
        """Mean of the last ``percentage`` % of the chain, per spaxel (lib/run.py:581-593)."""
        if '_chains_device' in self.__dict__ and 'chains' not in self.__dict__:
            rows = self._chains_device.shape[1]
            start = (100. - percentage) * rows / 100.
            self.parameters_all = self._ctx.chain_mean(self._chains_device, int(start))
            return self.parameters_all[0]
        start = (100. - percentage) * self.chain.shape[0] / 100.
        return np.mean(self.chain[int(start):, ...], 0)

    # -- simulators (GPU) ----------------------------------------------------------------------

    def simulate_clean(self, shape, parameters):
        """Cube of the un-convolved lines (lib/run.py:597-621)."""
        self._check_shape(shape)
        return self._ctx.simulate_clean(np.asarray(parameters, dtype=np.float64)[None])[0]

    def simulate_convolved(self, shape, parameters):
        """Cube of the lines convolved by the LSF and the FSF (lib/run.py:623-652)."""
        self._check_shape(shape)
        return self._ctx.simulate(np.asarray(parameters, dtype=np.float64)[None])[0]

    def _check_shape(self, shape):
        if tuple(shape) != tuple(self.cube.data.shape):
            raise ValueError("shape must be the shape of the cube of this run, %s"
                             % str(self.cube.data.shape))

    def contribution_of_spaxel(self, x, y, parameters, cube_width, cube_height, cube_depth,
                               fsf, lsf, lsf_fft=None):
        """Contribution cube of the line ``parameters`` at spaxel (x, y): FSF (x) (LSF * line),
        pasted into a zero cube and clipped at the borders (lib/run.py:654-708).
        Returns (cube, lsf_fft)."""
        line = self.model.modelize(self, range(0, cube_depth), parameters)
        if lsf is None:
            spread = line
        elif lsf_fft is None:
            spread, lsf_fft = convolve_1d(line, lsf)
        else:
            spread, _ = convolve_1d(line, lsf_fft, compute_fourier=False)
        fh, fw = fsf.shape
        hy, hx = (fh - 1) // 2, (fw - 1) // 2
        out = np.zeros((cube_depth, cube_height, cube_width))
        ys, ye = max(y - hy, 0), min(y + hy + 1, cube_height)
        xs, xe = max(x - hx, 0), min(x + hx + 1, cube_width)
        stamp = fsf[ys - (y - hy):ye - (y - hy), xs - (x - hx):xe - (x - hx)]
        out[:, ys:ye, xs:xe] = stamp[None, :, :] * spread[:, None, None]
        return out, lsf_fft

    # -- saves (lib/run.py:742-840; plotting needs matplotlib, which is optional) -----------------

    def save(self, name, clobber=False):
        """Writes <name>_parameters.npy, <name>_convolved_cube.fits, <name>_clean_cube.fits
        (and <name>_images.png when matplotlib is available)."""
        self.save_parameters_npy("%s_parameters.npy" % name)
        self.convolved_cube.to_fits("%s_convolved_cube.fits" % name, clobber=clobber)
        self.clean_cube.to_fits("%s_clean_cube.fits" % name, clobber=clobber)
        try:
            self.plot_images("%s_images.png" % name)
        except ImportError:
            self.logger.warning("matplotlib is not available: %s_images.png not written" % name)

    def save_parameters_npy(self, filepath):
        np.save(filepath, self.parameters)

    def save_chain_npy(self, filepath):
        np.save(filepath, self.chain)

    def save_matlab(self, filepath):
        import scipy.io
        scipy.io.savemat(filepath, {'chain': self.chain, 'parameters': self.parameters,
                                    'likelihoods': self.likelihoods})

    def _check_image_filepath(self, filepath):
        if filepath is not None:
            _, extension = splitext(filepath)
            if extension not in ['.png', '.pdf']:
                raise ValueError("Extension '%s' is not supported, you may use one of %s"
                                 % (extension, ', '.join(['.png', '.pdf'])))

    def plot_chain(self, filepath=None):
        self._check_image_filepath(filepath)
        from matplotlib import pyplot as plot
        names = self.model.parameters()
        fig = plot.figure()
        for i, n in enumerate(names):
            ax = fig.add_subplot(len(names), 1, i + 1)
            ax.plot(self.chain[:, :, :, i].reshape(self.chain.shape[0], -1))
            ax.set_ylabel(n)
        if filepath is None:
            plot.show()
        else:
            fig.savefig(filepath)
        return fig

    def plot_images(self, filepath=None):
        self._check_image_filepath(filepath)
        from matplotlib import pyplot as plot
        fig = plot.figure()
        cubes = [('data', self.cube.data), ('convolved', self.convolved_cube.data),
                 ('clean', self.clean_cube.data)]
        for i, (title, c) in enumerate(cubes):
            ax = fig.add_subplot(1, len(cubes), i + 1)
            ax.imshow(np.nansum(c, 0), interpolation='nearest', origin='lower')
            ax.set_title(title)
        if filepath is None:
            plot.show()
        else:
            fig.savefig(filepath)
        return fig
