"""
ctypes binding of libdeconv3d_b200.so (include/deconv3d_b200.h).

There is NO CPU fallback: if the library cannot be loaded, or no CUDA device
is present when a context is created, a ``NativeError`` is raised.
"""
import ctypes
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, 'libdeconv3d_b200.so')

F32, F64 = 0, 1
SEQ_EXACT, COLOURED = 0, 1
VAR_SCALAR, VAR_CUBE = 0, 1
EINVAL, ECUDA, ESTATE, ENUMERIC, ENOMEM = -1, -2, -3, -4, -5

# every symbol include/deconv3d_b200.h declares
SYMBOLS = [
    'd3d_abi_version', 'd3d_last_error', 'd3d_ctx_create', 'd3d_ctx_destroy',
    'd3d_ctx_set_stream', 'd3d_ctx_synchronize', 'd3d_set_problem', 'd3d_set_rtnorm_tables',
    'd3d_set_rng', 'd3d_set_params', 'd3d_get_params', 'd3d_init_params_uniform',
    'd3d_forward', 'd3d_simulate', 'd3d_simulate_clean', 'd3d_get_residual', 'd3d_conv1d',
    'd3d_rtnorm', 'd3d_delta_logl', 'd3d_sweep', 'd3d_get_counters',
    'd3d_set_tile', 'd3d_tile_record_slots', 'd3d_colour_begin', 'd3d_colour_phase',
    'd3d_apply_records', 'd3d_get_likelihoods', 'd3d_get_chain_control', 'd3d_chain_mean',
    'd3d_tile_fused_init', 'd3d_tile_fused_export', 'd3d_tile_fused_connect', 'd3d_colour_phase_fused', 'd3d_sweep_fused',
    'd3d_fp64_peak', 'd3d_last_kernel', 'd3d_set_line_model',
]
RECORD_DOUBLES = 8
ABI_VERSION = 2          # D3D_ABI_VERSION of include/deconv3d_b200.h


class NativeError(RuntimeError):
    def __init__(self, code, message):
        RuntimeError.__init__(self, message)
        self.code = code


_lib = None


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, int):              # raw (device) address
        return ctypes.c_void_p(a)
    if hasattr(a, 'data_ptr'):          # torch tensor
        return ctypes.c_void_p(a.data_ptr())
    return a.ctypes.data_as(ctypes.c_void_p)


def load():
    """Loads (building it first if the sources are newer and nvcc is around)
    the native library.  Raises NativeError when it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        try:
            from . import build_native
            build_native.build()
        except Exception as e:      # noqa: BLE001 - report the real reason
            raise NativeError(ECUDA, 'libdeconv3d_b200.so is missing and could not be built '
                                     '(%s); deconv3d_b200 has no CPU fallback' % e)
    try:
        lib = ctypes.CDLL(LIB_PATH)
    except OSError as e:
        raise NativeError(ECUDA, 'cannot load %s: %s; deconv3d_b200 has no CPU fallback'
                          % (LIB_PATH, e))
    vp, ci, cd = ctypes.c_void_p, ctypes.c_int, ctypes.c_double
    i64, u64, u32 = ctypes.c_int64, ctypes.c_uint64, ctypes.c_uint32
    lib.d3d_abi_version.restype = ci
    lib.d3d_last_error.restype = ctypes.c_char_p
    lib.d3d_ctx_create.argtypes = [ctypes.POINTER(vp), ci, ci]
    lib.d3d_ctx_destroy.argtypes = [vp]
    lib.d3d_ctx_set_stream.argtypes = [vp, vp]
    lib.d3d_ctx_synchronize.argtypes = [vp]
    lib.d3d_set_problem.argtypes = [vp, ci, ci, ci, ci, ci, vp, vp, ci, vp, vp, ci, ci, vp,
                                    vp, vp, vp, vp]
    lib.d3d_set_rtnorm_tables.argtypes = [vp, vp, ci, vp, ci, vp, ci]
    lib.d3d_set_rng.argtypes = [vp, u64, u32]
    lib.d3d_set_params.argtypes = [vp, vp]
    lib.d3d_get_params.argtypes = [vp, vp]
    lib.d3d_init_params_uniform.argtypes = [vp]
    lib.d3d_forward.argtypes = [vp, vp, ci, vp]
    lib.d3d_simulate.argtypes = [vp, vp, ci, vp]
    lib.d3d_simulate_clean.argtypes = [vp, vp, ci, vp]
    lib.d3d_get_residual.argtypes = [vp, vp]
    lib.d3d_conv1d.argtypes = [vp, vp, ci, ci, vp, vp]
    lib.d3d_rtnorm.argtypes = [vp, ci, vp, vp, vp, vp, u64, u32, u32, vp, vp]
    lib.d3d_delta_logl.argtypes = [vp, ci, ci, ci, vp, vp]
    lib.d3d_sweep.argtypes = [vp, i64, i64, ci, ci, ci, cd, vp, vp, i64, vp, vp, vp]
    lib.d3d_get_counters.argtypes = [vp, vp, vp, vp]
    lib.d3d_set_tile.argtypes = [vp, ci, ci, ci, ci]
    lib.d3d_tile_record_slots.argtypes = [vp, vp]
    lib.d3d_colour_begin.argtypes = [vp, i64, cd]
    lib.d3d_colour_phase.argtypes = [vp, i64, ci, ci, vp]
    lib.d3d_apply_records.argtypes = [vp, vp, i64]
    lib.d3d_get_likelihoods.argtypes = [vp, vp]
    lib.d3d_get_chain_control.argtypes = [vp, vp, vp, vp]
    lib.d3d_chain_mean.argtypes = [vp, vp, i64, i64, vp]
    lib.d3d_tile_fused_init.argtypes = [vp, ci, ci, vp, vp]
    lib.d3d_tile_fused_export.argtypes = [vp, vp]
    lib.d3d_tile_fused_connect.argtypes = [vp, ci, vp, vp]
    lib.d3d_colour_phase_fused.argtypes = [vp, i64, ci, ci, i64]
    lib.d3d_sweep_fused.argtypes = [vp, i64, i64, cd]
    lib.d3d_fp64_peak.argtypes = [vp, vp]
    lib.d3d_set_line_model.argtypes = [vp, ci, vp, vp]
    lib.d3d_last_kernel.argtypes = [vp]
    for name in SYMBOLS:
        fn = getattr(lib, name)
        if name not in ('d3d_last_error', 'd3d_last_kernel'):
            fn.restype = ci
    lib.d3d_last_kernel.restype = ctypes.c_char_p
    if lib.d3d_abi_version() != ABI_VERSION:
        raise NativeError(EINVAL, 'libdeconv3d_b200.so has ABI %d, expected %d'
                          % (lib.d3d_abi_version(), ABI_VERSION))
    _lib = lib
    return lib


def _check(rc):
    if rc != 0:
        raise NativeError(rc, load().d3d_last_error().decode('utf-8', 'replace'))


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class Context(object):
    """One GPU context of the native library (thin, explicit wrapper)."""

    def __init__(self, device=0, dtype=F64):
        self.lib = load()
        self.h = ctypes.c_void_p()
        _check(self.lib.d3d_ctx_create(ctypes.byref(self.h), int(device), int(dtype)))
        self.dtype = dtype
        self.shape = None
        self.n_chains = 0
        self._tables = False

    def close(self):
        if self.h:
            self.lib.d3d_ctx_destroy(self.h)
            self.h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:       # noqa: BLE001
            pass

    def set_stream(self, cuda_stream_ptr):
        _check(self.lib.d3d_ctx_set_stream(self.h, ctypes.c_void_p(cuda_stream_ptr or 0)))

    def synchronize(self):
        _check(self.lib.d3d_ctx_synchronize(self.h))

    def set_problem(self, data, variance, fsf, lsf, pmin, pmax, jump_amp, prior_var,
                    mask=None, chains_per_cube=1):
        """data: [n,D,H,W] or [D,H,W]; variance: same shape (cube) or [n] / scalar."""
        data = _f64(data)
        if data.ndim == 3:
            data = data[None]
        n, D, H, W = data.shape
        variance = _f64(variance)
        if variance.size == data.size:
            var_kind = VAR_CUBE
            variance = variance.reshape(data.shape)
        elif variance.size == n:
            var_kind = VAR_SCALAR
            variance = variance.reshape(n)
        else:
            raise ValueError('Provided variance has not the correct shape.')
        fsf = _f64(fsf)
        lsf = None if lsf is None else _f64(lsf)
        if lsf is not None and lsf.shape != (D,):
            raise ValueError('LSF vector must have the spectral length of the cube')
        pmin = np.ascontiguousarray(np.broadcast_to(_f64(pmin), (n, 3)))
        pmax = np.ascontiguousarray(np.broadcast_to(_f64(pmax), (n, 3)))
        prior = np.ascontiguousarray(np.broadcast_to(_f64(prior_var), (n,)))
        jump = np.ascontiguousarray(np.broadcast_to(_f64(jump_amp), (3,)))
        m = None
        if mask is not None:
            m = np.ascontiguousarray(np.broadcast_to(np.asarray(mask) == 1, (n, H, W)),
                                     dtype=np.uint8)
        _check(self.lib.d3d_set_problem(
            self.h, n, int(chains_per_cube), D, H, W, _ptr(data), _ptr(variance), var_kind,
            _ptr(m), _ptr(fsf), fsf.shape[0], fsf.shape[1], _ptr(lsf), _ptr(pmin), _ptr(pmax),
            _ptr(jump), _ptr(prior)))
        self.shape = (D, H, W)
        self.n_cubes = n
        self.n_chains = n * int(chains_per_cube)

    def set_rtnorm_tables(self, x, yu, ncell):
        x, yu = _f64(x), _f64(yu)
        ncell = np.ascontiguousarray(ncell, dtype=np.int32)
        _check(self.lib.d3d_set_rtnorm_tables(self.h, _ptr(x), x.size, _ptr(yu), yu.size,
                                              _ptr(ncell), ncell.size))
        self._tables = True

    def set_rng(self, seed, first_chain_id=0):
        _check(self.lib.d3d_set_rng(self.h, int(seed) & (2 ** 64 - 1), int(first_chain_id)))

    def set_params(self, params):
        D, H, W = self.shape
        params = _f64(params).reshape(self.n_chains, H, W, 3)
        _check(self.lib.d3d_set_params(self.h, _ptr(params)))

    def get_params(self):
        D, H, W = self.shape
        out = np.empty((self.n_chains, H, W, 3))
        _check(self.lib.d3d_get_params(self.h, _ptr(out)))
        return out

    def init_params_uniform(self):
        _check(self.lib.d3d_init_params_uniform(self.h))

    def forward(self, want_sim=False, write_err=True, want_chi2=False):
        D, H, W = self.shape
        sim = np.empty((self.n_chains, D, H, W)) if want_sim else None
        chi2 = np.empty(self.n_chains) if want_chi2 else None
        _check(self.lib.d3d_forward(self.h, _ptr(sim), 1 if write_err else 0, _ptr(chi2)))
        return sim, chi2

    def simulate(self, params):
        """params [n, H, W, 3] (n <= n_chains) -> convolved cubes [n, D, H, W]."""
        D, H, W = self.shape
        params = _f64(params).reshape(-1, H, W, 3)
        sim = np.empty((params.shape[0], D, H, W))
        _check(self.lib.d3d_simulate(self.h, _ptr(params), params.shape[0], _ptr(sim)))
        return sim

    def simulate_clean(self, params):
        D, H, W = self.shape
        params = _f64(params).reshape(-1, H, W, 3)
        sim = np.empty((params.shape[0], D, H, W))
        _check(self.lib.d3d_simulate_clean(self.h, _ptr(params), params.shape[0], _ptr(sim)))
        return sim

    def get_residual(self):
        D, H, W = self.shape
        out = np.empty((self.n_chains, D, H, W))
        _check(self.lib.d3d_get_residual(self.h, _ptr(out)))
        return out

    def conv1d(self, lines, lsf):
        lines = _f64(lines)
        lsf = _f64(lsf)
        flat = lines.reshape(-1, lines.shape[-1])
        out = np.empty_like(flat)
        _check(self.lib.d3d_conv1d(self.h, _ptr(flat), flat.shape[1], flat.shape[0], _ptr(lsf),
                                   _ptr(out)))
        return out.reshape(lines.shape)

    def rtnorm_batch(self, a, b, mu, sigma, seed=0, chain=0, sweep=0):
        """Truncated-normal variates, one per entry; returns (values, draws_used)."""
        if not self._tables:
            from . import rtnorm_tables
            self.set_rtnorm_tables(*rtnorm_tables.tables())
        a, b, mu, sigma = [_f64(v).reshape(-1) for v in (a, b, mu, sigma)]
        out = np.empty(a.size)
        used = np.zeros(a.size, dtype=np.int32)
        _check(self.lib.d3d_rtnorm(self.h, a.size, _ptr(a), _ptr(b), _ptr(mu), _ptr(sigma),
                                   int(seed) & (2 ** 64 - 1), int(chain), int(sweep),
                                   _ptr(out), _ptr(used)))
        return out, used

    def delta_logl(self, chain, y, x, p_new):
        p = _f64(p_new).reshape(3)
        out = np.empty(3)
        _check(self.lib.d3d_delta_logl(self.h, int(chain), int(y), int(x), _ptr(p), _ptr(out)))
        return out

    def sweep(self, first_iteration, n_iterations, mode=SEQ_EXACT, keep_one_in=1,
              refresh_every=1000, min_acceptance_rate=0.01, chain_out=None, lik_out=None):
        """Runs iterations [first, first+n). chain_out [n_chains,rows,H,W,3] / lik_out
        [n_chains,rows,H,W] are float64 C-contiguous numpy arrays (or None).
        Returns (accepted[n_chains], iterations[n_chains], elapsed_ms)."""
        n_rows = 0
        for a in (chain_out, lik_out):
            if a is not None:
                if hasattr(a, 'data_ptr'):          # torch tensor (device or host)
                    assert str(a.dtype) == 'torch.float64' and a.is_contiguous()
                else:
                    assert a.dtype == np.float64 and a.flags['C_CONTIGUOUS']
                assert a.shape[0] == self.n_chains
                n_rows = a.shape[1]
        acc = np.zeros(self.n_chains, dtype=np.int64)
        its = np.zeros(self.n_chains, dtype=np.int64)
        ms = ctypes.c_float(0.0)
        _check(self.lib.d3d_sweep(self.h, int(first_iteration), int(n_iterations), int(mode),
                                  int(keep_one_in), int(refresh_every),
                                  float(min_acceptance_rate), _ptr(chain_out), _ptr(lik_out),
                                  int(n_rows), _ptr(acc), _ptr(its), ctypes.byref(ms)))
        return acc, its, ms.value

    # ---- one cube tiled over several contexts (include/deconv3d_b200.h, "tiled") ----
    def set_tile(self, y0, y1, x0, x1):
        _check(self.lib.d3d_set_tile(self.h, int(y0), int(y1), int(x0), int(x1)))

    def record_slots(self):
        n = ctypes.c_int64(0)
        _check(self.lib.d3d_tile_record_slots(self.h, ctypes.byref(n)))
        return n.value

    def colour_begin(self, iteration, min_acceptance_rate=0.0):
        _check(self.lib.d3d_colour_begin(self.h, int(iteration), float(min_acceptance_rate)))

    def colour_phase(self, iteration, cy, cx, records=None):
        """records: None, a float64 numpy array [slots, 8] (filled synchronously) or an int
        device address (e.g. ``tensor.data_ptr()``; filled asynchronously on the stream)."""
        ptr = ctypes.c_void_p(records) if isinstance(records, int) else _ptr(records)
        _check(self.lib.d3d_colour_phase(self.h, int(iteration), int(cy), int(cx), ptr))

    def apply_records(self, records, n_records=None):
        if isinstance(records, int):
            ptr, n = ctypes.c_void_p(records), int(n_records)
        else:
            records = _f64(records).reshape(-1, RECORD_DOUBLES)
            ptr, n = _ptr(records), records.shape[0]
        _check(self.lib.d3d_apply_records(self.h, ptr, n))

    # fused exchange over peer memory
    def fused_init(self, n_tiles, my_index):
        """Returns the device address of this context's box."""
        box = ctypes.c_void_p()
        nbytes = ctypes.c_int64(0)
        _check(self.lib.d3d_tile_fused_init(self.h, int(n_tiles), int(my_index), ctypes.byref(box),
                                            ctypes.byref(nbytes)))
        return box.value

    def fused_export(self):
        buf = (ctypes.c_ubyte * 64)()
        _check(self.lib.d3d_tile_fused_export(self.h, buf))
        return bytes(buf)

    def fused_connect(self, index, box=None, handle=None):
        hb = (ctypes.c_ubyte * 64).from_buffer_copy(handle) if handle is not None else None
        _check(self.lib.d3d_tile_fused_connect(self.h, int(index), ctypes.c_void_p(box) if box else None, hb))

    def colour_phase_fused(self, iteration, cy, cx, phase_index):
        _check(self.lib.d3d_colour_phase_fused(self.h, int(iteration), int(cy), int(cx), int(phase_index)))

    def sweep_fused(self, first_iteration, n_iterations, min_acceptance_rate=0.0):
        _check(self.lib.d3d_sweep_fused(self.h, int(first_iteration), int(n_iterations),
                                        float(min_acceptance_rate)))

    def get_likelihoods(self):
        D, H, W = self.shape
        out = np.empty((self.n_chains, H, W))
        _check(self.lib.d3d_get_likelihoods(self.h, _ptr(out)))
        return out

    def chain_control(self):
        acc = np.zeros(self.n_chains, dtype=np.int64)
        its = np.zeros(self.n_chains, dtype=np.int64)
        act = np.zeros(self.n_chains, dtype=np.int32)
        _check(self.lib.d3d_get_chain_control(self.h, _ptr(acc), _ptr(its), _ptr(act)))
        return acc, its, act

    def chain_mean(self, chain, first_row):
        """Mean over rows [first_row:] of chain [n_chains, rows, H, W, 3] (numpy or torch tensor,
        host or device) -> numpy [n_chains, H, W, 3]."""
        D, H, W = self.shape
        out = np.empty((self.n_chains, H, W, 3))
        _check(self.lib.d3d_chain_mean(self.h, _ptr(chain), int(chain.shape[1]), int(first_row), _ptr(out)))
        return out

    def set_line_model(self, offsets=None, ratios=None):
        """Tied multiplet of Gaussians (offsets in channels, relative amplitudes); None: one Gaussian."""
        if offsets is None:
            _check(self.lib.d3d_set_line_model(self.h, 1, None, None))
            return
        off, rat = _f64(offsets).ravel(), _f64(ratios).ravel()
        assert off.shape == rat.shape
        _check(self.lib.d3d_set_line_model(self.h, int(off.size), _ptr(off), _ptr(rat)))

    def last_kernel(self):
        """Name of the sweep kernel the latest sweep() launched."""
        return self.lib.d3d_last_kernel(self.h).decode()

    def fp64_peak(self):
        """Measured FP64 FMA peak of the device in TFLOP/s (bench support)."""
        v = ctypes.c_double(0.0)
        _check(self.lib.d3d_fp64_peak(self.h, ctypes.byref(v)))
        return v.value

    def counters(self):
        a, b, c = ctypes.c_int64(0), ctypes.c_int64(0), ctypes.c_int64(0)
        _check(self.lib.d3d_get_counters(self.h, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c)))
        return dict(kernel_launches=a.value, last_sweep_bytes=b.value,
                    last_sweep_site_updates=c.value)
