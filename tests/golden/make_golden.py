#!/usr/bin/env python
"""
tests/golden/make_golden.py -- generates the golden fixtures in this directory
by running the REFERENCE'S OWN SOURCES (read from /root/reference at run time,
never copied) under Python 3.

Run it in the build container (the only place /root/reference exists):

    python tests/golden/make_golden.py

The reference is Python 2.7 and depends on astropy / hyperspectral /
matplotlib, all absent here.  Its modules are therefore compiled from their
source text after the *textual* py2->py3 shims listed in ``SHIMS`` below (each
one is a syntax / integer-division / indexing fix, none touches arithmetic),
with stub modules standing in for the absent I/O-only dependencies
(oracle/minifits.py provides the duck-typed HyperspectralCube).  What comes
out is what lib/run.py, lib/convolution.py, lib/spread_functions.py,
lib/rtnorm.py compute, under ``numpy.random.seed``-ed global state.

Outputs (all read by tests that run without /root/reference):
    ref_run_A.npz .. ref_run_D.npz  full ``Run`` results (chain, likelihoods,
                                    parameters, convolved/clean cubes) + inputs
    ref_conv1d.npz                  convolve_1d outputs for several depths
    ref_spread.npz                  FSF images / LSF vectors of the generators
    ref_rtnorm.npz                  rtnorm draws with injected random numbers
    ref_rtnorm_tables.npz           sparse samples of the reference tables
    mat_kat.npz                     the .mat known-answer fixture (data, variance,
                                    FSF, ground-truth parameters)
    muse_cube_01.npz                tests/input/test_cube_01.fits as an array
"""

import os
import sys
import types
import logging

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = '/root/reference'
sys.path.insert(0, ROOT)

from oracle import minifits          # noqa: E402
from oracle import streams           # noqa: E402

# (module, [(old, new), ...]) -- every occurrence is replaced; an absent
# pattern is an error so that a changed reference cannot go unnoticed.
SHIMS = {
    'convolution': [
        ('half = diff / 2 + 1', 'half = diff // 2 + 1'),
        ('half = diff / 2\n', 'half = diff // 2\n'),
        ('cube_padded[cube_slices]', 'cube_padded[tuple(cube_slices)]'),
        ('fft[boxcube]', 'fft[tuple(boxcube)]'),
        # a 1-element ndarray is no longer accepted as a slice bound
        ('np.shape(cubep)[slice(axis, axis + 1)]',
         'np.shape(cubep)[slice(int(axis[0]), int(axis[0]) + 1)]'),
    ],
    'spread_functions': [
        ('xo = (shape[1] - 1) / 2 - (shape[1] % 2 - 1)',
         'xo = (shape[1] - 1) // 2 - (shape[1] % 2 - 1)'),
        ('yo = (shape[0] - 1) / 2 - (shape[0] % 2 - 1)',
         'yo = (shape[0] - 1) // 2 - (shape[0] % 2 - 1)'),
        ('z_center_index = (depth - 1) / 2 - (depth % 2 - 1)',
         'z_center_index = (depth - 1) // 2 - (depth % 2 - 1)'),
    ],
    'instruments': [],
    'line_models': [],
    'math_utils': [],
    'rtnorm': [],
    'masks': [],
    'run': [
        ('MAXIMUM = sys.maxint', 'MAXIMUM = sys.maxsize'),
        ('fhh = (fh - 1) / 2  # FSF half height', 'fhh = (fh - 1) // 2'),
        ('fhw = (fw - 1) / 2  # FSF half width', 'fhw = (fw - 1) // 2'),
        ('fhh = (fh-1)/2  # FSF half height', 'fhh = (fh-1)//2'),
        ('fhw = (fw-1)/2  # FSF half width', 'fhw = (fw-1)//2'),
        ('i = cur_iteration / keep_one_in', 'i = cur_iteration // keep_one_in'),
        ('print e\n', 'print(e)\n'),
    ],
}


def _install_stubs():
    hyp = types.ModuleType('hyperspectral')
    hyp.HyperspectralCube = minifits.Cube
    hyp.Axis = lambda *a, **k: None
    sys.modules['hyperspectral'] = hyp

    astropy = types.ModuleType('astropy')
    units = types.ModuleType('astropy.units')

    class _U(object):
        def __init__(self, name):
            self.name = name
    units.um = _U('um')
    units.Unit = _U
    astropy.units = units
    io = types.ModuleType('astropy.io')
    fits = types.ModuleType('astropy.io.fits')
    fits.Header = dict
    fits.getdata = lambda path: minifits.read_fits(path)[0]
    io.fits = fits
    astropy.io = io
    sys.modules.update({'astropy': astropy, 'astropy.units': units,
                        'astropy.io': io, 'astropy.io.fits': fits})

    mpl = types.ModuleType('matplotlib')
    pyplot = types.ModuleType('matplotlib.pyplot')
    mpl.pyplot = pyplot
    sys.modules.update({'matplotlib': mpl, 'matplotlib.pyplot': pyplot})


def load_reference():
    """Compile the reference's lib/*.py (shimmed in memory) as top-level
    modules named the way its py2 implicit-relative imports expect."""
    _install_stubs()
    mods = {}
    for name in ['convolution', 'math_utils', 'line_models', 'rtnorm',
                 'spread_functions', 'instruments', 'masks', 'run']:
        path = os.path.join(REF, 'lib', name + '.py')
        src = open(path, encoding='utf-8').read()
        for old, new in SHIMS[name]:
            if old not in src:
                raise RuntimeError('shim pattern not found in %s: %r' % (name, old))
            src = src.replace(old, new)
        mod = types.ModuleType(name)
        mod.__file__ = path
        mod.__dict__['basestring'] = str            # py2 builtin
        sys.modules[name] = mod
        exec(compile(src, path, 'exec'), mod.__dict__)
        mods[name] = mod
    return mods


def synthetic_cube(D, H, W, seed, amp=8.0, noise=0.05, nan_spaxel=None):
    rng = np.random.RandomState(seed)
    yy, xx = np.mgrid[0:H, 0:W]
    a = amp * np.exp(-((yy - H / 2.) ** 2 + (xx - W / 2.) ** 2) / (2. * (0.35 * H) ** 2))
    c = D / 2. + 0.15 * D * np.tanh((xx - W / 2.) / (0.3 * W))
    w = 1.2 + 0.05 * yy
    z = np.arange(D)[:, None, None]
    clean = a * np.exp(-(z - c) ** 2 / (2 * w ** 2))
    data = clean + noise * rng.randn(D, H, W)
    if nan_spaxel is not None:
        data[3, nan_spaxel[0], nan_spaxel[1]] = np.nan
    return data


def run_reference(mods, name, data, seed, inst_kwargs, run_kwargs):
    Run = mods['run'].Run
    MUSE = mods['instruments'].MUSE
    cube = minifits.muse_cube(data.copy())
    kw = dict(run_kwargs)
    if kw.get('mask') is not None:
        kw['mask'] = kw['mask'].copy()            # Run mutates it (lib/run.py:162)
    np.random.seed(seed)
    run = Run(cube, MUSE(**inst_kwargs), **kw)
    out = dict(
        data=data, seed=seed,
        fsf=run.fsf, lsf=run.lsf, mask=np.asarray(run.mask, dtype=float),
        variance_cube=run.variance_cube,
        chain=run.chain, likelihoods=run.likelihoods, parameters=run.parameters,
        convolved_cube=run.convolved_cube.data, clean_cube=run.clean_cube.data,
        max_iterations=run_kwargs.get('max_iterations'),
        keep_one_in=run_kwargs.get('keep_one_in', 1),
    )
    for k in ('initial_parameters', 'variance', 'jump_amplitude',
              'gibbs_apriori_variance'):
        if run_kwargs.get(k) is not None:
            out['in_' + k] = np.asarray(run_kwargs[k])
    if run_kwargs.get('mask') is not None:
        out['in_mask'] = np.asarray(run_kwargs['mask'], dtype=float)
    for k, v in inst_kwargs.items():
        out['inst_' + k] = v
    # garbage-free copies: chain / likelihoods are np.ndarray() (uninitialised,
    # lib/run.py:270,281) outside the mask and for likelihoods[0]
    m = out['mask'] == 1
    out['chain'] = np.where(m[None, :, :, None], out['chain'], 0.0)
    out['likelihoods'] = np.where(m[None, :, :], out['likelihoods'], 0.0)
    out['likelihoods'][0] = 0.0
    out['parameters'] = np.where(m[:, :, None], out['parameters'], 0.0)
    np.savez_compressed(os.path.join(HERE, name + '.npz'), **out)
    print(name, 'chain', run.chain.shape, 'fsf', run.fsf.shape)


def main():
    logging.disable(logging.INFO)
    mods = load_reference()

    # ---- full Run goldens ------------------------------------------------
    # A: small cube, 7x7 Gaussian FSF, scalar variance guessed by median_clip
    run_reference(mods, 'ref_run_A', synthetic_cube(12, 9, 10, 1), 7,
                  dict(fsf_fwhm=0.5), dict(max_iterations=25))
    # B: the bundled MUSE cube (x1e20 so that lib/run.py:140-143 passes), MUSE()
    #    defaults (13x13 FSF, D=30 -> spectral wrap), 3 iterations
    muse, _ = minifits.read_fits(os.path.join(REF, 'tests/input/test_cube_01.fits'))
    np.savez_compressed(os.path.join(HERE, 'muse_cube_01.npz'), data=muse)
    run_reference(mods, 'ref_run_B', muse * 1e20, 11, dict(), dict(max_iterations=3))
    # C: mask + variance cube + initial parameters + keep_one_in, D=16 (full wrap),
    #    elliptical rotated FSF, custom jump amplitude and Gibbs prior variance
    dC = synthetic_cube(16, 8, 9, 2)   # (a NaN voxel trips lib/run.py:140-143: max(data) is nan)
    rs = np.random.RandomState(5)
    maskC = (rs.rand(8, 9) > 0.25).astype(float)
    varC = 0.05 ** 2 * (1.0 + rs.rand(16, 8, 9))
    ipC = np.dstack([rs.rand(8, 9) * 5, 2 + rs.rand(8, 9) * 11, 0.5 + rs.rand(8, 9) * 3])
    ipC[2, 3, 0] = 0.0          # amplitude 0 -> the lib/run.py:473-488 branch
    run_reference(mods, 'ref_run_C', dC, 3,
                  dict(fsf_fwhm=0.6, fsf_pa=30., fsf_ba=0.7, lsf_fwhm=0.0004),
                  dict(max_iterations=12, keep_one_in=2, mask=maskC, variance=varC,
                       initial_parameters=ipC, jump_amplitude=0.3,
                       gibbs_apriori_variance=50.0))
    # D: > 1000 iterations to cross the residual refresh at lib/run.py:525-534
    run_reference(mods, 'ref_run_D', synthetic_cube(8, 7, 6, 4, noise=0.3), 13,
                  dict(fsf_fwhm=0.2), dict(max_iterations=1003, keep_one_in=50))

    # ---- convolve_1d ------------------------------------------------------
    conv = {}
    rs = np.random.RandomState(21)
    for D in (2, 3, 8, 16, 21, 30, 31, 32, 33, 40, 41, 63, 64, 65):
        line = rs.rand(D)
        lsf = mods['spread_functions'].GaussianLineSpreadFunction(0.0002675) \
            .as_vector(minifits.muse_cube(np.zeros((D, 2, 2))))
        lsf_rand = rs.rand(D)
        out, fftpsf = mods['convolution'].convolve_1d(line, lsf)
        out2, _ = mods['convolution'].convolve_1d(line, fftpsf, compute_fourier=False)
        out3, _ = mods['convolution'].convolve_1d(line, lsf_rand)
        assert np.array_equal(out, out2)
        conv['line_%d' % D] = line
        conv['lsf_%d' % D] = lsf
        conv['out_%d' % D] = out
        conv['lsfrand_%d' % D] = lsf_rand
        conv['outrand_%d' % D] = out3
    np.savez_compressed(os.path.join(HERE, 'ref_conv1d.npz'), **conv)

    # ---- spread-function generators ----------------------------------------
    sf = mods['spread_functions']
    cube40 = minifits.muse_cube(np.zeros((40, 41, 39)))
    spread = dict(
        gauss_default=sf.GaussianFieldSpreadFunction(1.0).as_image(cube40),
        gauss_08=sf.GaussianFieldSpreadFunction(0.8).as_image(cube40),
        gauss_ell=sf.GaussianFieldSpreadFunction(0.9, pa=25., ba=0.6).as_image(cube40),
        moffat_41x39=sf.MoffatFieldSpreadFunction(fwhm=0.8, beta=2.5, pa=0., ba=1.0)
        .as_image(cube40),
        moffat_alpha=sf.MoffatFieldSpreadFunction(alpha=0.5, beta=3.0, pa=10., ba=0.8)
        .as_image(cube40),
        lsf_40=sf.GaussianLineSpreadFunction(0.0002675).as_vector(cube40),
        lsf_30=sf.GaussianLineSpreadFunction(0.0002675).as_vector(
            minifits.muse_cube(np.zeros((30, 3, 3)))),
        lsf_21_delta=sf.GaussianLineSpreadFunction(0.0).as_vector(
            minifits.muse_cube(np.zeros((21, 3, 3)))),
    )
    np.savez_compressed(os.path.join(HERE, 'ref_spread.npz'), **spread)

    # ---- rtnorm with injected random numbers -------------------------------
    rt = mods['rtnorm']
    cases = []
    rs = np.random.RandomState(99)
    for i in range(400):
        kind = i % 8
        if kind == 0:      # generic
            mu, sg = rs.randn() * 3, 0.1 + rs.rand() * 3
            a, b = 0.0, 5.0 + rs.rand() * 50
        elif kind == 1:    # left tail: a far below the mean -> Gaussian proposal
            mu, sg, a, b = 5 + rs.rand() * 5, 0.05 + rs.rand() * 0.2, 0.0, 20.0
        elif kind == 2:    # right tail: mean below a
            mu, sg, a, b = -1 - rs.rand() * 3, 0.2 + rs.rand() * 0.2, 0.0, 30.0
        elif kind == 3:    # narrow interval -> small-range branch
            mu, sg = rs.randn(), 1.0
            a = rs.randn()
            b = a + 1e-3 + rs.rand() * 2e-3
        elif kind == 4:    # mirrored
            mu, sg, a, b = rs.rand() * 2, 1.0, -40.0, 0.5
        elif kind == 5:    # table method, b beyond xmax
            mu, sg, a, b = 0.0, 1.0, -1.5 + rs.rand() * 3, 1e6
        elif kind == 6:    # table method, both ends inside
            mu, sg = 0.0, 1.0
            a = -1.9 + rs.rand() * 2
            b = a + 0.05 + rs.rand() * 3
        else:              # standard-normal shortcut (mu=0, sigma=1)
            mu, sg, a, b = 0.0, 1.0, rs.rand(), rs.rand() + 2
        cases.append((a, b, mu, sg))
    cases = np.array(cases)
    outs = np.zeros(len(cases))
    used = np.zeros(len(cases), dtype=np.int64)
    for i, (a, b, mu, sg) in enumerate(cases):
        st = streams.PhiloxStream(seed=2024, chain=1)
        st.begin_site(3, i)
        rt.rand = lambda low=0.0, st=st: st.rand(low)
        rt.randn = lambda st=st: st.randn()
        rt.randi = lambda low, high, st=st: st.randi(low, high)
        outs[i] = rt.rtnorm(a, b, mu=mu, sigma=sg)[0]
        used[i] = st.k
    np.savez_compressed(os.path.join(HERE, 'ref_rtnorm.npz'),
                        cases=cases, out=outs, used=used, seed=2024, chain=1, sweep=3)
    x = np.asarray(rt.x)
    yu = np.asarray(rt.yu)
    nc = np.asarray(rt.ncell)
    idx = np.unique(np.concatenate([np.arange(0, 4002, 97), [0, 1, 1953, 1954, 1955, 4000, 4001]]))
    idy = idx[idx < 4001]
    idn = np.unique(np.concatenate([np.arange(0, 8961, 53), [0, 3270, 3271, 3272, 8960]]))
    np.savez_compressed(
        os.path.join(HERE, 'ref_rtnorm_tables.npz'),
        x_idx=idx, x_val=x[idx], yu_idx=idy, yu_val=yu[idy],
        ncell_idx=idn, ncell_val=nc[idn],
        ncell_crc=np.array([int(np.sum(nc * (np.arange(len(nc)) % 251 + 1)))]),
        x_sum=np.array([x.sum()]), yu_sum=np.array([yu.sum()]))

    # ---- the .mat known-answer fixture -------------------------------------
    import scipy.io
    mat = scipy.io.loadmat(os.path.join(REF, 'tests/input/data14forAntoine.mat'))
    par = scipy.io.loadmat(os.path.join(REF, 'tests/input/Parametres_theoriques.mat'))[
        'Parametres_theoriques']
    # tests/read_mat.py:33-35, 49-66: transpose to (z, y, x); c is 1-based
    a = np.transpose(par[:, :, 0])
    c = np.transpose(par[:, :, 1]) - 1
    w = np.transpose(par[:, :, 2])
    np.savez_compressed(
        os.path.join(HERE, 'mat_kat.npz'),
        data=np.transpose(mat['data_noise']), variance=np.transpose(mat['varNoise']),
        fsf=np.transpose(mat['FSF']), params=np.dstack((a, c, w)))
    print('done')


if __name__ == '__main__':
    main()
