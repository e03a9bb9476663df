"""
Pins the oracle (oracle/reference_port.py, oracle/rtnorm_port.py) to outputs of
the reference's own code (tests/golden/*.npz, made by tests/golden/make_golden.py
from /root/reference).  CPU only.
"""
import importlib.util
import os

import numpy as np
import pytest

from oracle import reference_port as port
from oracle import rtnorm_port, streams
from conftest import load_golden


REF_RTNORM = '/root/reference/lib/rtnorm.py'


def _reference_tables():
    """The reference's own (x, yu, ncell), read from its module when the
    reference tree is present (build container); None elsewhere."""
    if not os.path.exists(REF_RTNORM):
        return None
    spec = importlib.util.spec_from_file_location('_ref_rtnorm', REF_RTNORM)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return (np.asarray(mod.x), np.asarray(mod.yu), np.asarray(mod.ncell))


def _run_port_like_golden(g, tables=None):
    kw = {'rtnorm_tables': tables}
    if 'in_mask' in g:
        kw['mask'] = g['in_mask'].copy()
    if 'in_variance' in g:
        kw['variance_cube'] = g['in_variance']
    if 'in_initial_parameters' in g:
        kw['initial_parameters'] = g['in_initial_parameters']
    if 'in_jump_amplitude' in g:
        kw['jump_amplitude'] = g['in_jump_amplitude']
    if 'in_gibbs_apriori_variance' in g:
        kw['gibbs_apriori_variance'] = float(g['in_gibbs_apriori_variance'])
    np.random.seed(int(g['seed']))
    return port.run_chain(g['data'], g['fsf'], g['lsf'], streams.NumpyGlobalStream(),
                          max_iterations=int(g['max_iterations']),
                          keep_one_in=int(g['keep_one_in']), **kw)


@pytest.mark.parametrize('name', ['ref_run_A', 'ref_run_B', 'ref_run_C', 'ref_run_D'])
def test_run_chain_bit_exact_vs_reference(name):
    """With the reference's own rtnorm tables the port reproduces the reference
    chain bit for bit; with the regenerated tables (the only ones available
    off the build container) to the tables' 5e-12 printing precision."""
    g = load_golden(name)
    tables = _reference_tables()
    out = _run_port_like_golden(g, tables)
    m = g['mask'] == 1
    if tables is None:
        chain = np.where(m[None, :, :, None], out['chain'], 0.0)
        np.testing.assert_allclose(chain, g['chain'], rtol=1e-7, atol=1e-9)
        return
    assert np.array_equal(out['mask'], g['mask'])
    assert np.array_equal(out['variance_cube'], g['variance_cube'])
    chain = np.where(m[None, :, :, None], out['chain'], 0.0)
    lik = np.where(m[None, :, :], out['likelihoods'], 0.0)
    lik[0] = 0.0
    par = np.where(m[:, :, None], out['parameters'], 0.0)
    assert np.array_equal(chain, g['chain'])
    assert np.array_equal(lik, g['likelihoods'])
    assert np.array_equal(par, g['parameters'])
    # final cubes (lib/run.py:542-549) from the extracted parameters; masked-out
    # spaxels are never read by the simulators
    conv = port.simulate_convolved(g['data'].shape, g['parameters'], g['mask'],
                                   g['fsf'], g['lsf'])
    clean = port.simulate_clean(g['data'].shape, g['parameters'], g['mask'])
    assert np.array_equal(conv, g['convolved_cube'])
    assert np.array_equal(clean, g['clean_cube'])


@pytest.mark.parametrize('name', ['ref_run_A', 'ref_run_C', 'ref_run_D'])
def test_run_chain_regenerated_tables_close(name):
    g = load_golden(name)
    out = _run_port_like_golden(g, None)
    m = g['mask'] == 1
    chain = np.where(m[None, :, :, None], out['chain'], 0.0)
    lik = np.where(m[None, :, :], out['likelihoods'], 0.0)
    lik[0] = 0.0
    np.testing.assert_allclose(chain, g['chain'], rtol=1e-7, atol=1e-9)
    np.testing.assert_allclose(lik, g['likelihoods'], rtol=1e-6, atol=1e-7)


def test_convolve_1d_vs_reference_and_direct_form():
    g = load_golden('ref_conv1d')
    for D in (2, 3, 8, 16, 21, 30, 31, 32, 33, 40, 41, 63, 64, 65):
        line, lsf = g['line_%d' % D], g['lsf_%d' % D]
        out, fftpsf = port.convolve_1d(line, lsf)
        assert np.array_equal(out, g['out_%d' % D])
        out2, _ = port.convolve_1d(line, fftpsf, compute_fourier=False)
        assert np.array_equal(out2, out)
        outr, _ = port.convolve_1d(line, g['lsfrand_%d' % D])
        assert np.array_equal(outr, g['outrand_%d' % D])
        # the wrap-around rule the CUDA spectral pass implements
        np.testing.assert_allclose(port.convolve_1d_direct(line, lsf), out,
                                   rtol=0, atol=2e-15)
        np.testing.assert_allclose(port.convolve_1d_direct(line, g['lsfrand_%d' % D]),
                                   outr, rtol=0, atol=2e-14)


def test_spread_function_generators_vs_reference():
    g = load_golden('ref_spread')
    step = 5.5555555555555e-05 * 3600.0
    assert np.array_equal(port.gaussian_fsf_image(1.0, step), g['gauss_default'])
    assert g['gauss_default'].shape == (13, 13)
    assert np.array_equal(port.gaussian_fsf_image(0.8, step), g['gauss_08'])
    assert np.array_equal(port.gaussian_fsf_image(0.9, step, pa=25., ba=0.6), g['gauss_ell'])
    assert np.array_equal(port.moffat_fsf_image((41, 39), step, fwhm_arcsec=0.8, beta=2.5),
                          g['moffat_41x39'])
    assert np.array_equal(port.moffat_fsf_image((41, 39), step, alpha_arcsec=0.5, beta=3.0,
                                                pa=10., ba=0.8), g['moffat_alpha'])
    assert np.array_equal(port.gaussian_lsf_vector(0.0002675, 1.25e-4, 40), g['lsf_40'])
    assert np.array_equal(port.gaussian_lsf_vector(0.0002675, 1.25e-4, 30), g['lsf_30'])
    assert np.array_equal(port.gaussian_lsf_vector(0.0, 1.25e-4, 21), g['lsf_21_delta'])


def test_rtnorm_tables_vs_reference_samples():
    g = load_golden('ref_rtnorm_tables')
    x, yu, ncell = rtnorm_port.build_tables()
    assert x.shape == (4002,) and yu.shape == (4001,) and ncell.shape == (8961,)
    np.testing.assert_allclose(x[g['x_idx']], g['x_val'], rtol=0, atol=1e-11)
    np.testing.assert_allclose(yu[g['yu_idx']], g['yu_val'], rtol=2e-11, atol=0)
    assert np.array_equal(ncell[g['ncell_idx']], g['ncell_val'])
    crc = int(np.sum(ncell * (np.arange(len(ncell)) % 251 + 1)))
    assert crc == int(g['ncell_crc'][0])             # the whole integer table
    assert abs(x.sum() - g['x_sum'][0]) < 1e-8
    assert abs(yu.sum() - g['yu_sum'][0]) < 1e-8


def test_rtnorm_vs_reference_same_draws():
    g = load_golden('ref_rtnorm')
    cases, ref, used = g['cases'], g['out'], g['used']
    for i, (a, b, mu, sg) in enumerate(cases):
        st = streams.PhiloxStream(seed=int(g['seed']), chain=int(g['chain']))
        st.begin_site(int(g['sweep']), i)
        r = rtnorm_port.rtnorm(a, b, mu=mu, sigma=sg, rng=st)
        assert st.k == used[i], (i, a, b, mu, sg)          # same branch, same #draws
        # tables are regenerated to the reference's printing precision (5e-12)
        assert abs(r - ref[i]) <= 1e-9 * max(1.0, abs(ref[i])), (i, r, ref[i])
        assert a <= r <= b


def test_mat_known_answer():
    """forward-modelling the ground-truth parameters of the
    reference's .mat fixture gives chi^2/N = 0.999; off by one channel: 3.6."""
    g = load_golden('mat_kat')
    data, var, fsf, params = g['data'], g['variance'], g['fsf'], g['params']
    D = data.shape[0]
    delta = port.gaussian_lsf_vector(0.0, 1.25e-4, D)       # tests/read_mat.py:93-94
    mask = np.ones(data.shape[1:])
    err = port.compute_error_in_one_step(data, params, fsf, delta, mask)
    chi2 = np.sum(err ** 2 / var) / err.size
    assert abs(chi2 - 0.999) < 2e-3
    shifted = params.copy()
    shifted[:, :, 1] += 1
    err1 = port.compute_error_in_one_step(data, shifted, fsf, delta, mask)
    assert np.sum(err1 ** 2 / var) / err1.size > 3.0
    # the O(HW * DHW) simulator is the same cube (lib/run.py:623-652 vs 999-1031)
    sim = port.simulate_convolved(data.shape, params, mask, fsf, delta)
    np.testing.assert_allclose(data - sim, err, rtol=0, atol=1e-12)
