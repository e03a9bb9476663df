"""
The ``Run`` drop-in on the GPU, written after the reference's own tests
(tests/run_test.py).  ``-m gpu``.
"""
import os

import numpy as np
import pytest

from conftest import load_golden

pytestmark = pytest.mark.gpu


def _muse_cube():
    from deconv3d_b200 import MUSE
    data = load_golden('muse_cube_01')['data'] * 1e20     # lib/run.py:140-143 needs > 1e-10
    return MUSE().build_cube(data)


def test_init_with_empty_cube():
    """tests/run_test.py:38-45"""
    from deconv3d_b200 import Run, MUSE, Cube
    cube = Cube()
    assert cube.is_empty()
    with pytest.raises(ValueError):
        Run(cube, MUSE())


def test_argument_errors():
    from deconv3d_b200 import Run, MUSE, Cube
    cube = _muse_cube()
    with pytest.raises(TypeError):
        Run(np.zeros((3, 3, 3)), MUSE())                    # not a Cube, lib/run.py:134
    with pytest.raises(TypeError):
        Run(cube, instrument='MUSE')                        # lib/run.py:204
    with pytest.raises(TypeError):
        Run(cube, MUSE(), variance=3.0)                     # lib/run.py:184
    with pytest.raises(ValueError):
        Run(cube, MUSE(), variance=np.ones((3, 3, 3)))      # lib/run.py:196
    with pytest.raises(AssertionError):
        Run(cube, MUSE(), max_iterations=0)                 # lib/run.py:114
    with pytest.raises(AssertionError):
        Run(Cube(data=cube.data * 1e-20, meta=cube.meta), MUSE())   # lib/run.py:141
    with pytest.raises(ValueError):
        Run(cube, MUSE(), initial_parameters=np.zeros((3, 3, 3)))   # lib/run.py:302
    from deconv3d_b200 import ImageFieldSpreadFunction
    with pytest.raises(ValueError):
        Run(cube, MUSE(fsf=ImageFieldSpreadFunction(np.ones((4, 5)) / 20.)))  # lib/run.py:211


def test_init_with_muse_cube_and_save(tmp_path):
    """tests/run_test.py:47-71 (cfg1: bundled cube, MUSE() defaults)."""
    from deconv3d_b200 import Run, MUSE
    cube = _muse_cube()
    run = Run(cube, instrument=MUSE(), max_iterations=200, seed=1)
    assert run.chain.shape == (200, 30, 30, 3)
    assert run.likelihoods.shape == (200, 30, 30)
    assert run.parameters.shape == (30, 30, 3)
    assert run.convolved_cube.data.shape == cube.data.shape
    assert np.isfinite(run.chain).all()
    # the fit explains the data: chi^2 per voxel drops to O(1)
    res = cube.data - run.convolved_cube.data
    assert np.mean(res ** 2 / run.variance_cube) < 5.0
    name = str(tmp_path / 'run_test')
    run.save(name, clobber=True)
    for suffix in ('_parameters.npy', '_convolved_cube.fits', '_clean_cube.fits'):
        assert os.path.isfile(name + suffix)
    from deconv3d_b200 import Cube
    back = Cube.from_fits(name + '_convolved_cube.fits')
    assert np.array_equal(back.data, run.convolved_cube.data)


def test_initial_parameters():
    """tests/run_test.py:73-98: max_iterations=1 => extract_parameters == initial."""
    from deconv3d_b200 import Run, MUSE, SingleGaussianLineModel
    cube = _muse_cube()
    m = SingleGaussianLineModel()

    class R(object):
        pass
    r = R()
    r.cube = cube
    r.fsf = MUSE().fsf.as_image(cube)
    minp = np.array(m.min_boundaries(r))
    maxp = np.array(m.max_boundaries(r))
    p1d = minp + (maxp - minp) * np.random.RandomState(0).rand(3)
    p3d = np.resize(p1d, (cube.shape[1], cube.shape[2], 3))
    run = Run(cube, MUSE(), initial_parameters=p3d, max_iterations=1)
    assert (run.extract_parameters() == p3d).all()
    np.save('test_ip.npy', p3d)
    try:
        run = Run(cube, MUSE(), initial_parameters='test_ip.npy', max_iterations=1)
        assert (run.extract_parameters() == p3d).all()
    finally:
        os.remove('test_ip.npy')


def test_masks():
    """tests/run_test.py:108-116"""
    from deconv3d_b200 import Run, MUSE, above_percentile
    cube = _muse_cube()
    mask = above_percentile(cube)
    run = Run(cube, MUSE(), mask=mask, max_iterations=42, seed=3)
    assert run.mask is mask
    on = mask == 1
    assert on.sum() == int(round(0.7 * 900))
    # spaxels outside the mask never move and never contribute
    assert np.array_equal(run.chain[-1][~on], run.chain[0][~on])
    assert np.abs(run.clean_cube.data[:, ~on]).max() == 0.0


def test_run_matches_oracle_chain():
    """Whole-``Run`` parity in sequential mode with the oracle fed by the same stream."""
    from deconv3d_b200 import Run, MUSE, rtnorm_tables
    from oracle import reference_port as port, streams
    g = load_golden('ref_run_A')
    inst = MUSE(fsf_fwhm=0.5)
    cube = inst.build_cube(g['data'])
    run = Run(cube, inst, max_iterations=10, seed=77)
    assert np.array_equal(run.fsf, g['fsf']) and np.array_equal(run.lsf, g['lsf'])
    x, yu, nc = rtnorm_tables.tables()
    ref = port.run_chain(g['data'], g['fsf'], g['lsf'], streams.PhiloxStream(77, 0),
                         max_iterations=10, rtnorm_tables=(x, yu, nc.astype(np.int64)))
    assert np.array_equal(run.variance_cube, ref['variance_cube'])
    np.testing.assert_allclose(run.chain, ref['chain'], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(run.parameters, ref['parameters'], rtol=1e-9, atol=1e-12)
    conv = port.simulate_convolved(g['data'].shape, ref['parameters'], ref['mask'], g['fsf'], g['lsf'])
    np.testing.assert_allclose(run.convolved_cube.data, conv, rtol=1e-9, atol=1e-12)
    c, _ = run.contribution_of_spaxel(3, 4, ref['parameters'][4, 3], 10, 9, 12, run.fsf, run.lsf)
    cr, _ = port.contribution_of_spaxel(3, 4, ref['parameters'][4, 3], 10, 9, 12, g['fsf'], g['lsf'])
    np.testing.assert_allclose(c, cr, rtol=1e-12, atol=1e-15)


def test_custom_model_is_refused():
    from deconv3d_b200 import Run, MUSE, SingleGaussianLineModel

    class Lorentz(SingleGaussianLineModel):
        def modelize(self, runner, x, parameters):
            return parameters[0] / (1 + ((np.asarray(x) - parameters[1]) / parameters[2]) ** 2)
    with pytest.raises(NotImplementedError):
        Run(_muse_cube(), MUSE(), model=Lorentz, max_iterations=2)


def test_coloured_mode_and_multi_chain_run():
    from deconv3d_b200 import Run, MUSE
    cube = _muse_cube()
    run = Run(cube, MUSE(), max_iterations=60, seed=9, mode='coloured', n_chains=3)
    assert run.chains.shape == (3, 60, 30, 30, 3)
    res = cube.data - run.convolved_cube.data
    assert np.mean(res ** 2 / run.variance_cube) < 5.0
    assert not np.array_equal(run.chains[0, -1], run.chains[1, -1])


def test_chain_kept_on_device_gives_the_same_outputs():
    """``chain_on_device=True`` keeps the chain in HBM, the posterior mean
    (extract_parameters, lib/run.py:581-593) is reduced on the device; ``run.chain`` is copied to
    the host only when asked for."""
    from deconv3d_b200 import Run, MUSE
    cube = _muse_cube()
    mask = np.ones((30, 30))
    mask[:3, :] = 0
    kw = dict(max_iterations=40, keep_one_in=2, seed=11, n_chains=2, mask=mask)
    host = Run(cube, MUSE(), **kw)
    dev = Run(cube, MUSE(), chain_on_device=True, **kw)
    assert 'chain' not in dev.__dict__
    np.testing.assert_allclose(dev.parameters, host.parameters, rtol=1e-13, atol=1e-13)
    np.testing.assert_allclose(dev.convolved_cube.data, host.convolved_cube.data, rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(dev.extract_parameters(50.), host.extract_parameters(50.), rtol=1e-13, atol=1e-13)
    assert np.array_equal(dev.chain, host.chain)                    # fetched now
    assert np.array_equal(dev.chains, host.chains)
    assert np.array_equal(dev.likelihoods[1:], host.likelihoods[1:])
    assert np.array_equal(dev.extract_parameters(), host.extract_parameters())   # host path after the fetch


def test_chain_mean_entry_point_host_and_device_buffers():
    import torch
    from deconv3d_b200 import _native
    rs = np.random.RandomState(0)
    ctx = _native.Context(0)
    D, H, W = 6, 5, 7
    ctx.set_problem(rs.rand(1, D, H, W) + 1, np.ones(1), np.ones((3, 3)) / 9., None, np.zeros((1, 3)),
                    np.array([[9., D - 1, D]]), [0, .1, .1], np.ones(1), chains_per_cube=2)
    chain = rs.randn(2, 11, H, W, 3)
    ref = chain[:, 4:].mean(axis=1)
    np.testing.assert_allclose(ctx.chain_mean(chain, 4), ref, rtol=1e-13, atol=1e-14)
    np.testing.assert_allclose(ctx.chain_mean(torch.from_numpy(chain).cuda(), 4), ref, rtol=1e-13, atol=1e-14)
    with pytest.raises(_native.NativeError):
        ctx.chain_mean(chain, 11)


@pytest.mark.parametrize('min_rate', [0.0, 0.9])
def test_page_locked_chain_rows_equal_ordinary_ones(monkeypatch, min_rate):
    """``Run`` keeps the chain rows in page-locked host memory (not zeroed when no chain can stop).
    Same chains, likelihoods and parameters as with ordinary numpy arrays, bit for bit -- also when
    every chain stops early on min_acceptance_rate and the unwritten rows must read zero
    (lib/run.py:270-271, 344-359), and with spaxels off the mask."""
    from deconv3d_b200 import Run, MUSE
    cube = _muse_cube()
    mask = np.ones((30, 30))
    mask[3:9, 10:20] = 0
    out = []
    for off in (None, '1'):
        if off:
            monkeypatch.setenv('D3D_NO_PINNED_ROWS', off)
        else:
            monkeypatch.delenv('D3D_NO_PINNED_ROWS', raising=False)
        run = Run(cube, instrument=MUSE(), mask=mask.copy(), max_iterations=60, seed=3, n_chains=3,
                  min_acceptance_rate=min_rate)
        assert run.chains.nbytes > (1 << 20)                       # above the pinning threshold
        out.append((run.chains.copy(), run.all_likelihoods.copy(), run.parameters.copy(),
                    np.array(run.iterations_done)))
    for a, b in zip(out[0], out[1]):
        assert np.array_equal(a, b)
    if min_rate > 0:
        assert (out[0][3] < 60).all()                              # the chains did stop ...
        assert not out[0][0][:, -1, mask == 1].any()               # ... and their last rows are zeros
