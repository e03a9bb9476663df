"""
GPU tests (``-m gpu``) of the edge cases of the hot path: NaN voxels (nansum semantics of
lib/run.py:24-27, 420-425), degenerate shapes, empty masks, early stop on the acceptance rate
(lib/run.py:344-359), row bookkeeping with keep_one_in.
"""
import numpy as np
import pytest

from test_gpu_parity import make_ctx, synthetic, _oracle, _tables   # noqa: F401

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def nat():
    import torch
    if not torch.cuda.is_available():
        pytest.fail('these tests need a CUDA device (there is no CPU fallback)')
    from deconv3d_b200 import _native
    return _native


def test_nan_voxels_drop_out_of_the_window_sums(nat):
    """NaN data voxels inside a window contribute nothing (the reference sums with nansum); the
    spaxels that contain them are masked out by Run (lib/run.py:162)."""
    port, _, _ = _oracle()
    D, H, W = 14, 11, 12
    rs = np.random.RandomState(2)
    data = synthetic(D, H, W, 4)
    nan_at = [(3, 4, 5), (0, 4, 5), (7, 9, 2)]
    for z, y, x in nan_at:
        data[z, y, x] = np.nan
    fsf = port.gaussian_fsf_image(0.6, 0.2)                     # 9x9
    lsf = port.gaussian_lsf_vector(0.0002675, 1.25e-4, D)
    var = 0.05 ** 2 * (1 + rs.rand(D, H, W))
    mask = np.ones((H, W))
    mask[np.isnan(data.sum(0))] = 0
    params = np.dstack([rs.rand(H, W) * 6, 3 + rs.rand(H, W) * 8, 0.8 + rs.rand(H, W) * 2])
    clean = np.where(np.isnan(data), 0.0, data)
    pmin, pmax = port.single_gaussian_boundaries(clean, fsf)
    ctx = nat.Context(0)
    ctx.set_rtnorm_tables(*_tables())
    ctx.set_rng(5, 0)
    ctx.set_problem(data, var, fsf, lsf, pmin, pmax, (0., .1, .1), float(pmax[0]) ** 2, mask=mask)
    ctx.set_params(params[None])
    ctx.forward(write_err=True)
    err_old = port.compute_error_in_one_step(data, params, fsf, lsf, mask)      # NaN where data is
    fhh = (fsf.shape[0] - 1) // 2
    for (y, x) in [(4, 6), (5, 5), (9, 3), (2, 2)]:
        p_new = params[y, x] + np.array([0.0, 0.4, -0.2])
        c_old, _ = port.contribution_of_spaxel(x, y, params[y, x], W, H, D, fsf, lsf)
        c_new, _ = port.contribution_of_spaxel(x, y, p_new, W, H, D, fsf, lsf)
        err_new = (err_old + c_old) - c_new
        sl = (slice(None), slice(max(y - fhh, 0), min(y + fhh + 1, H)),
              slice(max(x - fhh, 0), min(x + fhh + 1, W)))
        ar_old = 0.5 * np.nansum(err_old[sl] ** 2 / var[sl])
        ar_new = 0.5 * np.nansum(err_new[sl] ** 2 / var[sl])
        out = ctx.delta_logl(0, y, x, p_new)
        assert abs(out[0] - (ar_old - ar_new)) <= 1e-6 * abs(ar_old - ar_new) + 1e-11 * ar_old
        assert abs(out[1] - ar_old) <= 1e-10 * ar_old
    # sweeps stay finite, the masked spaxels are never touched, the residual matches its forward model
    chain = np.zeros((1, 6, H, W, 3))
    acc, its, _ = ctx.sweep(1, 5, chain_out=chain, min_acceptance_rate=0.0)
    assert np.isfinite(chain).all() and its[0] == 6
    assert (chain[0, 1:][:, mask == 0] == 0).all()
    res = ctx.get_residual()[0]
    ctx.forward(write_err=True)
    np.testing.assert_allclose(res, ctx.get_residual()[0], rtol=0, atol=1e-9)


def test_run_refuses_nan_cubes_like_the_reference():
    """np.max of a cube with a NaN is NaN, so the reference's signal assertion (lib/run.py:140-143)
    rejects it; the drop-in keeps that behaviour."""
    from deconv3d_b200 import Run, MUSE
    data = synthetic(16, 12, 13, 7) + 1.0
    data[5, 6, 6] = np.nan
    with pytest.raises(AssertionError):
        Run(MUSE().build_cube(data), MUSE(fsf_fwhm=0.6), max_iterations=12, seed=3)


@pytest.mark.parametrize('shape', [(2, 1, 9), (5, 8, 1), (1, 4, 4), (3, 1, 1)])
def test_degenerate_shapes_vs_oracle(nat, shape):
    """One-row / one-column / one-channel cubes: every window is clipped to the field."""
    from test_gpu_parity import _compare_chain
    port, _, _ = _oracle()
    D, H, W = shape
    rs = np.random.RandomState(6)
    data = 2.0 + rs.rand(D, H, W)
    fsf = port.gaussian_fsf_image(0.4, 0.2)                     # 5x5
    lsf = port.gaussian_lsf_vector(0.0002675, 1.25e-4, D)
    init = np.dstack([1 + rs.rand(H, W), rs.rand(H, W) * max(D - 1, 0), 0.5 + rs.rand(H, W)])
    var = np.array([0.3 ** 2])
    _compare_chain(nat, data, fsf, lsf, var, None, init, 5, 1, seed=3)
    _compare_chain(nat, data, fsf, lsf, var, None, init, 4, 1, seed=4, mode='colour')


def test_empty_mask_is_a_no_op(nat):
    port, _, _ = _oracle()
    D, H, W = 8, 6, 7
    data = synthetic(D, H, W, 1)
    fsf = port.gaussian_fsf_image(0.4, 0.2)
    lsf = port.gaussian_lsf_vector(0.0002675, 1.25e-4, D)
    ctx, _, _ = make_ctx(nat, data, np.array([0.01]), fsf, lsf, mask=np.zeros((H, W)))
    ctx.init_params_uniform()
    ctx.forward(write_err=True)
    chain = np.zeros((1, 4, H, W, 3))
    for mode in (nat.SEQ_EXACT, nat.COLOURED):
        acc, its, _ = ctx.sweep(1, 3, mode=mode, chain_out=chain, min_acceptance_rate=0.0)
        assert acc[0] == 0 and (chain == 0).all()
    np.testing.assert_allclose(ctx.get_residual()[0], data, rtol=0, atol=1e-12)   # no source at all


def test_early_stop_on_acceptance_rate_matches_oracle(nat):
    """A chain stops when accepted/(spaxels*iteration) falls to min_acceptance_rate
    (lib/run.py:344-359): same iteration count as the oracle, later rows untouched."""
    port, streams, _ = _oracle()
    D, H, W = 10, 7, 8
    rs = np.random.RandomState(9)
    data = synthetic(D, H, W, 3)
    fsf = port.gaussian_fsf_image(0.4, 0.2)
    lsf = port.gaussian_lsf_vector(0.0002675, 1.25e-4, D)
    init = np.dstack([rs.rand(H, W) * 4, 2 + rs.rand(H, W) * 5, 0.7 + rs.rand(H, W) * 2])
    var = np.full(data.shape, 0.05 ** 2)
    for min_rate in (0.9, 0.75):
        ref = port.run_chain(data, fsf, lsf, streams.PhiloxStream(23, 0), variance_cube=var,
                             initial_parameters=init, max_iterations=40, min_acceptance_rate=min_rate,
                             rtnorm_tables=_tables())
        ctx, _, _ = make_ctx(nat, data, var, fsf, lsf, seed=23, prior=ref['gibbs_apriori_variance'])
        ctx.set_params(init[None])
        ctx.forward(write_err=True)
        chain = np.zeros((1, 40, H, W, 3))
        acc, its, _ = ctx.sweep(1, 39, chain_out=chain, min_acceptance_rate=min_rate)
        assert its[0] == ref['iterations'] and acc[0] == ref['accepted_count']
        assert ref['iterations'] < 40, 'the case must actually stop early'
        assert (chain[0, its[0]:] == 0).all()
        np.testing.assert_allclose(chain[0, 1:its[0]], ref['chain'][1:its[0]], rtol=1e-9, atol=1e-9)


def test_keep_one_in_rows_across_calls(nat):
    """Rows it // keep for it % keep == 0, also when the iterations are split over several calls."""
    port, _, _ = _oracle()
    D, H, W = 8, 6, 7
    data = synthetic(D, H, W, 1)
    fsf = port.gaussian_fsf_image(0.4, 0.2)
    lsf = port.gaussian_lsf_vector(0.0002675, 1.25e-4, D)
    out = []
    for splits in ([20], [3, 4, 6, 7], [1] * 20):
        ctx, _, _ = make_ctx(nat, data, np.array([0.01]), fsf, lsf, seed=6)
        ctx.init_params_uniform()
        ctx.forward(write_err=True)
        chain = np.zeros((1, 5, H, W, 3))
        it = 1
        for n in splits:
            ctx.sweep(it, n, keep_one_in=5, chain_out=chain, min_acceptance_rate=0.0)
            it += n
        out.append(chain)
    assert (out[0][0, 1:] != 0).any()
    assert np.array_equal(out[0], out[1]) and np.array_equal(out[0], out[2])
    with pytest.raises(nat.NativeError):
        ctx.sweep(21, 5, keep_one_in=5, chain_out=chain)         # iteration 25 needs row 5
