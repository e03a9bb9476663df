"""
GPU parity tests (run with ``-m gpu`` on a B200): the CUDA path, called through
the C ABI (deconv3d_b200._native / the ``Run`` drop-in), against the CPU oracle
on the same seeded inputs and against the golden vectors made by the
reference's own code.

Tolerances (BASELINE.json north_star): convolved cubes 1e-12 relative (fp64) /
1e-5 (fp32); per-proposal delta-logL 1e-6 relative; identical accept/reject
decisions in sequential-exact mode given the same uniform stream.
"""
import numpy as np
import pytest

from conftest import load_golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def nat():
    import torch
    if not torch.cuda.is_available():
        pytest.fail('these tests need a CUDA device (there is no CPU fallback)')
    from deconv3d_b200 import _native
    return _native


def _oracle():
    from oracle import reference_port, streams, rtnorm_port
    return reference_port, streams, rtnorm_port


def _tables():
    from deconv3d_b200 import rtnorm_tables
    x, yu, nc = rtnorm_tables.tables()
    return x, yu, nc.astype(np.int64)


def synthetic(D, H, W, seed, noise=0.05):
    rs = np.random.RandomState(seed)
    yy, xx = np.mgrid[0:H, 0:W]
    a = 8.0 * np.exp(-((yy - H / 2.) ** 2 + (xx - W / 2.) ** 2) / (2. * (0.35 * H) ** 2))
    c = D / 2. + 0.15 * D * np.tanh((xx - W / 2.) / (0.3 * W))
    w = 1.2 + 0.05 * yy
    z = np.arange(D)[:, None, None]
    return a * np.exp(-(z - c) ** 2 / (2 * w ** 2)) + noise * rs.randn(D, H, W)


def make_ctx(nat, data, var, fsf, lsf, mask=None, dtype=None, chains=1, seed=42, first_chain=0,
             jump=(0.0, 0.1, 0.1), prior=None):
    port, _, _ = _oracle()
    pmin, pmax = port.single_gaussian_boundaries(data, fsf)
    ctx = nat.Context(0, nat.F64 if dtype is None else dtype)
    ctx.set_rtnorm_tables(*_tables())
    ctx.set_rng(seed, first_chain)
    if prior is None:
        prior = float(pmax[0]) ** 2
    ctx.set_problem(data, var, fsf, lsf, pmin, pmax, jump, prior, mask=mask,
                    chains_per_cube=chains)
    return ctx, np.array(pmin, float), np.array(pmax, float)


# ---------------------------------------------------------------------------
def test_conv1d_vs_reference_golden(nat):
    g = load_golden('ref_conv1d')
    ctx = nat.Context(0, nat.F64)
    for D in (2, 3, 8, 16, 21, 30, 31, 32, 33, 40, 41, 63, 64, 65):
        out = ctx.conv1d(g['line_%d' % D], g['lsf_%d' % D])
        np.testing.assert_allclose(out, g['out_%d' % D], rtol=1e-12, atol=1e-15)
        out = ctx.conv1d(g['line_%d' % D], g['lsfrand_%d' % D])
        np.testing.assert_allclose(out, g['outrand_%d' % D], rtol=1e-12, atol=1e-14)
    # batched + the drop-in function with the reference's (conv, fftpsf) protocol
    from deconv3d_b200 import convolve_1d
    rs = np.random.RandomState(0)
    lines = rs.rand(5, 7, 30)
    lsf = g['lsf_30']
    out, fftpsf = convolve_1d(lines, lsf, axis=2)
    port, _, _ = _oracle()
    for i in range(5):
        for j in range(7):
            np.testing.assert_allclose(out[i, j], port.convolve_1d(lines[i, j], lsf)[0],
                                       rtol=1e-12, atol=1e-15)
    out2, _ = convolve_1d(lines[0, 0], fftpsf, compute_fourier=False)
    np.testing.assert_allclose(out2, out[0, 0], rtol=1e-12, atol=1e-15)
    assert np.allclose(fftpsf, port.convolve_1d(lines[0, 0], lsf)[1])


def test_rtnorm_vs_reference_golden(nat):
    g = load_golden('ref_rtnorm')
    ctx = nat.Context(0, nat.F64)
    ctx.set_rtnorm_tables(*_tables())
    c = g['cases']
    out, used = ctx.rtnorm_batch(c[:, 0], c[:, 1], c[:, 2], c[:, 3], seed=int(g['seed']),
                                 chain=int(g['chain']), sweep=int(g['sweep']))
    assert np.array_equal(used, g['used'])               # same branch, same number of draws
    np.testing.assert_allclose(out, g['out'], rtol=1e-9, atol=1e-9)
    assert np.all(out >= c[:, 0]) and np.all(out <= c[:, 1])


def test_rtnorm_moments(nat):
    from deconv3d_b200 import rtnorm
    r = rtnorm(1., 17., mu=7., sigma=5., size=200000, seed=5)     # tests/rtnorm_test.py:16-34
    assert r.min() >= 1. and r.max() <= 17.
    from scipy.stats import truncnorm
    tn = truncnorm((1. - 7.) / 5., (17. - 7.) / 5., loc=7., scale=5.)
    assert abs(r.mean() - tn.mean()) < 0.05
    assert abs(r.std() - tn.std()) < 0.05
    r = rtnorm(0, 2 ** 63 - 1, size=42, seed=1)                    # tests/rtnorm_test.py:36-56
    assert isinstance(r, np.ndarray) and len(r) == 42 and (r > 0).all()
    with pytest.raises(Exception):
        rtnorm(2., 1.)


@pytest.mark.parametrize('dtype_name,rtol', [('f64', 1e-12), ('f32', 1e-5)])
def test_forward_mat_known_answer(nat, dtype_name, rtol):
    port, _, _ = _oracle()
    g = load_golden('mat_kat')
    data, var, fsf, params = g['data'], g['variance'], g['fsf'], g['params']
    delta = port.gaussian_lsf_vector(0.0, 1.25e-4, data.shape[0])
    ctx, _, _ = make_ctx(nat, data, var, fsf, delta,
                         dtype=nat.F64 if dtype_name == 'f64' else nat.F32)
    ctx.set_params(params[None])
    sim, chi2 = ctx.forward(want_sim=True, write_err=True, want_chi2=True)
    mask = np.ones(data.shape[1:])
    err_ref = port.compute_error_in_one_step(data, params, fsf, delta, mask)
    scale = np.abs(data - err_ref).max()
    np.testing.assert_allclose(sim[0], data - err_ref, rtol=1e-12, atol=1e-12 * scale)
    np.testing.assert_allclose(ctx.get_residual()[0], err_ref, rtol=rtol,
                               atol=rtol * np.abs(data).max())
    chi2_ref = 0.5 * np.sum(err_ref ** 2 / var)
    assert abs(chi2[0] - chi2_ref) <= max(rtol, 1e-10) * chi2_ref * 10
    assert abs(2 * chi2[0] / data.size - 0.999) < 2e-3


@pytest.mark.parametrize('shape,fsf_kind,lsf_fwhm', [
    ((30, 11, 12), 'gauss13', 0.0002675),      # D=30: spectral wrap
    ((32, 9, 9), 'moffat7', 0.0004),           # P=32 full wrap
    ((40, 16, 14), 'ell9', 0.0002675),
    ((21, 8, 15), 'rect5x9', None),            # no LSF (lib/run.py:675-676), non-square FSF
    ((64, 6, 7), 'gauss13', 0.0006),
    ((10, 37, 45), 'moffat23', 0.0002675),     # widths without a dedicated instantiation: the
    ((6, 20, 70), 'rect31x41', 0.0002675),     # chunked wide stencil (partial last chunk, 2 tiles)
    ((7, 18, 19), 'rect9x19', None),
    ((70, 6, 9), 'moffat7', 0.0006),           # P=128: four channels per lane in the warp-shuffle pass
    ((128, 5, 6), 'moffat7', 0.0004),          # P=128 full wrap
    ((130, 5, 6), 'moffat7', 0.0004),          # P=256: shared-memory spectral kernel
])
def test_forward_random_params_vs_oracle(nat, shape, fsf_kind, lsf_fwhm):
    port, _, _ = _oracle()
    D, H, W = shape
    rs = np.random.RandomState(D + H)
    step = 0.2
    fsf = {
        'gauss13': lambda: port.gaussian_fsf_image(1.0, step),
        'moffat7': lambda: port.moffat_fsf_image((7, 7), step, fwhm_arcsec=0.8, beta=2.5),
        'ell9': lambda: port.gaussian_fsf_image(0.6, step, pa=30., ba=0.7),
        'rect5x9': lambda: (lambda f: f / f.sum())(rs.rand(5, 9)),
        'moffat23': lambda: port.moffat_fsf_image((23, 23), step, fwhm_arcsec=1.2, beta=2.5),
        'rect31x41': lambda: (lambda f: f / f.sum())(rs.rand(31, 41)),
        'rect9x19': lambda: (lambda f: f / f.sum())(rs.rand(9, 19)),
    }[fsf_kind]()
    lsf = None if lsf_fwhm is None else port.gaussian_lsf_vector(lsf_fwhm, 1.25e-4, D)
    data = synthetic(D, H, W, 3)
    mask = (rs.rand(H, W) > 0.2).astype(float)
    params = np.dstack([rs.rand(H, W) * 9, rs.rand(H, W) * (D - 1), 0.3 + rs.rand(H, W) * 4])
    ctx, _, _ = make_ctx(nat, data, np.array([0.01]), fsf, lsf, mask=mask)
    ctx.set_params(params[None])
    sim, _ = ctx.forward(want_sim=True, write_err=True)
    err_ref = port.compute_error_in_one_step(data, params, fsf, lsf, mask)
    sim_ref = data - err_ref
    np.testing.assert_allclose(sim[0], sim_ref, rtol=1e-12, atol=1e-12 * np.abs(sim_ref).max())
    np.testing.assert_allclose(ctx.get_residual()[0], err_ref, rtol=1e-12,
                               atol=1e-12 * np.abs(data).max())
    # simulate_clean / simulate with explicit parameters
    np.testing.assert_allclose(ctx.simulate(params[None])[0], sim_ref, rtol=1e-12,
                               atol=1e-12 * np.abs(sim_ref).max())
    np.testing.assert_allclose(ctx.simulate_clean(params[None])[0],
                               port.simulate_clean(data.shape, params, mask),
                               rtol=1e-12, atol=1e-300)


@pytest.mark.parametrize('D', [12, 16, 30, 32, 40, 64, 100])
def test_spectral_pass_variants_agree(nat, monkeypatch, D):
    """Three spectral passes: one thread per spaxel with the spectrum in registers
    (lines_lane_kernel, the default where it is instantiated), one warp per spaxel
    (lines_warp_kernel, registers + __shfl) and the shared-memory one (lines_kernel).  The last
    two keep the same summation order: bit-identical cubes; the first one sums the taps in
    window order and uses a table-driven exp: equal to 1e-13 of the cube's scale."""
    port, _, _ = _oracle()
    H, W = 7, 9
    rs = np.random.RandomState(D)
    fsf = port.moffat_fsf_image((5, 5), 0.2, fwhm_arcsec=0.8, beta=2.5)
    lsf = port.gaussian_lsf_vector(0.0005 if D > 16 else 0.00025, 1.25e-4, D)
    data = synthetic(D, H, W, 5)
    mask = (rs.rand(H, W) > 0.2).astype(float)
    params = np.dstack([rs.rand(H, W) * 9, rs.rand(H, W) * (D - 1), 0.3 + rs.rand(H, W) * 4])
    params[0, 0] = [3.0, D / 2., 0.02]                # a line far narrower than a channel
    params[0, 1] = [3.0, D - 1.0, 40.0]               # and one far wider than the cube
    out = []
    for env in (None, 'D3D_LINES_WARP', 'D3D_LINES_SMEM'):
        if env:
            monkeypatch.setenv(env, '1')
        ctx, _, _ = make_ctx(nat, data, np.array([0.01]), fsf, lsf, mask=mask)
        ctx.set_params(params[None])
        sim, _ = ctx.forward(want_sim=True, write_err=True)
        out.append((sim.copy(), ctx.get_residual().copy()))
        ctx.close()
    assert np.array_equal(out[1][0], out[2][0])
    assert np.array_equal(out[1][1], out[2][1])
    scale = np.abs(out[2][0]).max()
    np.testing.assert_allclose(out[0][0], out[2][0], rtol=0, atol=1e-13 * scale)
    np.testing.assert_allclose(out[0][1], out[2][1], rtol=0, atol=1e-13 * scale)
    err_ref = port.compute_error_in_one_step(data, params, fsf, lsf, mask)
    np.testing.assert_allclose(out[0][1][0], err_ref, rtol=0, atol=1e-12 * max(scale, np.abs(data).max()))


@pytest.mark.parametrize('var_kind', ['scalar', 'cube'])
def test_delta_logl_vs_oracle(nat, var_kind):
    """Per-proposal (delta, ar_old, ar_new) against the reference's windowed sums
    (lib/run.py:400-426) on the same residual."""
    port, _, _ = _oracle()
    D, H, W = 30, 12, 13
    rs = np.random.RandomState(8)
    data = synthetic(D, H, W, 5)
    fsf = port.gaussian_fsf_image(0.7, 0.2)                     # 9x9
    lsf = port.gaussian_lsf_vector(0.0002675, 1.25e-4, D)
    var = np.array([0.05 ** 2]) if var_kind == 'scalar' else 0.05 ** 2 * (1 + rs.rand(D, H, W))
    var_cube = np.ones((D, H, W)) * var if var_kind == 'scalar' else var
    mask = np.ones((H, W))
    params = np.dstack([rs.rand(H, W) * 9, 5 + rs.rand(H, W) * 20, 0.8 + rs.rand(H, W) * 3])
    ctx, _, _ = make_ctx(nat, data, var, fsf, lsf)
    ctx.set_params(params[None])
    ctx.forward(write_err=True)
    err_old = port.compute_error_in_one_step(data, params, fsf, lsf, mask)
    fhh = (fsf.shape[0] - 1) // 2
    for (y, x) in [(0, 0), (5, 6), (11, 12), (3, 12), (11, 0), (6, 1)]:
        for trial in range(3):
            p_new = params[y, x].copy()
            p_new[1] += 0.3 * rs.randn()
            p_new[2] += 0.2 * rs.randn()
            if trial == 2:
                p_new[0] *= 1.1                                  # amplitude may move too
            c_old, _ = port.contribution_of_spaxel(x, y, params[y, x], W, H, D, fsf, lsf)
            c_new, _ = port.contribution_of_spaxel(x, y, p_new, W, H, D, fsf, lsf)
            err_new = (err_old + c_old) - c_new
            sl = (slice(None), slice(max(y - fhh, 0), min(y + fhh + 1, H)),
                  slice(max(x - fhh, 0), min(x + fhh + 1, W)))
            ar_old = 0.5 * np.nansum(err_old[sl] ** 2 / var_cube[sl])
            ar_new = 0.5 * np.nansum(err_new[sl] ** 2 / var_cube[sl])
            out = ctx.delta_logl(0, y, x, p_new)
            delta = ar_old - ar_new
            # the reference value is a difference of two O(N/2) sums: allow its own
            # cancellation noise on top of the 1e-6 relative bar
            assert abs(out[0] - delta) <= 1e-6 * abs(delta) + 1e-11 * ar_old, (y, x, out, delta)
            assert abs(out[1] - ar_old) <= 1e-10 * ar_old
            assert abs(out[2] - ar_new) <= 1e-10 * ar_new + 1e-11 * ar_old


def _compare_chain(nat, data, fsf, lsf, var_in, mask_in, init, max_it, keep, seed, mode='seq',
                   jump=0.1, prior=None, dtype=None, rtol=1e-9):
    port, streams, _ = _oracle()
    tables = _tables()
    D, H, W = data.shape
    var_cube = None
    if var_in is not None:
        var_cube = var_in if var_in.ndim == 3 else np.ones(data.shape) * var_in[0]
    trace = {}
    order = None
    m_for_order = port.prepare_mask(data, None if mask_in is None else mask_in.copy())
    if mode == 'colour':
        order = port.colour_class_order(m_for_order, fsf.shape[0], fsf.shape[1])
    ref = port.run_chain(data, fsf, lsf, streams.PhiloxStream(seed, 0),
                         mask=None if mask_in is None else mask_in.copy(),
                         variance_cube=var_cube, initial_parameters=init,
                         jump_amplitude=jump, gibbs_apriori_variance=prior,
                         max_iterations=max_it, keep_one_in=keep, trace=trace,
                         rtnorm_tables=tables, site_order=order)
    var_dev = ref['variance_cube'] if var_in is None or var_in.ndim == 3 else var_in
    if var_in is None:
        var_dev = np.array([ref['variance_cube'].flat[0]])
    jump_vec = np.ones(3) * np.array(jump)
    jump_vec[0] = 0
    ctx, pmin, pmax = make_ctx(nat, data, var_dev, fsf, lsf, mask=ref['mask'], seed=seed,
                               jump=jump_vec, prior=ref['gibbs_apriori_variance'], dtype=dtype)
    n_saved = ref['chain'].shape[0]
    chain = np.zeros((1, n_saved, H, W, 3))
    lik = np.zeros((1, n_saved, H, W))
    if init is not None:
        ctx.set_params(np.asarray(init, float)[None])
    else:
        ctx.init_params_uniform()
    chain[0, 0] = ctx.get_params()[0]
    ctx.forward(write_err=True)
    acc, its, ms = ctx.sweep(1, max_it - 1, mode=nat.SEQ_EXACT if mode == 'seq' else nat.COLOURED,
                             keep_one_in=keep, chain_out=chain, lik_out=lik)
    m = ref['mask'] == 1
    assert its[0] == ref['iterations']
    # accept / reject decisions: identical, proposal by proposal
    n_acc_ref = sum(1 for v in trace.values() if v[3])
    assert acc[0] - m.sum() == n_acc_ref, (acc[0], n_acc_ref)
    if keep == 1:
        for it in range(1, max_it):
            moved = (chain[0, it, :, :, 1] != chain[0, it - 1, :, :, 1]) | \
                    (chain[0, it, :, :, 2] != chain[0, it - 1, :, :, 2])
            ref_acc = np.zeros((H, W), bool)
            for (y, x) in zip(*np.nonzero(m)):
                ref_acc[y, x] = trace[(it, y, x)][3]
            assert np.array_equal(moved & m, ref_acc), 'decisions differ at iteration %d' % it
    # amplitudes drawn next to the lower bound (a ~ 1e-5 when a_max ~ 1e3) carry the absolute
    # rounding of the posterior mean: tolerance relative to the parameter scale
    scale = np.abs(ref['chain'][:, m]).max()
    # Degenerate Gibbs draws: when a line is so narrow that its footprint vanishes (w << 1
    # channel) the posterior mean mu is pure rounding noise of either sign, and the table
    # sampler branches on floor(-mu/sigma * INVH) (lib/rtnorm.py:144): the draw is then not a
    # continuous function of the inputs.  Such draws (they have no effect on the residual)
    # are excluded from the value comparison; they must stay rare.
    ok = np.ones(ref['chain'].shape, bool)
    for (it, y, x), v in trace.items():
        mu, ro = v[5], v[6]
        if it % keep == 0 and abs(mu) <= 1e-9 * np.sqrt(ro):
            ok[it // keep, y, x, 0] = False
    assert (~ok).sum() <= 0.005 * ok.size
    sel = ok & m[None, :, :, None]
    np.testing.assert_allclose(chain[0][sel], ref['chain'][sel], rtol=rtol, atol=1e-12 * scale)
    np.testing.assert_allclose(lik[0][1:, m], ref['likelihoods'][1:, m], rtol=max(rtol, 1e-6),
                               atol=1e-9)
    np.testing.assert_allclose(ctx.get_residual()[0], ref['err'], rtol=0,
                               atol=1e-9 * np.abs(data).max())
    return ref, chain, lik


def test_seq_exact_small_scalar_variance(nat):
    """Case A of the goldens: 12x9x10, 7x7 Gaussian FSF, default (scalar) variance,
    random initial parameters drawn on the device (sweep 0 of the stream)."""
    g = load_golden('ref_run_A')
    _compare_chain(nat, g['data'], g['fsf'], g['lsf'], None, None, None, 25, 1, seed=7)


def test_seq_exact_mask_variance_cube_wrap(nat):
    """Case C: mask, variance cube, initial parameters (one amplitude = 0), keep_one_in,
    D=16 (full spectral wrap), rotated elliptical FSF, custom jump and Gibbs prior."""
    g = load_golden('ref_run_C')
    _compare_chain(nat, g['data'], g['fsf'], g['lsf'], g['in_variance'], g['in_mask'],
                   g['in_initial_parameters'], 12, 2, seed=3, jump=0.3, prior=50.0)
    _compare_chain(nat, g['data'], g['fsf'], g['lsf'], g['in_variance'], g['in_mask'],
                   g['in_initial_parameters'], 9, 1, seed=4, jump=0.3, prior=50.0)


def test_seq_exact_muse_cube(nat):
    """cfg1: the bundled MUSE cube x1e20, MUSE() defaults (13x13 FSF, D=30 wrap)."""
    g = load_golden('ref_run_B')
    _compare_chain(nat, g['data'], g['fsf'], g['lsf'], None, None, None, 4, 1, seed=11)


def test_seq_exact_crosses_refresh(nat):
    """> 1000 iterations: the residual refresh of lib/run.py:525-534 happens inside."""
    g = load_golden('ref_run_D')
    _compare_chain(nat, g['data'], g['fsf'], g['lsf'], None, None, None, 1003, 50, seed=13,
                   rtol=1e-7)


def test_coloured_mode_vs_oracle_same_order(nat):
    g = load_golden('ref_run_A')
    _compare_chain(nat, g['data'], g['fsf'], g['lsf'], None, None, None, 12, 1, seed=21,
                   mode='colour')
    g = load_golden('ref_run_C')
    _compare_chain(nat, g['data'], g['fsf'], g['lsf'], g['in_variance'], g['in_mask'],
                   g['in_initial_parameters'], 8, 1, seed=5, jump=0.3, prior=50.0, mode='colour')


@pytest.mark.parametrize('fsf_shape', [(13, 13), (23, 23)])
def test_coloured_chain_per_cta_equals_launch_per_class(nat, monkeypatch, fsf_shape):
    """Coloured mode has two schedules: one launch per colour class (few chains) and one CTA per
    chain walking the sites in colour-class order (many chains).  Same decisions, same chains."""
    port, _, _ = _oracle()
    rs = np.random.RandomState(12)
    D, H, W = 12, 18, 20
    data = synthetic(D, H, W, 5)
    fsf = port.moffat_fsf_image(fsf_shape, 0.2, fwhm_arcsec=0.8, beta=2.5)
    lsf = port.gaussian_lsf_vector(0.0002675, 1.25e-4, D)
    var = 0.05 ** 2 * (1 + rs.rand(D, H, W))
    mask = (rs.rand(H, W) > 0.2).astype(float)
    out = []
    for by_chain in ('0', '1'):
        monkeypatch.setenv('D3D_COLOUR_BY_CHAIN', by_chain)
        ctx, _, _ = make_ctx(nat, data, var, fsf, lsf, mask=mask, chains=3, seed=19)
        ctx.init_params_uniform()
        ctx.forward(write_err=True)
        chain = np.zeros((3, 5, H, W, 3))
        lik = np.zeros((3, 5, H, W))
        acc, its, _ = ctx.sweep(1, 4, mode=nat.COLOURED, refresh_every=2, min_acceptance_rate=0.0,
                                chain_out=chain, lik_out=lik)
        out.append((chain, lik, acc, its, ctx.get_residual()))
    m = mask == 1
    # same decisions and the same chains; the two schedules use different kernels (launch per
    # class: row-mapped / generic; chain per CTA: sliding-window kernel), i.e. different
    # summation orders, hence a rounding-level tolerance instead of bit equality
    assert np.array_equal(out[0][2], out[1][2]) and np.array_equal(out[0][3], out[1][3])
    scale = np.abs(out[0][0][:, :, m]).max()
    np.testing.assert_allclose(out[0][0][:, 1:][:, :, m], out[1][0][:, 1:][:, :, m], rtol=1e-8, atol=1e-10 * scale)
    np.testing.assert_allclose(out[0][1][:, 1:][:, :, m], out[1][1][:, 1:][:, :, m], rtol=1e-6, atol=1e-8)
    np.testing.assert_allclose(out[0][4], out[1][4], rtol=0, atol=1e-8 * np.abs(data).max())


def test_fp32_storage_chain_close(nat):
    """float32 storage: same decisions are not required; the chain must stay close over
    a few sweeps and the residual must match its own forward model."""
    g = load_golden('ref_run_A')
    port, streams, _ = _oracle()
    ref = port.run_chain(g['data'], g['fsf'], g['lsf'], streams.PhiloxStream(7, 0),
                         max_iterations=3, rtnorm_tables=_tables())
    ctx, _, _ = make_ctx(nat, g['data'], np.array([ref['variance_cube'].flat[0]]), g['fsf'],
                         g['lsf'], seed=7, dtype=nat.F32)
    ctx.init_params_uniform()
    ctx.forward(write_err=True)
    chain = np.zeros((1, 3, 9, 10, 3))
    ctx.sweep(1, 2, chain_out=chain)
    np.testing.assert_allclose(chain[0, 1], ref['chain'][1], rtol=2e-3, atol=2e-3)
    res = ctx.get_residual()[0]
    ctx.forward(write_err=True)
    np.testing.assert_allclose(res, ctx.get_residual()[0], rtol=0, atol=1e-4)


def test_many_chains_equal_single_chain_runs(nat):
    """Chain k of a multi-chain context is bit-identical to a single-chain context whose
    stream is (seed, first_chain_id = k): chains are independent units (no collective)."""
    g = load_golden('ref_run_A')
    data, fsf, lsf = g['data'], g['fsf'], g['lsf']
    var = np.array([0.01])
    out = []
    ctx, _, _ = make_ctx(nat, data, var, fsf, lsf, chains=5, seed=99)
    ctx.init_params_uniform()
    ctx.forward(write_err=True)
    chain = np.zeros((5, 6, 9, 10, 3))
    ctx.sweep(1, 5, chain_out=chain)
    for k in (0, 3, 4):
        c1, _, _ = make_ctx(nat, data, var, fsf, lsf, chains=1, seed=99, first_chain=k)
        c1.init_params_uniform()
        c1.forward(write_err=True)
        ch = np.zeros((1, 6, 9, 10, 3))
        c1.sweep(1, 5, chain_out=ch)
        ch[0, 0] = c1.get_params()[0] * 0 + ch[0, 0]
        assert np.array_equal(ch[0, 1:], chain[k, 1:])
    assert not np.array_equal(chain[0, 1:], chain[1, 1:])


def test_multi_cube_galaxies(nat):
    """Survey-batch layout (cfg5): independent cubes with their own data, variance,
    mask and boundaries in one context equal the per-cube contexts."""
    port, _, _ = _oracle()
    rs = np.random.RandomState(2)
    fsf = port.gaussian_fsf_image(0.5, 0.2)
    D, H, W = 16, 8, 9
    lsf = port.gaussian_lsf_vector(0.0002675, 1.25e-4, D)
    cubes = np.stack([synthetic(D, H, W, s) for s in (1, 2, 3)])
    var = 0.05 ** 2 * (1 + rs.rand(3, D, H, W))
    masks = (rs.rand(3, H, W) > 0.3).astype(np.uint8)
    pmin = np.zeros((3, 3))
    pmax = np.array([[cubes[i].max() / fsf.max(), D - 1, D] for i in range(3)])
    prior = pmax[:, 0] ** 2
    ctx = nat.Context(0, nat.F64)
    ctx.set_rtnorm_tables(*_tables())
    ctx.set_rng(5, 0)
    ctx.set_problem(cubes, var, fsf, lsf, pmin, pmax, [0, 0.1, 0.1], prior, mask=masks,
                    chains_per_cube=2)
    ctx.init_params_uniform()
    ctx.forward(write_err=True)
    chain = np.zeros((6, 4, H, W, 3))
    ctx.sweep(1, 3, chain_out=chain)
    for i in range(3):
        c1 = nat.Context(0, nat.F64)
        c1.set_rtnorm_tables(*_tables())
        c1.set_rng(5, 2 * i)
        c1.set_problem(cubes[i], var[i], fsf, lsf, pmin[i], pmax[i], [0, 0.1, 0.1], prior[i],
                       mask=masks[i], chains_per_cube=2)
        c1.init_params_uniform()
        c1.forward(write_err=True)
        ch = np.zeros((2, 4, H, W, 3))
        c1.sweep(1, 3, chain_out=ch)
        assert np.array_equal(ch[:, 1:], chain[2 * i:2 * i + 2, 1:])


@pytest.mark.parametrize('pipe', ['0', '1'])
def test_balanced_schedule_more_chains_than_sms(nat, monkeypatch, pipe):
    """More chains than SMs: the chain x sweep rectangle is laid over the SMs by the
    wrap-around rule and some chains are handed from one CTA to another mid-call.  Every
    chain must still equal its own single-chain run bit for bit -- with either sweep kernel
    (D3D_PIPE=0 sliding-window, 1 pipelined; two different kernels agree to rounding, not to
    the bit)."""
    monkeypatch.setenv('D3D_PIPE', pipe)
    g = load_golden('ref_run_A')
    data, fsf, lsf = g['data'], g['fsf'], g['lsf']
    var = np.array([0.01])
    n = 333
    ctx, _, _ = make_ctx(nat, data, var, fsf, lsf, chains=n, seed=5)
    ctx.init_params_uniform()
    ctx.forward(write_err=True)
    chain = np.zeros((n, 8, 9, 10, 3))
    acc, its, _ = ctx.sweep(1, 7, chain_out=chain)
    assert (its == 8).all()
    res = ctx.get_residual()
    for k in (0, 1, 147, 148, 149, 200, 332):
        c1, _, _ = make_ctx(nat, data, var, fsf, lsf, chains=1, seed=5, first_chain=k)
        c1.init_params_uniform()
        c1.forward(write_err=True)
        ch = np.zeros((1, 8, 9, 10, 3))
        a1, _, _ = c1.sweep(1, 7, chain_out=ch)
        assert np.array_equal(ch[0, 1:], chain[k, 1:]), k
        assert a1[0] == acc[k]
        assert np.array_equal(c1.get_residual()[0], res[k])
    # a second call continues every chain where it stopped
    acc2, its2, _ = ctx.sweep(8, 3)
    assert (its2 == 11).all() and (acc2 >= acc).all()


def test_profile_cache_of_the_pipelined_sweep_changes_nothing(nat, monkeypatch):
    """The pipelined kernel keeps the unit line profile every site ended its last visit with and
    reads it back as the next sweep's OLD profile instead of recomputing it from the same (c, w)
    (lib/run.py:402).  With the cache and without it (D3D_NO_LUCACHE=1) the chains are the same
    to the bit -- over several calls, after set_params() between calls (cache invalid: the
    first sweep recomputes), after a coloured sweep in between, and for chains handed over
    between CTAs (more chains than SMs)."""
    g = load_golden('ref_run_A')
    data, fsf, lsf = g['data'], g['fsf'], g['lsf']
    var = np.array([0.01])
    n = 160

    def run(no_cache):
        if no_cache:
            monkeypatch.setenv('D3D_NO_LUCACHE', '1')
        else:
            monkeypatch.delenv('D3D_NO_LUCACHE', raising=False)
        ctx, _, _ = make_ctx(nat, data, var, fsf, lsf, chains=n, seed=11)
        ctx.init_params_uniform()
        ctx.forward(write_err=True)
        out = []
        chain = np.zeros((n, 13, 9, 10, 3))
        ctx.sweep(1, 4, chain_out=chain)                     # first call: sweep 1 fills the cache
        ctx.sweep(5, 3, chain_out=chain)                     # second call: cache valid from its first sweep
        out.append(chain[:, 1:8].copy())
        p = ctx.get_params()
        p[..., 1] += 0.25                                    # parameters moved from outside
        ctx.set_params(p)
        ctx.forward(write_err=True)
        ctx.sweep(8, 2, chain_out=chain)
        ctx.sweep(10, 1, mode=nat.COLOURED, chain_out=chain)  # another kernel moves them
        ctx.sweep(11, 2, chain_out=chain)
        out.append(chain[:, 8:13].copy())
        out.append(ctx.get_residual())
        kern = ctx.last_kernel()
        ctx.close()
        return out, kern

    a, kern = run(False)
    b, _ = run(True)
    assert kern.startswith('sweep_seq_pipe_kernel')
    for x, y in zip(a, b):
        assert np.array_equal(x, y)


@pytest.mark.parametrize('shape,fs,sets', [((40, 40, 40), 13, 3), ((30, 17, 23), 7, 2), ((12, 5, 50), 3, 1),
                                           ((64, 9, 9), 17, 2), ((16, 33, 16), 11, 1)])
def test_stencil_tile_by_tma_and_by_cp_async_agree(nat, monkeypatch, shape, fs, sets):
    """The register-tiled spatial pass stages its halo tile either with one TMA tensor copy (the
    borders of the field come from the TMA unit's zero fill of out-of-range coordinates) or with
    the cp.async loop (explicit zero stores): same cubes and same residuals to the bit, same chi^2, for fields smaller and larger than a tile, several parameter sets, D not a multiple of
    the 16-channel chunk, and against the oracle's forward model at 1e-12."""
    port, _, _ = _oracle()
    D, H, W = shape
    rs = np.random.RandomState(D + fs)
    fsf = port.moffat_fsf_image((fs, fs), 0.2, fwhm_arcsec=0.8, beta=2.5)
    lsf = port.gaussian_lsf_vector(0.0005 if D > 16 else 0.00025, 1.25e-4, D)
    data = synthetic(D, H, W, 5)
    var = 0.01 * (1 + rs.rand(D, H, W))
    params = np.stack([np.dstack([rs.rand(H, W) * 9, rs.rand(H, W) * (D - 1), 0.3 + rs.rand(H, W) * 4])
                       for _ in range(sets)])
    out = []
    for off in (None, '1'):
        if off:
            monkeypatch.setenv('D3D_STENCIL_NO_TMA', off)
        else:
            monkeypatch.delenv('D3D_STENCIL_NO_TMA', raising=False)
        ctx, _, _ = make_ctx(nat, data, var, fsf, lsf, chains=sets)
        ctx.set_params(params)
        sim, chi = ctx.forward(want_sim=True, write_err=True, want_chi2=True)
        out.append((sim.copy(), ctx.get_residual().copy(), np.array(chi)))
        ctx.close()
    assert np.array_equal(out[0][0], out[1][0])
    assert np.array_equal(out[0][1], out[1][1])
    np.testing.assert_allclose(out[0][2], out[1][2], rtol=1e-12)    # (block sums meet in atomics: order varies)
    err_ref = port.compute_error_in_one_step(data, params[0], fsf, lsf, np.ones((H, W)))
    scale = max(np.abs(out[0][0]).max(), np.abs(data).max())
    np.testing.assert_allclose(out[0][1][0], err_ref, rtol=0, atol=1e-12 * scale)


FUZZ = [  # D, H, W, fh, fw, variance, masked, chains, lsf sigma (0: none), dtype
    (16, 9, 10, 7, 7, 'cube', False, 1, 0.9, 'f64'),
    (12, 5, 37, 5, 5, 'scalar', True, 3, 0.7, 'f64'),        # long rows, holes: runs of every length
    (40, 14, 9, 13, 13, 'cube', True, 2, 1.1, 'f64'),         # field narrower than the stamp
    (30, 20, 33, 11, 11, 'scalar', False, 150, 0.9, 'f64'),   # more chains than SMs: hand-over
    (32, 8, 64, 9, 13, 'cube', True, 5, 0.0, 'f64'),          # non-square stamp, no LSF, W > ring
    (20, 3, 70, 3, 3, 'cube', False, 2, 0.8, 'f64'),          # tiny stamp
    (64, 10, 12, 7, 7, 'cube', False, 2, 1.0, 'f64'),         # D = 64: full wrap of the LSF
    (24, 12, 41, 13, 13, 'cube', True, 4, 0.9, 'f32'),
    (10, 1, 50, 7, 7, 'scalar', False, 2, 0.6, 'f64'),        # one row
    (18, 33, 2, 5, 5, 'cube', False, 2, 0.9, 'f64'),          # two columns: runs of two
]


@pytest.mark.parametrize('case', range(len(FUZZ)))
def test_pipelined_and_sliding_window_sweeps_agree(nat, monkeypatch, case):
    """The two sequential sweep kernels are independent implementations of lib/run.py:367-519 (one
    keeps the window sums fresh, the other corrects stale sums through the cross tables; their
    producers evaluate exp differently): on shapes chosen to stress the pipelined kernel's
    protocol -- runs shorter than a producer batch, fewer sites than ring stages, masks, rows
    longer than the ring, non-square stamps, several calls -- they must take the SAME decisions
    and end with the same chains and residuals to rounding.  Where the library does not run the
    pipelined kernel for a shape, the case still checks the kernel that runs."""
    D, H, W, fh, fw, vk, masked, n, sig, dt = FUZZ[case]
    rs = np.random.RandomState(100 + case)
    data = synthetic(D, H, W, seed=case)
    yy, xx = np.mgrid[0:fh, 0:fw]
    fsf = np.exp(-((yy - fh // 2) ** 2 + (xx - fw // 2) ** 2) / (2 * 1.3 ** 2)) * (1 + 0.1 * rs.rand(fh, fw))
    fsf /= fsf.sum()
    mid = (D - 1) // 2 - (D % 2 - 1)                     # centre of an LSF vector (lib/spread_functions.py:212-228)
    lsf = np.zeros(D)
    if sig > 0:
        lsf = np.exp(-(np.arange(D) - mid) ** 2 / (2 * sig ** 2))
        lsf[np.abs(np.arange(D) - mid) > 7] = 0.0
        lsf /= lsf.sum()
    else:
        lsf[mid] = 1.0
    var = 0.05 ** 2 * (1 + rs.rand(D, H, W)) if vk == 'cube' else np.array([0.05 ** 2])
    mask = (rs.rand(H, W) > 0.25).astype(np.float64) if masked else None
    dtype = nat.F32 if dt == 'f32' else nat.F64
    out = {}
    for pipe in ('0', '1'):
        monkeypatch.setenv('D3D_PIPE', pipe)
        ctx, _, _ = make_ctx(nat, data, var, fsf, lsf, mask=mask, dtype=dtype, chains=n, seed=3 + case)
        ctx.init_params_uniform()
        ctx.forward(write_err=True)
        chain = np.zeros((n, 7, H, W, 3))
        lik = np.zeros((n, 7, H, W))
        a1, _, _ = ctx.sweep(1, 4, chain_out=chain, lik_out=lik, min_acceptance_rate=0.0)
        a2, i2, _ = ctx.sweep(5, 2, chain_out=chain, lik_out=lik, min_acceptance_rate=0.0)
        out[pipe] = (chain.copy(), lik.copy(), a2.copy(), ctx.get_residual(), ctx.last_kernel())
        assert (i2 == 7).all()
        ctx.close()
    c0, l0, a0, r0, k0 = out['0']
    c1, l1, a1, r1, k1 = out['1']
    if dt == 'f32':
        # single precision: the kernels sum in different orders (1e-5 relative on delta-logL), a
        # decision next to its threshold may flip and the chains then part: same statistics only
        assert np.all(np.abs(a0 - a1) <= 0.05 * np.maximum(a0, a1) + 8), (a0, a1)
        assert np.isfinite(c1).all() and np.isfinite(r1).all()
        return
    assert np.array_equal(a0, a1), (k0, k1)                    # identical accept counts
    # centres and widths: every entry.  Amplitudes: every entry except where the line of that
    # site is narrower than 0.05 channels -- its profile is then below 1e-40 everywhere (often
    # denormal or zero), the site carries no flux, the amplitude is drawn from the prior alone and
    # WHICH cell of the truncated-normal table the draw starts from hangs on the sign of a
    # ~1e-305 number (as = -mu / sigma, lib/rtnorm.py:74-76), i.e. on how exp() rounds below
    # 1e-300.  Nothing else depends on such an amplitude: the residuals below still agree.
    np.testing.assert_allclose(c1[..., 1:], c0[..., 1:], rtol=1e-8, atol=1e-8)
    flux = c0[..., 2] >= 0.05
    np.testing.assert_allclose(c1[..., 0][flux], c0[..., 0][flux], rtol=1e-8, atol=1e-8)
    assert flux.mean() > 0.5
    np.testing.assert_allclose(r1, r0, rtol=0, atol=1e-9 * max(1.0, np.abs(data).max()))
    # delta-logL of a proposal is taken at the site's CURRENT amplitude (lib/run.py:400-426): the
    # exemption carries over to the sweep after a zero-flux visit
    prev = np.concatenate([flux[:, :1], flux[:, :-1]], axis=1)
    ok = flux & prev
    scale = np.abs(l0).max() + 1.0
    np.testing.assert_allclose(l1[ok], l0[ok], rtol=1e-6, atol=1e-9 * scale)


def _cfg2_problem(nat, chains, dtype=None, fsf_size=13, seed=42):
    """BASELINE cfg2 at full size: 40x40x40 cube, Moffat 13x13, MUSE LSF, variance cube."""
    from deconv3d_b200 import synthetic, MUSE
    D = H = W = 40
    inst = synthetic.muse_wfm_instrument('moffat', fsf_size)
    cube0 = MUSE().build_cube(np.zeros((D, H, W)))
    fsf = np.asarray(inst.fsf.as_image(cube0))
    lsf = inst.lsf.as_vector(cube0)
    truth = synthetic.halpha_truth(D, H, W)
    tmp = nat.Context(0, nat.F64)
    tmp.set_problem(np.ones((D, H, W)), np.ones(1), fsf, lsf, np.zeros(3), [100., D - 1, D],
                    [0, .1, .1], 1.0)
    data = tmp.simulate(truth[None])[0] + synthetic.noise((D, H, W), 0.05, 1234)
    tmp.close()
    var = np.full(data.shape, 0.05 ** 2)
    ctx, pmin, pmax = make_ctx(nat, data, var, fsf, lsf, chains=chains, dtype=dtype, seed=seed)
    return ctx, data, var, fsf, lsf, truth


def test_full_size_cfg2_invariants(nat):
    """Size-independent properties at BASELINE's full cfg2 size (the oracle needs ~1 ms per
    proposal there): (1) the incrementally updated residual equals data - forward(parameters)
    after many sweeps (the drift the reference squashes every 1000 iterations,
    lib/run.py:521-534, stays at rounding level); (2) chi^2 falls to ~N/2 from a poor start."""
    ctx, data, var, fsf, lsf, truth = _cfg2_problem(nat, chains=3)
    rs = np.random.RandomState(0)
    start = truth * np.array([0.3, 1.0, 1.5]) + np.array([0.0, 1.5, 0.0]) * rs.randn(40, 40, 1)
    ctx.set_params(np.broadcast_to(start, (3, 40, 40, 3)).copy())
    _, chi0 = ctx.forward(write_err=True, want_chi2=True)
    n_it = 400
    chain = np.zeros((3, n_it // 20 + 1, 40, 40, 3))
    acc, its, ms = ctx.sweep(1, n_it, keep_one_in=20, refresh_every=0, chain_out=chain)
    assert (its == n_it + 1).all()
    res_inc = ctx.get_residual()
    _, chi1 = ctx.forward(write_err=True, want_chi2=True)
    res_fresh = ctx.get_residual()
    assert np.abs(res_inc - res_fresh).max() < 1e-9 * np.abs(data).max()
    n_vox = data.size
    assert (chi0 > 5 * n_vox).all()
    assert (np.abs(2 * chi1 / n_vox - 1.0) < 0.1).all(), chi1 / n_vox
    # (individual spaxel parameters are ill-constrained by construction -- the PSF mixes
    # neighbours -- so no per-spaxel recovery is asserted; chi^2 of the convolved model is)
    assert np.isfinite(chain).all() and (acc > 0).all()


def test_full_size_cfg2_float32_storage(nat):
    ctx, data, var, fsf, lsf, truth = _cfg2_problem(nat, chains=2, dtype=nat.F32)
    start = truth * np.array([0.3, 1.0, 1.5]) + np.array([0.0, 1.0, 0.0])
    ctx.set_params(np.broadcast_to(start, (2, 40, 40, 3)).copy())
    ctx.forward(write_err=True)
    ctx.sweep(1, 200, refresh_every=100)
    res_inc = ctx.get_residual()
    _, chi = ctx.forward(write_err=True, want_chi2=True)
    assert np.abs(res_inc - ctx.get_residual()).max() < 2e-4 * np.abs(data).max()
    assert (np.abs(2 * chi / data.size - 1.0) < 0.1).all()


@pytest.mark.parametrize('mode', ['seq', 'colour'])
def test_large_fsf_generic_kernels_vs_oracle(nat, mode):
    """FSF larger than the row-mapped kernels take (23x23 on a 14x15 field: every window is
    clipped): the generic kernels against the oracle, decisions included."""
    port, _, _ = _oracle()
    rs = np.random.RandomState(3)
    D, H, W = 10, 14, 15
    data = synthetic(D, H, W, 9)
    fsf = port.moffat_fsf_image((23, 23), 0.2, fwhm_arcsec=1.2, beta=2.5)
    lsf = port.gaussian_lsf_vector(0.0002675, 1.25e-4, D)
    var = 0.05 ** 2 * (1 + rs.rand(D, H, W))
    init = np.dstack([rs.rand(H, W) * 4, 2 + rs.rand(H, W) * 6, 0.7 + rs.rand(H, W) * 2])
    _compare_chain(nat, data, fsf, lsf, var, None, init, 6, 1, seed=17, mode=mode)


@pytest.mark.parametrize('cluster', ['2', '8'])
def test_cluster_split_colour_kernel_vs_oracle(nat, monkeypatch, cluster):
    """Big windows are worked by a thread-block cluster per site (DSMEM reduction): forced here on
    a small problem so that the oracle can follow; decisions included."""
    monkeypatch.setenv('D3D_CLUSTER', cluster)
    port, _, _ = _oracle()
    rs = np.random.RandomState(3)
    D, H, W = 12, 17, 19
    data = synthetic(D, H, W, 9)
    fsf = port.moffat_fsf_image((23, 23), 0.2, fwhm_arcsec=1.2, beta=2.5)
    lsf = port.gaussian_lsf_vector(0.0002675, 1.25e-4, D)
    var = 0.05 ** 2 * (1 + rs.rand(D, H, W))
    init = np.dstack([rs.rand(H, W) * 4, 2 + rs.rand(H, W) * 6, 0.7 + rs.rand(H, W) * 2])
    _compare_chain(nat, data, fsf, lsf, var, None, init, 5, 1, seed=17, mode='colour')
    _compare_chain(nat, data, fsf, lsf, np.array([0.05 ** 2]), None, init, 4, 1, seed=18, mode='colour')


def test_mat_fixture_chain_vs_oracle(nat):
    """The reference's own .mat fixture (tests/input/data14forAntoine.mat: 21x30x24 cube with
    its per-voxel variance and 15x15 FSF, delta LSF as in tests/read_mat.py:93-94), started
    from its ground-truth parameters: odd depth (padded channels), 15-row FSF (row-mapped
    kernel with 21 register rows), non-square field."""
    port, _, _ = _oracle()
    g = load_golden('mat_kat')
    data, var, fsf, params = g['data'], g['variance'], g['fsf'], g['params']
    delta = port.gaussian_lsf_vector(0.0, 1.25e-4, data.shape[0])
    mask = (port.above_percentile(data, 60) == 1).astype(float)        # tests/read_mat.py mask
    _compare_chain(nat, data, fsf, delta, var, mask, params, 4, 1, seed=31)


def test_tiny_field_smaller_than_fsf(nat):
    """Field smaller than the FSF (H, W < fh, fw): every window is clipped on both sides."""
    port, _, _ = _oracle()
    rs = np.random.RandomState(4)
    D, H, W = 12, 5, 6
    data = synthetic(D, H, W, 2)
    fsf = port.gaussian_fsf_image(1.0, 0.2)                            # 13x13
    lsf = port.gaussian_lsf_vector(0.0002675, 1.25e-4, D)
    init = np.dstack([rs.rand(H, W) * 4, 2 + rs.rand(H, W) * 6, 0.7 + rs.rand(H, W) * 2])
    var = np.array([0.05 ** 2])      # (the default variance guess needs H > 6, lib/run.py:187)
    _compare_chain(nat, data, fsf, lsf, var, None, init, 6, 1, seed=8)
    _compare_chain(nat, data, fsf, lsf, var, None, init, 5, 1, seed=9, mode='colour')


@pytest.mark.parametrize('cluster', [None, '4', '7'])
def test_colour_phases_with_and_without_dependent_launch_agree(nat, monkeypatch, cluster):
    """The phase kernels of one coloured sweep are launched as programmatic dependents
    (`griddepcontrol.launch_dependents` / `.wait` in sweep_colour_cluster_kernel: phase p+1 sets up
    while phase p drains and waits in front of its first residual access).  A phase that read the
    residual too early would take other decisions: chains, likelihood rows, residual and counters
    equal the plainly serialised launches (D3D_NO_PDL=1) to the bit over 529 phases per sweep, with
    the cluster size the library picks and with forced ones (an odd size included)."""
    port, _, _ = _oracle()
    rs = np.random.RandomState(5)
    D, H, W = 64, 40, 44                   # 23 x 23 x 64 window: the library takes the cluster kernel by itself
    data = synthetic(D, H, W, 13)
    fsf = port.moffat_fsf_image((23, 23), 0.2, fwhm_arcsec=1.2, beta=2.5)
    lsf = port.gaussian_lsf_vector(0.0002675, 1.25e-4, D)
    var = 0.05 ** 2 * (1 + rs.rand(D, H, W))
    init = np.dstack([rs.rand(H, W) * 4, 20 + rs.rand(H, W) * 24, 0.7 + rs.rand(H, W) * 2])
    if cluster:
        monkeypatch.setenv('D3D_CLUSTER', cluster)
    else:
        monkeypatch.delenv('D3D_CLUSTER', raising=False)
    out = []
    for no_pdl in (None, '1'):
        if no_pdl:
            monkeypatch.setenv('D3D_NO_PDL', no_pdl)
        else:
            monkeypatch.delenv('D3D_NO_PDL', raising=False)
        ctx, _, _ = make_ctx(nat, data, var, fsf, lsf, seed=23)
        ctx.set_params(init[None])
        ctx.forward(write_err=True)
        chain = np.zeros((1, 4, H, W, 3))
        lik = np.zeros((1, 4, H, W))
        acc, its, _ = ctx.sweep(1, 3, mode=nat.COLOURED, keep_one_in=1, min_acceptance_rate=0.0,
                                chain_out=chain, lik_out=lik)
        assert ctx.last_kernel() == 'sweep_colour_cluster_kernel'
        out.append((chain, lik, ctx.get_residual().copy(), acc.copy(), its.copy()))
        ctx.close()
    for a, b in zip(out[0], out[1]):
        assert np.array_equal(a, b)
    assert out[0][3][0] > 0                                       # something was accepted


@pytest.mark.parametrize('fs,chains', [(13, 1), (7, 5), (23, 2)])
def test_row_mapped_colour_phases_with_and_without_dependent_launch_agree(nat, monkeypatch, fs, chains):
    """Same check for the launch-per-class coloured mode on the row-mapped (7x7, 13x13) and generic
    (23x23) window kernels: the phases of a sweep are programmatic dependents of each other there
    too (D3D_COLOUR_PROLOGUE); identical chains, likelihood rows, residuals and counters."""
    port, _, _ = _oracle()
    rs = np.random.RandomState(fs)
    D, H, W = 24, 30, 33
    data = synthetic(D, H, W, 21)
    fsf = port.moffat_fsf_image((fs, fs), 0.2, fwhm_arcsec=0.8, beta=2.5)
    lsf = port.gaussian_lsf_vector(0.0002675, 1.25e-4, D)
    var = 0.05 ** 2 * (1 + rs.rand(D, H, W))
    init = np.dstack([rs.rand(H, W) * 4, 6 + rs.rand(H, W) * 12, 0.7 + rs.rand(H, W) * 2])
    monkeypatch.setenv('D3D_COLOUR_BY_CHAIN', '0')                 # one launch per colour class
    monkeypatch.setenv('D3D_CLUSTER', '0')
    out = []
    for no_pdl in (None, '1'):
        if no_pdl:
            monkeypatch.setenv('D3D_NO_PDL', no_pdl)
        else:
            monkeypatch.delenv('D3D_NO_PDL', raising=False)
        ctx, _, _ = make_ctx(nat, data, var, fsf, lsf, chains=chains, seed=29)
        ctx.set_params(np.broadcast_to(init, (chains, H, W, 3)).copy())
        ctx.forward(write_err=True)
        chain = np.zeros((chains, 5, H, W, 3))
        lik = np.zeros((chains, 5, H, W))
        acc, its, _ = ctx.sweep(1, 4, mode=nat.COLOURED, keep_one_in=1, min_acceptance_rate=0.0,
                                chain_out=chain, lik_out=lik)
        assert ctx.last_kernel() in ('sweep_colour_kernel', 'sweep_colour_generic_kernel')
        out.append((chain, lik, ctx.get_residual().copy(), acc.copy(), its.copy()))
        ctx.close()
    for a, b in zip(out[0], out[1]):
        assert np.array_equal(a, b)
    assert (out[0][3] > 0).all()
