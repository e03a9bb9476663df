"""
GPU <-> oracle parity ON THE CONFIGURATIONS THAT ARE BENCHED (run with ``-m gpu`` on a B200).

tests/test_gpu_parity.py pins the kernels on small cases; here the oracle
(oracle.reference_port.run_chain, the literal restatement of lib/run.py:367-519, fed by the
same Philox stream) follows the full-size BASELINE.json configurations for a few sweeps:

  (i)   the exact problem of ``bench.build_workload('cfg2x256')`` -- 40^3, Moffat 13x13, variance
        cube, row-major order, 256 chains in one balanced launch -- chains 0, 147, 148, 255
        (first / last chain of the first wave of SMs, first / last chain that is handed over);
  (ii)  one cfg5 galaxy (32^3, FSF 11x11, its own data + variance) inside a multi-galaxy context;
  (iii) cfg3: the cfg2 cube in COLOURED mode at full size, oracle run in colour-class order;
  (iv)  cfg2 with the 21x21 stamp (SURVEY.md 8d "secondary");
  (v)   a D = 64 cube (P = 64: full spectral wrap, cfg4's depth).

Bar (BASELINE.json north_star): identical accept/reject decisions proposal by proposal, chains
to 1e-9, per-proposal delta-logL to 1e-6 relative, final residual to 1e-9 max|data|.
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def nat():
    import torch
    if not torch.cuda.is_available():
        pytest.fail('these tests need a CUDA device (there is no CPU fallback)')
    from deconv3d_b200 import _native
    return _native


def _tables():
    from deconv3d_b200 import rtnorm_tables
    x, yu, nc = rtnorm_tables.tables()
    return x, yu, nc.astype(np.int64)


def _check_chain(ref, trace, chain, lik, acc, its, residual, data, keep=1, rtol=1e-9):
    """One device chain (rows [n_saved,H,W,3]) against one oracle run of the same stream."""
    m = ref['mask'] == 1
    H, W = m.shape
    assert its == ref['iterations']
    n_acc_ref = sum(1 for v in trace.values() if v[3])
    assert acc - m.sum() == n_acc_ref, (acc, n_acc_ref)
    max_it = ref['iterations']
    if keep == 1:
        for it in range(1, max_it):
            moved = (chain[it, :, :, 1] != chain[it - 1, :, :, 1]) | \
                    (chain[it, :, :, 2] != chain[it - 1, :, :, 2])
            ref_acc = np.zeros((H, W), bool)
            for (y, x) in zip(*np.nonzero(m)):
                ref_acc[y, x] = trace[(it, y, x)][3]
            assert np.array_equal(moved & m, ref_acc), 'decisions differ at iteration %d' % it
    scale = np.abs(ref['chain'][:, m]).max()
    ok = np.ones(ref['chain'].shape, bool)
    for (it, y, x), v in trace.items():            # degenerate draws, see test_gpu_parity._compare_chain
        mu, ro = v[5], v[6]
        if it % keep == 0 and abs(mu) <= 1e-9 * np.sqrt(ro):
            ok[it // keep, y, x, 0] = False
    assert (~ok).sum() <= 0.005 * ok.size
    sel = ok & m[None, :, :, None]
    # chains to 1e-9 relative.  Right after a random start the window sums are ~1e9 and cancel
    # down to the posterior mean, so the amplitude carries their rounding: on top of rtol the Gibbs
    # draw may deviate by 1e-7 of its own posterior standard deviation sqrt(ro) (lib/run.py:492-496)
    atol = np.full(ref['chain'].shape, 1e-12 * scale)
    for (it, y, x), v in trace.items():
        if it % keep == 0:
            atol[it // keep, y, x, 0] += 1e-7 * np.sqrt(abs(v[6]))
    dev = np.abs(chain - ref['chain'])
    bad = sel & (dev > rtol * np.abs(ref['chain']) + atol)
    assert not bad.any(), (np.argwhere(bad)[:5], dev[bad][:5], ref['chain'][bad][:5])
    # delta-logL to 1e-6 relative; the REFERENCE forms it as the difference of two O(N) sums
    # (lib/run.py:423-426), so its own value carries ~eps * ar_old of cancellation noise (ar_old is
    # ~5e8 right after a random start): that noise is allowed on top, element by element
    ar_old = np.zeros(lik.shape)
    for (it, y, x), v in trace.items():
        if it % keep == 0:
            ar_old[it // keep, y, x] = abs(v[1])
    d_dev, d_ref = lik[1:, m], ref['likelihoods'][1:, m]
    tol = max(rtol, 1e-6) * np.abs(d_ref) + 1e-9 + 1e-14 * ar_old[1:, m]
    bad = np.abs(d_dev - d_ref) > tol
    assert not bad.any(), (np.abs(d_dev - d_ref)[bad][:5], d_ref[bad][:5], ar_old[1:, m][bad][:5])
    if residual is not None:
        np.testing.assert_allclose(residual, ref['err'], rtol=0, atol=1e-9 * np.abs(data).max())


def _oracle_chain(data, var, fsf, lsf, seed, chain_id, init, max_it, order=None, prior=None):
    from oracle import reference_port as port, streams
    trace = {}
    ref = port.run_chain(data, fsf, lsf, streams.PhiloxStream(seed, chain_id),
                         variance_cube=var, initial_parameters=init,
                         gibbs_apriori_variance=prior, max_iterations=max_it, trace=trace,
                         rtnorm_tables=_tables(), site_order=order, min_acceptance_rate=0.0,
                         refresh_every=0)
    return ref, trace


def _bench_arrays(name, chains):
    import bench
    wl = bench.build_workload(name, chains)
    return wl, bench.realise(wl, 0)


def _device_run(nat, arrays, chains_per_cube, max_it, mode, seed=42, first_chain=0, n_cubes=1):
    ctx = nat.Context(0, nat.F64)
    ctx.set_rtnorm_tables(*_tables())
    ctx.set_rng(seed, first_chain)
    ctx.set_problem(arrays['data'], arrays['var'], arrays['fsf'], arrays['lsf'], arrays['pmin'],
                    arrays['pmax'], [0, 0.1, 0.1], arrays['prior'], chains_per_cube=chains_per_cube)
    ctx.init_params_uniform()
    n = n_cubes * chains_per_cube
    D, H, W = arrays['data'].shape[-3:]
    chain = np.zeros((n, max_it, H, W, 3))
    lik = np.zeros((n, max_it, H, W))
    chain[:, 0] = ctx.get_params()
    ctx.forward(write_err=True)
    acc, its, _ = ctx.sweep(1, max_it - 1, mode=mode, refresh_every=1000, min_acceptance_rate=0.0,
                            chain_out=chain, lik_out=lik)
    res = ctx.get_residual()
    ctx.close()
    return chain, lik, acc, its, res


@pytest.mark.parametrize('pipe', ['1', '0'])
def test_cfg2x256_benched_launch_vs_oracle(nat, monkeypatch, pipe):
    """(i) The headline workload itself: 256 chains, balanced (wrap-around) launch, row-major
    order, <double, variance cube, 13 rows>: with the kernel the library picks at this chain
    count (D3D_PIPE=1: the pipelined kernel) and with the sliding-window kernel (D3D_PIPE=0)
    on the same balanced launch (work items, hand-over between CTAs)."""
    monkeypatch.setenv('D3D_PIPE', pipe)
    wl, arrays = _bench_arrays('cfg2x256', 256)
    max_it = 4
    chain, lik, acc, its, res = _device_run(nat, arrays, 256, max_it, nat.SEQ_EXACT)
    for k in (0, 147, 148, 255):
        ref, trace = _oracle_chain(arrays['data'][0], arrays['var'][0], arrays['fsf'], arrays['lsf'],
                                   42, k, chain[k, 0], max_it, prior=float(arrays['prior'][0]))
        _check_chain(ref, trace, chain[k], lik[k], int(acc[k]), int(its[k]), res[k], arrays['data'][0])


def test_overlapped_row_copies_change_nothing(nat, monkeypatch):
    """keep_one_in=1 on the benched workload sends ~100 MB of chain and likelihood rows to host
    memory per call: d3d_sweep then launches the sweeps in chunks and copies the rows of a chunk
    while the next one computes.  Same rows, same counters as the single launch + single copy
    (D3D_NO_COPY_OVERLAP=1), to the bit."""
    wl, arrays = _bench_arrays('cfg2x256', 256)
    out = []
    for off in ('1', None):
        if off:
            monkeypatch.setenv('D3D_NO_COPY_OVERLAP', off)
        else:
            monkeypatch.delenv('D3D_NO_COPY_OVERLAP', raising=False)
        chain, lik, acc, its, res = _device_run(nat, arrays, 256, 10, nat.SEQ_EXACT)
        out.append((chain, lik, acc, its, res))
    for a, b in zip(out[0], out[1]):
        assert np.array_equal(a, b)
    assert (out[1][3] == 10).all() and np.abs(out[1][0][:, 9]).max() > 0


def test_soak_benched_workload_across_refreshes(nat):
    """1 500 sweeps of the benched workload (256 chains, balanced launch, pipelined kernel) in
    three calls that cross the residual refresh of lib/run.py:521-534 at iteration 1000: every
    chain reaches the last iteration, no bounded wait gives up, chi^2 per voxel settles at 1, and
    500 sweeps after the refresh the incrementally updated residual still equals
    data - forward(parameters) to rounding (7e8 site updates; the barrier phases of a CTA have
    wrapped 1e5 times)."""
    wl, arrays = _bench_arrays('cfg2x256', 256)
    ctx = nat.Context(0, nat.F64)
    ctx.set_rtnorm_tables(*_tables())
    ctx.set_rng(7, 0)
    ctx.set_problem(arrays['data'], arrays['var'], arrays['fsf'], arrays['lsf'], arrays['pmin'],
                    arrays['pmax'], [0, 0.1, 0.1], arrays['prior'], chains_per_cube=256)
    ctx.init_params_uniform()
    ctx.forward(write_err=True)
    it = 1
    for n in (700, 650, 150):
        acc, its, _ = ctx.sweep(it, n, refresh_every=1000, min_acceptance_rate=0.0)
        it += n
        assert (its == it).all()
    res = ctx.get_residual()
    sim = ctx.simulate(ctx.get_params())
    data = arrays['data'][0]
    assert np.abs((data[None] - sim) - res).max() < 1e-9 * np.abs(data).max()
    chi2 = (res ** 2 / arrays['var'][0][None]).sum(axis=(1, 2, 3)) / data.size
    assert 0.95 < chi2.min() and chi2.max() < 1.1
    assert ctx.last_kernel().startswith('sweep_seq_pipe_kernel')
    ctx.close()


def test_cfg5_galaxy_vs_oracle(nat):
    """(ii) Survey batch: galaxy 2 of a 3-galaxy context (32^3, FSF 11x11, own data/variance)."""
    wl, arrays = _bench_arrays('cfg5', 3)
    max_it = 4
    chain, lik, acc, its, res = _device_run(nat, arrays, 1, max_it, nat.SEQ_EXACT, n_cubes=3)
    g = 2
    ref, trace = _oracle_chain(arrays['data'][g], arrays['var'][g], arrays['fsf'], arrays['lsf'],
                               42, g, chain[g, 0], max_it, prior=float(arrays['prior'][g]))
    # boundaries of the context = the reference's own for that galaxy
    np.testing.assert_allclose(ref['max_boundaries'], arrays['pmax'][g])
    _check_chain(ref, trace, chain[g], lik[g], int(acc[g]), int(its[g]), res[g], arrays['data'][g])


@pytest.mark.parametrize('by_chain', ['0', '1'])
def test_cfg3_coloured_full_size_vs_oracle(nat, monkeypatch, by_chain):
    """(iii) cfg3: colour-class order at full size, both schedules (launch per class / chain per
    CTA on the colour-ordered site list)."""
    from oracle import reference_port as port
    monkeypatch.setenv('D3D_COLOUR_BY_CHAIN', by_chain)
    wl, arrays = _bench_arrays('cfg2x256', 2)
    max_it = 3
    chain, lik, acc, its, res = _device_run(nat, arrays, 2, max_it, nat.COLOURED)
    fh, fw = arrays['fsf'].shape
    order = port.colour_class_order(np.ones((40, 40)), fh, fw)
    k = 1
    ref, trace = _oracle_chain(arrays['data'][0], arrays['var'][0], arrays['fsf'], arrays['lsf'],
                               42, k, chain[k, 0], max_it, order=order, prior=float(arrays['prior'][0]))
    _check_chain(ref, trace, chain[k], lik[k], int(acc[k]), int(its[k]), res[k], arrays['data'][0])


def test_cfg2_fsf21_vs_oracle(nat):
    """(iv) cfg2 with the Moffat stamp truncated to 21x21."""
    from deconv3d_b200 import synthetic
    import bench
    wl = bench.build_workload('cfg2', 1)
    wl['inst'] = synthetic.muse_wfm_instrument('moffat', 21)
    arrays = bench.realise(wl, 0)
    assert arrays['fsf'].shape == (21, 21)
    max_it = 3
    chain, lik, acc, its, res = _device_run(nat, arrays, 1, max_it, nat.SEQ_EXACT)
    ref, trace = _oracle_chain(arrays['data'][0], arrays['var'][0], arrays['fsf'], arrays['lsf'],
                               42, 0, chain[0, 0], max_it, prior=float(arrays['prior'][0]))
    _check_chain(ref, trace, chain[0], lik[0], int(acc[0]), int(its[0]), res[0], arrays['data'][0])


@pytest.mark.parametrize('fsf_size', [13, 21])
def test_depth64_vs_oracle(nat, fsf_size):
    """(v) D = 64 (P = 64: every channel wraps, cfg4's depth) on a 20x22 field."""
    from deconv3d_b200 import synthetic, MUSE
    from oracle import reference_port as port
    D, H, W = 64, 20, 22
    inst = synthetic.muse_wfm_instrument('moffat', fsf_size)
    cube0 = MUSE().build_cube(np.zeros((D, H, W)))
    fsf = np.asarray(inst.fsf.as_image(cube0), dtype=np.float64)
    lsf = inst.lsf.as_vector(cube0)
    truth = synthetic.halpha_truth(D, H, W)
    mask = np.ones((H, W))
    data = -port.compute_error_in_one_step(np.zeros((D, H, W)), truth, fsf, lsf, mask) \
        + synthetic.noise((D, H, W), 0.05, 77)
    rs = np.random.RandomState(3)
    var = 0.05 ** 2 * (1 + rs.rand(D, H, W))
    pmax = np.array([[data.max() / fsf.max(), D - 1, D]])
    arrays = dict(data=data[None], var=var[None], fsf=fsf, lsf=lsf, pmin=np.zeros((1, 3)), pmax=pmax,
                  prior=pmax[:, 0] ** 2)
    max_it = 3
    chain, lik, acc, its, res = _device_run(nat, arrays, 2, max_it, nat.SEQ_EXACT, seed=9)
    k = 1
    ref, trace = _oracle_chain(data, var, fsf, lsf, 9, k, chain[k, 0], max_it, prior=float(pmax[0, 0] ** 2))
    _check_chain(ref, trace, chain[k], lik[k], int(acc[k]), int(its[k]), res[k], data)


@pytest.mark.parametrize('pipe', ['0', '1'])
def test_balanced_schedule_with_stopped_chains(nat, monkeypatch, pipe):
    """(both sweep kernels, one at a time -- D3D_PIPE=0: sliding-window kernel everywhere, 1: the
    pipelined kernel wherever it can run, also on the balanced launch; bit equality holds per kernel)
    More chains than SMs, min_acceptance_rate > 0 and chains that stop early, followed by a
    second d3d_sweep call: a CTA whose FIRST work item is a stopped chain must still serve its
    remaining chains correctly (the truncated-normal tables are loaded once per CTA, not inside
    the first item).  Every chain equals its own single-chain run."""
    from conftest import load_golden
    monkeypatch.setenv('D3D_PIPE', pipe)
    g = load_golden('ref_run_A')
    data, fsf, lsf = g['data'], g['fsf'], g['lsf']
    from oracle import reference_port as port
    pmin, pmax = port.single_gaussian_boundaries(data, fsf)
    n = 333

    def mk(chains, first):
        ctx = nat.Context(0, nat.F64)
        ctx.set_rtnorm_tables(*_tables())
        ctx.set_rng(5, first)
        # a tiny jump on c, w and a huge one would both do; a high threshold stops chains for sure
        ctx.set_problem(data, np.array([0.01]), fsf, lsf, pmin, pmax, [0, 0.1, 0.1], float(pmax[0]) ** 2,
                        chains_per_cube=chains)
        ctx.init_params_uniform()
        ctx.forward(write_err=True)
        return ctx

    H, W = data.shape[1:]
    ns = H * W
    # dry run without a threshold, one sweep per call: accepted counts after every sweep give the
    # acceptance rate every chain will show at every iteration (lib/run.py:356-359)
    dry = mk(n, 0)
    A = [np.full(n, ns, dtype=np.int64)]                 # accepted_count starts at n_sites (:341)
    for it in range(1, 13):
        a, _, _ = dry.sweep(it, 1, min_acceptance_rate=0.0)
        A.append(np.array(a, dtype=np.int64))
    dry.close()
    # the test before sweep `it` uses the rate formed at it-1: A[it-2] / (ns * (it-1))
    low = np.min([A[j - 1] / float(ns * j) for j in range(3, 12)], axis=0)
    srt = np.sort(low)
    mid = n // 2
    gaps = srt[mid - 20:mid + 20][1:] - srt[mid - 20:mid + 20][:-1]
    q = int(np.argmax(gaps))
    thr = 0.5 * (srt[mid - 20 + q] + srt[mid - 20 + q + 1])   # inside the widest gap near the median

    ctx = mk(n, 0)
    chain = np.zeros((n, 13, H, W, 3))
    acc0, its0, _ = ctx.sweep(1, 3, chain_out=chain, min_acceptance_rate=0.0)
    acc1, its1, _ = ctx.sweep(4, 4, chain_out=chain, min_acceptance_rate=thr)
    acc2, its2, _ = ctx.sweep(8, 5, chain_out=chain, min_acceptance_rate=thr)
    stopped = its2 < 13
    assert stopped.any() and (~stopped).any(), (its2.min(), its2.max())
    res = ctx.get_residual()
    picks = [0, 1, 147, 148, 149, 332] + list(np.nonzero(stopped)[0][:3]) + list(np.nonzero(~stopped)[0][:3])
    for k in sorted(set(int(p) for p in picks)):
        c1 = mk(1, k)
        ch = np.zeros((1, 13, H, W, 3))
        c1.sweep(1, 3, chain_out=ch, min_acceptance_rate=0.0)
        c1.sweep(4, 4, chain_out=ch, min_acceptance_rate=thr)
        a1, i1, _ = c1.sweep(8, 5, chain_out=ch, min_acceptance_rate=thr)
        assert i1[0] == its2[k] and a1[0] == acc2[k], k
        nrow = int(its2[k])
        assert np.array_equal(ch[0, 1:nrow], chain[k, 1:nrow]), k
        assert np.array_equal(c1.get_residual()[0], res[k]), k
        c1.close()


# ---------------------------------------------------------------------------------------------
# f-4: line models beyond the single Gaussian (lib/line_models.py:4-61): tied multiplets
# ---------------------------------------------------------------------------------------------
MULTIPLETS = [([0.0, 3.2], [1.0, 0.6]),                       # a doublet ([OII]-like)
              ([0.0, -4.5, 6.1], [1.0, 0.11, 0.33])]          # [NII] - Halpha - [NII]-like triplet


@pytest.mark.parametrize('comp', MULTIPLETS)
def test_multiplet_forward_model_vs_oracle(nat, comp):
    """Forward model (lib/run.py:999-1031) of a tied multiplet, 1e-12 relative (fp64)."""
    from oracle import reference_port as port
    from conftest import load_golden
    g = load_golden('ref_run_A')
    data, fsf, lsf = g['data'], g['fsf'], g['lsf']
    D, H, W = data.shape
    rs = np.random.RandomState(4)
    params = np.dstack([rs.rand(H, W) * 5, 2 + rs.rand(H, W) * (D - 4), 0.6 + rs.rand(H, W) * 2])
    pmin, pmax = port.single_gaussian_boundaries(data, fsf)
    ctx = nat.Context(0, nat.F64)
    ctx.set_problem(data, np.array([0.01]), fsf, lsf, pmin, pmax, [0, .1, .1], float(pmax[0]) ** 2)
    ctx.set_line_model(*comp)
    ctx.set_params(params[None])
    sim, _ = ctx.forward(want_sim=True, write_err=True)
    with port.line_model(*comp):
        err_ref = port.compute_error_in_one_step(data, params, fsf, lsf, np.ones((H, W)))
        clean_ref = port.simulate_clean(data.shape, params, np.ones((H, W)))
    scale = np.abs(data - err_ref).max()
    np.testing.assert_allclose(sim[0], data - err_ref, rtol=1e-12, atol=1e-12 * scale)
    np.testing.assert_allclose(ctx.get_residual()[0], err_ref, rtol=0, atol=1e-12 * np.abs(data).max())
    np.testing.assert_allclose(ctx.simulate_clean(params[None])[0], clean_ref, rtol=1e-12, atol=1e-14)
    # one Gaussian again: the model is a property of the context, not of the build
    ctx.set_line_model(None)
    sim1, _ = ctx.forward(want_sim=True, write_err=False)
    err1 = port.compute_error_in_one_step(data, params, fsf, lsf, np.ones((H, W)))
    np.testing.assert_allclose(sim1[0], data - err1, rtol=1e-12, atol=1e-12 * scale)
    ctx.close()


@pytest.mark.parametrize('mode,pipe', [('seq', '0'), ('seq', '1'), ('colour', '1')])
@pytest.mark.parametrize('comp', MULTIPLETS[:1] + MULTIPLETS[1:])
def test_multiplet_chain_vs_oracle(nat, monkeypatch, mode, pipe, comp):
    """Sweeps with a tied multiplet against the oracle running the same model: identical
    decisions, chains to 1e-9 -- sliding-window kernel, pipelined kernel and coloured mode."""
    from oracle import reference_port as port
    from conftest import load_golden
    monkeypatch.setenv('D3D_PIPE', pipe)
    g = load_golden('ref_run_C')                         # mask + variance cube + D = 16 (full wrap)
    data, fsf, lsf = g['data'], g['fsf'], g['lsf']
    var, mask_in, init = g['in_variance'], g['in_mask'], g['in_initial_parameters']
    D, H, W = data.shape
    max_it = 8
    order = None
    m_for_order = port.prepare_mask(data, mask_in.copy())
    if mode == 'colour':
        order = port.colour_class_order(m_for_order, fsf.shape[0], fsf.shape[1])
    with port.line_model(*comp):
        from oracle import streams
        trace = {}
        ref = port.run_chain(data, fsf, lsf, streams.PhiloxStream(3, 0), mask=mask_in.copy(),
                             variance_cube=var, initial_parameters=init, jump_amplitude=0.3,
                             gibbs_apriori_variance=50.0, max_iterations=max_it, trace=trace,
                             rtnorm_tables=_tables(), site_order=order, min_acceptance_rate=0.0,
                             refresh_every=0)
    pmin, pmax = port.single_gaussian_boundaries(data, fsf)
    ctx = nat.Context(0, nat.F64)
    ctx.set_rtnorm_tables(*_tables())
    ctx.set_rng(3, 0)
    ctx.set_problem(data, var, fsf, lsf, pmin, pmax, [0, .3, .3], 50.0, mask=ref['mask'])
    ctx.set_line_model(*comp)
    ctx.set_params(np.asarray(init, float)[None])
    chain = np.zeros((1, max_it, H, W, 3))
    lik = np.zeros((1, max_it, H, W))
    chain[0, 0] = ctx.get_params()[0]
    ctx.forward(write_err=True)
    acc, its, _ = ctx.sweep(1, max_it - 1, mode=nat.SEQ_EXACT if mode == 'seq' else nat.COLOURED,
                            refresh_every=0, min_acceptance_rate=0.0, chain_out=chain, lik_out=lik)
    _check_chain(ref, trace, chain[0], lik[0], int(acc[0]), int(its[0]), ctx.get_residual()[0], data)
    ctx.close()


def test_run_dropin_with_tied_multiplet(nat):
    """``Run(model=TiedGaussiansLineModel(...))`` through the reference's plug-in argument
    (lib/run.py:95-109 `model=`): convolved output cube = forward model of the parameters."""
    from deconv3d_b200 import Run, MUSE, TiedGaussiansLineModel
    from oracle import reference_port as port
    from conftest import load_golden
    g = load_golden('ref_run_A')
    data = g['data']
    inst = MUSE(fsf_fwhm=0.5)
    model = TiedGaussiansLineModel([0.0, 3.2], [1.0, 0.6])
    run = Run(inst.build_cube(data), inst, model=model, max_iterations=6, seed=5)
    with port.line_model([0.0, 3.2], [1.0, 0.6]):
        ref = data - port.compute_error_in_one_step(data, run.parameters, run.fsf, run.lsf, run.mask)
    np.testing.assert_allclose(run.convolved_cube.data, ref, rtol=1e-12, atol=1e-12 * np.abs(ref).max())
    line = run.model.modelize(run, range(0, data.shape[0]), run.parameters[3, 4])
    with port.line_model([0.0, 3.2], [1.0, 0.6]):
        np.testing.assert_allclose(line, port.modelize(data.shape[0], run.parameters[3, 4]), rtol=1e-14)
