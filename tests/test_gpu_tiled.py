"""
GPU tests (``-m gpu``) of the spatially tiled coloured sweep of ONE cube:
several contexts on the same GPU stand for the ranks, the records travel through the same
buffers the NCCL all-gather fills on a multi-GPU box.  Property: the result does not depend on
the tiling, bit for bit (draws are addressed by (seed, chain, sweep, site)).
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


def _problem(D, H, W, fsf_shape, seed):
    port, _, _ = _oracle()
    rs = np.random.RandomState(seed)
    data = synthetic(D, H, W, seed)
    fsf = port.moffat_fsf_image(fsf_shape, 0.2, fwhm_arcsec=0.8, beta=2.5)
    lsf = port.gaussian_lsf_vector(0.0002675, 1.25e-4, D)
    var = 0.05 ** 2 * (1 + rs.rand(D, H, W))
    mask = (rs.rand(H, W) > 0.1).astype(float)
    init = np.dstack([rs.rand(H, W) * 4, 2 + rs.rand(H, W) * (D - 4), 0.7 + rs.rand(H, W) * 2])
    return data, var, fsf, lsf, mask, init


def _single(nat, prob, chains, n_it, dtype=None):
    data, var, fsf, lsf, mask, init = prob
    D, H, W = data.shape
    ctx, _, _ = make_ctx(nat, data, var, fsf, lsf, mask=mask, chains=chains, seed=77, dtype=dtype)
    ctx.set_params(np.broadcast_to(init, (chains, H, W, 3)))
    ctx.forward(write_err=True)
    chain = np.zeros((chains, n_it + 1, H, W, 3))
    lik = np.zeros((chains, n_it + 1, H, W))
    acc, its, _ = ctx.sweep(1, n_it, mode=nat.COLOURED, refresh_every=0, min_acceptance_rate=0.0,
                            chain_out=chain, lik_out=lik)
    return chain, lik, acc, ctx.get_residual()


def _tiled(nat, prob, chains, n_it, n_tiles, dtype=None, refresh_every=0, fused=False):
    from deconv3d_b200 import dist
    data, var, fsf, lsf, mask, init = prob
    D, H, W = data.shape
    ctxs = []
    for _ in range(n_tiles):
        ctx, _, _ = make_ctx(nat, data, var, fsf, lsf, mask=mask, chains=chains, seed=77, dtype=dtype)
        ctx.set_params(np.broadcast_to(init, (chains, H, W, 3)))
        ctx.forward(write_err=True)
        ctxs.append(ctx)
    sw = dist.TiledSweeper(ctxs, (H, W), fsf.shape, fused=fused)
    chain = np.zeros((chains, n_it + 1, H, W, 3))
    lik = np.zeros((chains, n_it + 1, H, W))
    acc, its = sw.sweep(1, n_it, refresh_every=refresh_every, chain_out=chain, lik_out=lik)
    regions = [(c_, t) for c_, t in zip(ctxs, sw.tiles)]
    partial = [c_.get_residual() for c_ in ctxs]
    sw.finish()
    return chain, lik, acc, [c_.get_residual() for c_ in ctxs], partial, sw, regions


@pytest.mark.parametrize('n_tiles', [2, 4, 6])
def test_tiling_does_not_change_the_chain(nat, n_tiles):
    """13x13 FSF (row-mapped kernels), 2 chains, masked sites, variance cube."""
    prob = _problem(16, 30, 34, (13, 13), 3)
    chain1, lik1, acc1, res1 = _single(nat, prob, 2, 4)
    chain, lik, acc, res, partial, sw, regions = _tiled(nat, prob, 2, 4, n_tiles)
    m = prob[4] == 1
    assert np.array_equal(chain[:, 1:][:, :, m], chain1[:, 1:][:, :, m])
    assert np.array_equal(lik[:, 1:][:, :, m], lik1[:, 1:][:, :, m])
    assert np.array_equal(acc, acc1)
    # inside its region (tile grown by the FSF half-size) every context kept the residual of the
    # single-context run; after finish() the whole residual is rebuilt from the parameters
    fhh = 6
    for (ctx, (y0, y1, x0, x1)), part in zip(regions, partial):
        ys = slice(max(y0 - fhh, 0), y1 + fhh)
        xs = slice(max(x0 - fhh, 0), x1 + fhh)
        assert np.array_equal(part[:, :, ys, xs], res1[:, :, ys, xs])
    for r in res:
        np.testing.assert_allclose(r, res1, rtol=0, atol=1e-9)


def test_tiling_large_fsf_generic_kernel(nat):
    """41x41 FSF like cfg4 (generic kernels; windows span several tiles), D = 64 (P = 64)."""
    prob = _problem(64, 44, 50, (41, 41), 5)
    chain1, lik1, acc1, res1 = _single(nat, prob, 1, 2)
    chain, lik, acc, res, partial, sw, regions = _tiled(nat, prob, 1, 2, 4)
    m = prob[4] == 1
    assert np.array_equal(chain[:, 1:][:, :, m], chain1[:, 1:][:, :, m])
    assert np.array_equal(acc, acc1)
    for r in res:
        np.testing.assert_allclose(r, res1, rtol=0, atol=1e-9)


@pytest.mark.parametrize('triage', ['0', '1'])
@pytest.mark.parametrize('case', ['rowsite-13x13', 'cluster-41x41'])
def test_fused_peer_memory_exchange_equals_one_context(nat, monkeypatch, case, triage):
    """The records go straight into the other tiles' boxes and the appliers wait on flags: no
    collective, no host round trip per phase (here the tiles share one GPU, each on its own
    stream; on a multi-GPU box the same stores travel over NVLink).  Both appliers: a cluster (or
    CTA) per slot of every tile, and the triage pass (one thread per record) ahead of clusters for
    the records that reach into the tile's region -- the library picks by grid size, forced here."""
    monkeypatch.setenv('D3D_TILE_TRIAGE', triage)
    if case == 'rowsite-13x13':
        prob, chains, n_it, n_tiles = _problem(16, 30, 34, (13, 13), 3), 2, 3, 3
    else:
        prob, chains, n_it, n_tiles = _problem(64, 44, 50, (41, 41), 5), 1, 2, 2
    chain1, lik1, acc1, res1 = _single(nat, prob, chains, n_it)
    chain, lik, acc, res, partial, sw, regions = _tiled(nat, prob, chains, n_it, n_tiles, fused=True)
    m = prob[4] == 1
    assert sw.fused
    assert np.array_equal(chain[:, 1:][:, :, m], chain1[:, 1:][:, :, m])
    assert np.array_equal(lik[:, 1:][:, :, m], lik1[:, 1:][:, :, m])
    assert np.array_equal(acc, acc1)
    for r in res:
        np.testing.assert_allclose(r, res1, rtol=0, atol=1e-9)


@pytest.mark.parametrize('triage', ['0', '1'])
def test_fused_exchange_gives_up_on_a_silent_peer(nat, monkeypatch, triage):
    """A peer that never publishes its phase must not hang the GPU: the applier's wait is bounded
    (D3D_TILE_TIMEOUT_S, 30 s by default; 2 s here), later phases give up at once, and the error
    surfaces through the C ABI."""
    import time
    monkeypatch.setenv('D3D_TILE_TIMEOUT_S', '2')
    monkeypatch.setenv('D3D_TILE_TRIAGE', triage)
    prob = _problem(8, 12, 14, (5, 5), 2)
    data, var, fsf, lsf, mask, init = prob
    ctx, _, _ = make_ctx(nat, data, var, fsf, lsf, mask=mask, seed=3)
    ctx.set_params(init[None])
    ctx.forward(write_err=True)
    ctx.set_tile(0, 12, 0, 7)
    box = ctx.fused_init(2, 0)
    ctx.fused_connect(1, box=box)              # "tile 1" never runs: nobody raises its flag
    t0 = time.perf_counter()
    ctx.sweep_fused(1, 2)                      # 50 phases, every applier waits on tile 1
    with pytest.raises(nat.NativeError) as err:
        ctx.chain_control()
    assert 'timed out' in str(err.value)
    assert time.perf_counter() - t0 < 30.0


def test_tiling_float32_storage_and_refresh(nat):
    prob = _problem(12, 20, 22, (7, 7), 8)
    chain1, lik1, acc1, res1 = _single(nat, prob, 1, 3, dtype=nat.F32)
    chain, lik, acc, res, partial, sw, regions = _tiled(nat, prob, 1, 3, 2, dtype=nat.F32)
    m = prob[4] == 1
    assert np.array_equal(chain[:, 1:][:, :, m], chain1[:, 1:][:, :, m])
    # with a refresh after every sweep the chain changes only by the rounding of the refresh
    chain_r = _tiled(nat, prob, 1, 3, 2, refresh_every=1)[0]
    ref = _single(nat, prob, 1, 1)[0]
    np.testing.assert_allclose(chain_r[:, 1][:, m], ref[:, 1][:, m], rtol=1e-3, atol=1e-3)


def test_tile_api_errors(nat):
    prob = _problem(8, 10, 10, (5, 5), 1)
    data, var, fsf, lsf, mask, init = prob
    ctx, _, _ = make_ctx(nat, data, var, fsf, lsf, mask=mask)
    with pytest.raises(nat.NativeError):
        ctx.colour_phase(1, 0, 0)                      # no parameters yet
    ctx.set_params(init[None])
    ctx.forward(write_err=True)
    with pytest.raises(nat.NativeError):
        ctx.colour_phase(1, 0, 0)                      # no tile yet
    with pytest.raises(nat.NativeError):
        ctx.set_tile(0, 11, 0, 10)
    ctx.set_tile(0, 10, 0, 10)
    with pytest.raises(nat.NativeError):
        ctx.colour_phase(1, 5, 0)
    rec = np.zeros((ctx.record_slots(), 8))
    ctx.colour_phase(1, 1, 2, rec)
    sites = rec[:, 0][rec[:, 0] >= 0].astype(int)
    ys, xs = np.divmod(sites, 10)
    assert sites.size and (ys % 5 == 1).all() and (xs % 5 == 2).all()
    assert (mask[ys, xs] == 1).all()
    p = ctx.get_params()[0]
    assert np.array_equal(rec[rec[:, 0] >= 0][:, 2:5], p[ys, xs])


def test_tiled_nccl_multi_gpu(nat):
    """One rank per GPU, records exchanged by NCCL (needs >= 2 GPUs; the single-GPU box of the
    round-end run skips it)."""
    import os
    import subprocess
    import sys
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip('needs at least 2 GPUs')
    n = 2 if n < 4 else 4
    here = os.path.dirname(os.path.abspath(__file__))
    out = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1',
                          '--nproc-per-node', str(n), '--master-addr', '127.0.0.1',
                          '--master-port', '29533', os.path.join(here, 'run_tiled_nccl.py')],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and 'TILED NCCL OK' in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
