"""
Host-side multi-GPU logic on CPU: world_size-2 gloo process group (SURVEY.md section 8e:
chains/galaxies are sharded by unit, no collective in the sweep, one gather at the end).
"""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import ROOT


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_units, q):
    sys.path.insert(0, ROOT)
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    import torch.distributed as dist
    from deconv3d_b200 import dist as d3dist
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        first, count = d3dist.shard_range(n_units, world, rank)

        class FakeRun(object):
            """Stands in for Run(n_chains=count, first_chain_id=first): chain k's samples
            are a deterministic function of its global id, as with the Philox streams."""
            def __init__(self, first, count):
                ids = np.arange(first, first + count, dtype=np.float64)
                rows = 10
                self.chains = (ids[:, None, None, None, None] * 100.0 +
                               np.arange(rows)[None, :, None, None, None] +
                               np.zeros((count, rows, 3, 4, 3)))
        out = d3dist.run_chains_sharded(FakeRun, n_units)
        local = np.arange(first, first + count, dtype=np.float64)[:, None] * np.ones((count, 5))
        gathered = d3dist.gather_units(local, n_units)
        q.put((rank, first, count, out, gathered))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('n_units', [7, 2, 1])
def test_chain_sharding_and_gather_world2(n_units):
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_units, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res.sort(key=lambda t: t[0])
    firsts = [r[1] for r in res]
    counts = [r[2] for r in res]
    assert sum(counts) == n_units and firsts == [0, counts[0]]
    expect_mean = np.arange(n_units)[:, None, None, None] * 100.0 + 8.5 + np.zeros((n_units, 3, 4, 3))
    for r in res:
        np.testing.assert_allclose(r[3], expect_mean)                  # same on every rank
        np.testing.assert_array_equal(r[4][:, 0], np.arange(n_units))  # unit order preserved


def test_shard_sizes_properties():
    from deconv3d_b200.dist import shard_sizes, shard_range
    for n in (0, 1, 7, 256, 4096):
        for w in (1, 2, 4, 8):
            s = shard_sizes(n, w)
            assert sum(s) == n and max(s) - min(s) <= 1
            pos = 0
            for r in range(w):
                assert shard_range(n, w, r) == (pos, s[r])
                pos += s[r]


# ---------------------------------------------------------------------------------------
# one cube, spatial tiles: the per-phase exchange of the tiled coloured sweep over gloo
# ---------------------------------------------------------------------------------------
def _tile_problem():
    from oracle import reference_port as port
    rs = np.random.RandomState(11)
    D, H, W = 8, 9, 10
    yy, xx = np.mgrid[0:H, 0:W]
    z = np.arange(D)[:, None, None]
    data = 6.0 * np.exp(-(z - 3.5 - 0.1 * xx) ** 2 / (2 * 1.3 ** 2)) * np.exp(
        -((yy - 4) ** 2 + (xx - 5) ** 2) / 30.0) + 0.05 * rs.randn(D, H, W)
    fsf = port.gaussian_fsf_image(0.25, 0.2)            # 5x5
    lsf = port.gaussian_lsf_vector(0.0002675, 1.25e-4, D)
    var = 0.05 ** 2 * (1 + rs.rand(D, H, W))
    mask = (rs.rand(H, W) > 0.15).astype(float)
    init = np.dstack([rs.rand(H, W) * 4, 1 + rs.rand(H, W) * 5, 0.7 + rs.rand(H, W) * 2])
    return data, var, fsf, lsf, mask, init


def _tile_worker(rank, world, port_no, n_it, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port_no)
    import torch.distributed as dist
    from deconv3d_b200 import dist as d3dist
    from oracle_tile_ctx import OracleTileCtx
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        data, var, fsf, lsf, mask, init = _tile_problem()
        ctx = OracleTileCtx(data, var, fsf, lsf, mask, init, seed=13)
        sw = d3dist.TiledSweeper([ctx], data.shape[1:], fsf.shape)
        chain = np.zeros((1, n_it + 1) + init.shape)
        lik = np.zeros((1, n_it + 1) + init.shape[:2])
        acc, its = sw.sweep(1, n_it, refresh_every=0, chain_out=chain, lik_out=lik)
        q.put((rank, sw.tiles[0], chain, lik, int(acc[0]), int(its[0]), sw.exchanges))
    finally:
        dist.destroy_process_group()


def test_tiled_coloured_sweep_world2_matches_single_process():
    import torch.multiprocessing as mp
    from oracle import reference_port as port, streams
    n_it = 2
    data, var, fsf, lsf, mask, init = _tile_problem()
    order = port.colour_class_order(mask, fsf.shape[0], fsf.shape[1])
    ref = port.run_chain(data, fsf, lsf, streams.PhiloxStream(13, 0), mask=mask.copy(),
                         variance_cube=var, initial_parameters=init, max_iterations=n_it + 1,
                         min_acceptance_rate=0.0, refresh_every=0, site_order=order)
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port_no = _free_port()
    procs = [ctx.Process(target=_tile_worker, args=(r, 2, port_no, n_it, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res.sort(key=lambda t: t[0])
    assert res[0][1] == (0, 9, 0, 5) and res[1][1] == (0, 9, 5, 10)
    m = mask == 1
    for r in res:
        # every rank ends with the complete chain; the tiled walk differs from the single
        # process only by the residual being rebuilt instead of updated in place
        np.testing.assert_allclose(r[2][0, 1:][:, m], ref['chain'][1:][:, m], rtol=1e-7, atol=1e-7)
        np.testing.assert_allclose(r[3][0, 1:][:, m], ref['likelihoods'][1:][:, m], rtol=1e-6, atol=1e-6)
        assert r[4] == int(ref['accepted_count']) and r[5] == n_it + 1
        assert r[6] == n_it * fsf.shape[0] * fsf.shape[1]      # one exchange per colour phase
    np.testing.assert_array_equal(res[0][2], res[1][2])


def test_tile_partition_properties():
    from deconv3d_b200.dist import tile_grid, tile_bounds
    assert tile_grid(256, 256, 8) == (2, 4)          # 2 x 4 tiles of 128 x 64
    assert tile_bounds(256, 256, 8, 5) == (128, 256, 64, 128)
    for (H, W) in ((40, 40), (9, 10), (256, 64), (7, 300)):
        for n in (1, 2, 3, 4, 6, 8):
            cover = np.zeros((H, W), dtype=int)
            for i in range(n):
                y0, y1, x0, x1 = tile_bounds(H, W, n, i)
                cover[y0:y1, x0:x1] += 1
            assert (cover == 1).all()
