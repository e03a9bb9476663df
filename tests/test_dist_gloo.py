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
