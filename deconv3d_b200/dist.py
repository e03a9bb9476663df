"""
Multi-GPU partitioning of the hot path.

Chains (and survey galaxies) are independent units: they are sharded across the
ranks of one node with NO collective inside the sweep; every rank owns a
contiguous block of unit ids and draws from the streams (seed, unit id), so the
union of the shards is bit-identical to a single-GPU run over all units.  The
only communication is the final gather of the (small) per-unit summaries.

``torch.distributed`` is plumbing here: NCCL on the GPU box, gloo in the CPU
tests.
"""
import numpy as np

__all__ = ['shard_range', 'shard_sizes', 'gather_units', 'run_chains_sharded']


def shard_sizes(n_units, world):
    """Balanced contiguous partition: the first n_units % world ranks get one more."""
    base, extra = divmod(int(n_units), int(world))
    return [base + (1 if r < extra else 0) for r in range(world)]


def shard_range(n_units, world, rank):
    """(first_unit, count) of ``rank``."""
    sizes = shard_sizes(n_units, world)
    return sum(sizes[:rank]), sizes[rank]


def gather_units(local, n_units, group=None):
    """All-gathers per-unit arrays (leading axis = the rank's units, in unit order) into
    the full [n_units, ...] array on every rank.  Works with NCCL (cuda tensors) and gloo."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return np.asarray(local)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    sizes = shard_sizes(n_units, world)
    local = np.ascontiguousarray(local)
    assert local.shape[0] == sizes[rank], 'rank %d holds %d units, expected %d' % (
        rank, local.shape[0], sizes[rank])
    cuda = dist.get_backend(group) == 'nccl'
    dev = torch.device('cuda', torch.cuda.current_device()) if cuda else torch.device('cpu')
    pad = max(sizes)
    buf = torch.zeros((pad,) + local.shape[1:], dtype=torch.from_numpy(local).dtype, device=dev)
    if sizes[rank]:
        buf[:sizes[rank]] = torch.from_numpy(local).to(dev)
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf, group=group)
    return np.concatenate([o[:sizes[r]].cpu().numpy() for r, o in enumerate(out)], axis=0)


def run_chains_sharded(make_run, n_chains, group=None):
    """Runs ``n_chains`` independent chains over the ranks of the process group.

    make_run(first_chain_id, count) -> object with ``chains`` [count, rows, H, W, P] (a
    ``deconv3d_b200.Run`` built with ``n_chains=count, first_chain_id=first_chain_id``).
    Returns the per-chain posterior means [n_chains, H, W, P] (mean of the last 20 % of
    each chain, lib/run.py:581-593) on every rank."""
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    first, count = shard_range(n_chains, world, rank)
    if count:
        run = make_run(first, count)
        rows = run.chains.shape[1]
        local = run.chains[:, int(0.8 * rows):].mean(axis=1)
    else:
        local = None
    if local is None:
        # shape is needed for the gather: ask a peer
        import torch
        shp = [None]
        dist.broadcast_object_list(shp, src=0, group=group)
        local = np.zeros((0,) + tuple(shp[0]))
    elif dist.is_initialized() and n_chains < world:
        shp = [local.shape[1:]]
        dist.broadcast_object_list(shp, src=0, group=group)
    return gather_units(local, n_chains, group)
