"""
Multi-GPU partitioning of the hot path.

Two cases.  (1) Chains (and survey galaxies) are independent units: they are sharded across the
ranks of one node with NO collective inside the sweep; every rank owns a
contiguous block of unit ids and draws from the streams (seed, unit id), so the
union of the shards is bit-identical to a single-GPU run over all units.  The
only communication is the final gather of the (small) per-unit summaries.

(2) ONE oversized cube (cfg4) does not shard by unit: its field is cut into
rectangular tiles, one per context/GPU, and the coloured sweep is driven phase
by phase (``TiledSweeper``).  The colour lattice is global, so the sites of a
phase have disjoint windows across tiles; after each phase the ranks all-gather
the 64-byte OUTCOME records of their site updates -- the one real exchange step
of the path -- and every context rebuilds the rank-1 residual change of the
remote records inside the region it reads (include/deconv3d_b200.h, "tiled").

``torch.distributed`` is plumbing here: NCCL on the GPU box, gloo in the CPU
tests.
"""
import numpy as np

__all__ = ['shard_range', 'shard_sizes', 'gather_units', 'run_chains_sharded',
           'tile_grid', 'tile_bounds', 'TiledSweeper']


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


# ---------------------------------------------------------------------------------------
# (2) one cube, spatial tiles
# ---------------------------------------------------------------------------------------
RECORD_DOUBLES = 8          # include/deconv3d_b200.h D3D_RECORD_DOUBLES


def tile_grid(H, W, n_tiles):
    """(rows, cols) with rows * cols == n_tiles whose tiles are closest to square
    (8 tiles of a square field -> 2 x 4)."""
    best = None
    for rows in range(1, int(n_tiles) + 1):
        if n_tiles % rows:
            continue
        cols = n_tiles // rows
        th, tw = H / float(rows), W / float(cols)
        score = max(th, tw) / max(min(th, tw), 1e-9)
        if best is None or score < best[0] - 1e-12:
            best = (score, rows, cols)
    return best[1], best[2]


def tile_bounds(H, W, n_tiles, index):
    """(y0, y1, x0, x1) of tile ``index`` (row-major over tile_grid); the tiles partition
    the field, sizes differ by at most one row / column."""
    rows, cols = tile_grid(H, W, n_tiles)
    ty, tx = divmod(int(index), cols)
    y0, ny = shard_range(H, rows, ty)
    x0, nx = shard_range(W, cols, tx)
    return y0, y0 + ny, x0, x0 + nx


class TiledSweeper(object):
    """Coloured MH-within-Gibbs sweeps (lib/run.py:344-537 in colour-class order) of ONE cube
    whose sites are split into tiles: ``contexts`` are this process' contexts (normally one,
    on its GPU; several in the single-GPU tests), all set up with the SAME problem, the same
    parameters and the same RNG; with a process group the tiles of all ranks form the grid.

    The result is independent of the tiling: the random numbers of a site update are
    addressed by (seed, chain, sweep, site), not by who draws them.
    """

    def __init__(self, contexts, field_hw, fsf_hw, group=None, use_torch=True, fused=False):
        self.ctxs = list(contexts)
        self.H, self.W = int(field_hw[0]), int(field_hw[1])
        self.fh, self.fw = int(fsf_hw[0]), int(fsf_hw[1])
        self.group = group
        self.dist = None
        self.world, self.rank = 1, 0
        try:
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized():
                self.dist = dist
                self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        except ImportError:
            pass
        self.nlocal = len(self.ctxs)
        self.n_tiles = self.world * self.nlocal
        self.tiles = []
        for i, ctx in enumerate(self.ctxs):
            b = tile_bounds(self.H, self.W, self.n_tiles, self.rank * self.nlocal + i)
            ctx.set_tile(*b)
            self.tiles.append(b)
        self.slots = self.ctxs[0].record_slots()
        self.device_records = bool(getattr(self.ctxs[0], 'device_records', True)) and use_torch
        if self.device_records:
            import torch
            dev = torch.device('cuda', torch.cuda.current_device())
            # phase kernels, NCCL and appliers all go to ONE explicit stream (the legacy default
            # stream has no handle a library context could be bound to)
            self.torch = torch
            self.stream = torch.cuda.Stream(device=dev)
            for ctx in self.ctxs:
                ctx.set_stream(self.stream.cuda_stream)
            self.local = torch.empty((self.nlocal * self.slots, RECORD_DOUBLES),
                                     dtype=torch.float64, device=dev)
            self.all = (torch.empty((self.world * self.nlocal * self.slots, RECORD_DOUBLES),
                                    dtype=torch.float64, device=dev)
                        if self.world > 1 else self.local)
        else:
            self.local = np.empty((self.nlocal * self.slots, RECORD_DOUBLES))
            self.all = self.local
        self.exchanges = 0
        self.fused = bool(fused) and self.n_tiles > 1
        self.n_classes = min(self.fh, self.H) * min(self.fw, self.W)
        if self.fused:
            self._connect_boxes()

    def _connect_boxes(self):
        """Fused exchange (include/deconv3d_b200.h): every context stores its records straight
        into the boxes of all the others -- device addresses inside this process, CUDA IPC handles
        across the ranks -- and its applier waits on flags in its own box.  No collective and no
        host synchronisation inside a phase."""
        if not self.device_records:
            raise ValueError('the fused exchange needs device contexts')
        if self.nlocal > 1:
            # several tiles on one GPU (tests): each context keeps its own stream, their kernels
            # must be able to run side by side (an applier spins until its peers have pushed)
            for ctx in self.ctxs:
                ctx.set_stream(0)
        first = self.rank * self.nlocal
        boxes = [ctx.fused_init(self.n_tiles, first + i) for i, ctx in enumerate(self.ctxs)]
        handles = [[ctx.fused_export() for ctx in self.ctxs]]
        if self.world > 1:
            gathered = [None] * self.world
            self.dist.all_gather_object(gathered, handles[0], group=self.group)
            handles = gathered
        for i, ctx in enumerate(self.ctxs):
            for g in range(self.n_tiles):
                r, k = divmod(g, self.nlocal)
                if g == first + i:
                    continue
                if r == self.rank:
                    ctx.fused_connect(g, box=boxes[k])
                else:
                    ctx.fused_connect(g, handle=handles[r][k])
        for ctx in self.ctxs:
            ctx.synchronize()
        if self.world > 1:
            self.dist.barrier(group=self.group)          # nobody pushes into a box that is not ready

    # -- one colour phase on every local tile, then the exchange, then the appliers --
    def _phase(self, it, cy, cx):
        if self.fused:
            phase = it * self.n_classes + cy * min(self.fw, self.W) + cx
            for ctx in self.ctxs:
                ctx.colour_phase_fused(it, cy, cx, phase)
            self.exchanges += 1
            return
        nb = self.slots * RECORD_DOUBLES * 8
        for i, ctx in enumerate(self.ctxs):
            if self.device_records:
                ctx.colour_phase(it, cy, cx, self.local.data_ptr() + i * nb)
            else:
                ctx.colour_phase(it, cy, cx, self.local[i * self.slots:(i + 1) * self.slots])
        if self.n_tiles == 1:
            return
        allrec = self.local
        if self.world > 1:
            if self.device_records:
                with self.torch.cuda.stream(self.stream):
                    self.dist.all_gather_into_tensor(self.all, self.local, group=self.group)
                allrec = self.all
            else:
                import torch
                mine = torch.from_numpy(self.local)
                parts = [torch.empty_like(mine) for _ in range(self.world)]
                self.dist.all_gather(parts, mine, group=self.group)
                allrec = np.concatenate([p.numpy() for p in parts], axis=0)
            self.exchanges += 1
        n = self.n_tiles * self.slots
        for ctx in self.ctxs:
            if self.device_records:
                ctx.apply_records(allrec.data_ptr(), n)
            else:
                ctx.apply_records(allrec, n)

    def sweep(self, first_iteration, n_iterations, keep_one_in=1, refresh_every=1000,
              min_acceptance_rate=0.0, chain_out=None, lik_out=None):
        """Iterations [first, first + n).  chain_out [n_chains, rows, H, W, 3] and lik_out
        [n_chains, rows, H, W] (numpy, optional) get row it // keep_one_in when
        it % keep_one_in == 0, complete on every rank.  Returns (accepted, iterations) per
        chain, identical on every rank."""
        lead = self.ctxs[0]
        for it in range(int(first_iteration), int(first_iteration + n_iterations)):
            if self.fused and self.nlocal == 1:
                # one context per process: the whole iteration is enqueued by the library
                lead.sweep_fused(it, 1, min_acceptance_rate)
                self.exchanges += self.n_classes
            else:
                for ctx in self.ctxs:
                    ctx.colour_begin(it, min_acceptance_rate)
                for cy in range(min(self.fh, self.H)):
                    for cx in range(min(self.fw, self.W)):
                        self._phase(it, cy, cx)
            if it % keep_one_in == 0:
                row = it // keep_one_in
                if chain_out is not None:
                    chain_out[:, row] = lead.get_params()
                if lik_out is not None:
                    lik_out[:, row] = lead.get_likelihoods()
            if refresh_every and it % refresh_every == 0:         # lib/run.py:525-534
                for ctx in self.ctxs:
                    ctx.forward()
            if min_acceptance_rate > 0.0:
                if not lead.chain_control()[2].any():
                    break
        acc, its, _ = lead.chain_control()
        return acc, its

    def finish(self):
        """Rebuilds the full residual of every context from the (complete) parameters: during
        the sweeps each context only maintains the part of it that its tile reads."""
        for ctx in self.ctxs:
            ctx.forward()
