"""
Multi-GPU check of the tiled coloured sweep (run under torchrun, one rank per GPU):
the chain produced with the sites tiled over the ranks and the records exchanged by NCCL must
equal, bit for bit, the chain rank 0 gets alone on one GPU.  Launched by
tests/test_gpu_tiled.py::test_tiled_nccl_multi_gpu when the box has >= 2 GPUs:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29533 tests/run_tiled_nccl.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))


def main():
    import torch
    import torch.distributed as dist
    from deconv3d_b200 import _native as nat
    from deconv3d_b200 import dist as d3dist
    from test_gpu_tiled import _problem, make_ctx
    rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
    local = int(os.environ.get('LOCAL_RANK', rank))
    torch.cuda.set_device(local)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    ok = True
    for fused in (False, True, 'triage'):
      # fused exchange with both appliers: cluster per slot (D3D_TILE_TRIAGE=0) and triage + list (=1)
      if fused:
          os.environ['D3D_TILE_TRIAGE'] = '1' if fused == 'triage' else '0'
      for shape, fsf_shape, chains, n_it in (((16, 30, 34), (13, 13), 2, 3), ((64, 44, 50), (41, 41), 1, 2)):
        prob = _problem(shape[0], shape[1], shape[2], fsf_shape, 3)
        data, var, fsf, lsf, mask, init = prob
        D, H, W = data.shape

        def make(device):
            ctx = nat.Context(device)
            from test_gpu_parity import _tables, _oracle
            port, _, _ = _oracle()
            pmin, pmax = port.single_gaussian_boundaries(data, fsf)
            ctx.set_rtnorm_tables(*_tables())
            ctx.set_rng(77, 0)
            ctx.set_problem(data, var, fsf, lsf, pmin, pmax, (0.0, 0.1, 0.1), float(pmax[0]) ** 2,
                            mask=mask, chains_per_cube=chains)
            ctx.set_params(np.broadcast_to(init, (chains, H, W, 3)))
            ctx.forward(write_err=True)
            return ctx

        ctx = make(local)
        sw = d3dist.TiledSweeper([ctx], (H, W), fsf.shape, fused=bool(fused))
        chain = np.zeros((chains, n_it + 1, H, W, 3))
        lik = np.zeros((chains, n_it + 1, H, W))
        acc, its = sw.sweep(1, n_it, refresh_every=0, chain_out=chain, lik_out=lik)
        m = mask == 1
        if rank == 0:
            solo = make(local)
            chain1 = np.zeros_like(chain)
            lik1 = np.zeros_like(lik)
            acc1, _, _ = solo.sweep(1, n_it, mode=nat.COLOURED, refresh_every=0, min_acceptance_rate=0.0,
                                    chain_out=chain1, lik_out=lik1)
            same = (np.array_equal(chain[:, 1:][:, :, m], chain1[:, 1:][:, :, m]) and
                    np.array_equal(lik[:, 1:][:, :, m], lik1[:, 1:][:, :, m]) and
                    np.array_equal(acc, acc1))
            print('tiled over %d GPUs (%s), field %dx%d fsf %dx%d: %s (exchanges %d)'
                  % (world, ('fused P2P + triage' if fused == 'triage' else 'fused P2P') if fused else 'NCCL all-gather', H, W, fsf_shape[0], fsf_shape[1],
                     'IDENTICAL' if same else 'DIFFERENT', sw.exchanges))
            ok = ok and same
        # every rank must hold the same complete chain
        t = torch.from_numpy(chain).cuda()
        ref = t.clone()
        dist.broadcast(ref, 0)
        if not torch.equal(t, ref):
            print('rank %d: chain differs from rank 0' % rank)
            ok = False
    flag = torch.tensor([0 if ok else 1], device='cuda')
    dist.all_reduce(flag)
    dist.destroy_process_group()
    if flag.item():
        raise SystemExit(1)
    if rank == 0:
        print('TILED NCCL OK')


if __name__ == '__main__':
    main()
