#!/usr/bin/env python
"""
bench.py -- logL evaluations/s and Gibbs sweeps/s of the deconv3d likelihood hot
path on B200 (BASELINE.json metric), with roofline, CPU baseline and clocks.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload cfg2x256|cfg2|cfg1|cfg5] [--mode sequential|coloured]
                    [--chains C] [--sweeps S] [--dtype f64|f32]

One "step" = S Gibbs sweeps (default 20) of every chain on the GPU: for each masked
spaxel of each chain one proposal + delta-logL evaluation + accept test + Gibbs
amplitude draw + residual update.  value = logL evaluations (= site updates) per
second summed over all ranks, inputs resident in HBM; e2e = the same metric through
the public ``Run(...)`` API with host (numpy) buffers in and chain rows out.

For N > 1 launch under torchrun (one rank per GPU); chains are sharded by rank, there
is no collective in the sweep (scaling: weak).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = 'logL evals/s'
UNIT = 'evals/s'


# ----------------------------------------------------------------------------------
def build_workload(name, chains):
    """Returns dict(data [n,D,H,W] builder inputs...) describing the workload on the host."""
    from deconv3d_b200 import synthetic
    if name in ('cfg2', 'cfg2x256', 'cfg2_fsf21'):
        D = H = W = 40
        fs = 21 if name == 'cfg2_fsf21' else 13
        inst = synthetic.muse_wfm_instrument('moffat', fs)
        truth = synthetic.halpha_truth(D, H, W)
        return dict(name=name, D=D, H=H, W=W, inst=inst, truth=truth[None], n_cubes=1,
                    chains_per_cube=chains, var_kind='cube', sigma=0.05,
                    desc='synthetic MUSE WFM Halpha cube 40x40x40, Moffat FWHM 0.8" beta 2.5 '
                         '%dx%d, MUSE LSF, variance cube 0.05^2, %d chain(s)/GPU' % (fs, fs, chains))
    if name == 'cfg5':
        D = H = W = 32
        inst = synthetic.muse_wfm_instrument('moffat', 11)
        rs = np.random.RandomState(5)
        truth = np.stack([synthetic.halpha_truth(D, H, W) * np.array([0.5 + rs.rand(), 1.0, 1.0])
                          for _ in range(chains)])
        return dict(name=name, D=D, H=H, W=W, inst=inst, truth=truth, n_cubes=chains,
                    chains_per_cube=1, var_kind='cube', sigma=0.05,
                    desc='survey batch: %d independent synthetic galaxies 32x32x32 per GPU, '
                         'FSF 11x11, per-galaxy data + variance' % chains)
    if name == 'cfg4':
        D, H, W = 64, chains, chains            # `chains` carries the field size here (1 chain)
        # 41x41 Moffat stamp of FWHM 3 px, beta 2 (the cube keeps the 0.2"/px
        # metadata of build_cube, so 3 px = 0.6")
        inst = synthetic.muse_wfm_instrument('moffat', 41, fsf_fwhm=0.6, beta=2.0)
        truth = synthetic.narrow_field_truth(D, H, W)
        return dict(name=name, D=D, H=H, W=W, inst=inst, truth=truth[None], n_cubes=1,
                    chains_per_cube=1, var_kind='cube', sigma=0.05,
                    desc='one oversized synthetic cube %dx%dx64, Moffat FWHM 3 px beta 2 stamp 41x41, '
                         'MUSE LSF (P = 64: full wrap), variance cube, 1 chain, spatial tiles' % (H, W))
    if name == 'cfg1':
        data = np.load(os.path.join(ROOT, 'tests', 'golden', 'muse_cube_01.npz'))['data'] * 1e20
        from deconv3d_b200 import MUSE
        return dict(name=name, D=30, H=30, W=30, inst=MUSE(), data=data[None], n_cubes=1,
                    chains_per_cube=chains, var_kind='scalar',
                    desc='bundled MUSE test cube 30x30x30 x1e20, MUSE() defaults (Gaussian FSF '
                         '13x13), scalar variance, %d chain(s)/GPU' % chains)
    raise SystemExit('unknown workload %s' % name)


def realise(wl, rank):
    """Host arrays of the workload: data, variance, fsf, lsf, boundaries."""
    from deconv3d_b200 import _native, MUSE
    from deconv3d_b200.math_utils import median_clip
    D, H, W = wl['D'], wl['H'], wl['W']
    cube0 = MUSE().build_cube(np.zeros((D, H, W)))
    fsf = np.asarray(wl['inst'].fsf.as_image(cube0), dtype=np.float64)
    lsf = wl['inst'].lsf.as_vector(cube0)
    if 'data' in wl:
        data = wl['data']
    else:
        # noiseless cube from the product's own forward model, then Gaussian noise
        n = wl['n_cubes']
        ctx = _native.Context(0 if 'LOCAL_RANK' not in os.environ else int(os.environ['LOCAL_RANK']))
        pmax = np.array([[100., D - 1, D]] * n)
        ctx.set_problem(np.ones((n, D, H, W)), np.ones(n), fsf, lsf, np.zeros((n, 3)), pmax,
                        [0, .1, .1], np.ones(n))
        clean = ctx.simulate(wl['truth'])
        ctx.close()
        from deconv3d_b200 import synthetic
        data = clean + synthetic.noise(clean.shape, wl['sigma'], 1234 + rank)
    n = data.shape[0]
    if wl['var_kind'] == 'cube':
        var = np.full(data.shape, wl['sigma'] ** 2)
    else:
        var = np.zeros(n)
        for i in range(n):
            _, s, _ = median_clip(np.copy(data[i][2:-2, 2:-4, 2:4]), 2.5)
            var[i] = (s if s != 0 else 1e-20) ** 2
    pmin = np.zeros((n, 3))
    pmax = np.array([[data[i].max() / fsf.max(), D - 1, D] for i in range(n)])
    return dict(data=data, var=var, fsf=fsf, lsf=lsf, pmin=pmin, pmax=pmax,
                prior=pmax[:, 0] ** 2)


# ----------------------------------------------------------------------------------
class ClockSampler(object):
    """Samples SM clocks and throttle reasons with nvidia-smi during the timed region."""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index = index
        self.samples = []
        self.stop = threading.Event()
        self.thread = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self.stop.is_set():
            try:
                out = subprocess.run(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                      '--format=csv,noheader,nounits'], capture_output=True,
                                     text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([v.strip() for v in out.split(',')])
            except Exception:           # noqa: BLE001
                pass
            self.stop.wait(0.2)

    def __enter__(self):
        self.thread.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.thread.join(timeout=6)

    def summary(self):
        if not self.samples:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': [], 'samples': 0}
        sm = sorted(float(s[0]) for s in self.samples)
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = [n for i, n in enumerate(names)
                   if any(s[3 + i].lower().startswith('active') for s in self.samples)]
        return {'sm_mhz': sm[len(sm) // 2], 'sm_max_mhz': float(self.samples[0][1]),
                'reasons': reasons, 'samples': len(self.samples),
                'power_w_max': max(float(s[2]) for s in self.samples)}


def measured_traffic(workload, chains, sweeps, dtype, mode, kernel):
    """DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture
    (profiles/*_traffic.json), only when it was taken on the same workload AND the same kernel
    (the library picks the sweep kernel by problem shape and chain count; d3d_last_kernel)."""
    import glob
    for path in sorted(glob.glob(os.path.join(ROOT, 'profiles', '*_traffic.json')), reverse=True):
        try:
            t = json.load(open(path))
            m = t.get('match', {})
            if (m.get('workload') == workload and m.get('chains') == chains and
                    m.get('sweeps') == sweeps and m.get('dtype') == dtype and m.get('mode') == mode and
                    str(t.get('kernel', '')).split('<')[0] == str(kernel).split('<')[0]):
                return float(t['dram_bytes_per_launch']), os.path.relpath(path, ROOT)
        except Exception:               # noqa: BLE001
            continue
    return None, None


def measured_profile(kernel):
    """Issue-slot utilisation of the dominant kernel from the committed ncu --set full capture
    (profiles/*_sweep_metrics.json) of that same kernel."""
    import glob
    for path in sorted(glob.glob(os.path.join(ROOT, 'profiles', '*_sweep_metrics.json')), reverse=True):
        try:
            t = json.load(open(path))
            if str(t.get('kernel', '')).split('<')[0] != str(kernel).split('<')[0]:
                continue
            return {'issue_slot_frac': float(t['issue_slot_frac']), 'source': os.path.relpath(path, ROOT)}
        except Exception:               # noqa: BLE001
            continue
    return {}


def kernel_name(ctx, mode):
    return ctx.last_kernel()


def measured_peak():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        try:
            return float(json.load(open(p))['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
        except Exception:               # noqa: BLE001
            pass
    return 6650.0, 'fallback (B200_PROFILING.md)'


# ----------------------------------------------------------------------------------
def cpu_reference(wl_name, sweeps, n_procs, sample_sites=None):
    """The reference's CPU path (oracle literal port, lib/run.py:344-519 incl. the
    full-cube temporaries and numpy-FFT spectral convolution) on the same workload;
    n_procs independent chains, one per process."""
    import multiprocessing as mp
    ctx = mp.get_context('fork')
    with ctx.Pool(n_procs) as pool:
        res = pool.map(_cpu_chain, [(wl_name, sweeps, k) for k in range(n_procs)])
    wall = max(r[0] for r in res)
    updates = sum(r[1] for r in res)
    return updates / wall, wall, updates


def _cpu_chain(args):
    wl_name, sweeps, k = args
    os.environ['OMP_NUM_THREADS'] = '1'
    from oracle import reference_port as port, streams
    from deconv3d_b200 import synthetic, MUSE
    # the same cube family as the GPU arm, built on the CPU by the oracle's forward model
    if wl_name in ('cfg2', 'cfg2x256'):
        D = H = W = 40
        inst = synthetic.muse_wfm_instrument('moffat', 13)
        truth = synthetic.halpha_truth(D, H, W)
    elif wl_name == 'cfg5':
        D = H = W = 32
        inst = synthetic.muse_wfm_instrument('moffat', 11)
        truth = synthetic.halpha_truth(D, H, W)
    else:
        D = H = W = 30
        inst = MUSE()
        truth = None
    cube0 = MUSE().build_cube(np.zeros((D, H, W)))
    fsf = np.asarray(inst.fsf.as_image(cube0))
    lsf = inst.lsf.as_vector(cube0)
    mask = np.ones((H, W))
    if truth is None:
        data = np.load(os.path.join(ROOT, 'tests', 'golden', 'muse_cube_01.npz'))['data'] * 1e20
        var = None
    else:
        data = -port.compute_error_in_one_step(np.zeros((D, H, W)), truth, fsf, lsf, mask) \
            + synthetic.noise((D, H, W), 0.05, 1234)
        var = np.full(data.shape, 0.05 ** 2)
    st = streams.PhiloxStream(42, k)
    # warm-up sweep excluded from the timing: run 2 iterations first (1 sweep)
    t0 = time.perf_counter()
    port.run_chain(data, fsf, lsf, st, variance_cube=var, max_iterations=2, refresh_every=0)
    t_setup = time.perf_counter() - t0
    t0 = time.perf_counter()
    port.run_chain(data, fsf, lsf, st, variance_cube=var, max_iterations=sweeps + 1,
                   refresh_every=0)
    t_all = time.perf_counter() - t0
    # both calls pay the same initial simulation; the difference is (sweeps-1) sweeps
    dt = max(t_all - t_setup, 1e-9)
    return dt, (sweeps - 1) * H * W


# ----------------------------------------------------------------------------------
_STDOUT_FD = None


def quiet_stdout():
    """Libraries may write to stdout (NCCL prints its version line there when NCCL_DEBUG is set):
    send everything to stderr until the result line, so that stdout carries ONE JSON line."""
    global _STDOUT_FD
    if _STDOUT_FD is None:
        sys.stdout.flush()
        _STDOUT_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    sys.stdout.flush()
    if _STDOUT_FD is not None:
        os.dup2(_STDOUT_FD, 1)
    print(json.dumps(line))
    sys.stdout.flush()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='cfg2x256', choices=['cfg2x256', 'cfg2', 'cfg1', 'cfg5', 'cfg4'])
    ap.add_argument('--field', type=int, default=256, help='cfg4: field side in spaxels')
    ap.add_argument('--exchange', default='fused', choices=['fused', 'nccl'], help='cfg4: record exchange')
    ap.add_argument('--mode', default='sequential', choices=['sequential', 'coloured'])
    ap.add_argument('--chains', type=int, default=None, help='chains (or galaxies) per GPU')
    ap.add_argument('--sweeps', type=int, default=None, help='Gibbs sweeps per step')
    ap.add_argument('--dtype', default='f64', choices=['f64', 'f32'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-probes', action='store_true', help='skip the sub-records of the other configurations')
    ap.add_argument('--cpu-sweeps', type=int, default=None)
    args = ap.parse_args()
    quiet_stdout()
    args.warmup = max(args.warmup, 0)

    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    chains = args.chains if args.chains is not None else \
        {'cfg2x256': 256, 'cfg2': 1, 'cfg1': 1, 'cfg5': 512, 'cfg4': 1}[args.workload]
    sweeps = args.sweeps if args.sweeps is not None else \
        {'cfg2x256': 20, 'cfg2': 200, 'cfg1': 200, 'cfg5': 20, 'cfg4': 1}[args.workload]
    if args.workload == 'cfg4':
        return bench_tiled(args, rank, local_rank, world, sweeps)

    if args.impl == 'reference':
        if rank != 0:
            return
        n_procs = os.cpu_count() or 1
        cs = args.cpu_sweeps or 3
        per_step = []
        for _ in range(max(args.warmup, 0) and 1):
            cpu_reference(args.workload, 2, n_procs)
        for _ in range(args.steps):
            v, wall, upd = cpu_reference(args.workload, cs, n_procs)
            per_step.append((v, wall))
        v = float(np.mean([p[0] for p in per_step]))
        wl = build_workload(args.workload, n_procs)          # the chains THIS arm runs: one per host core
        wl['desc'] += ' (reference arm: %d independent chain(s), one process per host core; the GPU arm ' \
                      'runs %d per GPU on the same cube)' % (n_procs, chains)
        line = {
            'impl': 'reference', 'metric': METRIC, 'value': v, 'unit': UNIT, 'n_gpus': args.gpus,
            'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': 1e3 * float(np.mean([p[1] for p in per_step])),
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64',
            'data': 'synthetic', 'config': {'workload': wl['desc'], 'mode': 'sequential (reference order)'},
            'cpu_baseline': {'value': v, 'unit': UNIT, 'cores': n_procs, 'kind': 'port',
                             'sample': '%d independent chains (one process per host core), %d timed '
                                       'sweeps each, oracle literal numpy/FFT path' % (n_procs, cs - 1)},
            'e2e': {'value': v, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'sweeps_per_s': v / (wl['H'] * wl['W']),
        }
        emit(line)
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device: deconv3d_b200 has no CPU fallback')
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))

    from deconv3d_b200 import _native, rtnorm_tables
    wl = build_workload(args.workload, chains)
    arrays = realise(wl, rank)
    D, H, W = wl['D'], wl['H'], wl['W']
    n_sites = H * W
    n_chains = wl['n_cubes'] * wl['chains_per_cube']
    mode = _native.SEQ_EXACT if args.mode == 'sequential' else _native.COLOURED

    ctx = _native.Context(local_rank, _native.F64 if args.dtype == 'f64' else _native.F32)
    # an explicit stream: the legacy default stream has handle 0, which d3d_ctx_set_stream reads
    # as "use the library's own stream" -- events and the L2 flush would then be unordered with
    # the sweep kernels
    stream = torch.cuda.Stream(device=local_rank)
    ctx.set_stream(stream.cuda_stream)
    ctx.set_rtnorm_tables(*rtnorm_tables.tables())
    ctx.set_rng(42, rank * n_chains)
    ctx.set_problem(arrays['data'], arrays['var'], arrays['fsf'], arrays['lsf'], arrays['pmin'],
                    arrays['pmax'], [0, 0.1, 0.1], arrays['prior'],
                    chains_per_cube=wl['chains_per_cube'])
    ctx.init_params_uniform()
    ctx.forward(write_err=True)

    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device='cuda')
    it = 1
    for _ in range(args.warmup):
        ctx.sweep(it, sweeps, mode=mode, refresh_every=1000, min_acceptance_rate=0.0)
        it += sweeps
    launches0 = ctx.counters()['kernel_launches']

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    step_ms, kern_ms, bytes_algo, updates = [], [], 0, 0
    barrier()
    with ClockSampler(local_rank) as clocks:
        for _ in range(args.steps):
            with torch.cuda.stream(stream):
                flush.fill_(1)                   # L2 flush between timed iterations (same stream)
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            _, _, ms = ctx.sweep(it, sweeps, mode=mode, refresh_every=1000,
                                 min_acceptance_rate=0.0)
            e1.record(stream)
            e1.synchronize()
            step_ms.append(e0.elapsed_time(e1))
            kern_ms.append(ms)
            c = ctx.counters()
            bytes_algo += c['last_sweep_bytes']
            updates += c['last_sweep_site_updates']
            it += sweeps
        barrier()
    launches = ctx.counters()['kernel_launches'] - launches0
    total_ms = float(np.sum(step_ms))
    t = torch.tensor([total_ms], dtype=torch.float64, device='cuda')
    u = torch.tensor([float(updates)], dtype=torch.float64, device='cuda')
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(u, op=dist.ReduceOp.SUM)
    total_ms_max, updates_all = float(t.item()), float(u.item())
    value = updates_all / (total_ms_max * 1e-3)

    # ---- end-to-end through the public API: host buffers in, chain rows out -------
    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(args, wl, arrays, rank, local_rank, world, sweeps)

    # cfg4 strong-scaling probe: every rank takes part (the one path with an exchange step)
    cfg4 = cfg4_big = None
    if args.workload == 'cfg2x256' and args.mode == 'sequential' and not args.no_probes:
        cfg4 = tiled_probe(args, rank, local_rank, world)
        # the same path on a field where a colour phase is several waves of CTAs on one GPU: the
        # size from which tiling over GPUs pays (749 ms per sweep on one B200, 326 on four, 252 on eight)
        cfg4_big = tiled_probe(args, rank, local_rank, world, field=1024)

    if world > 1 and rank != 0:
        dist.destroy_process_group()
        return

    peak, peak_src = measured_peak()
    traffic, traffic_src = measured_traffic(args.workload, n_chains, sweeps, args.dtype, args.mode,
                                            kernel_name(ctx, args.mode))
    prof = measured_profile(kernel_name(ctx, args.mode))
    kern_total_ms = float(np.sum(kern_ms))
    achieved = bytes_algo / (kern_total_ms * 1e-3) / 1e9
    state_mb = n_chains * D * H * W * (8 if args.dtype == 'f64' else 4) / 1e6
    fp64_peak = ctx.fp64_peak()
    # FP64 work of one site update: per window voxel one multiply + one FMA for the sums and one
    # FMA for the residual update (5 flop; the O(D) scalar work of the decision is not counted)
    vox_per_update = (bytes_algo / ((3 if wl['var_kind'] == 'cube' else 2) * (8 if args.dtype == 'f64' else 4))) \
        / max(1.0, float(updates))
    fp64_tflops = 5.0 * vox_per_update * float(updates) / (kern_total_ms * 1e-3) / 1e12
    line = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': total_ms_max / args.steps,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': args.dtype, 'data': 'synthetic',
        'config': {'workload': wl['desc'], 'mode': args.mode, 'sweeps_per_step': sweeps,
                   'chains_per_gpu': n_chains, 'sites_per_sweep': n_sites,
                   'residual_state_MB_per_gpu': state_mb,
                   'l2': 'explicit 256 MiB L2 flush between timed steps'},
        'sweeps_per_s': value / n_sites,
        'gpu_launches': int(launches),
        'clocks': clocks.summary(),
        'roofline': {
            # The byte model of SURVEY.md 8d (achieved / frac / nominal_hbm_frac) is kept as the
            # contract asks, but the kernel does not move those bytes: the window lives in
            # registers.  `bound` names what limits it (ncu: issue slots / dependent latency).
            'bound': 'issue', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s',
            'frac': achieved / peak, 'nominal_hbm_frac': achieved / peak,
            'traffic': traffic, 'traffic_source': traffic_src,
            'dram_frac': (traffic / (kern_total_ms / max(1, args.steps) * 1e-3) / 1e9 / peak) if traffic else None,
            'fp64_pipe_frac': fp64_tflops / fp64_peak, 'fp64_tflops': fp64_tflops,
            'fp64_peak_tflops': fp64_peak,
            'issue_slot_frac': prof.get('issue_slot_frac'), 'profile_source': prof.get('source'),
            'peak_source': peak_src,
            'kernel': kernel_name(ctx, args.mode),
            'algorithmic_bytes_per_launch': bytes_algo / max(1, args.steps),
            'kernel_ms_per_launch': kern_total_ms / max(1, args.steps),
            'note': 'achieved = algorithmic bytes / kernel time; algorithmic bytes = (3 with a variance '
                    'cube | 2 with a scalar variance) * s * D * sum_sites wh*ww per chain per sweep '
                    '(SURVEY.md 8d). The register-resident window re-uses 12 of 13 window columns, so '
                    'real DRAM traffic (traffic, dram_frac) is orders of magnitude lower; the kernel '
                    'is bound by instruction issue and the serial decision chain (issue_slot_frac, '
                    'fp64_pipe_frac = 5 flop per window voxel against the DFMA peak measured in this run)',
        },
    }
    if e2e is not None:
        line['e2e'] = e2e.pop('main')
        line.update(e2e)                                  # e2e_keep1, e2e_chain_on_device
    if cfg4 is not None:
        line['cfg4_1gpu' if world == 1 else 'cfg4_tiled'] = cfg4
    if cfg4_big is not None:
        line['cfg4_1024_1gpu' if world == 1 else 'cfg4_1024_tiled'] = cfg4_big
    if world == 1 and args.workload == 'cfg2x256' and args.mode == 'sequential' and not args.no_probes:
        # every other BASELINE.json configuration as a sub-record (CUDA events, state resident)
        line['cfg2_single_chain'] = sweep_probe(args, local_rank, stream, 'cfg2', 1, 200, 50,
                                                note='BASELINE configs[1] taken literally: ONE sequential-exact chain')
        line['cfg1'] = sweep_probe(args, local_rank, stream, 'cfg1', 1, 100, 20,
                                   note='configs[0]: bundled MUSE cube x1e20, MUSE() defaults, scalar variance, 1 chain')
        line['cfg3_coloured'] = sweep_probe(args, local_rank, stream, 'cfg2x256', 256, 10, 2, mode='coloured',
                                            note='configs[2]: colour-class updates, 256 chains')
        line['cfg5'] = sweep_probe(args, local_rank, stream, 'cfg5', 512, 10, 2,
                                   note='configs[4]: survey batch, 512 galaxies per GPU')
        line['cfg2_fsf21'] = sweep_probe(args, local_rank, stream, 'cfg2_fsf21', 148, 5, 1,
                                         note='cfg2 with the 21x21 stamp (SURVEY.md 8d secondary), 148 chains')
        line['forward_microbench'] = forward_microbench(args, local_rank, stream, fp64_peak)
    if world == 1 and not args.no_cpu_baseline:
        cs = args.cpu_sweeps or 3
        n_procs = min(os.cpu_count() or 1, 8)
        v, wall, upd = cpu_reference(args.workload, cs, 1)
        line['cpu_baseline'] = {
            'value': v, 'unit': UNIT, 'cores': 1, 'kind': 'port',
            'sample': '1 chain, %d timed sweeps (%d site updates) of the same cube on one host core '
                      '(numpy is single-threaded on this path); host has %d cores'
                      % (cs - 1, upd, os.cpu_count() or 1)}
    emit(line)
    if world > 1:
        dist.destroy_process_group()


def sweep_probe(args, local_rank, stream, workload, chains, sweeps, warm, mode='sequential', note=None):
    """One BASELINE.json configuration as a sub-record of the default line: set the problem up,
    `warm` untimed sweeps, then `sweeps` sweeps timed with CUDA events on the library's stream.
    Not part of the timed steps of the headline."""
    import torch
    from deconv3d_b200 import _native, rtnorm_tables
    try:
        wl = build_workload(workload, chains)
        arrays = realise(wl, 0)
        ctx = _native.Context(local_rank, _native.F64 if args.dtype == 'f64' else _native.F32)
        ctx.set_stream(stream.cuda_stream)
        ctx.set_rtnorm_tables(*rtnorm_tables.tables())
        ctx.set_rng(42, 0)
        ctx.set_problem(arrays['data'], arrays['var'], arrays['fsf'], arrays['lsf'], arrays['pmin'],
                        arrays['pmax'], [0, 0.1, 0.1], arrays['prior'], chains_per_cube=wl['chains_per_cube'])
        ctx.init_params_uniform()
        ctx.forward(write_err=True)
        nmode = _native.SEQ_EXACT if mode == 'sequential' else _native.COLOURED
        ctx.sweep(1, warm, mode=nmode, refresh_every=1000, min_acceptance_rate=0.0)
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        ctx.sweep(1 + warm, sweeps, mode=nmode, refresh_every=1000, min_acceptance_rate=0.0)
        e1.record(stream)
        e1.synchronize()
        ms = e0.elapsed_time(e1)
        c = ctx.counters()
        upd = c['last_sweep_site_updates']
        peak, _ = measured_peak()
        ctx.close()
        n_units = wl['n_cubes'] * wl['chains_per_cube']
        rec = {'value': upd / (ms * 1e-3), 'unit': UNIT,
               'sweeps_per_s': n_units * sweeps / (ms * 1e-3),
               'us_per_site_update': ms * 1e3 / upd, 'sweeps': sweeps, 'units': n_units,
               'workload': wl['desc'], 'mode': mode,
               'nominal_hbm_frac': c['last_sweep_bytes'] / (ms * 1e-3) / 1e9 / peak}
        if note:
            rec['note'] = note
        return rec
    except Exception as e:                              # noqa: BLE001 - a probe never kills the line
        return {'error': '%s: %s' % (type(e).__name__, str(e)[:300])}


def forward_microbench(args, local_rank, stream, fp64_peak_tflops, n=512):
    """Forward model (lib/run.py:999-1031: spectral LSF pass + spatial FSF pass + residual) of a
    survey batch of `n` cubes 32^3 for four FSF sizes: ms per call, fraction of the measured FP64
    FMA peak (2 D H W fh fw flop of the FSF pass) and of the measured HBM peak (B_fwd = 4 s D H W:
    lines written + read, data read, residual written)."""
    import torch
    from deconv3d_b200 import _native, MUSE
    from deconv3d_b200.spread_functions import MoffatFieldSpreadFunction
    out = []
    peak, _ = measured_peak()
    D = H = W = 32
    rs = np.random.RandomState(0)
    data = rs.rand(n, D, H, W)
    var = np.full(data.shape, 0.01)
    p = np.dstack([rs.rand(H, W) * 9, rs.rand(H, W) * (D - 1), 0.5 + rs.rand(H, W) * 3])
    params = np.broadcast_to(p, (n, H, W, 3)).copy()
    for fs in (3, 13, 21, 41):
        try:
            inst = MUSE(fsf=MoffatFieldSpreadFunction(fwhm=0.8, beta=2.5, size=fs))
            cube0 = MUSE().build_cube(np.zeros((D, H, W)))
            fsf = np.asarray(inst.fsf.as_image(cube0))
            lsf = inst.lsf.as_vector(cube0)
            ctx = _native.Context(local_rank, _native.F64)
            ctx.set_stream(stream.cuda_stream)
            ctx.set_problem(data, var, fsf, lsf, np.zeros((n, 3)), np.tile([100., D - 1, D], (n, 1)),
                            [0, .1, .1], np.ones(n))
            ctx.set_params(params)
            for _ in range(2):
                ctx.forward(write_err=True)
            reps = 5
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(reps):
                ctx.forward(write_err=True)
            e1.record(stream)
            e1.synchronize()
            ms = e0.elapsed_time(e1) / reps
            ctx.close()
            vox = float(n) * D * H * W
            out.append({'fsf': '%dx%d' % (fs, fs), 'ms': ms,
                        'fp64_tflops': 2.0 * vox * fs * fs / ms / 1e9,
                        'fp64_frac': 2.0 * vox * fs * fs / ms / 1e9 / fp64_peak_tflops,
                        'hbm_gbs': 4 * 8 * vox / ms / 1e6, 'hbm_frac': 4 * 8 * vox / ms / 1e6 / peak})
        except Exception as e:                          # noqa: BLE001
            out.append({'fsf': '%dx%d' % (fs, fs), 'error': '%s: %s' % (type(e).__name__, str(e)[:200])})
    return {'workload': 'forward model of %d cubes 32x32x32, f64, LSF sigma 0.9 px' % n, 'points': out,
            'fp64_peak_tflops': fp64_peak_tflops, 'hbm_peak_gbs': peak}


def tiled_probe(args, rank, local_rank, world, field=256, sweeps=1, warm=1):
    """cfg4 as a sub-record: ONE cube field x field x 64 (FSF 41x41), 1 chain, coloured sweep; with
    world > 1 the sites are tiled over the ranks and the outcome records of every colour phase
    are exchanged over NVLink peer memory (strong scaling: the cube is fixed).  Every rank takes
    part; the record is returned on every rank (max over ranks of the device time)."""
    import torch
    import torch.distributed as dist
    from deconv3d_b200 import _native, rtnorm_tables
    from deconv3d_b200 import dist as d3dist
    try:
        wl = build_workload('cfg4', field)
        arrays = realise(wl, 0)                       # the same cube on every rank
        D, H, W = wl['D'], wl['H'], wl['W']
        fh, fw = arrays['fsf'].shape
        ctx = _native.Context(local_rank, _native.F64 if args.dtype == 'f64' else _native.F32)
        ctx.set_rtnorm_tables(*rtnorm_tables.tables())
        ctx.set_rng(42, 0)
        ctx.set_problem(arrays['data'], arrays['var'], arrays['fsf'], arrays['lsf'], arrays['pmin'],
                        arrays['pmax'], [0, 0.1, 0.1], arrays['prior'], chains_per_cube=1)
        ctx.init_params_uniform()
        ctx.forward(write_err=True)
        if world == 1:
            # one GPU: the library's own coloured sweep (one launch per colour class, C loop)
            stream = torch.cuda.Stream(device=local_rank)
            ctx.set_stream(stream.cuda_stream)
            ctx.sweep(1, warm, mode=_native.COLOURED, refresh_every=0, min_acceptance_rate=0.0)
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            ctx.sweep(1 + warm, sweeps, mode=_native.COLOURED, refresh_every=0, min_acceptance_rate=0.0)
            e1.record(stream)
            e1.synchronize()
            ms = e0.elapsed_time(e1)
            exchange = 'none (one GPU)'
        else:
            sw = d3dist.TiledSweeper([ctx], (H, W), (fh, fw), fused=args.exchange == 'fused')
            stream = sw.stream
            sw.sweep(1, warm, refresh_every=0)
            torch.cuda.synchronize()
            dist.barrier()
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            sw.sweep(1 + warm, sweeps, refresh_every=0)
            e1.record(stream)
            e1.synchronize()
            t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device='cuda')
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
            exchange = ('P2P stores of outcome records into every peer + flags (no collective)' if sw.fused
                        else 'NCCL all-gather of outcome records per phase')
            sw.finish()
        ctx.close()
        return {'value': float(H * W) * sweeps / (ms * 1e-3), 'unit': UNIT, 'ms_per_sweep': ms / sweeps,
                'us_per_phase': ms * 1e3 / sweeps / (min(fh, H) * min(fw, W)), 'n_gpus': world,
                'tiles': '%dx%d' % d3dist.tile_grid(H, W, world), 'exchange': exchange,
                'scaling': 'strong', 'workload': wl['desc'], 'mode': 'coloured'}
    except Exception as e:                              # noqa: BLE001
        return {'error': '%s: %s' % (type(e).__name__, str(e)[:300])}


def bench_tiled(args, rank, local_rank, world, sweeps):
    """cfg4: ONE oversized cube, sites tiled over the ranks, coloured sweep with one exchange of
    outcome records (NCCL all-gather) per colour phase.  Strong scaling: the cube is fixed."""
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device: deconv3d_b200 has no CPU fallback')
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    from deconv3d_b200 import _native, rtnorm_tables
    from deconv3d_b200 import dist as d3dist
    wl = build_workload('cfg4', args.field)
    arrays = realise(wl, 0)                       # the same cube on every rank
    D, H, W = wl['D'], wl['H'], wl['W']
    fh, fw = arrays['fsf'].shape
    ctx = _native.Context(local_rank, _native.F64 if args.dtype == 'f64' else _native.F32)
    ctx.set_rtnorm_tables(*rtnorm_tables.tables())
    ctx.set_rng(42, 0)
    t0 = time.perf_counter()
    ctx.set_problem(arrays['data'], arrays['var'], arrays['fsf'], arrays['lsf'], arrays['pmin'],
                    arrays['pmax'], [0, 0.1, 0.1], arrays['prior'], chains_per_cube=1)
    ctx.init_params_uniform()                     # Philox-addressed: identical on every rank
    ctx.forward(write_err=True)
    setup_s = time.perf_counter() - t0
    sw = d3dist.TiledSweeper([ctx], (H, W), (fh, fw), fused=args.exchange == 'fused')
    stream = sw.stream
    it = 1
    for _ in range(args.warmup):
        sw.sweep(it, sweeps, refresh_every=0)
        it += sweeps

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    launches0 = ctx.counters()['kernel_launches']
    step_ms = []
    barrier()
    with ClockSampler(local_rank) as clocks:
        for _ in range(args.steps):
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            sw.sweep(it, sweeps, refresh_every=0)
            e1.record(stream)
            e1.synchronize()
            step_ms.append(e0.elapsed_time(e1))
            it += sweeps
        barrier()
    launches = ctx.counters()['kernel_launches'] - launches0
    # end to end: one sweep with the chain row and the likelihoods read back to the host
    chain = np.zeros((1, it + sweeps, H, W, 3))
    lik = np.zeros((1, it + sweeps, H, W))
    barrier()
    t0 = time.perf_counter()
    sw.sweep(it, sweeps, keep_one_in=1, refresh_every=0, chain_out=chain, lik_out=lik)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([float(np.sum(step_ms)), e2e_s], dtype=torch.float64, device='cuda')
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, e2e_s = float(t[0].item()), float(t[1].item())
    if rank != 0:
        dist.destroy_process_group()
        return
    n_sites = H * W
    updates = float(n_sites) * sweeps * args.steps
    value = updates / (total_ms * 1e-3)
    s = 8 if args.dtype == 'f64' else 4
    fhh, fhw = (fh - 1) // 2, (fw - 1) // 2
    sum_wh = sum(min(y + fhh + 1, H) - max(y - fhh, 0) for y in range(H))
    sum_ww = sum(min(x + fhw + 1, W) - max(x - fhw, 0) for x in range(W))
    bytes_sweep = 3.0 * s * D * sum_wh * sum_ww
    peak, peak_src = measured_peak()
    achieved = bytes_sweep * sweeps * args.steps / (total_ms * 1e-3) / 1e9
    line = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': total_ms / args.steps, 'higher_is_better': True,
        'scaling': 'strong', 'vs_baseline': None, 'dtype': args.dtype, 'data': 'synthetic',
        'config': {'workload': wl['desc'], 'mode': 'coloured, tiles %dx%d' % d3dist.tile_grid(H, W, world),
                   'sweeps_per_step': sweeps, 'sites_per_sweep': n_sites,
                   'phases_per_sweep': min(fh, H) * min(fw, W),
                   'exchange': ('P2P stores of %d records x 64 B into every peer + flags (no collective)' if sw.fused
                                else 'NCCL all-gather of %d records x 64 B per rank and phase') % sw.slots,
                   'l2': 'window traffic per sweep (%.0f GB) exceeds L2' % (bytes_sweep / 1e9)},
        'sweeps_per_s': value / n_sites, 'gpu_launches': int(launches),
        'clocks': clocks.summary(),
        'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': peak * world, 'unit': 'GB/s',
                     'frac': achieved / (peak * world), 'traffic': None, 'peak_source': peak_src,
                     'kernel': 'sweep_colour_generic_kernel',
                     'note': 'whole-step figure (phase kernels + exchange + appliers), all GPUs'},
        'e2e': {'value': n_sites * sweeps / e2e_s, 'unit': UNIT,
                'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': int(n_sites * 4 * 8 * sweeps),
                'ms_per_step': e2e_s * 1e3, 'setup_s': setup_s,
                'api': 'TiledSweeper.sweep(..., chain_out, lik_out)'},
    }
    emit(line)
    if world > 1:
        dist.destroy_process_group()


def run_e2e(args, wl, arrays, rank, local_rank, world, sweeps):
    """Same metric through ``Run(...)``: numpy cube/variance in, results out.  Three variants:
    main                  keep_one_in = sweeps/2: two chain rows per chain come back
    e2e_keep1             the reference's default keep_one_in = 1: EVERY row comes back (262 MB per
                          step at the default workload)
    e2e_chain_on_device   the chain stays in HBM, the posterior mean is reduced there
                          (d3d_chain_mean); parameters + both output cubes come back"""
    import logging
    import torch
    import torch.distributed as dist
    from deconv3d_b200 import Run, MUSE
    if wl['n_cubes'] != 1:
        return None
    logging.getLogger('deconv3d').setLevel(logging.WARNING)
    cube = MUSE().build_cube(arrays['data'][0])
    var = arrays['var'][0] if wl['var_kind'] == 'cube' else None
    chains = wl['chains_per_cube']
    h2d = arrays['data'][0].nbytes + (var.nbytes if var is not None else 8) + arrays['fsf'].nbytes \
        + arrays['lsf'].nbytes
    updates = sweeps * wl['H'] * wl['W'] * chains * world

    def one(keep, on_device, reps):
        kw = dict(variance=var, max_iterations=sweeps + 1, keep_one_in=keep, n_chains=chains,
                  seed=42, first_chain_id=rank * chains, device=local_rank,
                  mode=args.mode, dtype='float64' if args.dtype == 'f64' else 'float32',
                  min_acceptance_rate=0.0)
        if on_device:
            kw['chain_on_device'] = True
        Run(cube, wl['inst'], **kw)                       # warm-up
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        # median of `reps` calls: the host side of a shared box is noisy (page faults of the fresh
        # chain arrays, other tenants), the device side is not
        times = []
        for _ in range(reps):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            run = Run(cube, wl['inst'], **kw)
            torch.cuda.synchronize()
            times.append(time.perf_counter() - t0)
        dt = float(np.median(times))
        t = torch.tensor([dt], dtype=torch.float64, device='cuda')
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
        if on_device:
            d2h = run.parameters_all.nbytes + 2 * arrays['data'][0].nbytes   # posterior means + both output cubes
        else:
            d2h = run.chains[:, 1:].nbytes + run.all_likelihoods[:, 1:].nbytes + run.chains[:, 0].nbytes
        return {'value': updates / dt, 'unit': UNIT, 'h2d_bytes_per_step': int(h2d),
                'd2h_bytes_per_step': int(d2h), 'ms_per_step': dt * 1e3,
                'ms_per_call_all': [round(t * 1e3, 1) for t in times],
                'api': 'Run(cube, instrument, variance=..., max_iterations=%d, keep_one_in=%d, '
                       'n_chains=%d%s)' % (sweeps + 1, keep, chains, ', chain_on_device=True' if on_device else '')}

    out = {'main': one(max(1, sweeps // 2), False, 5)}
    for key, keep, dev in (('e2e_keep1', 1, False), ('e2e_chain_on_device', 1, True)):
        try:
            out[key] = one(keep, dev, 3)
        except Exception as e:                          # noqa: BLE001
            out[key] = {'error': '%s: %s' % (type(e).__name__, str(e)[:300])}
    return out


if __name__ == '__main__':
    main()
