#!/usr/bin/env python
"""
Same-box A/B timing of builds of the sequential sweep kernels (the tables of profiles/*_notes.md).

    python profiles/tools/ab_slide.py LIB.so[,ENV=VAL...] [LIB2.so ...] [--chains 1,148,256]
                                      [--workload cfg2x256|cfg5|cfg1] [--sweeps 20]

Every library (built with profiles/tools/build_variant.sh; the production one is
deconv3d_b200/libdeconv3d_b200.so) is timed in its own subprocess on the same GPU: cycles per
site update of one chain slot = kernel time / (sweeps * sites) at 1.965 GHz.  ENV=VAL pairs
after a comma are set for that run (e.g. D3D_PIPE=0 selects the sliding-window kernel).  A build
with -DD3D_PIPE_PROF also prints the per-warp wait accounting of CTA 0.
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def child(lib, chains, workload, sweeps):
    sys.path.insert(0, ROOT)
    import ctypes
    import numpy as np
    from deconv3d_b200 import _native, rtnorm_tables
    _native.LIB_PATH = os.path.abspath(lib)
    import bench
    wl = bench.build_workload(workload, chains)
    arr = bench.realise(wl, 0)
    ctx = _native.Context(0)
    ctx.set_rtnorm_tables(*rtnorm_tables.tables())
    ctx.set_rng(42, 0)
    ctx.set_problem(arr['data'], arr['var'], arr['fsf'], arr['lsf'], arr['pmin'], arr['pmax'],
                    [0, .1, .1], arr['prior'], chains_per_cube=wl['chains_per_cube'])
    ctx.init_params_uniform()
    ctx.forward()
    n_sites = wl['H'] * wl['W']
    lib_h = _native.load()
    prof = getattr(lib_h, 'd3d_debug_pipe_prof', None) if hasattr(lib_h, 'd3d_debug_pipe_prof') else None
    try:
        ctx.sweep(1, sweeps, refresh_every=0, min_acceptance_rate=0.0)
        if prof is not None:
            prof(None, 1)
        _, _, ms = ctx.sweep(1 + sweeps, sweeps, refresh_every=0, min_acceptance_rate=0.0)
    except Exception as e:                                  # noqa: BLE001
        print('%-40s chains %4d  ERROR %s' % (os.path.basename(lib), chains, str(e)[:400]))
        return
    n_units = wl['n_cubes'] * wl['chains_per_cube']
    slots = -(-n_units // 148) if n_units > 148 else 1      # chain slots per SM (balanced launch)
    us = ms * 1e3 / (sweeps * n_sites) / (n_units / 148.0 if n_units > 148 else 1.0)
    print('%-40s %s units %4d  %8.3f ms  %.3f us/site/SM = %6.0f cycles  (%.2f M evals/s)'
          % (os.path.basename(lib), workload, n_units, ms, us, us * 1965,
             n_units * sweeps * n_sites / ms / 1e3))
    if prof is not None:
        buf = (ctypes.c_ulonglong * 512)()
        prof(buf, 0)
        a = np.array(buf[:], dtype=np.float64).reshape(32, 16)
        cta_sweeps = sweeps * (n_units / 148.0 if n_units > 148 else 1.0)   # sweeps CTA 0 worked
        per_site = cta_sweeps * n_sites
        names = ['total', 'R:HSUM', 'W:DEC', 'B:SCAL', 'B:PART', 'X:PROF', '-', 'AP:FREE', 'R:PROF', 'W:switch', 'W:sums', 'W:store', 'W:update', 'W:head']
        print('   per-warp cycles per site of CTA 0 (items of CTA 0 only): ' + ' '.join('%8s' % n for n in names))
        # CTA 0 works ceil/floor(units*sweeps/148) sweeps; normalise by its own total of sweeps
        for w in range(32):
            if a[w, 0] == 0:
                continue
            print('   warp %2d ' % w + ' '.join('%8.0f' % (a[w, k] / per_site) for k in range(14)))
        if hasattr(lib_h, 'd3d_debug_pipe_cta') and n_units <= 148:
            cb = (ctypes.c_ulonglong * 2048)()
            lib_h.d3d_debug_pipe_cta(cb)
            c = np.array(cb[:], dtype=np.float64).reshape(1024, 2)[:n_units]
            cyc = c[:, 0] / (sweeps * n_sites)
            order = np.argsort(cyc)
            print('   per-CTA cycles per site (whole kernel): min %.0f  median %.0f  max %.0f' %
                  (cyc.min(), np.median(cyc), cyc.max()))
            print('   (cycles, smid) sorted: ' + ' '.join('%.0f@%d' % (cyc[i], int(c[i, 1])) for i in order))


if __name__ == '__main__':
    if len(sys.argv) > 1 and sys.argv[1] == '--child':
        child(sys.argv[2], int(sys.argv[3]), sys.argv[4], int(sys.argv[5]))
        sys.exit(0)
    libs, chains, workload, sweeps = [], [1, 148, 256], 'cfg2x256', 20
    args = sys.argv[1:]
    while args:
        a = args.pop(0)
        if a == '--chains':
            chains = [int(v) for v in args.pop(0).split(',')]
        elif a == '--workload':
            workload = args.pop(0)
        elif a == '--sweeps':
            sweeps = int(args.pop(0))
        else:
            libs.append(a)
    for spec in libs:
        parts = spec.split(',')
        env = dict(os.environ)
        for kv in parts[1:]:
            k, v = kv.split('=', 1)
            env[k] = v
        for c in chains:
            sys.stdout.flush()
            subprocess.run([sys.executable, os.path.abspath(__file__), '--child', parts[0], str(c), workload,
                            str(sweeps)], env=env)
