import sys, os, time, numpy as np
sys.path.insert(0, '/root/repo')
import bench
from deconv3d_b200 import _native, rtnorm_tables
if os.environ.get('D3D_LIB'): _native.LIB_PATH = os.environ['D3D_LIB']
field = int(sys.argv[1]) if len(sys.argv) > 1 else 256
wl = bench.build_workload('cfg4', field); arr = bench.realise(wl, 0)
ctx = _native.Context(0); ctx.set_rtnorm_tables(*rtnorm_tables.tables()); ctx.set_rng(42, 0)
ctx.set_problem(arr['data'], arr['var'], arr['fsf'], arr['lsf'], arr['pmin'], arr['pmax'], [0,.1,.1], arr['prior'], chains_per_cube=1)
ctx.init_params_uniform(); ctx.forward()
ctx.sweep(1, 1, mode=_native.COLOURED, refresh_every=0, min_acceptance_rate=0.0)
t0 = time.perf_counter()
_, _, ms = ctx.sweep(2, 2, mode=_native.COLOURED, refresh_every=0, min_acceptance_rate=0.0)
wall = time.perf_counter() - t0
print('launches', ctx.counters()['kernel_launches'], end=' ')
print('cluster=%s field %d: d3d_sweep coloured %.2f ms/sweep (device), wall %.2f ms/sweep, per phase %.1f us' % (os.environ.get('D3D_CLUSTER', 'auto'), field, ms / 2, wall * 500, ms / 2 / 1681 * 1e3))
