import sys, os, ctypes, numpy as np
sys.path.insert(0, '/root/repo')
from deconv3d_b200 import _native
_native.LIB_PATH = os.environ.get('D3D_TRACE_LIB', '/root/repo/scratch/libd3d_trace.so')   # nvcc ... -DD3D_TRACE -o that.so deconv3d_b200/csrc/d3d_api.cu
import bench
chains = int(sys.argv[1]) if len(sys.argv) > 1 else 1
wl = bench.build_workload('cfg2x256', chains); arr = bench.realise(wl, 0)
from deconv3d_b200 import rtnorm_tables
ctx = _native.Context(0); ctx.set_rtnorm_tables(*rtnorm_tables.tables()); ctx.set_rng(42, 0)
ctx.set_problem(arr['data'], arr['var'], arr['fsf'], arr['lsf'], arr['pmin'], arr['pmax'], [0,.1,.1], arr['prior'], chains_per_cube=chains)
ctx.init_params_uniform(); ctx.forward()
ctx.sweep(1, 60, refresh_every=0, min_acceptance_rate=0.0)
_, _, ms = ctx.sweep(61, 10, refresh_every=0, min_acceptance_rate=0.0)
print('us/site %.3f = %.0f cycles' % (ms * 1e3 / 16000, ms * 1e3 / 16000 * 1965))
lib = _native.load()
out = (ctypes.c_ulonglong * (16 * 16 * 16))()
lib.d3d_debug_trace2(out)
T = np.array(out, dtype=np.int64).reshape(16, 16, 16)   # site, warp, ev
nww = 9
names = {0: 'start', 1: 'sums done', 3: 'partials posted', 6: 'switch done', 4: 'decision seen', 5: 'update done'}
base = T[3, 0, 0]
for j in range(3, 7):
    print('--- site', 700 + j, '(x = %d)' % ((700 + j) % 40))
    for ev in (0, 1, 3, 6, 4, 5):
        row = T[j, :nww, ev] - base
        print('  W %-16s' % names[ev], ' '.join('%6d' % v for v in row), '  max %d' % row.max())
    b = T[j, nww + 2]
    print('  B totals read %6d | props read %6d | decision written %6d' % (b[8] - base, b[10] - base, b[9] - base))
    a = T[j, nww]; p = T[j, nww + 1]
    print('  A start %6d done %6d | P start %6d done %6d   (site they work on = this j)' % (a[12] - base, a[13] - base, p[12] - base, p[13] - base))
per = (T[4:14, 0, 0] - T[3:13, 0, 0])
print('site periods', per)
W = T[3:14]
print('mean over sites: sums %.0f | post %.0f | wait decision %.0f | update %.0f | next start %.0f' % (
    (W[:, :nww, 1].max(1) - W[:, :nww, 0].min(1)).mean(), (W[:, :nww, 3].max(1) - W[:, :nww, 1].max(1)).mean(),
    (W[:, :nww, 4].min(1) - W[:, :nww, 3].max(1)).mean(), (W[:, :nww, 5].max(1) - W[:, :nww, 4].min(1)).mean(),
    (W[1:, :nww, 0].min(1) - W[:-1, :nww, 5].max(1)).mean()))
Bw = T[3:14, nww + 2]
print('B: partials posted(max W) -> totals read %.0f | totals -> decision written %.0f | decision written -> W sees %.0f' % (
    (Bw[:, 8] - W[:, :nww, 3].max(1)).mean(), (Bw[:, 9] - Bw[:, 8]).mean(), (W[:, :nww, 4].min(1) - Bw[:, 9]).mean()))
