#!/usr/bin/env python
"""cfg5 convolution micro-benchmark: the full forward model of a survey batch
(1024 galaxies of 32^3, f64: spectral LSF pass + spatial FSF pass + residual) for FSF sizes
{3..41}^2 and LSF widths {0.5, 0.9, 1.5, 3} px.  CUDA events around d3d_forward, inputs resident.
usage (GPU box): python profiles/tools/conv_microbench.py > profiles/rNN_conv_microbench.txt"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch                                                    # noqa: E402
from deconv3d_b200 import _native, MUSE                         # noqa: E402
from deconv3d_b200.spread_functions import MoffatFieldSpreadFunction   # noqa: E402


def run(n, D, H, W, fs, lsf_sigma_px):
    inst = MUSE(fsf=MoffatFieldSpreadFunction(fwhm=0.8, beta=2.5, size=fs),
                lsf_fwhm=lsf_sigma_px * 2.35482 * 1.25e-4)
    cube0 = MUSE().build_cube(np.zeros((D, H, W)))
    fsf = np.asarray(inst.fsf.as_image(cube0))
    lsf = inst.lsf.as_vector(cube0)
    rs = np.random.RandomState(0)
    data = rs.rand(n, D, H, W)
    ctx = _native.Context(0, _native.F64)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    ctx.set_problem(data, np.full(data.shape, 0.01), fsf, lsf, np.zeros((n, 3)),
                    np.tile([100., D - 1, D], (n, 1)), [0, .1, .1], np.ones(n))
    p = np.dstack([rs.rand(H, W) * 9, rs.rand(H, W) * (D - 1), 0.5 + rs.rand(H, W) * 3])
    ctx.set_params(np.broadcast_to(p, (n, H, W, 3)).copy())
    for _ in range(3):
        ctx.forward(write_err=True)
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    reps = 10
    e0.record()
    for _ in range(reps):
        ctx.forward(write_err=True)
    e1.record()
    e1.synchronize()
    ms = e0.elapsed_time(e1) / reps
    vox = float(n) * D * H * W
    flop = 2.0 * vox * fs * fs
    byts = vox * 8 * 4                       # lines written + read, data read, residual written
    print('| %2dx%-2d | %.1f | %7.3f | %6.2f | %6.0f | %s |' % (
        fs, fs, lsf_sigma_px, ms, flop / ms / 1e9, byts / ms / 1e6,
        'wide (chunked)' if os.environ.get('D3D_STENCIL_WIDE') or fs not in (3, 5, 7, 9, 11, 13, 15, 17)
        else 'tiled<FW>'))
    ctx.close()


if __name__ == '__main__':
    if len(sys.argv) == 3:                   # one configuration (for ncu): FSF size, LSF sigma in px
        run(1024, 32, 32, 32, int(sys.argv[1]), float(sys.argv[2]))
        sys.exit(0)
    print('forward model, 1024 cubes of 32x32x32, f64, one B200; B_fwd counted as 4 x 8 bytes per voxel')
    print('| FSF | LSF sigma px | ms | TFLOP/s fp64 (FSF pass) | GB/s | spatial kernel |')
    print('|---|---|---|---|---|---|')
    for fs in (3, 5, 7, 9, 11, 13, 17, 21, 31, 41):
        run(1024, 32, 32, 32, fs, 0.9)
    for sg in (0.5, 1.5, 3.0):
        run(1024, 32, 32, 32, 13, sg)
