"""
Builds libdeconv3d_b200.so (the CUDA kernels + C ABI) in-tree for sm_100a.

    python -m deconv3d_b200.build_native [--force] [--verbose]

nvcc cross-compiles without a GPU; the resulting .so sits next to this file so
that it travels with the source tree (it is git-ignored, not gpurun-ignored).
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'libdeconv3d_b200.so')
SOURCES = [os.path.join(CSRC, 'd3d_api.cu')]
DEPS = SOURCES + sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.cuh', '.h'))) + \
    [os.path.join(os.path.dirname(HERE), 'include', 'deconv3d_b200.h')]


def find_nvcc():
    for cand in (os.environ.get('NVCC'), shutil.which('nvcc'), '/usr/local/cuda/bin/nvcc'):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError('nvcc not found; deconv3d_b200 needs the CUDA toolkit to build its kernels')


def up_to_date():
    if not os.path.exists(LIB):
        return False
    t = os.path.getmtime(LIB)
    return all(os.path.getmtime(d) <= t for d in DEPS)


def build(force=False, verbose=False):
    if not force and up_to_date():
        return LIB
    cmd = [find_nvcc(), '-shared', '-Xcompiler', '-fPIC', '-O3', '-std=c++17', '-lineinfo',
           '-gencode', 'arch=compute_100a,code=sm_100a',
           '-o', LIB] + SOURCES
    if verbose:
        cmd.insert(1, '-Xptxas=-v')
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError('nvcc failed:\n%s\n%s\n%s' % (' '.join(cmd), res.stdout, res.stderr))
    if verbose:
        sys.stderr.write(res.stderr)
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='--verbose' in sys.argv))
