"""
CPU-only tests of the host-side mirror of the reference interface and of the
C-ABI library's presence (no compute calls: there is no GPU here).
"""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import load_golden, ROOT


def test_library_builds_loads_and_exports_every_declared_symbol():
    from deconv3d_b200 import build_native, _native
    lib_path = build_native.build()
    assert os.path.exists(lib_path)
    lib = ctypes.CDLL(lib_path)
    header = open(os.path.join(ROOT, 'include', 'deconv3d_b200.h')).read()
    declared = set(re.findall(r'\b(d3d_[a-z0-9_]+)\s*\(', header))
    assert declared, 'no declarations found in the header'
    assert declared == set(_native.SYMBOLS)
    assert int(re.search(r'#define D3D_ABI_VERSION (\d+)', header).group(1)) == _native.ABI_VERSION
    for name in declared:
        assert hasattr(lib, name), 'library does not export %s' % name
    assert _native.load().d3d_abi_version() == _native.ABI_VERSION


def test_no_silent_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip('a GPU is present')
    from deconv3d_b200 import _native, Run, MUSE
    with pytest.raises(_native.NativeError):
        _native.Context(0)
    data = load_golden('ref_run_A')['data']
    with pytest.raises(_native.NativeError):
        Run(MUSE().build_cube(data), MUSE(fsf_fwhm=0.5), max_iterations=3)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, 'deconv3d_b200')
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                text = open(os.path.join(base, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle\b', text, re.M), f
                assert 'oracle.' not in text.replace('oracle/', ''), f


def test_spread_functions_match_reference_golden():
    from deconv3d_b200 import (MUSE, GaussianFieldSpreadFunction, MoffatFieldSpreadFunction,
                               GaussianLineSpreadFunction)
    g = load_golden('ref_spread')
    cube = MUSE().build_cube(np.zeros((40, 41, 39)))
    assert np.array_equal(GaussianFieldSpreadFunction(1.0).as_image(cube), g['gauss_default'])
    assert np.array_equal(GaussianFieldSpreadFunction(0.8).as_image(cube), g['gauss_08'])
    assert np.array_equal(GaussianFieldSpreadFunction(0.9, pa=25., ba=0.6).as_image(cube),
                          g['gauss_ell'])
    assert np.array_equal(MoffatFieldSpreadFunction(fwhm=0.8, beta=2.5, pa=0., ba=1.0)
                          .as_image(cube), g['moffat_41x39'])
    assert np.array_equal(MoffatFieldSpreadFunction(alpha=0.5, beta=3.0, pa=10., ba=0.8)
                          .as_image(cube), g['moffat_alpha'])
    assert np.array_equal(GaussianLineSpreadFunction(0.0002675).as_vector(cube), g['lsf_40'])
    assert np.array_equal(GaussianLineSpreadFunction(0.0002675).as_vector(
        MUSE().build_cube(np.zeros((30, 3, 3)))), g['lsf_30'])
    assert np.array_equal(GaussianLineSpreadFunction(0.0).as_vector(
        MUSE().build_cube(np.zeros((21, 3, 3)))), g['lsf_21_delta'])
    stamp = MoffatFieldSpreadFunction(fwhm=0.8, beta=2.5, size=13).as_image(cube)
    assert stamp.shape == (13, 13) and abs(stamp.sum() - 1) < 1e-14
    assert np.unravel_index(stamp.argmax(), stamp.shape) == (6, 6)
    inst = MUSE()
    assert np.array_equal(inst.fsf.as_image(cube), g['gauss_default'])


def test_instrument_type_checks():
    from deconv3d_b200 import Instrument, GaussianLineSpreadFunction, GaussianFieldSpreadFunction
    with pytest.raises(ValueError):
        Instrument(lsf=None, fsf=GaussianFieldSpreadFunction(1.))
    with pytest.raises(ValueError):
        Instrument(lsf=GaussianLineSpreadFunction(1e-4), fsf='x')


def test_rtnorm_tables_match_reference_samples():
    from deconv3d_b200 import rtnorm_tables
    g = load_golden('ref_rtnorm_tables')
    x, yu, ncell = rtnorm_tables.tables()
    assert x.shape == (4002,) and yu.shape == (4001,) and ncell.shape == (8961,)
    np.testing.assert_allclose(x[g['x_idx']], g['x_val'], rtol=0, atol=1e-11)
    np.testing.assert_allclose(yu[g['yu_idx']], g['yu_val'], rtol=2e-11, atol=0)
    assert np.array_equal(ncell[g['ncell_idx']], g['ncell_val'])
    crc = int(np.sum(ncell.astype(np.int64) * (np.arange(len(ncell)) % 251 + 1)))
    assert crc == int(g['ncell_crc'][0])
    assert np.all(np.diff(x) > 0) and np.all(np.diff(ncell) >= 0)


def test_masks_math_utils_line_model_vs_oracle():
    from deconv3d_b200 import above_percentile, MUSE, SingleGaussianLineModel
    from deconv3d_b200.math_utils import median_clip
    from oracle import reference_port as port
    data = load_golden('muse_cube_01')['data'] * 1e20
    cube = MUSE().build_cube(data)
    assert np.array_equal(above_percentile(cube), port.above_percentile(data))
    assert np.array_equal(above_percentile(cube, 60), port.above_percentile(data, 60))
    sub = data[2:-2, 2:-4, 2:4]
    assert median_clip(sub, 2.5) == port.median_clip(sub, 2.5)
    m = SingleGaussianLineModel()
    p = [3.2, 11.4, 1.7]
    assert np.array_equal(m.modelize(None, range(0, 30), p), port.modelize(30, p))

    class R(object):
        pass
    r = R()
    r.cube, r.fsf = cube, MUSE().fsf.as_image(cube)
    assert m.max_boundaries(r) == port.single_gaussian_boundaries(data, r.fsf)[1]
    assert m.min_boundaries(r) == [0, 0, 0] and m.gibbs_parameter_index() == 0


def test_padding_matches_oracle():
    from deconv3d_b200 import padding
    from oracle import reference_port as port
    rs = np.random.RandomState(0)
    for shape, axes in [((30,), [0]), ((5, 21), [1]), ((6, 7, 3), None), ((33, 4), [0])]:
        a = rs.rand(*shape)
        p1, s1 = padding(a, axes)
        p2, s2 = port.padding(a, axes)
        assert np.array_equal(p1, p2) and list(s1) == list(s2)


def test_fits_round_trip(tmp_path):
    from deconv3d_b200 import Cube, MUSE
    data = np.random.RandomState(1).rand(5, 4, 3)
    cube = MUSE().build_cube(data)
    path = str(tmp_path / 'c.fits')
    cube.to_fits(path)
    back = Cube.from_fits(path)
    assert np.array_equal(back.data, data)
    assert back.get_step(1).to('arcsec').value == cube.get_step(1).to('arcsec').value
    assert abs(cube.get_step(1).to('arcsec').value - 0.2) < 1e-9
    assert abs(cube.get_step(0).to('um').value - 1.25e-4) < 1e-18
    with pytest.raises(IOError):
        cube.to_fits(path)


def test_tied_gaussians_line_model_vs_oracle():
    """TiedGaussiansLineModel (the natively evaluated family of user line models,
    lib/line_models.py:4-61) against the oracle's multiplet, and the interface it keeps."""
    from deconv3d_b200 import TiedGaussiansLineModel, SingleGaussianLineModel, LineModel
    from oracle import reference_port as port
    m = TiedGaussiansLineModel([10.0, 13.2, 5.0], [2.0, 1.2, 0.5])
    assert isinstance(m, LineModel) and m.parameters() == ['a', 'c', 'w'] and m.gibbs_parameter_index() == 0
    off, rat = m.native_components()
    assert np.allclose(off, [0, 3.2, -5.0]) and np.allclose(rat, [1, 0.6, 0.25])
    p = [3.5, 14.2, 1.7]
    with port.line_model([10.0, 13.2, 5.0], [2.0, 1.2, 0.5]):
        assert np.allclose(m.modelize(None, range(0, 30), p), port.modelize(30, p), rtol=1e-15, atol=0)
    # one component = the single Gaussian, bit for bit
    one = TiedGaussiansLineModel([0.0], [1.0])
    assert np.array_equal(one.modelize(None, range(0, 30), p), SingleGaussianLineModel().modelize(None, range(0, 30), p))
    import pytest
    with pytest.raises(ValueError):
        TiedGaussiansLineModel([0, 1], [0.0, 1.0])
    with pytest.raises(ValueError):
        TiedGaussiansLineModel([0, 1, 2, 3, 4], [1, 1, 1, 1, 1])


def test_chain_row_arrays_fall_back_to_ordinary_memory_without_a_gpu():
    """`Run` asks for page-locked chain arrays (DMA at PCIe speed); where pinning is impossible
    (no CUDA device, D3D_NO_PINNED_ROWS) it gets ordinary zeroed memory of the same shape."""
    import os
    from deconv3d_b200.run import _host_rows
    for zeroed in (True, False):
        a = _host_rows((2, 70, 20, 20, 3), zeroed)              # 1.3 MB: above the pinning threshold
        assert a.shape == (2, 70, 20, 20, 3) and a.dtype == np.float64 and a.flags['C_CONTIGUOUS']
        a[...] = 1.0                                            # writable
    os.environ['D3D_NO_PINNED_ROWS'] = '1'
    try:
        assert not _host_rows((2, 70, 20, 20, 3), False).any()  # ordinary memory is always zeroed
    finally:
        del os.environ['D3D_NO_PINNED_ROWS']
    assert not _host_rows((3, 4), True).any()                   # small arrays: never pinned
