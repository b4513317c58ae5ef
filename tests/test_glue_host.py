"""The glue kernel's per-element functions (csrc/step_glue.cu: synthesis, its adjoint, Adam) run on HOST memory through a
test-only build (-DLFGC_GLUE_HOST_TEST, never part of liblfgc.so) and are checked against the numpy oracle: the index
arithmetic of the cooperative kernel is verified without a GPU.  (The GPU run of the same functions is
tests/test_gpu_trainer.py::test_glue_step_equals_separate_kernels.)"""
import ctypes as ct
import os
import shutil
import subprocess

import numpy as np
import pytest

from oracle import fvsrn_numpy as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, 'latent_feature_grid_compression_b200', 'csrc')


@pytest.fixture(scope='module')
def hostlib(tmp_path_factory):
    nvcc = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    if not os.path.exists(nvcc):
        pytest.skip('nvcc not available')
    out = str(tmp_path_factory.mktemp('glue') / 'libglue_host.so')
    cmd = [nvcc, '-shared', '-Xcompiler', '-fPIC', '-std=c++17', '-O2', '-gencode', 'arch=compute_100a,code=sm_100a',
           '-DLFGC_GLUE_HOST_TEST', '-I', os.path.join(ROOT, 'include'), '-I', CSRC,
           os.path.join(CSRC, 'step_glue.cu'), os.path.join(CSRC, 'api.cu'), '-o', out]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr[-2000:]
    return ct.CDLL(out)


def _ptr(a):
    return ct.c_void_p(a.ctypes.data)


def _ptr_array(arrs):
    return (ct.c_void_p * len(arrs))(*[a.ctypes.data for a in arrs])


def _run(lib, geom, coeffs, gcoeffs, scratch, grad_grid, grid_cl, also_zero, p, g, m, v, lr, step, phases):
    from latent_feature_grid_compression_b200 import _lib as L
    fn = lib.lfgc_step_glue_host
    fn.restype = ct.c_int
    fn.argtypes = [ct.POINTER(L.WaveletDesc), ct.c_int, ct.POINTER(ct.c_void_p), ct.POINTER(ct.c_void_p)] + \
                  [ct.c_void_p] * 8 + [ct.c_int64, ct.c_void_p, ct.c_void_p] + [ct.c_double] * 4 + [ct.c_int]
    rc = fn(ct.byref(geom.wavelet_desc), geom.Cp, _ptr_array(coeffs), _ptr_array(gcoeffs), _ptr(scratch),
            _ptr(grad_grid), _ptr(grid_cl), _ptr(also_zero), _ptr(p), _ptr(g), _ptr(m), _ptr(v), p.size, _ptr(lr),
            _ptr(step), 0.9, 0.999, 1e-8, 1.0, phases)
    assert rc == 0


@pytest.mark.parametrize('C,G,wavelet', [(3, 15, 'db2'), (2, 16, 'haar'), (5, 33, 'db2'), (2, 5, 'db2')])
def test_glue_phases_against_the_oracle(hostlib, C, G, wavelet):
    from latent_feature_grid_compression_b200 import ops
    rng = np.random.default_rng(C * 100 + G)
    grid = rng.uniform(0, 1, size=(C, G, G, G))
    coeffs64, shapes = O.encode_volume(grid, wavelet)
    n_coeff = len(coeffs64)
    dims = [c.shape[-3:] for c in coeffs64]
    geom = ops.Geometry(C, (G, G, G), 32, 4, 2, wavelet, dims, np.asarray(shapes).reshape(-1, 3))
    Cp = geom.Cp
    # flat parameter / gradient buffers holding the coefficient tensors back to back (as FastTrainer lays them out)
    sizes = [c.size for c in coeffs64]
    n = sum(sizes)
    p = np.concatenate([c.astype(np.float32).ravel() for c in coeffs64])
    g = np.zeros(n, np.float32)
    offs = np.cumsum([0] + sizes)
    coeffs = [p[offs[i]:offs[i + 1]].reshape(coeffs64[i].shape) for i in range(n_coeff)]
    gcoeffs = [g[offs[i]:offs[i + 1]].reshape(coeffs64[i].shape) for i in range(n_coeff)]
    scratch = np.zeros(max(geom.decode_scratch_bytes // 4, 4), np.float32)
    grid_cl = np.full((G, G, G, Cp), 7.0, np.float32)
    also_zero = np.full((G, G, G, Cp), 3.0, np.float32)
    grad_grid = np.zeros((G, G, G, Cp), np.float32)
    grad_grid[..., :C] = rng.standard_normal((G, G, G, C)).astype(np.float32)
    m = rng.standard_normal(n).astype(np.float32) * 0.01
    v = (rng.standard_normal(n).astype(np.float32) * 0.01) ** 2
    lr = np.array([0.008], np.float32)
    step = np.array([4, 0], np.int32)

    # synthesis (phase 4)
    _run(hostlib, geom, coeffs, gcoeffs, scratch, grad_grid, grid_cl, also_zero, p, g, m, v, lr, step, 4)
    ref = O.decode_volume([c.astype(np.float64) for c in coeffs], [None] * n_coeff, shapes, wavelet)   # (C,G,G,G)
    assert np.abs(grid_cl[..., :C] - np.moveaxis(ref, 0, -1)).max() <= 2e-6 * np.abs(ref).max()
    assert not grid_cl[..., C:].any() and not also_zero.any()

    # adjoint (phase 1)
    _run(hostlib, geom, coeffs, gcoeffs, scratch, grad_grid, grid_cl, also_zero, p, g, m, v, lr, step, 1)
    gref, _ = O.decode_volume_adjoint(np.moveaxis(grad_grid[..., :C].astype(np.float64), -1, 0),
                                      [c.astype(np.float64) for c in coeffs], [None] * n_coeff, shapes, wavelet)
    for a, b in zip(gcoeffs, gref):
        assert np.abs(a - b).max() <= 2e-6 * max(np.abs(b).max(), 1e-30)

    # Adam (phase 2) on the gradients just produced; the step counter advances
    p0, m0, v0 = p.astype(np.float64), m.astype(np.float64), v.astype(np.float64)
    _run(hostlib, geom, coeffs, gcoeffs, scratch, grad_grid, grid_cl, also_zero, p, g, m, v, lr, step, 2)
    pr, mr, vr = O.adam_step(p0, g.astype(np.float64), m0, v0, 5, float(lr[0]))
    assert int(step[0]) == 5
    assert np.abs(p - pr).max() <= 1e-5 * np.abs(pr).max()
    # the betas cross the C ABI as doubles, so 1 - beta2 is fl32(0.001) as in torch (1 - fl32(0.999) would be 4.7e-5 off)
    assert np.abs(m - mr).max() <= 1e-6 * np.abs(mr).max() and np.abs(v - vr).max() <= 1e-6 * np.abs(vr).max()

    # all phases in one call == the three calls in sequence
    p2 = np.concatenate([c.astype(np.float32).ravel() for c in coeffs64])
    g2 = np.zeros(n, np.float32)
    c2 = [p2[offs[i]:offs[i + 1]].reshape(coeffs64[i].shape) for i in range(n_coeff)]
    gc2 = [g2[offs[i]:offs[i + 1]].reshape(coeffs64[i].shape) for i in range(n_coeff)]
    m2, v2 = m0.astype(np.float32), v0.astype(np.float32)
    step2 = np.array([4, 0], np.int32)
    grid2 = np.zeros_like(grid_cl)
    zero2 = np.ones_like(grid_cl)
    _run(hostlib, geom, c2, gc2, scratch, grad_grid, grid2, zero2, p2, g2, m2, v2, lr, step2, 7)
    assert np.array_equal(p2, p) and np.array_equal(g2, g) and int(step2[0]) == 5
    ref2 = O.decode_volume([c.astype(np.float64) for c in c2], [None] * n_coeff, shapes, wavelet)
    assert np.abs(grid2[..., :C] - np.moveaxis(ref2, 0, -1)).max() <= 2e-6 * np.abs(ref2).max()
    assert not zero2.any()
