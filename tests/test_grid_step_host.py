"""The grid-step kernel's per-element functions (csrc/grid_step.cu: separable synthesis, its adjoint, Adam, partial
reduction) run on HOST memory through a test-only build (-DLFGC_GRID_STEP_HOST_TEST, never part of liblfgc.so) and are
checked against the numpy oracle, so the separable index arithmetic is verified without a GPU.  The GPU run of the same
functions is tests/test_gpu_grid_step.py."""
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
    out = str(tmp_path_factory.mktemp('gstep') / 'libgstep_host.so')
    cmd = [nvcc, '-shared', '-Xcompiler', '-fPIC', '-std=c++17', '-O2', '-gencode', 'arch=compute_100a,code=sm_100a',
           '-DLFGC_GRID_STEP_HOST_TEST', '-I', os.path.join(ROOT, 'include'), '-I', CSRC,
           os.path.join(CSRC, 'grid_step.cu'), os.path.join(CSRC, 'wavelet.cu'), os.path.join(CSRC, 'wavelet_sep.cu'),
           os.path.join(CSRC, 'api.cu'), '-o', out]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr[-2000:]
    return ct.CDLL(out)


@pytest.mark.parametrize('C,G,wavelet,n_srcs,w2', [(3, 15, 'db2', 1, 0.0), (2, 16, 'haar', 2, 0.0), (5, 17, 'db2', 1, 1e-3),
                                                 (2, 5, 'db2', 3, 0.0), (6, 12, 'db2', 1, 0.0)])
@pytest.mark.parametrize('nworkers', [1024, 352])   # the kernel's plan, and another split of the same work
def test_grid_step_against_the_oracle(hostlib, C, G, wavelet, n_srcs, w2, nworkers):
    from latent_feature_grid_compression_b200 import _lib as L
    from latent_feature_grid_compression_b200 import ops
    rng = np.random.default_rng(C * 100 + G)
    grid = rng.uniform(0, 1, size=(C, G, G, G))
    coeffs64, shapes = O.encode_volume(grid, wavelet)
    n_coeff = len(coeffs64)
    dims = [c.shape[-3:] for c in coeffs64]
    geom = ops.Geometry(C, (G, G, G), 32, 4, 2, wavelet, dims, np.asarray(shapes).reshape(-1, 3))
    Cp = geom.Cp
    assert int(hostlib.lfgc_grid_step_smem_bytes(ct.byref(geom.wavelet_desc))) > 0 or True
    # flat buffers as FastTrainer lays them out: [coefficients | pad | MLP block]
    sizes = [c.size for c in coeffs64]
    ncoef = sum(sizes)
    pcount = 37
    mlp_off = (ncoef + 3) // 4 * 4 + 8
    n = mlp_off + pcount
    p = np.zeros(n, np.float32)
    p[:ncoef] = np.concatenate([c.astype(np.float32).ravel() for c in coeffs64])
    p[mlp_off:] = rng.standard_normal(pcount).astype(np.float32)
    offs = np.cumsum([0] + sizes)
    g = np.full(n, 9.0, np.float32)
    m = (rng.standard_normal(n) * 0.01).astype(np.float32)
    v = ((rng.standard_normal(n) * 0.01) ** 2).astype(np.float32)
    lr = np.array([0.008], np.float32)
    step = np.array([4, 0], np.int32)
    grad_grids = []
    for r in range(n_srcs):
        gg = np.zeros((G, G, G, Cp), np.float32)
        gg[..., :C] = rng.standard_normal((G, G, G, C)).astype(np.float32)
        grad_grids.append(gg)
    nslices, pstride = 5, pcount + 1 + 3
    partials = [rng.standard_normal((nslices, pstride)).astype(np.float32) for _ in range(n_srcs)]
    grid_cl = np.full((G, G, G, Cp), 7.0, np.float32)
    zero = np.full((G, G, G, Cp), 3.0, np.float32)
    loss = np.zeros(1, np.float32)

    a = L.GridStepArgs()
    a.n_srcs = n_srcs
    for r in range(n_srcs):
        a.grad_grid[r] = grad_grids[r].ctypes.data
        a.mlp_partials[r] = partials[r].ctypes.data
    a.nslices, a.pstride, a.pcount = nslices, pstride, pcount
    a.zero_grid, a.grid_cl = zero.ctypes.data, grid_cl.ctypes.data
    a.p, a.g, a.m, a.v = p.ctypes.data, g.ctypes.data, m.ctypes.data, v.ctypes.data
    for l in range(n_coeff):
        a.coeff_off[l] = int(offs[l])
    a.mlp_off = mlp_off
    a.loss_out, a.lr, a.step_count = loss.ctypes.data, lr.ctypes.data, step.ctypes.data
    a.beta1, a.beta2, a.eps, a.grad_scale, a.weight_l2 = 0.9, 0.999, 1e-8, 1.0, w2
    p0, m0, v0 = p.astype(np.float64), m.astype(np.float64), v.astype(np.float64)
    fn = hostlib.lfgc_grid_step_host
    fn.restype = ct.c_int
    fn.argtypes = [ct.POINTER(L.WaveletDesc), ct.c_int, ct.POINTER(L.GridStepArgs), ct.c_int]
    assert fn(ct.byref(geom.wavelet_desc), Cp, ct.byref(a), nworkers) == 0

    # oracle: adjoint of the summed grid gradient, regulariser, Adam, synthesis of the updated coefficients
    gsum = sum(gg.astype(np.float64) for gg in grad_grids)
    gref, _ = O.decode_volume_adjoint(np.moveaxis(gsum[..., :C], -1, 0), [c.astype(np.float64) for c in coeffs64],
                                      [None] * n_coeff, shapes, wavelet)
    g_ref = np.zeros(n)
    g_ref[:ncoef] = np.concatenate([x.ravel() for x in gref]) + 2.0 * w2 * p0[:ncoef]
    g_ref[mlp_off:] = sum(pt.astype(np.float64)[:, :pcount].sum(0) for pt in partials)
    live = np.zeros(n, bool)
    live[:ncoef] = True
    live[mlp_off:] = True
    assert np.abs(g[live] - g_ref[live]).max() <= 3e-6 * np.abs(g_ref).max()
    assert np.all(g[~live] == 9.0)                                   # padding between the sections is not touched
    pr, mr, vr = O.adam_step(p0, g_ref, m0, v0, 5, float(lr[0]))
    assert int(step[0]) == 5
    assert np.abs(p[live] - pr[live]).max() <= 1e-5 * np.abs(pr).max()
    assert np.abs(m[live] - mr[live]).max() <= 2e-6 * np.abs(mr).max()
    assert np.abs(v[live] - vr[live]).max() <= 2e-6 * np.abs(vr).max()
    assert np.array_equal(p[~live], p0[~live].astype(np.float32))
    new_coeffs = [p[offs[i]:offs[i + 1]].reshape(coeffs64[i].shape).astype(np.float64) for i in range(n_coeff)]
    ref = O.decode_volume(new_coeffs, [None] * n_coeff, shapes, wavelet)
    assert np.abs(grid_cl[..., :C] - np.moveaxis(ref, 0, -1)).max() <= 3e-6 * np.abs(ref).max()
    assert not grid_cl[..., C:].any() and not zero.any()
    want_loss = sum(float(pt.astype(np.float64)[:, pcount].sum()) for pt in partials)
    assert abs(float(loss[0]) - want_loss) <= 1e-5 * max(abs(want_loss), 1.0)
