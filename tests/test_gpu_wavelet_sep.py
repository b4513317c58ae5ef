"""The separable large-grid synthesis / adjoint kernels (csrc/wavelet_sep.cu) against the numpy oracle and against the
direct-sum kernels (csrc/wavelet.cu): small and odd extents forced through LFGC_WAVELET_SEP=1, and the wide grid of
BASELINE.json configs[4] (G = 64, four db2 levels) where they are the default.
Reference: model/Feature_Grid_Model.py:102-108, wavelet_transform/Torch_Wavelet_Transform.py:91-104."""
import numpy as np
import pytest
import torch

from oracle import fvsrn_numpy as O

pytestmark = pytest.mark.gpu


def _case(C, G, wavelet, seed):
    from latent_feature_grid_compression_b200 import ops
    rng = np.random.default_rng(seed)
    grid = rng.uniform(0, 1, size=(C, G, G, G))
    coeffs64, shapes = O.encode_volume(grid, wavelet)
    dims = [c.shape[-3:] for c in coeffs64]
    geom = ops.Geometry(C, (G, G, G), 32, 4, 2, wavelet, dims, np.asarray(shapes).reshape(-1, 3))
    coeffs = [torch.from_numpy(c.astype(np.float32)).cuda().contiguous() for c in coeffs64]
    return geom, grid, coeffs64, shapes, coeffs, rng


@pytest.mark.parametrize('C,G,wavelet', [(3, 15, 'db2'), (5, 17, 'db2'), (4, 16, 'haar'), (2, 33, 'db2'), (6, 12, 'db2')])
def test_separable_kernels_forced_on_small_grids(C, G, wavelet, monkeypatch):
    from latent_feature_grid_compression_b200 import ops
    geom, grid, coeffs64, shapes, coeffs, rng = _case(C, G, wavelet, C * 31 + G)
    if len(coeffs) < 2:
        pytest.skip('no wavelet level')
    n = len(coeffs)
    gg = torch.zeros((G, G, G, geom.Cp), device='cuda')
    gg[..., :C] = torch.from_numpy(rng.standard_normal((G, G, G, C)).astype(np.float32)).cuda()
    res = {}
    for flag in ('0', '1'):
        monkeypatch.setenv('LFGC_WAVELET_SEP', flag)
        zero = torch.ones((G, G, G, geom.Cp), device='cuda')
        out = ops.decode_fwd(geom, coeffs, [None] * n, also_zero=zero)
        gcs, _ = ops.decode_bwd(geom, gg, coeffs, [None] * n, [False] * n)
        torch.cuda.synchronize()
        assert float(zero.abs().max()) == 0.0
        res[flag] = (out.cpu().numpy(), [g.cpu().numpy() for g in gcs])
    ref = np.moveaxis(O.decode_volume([c.astype(np.float64) for c in coeffs64], [None] * n, shapes, wavelet), 0, -1)
    gref, _ = O.decode_volume_adjoint(np.moveaxis(gg[..., :C].cpu().numpy().astype(np.float64), -1, 0),
                                      [c.astype(np.float64) for c in coeffs64], [None] * n, shapes, wavelet)
    for flag in ('0', '1'):
        out, gcs = res[flag]
        assert np.abs(out[..., :C] - ref).max() <= 3e-6 * np.abs(ref).max(), flag
        assert not out[..., C:].any()
        for a, b in zip(gcs, gref):
            assert np.abs(a - b).max() <= 3e-6 * max(np.abs(b).max(), 1e-30), flag


def test_separable_kernels_are_the_default_on_the_wide_grid():
    """C = 8, G = 64 (four db2 levels, 2.1 M vertices): perfect reconstruction, oracle synthesis on a channel subset,
    adjointness <decode c, g> == <c, decode^T g>."""
    from latent_feature_grid_compression_b200 import _lib as L
    from latent_feature_grid_compression_b200 import ops
    import ctypes as ct
    C, G = 8, 64
    geom, grid, coeffs64, shapes, coeffs, rng = _case(C, G, 'db2', 5)
    n = len(coeffs)
    assert n == 5
    out = ops.decode_fwd(geom, coeffs, [None] * n).cpu().numpy()
    want = np.moveaxis(grid, 0, -1)
    assert np.abs(out[..., :C] - want).max() <= 1e-5      # encode (fp64 oracle) -> decode: the grid itself
    g = torch.randn((G, G, G, geom.Cp), device='cuda')
    g[..., C:] = 0
    gcs, _ = ops.decode_bwd(geom, g, coeffs, [None] * n, [False] * n)
    lhs = float((torch.from_numpy(out).cuda().double() * g.double()).sum())
    rhs = sum(float((a.double() * b.double()).sum()) for a, b in zip(coeffs, gcs))
    assert abs(lhs - rhs) <= 1e-5 * max(abs(lhs), 1.0)
    gref, _ = O.decode_volume_adjoint(np.moveaxis(g[..., :2].cpu().numpy().astype(np.float64), -1, 0),
                                      [c[:2].astype(np.float64) for c in coeffs64], [None] * n, shapes, 'db2')
    for a, b in zip(gcs, gref):
        assert np.abs(a[:2].cpu().numpy() - b).max() <= 3e-6 * np.abs(b).max()
