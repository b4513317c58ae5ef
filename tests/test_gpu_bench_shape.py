"""Oracle parity AT THE BENCHMARKED SHAPES (round-1 verdict, weak #1): the fused training step with the in-kernel
Philox sampler on a 255^3 volume at N = 32 768 (C16/G15/H32/L4: 256 tiles on 148 CTAs, i.e. the second-tile path of the
tcgen05 kernel with dW accumulating in TMEM across tiles) and at N = 262 144 on the wide grid (C32/G64, four wavelet
levels), both kernel families, against ``oracle.model_forward / model_backward``; and ``lfgc_reconstruct`` on the full
255^3 volume against the oracle on a 1 % voxel subset.

Gates (the usual ones): forward / loss 1e-5 relative, every gradient tensor max|d| <= 1e-5 * max|g|.
Reference: model/Feature_Grid_Model.py:50-80, training/training.py:89-138, visualization/OutputToVTK.py:7-47.
"""
import numpy as np
import pytest
import torch

from oracle import fvsrn_numpy as O
from tests.util import relerr

pytestmark = pytest.mark.gpu

GRAD_TOL = 1e-5


def _volume(R, seed=5):
    g = torch.Generator(device='cuda').manual_seed(seed)
    return (torch.rand(R, R, R, device='cuda', generator=g) * 2 - 1).contiguous()


def _model(C, G, seed):
    from latent_feature_grid_compression_b200.model.model_utils import setup_model
    torch.manual_seed(seed)
    m = setup_model(3, 32, 1, 4, 'fourier', 2, '', 0.1, 0.9, 'db2', C, G, '').cuda().train()
    with torch.no_grad():                      # non-zero biases: nn.Linear's default init is tiny at width 32
        for lyr in m.net_layers:
            lyr.bias.uniform_(-0.3, 0.3)
    return m


@pytest.mark.parametrize('family', ['tc', 'ffma2'])
@pytest.mark.parametrize('C,G,n', [(16, 15, 32768), (32, 64, 262144)])
def test_fused_train_step_at_bench_shape_vs_oracle(C, G, n, family, monkeypatch):
    from latent_feature_grid_compression_b200 import ops
    monkeypatch.setenv('LFGC_BACKWARD_TC', '1' if family == 'tc' else '0')
    monkeypatch.setenv('LFGC_FORWARD_TC', '1' if family == 'tc' else '0')
    R = 255
    vol = _volume(R)
    model = _model(C, G, 100 + C)
    geom = model.geometry()
    coeffs = [f.detach().contiguous() for f in model.feature_grid]
    grid_cl = ops.decode_fwd(geom, coeffs, [None] * len(coeffs))
    mlp = model.mlp_flat()
    ws = torch.empty(geom.backward_workspace_bytes // 4, device='cuda')
    gg = torch.zeros_like(grid_cl)
    gm = torch.empty(geom.mlp_param_count, device='cuda')
    loss = torch.zeros(1, device='cuda')
    seed, off = 1234, 3 * n
    ops.train_step(geom, vol, n, seed, off, 1.0 / n, grid_cl, mlp, gg, gm, loss, ws)
    torch.cuda.synchronize()
    # the samples the fused kernel drew: same Philox stream through the stand-alone sampler (bit-exact vs the reference's
    # coordinate formula in test_gpu_parity.py)
    raw, norm, gt = ops.sample(vol.shape, n, seed=seed, sample_offset=off, volume=vol, want_gt=True)
    sd = {k: v.detach().cpu().numpy() for k, v in model.state_dict().items()}
    spec = O.Spec(C, G, 32, 4, 2, 'db2', '')
    y, ctx = O.model_forward(sd, spec, norm.cpu().numpy(), training=True, keep=True)
    gt64 = gt.cpu().numpy().astype(np.float64)[:, None]
    want_loss = float(((y - gt64) ** 2).sum())
    assert abs(float(loss) - want_loss) <= 1e-5 * want_loss
    grads = O.model_backward((2.0 / n) * (y - gt64), ctx, spec)
    want_grid = np.transpose(grads['grid'], (1, 2, 3, 0))
    got_grid = gg.cpu().numpy()
    assert float(np.abs(got_grid[..., :C] - want_grid).max()) <= GRAD_TOL * np.abs(want_grid).max()
    if geom.Cp > C:
        assert float(np.abs(got_grid[..., C:]).max()) == 0.0
    o = 0
    for name, shape in geom.mlp_shapes():
        k = int(np.prod(shape))
        ref = grads[name].reshape(-1)
        d = float(np.abs(gm[o:o + k].cpu().numpy() - ref).max())
        assert d <= GRAD_TOL * max(float(np.abs(ref).max()), 1e-30), (name, d)
        o += k
    # and the forward kernel of the same family on the same positions
    pred = ops.sample_forward(geom, norm, grid_cl, mlp)
    assert relerr(pred.cpu().numpy(), y[:, 0]) < 1e-5


@pytest.mark.parametrize('family', ['tc', 'ffma2'])
def test_full_volume_reconstruction_vs_oracle_subset(family, monkeypatch):
    """field_from_net on all of 255^3 in one launch; 1 % of the voxels (plus the eight corners and a face) against the
    oracle evaluated at the coordinates the reference's tile loop would have fed the network."""
    from latent_feature_grid_compression_b200.data.IndexDataset import IndexDataset
    from latent_feature_grid_compression_b200.visualization.OutputToVTK import axis_tables, field_from_net
    monkeypatch.setenv('LFGC_FORWARD_TC', '1' if family == 'tc' else '0')
    R = 255
    model = _model(16, 15, 7).eval()
    ds = IndexDataset(torch.zeros(1, 1, 1).expand(R, R, R), 16)
    with torch.no_grad():
        full = field_from_net(ds, model, True, 32, to_cpu=False)
    assert tuple(full.shape) == (R, R, R)
    rng = np.random.default_rng(3)
    idx = rng.integers(0, R, size=(R ** 3 // 100, 3))
    corners = np.array([[a, b, c] for a in (0, R - 1) for b in (0, R - 1) for c in (0, R - 1)])
    face = np.stack([np.zeros(R, np.int64), np.arange(R), np.full(R, R - 1)], 1)
    idx = np.concatenate([idx, corners, face], 0)
    tabs = [t.cpu().numpy() for t in axis_tables(ds, 32)]
    coords = np.stack([tabs[a][idx[:, a]] for a in range(3)], 1).astype(np.float32)
    sd = {k: v.detach().cpu().numpy() for k, v in model.state_dict().items()}
    spec = O.Spec(16, 15, 32, 4, 2, 'db2', '')
    y = O.model_forward(sd, spec, coords, training=False)
    y = np.clip(np.asarray(y).reshape(-1), -1.0, 1.0)
    got = full[torch.from_numpy(idx[:, 0]).cuda(), torch.from_numpy(idx[:, 1]).cuda(),
               torch.from_numpy(idx[:, 2]).cuda()].cpu().numpy()
    assert float(np.abs(got - y).max()) <= 1e-5 * max(float(np.abs(y).max()), 1e-30)
