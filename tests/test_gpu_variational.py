"""Variational-dropout training path on the GPU: Variance_Model kernels against the reference fixture, the
log-likelihood variant of the fused training step, the KL-gradient kernel, and the graph-captured trainer against the
nn.Module + VariationalDropoutLoss + torch.optim.Adam recipe of training/training.py:80-84,116-137."""
import copy
import math

import numpy as np
import pytest
import torch

from tests.util import load, relerr

pytestmark = pytest.mark.gpu


def _volume(shape=(40, 36, 44)):
    xs = [torch.linspace(0, 1, s) for s in shape]
    v = torch.sin(6 * xs[0])[:, None, None] * torch.cos(4 * xs[1])[None, :, None] + 0.4 * torch.sin(9 * xs[2])[None, None, :]
    v = 2 * (v - v.min()) / (v.max() - v.min()) - 1
    return v.contiguous().cuda()


def _model(seed=3, C=8, drop='variational_dynamic'):
    from latent_feature_grid_compression_b200.model.model_utils import setup_model
    torch.manual_seed(seed)
    return setup_model(3, 32, 1, 4, 'fourier', 2, drop, 0.1, 0.9, 'db2', C, 15, '').cuda().train()


def test_variance_model_matches_reference_fixture():
    from latent_feature_grid_compression_b200.model.Variational_Dropout_Layer import Variance_Model
    g = load('variance_model')
    vm = Variance_Model()
    vm.load_state_dict({k[3:]: torch.from_numpy(v) for k, v in g.items() if k.startswith('sd.')})
    vm.cuda()
    x = torch.from_numpy(g['x']).cuda().requires_grad_(True)
    y = vm(x)
    assert tuple(y.shape) == tuple(g['y'].shape)
    assert relerr(y.detach().cpu().numpy(), g['y']) < 1e-5
    (y * torch.from_numpy(g['w']).cuda()).sum().backward()
    for name, prm in vm.named_parameters():
        ref = g['grad.' + name]
        assert float(np.abs(prm.grad.cpu().numpy() - ref).max()) <= 1e-5 * float(np.abs(ref).max()), name


@pytest.mark.parametrize('n', [1, 77, 5000, 40000])
def test_variance_model_against_torch_ops(n):
    """Ragged and multi-tile batches against the same MLP written with torch ops (fp64)."""
    from latent_feature_grid_compression_b200.model.Variational_Dropout_Layer import Variance_Model
    torch.manual_seed(n)
    vm = Variance_Model().cuda()
    x = (torch.rand(n, 3, device='cuda') * 2 - 1)
    w = torch.randn(n, 1, device='cuda') / n
    y = vm(x)
    (y * w).sum().backward()
    h = x.double()
    ps = [(l.weight.detach().double().requires_grad_(True), l.bias.detach().double().requires_grad_(True))
          for l in list(vm.net_layers) + [vm.final_layer]]
    for W, b in ps[:-1]:
        h = torch.relu(h @ W.t() + b)
    yr = h @ ps[-1][0].t() + ps[-1][1]
    (yr * w.double()).sum().backward()
    assert relerr(y.detach().cpu().numpy(), yr.detach().cpu().numpy()) < 1e-5
    for (W, b), layer in zip(ps, list(vm.net_layers) + [vm.final_layer]):
        for ref, got in ((W.grad, layer.weight.grad), (b.grad, layer.bias.grad)):
            assert float((got.double() - ref).abs().max()) <= 1e-5 * max(float(ref.abs().max()), 1e-30)


def test_variance_model_refuses_cpu_and_unsupported_shapes():
    from latent_feature_grid_compression_b200._lib import LfgcError
    from latent_feature_grid_compression_b200.model.Variational_Dropout_Layer import Variance_Model
    with pytest.raises(LfgcError):
        Variance_Model()(torch.zeros(4, 3))
    with pytest.raises(LfgcError):
        Variance_Model(size_layers=64)


def test_log_likelihood_train_step_matches_oracle_formulas():
    """lfgc_train_step_weighted: MLP / grid gradients equal the plain backward fed with the reference's
    d loss / d pred, and dlog_sigma equals d loss / d log_sigma (oracle.variational_loss)."""
    from latent_feature_grid_compression_b200 import ops
    from oracle import fvsrn_numpy as O
    m = _model(drop='')
    geom = m.geometry()
    grid = ops.decode_fwd(geom, [f.detach().contiguous() for f in m.feature_grid], [None] * len(m.feature_grid))
    mlp = m.mlp_flat()
    n = 3000
    torch.manual_seed(5)
    coords = torch.rand(n, 3, device='cuda') * 2 - 1
    gt = torch.rand(n, device='cuda') * 2 - 1
    v = torch.randn(n, device='cuda') * 0.3 - 0.5
    scale = 63360.0 / n
    gg = torch.zeros((*geom.G, geom.Cp), device='cuda')
    gm = torch.empty(geom.mlp_param_count, device='cuda')
    ls = torch.zeros(1, device='cuda')
    dv = torch.zeros(n, device='cuda')
    ws = torch.empty(geom.backward_workspace_bytes // 4, device='cuda')
    ops.train_step(geom, None, n, 0, 0, 0.5 * scale, grid, mlp, gg, gm, ls, ws, coords=coords, targets=gt, log_sigma=v,
                   dlog_sigma=dv)
    pred = ops.sample_forward(geom, coords, grid, mlp, clamp=False)
    a = 1.0 / (2.0 * np.exp(v.double().cpu().numpy()) ** 2)
    err = (gt.double() - pred.double()).cpu().numpy()
    g_pred = -scale * 2.0 * a * err
    g_v = -scale * (2.0 * a * err ** 2 - 1.0)
    assert relerr(dv.cpu().numpy(), g_v) < 1e-5
    assert abs(float(ls) - float((err ** 2).sum())) <= 1e-5 * float((err ** 2).sum())
    gg_ref, gm_ref = ops.sample_backward(geom, coords, torch.from_numpy(g_pred).float().cuda(), grid, mlp)
    assert float((gm - gm_ref).abs().max()) <= 1e-5 * float(gm_ref.abs().max())
    assert float((gg - gg_ref).abs().max()) <= 1e-5 * float(gg_ref.abs().max())


def test_dkl_gradient_kernel_matches_autograd_and_ramps_the_weight():
    from latent_feature_grid_compression_b200 import ops
    from latent_feature_grid_compression_b200.model.Variational_Dropout_Layer import VariationalDropout
    torch.manual_seed(2)
    sizes = [(6, 6, 6), (7, 6, 6, 6), (7, 9, 9, 9)]
    layers = [VariationalDropout(s, 0.1, 0.5).cuda() for s in sizes]
    for d in layers:
        d.log_thetas.data.normal_(0, 0.5)
        d.log_var.data.normal_(-2, 1.5)
    flat = torch.cat([torch.cat([d.log_thetas.detach().reshape(-1), d.log_var.detach().reshape(-1)]) for d in layers])
    grads = torch.randn_like(flat)
    g0 = grads.clone()
    w0, ramp, wmax, scale = 0.1, 1.0 + 3e-2, 0.104, 31.5
    w = torch.full((2,), w0, device='cuda', dtype=torch.float64)
    step = torch.zeros(2, device='cuda', dtype=torch.int32)
    ops.variational_dkl_grad(flat, grads, [d.log_thetas.numel() for d in layers], w, step, ramp, wmax, scale)
    total = sum(d.calculate_Dkl() for d in layers) * (w0 * ramp) * scale
    total.backward()
    ref = torch.cat([torch.cat([d.log_thetas.grad.reshape(-1), d.log_var.grad.reshape(-1)]) for d in layers])
    assert float(((grads - g0) - ref).abs().max()) <= 1e-5 * float(ref.abs().max())
    # ramp: w <- w * ramp while w < w_max, ping-pong slot of the step parity
    cur = w0
    for k in range(4):
        assert abs(float(w[k & 1]) - cur) < 1e-15
        cur = cur * ramp if cur < wmax else cur
        step[0] = k
        ops.variational_dkl_grad(flat, grads, [d.log_thetas.numel() for d in layers], w, step, ramp, wmax, scale)
        assert abs(float(w[(k + 1) & 1]) - cur) < 1e-15


@pytest.mark.parametrize('flavour', ['dynamic', 'static'])
def test_fast_trainer_variational_matches_module_path(flavour):
    from latent_feature_grid_compression_b200 import ops
    from latent_feature_grid_compression_b200.model.Variational_Dropout_Layer import VariationalDropoutLoss, Variance_Model
    from latent_feature_grid_compression_b200.training.fast_loop import FastTrainer
    vol = _volume()
    n_vox = vol.numel()
    n, steps, seed, lr = 2000, 8, 11, 0.008
    a, b = _model(3), _model(3)
    b.load_state_dict(copy.deepcopy(a.state_dict()))
    cfg = dict(n_voxels=n_vox, weight_dkl=0.1, weight_weights=2.0, weight_dkl_multiplier=3e-2)
    params_b = list(b.parameters())
    if flavour == 'dynamic':
        torch.manual_seed(9)
        va = Variance_Model().cuda()
        vb = Variance_Model().cuda()
        vb.load_state_dict(copy.deepcopy(va.state_dict()))
        cfg['variance_model'] = va
        params_b += list(vb.parameters())
    else:
        cfg['log_sigma'] = -0.7
    tr = FastTrainer(a, vol, n, lr=lr, seed=seed, use_graph=False, variational=cfg)
    tr.capture()                        # the warm-up steps draw noise too: keep them out of the seeded region
    opt = torch.optim.Adam(params_b, lr=lr)
    crit = VariationalDropoutLoss(n_vox, n, weight_dkl=0.1, weight_weights=2.0)
    mse_b = 0.0
    for s in range(steps):
        torch.manual_seed(100 + s)      # both paths draw the mask noise from the same generator state
        tr.step()
        raw, norm, gt = ops.sample(vol.shape, n, seed=seed, sample_offset=s * n, volume=vol, want_gt=True)
        torch.manual_seed(100 + s)
        opt.zero_grad()
        pred = b(norm).squeeze(-1)
        logsig = vb(norm).squeeze(-1) if flavour == 'dynamic' else torch.full_like(pred, -0.7)
        loss, ll, mse, dkl, wsum = crit(b, pred, gt, logsig, 3e-2)
        loss.backward()
        opt.step()
        mse_b = float(mse.detach())
    assert abs(tr.last_loss() - mse_b) <= 2e-4 * abs(mse_b)
    assert abs(float(tr.w_dkl[steps & 1]) - crit.weight_dkl) <= 1e-12 * crit.weight_dkl
    sa, sb = a.state_dict(), b.state_dict()
    for k in sb:
        ref = sb[k].float()
        assert float((sa[k].float() - ref).abs().max()) <= 2e-4 * max(float(ref.abs().max()), 1e-6), k
    if flavour == 'dynamic':
        for (k, pa), (_, pb) in zip(va.state_dict().items(), vb.state_dict().items()):
            assert float((pa - pb).abs().max()) <= 2e-4 * max(float(pb.abs().max()), 1e-6), k


def test_fast_trainer_variational_graph_replay_runs_the_whole_schedule():
    """Graph-captured variational step: fresh samples and noise per replay, the device-side KL-weight ramp, and the
    two-phase driver (mask baking + fine-tuning) end to end on a small volume."""
    from latent_feature_grid_compression_b200.model.Variational_Dropout_Layer import Variance_Model
    from latent_feature_grid_compression_b200.training.fast_loop import FastTrainer, train_volume
    vol = _volume()
    a = _model(5)
    torch.manual_seed(1)
    cfg = dict(n_voxels=vol.numel(), weight_dkl=0.1, weight_weights=2.0, weight_dkl_multiplier=1e-2,
               variance_model=Variance_Model().cuda())
    tr = FastTrainer(a, vol, 4096, lr=0.008, seed=3, variational=cfg)
    losses = []
    for s in range(60):
        tr.step()
        if s % 10 == 9:
            losses.append(tr.last_loss())
    assert all(math.isfinite(v) for v in losses) and losses[-1] < losses[0]
    assert int(tr.step_dev[0]) == 60
    assert abs(float(tr.w_dkl[0]) - 0.1 * (1.01 ** 60)) <= 1e-9
    args = dict(d_in=3, n_hidden_size=32, d_out=1, n_layers=4, embedding_type='fourier', n_embedding_freq=2,
                drop_type='variational_dynamic', drop_momentum=0.1, drop_threshold=0.9, wavelet_filter='db2',
                grid_features=8, grid_size=15, checkpoint_path='', batch_size=256, sample_size=16, max_pass=6,
                pass_decay=20, lr_decay=0.2, lr=0.008, lambda_drop_loss=0.1, lambda_weight_loss=2.0,
                variational_sigma=0, weight_dkl_multiplier=3e-5)
    info = train_volume(args, volume=vol.cpu(), seed=0)
    assert math.isfinite(info['psnr']) and info['psnr'] > 15.0
    assert info['compression_ratio'] > 0
