"""The graph-captured flat-buffer trainer against the nn.Module + torch.optim.Adam path on the same sample stream,
the two-phase training driver, and the on-disk format round trip."""
import copy

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _make(drop, seed=0, C=8):
    from latent_feature_grid_compression_b200.model.model_utils import setup_model
    torch.manual_seed(seed)
    m = setup_model(3, 32, 1, 4, 'fourier', 2, drop, 0.1, 0.9, 'db2', C, 15, '')
    return m.cuda().train()


def _volume(shape=(40, 36, 44)):
    xs = [torch.linspace(0, 1, s) for s in shape]
    v = torch.sin(6 * xs[0])[:, None, None] * torch.cos(4 * xs[1])[None, :, None] + 0.4 * torch.sin(9 * xs[2])[None, None, :]
    v = 2 * (v - v.min()) / (v.max() - v.min()) - 1
    return v.contiguous().cuda()


@pytest.mark.parametrize('drop,w1,w2', [('', 0.0, 0.0), ('smallify', 1e-4, 1e-5)])
@pytest.mark.parametrize('use_graph', [True, False])
def test_fast_trainer_matches_module_path(drop, w1, w2, use_graph):
    from latent_feature_grid_compression_b200 import ops
    from latent_feature_grid_compression_b200.model.Smallify_Dropout import SmallifyLoss
    from latent_feature_grid_compression_b200.training.fast_loop import FastTrainer
    vol = _volume()
    n, steps, seed, lr = 2000, 8, 77, 0.008
    a = _make(drop, 3)
    b = _make(drop, 3)
    b.load_state_dict(copy.deepcopy(a.state_dict()))
    for da, db in zip(a.drop, b.drop):
        if hasattr(da, 'tracker'):
            db.tracker.EMA, db.tracker.EMAVar = db.tracker.init_variance_data(db.betas)
            da.tracker.EMA, da.tracker.EMAVar = da.tracker.init_variance_data(da.betas)
    tr = FastTrainer(a, vol, n, lr=lr, seed=seed, weight_l1=w1, weight_l2=w2, use_graph=use_graph)
    opt = torch.optim.Adam(b.parameters(), lr=lr)
    reg = SmallifyLoss(w1, w2) if drop else None
    losses_b = []
    for s in range(steps):
        tr.step()
        raw, norm, gt = ops.sample(vol.shape, n, seed=seed, sample_offset=s * n, volume=vol, want_gt=True)
        opt.zero_grad()
        pred = b(norm).squeeze(-1)
        mse = torch.nn.functional.mse_loss(pred, gt)
        loss = mse + (reg(b) if reg is not None else 0.0)
        loss.backward()
        opt.step()
        losses_b.append(float(mse))
    assert tr.steps_done == steps and int(tr.step_dev[0]) == steps
    assert abs(tr.last_loss() - losses_b[-1]) <= 1e-4 * abs(losses_b[-1])
    sa, sb = a.state_dict(), b.state_dict()
    for k in sb:
        ref = sb[k].float()
        assert float((sa[k].float() - ref).abs().max()) <= 2e-4 * max(float(ref.abs().max()), 1e-6), k
    if drop:
        for da, db in zip(a.drop, b.drop):
            assert torch.allclose(da.tracker.EMAVar, db.tracker.EMAVar, atol=1e-6)
    # the model's parameters are live views of the trainer's flat buffer
    assert a.final_layer.bias.data_ptr() == tr.flat_p.data_ptr() + 4 * (tr.flat_p.numel() - 1)


def test_two_phase_training_learns_and_reports():
    """train_volume = the reference's training() schedule on the fast loop: PSNR must improve a lot over the
    untrained model and the Smallify run must actually prune."""
    from latent_feature_grid_compression_b200.training.fast_loop import train_volume
    vol = _volume((48, 48, 48)).cpu()
    args = dict(d_in=3, n_hidden_size=32, d_out=1, n_layers=4, embedding_type='fourier', n_embedding_freq=2,
                drop_type='smallify', drop_momentum=0.025, drop_threshold=0.75, wavelet_filter='db2', grid_features=8,
                grid_size=15, checkpoint_path='', lr=0.008, max_pass=12, pass_decay=20, lr_decay=0.2,
                lambda_drop_loss=1e-8, lambda_weight_loss=1e-8, batch_size=256, sample_size=16)
    torch.manual_seed(0)
    info = train_volume(args, volume=vol, seed=5)
    assert info['psnr'] > 30.0
    assert info['steps'] > 0 and info['compression_ratio'] > 0
    assert not any(k.startswith('drop.') for k in info['model'].state_dict())


def test_store_and_restore_model_round_trip(tmp_path):
    """Reference binary layout (header + fp32 first/last layer + 8-bit k-means codebooks + mask bitstream)."""
    import struct
    from latent_feature_grid_compression_b200.model.model_utils import restore_model, store_model_parameters
    m = _make('', 1, C=4)
    with torch.no_grad():  # prune some coefficients so that the mask stream matters
        for f in m.feature_grid:
            f[f.abs() < 0.05] = 0.0
    path = str(tmp_path / 'binary_model_file')
    store_model_parameters(m, path)
    raw = open(path, 'rb').read()
    assert struct.unpack('9B', raw[:9]) == (4, 32, 19, 3, 1, 8, 15, 3, 4)
    n_nonzero = struct.unpack('3I', raw[9:21])
    assert list(n_nonzero) == [int(torch.count_nonzero(f)) for f in m.feature_grid]
    r = restore_model(path).cuda().eval()
    m.eval()
    # first/last layers are stored exactly, zeros stay zeros, the rest within the 256-centre quantisation error
    assert torch.equal(r.net_layers[0].weight.cpu(), m.net_layers[0].weight.cpu())
    assert torch.equal(r.final_layer.weight.cpu(), m.final_layer.weight.cpu())
    for a, b in zip(r.feature_grid, m.feature_grid):
        assert torch.equal(a.cpu() == 0, b.cpu() == 0)
        assert float((a.cpu() - b.cpu()).abs().max()) < 0.05 * float(b.abs().max())
    tile = (torch.rand(1, 1, 6, 6, 6, 3, device='cuda') * 2 - 1)
    with torch.no_grad():
        assert float((r(tile) - m(tile)).abs().max()) < 0.2


def test_host_fed_step_equals_device_sampled_step():
    """FastTrainer.step_host on the samples the in-kernel sampler would draw == FastTrainer.step."""
    from latent_feature_grid_compression_b200 import ops
    from latent_feature_grid_compression_b200.training.fast_loop import FastTrainer
    vol = _volume()
    n, steps, seed = 3000, 6, 21
    a = _make('', 4)
    b = _make('', 4)
    b.load_state_dict(copy.deepcopy(a.state_dict()))
    ta = FastTrainer(a, vol, n, lr=0.008, seed=seed)
    tb = FastTrainer(b, vol, n, lr=0.008, seed=seed)
    for s in range(steps):
        ta.step()
        raw, norm, gt = ops.sample(vol.shape, n, seed=seed, sample_offset=s * n, volume=vol, want_gt=True)
        tb.step_host(norm.cpu().pin_memory(), gt.cpu().pin_memory())
        assert abs(ta.last_loss() - tb.last_loss()) <= 1e-5 * abs(ta.last_loss())
    assert float((ta.flat_p - tb.flat_p).abs().max()) <= 1e-5 * float(ta.flat_p.abs().max())


def test_pipelined_host_fed_step_equals_serial_one():
    """step_host_pipelined (copy stream + two staging buffer pairs + deferred loss read) == step_host, step by step."""
    from latent_feature_grid_compression_b200 import ops
    from latent_feature_grid_compression_b200.training.fast_loop import FastTrainer
    vol = _volume()
    n, steps, seed = 3000, 7, 5
    a = _make('', 4)
    b = _make('', 4)
    b.load_state_dict(copy.deepcopy(a.state_dict()))
    ta = FastTrainer(a, vol, n, lr=0.008, seed=seed)
    tb = FastTrainer(b, vol, n, lr=0.008, seed=seed)
    serial, piped = [], []
    for s in range(steps):
        raw, norm, gt = ops.sample(vol.shape, n, seed=seed, sample_offset=s * n, volume=vol, want_gt=True)
        hc, hg = norm.cpu().pin_memory(), gt.cpu().pin_memory()
        ta.step_host(hc, hg)
        serial.append(ta.last_loss())
        prev = tb.step_host_pipelined(hc, hg)
        assert (prev is None) == (s == 0)
        if prev is not None:
            piped.append(prev)
    piped.append(tb.flush_host_pipeline())
    assert len(piped) == steps
    for x, y in zip(serial, piped):
        assert abs(x - y) <= 1e-6 * abs(x)
    assert float((ta.flat_p - tb.flat_p).abs().max()) <= 1e-6 * float(ta.flat_p.abs().max())


@pytest.mark.parametrize('wavelet,G,C,w2', [('db2', 15, 16, 0.0), ('haar', 16, 8, 0.0), ('db2', 5, 6, 0.0), ('db2', 15, 8, 1e-4),
                                           ('db2', 17, 5, 0.0)])
@pytest.mark.parametrize('split', ['1', '0'])
def test_grid_step_path_equals_separate_kernels(wavelet, G, C, w2, split, monkeypatch):
    """lfgc_grid_step (partial reduction + adjoint + Adam + next synthesis in ONE launch, per-channel CTAs, separable
    levels) against the separate kernels over several optimiser steps, weight-decay term included."""
    from latent_feature_grid_compression_b200 import ops
    from latent_feature_grid_compression_b200.model.model_utils import setup_model
    from latent_feature_grid_compression_b200.training.fast_loop import FastTrainer
    vol = _volume()

    def make():
        torch.manual_seed(2)
        return setup_model(3, 32, 1, 4, 'fourier', 2, '', 0.1, 0.9, wavelet, C, G, '').cuda().train()
    a, b = make(), make()
    b.load_state_dict(copy.deepcopy(a.state_dict()))
    monkeypatch.setenv('LFGC_GRID_STEP', '0')
    ta = FastTrainer(a, vol, 3000, lr=0.008, seed=4, weight_l2=w2)
    monkeypatch.setenv('LFGC_GRID_STEP', '1')
    monkeypatch.setenv('LFGC_GRID_STEP_SPLIT', split)   # finest level on the whole GPU (default) / whole pyramid per CTA
    tb = FastTrainer(b, vol, 3000, lr=0.008, seed=4, weight_l2=w2)
    assert tb._gstep and not ta._gstep
    for s in range(7):
        ta.step()
        tb.step()
        if s == 0:      # first step: identical inputs, so the gradients themselves are comparable
            torch.cuda.synchronize()
            ga, gb = ta.flat_g, tb.flat_g
            assert float((ga - gb).abs().max()) <= 2e-6 * float(ga.abs().max())
    assert int(tb.step_dev[0]) == 7 and int(tb.step_dev[1]) == 0
    assert tb.launches_per_step == (4 if split == '1' and len(tb.coeff_params) > 1 else 2)
    assert abs(ta.last_loss() - tb.last_loss()) <= 1e-5 * abs(ta.last_loss())
    assert float((ta.flat_p - tb.flat_p).abs().max()) <= 1e-5 * float(ta.flat_p.abs().max())
    assert float((ta.flat_m - tb.flat_m).abs().max()) <= 1e-5 * float(ta.flat_m.abs().max())
    # the step leaves the grid of the UPDATED coefficients behind (the separate path decodes at the next step's start)
    fresh = ops.decode_fwd(tb.geom, [p.data for p in tb.coeff_params], [None] * len(tb.coeff_params))
    assert float((fresh - tb.grid_cl).abs().max()) <= 2e-6 * float(fresh.abs().max())
    assert float(tb.grad_grid.abs().max()) == 0.0
    # the parameters the model exposes are the updated ones (views into the flat buffer)
    assert b.feature_grid[0].data_ptr() == tb.flat_p.data_ptr()
    # the tensor-core operand image the step keeps current (csrc/tc_panels.cuh) is bit-identical to one rebuilt from the
    # updated parameters
    assert tb._tc_panels is not None
    rebuilt = ops.tc_panel_image(tb.geom, tb.mlp_flat)
    torch.cuda.synchronize()
    assert torch.equal(rebuilt.view(torch.int32), tb._tc_panels.view(torch.int32))


def test_grid_step_falls_back_when_the_pyramid_does_not_fit_or_masks_are_live():
    from latent_feature_grid_compression_b200.model.model_utils import setup_model
    from latent_feature_grid_compression_b200.training.fast_loop import FastTrainer
    vol = _volume()
    torch.manual_seed(1)
    big = setup_model(3, 32, 1, 4, 'fourier', 2, '', 0.1, 0.9, 'db2', 4, 96, '').cuda().train()   # 49^3 per channel below the finest level
    t = FastTrainer(big, vol, 2000, lr=0.008, seed=1)
    assert not t._gstep
    t.step()
    masked = _make('smallify', 4)
    t2 = FastTrainer(masked, vol, 2000, lr=0.008, seed=1, weight_l1=1e-6, weight_l2=1e-6)
    assert not t2._gstep
    t2.step()
    torch.cuda.synchronize()


def test_grid_step_path_notices_parameters_changed_from_outside():
    """The grid-step path carries the decoded grid and the tensor-core operand image from step to step; in-place torch
    writes to the parameters between steps must refresh both (FastTrainer._run's version check / refresh())."""
    from latent_feature_grid_compression_b200 import ops
    from latent_feature_grid_compression_b200.model.model_utils import setup_model
    from latent_feature_grid_compression_b200.training.fast_loop import FastTrainer
    torch.manual_seed(3)
    model = setup_model(3, 32, 1, 4, 'fourier', 2, '', 0.1, 0.9, 'db2', 16, 15, '').cuda().train()
    tr = FastTrainer(model, _volume(), 4096, lr=0.008, seed=1)
    assert tr._gstep
    for _ in range(3):
        tr.step()
    torch.cuda.synchronize()
    with torch.no_grad():
        for p in model.parameters():
            p.mul_(0.5)
    before = tr.flat_p.clone()
    calls = []
    prime = tr._prime_gstep
    tr._prime_gstep = lambda: (calls.append(1), prime())[1]
    tr.step()
    tr.step()
    torch.cuda.synchronize()
    assert len(calls) == 1      # refreshed once, for the step after the outside write, and not again
    # two Adam steps move a parameter by at most ~2 lr: the steps started from the halved parameters
    assert float((tr.flat_p - before).abs().max()) <= 0.05
    rebuilt = ops.tc_panel_image(tr.geom, tr.mlp_flat)
    fresh = ops.decode_fwd(tr.geom, [p.data for p in tr.coeff_params], [None] * len(tr.coeff_params))
    torch.cuda.synchronize()
    assert torch.equal(rebuilt.view(torch.int32), tr._tc_panels.view(torch.int32))
    assert float((fresh - tr.grid_cl).abs().max()) <= 2e-6 * float(fresh.abs().max())
