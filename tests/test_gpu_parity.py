"""GPU parity tests: the CUDA path (through the C ABI) against the reference's own outputs (golden fixtures) and
the numpy oracle.  Tolerances: forward 1e-5 relative (fp32), gradients max|d| <= 1e-5 * max|g| (atomic
re-ordering), integer/index work bit-exact."""
import numpy as np
import pytest
import torch

from oracle import fvsrn_numpy as O
from tests.util import MODEL_CASES, build_model, load, noise_of, relerr, replay_noise, state

pytestmark = pytest.mark.gpu

FWD_TOL = 1e-5
GRAD_TOL = 1e-5


@pytest.mark.parametrize('tag', list(MODEL_CASES))
def test_decode_volume(tag):
    model, g, cfg = build_model(tag)
    with replay_noise(noise_of(g)):
        grid = model.decode_volume()
    assert relerr(grid.cpu().numpy(), g['grid']) < 2e-6


@pytest.mark.parametrize('tag', list(MODEL_CASES))
def test_encode_volume_matches_oracle(tag):
    g = load('model_' + tag)
    if MODEL_CASES[tag]['mask']:
        pytest.skip('grid fixture is masked')
    from latent_feature_grid_compression_b200.model.model_utils import setup_model
    C, G, H, L, F, N = [int(v) for v in g['meta']]
    model = setup_model(3, H, 1, L, 'fourier', F, '', 0.1, 0.9, MODEL_CASES[tag]['wavelet'], C, G, '')
    feats, shapes = model.encode_volume(torch.from_numpy(g['grid']))
    ref, ref_shapes = O.encode_volume(g['grid'].astype(np.float64), MODEL_CASES[tag]['wavelet'])
    assert np.array_equal(np.asarray(shapes).reshape(-1, 3), ref_shapes)
    assert np.array_equal(np.asarray(model.shape_array).reshape(-1, 3), g['shape_array'].reshape(-1, 3))
    for a, b in zip(feats, ref):
        assert relerr(a.numpy(), b) < 5e-6  # fp32 analysis chain (up to 4 levels) against the fp64 oracle
    # stored coefficients of the reference model re-synthesise to the same grid: encode(grid) == coefficients
    sd = state(g)
    for i, a in enumerate(feats):
        assert relerr(a.numpy(), sd['feature_grid.%d' % i]) < 5e-6


@pytest.mark.parametrize('tag', list(MODEL_CASES))
def test_forward_train_and_eval(tag):
    model, g, cfg = build_model(tag)
    coords = torch.from_numpy(g['coords']).cuda()
    with replay_noise(noise_of(g)):
        y = model(coords)
    assert y.shape == (coords.shape[0], 1)
    assert relerr(y.detach().cpu().numpy(), g['y_train']) < FWD_TOL
    model.eval()
    tile = torch.from_numpy(g['tile']).cuda()
    with torch.no_grad(), replay_noise(noise_of(g, 'noise_eval')):
        ye = model(tile)
    assert tuple(ye.shape) == tuple(g['y_eval'].shape)
    assert relerr(ye.cpu().numpy(), g['y_eval']) < FWD_TOL
    assert float(ye.max()) <= 1.0 and float(ye.min()) >= -1.0


@pytest.mark.parametrize('tag', [t for t in MODEL_CASES if 'h64' not in t])
def test_backward_all_parameters(tag):
    model, g, cfg = build_model(tag)
    coords = torch.from_numpy(g['coords']).cuda().requires_grad_(True)
    wout = torch.from_numpy(g['wout']).cuda()
    with replay_noise(noise_of(g)):
        y = model(coords)
    (y * wout).sum().backward()
    checked = 0
    for name, prm in model.named_parameters():
        ref = g['grad.' + name]
        if ref.size == 0:
            assert prm.grad is None or float(prm.grad.abs().max()) == 0.0
            continue
        assert prm.grad is not None, name
        tol = GRAD_TOL * max(float(np.abs(ref).max()), 1e-30)
        assert float(np.abs(prm.grad.cpu().numpy() - ref).max()) <= tol, name
        checked += 1
    assert checked >= 2 * (int(g['meta'][3]) + 1) + len(model.feature_grid)


def test_backward_wide_hidden_is_refused_loudly():
    """H=64 forward is supported; the fused backward is built for H<=32 and must raise, not fall back."""
    from latent_feature_grid_compression_b200._lib import LfgcError
    model, g, cfg = build_model('basic_db2_c8_g17_h64_l3_f3')
    y = model(torch.from_numpy(g['coords']).cuda())
    with pytest.raises(LfgcError):
        y.sum().backward()


@pytest.mark.parametrize('tag', ['smallify_db2_c6_g15', 'variational_db2_c8_g15', 'maskedste_db2_c8_g15'])
def test_mask_baking(tag):
    model, g, cfg = build_model(tag)
    coords = torch.from_numpy(g['coords']).cuda()
    zeros = model.save_dropvalues_on_grid(torch.device('cuda'))
    assert abs(float(zeros) - float(g['bake.zeros'])) < 1e-3
    for i, f in enumerate(model.feature_grid):
        assert relerr(f.detach().cpu().numpy(), g['bake.feature_grid.%d' % i]) < 1e-6
        assert np.array_equal(model.drop[i].d_mask.float().cpu().numpy(), g['bake.d_mask.%d' % i])
    with torch.no_grad():
        yb = model(coords)
    assert relerr(yb.cpu().numpy(), g['y_baked_train']) < FWD_TOL
    model.remove_drop_layers(torch.device('cuda'))
    for i, f in enumerate(model.feature_grid):
        assert relerr(f.detach().cpu().numpy(), g['final.feature_grid.%d' % i]) < 1e-6
    with torch.no_grad():
        yf = model(coords)
    assert relerr(yf.cpu().numpy(), g['y_final_train']) < FWD_TOL
    assert not any(k.startswith('drop.') for k in model.state_dict())


def test_regulariser_losses():
    from latent_feature_grid_compression_b200.model.Smallify_Dropout import SmallifyLoss
    from latent_feature_grid_compression_b200.model.Variational_Dropout_Layer import VariationalDropoutLoss
    model, g, cfg = build_model('smallify_db2_c6_g15')
    v = SmallifyLoss(0.37, 1.9)(model)
    v.backward()
    assert abs(float(v) - float(g['smallify_loss'])) < 1e-5 * abs(float(g['smallify_loss']))
    for name, prm in model.named_parameters():
        if 'sgrad.' + name in g:
            assert relerr(prm.grad.cpu().numpy(), g['sgrad.' + name]) < 1e-5
    model, g, cfg = build_model('variational_db2_c8_g15')
    coords = torch.from_numpy(g['coords']).cuda()
    N = coords.shape[0]
    vl = VariationalDropoutLoss(size_volume=1000.0, batch_size=float(N), weight_dkl=1.3, weight_weights=0.7)
    with replay_noise(noise_of(g)):
        pred = model(coords).squeeze(-1)
    logsig = torch.from_numpy(g['vloss.logsig']).cuda().requires_grad_(True)
    tot, ll, mse, dkl, wsum = vl(model, pred, torch.from_numpy(g['vloss.gt']).cuda(), logsig, 5e-5)
    tot.backward()
    got = [float(tot), float(ll), float(mse), float(dkl), float(wsum), vl.weight_dkl]
    for a, b in zip(got, g['vloss.values']):
        assert abs(a - b) <= 3e-5 * abs(b)
    assert relerr(logsig.grad.cpu().numpy(), g['vloss.grad_logsig']) < 1e-4
    for name, prm in model.named_parameters():
        ref = g['vgrad.' + name]
        assert float(np.abs(prm.grad.cpu().numpy() - ref).max()) <= 2e-5 * float(np.abs(ref).max()), name


def test_smallify_tracker_device_side():
    from latent_feature_grid_compression_b200.model.Smallify_Dropout import SmallifyDropout
    g = load('smallify_tracker')
    mom, thr = [float(v) for v in g['momentum_threshold']]
    d = SmallifyDropout((3, 4, 5), mom, thr)
    with torch.no_grad():
        d.betas.copy_(torch.from_numpy(g['betas0']))
    d.tracker.EMA, d.tracker.EMAVar = d.tracker.init_variance_data(d.betas)
    d.cuda()
    d.train()
    x = torch.from_numpy(g['x']).cuda()
    for step in range(12):
        y = d(x)
        with torch.no_grad():
            flip = torch.from_numpy(g['flip.%d' % step]).cuda()
            d.betas[flip] *= -1.0
    assert np.allclose(y.detach().cpu().numpy(), g['y_last'], rtol=1e-6, atol=0)
    assert np.array_equal(d.tracker.EMA.cpu().numpy(), g['EMA'])        # same fp32 operation order: bit-exact
    assert np.array_equal(d.tracker.EMAVar.cpu().numpy(), g['EMAVar'])
    assert np.array_equal(d.calculate_pruning_mask(torch.device('cuda')).cpu().numpy(), g['mask'])
    assert relerr(d.multiply_values_with_dropout(x, torch.device('cuda')).cpu().numpy(), g['baked']) < 1e-6


def test_sampler_and_ground_truth_bit_exact():
    from latent_feature_grid_compression_b200.data.IndexDataset import IndexDataset
    from latent_feature_grid_compression_b200.data.Interpolation import trilinear_f_interpolation
    g = load('dataset')
    for tag in ('a', 'b'):
        vol = torch.from_numpy(g[tag + '.vol'])
        ds = IndexDataset(vol, 16)
        assert ds.n_voxels == int(g[tag + '.n_voxels'][0])
        assert np.array_equal(ds.max_idx.numpy(), g[tag + '.max_idx'])
        assert np.array_equal(ds.scales.numpy(), g[tag + '.scales'])
        idx = torch.from_numpy(g[tag + '.idx']).cuda()
        volc = vol.cuda()
        raw, norm, gt = ds.sample(idx.numel(), volume=volc, explicit_idx=idx)
        assert np.array_equal(raw.cpu().numpy(), g[tag + '.raw'])
        assert np.array_equal(norm.cpu().numpy(), g[tag + '.norm'])
        assert np.array_equal(gt.cpu().numpy(), g[tag + '.gt_int'])
        # the reference call, unchanged signature (training/training.py:107-109)
        gt2 = trilinear_f_interpolation(raw, volc, ds.min_idx.cuda(), ds.max_idx.cuda(), ds.vol_res.cuda())
        assert np.array_equal(gt2.cpu().numpy(), g[tag + '.gt_int'])
        gtf = trilinear_f_interpolation(torch.from_numpy(g[tag + '.pf']).cuda(), volc, ds.min_idx, ds.max_idx, ds.vol_res)
        assert float(np.abs(gtf.cpu().numpy() - g[tag + '.gt_float']).max()) <= 1e-6
        bb = g[tag + '.bb']
        gtb = trilinear_f_interpolation(torch.from_numpy(g[tag + '.pb']).cuda(), volc, torch.from_numpy(bb[0]),
                                        torch.from_numpy(bb[1]), ds.vol_res)
        assert float(np.abs(gtb.cpu().numpy() - g[tag + '.gt_bb']).max()) <= 1e-6
        # CPU-side __getitem__ keeps the reference's generator stream
        torch.manual_seed(17)
        r2, n2 = ds[0]
        assert np.array_equal(r2.numpy(), g[tag + '.raw']) and np.array_equal(n2.numpy(), g[tag + '.norm'])


def test_philox_sampler_properties():
    from latent_feature_grid_compression_b200 import ops
    shape = (255, 255, 255)
    n = 1 << 20
    raw, norm, _ = ops.sample(shape, n, seed=7, sample_offset=0)
    raw2, _, _ = ops.sample(shape, n, seed=7, sample_offset=0)
    assert torch.equal(raw, raw2)                                    # reproducible
    raw3, _, _ = ops.sample(shape, n // 2, seed=7, sample_offset=n // 2)
    assert torch.equal(raw3, raw[n // 2:])                           # counter-based: offset = position in the stream
    raw4, _, _ = ops.sample(shape, n, seed=8, sample_offset=0)
    assert not torch.equal(raw, raw4)
    assert float(raw.min()) >= 0 and float(raw.max()) <= 254
    flat = (raw[:, 0] * 255 + raw[:, 1]) * 255 + raw[:, 2]
    # uniformity: chi-square over 64 bins of the flat index
    hist = torch.histc(flat, bins=64, min=0, max=255 ** 3).cpu().numpy()
    chi2 = ((hist - n / 64) ** 2 / (n / 64)).sum()
    assert chi2 < 130.0                                              # 63 dof: P(chi2 > 130) ~ 1e-6
    # normalised coordinates obey the oracle's fp32 formula bit for bit
    _, want = O.sample_positions(flat.cpu().numpy().astype(np.int64)[:4096], shape)
    assert np.array_equal(norm[:4096].cpu().numpy(), want)


def test_reconstruction_volume_and_stats():
    from latent_feature_grid_compression_b200.data.IndexDataset import IndexDataset
    from latent_feature_grid_compression_b200.model.model_utils import setup_model
    from latent_feature_grid_compression_b200.visualization.OutputToVTK import (
        axis_tables, calculate_deviation_statistics, field_from_net, tiled_net_out)
    g = load('reconstruct')
    vol = torch.from_numpy(g['vol'])
    ds = IndexDataset(vol, 16)
    model = setup_model(3, 32, 1, 4, 'fourier', 2, '', 0.1, 0.9, 'db2', 4, 15, '')
    model.load_state_dict({k: torch.from_numpy(v) for k, v in state(g).items()})
    model.cuda().eval()
    tabs = axis_tables(ds, 32)
    for a in range(3):  # coordinates bit-identical to what the reference's field_from_net fed its network
        assert np.array_equal(tabs[a].cpu().numpy(), g['axis%d' % a])
    full = field_from_net(ds, model, True, 32)
    assert tuple(full.shape) == tuple(vol.shape)
    assert float(np.abs(full.numpy() - g['full']).max()) < 2e-5
    # slabs (the multi-GPU sharding unit) tile the volume exactly
    a = field_from_net(ds, model, True, 32, slab=(0, 17))
    b = field_from_net(ds, model, True, 32, slab=(17, 40))
    assert torch.equal(torch.cat([a, b], 0), full)
    stats = calculate_deviation_statistics(torch.from_numpy(g['full']), vol)
    for got, want_v in zip(stats, g['stats']):
        assert abs(got - want_v) <= 1e-5 * abs(want_v)
    psnr, l1, mse, rmse = tiled_net_out(ds, model, True, gt_vol=vol, evaluate=True, write_vols=False)
    assert abs(psnr - g['stats'][0]) < 1e-3 and model.training


def test_fused_train_step_matches_separate_path_and_oracle():
    """lfgc_train_step (in-kernel sampler + GT + MSE + backward) == lfgc_sample + lfgc_backward == oracle."""
    from latent_feature_grid_compression_b200 import ops
    model, g, cfg = build_model('basic_db2_c16_g15')
    geom = model.geometry()
    gen = torch.Generator().manual_seed(3)
    vol = (torch.rand(31, 29, 37, generator=gen) * 2 - 1).cuda()
    n = 1000  # not a multiple of the tile: exercises the tail
    idx = torch.randint(0, vol.numel(), (n,), generator=gen).cuda()
    coeffs = [f.detach().contiguous() for f in model.feature_grid]
    grid_cl = ops.decode_fwd(geom, coeffs, [None] * len(coeffs))
    mlp = model.mlp_flat()
    ws = torch.empty(geom.backward_workspace_bytes // 4, device='cuda')
    gg = torch.zeros_like(grid_cl)
    gm = torch.empty(geom.mlp_param_count, device='cuda')
    loss = torch.zeros(1, device='cuda')
    ops.train_step(geom, vol, n, 0, 0, 1.0 / n, grid_cl, mlp, gg, gm, loss, ws, explicit_idx=idx)
    # separate path
    raw, norm, gt = ops.sample(vol.shape, n, volume=vol, explicit_idx=idx, want_gt=True)
    pred = ops.sample_forward(geom, norm, grid_cl, mlp)
    gout = (2.0 / n) * (pred - gt)
    gg2, gm2 = ops.sample_backward(geom, norm, gout.contiguous(), grid_cl, mlp)
    assert float((gm - gm2).abs().max()) <= 2e-6 * float(gm2.abs().max())
    assert float((gg - gg2).abs().max()) <= 2e-6 * float(gg2.abs().max())
    assert abs(float(loss) - float(((pred - gt) ** 2).sum())) <= 1e-5 * float(loss)
    # oracle
    spec = O.Spec(16, 15, 32, 4, 2, 'db2', '')
    sd = state(g)
    y, ctx = O.model_forward(sd, spec, norm.cpu().numpy(), training=True, keep=True)
    assert relerr(pred.cpu().numpy(), y[:, 0]) < FWD_TOL
    grads = O.model_backward((2.0 / n) * (y - gt.cpu().numpy().astype(np.float64)[:, None]), ctx, spec)
    want_grid = np.transpose(grads['grid'], (1, 2, 3, 0))
    assert float(np.abs(gg.cpu().numpy()[..., :16] - want_grid).max()) <= GRAD_TOL * np.abs(want_grid).max()
    off = 0
    for name, shape in geom.mlp_shapes():
        k = int(np.prod(shape))
        ref = grads[name].reshape(-1)
        assert float(np.abs(gm[off:off + k].cpu().numpy() - ref).max()) <= GRAD_TOL * max(np.abs(ref).max(), 1e-30), name
        off += k


@pytest.mark.parametrize('n', [1000, 32768])
@pytest.mark.parametrize('tc', ['1', '0'])
def test_train_step_accumulate_adds_the_same_sums(n, tc, monkeypatch):
    """lfgc_train_step_accumulate (atomics from the tensor-core kernel's epilogue / the reduction kernel in add mode for the
    FFMA2 kernels) adds exactly what lfgc_train_step writes: MLP gradient, loss (last float), grid gradient."""
    from latent_feature_grid_compression_b200 import ops
    monkeypatch.setenv('LFGC_BACKWARD_TC', tc)
    model, g, cfg = build_model('basic_db2_c16_g15')
    geom = model.geometry()
    gen = torch.Generator().manual_seed(11)
    vol = (torch.rand(31, 29, 37, generator=gen) * 2 - 1).cuda()
    idx = torch.randint(0, vol.numel(), (n,), generator=gen).cuda()
    coeffs = [f.detach().contiguous() for f in model.feature_grid]
    grid_cl = ops.decode_fwd(geom, coeffs, [None] * len(coeffs))
    mlp = model.mlp_flat()
    ws = torch.empty(geom.backward_workspace_bytes // 4, device='cuda')
    gg = torch.zeros_like(grid_cl)
    gm = torch.empty(geom.mlp_param_count, device='cuda')
    loss = torch.zeros(1, device='cuda')
    ops.train_step(geom, vol, n, 0, 0, 1.0 / n, grid_cl, mlp, gg, gm, loss, ws, explicit_idx=idx)
    gg2 = torch.zeros_like(grid_cl)
    K = 1 if n == 1000 else 8    # rows the CTAs are spread over (the result is the sum of the rows)
    acc = torch.zeros(K * (geom.mlp_param_count + 1), device='cuda')
    for rep in (1, 2):     # a running sum: the second call doubles it
        ops.train_step_accumulate(geom, vol, n, 0, 0, 1.0 / n, grid_cl, mlp, gg2, acc, ws, explicit_idx=idx, n_slices=K)
        torch.cuda.synchronize()
        tot = acc.view(K, -1).sum(0)
        assert float((tot[:-1] - rep * gm).abs().max()) <= 3e-6 * rep * float(gm.abs().max())
        assert abs(float(tot[-1]) - rep * float(loss)) <= 1e-5 * rep * float(loss)
        assert float((gg2 - rep * gg).abs().max()) <= 3e-6 * rep * float(gg.abs().max())


def test_flat_adam_matches_torch_adam():
    from latent_feature_grid_compression_b200 import ops
    gen = torch.Generator().manual_seed(5)
    p0 = torch.randn(5000, generator=gen)
    p_ref = p0.clone().cuda().requires_grad_(True)
    opt = torch.optim.Adam([p_ref], lr=0.008)
    p = p0.clone().cuda()
    m = torch.zeros_like(p)
    v = torch.zeros_like(p)
    lr = torch.tensor([0.008], device='cuda')
    step = torch.zeros(2, dtype=torch.int32, device='cuda')
    for s in range(25):
        gr = torch.randn(5000, generator=gen).cuda() * (0.1 + s)
        p_ref.grad = gr.clone()
        opt.step()
        ops.adam(p, gr, m, v, lr, step)
        if s == 10:
            lr.mul_(0.2)
            opt.param_groups[0]['lr'] *= 0.2
    assert int(step[0]) == 25 and int(step[1]) == 0
    assert float((p - p_ref.detach()).abs().max()) <= 2e-6 * float(p_ref.abs().max())


@pytest.mark.parametrize('tag', ['basic', 'smallify'])
def test_training_trajectory_through_module_api(tag):
    """A dozen optimiser steps written exactly like training/training.py:89-138, on the reference's batches."""
    from latent_feature_grid_compression_b200.data.IndexDataset import IndexDataset
    from latent_feature_grid_compression_b200.data.Interpolation import trilinear_f_interpolation
    from latent_feature_grid_compression_b200.model.model_utils import setup_model
    from latent_feature_grid_compression_b200.model.Smallify_Dropout import SmallifyLoss
    g = load('trajectory_' + tag)
    device = torch.device('cuda')
    volume = torch.from_numpy(g['vol'])
    dataset = IndexDataset(volume, 16)
    volume = volume.to(device)
    model = setup_model(3, 32, 1, 4, 'fourier', 2, 'smallify' if tag == 'smallify' else '', 0.1, 0.9, 'db2', 8, 15, '')
    model.load_state_dict({k: torch.from_numpy(v) for k, v in state(g, 'sd0.').items()})
    if tag == 'smallify':
        for d in model.drop:  # the tracker was initialised from the pre-load betas: re-initialise like a fresh model
            d.tracker.EMA, d.tracker.EMAVar = d.tracker.init_variance_data(d.betas)
    model.to(device)
    model.train()
    optimizer = torch.optim.Adam(model.parameters(), lr=0.008)
    loss_criterion = torch.nn.MSELoss().to(device)
    drop_loss = SmallifyLoss(1e-4, 1e-5) if tag == 'smallify' else None
    losses = []
    for s in range(g['idx'].shape[0]):
        raw_positions = dataset.volume_indices[torch.from_numpy(g['idx'][s])]
        norm_positions = dataset.scales.unsqueeze(0) * (2.0 * (raw_positions - dataset.min_idx) / (dataset.max_idx - dataset.min_idx) - 1.0)
        raw_positions = raw_positions.to(device).view(-1, 3)
        norm_positions = norm_positions.to(device).view(-1, 3)
        norm_positions.requires_grad = True
        optimizer.zero_grad()
        predicted_volume = model(norm_positions).squeeze(-1)
        ground_truth_volume = trilinear_f_interpolation(raw_positions, volume, dataset.min_idx.to(device),
                                                        dataset.max_idx.to(device), dataset.vol_res.to(device))
        vol_loss = loss_criterion(predicted_volume, ground_truth_volume)
        complete_loss = vol_loss + (drop_loss(model) if drop_loss is not None else torch.zeros_like(vol_loss))
        complete_loss.backward()
        optimizer.step()
        losses.append(complete_loss.item())
    assert float(np.abs(np.asarray(losses) - g['losses']).max()) < 5e-4 * float(np.abs(g['losses']).max())
    for k, v in model.state_dict().items():
        ref = g['sd1.' + k]
        assert float(np.abs(v.cpu().numpy() - ref).max()) < 5e-3 * max(float(np.abs(ref).max()), 1e-6), k
    if tag == 'smallify':
        for i, d in enumerate(model.drop):
            assert np.allclose(d.tracker.EMA.cpu().numpy(), g['EMA.%d' % i], atol=1e-6)
            assert np.allclose(d.tracker.EMAVar.cpu().numpy(), g['EMAVar.%d' % i], atol=1e-6)


@pytest.mark.parametrize('C,G,H,Lyr,F,n', [(8, 15, 20, 3, 2, 5000), (4, 9, 4, 2, 1, 700), (24, 12, 31, 4, 2, 3000),
                                         (5, 15, 12, 1, 3, 1500)])
def test_narrow_hidden_widths_against_oracle(C, G, H, Lyr, F, n):
    """n_hidden_size below the 32-wide register tile (the NAS range is 4..32, Multi_Objective_NAS.py:127-132): forward
    and every parameter gradient against the numpy oracle."""
    from latent_feature_grid_compression_b200.model.model_utils import setup_model
    from oracle import fvsrn_numpy as O
    torch.manual_seed(C * 100 + H)
    model = setup_model(3, H, 1, Lyr, 'fourier', F, '', 0.1, 0.9, 'db2', C, G, '').cuda().train()
    with torch.no_grad():
        for lyr in model.net_layers:
            lyr.bias.uniform_(-0.5, 0.5)
    coords = (torch.rand(n, 3, device='cuda') * 2.1 - 1.05).requires_grad_(True)
    w = torch.randn(n, 1, device='cuda')
    y = model(coords)
    (y * w).sum().backward()
    sd = {k: v.detach().cpu().numpy() for k, v in model.state_dict().items()}
    spec = O.Spec(C, G, H, Lyr, F, 'db2', '')
    yo, ctx = O.model_forward(sd, spec, coords.detach().cpu().numpy(), training=True, keep=True)
    go = O.model_backward(w.cpu().numpy(), ctx, spec)
    assert relerr(y.detach().cpu().numpy(), yo) < FWD_TOL
    for name, p in model.named_parameters():
        ref = go[name]
        assert float(np.abs(p.grad.cpu().numpy() - ref).max()) <= GRAD_TOL * float(np.abs(ref).max()), name
