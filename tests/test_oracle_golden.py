"""Pin the numpy oracle against fixtures produced by the reference itself (tests/golden/make_golden.py)."""
import glob
import os

import numpy as np
import pytest

from oracle import fvsrn_numpy as O

GOLD = os.path.join(os.path.dirname(__file__), 'golden')

MODEL_CASES = {
    'basic_db2_c16_g15': dict(mask='', wavelet='db2', thr=0.9),
    'basic_haar_c4_g16': dict(mask='', wavelet='haar', thr=0.9),
    'basic_db2_c8_g17_h64_l3_f3': dict(mask='', wavelet='db2', thr=0.9),
    'basic_db2_c4_g5_nolevels': dict(mask='', wavelet='db2', thr=0.9),
    'smallify_db2_c6_g15': dict(mask='smallify', wavelet='db2', thr=0.75),
    'variational_db2_c8_g15': dict(mask='variational', wavelet='db2', thr=0.5),
    'maskedste_db2_c8_g15': dict(mask='masked_ste', wavelet='db2', thr=0.6),
    'bernoulli_db2_c8_g15': dict(mask='bernoulli', wavelet='db2', thr=0.5),
}


def load(tag):
    return dict(np.load(os.path.join(GOLD, tag + '.npz')))


def spec_of(g, cfg):
    C, G, H, L, F, N = [int(v) for v in g['meta']]
    return O.Spec(C, G, H, L, F, cfg['wavelet'], cfg['mask'], cfg['thr'], g['shape_array'])


def state(g, prefix='sd.'):
    return {k[len(prefix):]: v for k, v in g.items() if k.startswith(prefix)}


def relerr(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


def test_all_fixture_files_present():
    names = {os.path.basename(p) for p in glob.glob(os.path.join(GOLD, '*.npz'))}
    for tag in MODEL_CASES:
        assert 'model_%s.npz' % tag in names
    for n in ('dataset.npz', 'reconstruct.npz', 'smallify_tracker.npz', 'variance_model.npz',
              'trajectory_basic.npz', 'trajectory_smallify.npz'):
        assert n in names


@pytest.mark.parametrize('tag', list(MODEL_CASES))
def test_shape_array_and_level_count(tag):
    g = load('model_' + tag)
    cfg = MODEL_CASES[tag]
    C, G, H, L, F, N = [int(v) for v in g['meta']]
    derived = O.Spec(C, G, H, L, F, cfg['wavelet'], cfg['mask'], cfg['thr'])
    assert np.array_equal(derived.shape_array, g['shape_array'].reshape(-1, 3))
    sd = state(g)
    assert derived.n_levels == sum(1 for k in sd if k.startswith('feature_grid.'))


@pytest.mark.parametrize('tag', list(MODEL_CASES))
def test_decode_volume_matches_reference(tag):
    g = load('model_' + tag)
    cfg = MODEL_CASES[tag]
    spec = spec_of(g, cfg)
    noise = [g['noise.%d' % i] for i in range(int(g['n_noise'][0]))] or None
    _, ctx = O.model_forward(state(g), spec, g['coords'], noise=noise, training=True, keep=True)
    assert relerr(ctx['grid'], g['grid']) < 2e-6


@pytest.mark.parametrize('tag', list(MODEL_CASES))
def test_wavelet_perfect_reconstruction_and_encode(tag):
    """decode(encode(x)) == x (the only self-checking property the reference offers) and
    encode(decoded reference grid) reproduces the stored coefficients when no mask is active."""
    g = load('model_' + tag)
    cfg = MODEL_CASES[tag]
    if cfg['mask']:
        pytest.skip('masked: grid is not the plain synthesis of the coefficients')
    spec = spec_of(g, cfg)
    grid = g['grid'].astype(np.float64)
    coeffs, shapes = O.encode_volume(grid, cfg['wavelet'])
    assert np.array_equal(shapes, spec.shape_array)
    sd = state(g)
    for i, c in enumerate(coeffs):
        assert relerr(c, sd['feature_grid.%d' % i]) < 5e-6
    back = O.decode_volume(coeffs, [None] * len(coeffs), shapes, cfg['wavelet'])
    assert relerr(back, grid) < 2e-6


@pytest.mark.parametrize('tag', list(MODEL_CASES))
def test_forward_train_and_eval(tag):
    g = load('model_' + tag)
    cfg = MODEL_CASES[tag]
    spec = spec_of(g, cfg)
    noise = [g['noise.%d' % i] for i in range(int(g['n_noise'][0]))] or None
    y = O.model_forward(state(g), spec, g['coords'], noise=noise, training=True)
    assert relerr(y, g['y_train']) < 1e-5
    noise_e = [g['noise_eval.%d' % i] for i in range(int(g['n_noise_eval'][0]))] or None
    ye = O.model_forward(state(g), spec, g['tile'].reshape(-1, 3), noise=noise_e, training=False, clamp=True)
    assert relerr(ye.reshape(g['y_eval'].shape), g['y_eval']) < 1e-5


@pytest.mark.parametrize('tag', list(MODEL_CASES))
def test_backward_all_parameters(tag):
    g = load('model_' + tag)
    cfg = MODEL_CASES[tag]
    spec = spec_of(g, cfg)
    noise = [g['noise.%d' % i] for i in range(int(g['n_noise'][0]))] or None
    y, ctx = O.model_forward(state(g), spec, g['coords'], noise=noise, training=True, keep=True)
    grads = O.model_backward(g['wout'], ctx, spec)
    checked = 0
    for k, ref in g.items():
        if not k.startswith('grad.') or k == 'grad.coords':
            continue
        name = k[len('grad.'):]
        if ref.size == 0:  # parameter without gradient in the reference (Bernoulli mask values)
            assert name not in grads or grads[name] is None
            continue
        assert name in grads, name
        tol = 1e-5 * max(np.abs(ref).max(), 1e-30)  # the repo-wide gradient gate
        assert np.abs(grads[name] - ref).max() <= tol, name
        checked += 1
    assert checked >= 2 * (spec.L + 1) + spec.n_levels


def test_regulariser_losses():
    g = load('model_smallify_db2_c6_g15')
    spec = spec_of(g, MODEL_CASES['smallify_db2_c6_g15'])
    val, grads = O.smallify_loss(state(g), spec, 0.37, 1.9)
    assert abs(val - float(g['smallify_loss'])) < 1e-5 * abs(float(g['smallify_loss']))
    for k, ref in g.items():
        if k.startswith('sgrad.'):
            assert relerr(grads[k[len('sgrad.'):]], ref) < 1e-6

    g = load('model_maskedste_db2_c8_g15')
    spec = spec_of(g, MODEL_CASES['maskedste_db2_c8_g15'])
    val, grads = O.smallify_loss(state(g), spec, 0.37, 1.9)
    assert abs(val - float(g['smallify_loss'])) < 1e-5 * abs(float(g['smallify_loss']))

    g = load('model_variational_db2_c8_g15')
    spec = spec_of(g, MODEL_CASES['variational_db2_c8_g15'])
    sd = state(g)
    noise = [g['noise.%d' % i] for i in range(int(g['n_noise'][0]))]
    pred = O.model_forward(sd, spec, g['coords'], noise=noise, training=True)[:, 0]
    N = pred.shape[0]
    r = O.variational_loss(sd, spec, pred, g['vloss.gt'], g['vloss.logsig'], 1000.0, float(N), 1.3, 0.7, 5e-5)
    ref = g['vloss.values']
    for got, want in zip((r['loss'], r['ll'], r['mse'], r['dkl'], r['wsum'], r['weight_dkl']), ref):
        assert abs(got - want) <= 2e-5 * abs(want)
    assert relerr(r['g_logsig'], g['vloss.grad_logsig']) < 1e-5
    for i in range(spec.n_levels):
        assert abs(O.variational_dkl(sd['drop.%d.log_thetas' % i].astype(np.float64),
                                     sd['drop.%d.log_var' % i].astype(np.float64)) - float(g['dkl.%d' % i])) \
            < 1e-5 * abs(float(g['dkl.%d' % i]))
        assert relerr(O.variational_droprate(sd['drop.%d.log_thetas' % i].astype(np.float64),
                                             sd['drop.%d.log_var' % i].astype(np.float64)),
                      g['droprate.%d' % i]) < 1e-5


@pytest.mark.parametrize('tag', ['smallify_db2_c6_g15', 'variational_db2_c8_g15', 'maskedste_db2_c8_g15'])
def test_mask_baking(tag):
    """save_dropvalues_on_grid / remove_drop_layers (Feature_Grid_Model.py:110-140)."""
    g = load('model_' + tag)
    cfg = MODEL_CASES[tag]
    spec = spec_of(g, cfg)
    sd = state(g)
    zeros = 0.0
    mask_bits = 0
    for i in range(spec.n_levels):
        co = sd['feature_grid.%d' % i].astype(np.float64)
        if cfg['mask'] == 'smallify':
            # golden model never ran a forward, so the tracker is at its initial state: EMAVar = 0 < thr -> keep all
            d_mask = O.smallify_prune_mask(np.zeros(co.shape[1:]), cfg['thr'])
            baked = co * (d_mask * sd['drop.%d.betas' % i])[None]
            mask_bits += d_mask.size
        elif cfg['mask'] == 'variational':
            d_mask = O.variational_prune_mask(sd['drop.%d.log_thetas' % i].astype(np.float64),
                                              sd['drop.%d.log_var' % i].astype(np.float64), cfg['thr'])
            baked = co * (d_mask * np.exp(sd['drop.%d.log_thetas' % i].astype(np.float64)))[None]
            mask_bits += d_mask.size
        else:
            s = O.sigmoid(sd['drop.%d.mask_values' % i].astype(np.float64))
            d_mask = (s >= cfg['thr']).astype(np.float64)
            baked = co * d_mask[None]
            mask_bits += d_mask.size
        assert np.array_equal(d_mask, g['bake.d_mask.%d' % i].astype(np.float64))
        assert relerr(baked, g['bake.feature_grid.%d' % i]) < 1e-6
        zeros += baked.size - np.count_nonzero(g['bake.feature_grid.%d' % i])
    assert abs((zeros - mask_bits / 32.0) - float(g['bake.zeros'])) < 1e-3


def test_smallify_tracker():
    g = load('smallify_tracker')
    mom, thr = g['momentum_threshold']
    betas = g['betas0'].astype(np.float64).copy()
    ema = np.sign(betas)
    emavar = np.zeros_like(betas)
    for step in range(12):
        ema, emavar = O.smallify_tracker_step(betas, ema, emavar, mom)
        betas[g['flip.%d' % step]] *= -1.0
    assert relerr(ema, g['EMA']) < 1e-5
    assert relerr(emavar, g['EMAVar']) < 1e-5
    assert np.array_equal(O.smallify_prune_mask(g['EMAVar'], thr), g['mask'])


def test_dataset_sampler_and_ground_truth():
    g = load('dataset')
    for tag in ('a', 'b'):
        vol = O.normalize_volume(g[tag + '.vol_raw'])
        assert np.array_equal(vol, g[tag + '.vol'])  # bit-exact float32
        shape = vol.shape
        max_idx, scales = O.dataset_constants(shape)
        assert np.array_equal(max_idx, g[tag + '.max_idx'])
        assert np.array_equal(scales, g[tag + '.scales'])
        assert int(np.prod(np.asarray(shape, dtype=np.float32))) == int(g[tag + '.n_voxels'][0])
        raw, norm = O.sample_positions(g[tag + '.idx'], shape)
        assert np.array_equal(raw, g[tag + '.raw'])
        assert np.array_equal(norm, g[tag + '.norm'])  # bit-exact float32
        res = np.asarray(shape, dtype=np.float32)
        zero = np.zeros(3, np.float32)
        gt = O.trilinear_lookup(raw, vol, zero, max_idx, res)
        assert np.array_equal(gt, g[tag + '.gt_int'])
        i, j, k = np.unravel_index(g[tag + '.idx'], shape)
        assert np.array_equal(gt, vol[i, j, k])  # integer positions are an exact voxel lookup
        gtf = O.trilinear_lookup(g[tag + '.pf'], vol, zero, max_idx, res)
        assert np.abs(gtf - g[tag + '.gt_float']).max() <= 1e-6
        gtb = O.trilinear_lookup(g[tag + '.pb'], vol, g[tag + '.bb'][0], g[tag + '.bb'][1], res)
        assert np.abs(gtb - g[tag + '.gt_bb']).max() <= 1e-6


def test_reconstruction_coords_and_volume():
    g = load('reconstruct')
    vol = g['vol']
    coords = O.reconstruction_coords(vol.shape, 32)
    for a, sl in enumerate(((slice(None), 0, 0), (0, slice(None), 0), (0, 0, slice(None)))):
        mine = coords[sl + (a,)]
        assert np.abs(mine - g['axis%d' % a]).max() <= 1.2e-7  # <= 1 ulp of the reference's linspace (see oracle doc)
    spec = O.Spec(4, 15, 32, 4, 2, 'db2', '')
    y = O.model_forward(state(g), spec, coords.reshape(-1, 3), training=False, clamp=True)
    full = y.reshape(vol.shape)
    assert np.abs(full - g['full']).max() < 2e-5
    psnr, l1, mse, rmse = O.deviation_statistics(g['full'], vol)
    for got, want in zip((psnr, l1, mse, rmse), g['stats']):
        assert abs(got - want) <= 1e-5 * abs(want)


def test_variance_model():
    g = load('variance_model')
    W = [g['sd.net_layers.%d.weight' % i].astype(np.float64) for i in range(4)] + [g['sd.final_layer.weight'].astype(np.float64)]
    b = [g['sd.net_layers.%d.bias' % i].astype(np.float64) for i in range(4)] + [g['sd.final_layer.bias'].astype(np.float64)]
    y, hs = O.variance_model_forward(g['x'].astype(np.float64), W, b, keep=True)
    assert relerr(y, g['y']) < 1e-5
    _, gW, gb = O.variance_model_backward(g['w'].astype(np.float64), hs, W)
    for i in range(4):
        assert relerr(gW[i], g['grad.net_layers.%d.weight' % i]) < 1e-5
        assert relerr(gb[i], g['grad.net_layers.%d.bias' % i]) < 1e-5
    assert relerr(gW[4], g['grad.final_layer.weight']) < 1e-5


@pytest.mark.parametrize('tag', ['basic', 'smallify'])
def test_training_trajectory(tag):
    """A dozen Adam steps of the reference step loop (training/training.py:89-138) on fixed batches."""
    g = load('trajectory_' + tag)
    mask = 'smallify' if tag == 'smallify' else ''
    spec = O.Spec(8, 15, 32, 4, 2, 'db2', mask, 0.9)
    sd = {k: v.astype(np.float64) for k, v in state(g, 'sd0.').items() if not k.startswith('filter.')}
    vol = g['vol']
    m = {k: np.zeros_like(v) for k, v in sd.items()}
    v = {k: np.zeros_like(v) for k, v in sd.items()}
    losses = []
    zero = np.zeros(3, np.float32)
    max_idx, _ = O.dataset_constants(vol.shape)
    for s in range(g['idx'].shape[0]):
        raw, norm = O.sample_positions(g['idx'][s], vol.shape)
        gt = O.trilinear_lookup(raw, vol, zero, max_idx, np.asarray(vol.shape, np.float32)).astype(np.float64)
        y, ctx = O.model_forward(sd, spec, norm, training=True, keep=True)
        N = y.shape[0]
        loss = np.mean((y[:, 0] - gt) ** 2)
        grads = O.model_backward((2.0 / N) * (y - gt[:, None]), ctx, spec)
        if mask:
            lv, lg = O.smallify_loss(sd, spec, 1e-4, 1e-5)
            loss += lv
            for k, gv in lg.items():
                grads[k] = grads[k] + gv
        losses.append(loss)
        for k in sd:
            sd[k], m[k], v[k] = O.adam_step(sd[k], grads[k], m[k], v[k], s + 1, 0.008)
    assert np.abs(np.asarray(losses) - g['losses']).max() < 2e-4 * np.abs(g['losses']).max()
    for k in sd:
        ref = g['sd1.' + k]
        assert np.abs(sd[k] - ref).max() < 2e-3 * max(np.abs(ref).max(), 1e-6), k
