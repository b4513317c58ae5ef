"""CPU oracle: a plain numpy restatement of the reference's fV-SRN hot path.

TEST INFRASTRUCTURE ONLY -- this file is the *checker*, never the product.  Only
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl
reference`` legs of ``bench.py`` may import it.  The product package
(``latent_feature_grid_compression_b200``) never imports anything under
``oracle/`` and has no CPU fallback.

Parity status: **pinned**.  The reference has no golden vectors or asserting
tests of its own, so the oracle is pinned against outputs
of the reference itself: ``tests/golden/make_golden.py`` runs the unmodified
reference modules in the build container and commits their outputs as
``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks every function
below against those fixtures.

All ``file:line`` citations are relative to the reference repository root.
Everything is written for clarity (direct sums, explicit loops over filter
taps), works in float64 by default (``dtype=np.float32`` mimics the reference's
precision) and is independent of torch.
"""
from __future__ import annotations

import math

import numpy as np

# ---------------------------------------------------------------------------
# wavelet filter taps (third-party arithmetic: PyWavelets 1.4.1, Env.txt:178,
# not vendored in the reference; only the published 1-D taps are needed)
# ---------------------------------------------------------------------------


def wavelet_taps(name: str):
    """``pywt.Wavelet(name).filter_bank`` = (dec_lo, dec_hi, rec_lo, rec_hi).

    Used at wavelet_transform/Torch_Wavelet_Transform.py:41.  The reference
    stores the taps as float32 buffers (``torch.tensor(python floats)`` is
    float32), so they are rounded to float32 here as well.
    """
    s2, s3 = math.sqrt(2.0), math.sqrt(3.0)
    if name in ("haar", "db1"):
        rec_lo = [1 / s2, 1 / s2]
    elif name == "db2":
        rec_lo = [(1 + s3) / (4 * s2), (3 + s3) / (4 * s2), (3 - s3) / (4 * s2), (1 - s3) / (4 * s2)]
    else:
        raise ValueError("oracle knows haar/db1/db2, got %r" % (name,))
    dec_lo = rec_lo[::-1]
    rec_hi = [((-1) ** i) * dec_lo[i] for i in range(len(dec_lo))]
    dec_hi = rec_hi[::-1]
    f32 = lambda v: np.asarray(v, dtype=np.float32).astype(np.float64)  # noqa: E731
    return f32(dec_lo), f32(dec_hi), f32(rec_lo), f32(rec_hi)


def dwt_max_level(n: int, filter_len: int) -> int:
    """``pywt.dwt_max_level`` as used at model/Feature_Grid_Model.py:85."""
    if n < filter_len - 1:
        return 0
    return int(math.floor(math.log2(n / (filter_len - 1.0))))


# ---------------------------------------------------------------------------
# 3-D wavelet analysis / synthesis (wavelet_transform/Torch_Wavelet_Transform.py)
# ---------------------------------------------------------------------------

def _band_filters(lo, hi):
    """Sub-band k = 4a+2b+c uses filter a along D, b along H, c along W
    (Torch_Wavelet_Transform.py:44-57, outer-product construction)."""
    return [(a, b, c) for a in (0, 1) for b in (0, 1) for c in (0, 1)], (lo, hi)


def dwt_level(x, wavelet: str):
    """One analysis level: ``_WaveletFilterNd.encode`` (Torch_Wavelet_Transform.py:75-89).

    x: (C, d0, d1, d2).  Zero-pad (2L-3)//2 on the left and the same (+1 when
    the extent is odd) on the right (:59-67), correlate with the *flipped*
    decomposition filters (:56) at stride 2.  Returns coeffs (C, 8, e0, e1, e2)
    and the input shape.
    """
    dec_lo, dec_hi, _, _ = wavelet_taps(wavelet)
    L = len(dec_lo)
    f = (dec_lo[::-1], dec_hi[::-1])  # filter_fwd is built from flipped dec filters
    C = x.shape[0]
    shape = np.asarray(x.shape[1:])
    pad = (2 * L - 3) // 2
    pads = [(0, 0)] + [(pad, pad + int(s % 2 == 1)) for s in shape]
    xp = np.pad(x, pads)
    out_sz = [(xp.shape[1 + a] - L) // 2 + 1 for a in range(3)]
    out = np.zeros((C, 8, *out_sz), dtype=x.dtype)
    for k, (a, b, c) in enumerate(_band_filters(None, None)[0]):
        for ta in range(L):
            for tb in range(L):
                for tc in range(L):
                    w = f[a][ta] * f[b][tb] * f[c][tc]
                    out[:, k] += w * xp[:, ta:ta + 2 * out_sz[0]:2, tb:tb + 2 * out_sz[1]:2, tc:tc + 2 * out_sz[2]:2]
    return out, shape


def idwt_level(low, high, target_shape, wavelet: str):
    """One synthesis level: ``_WaveletFilterNd.decode`` (Torch_Wavelet_Transform.py:91-104).

    low: (C, d, d, d) running low-pass (sub-band 0), high: (C, 7, d, d, d) stored
    high bands (sub-bands 1..7).  Grouped transposed conv, stride 2, with the
    reconstruction filters gives the uncropped ``u[o] = sum_k sum_i c_k[i] *
    r_a[oz-2iz] r_b[oy-2iy] r_c[ox-2ix]`` of size 2d+L-2; the crop (:69-73)
    keeps ``u[floor(delta/2) : floor(delta/2)+target]``.
    """
    _, _, rec_lo, rec_hi = wavelet_taps(wavelet)
    r = (rec_lo, rec_hi)
    L = len(rec_lo)
    C = low.shape[0]
    d = low.shape[1:]
    full = [2 * s + L - 2 for s in d]
    u = np.zeros((C, *full), dtype=low.dtype)
    for k, (a, b, c) in enumerate(_band_filters(None, None)[0]):
        ck = low if k == 0 else high[:, k - 1]
        for ta in range(L):
            for tb in range(L):
                for tc in range(L):
                    w = r[a][ta] * r[b][tb] * r[c][tc]
                    u[:, ta:ta + 2 * d[0]:2, tb:tb + 2 * d[1]:2, tc:tc + 2 * d[2]:2] += w * ck
    off = [int(math.floor((full[a] - int(target_shape[a])) / 2)) for a in range(3)]
    t = [int(s) for s in target_shape]
    return u[:, off[0]:off[0] + t[0], off[1]:off[1] + t[1], off[2]:off[2] + t[2]]


def idwt_level_adjoint(g_out, d, wavelet: str):
    """Adjoint of :func:`idwt_level`: gradient w.r.t. (low, high) given d(out)."""
    _, _, rec_lo, rec_hi = wavelet_taps(wavelet)
    r = (rec_lo, rec_hi)
    L = len(rec_lo)
    C = g_out.shape[0]
    full = [2 * s + L - 2 for s in d]
    t = g_out.shape[1:]
    off = [int(math.floor((full[a] - int(t[a])) / 2)) for a in range(3)]
    gu = np.zeros((C, *full), dtype=g_out.dtype)
    gu[:, off[0]:off[0] + t[0], off[1]:off[1] + t[1], off[2]:off[2] + t[2]] = g_out
    g = np.zeros((C, 8, *d), dtype=g_out.dtype)
    for k, (a, b, c) in enumerate(_band_filters(None, None)[0]):
        for ta in range(L):
            for tb in range(L):
                for tc in range(L):
                    w = r[a][ta] * r[b][tb] * r[c][tc]
                    g[:, k] += w * gu[:, ta:ta + 2 * d[0]:2, tb:tb + 2 * d[1]:2, tc:tc + 2 * d[2]:2]
    return g[:, 0], g[:, 1:]


def encode_volume(grid, wavelet: str):
    """``Feature_Grid_Model.encode_volume`` (model/Feature_Grid_Model.py:83-99).

    Returns ([LLL_coarsest, high_coarsest, ..., high_finest], shape_array coarse->fine).
    """
    L = len(wavelet_taps(wavelet)[0])
    levels = min(dwt_max_level(s, L) for s in grid.shape[-3:])
    feats, shapes = [], []
    data = grid
    for _ in range(levels):
        co, shape = dwt_level(data, wavelet)
        feats.append(co[:, 1:])
        shapes.append(shape)
        data = co[:, 0]
    return [data] + feats[::-1], np.asarray(shapes[::-1], dtype=np.int64).reshape(-1, 3)


def decode_volume(coeffs, mults, shape_array, wavelet: str):
    """``Feature_Grid_Model.decode_volume`` (model/Feature_Grid_Model.py:102-108).

    ``mults[l]`` is the mask multiplier of level l (shape = coeffs[l].shape[1:],
    broadcast over the channel dim) or None for ``nn.Identity``.
    """
    def app(c, m):
        return c if m is None else c * m[None]
    restored = app(coeffs[0], mults[0])
    for hi, m, shape in zip(coeffs[1:], mults[1:], shape_array):
        restored = idwt_level(restored, app(hi, m), shape, wavelet)
    return restored


def decode_volume_adjoint(g_grid, coeffs, mults, shape_array, wavelet: str):
    """Gradients of ``decode_volume`` w.r.t. the coefficients and the multipliers.

    d coeff = g * mult;  d mult = sum_c coeff[c] * g[c]  (mask broadcast over channels).
    """
    n = len(coeffs)
    g_masked = [None] * n  # gradient w.r.t. (coeff * mult)
    g = g_grid
    for l in range(n - 1, 0, -1):
        d = coeffs[l].shape[2:]
        g, g_hi = idwt_level_adjoint(g, d, wavelet)
        g_masked[l] = g_hi
    g_masked[0] = g
    g_coeff, g_mult = [], []
    for l in range(n):
        m = mults[l]
        if m is None:
            g_coeff.append(g_masked[l])
            g_mult.append(None)
        else:
            g_coeff.append(g_masked[l] * m[None])
            g_mult.append((g_masked[l] * coeffs[l]).sum(axis=0))
    return g_coeff, g_mult


# ---------------------------------------------------------------------------
# mask layers
# ---------------------------------------------------------------------------

def sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def mask_multiplier(kind, params, noise=None, d_mask=None, threshold=0.5, training=True):
    """Value multiplier m and the derivative factors needed by the backward pass.

    Returns (m, dm) where ``dm`` maps parameter name -> d m / d param (elementwise),
    or (None, {}) when the layer is an identity in this mode.

    * ``smallify``  Smallify_Dropout.py:54-61: train & no d_mask -> m = betas; train & d_mask -> m = d_mask;
      eval -> identity.
    * ``variational`` Variational_Dropout_Layer.py:101-113: m = exp(log_thetas) + exp(log_var/2)*xi in train AND
      eval; d_mask once baked.
    * ``masked_ste`` Straight_Through_Dropout.py:53-61: value x*[sigmoid(v) >= t], gradient as if x*sigmoid(v).
    * ``bernoulli`` Straight_Through_Dropout.py:26-30: m = [u < v] (bool, no gradient path to v).
    """
    if kind in (None, '', 'identity'):
        return None, {}
    if kind == 'smallify':
        if not training:
            return None, {}
        if d_mask is not None:
            return d_mask, {}
        b = params['betas']
        return b, {'betas': np.ones_like(b)}
    if kind == 'variational':
        if d_mask is not None:
            return d_mask, {}
        th = np.exp(params['log_thetas'])
        sg = np.exp(params['log_var'] / 2.0)
        m = th + sg * noise
        return m, {'log_thetas': th, 'log_var': 0.5 * sg * noise}
    if kind == 'masked_ste':
        if not training:
            return None, {}
        if d_mask is not None:
            return d_mask.astype(params['mask_values'].dtype), {}
        s = sigmoid(params['mask_values'])
        m = (s >= threshold).astype(s.dtype)
        return m, {'mask_values': s * (1 - s), '_ste_grad_mult': s}
    if kind == 'bernoulli':
        if not training:
            return None, {}
        m = (noise < params['mask_values']).astype(params['mask_values'].dtype)
        return m, {}
    raise ValueError(kind)


def smallify_tracker_step(betas, ema, emavar, momentum):
    """``sign_variance_pruning_onlyVar`` (Smallify_Dropout.py:103-112)."""
    phi = np.sign(betas) - ema
    ema = ema + momentum * phi
    emavar = (1.0 - momentum) * (emavar + momentum * phi ** 2)
    return ema, emavar


def smallify_prune_mask(emavar, threshold):
    """Smallify_Dropout.py:114-118."""
    return np.where(emavar < threshold, 1.0, 0.0)


def variational_dkl(log_thetas, log_var):
    """``VariationalDropout.calculate_Dkl`` (Variational_Dropout_Layer.py:115-122)."""
    k1, k2, k3 = 0.63576, 1.87320, 1.48695
    la = log_var - 2.0 * log_thetas
    t1 = k1 * sigmoid(k2 + k3 * la)
    t2 = 0.5 * np.logaddexp(0.0, -la)
    return np.sum(-t1 + t2 + k1)


def variational_droprate(log_thetas, log_var):
    """Variational_Dropout_Layer.py:89-95."""
    a = np.exp(log_var - 2.0 * log_thetas)
    return a / (1.0 + a)


def variational_prune_mask(log_thetas, log_var, threshold):
    """Variational_Dropout_Layer.py:137-147 (keeps element 0 when nothing would be pruned)."""
    m = np.where(variational_droprate(log_thetas, log_var) < threshold, 1.0, 0.0)
    if m.size - np.count_nonzero(m) == 0:
        m[0] = 1.0
    return m


# ---------------------------------------------------------------------------
# per-sample path: trilinear gather, Fourier features, SnakeAlt MLP
# ---------------------------------------------------------------------------

def grid_sample_corners(coords, G, dtype=np.float64):
    """Corner indices/weights of ``F.grid_sample(mode='bilinear', padding_mode='zeros',
    align_corners=False)`` as called at model/Feature_Grid_Model.py:62-64.

    coords[:, 0] indexes the LAST grid dim (x = W), [:, 1] -> H, [:, 2] -> D.
    Returns idx (N, 8, 3) as (z, y, x), w (N, 8), valid (N, 8).
    """
    c = coords.astype(dtype)
    one, two = dtype(1.0), dtype(2.0)
    f = ((c + one) * dtype(G) - one) / two  # unnormalise, align_corners=False
    f0 = np.floor(f)
    t = f - f0
    i0 = f0.astype(np.int64)
    idx = np.zeros((c.shape[0], 8, 3), dtype=np.int64)
    w = np.ones((c.shape[0], 8), dtype=dtype)
    for corner in range(8):
        dz, dy, dx = (corner >> 2) & 1, (corner >> 1) & 1, corner & 1
        for axis, (dim, dd) in enumerate(((2, dz), (1, dy), (0, dx))):
            idx[:, corner, axis] = i0[:, dim] + dd
            w[:, corner] = w[:, corner] * (t[:, dim] if dd else (one - t[:, dim]))
    valid = np.all((idx >= 0) & (idx < G), axis=2)
    return idx, w, valid


def grid_sample(grid, coords, dtype=np.float64):
    """(C, G, G, G) x (N, 3) -> (N, C)."""
    C, G = grid.shape[0], grid.shape[1]
    idx, w, valid = grid_sample_corners(coords, G, dtype)
    out = np.zeros((coords.shape[0], C), dtype=dtype)
    ic = np.clip(idx, 0, G - 1)
    for corner in range(8):
        v = grid[:, ic[:, corner, 0], ic[:, corner, 1], ic[:, corner, 2]].T  # (N, C)
        out += (w[:, corner] * valid[:, corner])[:, None] * v
    return out


def grid_sample_adjoint(g_feat, coords, C, G, dtype=np.float64):
    """Scatter-add of d(features) (N, C) into d(grid) (C, G, G, G)."""
    idx, w, valid = grid_sample_corners(coords, G, dtype)
    gg = np.zeros((C, G, G, G), dtype=dtype)
    ic = np.clip(idx, 0, G - 1)
    for corner in range(8):
        contrib = (w[:, corner] * valid[:, corner])[:, None] * g_feat  # (N, C)
        np.add.at(gg, (slice(None), ic[:, corner, 0], ic[:, corner, 1], ic[:, corner, 2]), contrib.T)
    return gg


def fourier_embed(coords, n_freqs, dtype=np.float64):
    """``FourierEmbedding`` (model/Feature_Embedding.py:20-34): for k < F: [sin(x*w_k) (3), cos(x*w_k) (3)],
    w_k = float32(2^k) * 2 * pi rounded to float32; the argument x*w_k is rounded to float32 first."""
    freqs = (np.float32(2.0) ** np.linspace(0.0, n_freqs - 1, n_freqs, dtype=np.float32))
    freqs = (freqs * np.float32(2.0) * np.float32(np.pi)).astype(np.float32)
    outs = []
    for fr in freqs:
        arg = (coords.astype(np.float32) * fr).astype(dtype)  # float32 product, as in the reference
        outs.append(np.sin(arg))
        outs.append(np.cos(arg))
    return np.concatenate(outs, axis=-1) if outs else np.zeros((coords.shape[0], 0), dtype=dtype)


def fourier_freqs(n_freqs):
    f = (np.float32(2.0) ** np.linspace(0.0, n_freqs - 1, n_freqs, dtype=np.float32))
    return (f * np.float32(2.0) * np.float32(np.pi)).astype(np.float32)


def snake(x):
    """``SnakeAlt`` (model/Feature_Grid_Model.py:12-13)."""
    return 0.5 * x + np.sin(x) ** 2


def snake_grad(x):
    return 0.5 + np.sin(2.0 * x)


def mlp_forward(x, weights, biases, keep=False):
    """model/Feature_Grid_Model.py:72-75: L x SnakeAlt(Linear) + final Linear; weights are (out, in)."""
    zs, hs = [], [x]
    h = x
    for W, b in zip(weights[:-1], biases[:-1]):
        z = h @ W.T + b
        h = snake(z)
        zs.append(z)
        hs.append(h)
    y = h @ weights[-1].T + biases[-1]
    return (y, zs, hs) if keep else y


def mlp_backward(g_y, zs, hs, weights):
    """Gradients of :func:`mlp_forward`; returns (g_x, g_weights, g_biases)."""
    gW = [None] * len(weights)
    gb = [None] * len(weights)
    gW[-1] = g_y.T @ hs[-1]
    gb[-1] = g_y.sum(axis=0)
    gh = g_y @ weights[-1]
    for l in range(len(weights) - 2, -1, -1):
        gz = gh * snake_grad(zs[l])
        gW[l] = gz.T @ hs[l]
        gb[l] = gz.sum(axis=0)
        gh = gz @ weights[l]
    return gh, gW, gb


# ---------------------------------------------------------------------------
# whole model (state-dict keyed, reference names)
# ---------------------------------------------------------------------------

_MASK_PARAM_NAMES = {'smallify': ('betas',), 'variational': ('log_thetas', 'log_var'),
                     'masked_ste': ('mask_values',), 'bernoulli': ('mask_values',)}


class Spec:
    """Static description of a model instance (what ``setup_model`` fixes, model/model_utils.py:23-59)."""

    def __init__(self, C, G, H=32, L=4, F=2, wavelet='db2', mask='', threshold=0.5, shape_array=None):
        self.C, self.G, self.H, self.L, self.F = C, G, H, L, F
        self.wavelet = wavelet
        self.mask = mask
        self.threshold = threshold
        if shape_array is None:
            Lf = len(wavelet_taps(wavelet)[0])
            sizes = []
            s = G
            for _ in range(dwt_max_level(G, Lf)):
                sizes.append([s, s, s])
                pad = (2 * Lf - 3) // 2
                s = (s + 2 * pad + (s % 2) - Lf) // 2 + 1
            shape_array = np.asarray(sizes[::-1], dtype=np.int64).reshape(-1, 3)
        self.shape_array = np.asarray(shape_array, dtype=np.int64).reshape(-1, 3)
        self.n_levels = self.shape_array.shape[0] + 1  # number of coefficient tensors


def split_state(sd, spec: Spec, dtype=np.float64):
    coeffs = [np.asarray(sd['feature_grid.%d' % i], dtype=dtype) for i in range(spec.n_levels)]
    weights = [np.asarray(sd['net_layers.%d.weight' % i], dtype=dtype) for i in range(spec.L)]
    biases = [np.asarray(sd['net_layers.%d.bias' % i], dtype=dtype) for i in range(spec.L)]
    weights.append(np.asarray(sd['final_layer.weight'], dtype=dtype))
    biases.append(np.asarray(sd['final_layer.bias'], dtype=dtype))
    masks = []
    for i in range(spec.n_levels):
        names = _MASK_PARAM_NAMES.get(spec.mask, ())
        masks.append({n: np.asarray(sd['drop.%d.%s' % (i, n)], dtype=dtype) for n in names})
    return coeffs, masks, weights, biases


def model_forward(sd, spec: Spec, coords, noise=None, d_masks=None, training=True, dtype=np.float64,
                  keep=False, clamp=False):
    """``Feature_Grid_Model.forward`` (model/Feature_Grid_Model.py:50-80)."""
    coeffs, masks, weights, biases = split_state(sd, spec, dtype)
    mults, dms = [], []
    for i in range(spec.n_levels):
        m, dm = mask_multiplier(spec.mask, masks[i],
                                noise=None if noise is None else np.asarray(noise[i], dtype=dtype),
                                d_mask=None if d_masks is None else d_masks[i],
                                threshold=spec.threshold, training=training)
        mults.append(m)
        dms.append(dm)
    grid = decode_volume(coeffs, mults, spec.shape_array, spec.wavelet)
    c = np.asarray(coords, dtype=np.float32).reshape(-1, 3)
    feats = grid_sample(grid, c, dtype)
    x = np.concatenate([c.astype(dtype), fourier_embed(c, spec.F, dtype), feats], axis=-1)
    y, zs, hs = mlp_forward(x, weights, biases, keep=True)
    if clamp:
        y = np.clip(y, -1.0, 1.0)
    if keep:
        return y, dict(grid=grid, coeffs=coeffs, masks=masks, mults=mults, dms=dms, zs=zs, hs=hs,
                       weights=weights, coords=c)
    return y


def model_backward(g_y, ctx, spec: Spec, dtype=np.float64):
    """Gradient of ``sum(y * g_y)`` w.r.t. every parameter, keyed by state-dict name.
    The masked straight-through estimator is handled exactly: the forward value uses the hard mask, the
    backward uses d/dx and d/dv of x*sigmoid(v) (Straight_Through_Dropout.py:58)."""
    g_x, gW, gb = mlp_backward(np.asarray(g_y, dtype=dtype), ctx['zs'], ctx['hs'], ctx['weights'])
    n_in0 = 3 + 6 * spec.F
    g_feat = g_x[:, n_in0:]
    g_grid = grid_sample_adjoint(g_feat, ctx['coords'], spec.C, ctx['grid'].shape[1], dtype)
    # gradient w.r.t. the masked coefficients (coeff*mult), level by level
    n = spec.n_levels
    g_masked = [None] * n
    g = g_grid
    for l in range(n - 1, 0, -1):
        d = ctx['coeffs'][l].shape[2:]
        g, g_hi = idwt_level_adjoint(g, d, spec.wavelet)
        g_masked[l] = g_hi
    g_masked[0] = g
    out = {}
    for i in range(n):
        m = ctx['mults'][i]
        dm = ctx['dms'][i]
        co = ctx['coeffs'][i]
        if m is None:
            out['feature_grid.%d' % i] = g_masked[i]
            continue
        if '_ste_grad_mult' in dm:
            soft = dm['_ste_grad_mult']
            out['feature_grid.%d' % i] = g_masked[i] * soft[None]
            out['drop.%d.mask_values' % i] = (g_masked[i] * co).sum(axis=0) * dm['mask_values']
            continue
        out['feature_grid.%d' % i] = g_masked[i] * m[None]
        gm = (g_masked[i] * co).sum(axis=0)
        for pname, fac in dm.items():
            out['drop.%d.%s' % (i, pname)] = gm * fac
    for l in range(spec.L):
        out['net_layers.%d.weight' % l] = gW[l]
        out['net_layers.%d.bias' % l] = gb[l]
    out['final_layer.weight'] = gW[-1]
    out['final_layer.bias'] = gb[-1]
    out['grid'] = g_grid
    out['x'] = g_x
    return out


# ---------------------------------------------------------------------------
# regulariser losses
# ---------------------------------------------------------------------------

def smallify_loss(sd, spec: Spec, weight_l1, weight_l2, dtype=np.float64):
    """``SmallifyLoss`` (model/Smallify_Dropout.py:10-40): w1 * sum|mask param| + w2 * sum coeff^2.
    Returns (value, grads keyed by state-dict name)."""
    coeffs, masks, _, _ = split_state(sd, spec, dtype)
    val = 0.0
    grads = {}
    pname = {'smallify': 'betas', 'masked_ste': 'mask_values', 'bernoulli': 'mask_values'}.get(spec.mask)
    if weight_l1 > 0 and pname is not None:
        for i, m in enumerate(masks):
            val += weight_l1 * np.abs(m[pname]).sum()
            grads['drop.%d.%s' % (i, pname)] = weight_l1 * np.sign(m[pname])
    if weight_l2 > 0:
        for i, c in enumerate(coeffs):
            val += weight_l2 * (c ** 2).sum()
            grads['feature_grid.%d' % i] = 2.0 * weight_l2 * c
    return val, grads


def variational_loss(sd, spec: Spec, pred, gt, log_sigma, n_voxels, batch, weight_dkl, weight_weights,
                     weight_dkl_multiplier, weight_dkl_max=30.0, dtype=np.float64):
    """``VariationalDropoutLoss.forward`` (model/Variational_Dropout_Layer.py:54-69) with
    ``calculate_Log_Likelihood_variance`` (:24-30).  Returns dict of the five reference outputs plus the updated
    weight_dkl and d loss / d pred, d loss / d log_sigma."""
    coeffs, masks, _, _ = split_state(sd, spec, dtype)
    scale = n_voxels / batch
    if weight_dkl < weight_dkl_max:
        weight_dkl = weight_dkl * (1.0 + weight_dkl_multiplier)
    pred = np.asarray(pred, dtype=dtype)
    gt = np.asarray(gt, dtype=dtype)
    v = np.asarray(log_sigma, dtype=dtype)
    err2 = (gt - pred) ** 2
    a = 1.0 / (2.0 * np.exp(v) ** 2)
    ll_el = a * (-err2) - (math.log(2 * math.pi) + 2 * v) / 2
    mse = err2.sum() / pred.shape[0]
    ll = ll_el.sum() * scale
    dkl = weight_dkl * sum(variational_dkl(m['log_thetas'], m['log_var']) for m in masks) * scale
    wsum = weight_weights * sum((c ** 2).sum() for c in coeffs) * scale
    loss = -(ll - dkl - wsum)
    g_pred = -scale * (2.0 * a * (gt - pred))
    g_v = -scale * (2.0 * a * err2 - 1.0)
    return dict(loss=loss, ll=ll, mse=mse, dkl=dkl, wsum=wsum, weight_dkl=weight_dkl, g_pred=g_pred, g_logsig=g_v)


# ---------------------------------------------------------------------------
# sampler / ground truth / reconstruction
# ---------------------------------------------------------------------------

def normalize_volume(vol):
    """``normalize_volume(volume, min, max, -1, 1)`` in float32 (data/IndexDataset.py:7-8,15-17)."""
    v = np.asarray(vol, dtype=np.float32)
    mn, mx = v.min(), v.max()
    return (np.float32(2.0) * ((v - mn) / (mx - mn)) + np.float32(-1.0)).astype(np.float32)


def dataset_constants(vol_shape):
    """``IndexDataset.__init__`` (data/IndexDataset.py:52-65): max_idx, scales (float32)."""
    res = np.asarray(vol_shape, dtype=np.float32)
    max_idx = res - np.float32(1.0)
    scales = (max_idx / max_idx.max()).astype(np.float32)
    return max_idx, scales


def sample_positions(flat_idx, vol_shape):
    """``IndexDataset.__getitem__`` (data/IndexDataset.py:90-96) for given flat voxel indices:
    raw = (i, j, k) as float32; norm = scales * (2 * raw / max_idx - 1), float32 op order as in the reference."""
    max_idx, scales = dataset_constants(vol_shape)
    i, j, k = np.unravel_index(np.asarray(flat_idx, dtype=np.int64), vol_shape)
    raw = np.stack([i, j, k], axis=-1).astype(np.float32)
    # normalize_volume(raw, min=0, max=max_idx, -1, 1) = (1 - -1) * ((raw - 0) / (max_idx - 0)) + -1
    norm = np.float32(2.0) * ((raw - np.float32(0.0)) / (max_idx - np.float32(0.0))) + np.float32(-1.0)
    norm = (scales[None] * norm).astype(np.float32)
    return raw, norm


def trilinear_lookup(p, f, min_bb, max_bb, res):
    """``trilinear_f_interpolation`` (data/Interpolation.py:8-44): float32 lattice coordinates, float64 alphas,
    f[x, y, z] indexing (p[:, 0] -> dim 0), lerp order x -> y -> z in float32."""
    p = np.asarray(p, dtype=np.float32)
    f = np.asarray(f, dtype=np.float32)
    min_bb = np.asarray(min_bb, dtype=np.float32)
    max_bb = np.asarray(max_bb, dtype=np.float32)
    res = np.asarray(res, dtype=np.float32)
    normalized = ((p - min_bb[None]) / (max_bb - min_bb)[None]) * (res[None] - np.float32(1.0))
    lo = np.floor(normalized).astype(np.int64)
    hi = np.ceil(normalized).astype(np.int64)
    diff = np.maximum((hi - lo).astype(np.float64), 1e-12)
    alpha = ((normalized.astype(np.float64) - lo.astype(np.float64)) / diff).astype(np.float32)
    one_alpha = np.float32(1.0) - alpha

    def at(ix, iy, iz):
        return f[ix, iy, iz]
    x00 = one_alpha[:, 0] * at(lo[:, 0], lo[:, 1], lo[:, 2]) + alpha[:, 0] * at(hi[:, 0], lo[:, 1], lo[:, 2])
    x10 = one_alpha[:, 0] * at(lo[:, 0], hi[:, 1], lo[:, 2]) + alpha[:, 0] * at(hi[:, 0], hi[:, 1], lo[:, 2])
    x01 = one_alpha[:, 0] * at(lo[:, 0], lo[:, 1], hi[:, 2]) + alpha[:, 0] * at(hi[:, 0], lo[:, 1], hi[:, 2])
    x11 = one_alpha[:, 0] * at(lo[:, 0], hi[:, 1], hi[:, 2]) + alpha[:, 0] * at(hi[:, 0], hi[:, 1], hi[:, 2])
    y0 = one_alpha[:, 1] * x00 + alpha[:, 1] * x10
    y1 = one_alpha[:, 1] * x01 + alpha[:, 1] * x11
    return (one_alpha[:, 2] * y0 + alpha[:, 2] * y1).astype(np.float32)


def reconstruction_coords(vol_shape, tiled_res=32):
    """Normalised coordinates of every voxel as ``field_from_net`` builds them (visualization/OutputToVTK.py:11-37):
    per tile ``linspace(b/(R-1), (e-1)/(R-1), e-b) * 2 - 1``, times ``scales``.  Returns (R0, R1, R2, 3) float32."""
    max_idx, scales = dataset_constants(vol_shape)
    axes = []
    for a, R in enumerate(vol_shape):
        v = np.zeros(R, dtype=np.float32)
        for b in range(0, R, tiled_res):
            e = min(b + tiled_res, R)
            # min_bounds = 0 + (b/(R-1)) * (max_idx - 0); start = min_bounds / (max_idx - 0)  (all float32)
            start = np.float32(np.float32(np.float32(b / (R - 1)) * max_idx[a]) / max_idx[a])
            end = np.float32(np.float32(np.float32((e - 1) / (R - 1)) * max_idx[a]) / max_idx[a])
            v[b:e] = _torch_linspace_f32(start, end, e - b)
        axes.append((np.float32(2.0) * v - np.float32(1.0)) * scales[a])
    out = np.zeros((*vol_shape, 3), dtype=np.float32)
    out[..., 0] = axes[0][:, None, None]
    out[..., 1] = axes[1][None, :, None]
    out[..., 2] = axes[2][None, None, :]
    return out


def _torch_linspace_f32(start, end, steps):
    """torch.linspace in float32, scalar form (ATen RangeFactories: step = (end-start)/(steps-1); first half
    start + step*i, second half end - step*(n-1-i)).  ATen's vectorised kernel re-associates the additions, so the
    reference's values can differ from this by one ulp; the golden fixture (reconstruct.npz axis0/1/2) holds the
    reference's exact coordinates and the tests bound the difference."""
    start, end = np.float32(start), np.float32(end)
    if steps == 1:
        return np.asarray([start], dtype=np.float32)
    step = np.float32((end - start) / np.float32(steps - 1))
    idx = np.arange(steps)
    half = steps // 2
    lo = (start + step * idx.astype(np.float32)).astype(np.float32)
    hi = (end - step * (steps - 1 - idx).astype(np.float32)).astype(np.float32)
    return np.where(idx < half, lo, hi).astype(np.float32)


def deviation_statistics(pred, gt):
    """``calculate_deviation_statistics`` (visualization/OutputToVTK.py:53-60): psnr, l1, mse, rmse."""
    pred = np.asarray(pred, dtype=np.float64)
    gt = np.asarray(gt, dtype=np.float64)
    diff = gt - pred
    mse = np.mean(diff ** 2)
    psnr = 10.0 * np.log10((gt.max() - gt.min()) ** 2 / mse)
    return psnr, np.mean(np.abs(diff)), mse, np.sqrt(mse)


# ---------------------------------------------------------------------------
# Variance_Model and Adam (used by the whole-step checks)
# ---------------------------------------------------------------------------

def variance_model_forward(x, weights, biases, keep=False):
    """``Variance_Model.forward`` (model/Variational_Dropout_Layer.py:159-175): ReLU MLP 3->32x4->1."""
    hs = [x]
    h = x
    for W, b in zip(weights[:-1], biases[:-1]):
        h = np.maximum(h @ W.T + b, 0.0)
        hs.append(h)
    y = h @ weights[-1].T + biases[-1]
    return (y, hs) if keep else y


def variance_model_backward(g_y, hs, weights):
    gW = [None] * len(weights)
    gb = [None] * len(weights)
    gW[-1] = g_y.T @ hs[-1]
    gb[-1] = g_y.sum(axis=0)
    gh = g_y @ weights[-1]
    for l in range(len(weights) - 2, -1, -1):
        gz = gh * (hs[l + 1] > 0)
        gW[l] = gz.T @ hs[l]
        gb[l] = gz.sum(axis=0)
        gh = gz @ weights[l]
    return gh, gW, gb


def adam_step(p, g, m, v, step, lr, b1=0.9, b2=0.999, eps=1e-8):
    """torch.optim.Adam defaults (training/training.py:199): returns new (p, m, v)."""
    m = b1 * m + (1 - b1) * g
    v = b2 * v + (1 - b2) * g * g
    bc1 = 1 - b1 ** step
    bc2 = 1 - b2 ** step
    denom = np.sqrt(v) / math.sqrt(bc2) + eps
    p = p - (lr / bc1) * m / denom
    return p, m, v
