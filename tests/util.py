"""Shared helpers for the GPU parity tests (golden fixtures + oracle)."""
import contextlib
import os

import numpy as np
import torch

GOLD = os.path.join(os.path.dirname(__file__), 'golden')

MODEL_CASES = {
    'basic_db2_c16_g15': dict(drop='', mask='', wavelet='db2', p=0.1, thr=0.9),
    'basic_haar_c4_g16': dict(drop='', mask='', wavelet='haar', p=0.1, thr=0.9),
    'basic_db2_c8_g17_h64_l3_f3': dict(drop='', mask='', wavelet='db2', p=0.1, thr=0.9),
    'basic_db2_c4_g5_nolevels': dict(drop='', mask='', wavelet='db2', p=0.1, thr=0.9),
    'smallify_db2_c6_g15': dict(drop='smallify', mask='smallify', wavelet='db2', p=0.025, thr=0.75),
    'variational_db2_c8_g15': dict(drop='variational_dynamic', mask='variational', wavelet='db2', p=0.1, thr=0.5),
    'maskedste_db2_c8_g15': dict(drop='masked_straight_through', mask='masked_ste', wavelet='db2', p=0.5, thr=0.6),
    'bernoulli_db2_c8_g15': dict(drop='straight_through', mask='bernoulli', wavelet='db2', p=0.5, thr=0.5),
}


def load(tag):
    return dict(np.load(os.path.join(GOLD, tag + '.npz')))


def state(g, prefix='sd.'):
    return {k[len(prefix):]: v for k, v in g.items() if k.startswith(prefix)}


def relerr(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def build_model(tag, g=None):
    """Our model with the golden state dict loaded, on cuda:0, train mode."""
    from latent_feature_grid_compression_b200.model.model_utils import setup_model
    g = g or load('model_' + tag)
    cfg = MODEL_CASES[tag]
    C, G, H, L, F, N = [int(v) for v in g['meta']]
    model = setup_model(3, H, 1, L, 'fourier', F, cfg['drop'], cfg['p'], cfg['thr'], cfg['wavelet'], C, G, '')
    sd = {k: torch.from_numpy(v) for k, v in state(g).items()}
    model.load_state_dict(sd)
    model.cuda()
    model.train()
    return model, g, cfg


@contextlib.contextmanager
def replay_noise(draws):
    """Make torch.randn_like / torch.rand return the reference's recorded draws (on the GPU), in order."""
    it = iter([torch.from_numpy(np.ascontiguousarray(d)).cuda() for d in draws])
    randn_like, rand = torch.randn_like, torch.rand
    torch.randn_like = lambda *a, **k: next(it)
    torch.rand = lambda *a, **k: next(it)
    try:
        yield
    finally:
        torch.randn_like, torch.rand = randn_like, rand


def noise_of(g, prefix='noise'):
    n = int(g['n_' + prefix][0])
    return [g['%s.%d' % (prefix, i)] for i in range(n)]
