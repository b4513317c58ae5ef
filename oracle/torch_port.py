"""CPU baseline: a port of the reference's training / reconstruction step onto the same ATen operators the reference
itself calls, for TIMING on the host cores.

TEST / BENCH INFRASTRUCTURE ONLY (see oracle/fvsrn_numpy.py for the rules).  The reference is pure PyTorch and
cannot travel to the GPU box, so ``bench.py --impl reference`` and the ``cpu_baseline`` leg time this port instead
(``kind: "port"``).  It issues the operators the reference issues per step -- ``conv_transpose3d`` synthesis per
wavelet level (wavelet_transform/Torch_Wavelet_Transform.py:100-104), ``F.grid_sample`` (model/Feature_Grid_Model.py
:63), sin/cos embedding (:67), ``nn.functional.linear`` + SnakeAlt (:72-75), MSE, autograd backward and
``torch.optim.Adam`` (training/training.py:130-138) -- with every host thread torch can use.  The sampler is a
vectorised ``randint`` + gather (cheaper than the reference's DataLoader path, i.e. the baseline is flattered, not
handicapped).  ``tests/test_cpu_port.py`` checks the port against the numpy oracle.
"""
from __future__ import annotations

import math
import time

import numpy as np
import torch
import torch.nn.functional as F

from . import fvsrn_numpy as O


def _filter_bank_rev(wavelet: str):
    _, _, rec_lo, rec_hi = O.wavelet_taps(wavelet)
    f = (torch.tensor(rec_lo, dtype=torch.float32), torch.tensor(rec_hi, dtype=torch.float32))
    bank = [f[a][:, None, None] * f[b][None, :, None] * f[c][None, None, :]
            for a in (0, 1) for b in (0, 1) for c in (0, 1)]
    return torch.stack(bank).unsqueeze(1)  # (8, 1, L, L, L)


class CpuPort:
    """State + step functions of one model instance on the CPU."""

    def __init__(self, spec: O.Spec, state: dict, lr: float = 0.008):
        self.spec = spec
        self.params = {k: torch.tensor(np.asarray(v), dtype=torch.float32, requires_grad=True)
                       for k, v in state.items() if not k.startswith('filter.')}
        self.filt = _filter_bank_rev(spec.wavelet)
        self.opt = torch.optim.Adam(list(self.params.values()), lr=lr)
        freqs = O.fourier_freqs(spec.F)
        self.freqs = [torch.tensor(f) for f in freqs]

    def decode(self):
        sp, P = self.spec, self.params
        restored = P['feature_grid.0'].unsqueeze(0)
        C = sp.C
        w = self.filt.repeat(C, 1, 1, 1, 1)
        for l in range(1, sp.n_levels):
            data = torch.cat([restored.unsqueeze(2), P['feature_grid.%d' % l].unsqueeze(0)], dim=2)
            full = F.conv_transpose3d(torch.flatten(data, 1, 2), w, groups=C, stride=2)
            tgt = sp.shape_array[l - 1]
            sl = []
            for a in range(3):
                d = full.shape[2 + a] - int(tgt[a])
                lo = d // 2
                sl.append(slice(lo, lo + int(tgt[a])))
            restored = full[:, :, sl[0], sl[1], sl[2]]
        return restored[0]

    def forward(self, coords, clamp=False):
        sp, P = self.spec, self.params
        grid = self.decode()
        n = coords.shape[0]
        feats = F.grid_sample(grid.unsqueeze(0), coords.view(1, 1, 1, n, 3), mode='bilinear',
                              align_corners=False).reshape(sp.C, n).t()
        emb = []
        for f in self.freqs:
            emb += [torch.sin(coords * f), torch.cos(coords * f)]
        x = torch.cat([coords] + emb + [feats], dim=-1)
        for l in range(sp.L):
            z = F.linear(x, P['net_layers.%d.weight' % l], P['net_layers.%d.bias' % l])
            x = 0.5 * z + torch.sin(z) ** 2
        y = F.linear(x, P['final_layer.weight'], P['final_layer.bias'])
        return y.clamp(-1, 1) if clamp else y

    def train_step(self, volume: torch.Tensor, n: int, gen: torch.Generator):
        """sampler + GT + forward + MSE + backward + Adam on n samples; returns the loss value."""
        shape = volume.shape
        idx = torch.randint(0, volume.numel(), (n,), generator=gen)
        i = torch.div(idx, shape[1] * shape[2], rounding_mode='floor')
        j = torch.div(idx, shape[2], rounding_mode='floor') % shape[1]
        k = idx % shape[2]
        raw = torch.stack([i, j, k], dim=-1).float()
        max_idx = torch.tensor([s - 1.0 for s in shape])
        scales = max_idx / max_idx.max()
        norm = scales * (2.0 * (raw / max_idx) - 1.0)
        gt = volume.reshape(-1)[idx]
        self.opt.zero_grad(set_to_none=True)
        pred = self.forward(norm).squeeze(-1)
        loss = F.mse_loss(pred, gt)
        loss.backward()
        self.opt.step()
        return float(loss.detach())

    @torch.no_grad()
    def reconstruct(self, coords):
        return self.forward(coords, clamp=True)


def make_state(spec: O.Spec, seed: int = 0):
    """Random-init parameters of the reference architecture (numpy, keyed by state-dict names)."""
    rng = np.random.default_rng(seed)
    grid = rng.uniform(0.0, 1.0, size=(spec.C, spec.G, spec.G, spec.G))
    coeffs, _ = O.encode_volume(grid, spec.wavelet)
    sd = {'feature_grid.%d' % i: c.astype(np.float32) for i, c in enumerate(coeffs)}
    in0 = 3 + 6 * spec.F + spec.C
    dims = [(spec.H, in0)] + [(spec.H, spec.H)] * (spec.L - 1)
    for l, (o, i) in enumerate(dims):
        bound = 1.0 / math.sqrt(i)
        sd['net_layers.%d.weight' % l] = rng.uniform(-bound, bound, size=(o, i)).astype(np.float32)
        sd['net_layers.%d.bias' % l] = rng.uniform(-bound, bound, size=(o,)).astype(np.float32)
    bound = 1.0 / math.sqrt(spec.H)
    sd['final_layer.weight'] = rng.uniform(-bound, bound, size=(1, spec.H)).astype(np.float32)
    sd['final_layer.bias'] = rng.uniform(-bound, bound, size=(1,)).astype(np.float32)
    return sd


def time_train_steps(spec: O.Spec, volume: torch.Tensor, n: int, steps: int, warmup: int, seed: int = 0):
    """Returns (seconds for `steps` steps, threads used)."""
    port = CpuPort(spec, make_state(spec, seed))
    gen = torch.Generator().manual_seed(seed)
    for _ in range(warmup):
        port.train_step(volume, n, gen)
    t0 = time.perf_counter()
    for _ in range(steps):
        port.train_step(volume, n, gen)
    return time.perf_counter() - t0, torch.get_num_threads()
