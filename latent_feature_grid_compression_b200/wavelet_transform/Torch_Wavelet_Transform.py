"""3-D orthogonal wavelet filter bank -- API counterpart of the reference's
wavelet_transform/Torch_Wavelet_Transform.py (``_WaveletFilterNd`` / ``WaveletFilter3d``).

The reference runs analysis as a grouped ``conv3d`` and synthesis as a grouped ``conv_transpose3d`` with an
(8,1,L,L,L) outer-product filter bank.  Here both directions are hand-written CUDA (csrc/wavelet.cu):
``encode`` -> ``lfgc_dwt_level`` (model construction only), ``decode`` -> ``lfgc_decode_fwd`` for one level.
``filter_fwd`` / ``filter_rev`` are still registered as buffers with the reference's shapes and values so that
state dicts interchange (keys ``filter.filter_fwd`` / ``filter.filter_rev``).
"""
from __future__ import annotations

from typing import Union

import numpy as np
import torch
from torch import nn, Tensor

from .. import ops, wavelets


def _outer3(fa, fb, fc):
    return fa[:, None, None] * fb[None, :, None] * fc[None, None, :]


class _WaveletFilterNd(nn.Module):

    def __init__(self, wavelet: Union[str, object], dim: int = 3, padding='constant'):
        super().__init__()
        if dim != 3:
            raise NotImplementedError('only the 3-D filter bank of the fV-SRN path is implemented')
        if padding != 'constant':
            raise NotImplementedError('the reference path uses zero ("constant") padding')
        self.dim = dim
        self.padding = padding
        self.wavelet_name = wavelet if isinstance(wavelet, str) else getattr(wavelet, 'name')
        dec_lo, dec_hi, rec_lo, rec_hi = (torch.tensor(t) for t in wavelets.filter_bank(self.wavelet_name))
        if len(dec_lo) % 2:
            raise ValueError('uneven filter lengths are not supported')
        # sub-band k = 4a + 2b + c: filter a along D, b along H, c along W (lo = 0, hi = 1)
        fwd = (dec_lo.flip(-1), dec_hi.flip(-1))
        rev = (rec_lo, rec_hi)
        bank = lambda f: torch.stack([_outer3(f[a], f[b], f[c]) for a in (0, 1) for b in (0, 1) for c in (0, 1)])
        self.register_buffer('filter_fwd', bank(fwd).unsqueeze(1))
        self.register_buffer('filter_rev', bank(rev).unsqueeze(1))

    @property
    def filter_length(self):
        return self.filter_fwd.shape[-1]

    def encode(self, data: Tensor):
        """(batch, C, d0, d1, d2) -> coefficients (batch, C, 8, e0, e1, e2) and the input's spatial shape."""
        if data.dim() != 5:
            raise ValueError('encode expects (batch, channel, D, H, W)')
        shape = np.asarray(data.shape[-3:])
        out = [ops.dwt_level(x.contiguous().float(), self.wavelet_name) for x in data]
        return torch.stack(out, dim=0), shape

    def decode(self, data: Tensor, shape):
        """(batch, C, 8, d0, d1, d2) -> (batch, C, *shape): one synthesis level."""
        if data.dim() != 6:
            raise ValueError('decode expects (batch, channel, 8, D, H, W)')
        outs = []
        for x in data:
            C = x.shape[0]
            d = tuple(x.shape[-3:])
            geom = ops.Geometry(C, tuple(int(s) for s in shape), 1, 1, 0, self.wavelet_name, [d, d], [list(shape)])
            low = x[:, 0].contiguous().float()
            high = x[:, 1:].contiguous().float()
            cl = ops.decode_fwd(geom, [low, high], [None, None])
            outs.append(cl[..., :C].permute(3, 0, 1, 2).contiguous())
        return torch.stack(outs, dim=0)

    def forward(self, data: Tensor):
        return self.encode(data)[0]


class WaveletFilter3d(_WaveletFilterNd):

    def __init__(self, wavelet, padding='constant'):
        super().__init__(wavelet, 3, padding=padding)
