"""Fourier positional embedding -- API counterpart of the reference's model/Feature_Embedding.py.

Inside ``Feature_Grid_Model`` the embedding never exists as a tensor: the fused sample kernels compute
sin/cos(coord * omega_k) in registers (csrc/sample_common.cuh, ``stage_inputs``).  This class only carries the
metadata the model needs (``out_dim``, ``n_freqs``, ``freq_bands``) and offers ``embed`` for callers that want the
features on their own; that helper is plain tensor code and not part of the hot path.

Layout (reference Feature_Embedding.py:28-34): for k = 0..n_freqs-1: sin(x * w_k) over all input dims, then
cos(x * w_k) over all input dims, with w_k = 2^k * 2 * pi evaluated in fp32.
"""
from __future__ import annotations

import math
from functools import partial

import torch


def _periodic(fn, omega, x):
    return fn(x * omega)


class Embedder:
    """Base class: ``out_dim`` extra features produced by ``embed``."""

    def __init__(self):
        self.embed_functions = []
        self.out_dim = 0

    def embed(self, inputs):
        if not self.embed_functions:
            return inputs.new_zeros((*inputs.shape[:-1], 0))
        return torch.cat([f(inputs) for f in self.embed_functions], dim=-1)


class FourierEmbedding(Embedder):

    def __init__(self, n_freqs, input_dim):
        super().__init__()
        self.n_freqs = int(n_freqs)
        self.input_dim = int(input_dim)
        # fp32 arithmetic in the reference's order: (2 ** linspace) * 2 * pi
        octave = torch.pow(torch.tensor(2.0), torch.linspace(0.0, self.n_freqs - 1, steps=self.n_freqs))
        self.freq_bands = octave * 2.0 * math.pi
        for omega in self.freq_bands:
            self.embed_functions.append(partial(_periodic, torch.sin, omega))
            self.embed_functions.append(partial(_periodic, torch.cos, omega))
        self.out_dim = 2 * self.n_freqs * self.input_dim
