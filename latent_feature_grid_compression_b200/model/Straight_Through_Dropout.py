"""Straight-through mask layers (reference: model/Straight_Through_Dropout.py).

* ``Straight_Through_Dropout``: Bernoulli mask ``rand < mask_values`` (no gradient path to the mask values other
  than the L1 penalty).  The reference class lacks ``size_layer`` and therefore crashes when the masks are baked
  (Feature_Grid_Model.py:125); ``size_layer`` is provided here, the forward semantics are unchanged.
* ``MaskedWavelet_Straight_Through_Dropout``: hard threshold on sigmoid(mask_values) in the forward value,
  gradient of ``x * sigmoid(mask_values)``.
"""
from __future__ import annotations

import torch
from torch.nn import functional as F

from .. import _lib as L
from .Dropout_Layer import DropoutLayer, MaskSpec


class STEFunction(torch.autograd.Function):
    """Kept for API parity (Straight_Through_Dropout.py:10-17)."""

    @staticmethod
    def forward(ctx, input, thresh):
        return input < thresh

    @staticmethod
    def backward(ctx, grad_output):
        return F.hardtanh(grad_output)


class Straight_Through_Dropout(DropoutLayer):

    def __init__(self, size=(1, 1, 1), probability=0.5, threshold=0.5):
        super().__init__(size, probability, threshold)
        self.mask_values = torch.nn.Parameter(torch.ones(size), requires_grad=True)

    def mask_spec(self, training):
        if not training:
            return None
        u = torch.rand(self.c, device=self.mask_values.device)  # same draw as the reference (:28)
        return MaskSpec(L.MASK_BERNOULLI, self.mask_values, noise=u)

    def l1_loss(self):
        return torch.abs(self.mask_values).sum()

    def calculate_pruning_mask(self, device):
        return self.mask_values > self.threshold

    def multiply_values_with_dropout(self, input, device):
        with torch.no_grad():
            return input * self.calculate_pruning_mask(device)

    def size_layer(self):
        return self.mask_values.numel()


class MaskedWavelet_Straight_Through_Dropout(DropoutLayer):

    def __init__(self, size=(1, 1, 1), probability=0.5, threshold=0.5):
        super().__init__(size, probability, threshold)
        self.mask_values = torch.nn.Parameter(torch.ones(size), requires_grad=True)
        self.d_mask = None
        self._d_mask_f = None

    def mask_spec(self, training):
        if not training:
            return None
        if self.d_mask is None:
            return MaskSpec(L.MASK_STE_SIGMOID, self.mask_values, threshold=self.threshold,
                            grad_params=(self.mask_values,))
        if self._d_mask_f is None or self._d_mask_f.device != self.mask_values.device:
            self._d_mask_f = self.d_mask.to(device=self.mask_values.device, dtype=torch.float32).contiguous()
        return MaskSpec(L.MASK_DIRECT, self._d_mask_f)

    def l1_loss(self):
        return torch.abs(self.mask_values).sum()

    def calculate_pruning_mask(self, device):
        mask = torch.sigmoid(self.mask_values)
        self.d_mask = (mask >= self.threshold).to(device)
        self._d_mask_f = None
        return mask

    def multiply_values_with_dropout(self, input, device):
        with torch.no_grad():
            mask = self.calculate_pruning_mask(device)
            return (input * (mask >= self.threshold) - input * mask) + (input * mask)

    def size_layer(self):
        return self.mask_values.numel()
