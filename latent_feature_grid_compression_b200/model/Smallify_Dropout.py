"""Smallify mask layer, its sign-variance tracker and the Smallify regulariser
(reference: model/Smallify_Dropout.py).

Differences in mechanism, not in results: the multiplier ``coeff * betas`` is applied inside the fused CUDA
synthesis, and the tracker's EMA / EMAVar live on the device and are updated by ``lfgc_smallify_ema`` -- the
reference moves ``sign(betas)`` to the host on every forward (Smallify_Dropout.py:106-112).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import _lib as L
from .. import ops
from .Dropout_Layer import DropoutLayer, MaskSpec
from .Straight_Through_Dropout import MaskedWavelet_Straight_Through_Dropout, Straight_Through_Dropout


class SmallifyLoss(nn.Module):
    """lambda_1 * sum |mask parameters| + lambda_2 * sum coeff^2 over the model (Smallify_Dropout.py:10-40)."""

    def __init__(self, weight_l1: float = 1., weight_l2: float = 1.):
        super().__init__()
        self.weight_l1 = float(weight_l1)
        self.weight_l2 = float(weight_l2)

    def forward(self, model: nn.Module) -> torch.Tensor:
        from .Feature_Grid_Model import Feature_Grid_Model
        l1_terms, l2_terms = [], []
        for m in model.modules():
            if isinstance(m, (SmallifyDropout, MaskedWavelet_Straight_Through_Dropout, Straight_Through_Dropout)):
                l1_terms.append(m.l1_loss())
            if isinstance(m, Feature_Grid_Model):
                l2_terms.append(sum(torch.sum(torch.abs(f) ** 2) for f in m.feature_grid))
        loss = 0.
        if self.weight_l1 > 0.:
            loss = loss + self.weight_l1 * sum(l1_terms)
        if self.weight_l2 > 0.:
            loss = loss + self.weight_l2 * sum(l2_terms)
        return loss


class SmallifySignVarianceTracker:
    """EMA of sign(beta) and of its variance (Smallify_Dropout.py:78-118), kept on the device of ``betas``."""

    def __init__(self, c, sign_variance_momentum, threshold, betas):
        self.c = c
        self.sign_variance_momentum = sign_variance_momentum
        self.threshold = threshold
        self.EMA, self.EMAVar = self.init_variance_data(betas)

    def init_variance_data(self, betas):
        with torch.no_grad():
            return torch.sign(betas.detach()).clone(), torch.zeros_like(betas.detach())

    def _sync_device(self, betas):
        if self.EMA.device != betas.device:
            self.EMA = self.EMA.to(betas.device)
            self.EMAVar = self.EMAVar.to(betas.device)

    def sign_variance_pruning_onlyVar(self, betas):
        self._sync_device(betas)
        ops.smallify_ema(betas.detach(), self.EMA, self.EMAVar, self.sign_variance_momentum)

    def sign_variance_pruning(self, device, betas):
        self.sign_variance_pruning_onlyVar(betas)
        return self.calculate_pruning_mask(device)

    def calculate_pruning_mask(self, device):
        with torch.no_grad():
            keep = torch.where(self.EMAVar < self.threshold, 1.0, 0.0)
        return keep.to(device)


class SmallifyDropout(DropoutLayer):

    def __init__(self, size=(1, 1, 1), sign_variance_momentum=0.025, threshold=0.75):
        super().__init__(size, sign_variance_momentum, threshold)
        self.betas = torch.nn.Parameter(torch.empty(size).normal_(0, 1), requires_grad=True)
        self.tracker = SmallifySignVarianceTracker(self.c, sign_variance_momentum, self.threshold, self.betas)
        self.d_mask = None

    def mask_spec(self, training):
        if not training:
            return None  # identity in eval mode, baked or not (Smallify_Dropout.py:54-61)
        if self.d_mask is None:
            self.tracker.sign_variance_pruning_onlyVar(self.betas)
            return MaskSpec(L.MASK_DIRECT, self.betas, grad_params=(self.betas,))
        return MaskSpec(L.MASK_DIRECT, self.d_mask)

    def l1_loss(self):
        return torch.abs(self.betas).sum()

    def calculate_pruning_mask(self, device):
        mask = self.tracker.calculate_pruning_mask(device)
        self.d_mask = mask.contiguous()
        return mask

    def multiply_values_with_dropout(self, input, device):
        with torch.no_grad():
            mask = self.calculate_pruning_mask(device) * self.betas.unsqueeze(0)
            return input * mask

    def size_layer(self):
        return self.betas.numel()
