"""Mask-layer protocol (reference: model/Dropout_Layer.py:4-39).

Besides the reference protocol (``forward``, ``calculate_pruning_mask``, ``multiply_values_with_dropout``,
``size_layer``, ``create_instance``, class-level threshold list) every layer here answers ``mask_spec(training)``:
a description of its multiplier that ``Feature_Grid_Model`` hands to the fused CUDA synthesis
(``lfgc_mask_multiplier`` + ``lfgc_decode_fwd``) instead of materialising ``coeff * mask`` with torch ops.
"""
from __future__ import annotations

from typing import Optional

import torch


class MaskSpec:
    """How one mask layer multiplies its coefficient tensor in the current mode."""

    __slots__ = ('mode', 'p0', 'p1', 'noise', 'threshold', 'grad_params')

    def __init__(self, mode, p0, p1=None, noise=None, threshold=0.0, grad_params=()):
        self.mode = mode
        self.p0, self.p1, self.noise = p0, p1, noise
        self.threshold = float(threshold)
        self.grad_params = tuple(grad_params)  # nn.Parameters that receive a gradient through the multiplier


class DropoutLayer(torch.nn.Module):
    # class-level hooks kept for parity with the reference (Dropout_Layer.py:6-7,15-18,32-35)
    i = 0
    theshold_list = None

    def __init__(self, size=0, p: float = 0.5, threshold: float = 0.9):
        super().__init__()
        self.c = size
        self.p = p
        self.threshold = threshold
        if DropoutLayer.theshold_list is not None and DropoutLayer.i != 0:
            self.threshold = DropoutLayer.theshold_list[DropoutLayer.i - 1]
        DropoutLayer.i = DropoutLayer.i + 1

    # -- reference protocol ---------------------------------------------------------------------------------------
    def forward(self, x):
        spec = self.mask_spec(self.training)
        if spec is None:
            return x
        return apply_mask_standalone(x, spec)

    def calculate_pruning_mask(self, device):
        raise NotImplementedError

    def multiply_values_with_dropout(self, input, device):
        raise NotImplementedError

    def size_layer(self):
        raise NotImplementedError

    @classmethod
    def set_threshold_list(cls, list):
        cls.i = 0
        cls.theshold_list = list

    @classmethod
    def create_instance(cls, size, sign_variance_momentum=0.02, threshold=0.9):
        return cls(size, sign_variance_momentum, threshold)

    # -- fused-path protocol --------------------------------------------------------------------------------------
    def mask_spec(self, training: bool) -> Optional[MaskSpec]:
        """None = identity in this mode."""
        return None


class _StandaloneMask(torch.autograd.Function):
    """x * multiplier for a mask layer called on its own (outside Feature_Grid_Model): CUDA multiplier kernel,
    gradients through lfgc_mask_param_grad."""

    @staticmethod
    def forward(ctx, x, spec, *params):
        from .. import ops
        mult, aux = ops.mask_multiplier(spec.mode, spec.p0.detach(), None if spec.p1 is None else spec.p1.detach(),
                                        spec.noise, spec.threshold, want_aux=True)
        ctx.spec = spec
        ctx.save_for_backward(x, aux)
        return x * mult.unsqueeze(0)

    @staticmethod
    def backward(ctx, g):
        from .. import ops
        x, aux = ctx.saved_tensors
        spec = ctx.spec
        gx = g * aux.unsqueeze(0)
        grads = []
        if spec.grad_params:
            gm = (g * x).sum(dim=0).contiguous()
            g0, g1 = ops.mask_param_grad(spec.mode, spec.p0.detach(), None if spec.p1 is None else spec.p1.detach(),
                                         spec.noise, gm)
            grads = [g0] if len(spec.grad_params) == 1 else [g0, g1]
        return (gx, None, *grads)


def apply_mask_standalone(x, spec: MaskSpec):
    return _StandaloneMask.apply(x, spec, *spec.grad_params)
