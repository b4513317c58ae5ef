"""Variational dropout on the wavelet coefficients (reference: model/Variational_Dropout_Layer.py):
additive-noise reparameterisation w = theta + sigma * xi (Molchanov et al.), its KL term, the variational loss and
the small ReLU ``Variance_Model`` that predicts a per-sample log sigma.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import _lib as L
from .Dropout_Layer import DropoutLayer, MaskSpec


def inference_variational_model(mu, sigma):
    return torch.normal(mu, sigma)


def calculate_Log_Likelihood(loss_criterion, predicted_volume, ground_truth_volume, log_sigma):
    """Static-sigma Gaussian log likelihood (Variational_Dropout_Layer.py:15-21)."""
    x_mu_loss = loss_criterion(predicted_volume, ground_truth_volume)
    sigma = math.exp(log_sigma)
    return -x_mu_loss / (2 * (sigma ** 2)) - (math.log(2 * math.pi) + (2 * log_sigma)) / 2, x_mu_loss


def calculate_Log_Likelihood_variance(predicted_volume, ground_truth_volume, variance):
    """Per-sample log likelihood with ``variance`` = log sigma (Variational_Dropout_Layer.py:24-30)."""
    x_mu_loss = (ground_truth_volume - predicted_volume) ** 2
    sigma = torch.exp(variance)
    a = 1 / (2 * (sigma ** 2))
    b = - (math.log(2 * np.pi) + (2 * variance)) / 2
    return a * (-x_mu_loss) + b, x_mu_loss


class VariationalDropoutLoss(nn.Module):
    """-(LL*s - w_dkl*sum DKL*s - w_w*sum coeff^2*s), s = n_voxels / batch; w_dkl ramps by (1 + multiplier) per call
    until it reaches weight_dkl_max (Variational_Dropout_Layer.py:33-69)."""

    def __init__(self, size_volume: float, batch_size: float, weight_dkl: float = 1., weight_weights: float = 1.,
                 weight_dkl_max=30.0):
        super().__init__()
        self.batch_scale = (size_volume / batch_size)
        self.weight_dkl = float(weight_dkl)
        self.weight_dkl_max = weight_dkl_max
        self.weight_weights = float(weight_weights)

    def forward(self, model: nn.Module, predicted_volume, ground_truth_volume, log_sigma, weight_dkl_multiplier):
        from .Feature_Grid_Model import Feature_Grid_Model
        dkl_terms, weight_terms = [], []
        for m in model.modules():
            if isinstance(m, VariationalDropout):
                dkl_terms.append(m.calculate_Dkl())
            if isinstance(m, Feature_Grid_Model):
                weight_terms.append(sum(torch.sum(torch.abs(f) ** 2) for f in m.feature_grid))
        if self.weight_dkl < self.weight_dkl_max:
            self.weight_dkl = self.weight_dkl * (1.0 + weight_dkl_multiplier)
        ll, sq_err = calculate_Log_Likelihood_variance(predicted_volume, ground_truth_volume, log_sigma)
        mse = sq_err.sum() * (1 / predicted_volume.shape[0])
        ll = ll.sum() * self.batch_scale
        dkl_sum = self.weight_dkl * sum(dkl_terms) * self.batch_scale
        weight_sum = self.weight_weights * sum(weight_terms) * self.batch_scale
        loss = -(ll - dkl_sum - weight_sum)
        return loss, ll, mse, dkl_sum, weight_sum


class VariationalDropout(DropoutLayer):
    # constants of the KL approximation (Molchanov et al. 2017), Variational_Dropout_Layer.py:74-77
    k1 = 0.63576
    k2 = 1.87320
    k3 = 1.48695
    C = -k1

    def __init__(self, size=(1, 1, 1), init_dropout=0.5, threshold=0.9):
        super().__init__(size, init_dropout, threshold)
        self.log_thetas = torch.nn.Parameter(torch.zeros(size), requires_grad=True)
        log_alphas = math.log(init_dropout / (1 - init_dropout))
        self.log_var = torch.nn.Parameter(torch.empty(size).fill_(log_alphas), requires_grad=True)
        self.d_mask = None

    @property
    def alphas(self):
        return torch.exp(self.log_var - 2.0 * self.log_thetas)

    @property
    def dropout_rates(self):
        return self.alphas / (1.0 + self.alphas)

    @property
    def sigma(self):
        return torch.exp(self.log_var / 2.0)

    def mask_spec(self, training):
        # noise is drawn on every call, train and eval, baked or not -- as the reference does (:105-108)
        xi = torch.randn_like(self.log_thetas)
        if self.d_mask is None:
            return MaskSpec(L.MASK_VARIATIONAL, self.log_thetas, self.log_var, noise=xi,
                            grad_params=(self.log_thetas, self.log_var))
        return MaskSpec(L.MASK_DIRECT, self.d_mask)

    def calculate_Dkl(self):
        log_alphas = self.log_var - 2.0 * self.log_thetas
        t1 = self.k1 * torch.sigmoid(self.k2 + self.k3 * log_alphas)
        t2 = 0.5 * F.softplus(-log_alphas, beta=1.)
        return torch.sum(- t1 + t2 + self.k1)

    def calculate_Dropout_Entropy(self):
        r = self.dropout_rates
        return torch.sum(r * torch.log(r) + (1.0 - r) * torch.log(1 - r))

    def get_valid_fraction(self):
        not_dropped = torch.mean((self.dropout_rates < self.threshold).to(torch.float)).item()
        return not_dropped, self.dropout_rates

    def calculate_pruning_mask(self, device):
        with torch.no_grad():
            keep = torch.where(self.dropout_rates < self.threshold, 1.0, 0.0)
            if keep.numel() - torch.count_nonzero(keep) == 0:
                keep.data[0] = 1.0  # reference quirk (:142-143): only fires when nothing is pruned
            self.d_mask = keep.to(device).contiguous()
            return keep.to(device)

    def multiply_values_with_dropout(self, input, device):
        with torch.no_grad():
            mask = self.calculate_pruning_mask(device) * torch.exp(self.log_thetas)
            return input * mask

    def size_layer(self):
        return self.log_thetas.numel()


class _PlainMLPFunction(torch.autograd.Function):
    """lfgc_plain_mlp_forward / lfgc_plain_mlp_backward (forward recomputed in the backward; the coordinate gradient
    the reference's autograd would also produce is never read, training/training.py:99,119-121)."""

    @staticmethod
    def forward(ctx, x, width, n_layers, flat, *params):
        from .. import ops
        out = ops.plain_mlp_forward(width, n_layers, x, flat)
        ctx.saved = (x, flat, width, n_layers, [tuple(p.shape) for p in params])
        ctx.flat_version = flat._version      # the backward recomputes from live storage: check it by hand
        return out

    @staticmethod
    def backward(ctx, grad_out):
        from .. import ops
        x, flat, width, n_layers, shapes = ctx.saved
        if flat._version != ctx.flat_version:
            raise RuntimeError('one of the variables needed for gradient computation has been modified by an inplace '
                               'operation (a Variance_Model parameter changed between forward and backward)')
        g = ops.plain_mlp_backward(width, n_layers, x, grad_out.contiguous().float(), flat)
        grads, off = [], 0
        for shp in shapes:
            n = int(np.prod(shp))
            grads.append(g[off:off + n].view(shp))
            off += n
        return (None, None, None, None, *grads)


class Variance_Model(nn.Module):
    """3 -> 32 x 4 -> 1 ReLU MLP predicting log sigma per sample (Variational_Dropout_Layer.py:159-175).  Same
    constructor, ``net_layers`` / ``final_layer`` parameters and ``forward(input) -> (N, 1)``; the arithmetic is the
    fused coordinate-MLP kernel pair behind lfgc_plain_mlp_forward / _backward."""

    def __init__(self, input_ch=3, output_ch=1, n_layers=4, size_layers=32):
        super().__init__()
        if input_ch != 3 or output_ch != 1 or size_layers > 32 or not 1 <= n_layers <= 4:
            raise L.LfgcError('Variance_Model kernels are built for 3 -> (<=32) x (<=4) -> 1 (got %d -> %d x %d -> %d)'
                              % (input_ch, size_layers, n_layers, output_ch))
        self.net_layers = nn.ModuleList(
            [nn.Linear(input_ch, size_layers)] + [nn.Linear(size_layers, size_layers) for _ in range(n_layers - 1)])
        self.final_layer = nn.Linear(size_layers, output_ch)
        self.width, self.n_layers = size_layers, n_layers
        from .. import ops
        self._pack = ops.FlatPack()

    def _params(self):
        ps = []
        for layer in self.net_layers:
            ps += [layer.weight, layer.bias]
        return ps + [self.final_layer.weight, self.final_layer.bias]

    def mlp_flat(self):
        return self._pack.ensure(self._params())

    def forward(self, input):
        if not input.is_cuda:
            raise L.LfgcError('Variance_Model runs on CUDA only (no CPU fallback); got a %s tensor' % input.device)
        x = input.detach().reshape(-1, 3).contiguous().float()
        out = _PlainMLPFunction.apply(x, self.width, self.n_layers, self.mlp_flat(), *self._params())
        return out.view(*input.shape[:-1], 1)
