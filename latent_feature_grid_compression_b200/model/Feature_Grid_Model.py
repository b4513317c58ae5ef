"""fV-SRN latent-feature-grid model -- drop-in counterpart of the reference's model/Feature_Grid_Model.py.

Same constructor, attributes, parameter names and ``forward(coords)`` contract; the arithmetic runs in a handful of
hand-written sm_100a launches behind the C ABI (include/lfgc.h):

    lfgc_mask_multiplier + lfgc_decode_fwd   mask x wavelet coefficients -> synthesis -> channels-last grid
    lfgc_forward                             trilinear gather + Fourier features + SnakeAlt MLP, fused
    lfgc_backward + lfgc_decode_bwd          fused backward (forward recomputed), synthesis adjoint, mask gradients

Nothing here falls back to torch ops; without liblfgc.so or a CUDA device the model raises.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from .. import _lib as L
from .. import ops, wavelets
from .Dropout_Layer import DropoutLayer
from .Feature_Embedding import Embedder


def SnakeAlt(x):
    """0.5 x + sin^2 x -- the activation the fused kernels implement (kept for callers that import it)."""
    return 0.5 * x + torch.sin(x) ** 2


def _multipliers(specs):
    """(value multiplier, gradient multiplier) per level from the mask specs; None = identity."""
    mults, auxs = [], []
    for spec in specs:
        if spec is None:
            mults.append(None)
            auxs.append(None)
            continue
        need_aux = spec.mode == L.MASK_STE_SIGMOID
        m, a = ops.mask_multiplier(spec.mode, spec.p0.detach(), None if spec.p1 is None else spec.p1.detach(),
                                   spec.noise, spec.threshold, want_aux=need_aux)
        mults.append(m)
        auxs.append(a if need_aux else m)
    return mults, auxs


class _FeatureGridFunction(torch.autograd.Function):
    """decode + fused sample forward; backward = fused sample backward + synthesis adjoint + mask gradients."""

    @staticmethod
    def forward(ctx, coords, geom, specs, mlp_flat, clamp, n_coeff, *params):
        coeffs = [p.detach() for p in params[:n_coeff]]
        mults, auxs = _multipliers(specs)
        grid_cl = ops.decode_fwd(geom, coeffs, mults)
        out = ops.sample_forward(geom, coords, grid_cl, mlp_flat, clamp=clamp)
        ctx.geom, ctx.specs = geom, specs
        ctx.coords, ctx.grid_cl, ctx.mlp_flat, ctx.coeffs, ctx.auxs = coords, grid_cl, mlp_flat, coeffs, auxs
        # The backward pass recomputes the forward from live parameter storage (views into the flat buffers), which
        # bypasses autograd's version-counter check: do that check by hand, as torch would for saved tensors.
        watched = list(params[:n_coeff]) + [mlp_flat] + [t for s in specs if s is not None for t in (s.p0, s.p1) if t is not None]
        ctx.versions = [(t, t._version) for t in watched]
        return out

    @staticmethod
    def backward(ctx, grad_out):
        geom, specs = ctx.geom, ctx.specs
        for t, v in ctx.versions:
            if t._version != v:
                raise RuntimeError('one of the variables needed for gradient computation has been modified by an inplace '
                                   'operation (a parameter of Feature_Grid_Model changed between forward and backward)')
        gout = grad_out.contiguous().float()
        grad_grid_cl, grad_mlp = ops.sample_backward(geom, ctx.coords, gout, ctx.grid_cl, ctx.mlp_flat)
        want_gmult = [s is not None and len(s.grad_params) > 0 for s in specs]
        grad_coeffs, grad_mults = ops.decode_bwd(geom, grad_grid_cl, ctx.coeffs, ctx.auxs, want_gmult)
        grads = list(grad_coeffs)
        for spec, gm in zip(specs, grad_mults):
            if spec is None or not spec.grad_params:
                continue
            g0, g1 = ops.mask_param_grad(spec.mode, spec.p0.detach(), None if spec.p1 is None else spec.p1.detach(),
                                         spec.noise, gm)
            grads.append(g0)
            if len(spec.grad_params) == 2:
                grads.append(g1)
        off = 0
        for _, shape in geom.mlp_shapes():
            n = int(np.prod(shape))
            grads.append(grad_mlp[off:off + n].view(shape))
            off += n
        return (None, None, None, None, None, None, *grads)


class Feature_Grid_Model(nn.Module):

    def __init__(self, embedder: Embedder, feature_grid, drop_layer: DropoutLayer, wavelet_filter,
                 input_channel_data=3, hidden_channel=32, out_channel=1, num_layer=4):
        super().__init__()
        if input_channel_data != 3 or out_channel != 1:
            raise L.LfgcError('the fV-SRN kernels are built for d_in=3, d_out=1 (got %d, %d)'
                              % (input_channel_data, out_channel))
        self.embedder = embedder
        self.filter = wavelet_filter

        features, shapes = self.encode_volume(feature_grid)
        self.feature_grid = nn.ParameterList([nn.Parameter(f, requires_grad=True) for f in features])
        self.shape_array = shapes

        if drop_layer is None:
            self.drop = nn.ModuleList([nn.Identity() for _ in features])
        else:
            self.drop = nn.ModuleList(
                [drop_layer.create_instance(f.shape[1:], drop_layer.p, drop_layer.threshold) for f in features])

        self.input_channel = input_channel_data + embedder.out_dim + feature_grid.shape[0]
        self.hidden_width = hidden_channel
        self.output_channel = out_channel
        self.num_layer = num_layer
        self.d_in = input_channel_data

        self.net_layers = nn.ModuleList(
            [nn.Linear(self.input_channel, self.hidden_width)]
            + [nn.Linear(self.hidden_width, self.hidden_width) for _ in range(self.num_layer - 1)])
        self.final_layer = nn.Linear(self.hidden_width, self.output_channel)

        self._grid_channels = int(feature_grid.shape[0])
        self._grid_shape = tuple(int(s) for s in feature_grid.shape[-3:])
        self._geom = None
        self._mlp_pack = ops.FlatPack()

    # ---- geometry / packing ---------------------------------------------------------------------------------------
    def geometry(self) -> ops.Geometry:
        if self._geom is None:
            n_freq = getattr(self.embedder, 'n_freqs', self.embedder.out_dim // (2 * self.d_in))
            dims = [tuple(f.shape[-3:]) for f in self.feature_grid]
            self._geom = ops.Geometry(self._grid_channels, self._grid_shape, self.hidden_width, self.num_layer,
                                      n_freq, self.filter.wavelet_name, dims, self.shape_array)
        return self._geom

    def _mlp_params(self):
        ps = []
        for layer in self.net_layers:
            ps += [layer.weight, layer.bias]
        return ps + [self.final_layer.weight, self.final_layer.bias]

    def mlp_flat(self) -> torch.Tensor:
        """The packed MLP block the kernels read; the nn.Linear parameters are views into it."""
        return self._mlp_pack.ensure(self._mlp_params())

    def mask_specs(self):
        return [d.mask_spec(self.training) if isinstance(d, DropoutLayer) else None for d in self.drop]

    # ---- forward ---------------------------------------------------------------------------------------------------
    def forward(self, input):
        if not input.is_cuda:
            raise L.LfgcError('Feature_Grid_Model runs on CUDA only (no CPU fallback); got a %s tensor' % input.device)
        eval_mode = not self.training
        orig_shape = input.shape
        coords = input.detach().reshape(-1, 3).contiguous().float()
        specs = self.mask_specs()
        params = [f if f.is_contiguous() else f.contiguous() for f in self.feature_grid]
        for s in specs:
            if s is not None:
                params += list(s.grad_params)
        params += self._mlp_params()
        out = _FeatureGridFunction.apply(coords, self.geometry(), specs, self.mlp_flat(), eval_mode,
                                         len(self.feature_grid), *params)
        if eval_mode:
            return out.view(*orig_shape[:-1], 1)  # the intended shape of reference line 78 (broken under torch >= 2)
        return out.view(-1, 1)

    # ---- wavelet representation --------------------------------------------------------------------------------------
    def encode_volume(self, feature_volume, num_levels=None):
        """Spatial grid (C, G, G, G) -> [LLL_coarsest, high_coarsest, ..., high_finest], shape_array (coarse->fine).
        Runs lfgc_dwt_level on the GPU; the coefficients are returned on the input's device."""
        if num_levels is None:
            num_levels = min(wavelets.dwt_max_level(s, self.filter.filter_length) for s in feature_volume.shape[-3:])
        if not torch.cuda.is_available():
            raise L.LfgcError('Feature_Grid_Model needs a CUDA device to build its wavelet representation')
        home = feature_volume.device
        data = feature_volume.detach().to('cuda', torch.float32).contiguous()
        highs, shapes = [], []
        for _ in range(num_levels):
            shapes.append(np.asarray(data.shape[-3:]))
            co = ops.dwt_level(data, self.filter.wavelet_name)
            highs.append(co[:, 1:].contiguous())
            data = co[:, 0].contiguous()
        feats = [data] + highs[::-1]
        return [f.to(home) for f in feats], np.asarray(shapes[::-1], dtype=int).reshape(-1, 3)

    @torch.no_grad()
    def decode_volume(self) -> torch.Tensor:
        """Masked synthesis of the latent grid, returned in the reference's (C, G, G, G) layout."""
        geom = self.geometry()
        mults, _ = _multipliers(self.mask_specs())
        cl = ops.decode_fwd(geom, [f.detach().contiguous() for f in self.feature_grid], mults)
        return cl[..., :geom.C].permute(3, 0, 1, 2).contiguous()

    # ---- post-training mask baking (reference :110-140) ------------------------------------------------------------------
    def save_dropvalues_on_grid(self, device):
        if isinstance(self.drop[0], nn.Identity):
            return torch.tensor(0, dtype=torch.float32)
        baked = [d.multiply_values_with_dropout(g, device) for g, d in zip(self.feature_grid, self.drop)]
        self.feature_grid = nn.ParameterList([nn.Parameter(f.contiguous(), requires_grad=True) for f in baked])
        zeros = 0
        for g in baked:
            zeros += (g.numel() - torch.count_nonzero(g))
        mask_floats = torch.tensor(0, dtype=torch.float32)
        for d in self.drop:
            mask_floats += d.size_layer()
        return zeros - mask_floats / 32.0

    def remove_drop_layers(self, device):
        masks = []
        for d in self.drop:
            if isinstance(d, nn.Identity):
                return
            masks.append(d.calculate_pruning_mask(device))
        pruned = [g * m for g, m in zip(self.feature_grid, masks)]
        self.feature_grid = nn.ParameterList([nn.Parameter(f.contiguous(), requires_grad=True) for f in pruned])
        self.drop = nn.ModuleList([nn.Identity() for _ in self.drop])
