"""Tensor-level wrappers over the C ABI (torch supplies device memory and streams; all arithmetic is in liblfgc.so).

Every function takes/returns CUDA float32 tensors, launches on the current torch stream, never synchronises and
raises ``LfgcError`` on failure.  There is no CPU path.
"""
from __future__ import annotations

import ctypes as ct
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import _lib as L
from . import wavelets


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _p(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _req(t: torch.Tensor, name: str, dtype=torch.float32):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise L.LfgcError('%s must be a CUDA tensor (this framework has no CPU path)' % name)
    if t.dtype != dtype:
        raise L.LfgcError('%s must be %s, got %s' % (name, dtype, t.dtype))
    if not t.is_contiguous():
        raise L.LfgcError('%s must be contiguous' % name)
    return t


class Geometry:
    """Static shape of one model: what ``setup_model`` fixes (reference model/model_utils.py:23-59)."""

    def __init__(self, C: int, grid_shape: Sequence[int], H: int, n_layers: int, n_freq: int, wavelet: str,
                 coeff_dims: Sequence[Sequence[int]], shape_array):
        self.C = int(C)
        self.Cp = (self.C + 3) // 4 * 4
        self.G = tuple(int(g) for g in grid_shape)
        self.H, self.L, self.F = int(H), int(n_layers), int(n_freq)
        self.in0 = 3 + 6 * self.F + self.C
        self.wavelet = wavelet
        self.coeff_dims = [tuple(int(v) for v in d) for d in coeff_dims]
        self.shape_array = np.asarray(shape_array, dtype=np.int64).reshape(-1, 3)
        if len(self.coeff_dims) != self.shape_array.shape[0] + 1:
            raise ValueError('need one coefficient tensor more than synthesis levels')
        if len(self.coeff_dims) > L.MAX_LEVELS:
            raise L.LfgcError('too many wavelet levels (%d > %d)' % (len(self.coeff_dims), L.MAX_LEVELS))
        m = L.ModelDesc()
        m.C, m.Cp, m.H, m.L, m.F = self.C, self.Cp, self.H, self.L, self.F
        for a in range(3):
            m.G[a] = self.G[a]
        self.model_desc = m
        w = L.WaveletDesc()
        w.n_coeff = len(self.coeff_dims)
        w.C = self.C
        _, _, rec_lo, rec_hi = wavelets.filter_bank(wavelet)
        if len(rec_lo) > L.MAX_TAPS:
            raise L.LfgcError('wavelet %s has %d taps (> %d)' % (wavelet, len(rec_lo), L.MAX_TAPS))
        w.n_taps = len(rec_lo)
        for i, (a, b) in enumerate(zip(rec_lo, rec_hi)):
            w.rec_lo[i] = np.float32(a)  # the reference keeps the taps as fp32 buffers
            w.rec_hi[i] = np.float32(b)
        for l, d in enumerate(self.coeff_dims):
            for a in range(3):
                w.dims[l][a] = d[a]
                w.target[l][a] = int(self.shape_array[l - 1][a]) if l >= 1 else 0
        self.wavelet_desc = w
        lib = L.load()
        self.mlp_param_count = int(lib.lfgc_mlp_param_count(ct.byref(m)))
        self.decode_scratch_bytes = int(lib.lfgc_decode_scratch_bytes(ct.byref(w)))
        self.backward_workspace_bytes = int(lib.lfgc_backward_workspace_bytes(ct.byref(m)))

    def coeff_shape(self, l):
        d = self.coeff_dims[l]
        return (self.C, *d) if l == 0 else (self.C, 7, *d)

    def mask_shape(self, l):
        return self.coeff_shape(l)[1:]

    def mlp_shapes(self):
        """[(name, shape)] in packed order = state_dict order of net_layers.* then final_layer.*"""
        out = []
        for l in range(self.L):
            k = self.in0 if l == 0 else self.H
            out.append(('net_layers.%d.weight' % l, (self.H, k)))
            out.append(('net_layers.%d.bias' % l, (self.H,)))
        out.append(('final_layer.weight', (1, self.H)))
        out.append(('final_layer.bias', (1,)))
        return out


# --------------------------------------------------------------------------------------------------------------------
# mask layers
# --------------------------------------------------------------------------------------------------------------------

def mask_multiplier(mode: int, p0, p1=None, noise=None, threshold: float = 0.0, want_aux: bool = False):
    lib = L.load()
    _req(p0, 'p0')
    mult = torch.empty_like(p0)
    aux = torch.empty_like(p0) if want_aux else None
    L.check(lib.lfgc_mask_multiplier(mode, p0.numel(), _p(p0), _p(p1), _p(noise), float(threshold), _p(mult), _p(aux),
                                     _stream()), 'lfgc_mask_multiplier')
    return mult, aux


def mask_param_grad(mode: int, p0, p1, noise, gmult):
    lib = L.load()
    g0 = torch.empty_like(p0)
    g1 = torch.empty_like(p1) if p1 is not None else None
    L.check(lib.lfgc_mask_param_grad(mode, p0.numel(), _p(p0), _p(p1), _p(noise), _p(_req(gmult, 'gmult')), _p(g0),
                                     _p(g1), 0, _stream()), 'lfgc_mask_param_grad')
    return g0, g1


def smallify_ema(betas, ema, emavar, momentum: float):
    lib = L.load()
    L.check(lib.lfgc_smallify_ema(_p(_req(betas, 'betas')), _p(_req(ema, 'ema')), _p(_req(emavar, 'emavar')),
                                  betas.numel(), float(momentum), _stream()), 'lfgc_smallify_ema')


# --------------------------------------------------------------------------------------------------------------------
# wavelet analysis / synthesis
# --------------------------------------------------------------------------------------------------------------------

def dwt_level(x: torch.Tensor, wavelet: str):
    """One analysis level on (C, d0, d1, d2) -> (C, 8, e0, e1, e2)."""
    lib = L.load()
    _req(x, 'x')
    dec_lo, dec_hi, _, _ = wavelets.filter_bank(wavelet)
    n = len(dec_lo)
    lo = (ct.c_float * n)(*[np.float32(v) for v in dec_lo])
    hi = (ct.c_float * n)(*[np.float32(v) for v in dec_hi])
    e = (ct.c_int32 * 3)()
    d = L.int3(x.shape[1:])
    L.check(lib.lfgc_dwt_level(None, x.shape[0], d, n, lo, hi, None, e, None), 'lfgc_dwt_level(size)')
    out = torch.empty((x.shape[0], 8, e[0], e[1], e[2]), device=x.device, dtype=torch.float32)
    L.check(lib.lfgc_dwt_level(_p(x), x.shape[0], d, n, lo, hi, _p(out), e, _stream()), 'lfgc_dwt_level')
    return out


def decode_fwd(geom: Geometry, coeffs: Sequence[torch.Tensor], mults: Sequence[Optional[torch.Tensor]],
               scratch: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None,
               also_zero: Optional[torch.Tensor] = None):
    lib = L.load()
    dev = coeffs[0].device
    for i, c in enumerate(coeffs):
        _req(c, 'coeff[%d]' % i)
        if tuple(c.shape) != geom.coeff_shape(i):
            raise L.LfgcError('coeff[%d] has shape %s, expected %s' % (i, tuple(c.shape), geom.coeff_shape(i)))
    for i, m in enumerate(mults):
        if m is not None:
            _req(m, 'mult[%d]' % i)
    if scratch is None:
        scratch = torch.empty(max(geom.decode_scratch_bytes // 4, 4), device=dev, dtype=torch.float32)
    if out is None:
        out = torch.empty((*geom.G, geom.Cp), device=dev, dtype=torch.float32)
    L.check(lib.lfgc_decode_fwd(ct.byref(geom.wavelet_desc), L.ptr_array([_p(c) for c in coeffs]),
                                L.ptr_array([_p(m) for m in mults]), _p(scratch), _p(out), geom.Cp, _p(also_zero),
                                _stream()), 'lfgc_decode_fwd')
    return out


def decode_bwd(geom: Geometry, grad_grid_cl, coeffs, gmuls, want_gmult: Sequence[bool],
               scratch: Optional[torch.Tensor] = None, grad_coeffs=None, grad_mults=None, accumulate: int = 0):
    lib = L.load()
    dev = grad_grid_cl.device
    _req(grad_grid_cl, 'grad_grid_cl')
    if scratch is None:
        scratch = torch.empty(max(geom.decode_scratch_bytes // 4, 4), device=dev, dtype=torch.float32)
    if grad_coeffs is None:
        grad_coeffs = [torch.empty_like(c) for c in coeffs]
    if grad_mults is None:
        grad_mults = [torch.empty(geom.mask_shape(i), device=dev, dtype=torch.float32) if w else None
                      for i, w in enumerate(want_gmult)]
    L.check(lib.lfgc_decode_bwd(ct.byref(geom.wavelet_desc), _p(grad_grid_cl), geom.Cp,
                                L.ptr_array([_p(c) for c in coeffs]), L.ptr_array([_p(m) for m in gmuls]),
                                _p(scratch), L.ptr_array([_p(g) for g in grad_coeffs]),
                                L.ptr_array([_p(g) for g in grad_mults]), int(accumulate), _stream()), 'lfgc_decode_bwd')
    return grad_coeffs, grad_mults


# --------------------------------------------------------------------------------------------------------------------
# per-sample path
# --------------------------------------------------------------------------------------------------------------------

def sample_forward(geom: Geometry, coords, grid_cl, mlp_flat, clamp=False, out=None):
    lib = L.load()
    _req(coords, 'coords')
    n = coords.numel() // 3
    if out is None:
        out = torch.empty(n, device=coords.device, dtype=torch.float32)
    L.check(lib.lfgc_forward(ct.byref(geom.model_desc), _p(coords), n, _p(_req(grid_cl, 'grid_cl')),
                             _p(_req(mlp_flat, 'mlp')), _p(out), L.F_CLAMP if clamp else 0, _stream()), 'lfgc_forward')
    return out


def sample_backward(geom: Geometry, coords, grad_out, grid_cl, mlp_flat, grad_grid_cl=None, grad_mlp=None,
                    workspace=None, accumulate_mlp=False):
    lib = L.load()
    dev = coords.device
    n = coords.numel() // 3
    if grad_grid_cl is None:
        grad_grid_cl = torch.zeros((*geom.G, geom.Cp), device=dev, dtype=torch.float32)
    if grad_mlp is None:
        grad_mlp = torch.empty(geom.mlp_param_count, device=dev, dtype=torch.float32)
    if workspace is None:
        workspace = torch.empty(geom.backward_workspace_bytes // 4, device=dev, dtype=torch.float32)
    L.check(lib.lfgc_backward(ct.byref(geom.model_desc), _p(_req(coords, 'coords')), n, _p(_req(grad_out, 'grad_out')),
                              _p(_req(grid_cl, 'grid_cl')), _p(_req(mlp_flat, 'mlp')), _p(grad_grid_cl), _p(grad_mlp),
                              None, 1 if accumulate_mlp else 0, _p(workspace), workspace.numel() * 4, _stream()),
            'lfgc_backward')
    return grad_grid_cl, grad_mlp


def train_step(geom: Geometry, volume, n: int, seed: int, sample_offset: int, loss_scale: float, grid_cl, mlp_flat,
               grad_grid_cl, grad_mlp, loss_sum, workspace, explicit_idx=None, accumulate_mlp=False, step_dev=None,
               step_stride: int = 0, coords=None, targets=None, log_sigma=None, dlog_sigma=None):
    """Fused sampler + forward + MSE + backward.  With ``coords`` (n,3) and ``targets`` (n,) the caller supplies the
    samples (host-fed step) and ``volume`` may be None.  ``log_sigma`` (n,) switches the loss to the Gaussian log
    likelihood of VariationalDropoutLoss and ``dlog_sigma`` (n,) receives its gradient (lfgc_train_step_weighted)."""
    lib = L.load()
    if volume is not None:
        _req(volume, 'volume')
    if explicit_idx is not None:
        _req(explicit_idx, 'explicit_idx', torch.int64)
    if coords is not None:
        _req(coords, 'coords')
        _req(targets, 'targets')
    shape3 = L.int3(volume.shape) if volume is not None else None
    if log_sigma is not None:
        _req(log_sigma, 'log_sigma')
        if dlog_sigma is not None:
            _req(dlog_sigma, 'dlog_sigma')
        L.check(lib.lfgc_train_step_weighted(ct.byref(geom.model_desc), _p(volume), shape3, int(n), int(seed),
                                             int(sample_offset), _p(step_dev), int(step_stride), _p(explicit_idx),
                                             _p(coords), _p(targets), float(loss_scale), _p(log_sigma), _p(dlog_sigma),
                                             _p(_req(grid_cl, 'grid_cl')), _p(_req(mlp_flat, 'mlp')), _p(grad_grid_cl),
                                             _p(grad_mlp), _p(loss_sum), 1 if accumulate_mlp else 0, _p(workspace),
                                             workspace.numel() * 4, _stream()), 'lfgc_train_step_weighted')
        return
    L.check(lib.lfgc_train_step(ct.byref(geom.model_desc), _p(volume), shape3, int(n), int(seed),
                                int(sample_offset), _p(step_dev), int(step_stride), _p(explicit_idx), _p(coords),
                                _p(targets), float(loss_scale), _p(_req(grid_cl, 'grid_cl')),
                                _p(_req(mlp_flat, 'mlp')), _p(grad_grid_cl), _p(grad_mlp), _p(loss_sum),
                                1 if accumulate_mlp else 0, _p(workspace), workspace.numel() * 4, _stream()),
            'lfgc_train_step')


# --------------------------------------------------------------------------------------------------------------------
# sampler, ground truth, reconstruction, statistics, optimiser
# --------------------------------------------------------------------------------------------------------------------

def plain_mlp_param_count(H: int, L_: int) -> int:
    return 3 * H + H + (L_ - 1) * (H * H + H) + H + 1


def plain_mlp_forward(H: int, n_layers: int, x, mlp_flat, out=None):
    """Variance_Model forward (lfgc_plain_mlp_forward): x (n,3) -> (n,)."""
    lib = L.load()
    _req(x, 'x')
    _req(mlp_flat, 'mlp')
    n = x.shape[0]
    if out is None:
        out = torch.empty(n, device=x.device, dtype=torch.float32)
    L.check(lib.lfgc_plain_mlp_forward(int(H), int(n_layers), _p(x), int(n), _p(mlp_flat), _p(out), _stream()),
            'lfgc_plain_mlp_forward')
    return out


def plain_mlp_workspace_floats(H: int, n_layers: int) -> int:
    return int(L.load().lfgc_plain_mlp_workspace_bytes(int(H), int(n_layers))) // 4


def plain_mlp_backward(H: int, n_layers: int, x, grad_out, mlp_flat, grad_mlp=None, accumulate=False, workspace=None):
    """Parameter gradients of the Variance_Model for d(loss)/d(out) = grad_out (lfgc_plain_mlp_backward)."""
    lib = L.load()
    _req(x, 'x')
    _req(grad_out, 'grad_out')
    _req(mlp_flat, 'mlp')
    if grad_mlp is None:
        grad_mlp = torch.empty(plain_mlp_param_count(H, n_layers), device=x.device, dtype=torch.float32)
    if workspace is None:
        workspace = torch.empty(plain_mlp_workspace_floats(H, n_layers), device=x.device, dtype=torch.float32)
    L.check(lib.lfgc_plain_mlp_backward(int(H), int(n_layers), _p(x), int(x.shape[0]), _p(grad_out), _p(mlp_flat),
                                        _p(grad_mlp), 1 if accumulate else 0, _p(workspace), workspace.numel() * 4,
                                        _stream()), 'lfgc_plain_mlp_backward')
    return grad_mlp


def variational_dkl_grad(mask_params, mask_grads, layer_sizes, w_dkl, step_dev, ramp: float, w_max: float, scale: float):
    """grad += w_dkl * scale * d DKL/d(log_thetas, log_var), with the per-step ramp of w_dkl on the device
    (lfgc_variational_dkl_grad).  ``w_dkl`` is a float64[2] device tensor."""
    lib = L.load()
    _req(mask_params, 'mask_params')
    _req(mask_grads, 'mask_grads')
    _req(w_dkl, 'w_dkl', torch.float64)
    sizes = (ct.c_int64 * len(layer_sizes))(*[int(v) for v in layer_sizes])
    L.check(lib.lfgc_variational_dkl_grad(_p(mask_params), _p(mask_grads), len(layer_sizes), sizes, _p(w_dkl),
                                          _p(step_dev), float(ramp), float(w_max), float(scale), _stream()),
            'lfgc_variational_dkl_grad')


def variational_multiplier(mask_params, noise, layer_sizes, mult_out, zero_out=None):
    """Multipliers of all live variational mask layers in one launch (lfgc_variational_multiplier)."""
    sizes = (ct.c_int64 * len(layer_sizes))(*[int(v) for v in layer_sizes])
    L.check(L.load().lfgc_variational_multiplier(_p(_req(mask_params, 'mask_params')), _p(_req(noise, 'noise')),
                                                 len(layer_sizes), sizes, _p(_req(mult_out, 'mult_out')), _p(zero_out),
                                                 _stream()), 'lfgc_variational_multiplier')


def variational_param_grad(mask_params, noise, gmult, layer_sizes, mask_grads):
    """d loss / d(log_thetas, log_var) of all live variational mask layers in one launch, written into the flat mask
    gradient section (lfgc_variational_param_grad)."""
    sizes = (ct.c_int64 * len(layer_sizes))(*[int(v) for v in layer_sizes])
    L.check(L.load().lfgc_variational_param_grad(_p(_req(mask_params, 'mask_params')), _p(_req(noise, 'noise')),
                                                 _p(_req(gmult, 'gmult')), len(layer_sizes), sizes,
                                                 _p(_req(mask_grads, 'mask_grads')), _stream()),
            'lfgc_variational_param_grad')


def sample(volume_shape, n: int, seed: int = 0, sample_offset: int = 0, volume=None, explicit_idx=None,
           want_raw=True, want_norm=True, want_gt=False, device=None, step_dev=None, step_stride: int = 0, out=None):
    lib = L.load()
    device = device or (volume.device if volume is not None else torch.device('cuda'))
    if out is not None:
        raw, norm, gt = out
    else:
        raw = torch.empty((n, 3), device=device, dtype=torch.float32) if want_raw else None
        norm = torch.empty((n, 3), device=device, dtype=torch.float32) if want_norm else None
        gt = torch.empty(n, device=device, dtype=torch.float32) if want_gt else None
    if explicit_idx is not None:
        _req(explicit_idx, 'explicit_idx', torch.int64)
    if volume is not None:
        _req(volume, 'volume')
    if step_dev is not None:
        L.check(lib.lfgc_sample_stream(_p(volume), L.int3(volume_shape), int(n), int(seed), int(sample_offset),
                                       _p(step_dev), int(step_stride), _p(explicit_idx), _p(raw), _p(norm), _p(gt),
                                       _stream()), 'lfgc_sample_stream')
        return raw, norm, gt
    L.check(lib.lfgc_sample(_p(volume), L.int3(volume_shape), int(n), int(seed), int(sample_offset), _p(explicit_idx),
                            _p(raw), _p(norm), _p(gt), _stream()), 'lfgc_sample')
    return raw, norm, gt


def trilinear(p, volume, min_bb, max_bb):
    lib = L.load()
    _req(p, 'p')
    _req(volume, 'volume')
    n = p.shape[0]
    out = torch.empty(n, device=p.device, dtype=torch.float32)
    L.check(lib.lfgc_trilinear(_p(p), n, _p(volume), L.int3(volume.shape), L.float3(min_bb), L.float3(max_bb),
                               _p(out), _stream()), 'lfgc_trilinear')
    return out


def reconstruct(geom: Geometry, grid_cl, mlp_flat, vol_shape, axes, slab_begin: int, slab_end: int, clamp=True,
                out=None):
    lib = L.load()
    dev = grid_cl.device
    if out is None:
        out = torch.empty((slab_end - slab_begin, vol_shape[1], vol_shape[2]), device=dev, dtype=torch.float32)
    for a, R in zip(axes, vol_shape):
        _req(a, 'axis')
        if a.numel() != R:
            raise L.LfgcError('axis table length %d != extent %d' % (a.numel(), R))
    L.check(lib.lfgc_reconstruct(ct.byref(geom.model_desc), _p(_req(grid_cl, 'grid_cl')), _p(_req(mlp_flat, 'mlp')),
                                 L.int3(vol_shape), _p(axes[0]), _p(axes[1]), _p(axes[2]), int(slab_begin),
                                 int(slab_end), _p(out), L.F_CLAMP if clamp else 0, _stream()), 'lfgc_reconstruct')
    return out


def deviation_stats_accumulate(pred, gt, acc):
    """acc: float64[4] = [sum sq, sum abs, max gt, min gt] (initialise to [0, 0, -inf, +inf])."""
    lib = L.load()
    _req(acc, 'acc', torch.float64)
    L.check(lib.lfgc_deviation_stats(_p(_req(pred, 'pred')), _p(_req(gt, 'gt')), pred.numel(), _p(acc), _stream()),
            'lfgc_deviation_stats')


def launch_count() -> int:
    """Kernels launched through the C ABI so far in this process."""
    return int(L.load().lfgc_launch_count())


def adam(p, g, m, v, lr_dev, step_dev, beta1=0.9, beta2=0.999, eps=1e-8, grad_scale=1.0):
    lib = L.load()
    _req(step_dev, 'step', torch.int32)
    L.check(lib.lfgc_adam(_p(_req(p, 'p')), _p(_req(g, 'g')), _p(_req(m, 'm')), _p(_req(v, 'v')), p.numel(),
                          _p(_req(lr_dev, 'lr')), _p(step_dev), beta1, beta2, eps, grad_scale, _stream()), 'lfgc_adam')


def adam_reg(p, g, m, v, lr_dev, step_dev, l2_range, weight_l2, l1_range, weight_l1, beta1=0.9, beta2=0.999, eps=1e-8,
             ema=None, emavar=None, momentum=0.0, zero_l1_grad=False):
    """Adam with the SmallifyLoss gradient terms folded in (lfgc_adam_reg): + 2 weight_l2 p on ``l2_range`` (begin, end),
    + weight_l1 sign(p) on ``l1_range``; optionally the Smallify tracker update (``ema``, ``emavar``) and the clearing of
    the gradient on ``l1_range`` in the same pass."""
    lib = L.load()
    _req(step_dev, 'step', torch.int32)
    L.check(lib.lfgc_adam_reg(_p(_req(p, 'p')), _p(_req(g, 'g')), _p(_req(m, 'm')), _p(_req(v, 'v')), p.numel(),
                              _p(_req(lr_dev, 'lr')), _p(step_dev), beta1, beta2, eps, 1.0, int(l2_range[0]),
                              int(l2_range[1]), float(weight_l2), int(l1_range[0]), int(l1_range[1]), float(weight_l1),
                              _p(ema), _p(emavar), float(momentum), 1 if zero_l1_grad else 0, _stream()), 'lfgc_adam_reg')


def train_step_partials(geom: Geometry, volume, n: int, seed: int, sample_offset: int, loss_scale: float, grid_cl,
                        mlp_flat, grad_grid_cl, workspace, step_dev=None, step_stride: int = 0, coords=None,
                        targets=None, explicit_idx=None, tc_panels=None) -> int:
    """``train_step`` that leaves the MLP-gradient partial sums (and the loss partials) in ``workspace`` for
    ``grid_step`` to reduce; returns the number of partial rows (lfgc_train_step_partials)."""
    lib = L.load()
    if volume is not None:
        _req(volume, 'volume')
    if coords is not None:
        _req(coords, 'coords')
        _req(targets, 'targets')
    if explicit_idx is not None:
        _req(explicit_idx, 'explicit_idx', torch.int64)
    shape3 = L.int3(volume.shape) if volume is not None else None
    ns = ct.c_int32(0)
    L.check(lib.lfgc_train_step_partials(ct.byref(geom.model_desc), _p(volume), shape3, int(n), int(seed),
                                         int(sample_offset), _p(step_dev), int(step_stride), _p(explicit_idx),
                                         _p(coords), _p(targets), float(loss_scale), _p(_req(grid_cl, 'grid_cl')),
                                         _p(_req(mlp_flat, 'mlp')), _p(_req(grad_grid_cl, 'grad_grid_cl')),
                                         _p(tc_panels), _p(_req(workspace, 'workspace')), workspace.numel() * 4,
                                         ct.byref(ns), _stream()), 'lfgc_train_step_partials')
    return int(ns.value)


def train_step_accumulate(geom: Geometry, volume, n: int, seed: int, sample_offset: int, loss_scale: float, grid_cl,
                          mlp_flat, grad_grid_cl, grad_mlp_loss, workspace, step_dev=None, step_stride: int = 0,
                          coords=None, targets=None, explicit_idx=None, announce=None, n_slices: int = 1, tc_panels=None):
    """``train_step`` that ADDS its MLP-gradient sums and loss sum to ``grad_mlp_loss`` (mlp_param_count + 1 floats, a
    per row, ``n_slices`` rows whose SUM is the result: a running sum the caller cleared) -- atomics from the tensor-core
    kernel's epilogue, its CTAs spread over the rows, no reduction launch
    (lfgc_train_step_accumulate).  ``announce``: a ``peer_announce(...)`` block; the kernel's last CTA then stores this
    rank's epoch flags for ``peer_sum(..., announced=True)``."""
    lib = L.load()
    if volume is not None:
        _req(volume, 'volume')
    if coords is not None:
        _req(coords, 'coords')
        _req(targets, 'targets')
    if explicit_idx is not None:
        _req(explicit_idx, 'explicit_idx', torch.int64)
    if grad_mlp_loss.numel() < n_slices * (geom.mlp_param_count + 1):
        raise L.LfgcError('train_step_accumulate: grad_mlp_loss holds %d floats, %d needed'
                          % (grad_mlp_loss.numel(), n_slices * (geom.mlp_param_count + 1)))
    shape3 = L.int3(volume.shape) if volume is not None else None
    L.check(lib.lfgc_train_step_accumulate(ct.byref(geom.model_desc), _p(volume), shape3, int(n), int(seed),
                                           int(sample_offset), _p(step_dev), int(step_stride), _p(explicit_idx),
                                           _p(coords), _p(targets), float(loss_scale), _p(_req(grid_cl, 'grid_cl')),
                                           _p(_req(mlp_flat, 'mlp')), _p(_req(grad_grid_cl, 'grad_grid_cl')),
                                           _p(_req(grad_mlp_loss, 'grad_mlp_loss')), int(n_slices),
                                           ct.byref(announce) if announce is not None else None, _p(tc_panels),
                                           _p(_req(workspace, 'workspace')), workspace.numel() * 4, _stream()),
            'lfgc_train_step_accumulate')


def peer_announce(flag_addrs, rank: int, epoch, ticket):
    """lfgc_peer_announce block: the flag arrays of all ranks (raw device addresses), this rank, the epoch counter peer_sum
    maintains and a zeroed int32 ticket.  Keep the tensors alive as long as the block is in use."""
    _req(epoch, 'epoch', torch.int32)
    _req(ticket, 'ticket', torch.int32)
    a = L.PeerAnnounce()
    a.n_peers = len(flag_addrs)
    a.rank = int(rank)
    for i, addr in enumerate(flag_addrs):
        a.flags[i] = int(addr)
    a.epoch = epoch.data_ptr()
    a.ticket = ticket.data_ptr()
    return a


def peer_sum(src_addrs, flag_addrs, rank: int, epoch, out, zero=None, announced=False):
    """out = sum over the ranks' buffers (raw device addresses, peer memory) behind the in-kernel barrier
    (lfgc_peer_sum); ``epoch``: int32[2] device tensor; ``zero``: buffer cleared in the same pass."""
    lib = L.load()
    _req(epoch, 'epoch', torch.int32)
    _req(out, 'out')
    L.check(lib.lfgc_peer_sum(L.ptr_array([int(a) for a in src_addrs]), L.ptr_array([int(a) for a in flag_addrs]),
                              len(src_addrs), int(rank), _p(epoch), _p(out), _p(zero), out.numel(),
                              1 if announced else 0, _stream()),
            'lfgc_peer_sum')


def grid_step_supported(geom: Geometry) -> bool:
    """True when lfgc_grid_step covers this model's wavelet pyramid (lfgc_grid_step_supported)."""
    return int(L.load().lfgc_grid_step_supported(ct.byref(geom.wavelet_desc))) == 1


def grid_step_scratch_floats(geom: Geometry) -> int:
    """Size of the scratch that lets lfgc_grid_step run the finest wavelet level on the whole GPU (split path)."""
    return int(L.load().lfgc_grid_step_scratch_bytes(ct.byref(geom.wavelet_desc))) // 4


def grid_step(geom: Geometry, grad_grids, mlp_partials, nslices: int, pstride: int, pcount: int, grid_cl, p, g, m, v,
              coeff_offs, mlp_off: int, lr_dev, step_dev, zero_grid=None, loss_out=None, beta1=0.9, beta2=0.999,
              eps=1e-8, grad_scale=1.0, weight_l2=0.0, scratch=None, tc_panels=None):
    """Everything of one optimiser step that is not per-sample, one launch (lfgc_grid_step): partial reduction +
    synthesis adjoint + Adam + synthesis of the updated coefficients.  ``grad_grids`` / ``mlp_partials``: lists (one
    entry per gradient source: this rank, or every data-parallel rank in rank order) of tensors or raw device
    addresses.  ``scratch`` (grid_step_scratch_floats) enables the split path: finest level on the whole GPU.
    ``tc_panels`` (tc_panel_image): kept current with the MLP parameters this step updates."""
    lib = L.load()
    _req(step_dev, 'step', torch.int32)
    a = L.GridStepArgs()
    a.n_srcs = len(grad_grids)
    addr = lambda t: int(t) if isinstance(t, int) else _p(t)
    for r, (gg, mp) in enumerate(zip(grad_grids, mlp_partials)):
        a.grad_grid[r] = addr(gg)
        a.mlp_partials[r] = addr(mp) if mp is not None else None
    a.nslices, a.pstride, a.pcount = int(nslices), int(pstride), int(pcount)
    a.zero_grid = _p(zero_grid)
    a.grid_cl = _p(_req(grid_cl, 'grid_cl'))
    a.p, a.g, a.m, a.v = _p(_req(p, 'p')), _p(_req(g, 'g')), _p(_req(m, 'm')), _p(_req(v, 'v'))
    for l, o in enumerate(coeff_offs):
        a.coeff_off[l] = int(o)
    a.mlp_off = int(mlp_off)
    a.loss_out = _p(loss_out)
    a.lr = _p(_req(lr_dev, 'lr'))
    a.step_count = _p(step_dev)
    a.beta1, a.beta2, a.eps, a.grad_scale, a.weight_l2 = float(beta1), float(beta2), float(eps), float(grad_scale), \
        float(weight_l2)
    if scratch is not None:
        a.scratch = _p(_req(scratch, 'scratch'))
        a.scratch_bytes = scratch.numel() * 4
    if tc_panels is not None:
        a.panel_model = ct.pointer(geom.model_desc)
        a.panel_image = _p(_req(tc_panels, 'tc_panels'))
    L.check(lib.lfgc_grid_step(ct.byref(geom.wavelet_desc), geom.Cp, ct.byref(a), _stream()), 'lfgc_grid_step')


def tc_panel_image(geom: Geometry, mlp_flat, out=None):
    """Operand image of the tensor-core training kernel for the packed MLP block ``mlp_flat`` (lfgc_tc_panel_build); None
    when the tensor-core kernel does not cover the model.  ``out``: rebuild into an existing image."""
    lib = L.load()
    nbytes = int(lib.lfgc_tc_panel_bytes(ct.byref(geom.model_desc)))
    if nbytes == 0:
        return None
    if out is None:
        out = torch.zeros((nbytes + 3) // 4, device=mlp_flat.device, dtype=torch.float32)
    L.check(lib.lfgc_tc_panel_build(ct.byref(geom.model_desc), _p(_req(mlp_flat, 'mlp')), _p(_req(out, 'image')), _stream()),
            'lfgc_tc_panel_build')
    return out


def add_l2_grad(g, p, weight: float):
    L.check(L.load().lfgc_add_l2_grad(_p(_req(g, 'g')), _p(_req(p, 'p')), p.numel(), float(weight), _stream()),
            'lfgc_add_l2_grad')


def add_l1_grad(g, p, weight: float):
    L.check(L.load().lfgc_add_l1_grad(_p(_req(g, 'g')), _p(_req(p, 'p')), p.numel(), float(weight), _stream()),
            'lfgc_add_l1_grad')


# --------------------------------------------------------------------------------------------------------------------
# parameter packing: nn.Parameters as views into one flat buffer
# --------------------------------------------------------------------------------------------------------------------

class FlatPack:
    """Keeps a list of parameters as contiguous views into one flat fp32 buffer (what the kernels read).

    ``ensure(params)`` re-packs whenever the parameters stopped being views of the buffer (``module.to(device)``,
    ``parameters.data = ...`` as in the reference's restore_model, ParameterList replacement)."""

    def __init__(self):
        self.flat = None
        self._layout = None

    def ensure(self, params: List[torch.Tensor]) -> torch.Tensor:
        ok = self.flat is not None and self._layout is not None and len(params) == len(self._layout)
        if ok:
            base = self.flat.data_ptr()
            for p, (off, n) in zip(params, self._layout):
                if p.data_ptr() != base + 4 * off or p.numel() != n or p.device != self.flat.device \
                        or not p.is_contiguous():
                    ok = False
                    break
        if ok:
            return self.flat
        with torch.no_grad():
            dev = params[0].device
            flat = torch.cat([p.detach().reshape(-1).to(device=dev, dtype=torch.float32) for p in params])
            layout, off = [], 0
            for p in params:
                n = p.numel()
                p.data = flat[off:off + n].view(p.shape)
                layout.append((off, n))
                off += n
        self.flat, self._layout = flat, layout
        return flat
