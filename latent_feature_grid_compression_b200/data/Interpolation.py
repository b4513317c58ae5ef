"""Ground-truth lookup -- counterpart of the reference's data/Interpolation.py.

``trilinear_f_interpolation`` keeps the reference's signature and numerics (fp32 lattice coordinates, fp64 alphas,
``f[x, y, z]`` indexing, lerp order x -> y -> z) but runs as one CUDA launch (``lfgc_trilinear``) instead of ~40
ATen ops with 8 advanced-index gathers.  At the integer voxel positions the training loop passes it is an exact
voxel lookup.
"""
from __future__ import annotations

import torch

from .. import ops


def trilinear_f_interpolation(p, f, min_bb, max_bb, res):
    """p (N,3) positions, f (R0,R1,R2) volume, bounding box and resolution as in the reference -> (N,) values."""
    if not f.is_cuda:
        raise ops.L.LfgcError('trilinear_f_interpolation runs on CUDA only (no CPU fallback)')
    if tuple(int(v) for v in res.tolist()) != tuple(f.shape):
        raise ValueError('res %s does not match the volume shape %s' % (res.tolist(), tuple(f.shape)))
    p = p.detach().to(f.device, torch.float32).reshape(-1, 3).contiguous()
    return ops.trilinear(p, f.contiguous().float(), min_bb.tolist(), max_bb.tolist())


def finite_difference_trilinear_grad(p, f, min_bb, max_bb, res, scale=None):
    """Central differences of the interpolated volume (dead code in the reference, Interpolation.py:47-87)."""
    step = (max_bb - min_bb) / (res - 1)
    grads = []
    for a in range(3):
        e = torch.zeros(3, device=p.device, dtype=p.dtype)
        e[a] = step[a]
        hi = torch.minimum(p + e, max_bb.to(p.device).unsqueeze(0))
        lo = torch.maximum(p - e, min_bb.to(p.device).unsqueeze(0))
        span = 2 * (hi[:, a] - lo[:, a]) / (max_bb[a] - min_bb[a])
        if scale is not None:
            span = span * scale[a]
        grads.append((trilinear_f_interpolation(hi, f, min_bb, max_bb, res)
                      - trilinear_f_interpolation(lo, f, min_bb, max_bb, res)) / span)
    return torch.stack(grads, dim=1)
