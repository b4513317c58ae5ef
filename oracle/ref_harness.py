"""Harness that makes the UNMODIFIED reference importable in the build container.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package may import this file.
It is used by ``tests/golden/make_golden.py`` (run in the build container, where
``/root/reference`` is mounted) to produce the committed golden fixtures that pin
the numpy oracle (``oracle/fvsrn_numpy.py``).  ``/root/reference`` does not exist
on the GPU box, so nothing here may be touched by ``-m gpu`` tests, ``smoke()``
or ``bench.py``.

The reference needs five Python packages that are absent from this image
(SURVEY.md section 8c).  We register minimal stand-ins in ``sys.modules`` instead
of editing the reference:

* ``pywt``  (PyWavelets 1.4.1, Env.txt:178) -- only ``Wavelet(name).filter_bank``
  and ``dwt_max_level`` are used (model/Feature_Grid_Model.py:85,
  wavelet_transform/Torch_Wavelet_Transform.py:14,41).  The taps are the
  published Daubechies coefficients (closed form for db1/db2).
* ``pyevtk.hl.imageToVTK`` (visualization/OutputToVTK.py:3) -- no-op.
* ``configargparse`` (Feature_Grid_Training.py:5) -- argparse subclass.
* ``mlflow`` / ``matplotlib`` (visualization/pltUtils.py:1-3) -- empty stubs.

One run-time patch is applied (not a source edit): under torch >= 2 the eval
branch ``x.view(orig_shape[0:-1], 1)`` (model/Feature_Grid_Model.py:78) raises
TypeError; ``patched_eval_forward`` re-implements that single line with the
intended ``x.view(*orig_shape[0:-1], 1)``.
"""
from __future__ import annotations

import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'baseline'))
from ref_shims import REFERENCE_ROOT, install  # noqa: E402,F401  (the stand-ins live in baseline/ref_shims.py)
sys.path.pop(0)


def patched_eval_forward(model, tile):
    """Reference eval forward with the one-line torch>=2 fix (Feature_Grid_Model.py:56-80)."""
    import torch
    from torch.nn import functional as F
    from model.Feature_Grid_Model import SnakeAlt

    assert not model.training
    grid_vol = model.decode_volume()
    orig_shape = tile.shape
    inp = tile.squeeze()
    inp = inp.view(inp.shape[0] * inp.shape[1] * inp.shape[2], inp.shape[3])
    g = inp.view(1, 1, 1, *inp.shape)
    feats = F.grid_sample(grid_vol.unsqueeze(0), g, mode="bilinear", align_corners=False).squeeze().transpose_(0, 1)
    x = torch.cat([inp, model.embedder.embed(inp), feats], -1)
    for layer in model.net_layers:
        x = SnakeAlt(layer(x))
    x = model.final_layer(x)
    return x.view(*orig_shape[0:-1], 1).clamp(-1, 1)
