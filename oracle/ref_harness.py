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

import argparse
import math
import sys
import types

REFERENCE_ROOT = "/root/reference"


# ----------------------------------------------------------------------------
# pywt stand-in
# ----------------------------------------------------------------------------
def _db_rec_lo(name: str):
    s2 = math.sqrt(2.0)
    s3 = math.sqrt(3.0)
    if name in ("haar", "db1"):
        return [1.0 / s2, 1.0 / s2]
    if name == "db2":
        return [(1 + s3) / (4 * s2), (3 + s3) / (4 * s2), (3 - s3) / (4 * s2), (1 - s3) / (4 * s2)]
    raise ValueError("pywt stand-in knows haar/db1/db2 only, got %r" % (name,))


class _Wavelet:
    """Just enough of ``pywt.Wavelet``: ``filter_bank = (dec_lo, dec_hi, rec_lo, rec_hi)``."""

    def __init__(self, name):
        self.name = name
        rec_lo = _db_rec_lo(name)
        dec_lo = rec_lo[::-1]
        rec_hi = [((-1) ** i) * dec_lo[i] for i in range(len(dec_lo))]
        dec_hi = rec_hi[::-1]
        self.filter_bank = (dec_lo, dec_hi, rec_lo, rec_hi)
        self.dec_len = len(dec_lo)


def _dwt_max_level(data_len, filter_len):
    if isinstance(filter_len, _Wavelet):
        filter_len = filter_len.dec_len
    if filter_len < 2:
        raise ValueError("bad filter length")
    if data_len < filter_len - 1:
        return 0
    return int(math.floor(math.log2(data_len / (filter_len - 1.0))))


def _install_stub(name, **attrs):
    mod = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(mod, k, v)
    sys.modules[name] = mod
    return mod


class _ConfigArgParser(argparse.ArgumentParser):
    def add_argument(self, *a, **kw):
        kw.pop("is_config_file", None)
        return super().add_argument(*a, **kw)


def install(reference_root: str = REFERENCE_ROOT):
    """Register the stand-ins and put the reference root on ``sys.path``."""
    if "pywt" not in sys.modules:
        _install_stub("pywt", Wavelet=_Wavelet, dwt_max_level=_dwt_max_level)
    if "pyevtk" not in sys.modules:
        hl = _install_stub("pyevtk.hl", imageToVTK=lambda *a, **k: None)
        _install_stub("pyevtk", hl=hl)
    if "configargparse" not in sys.modules:
        _install_stub("configargparse", ArgumentParser=_ConfigArgParser)
    if "mlflow" not in sys.modules:
        tr = _install_stub("mlflow.tracking", MlflowClient=object)
        _install_stub("mlflow", tracking=tr)
    try:  # matplotlib is only imported at module level by pltUtils
        import matplotlib  # noqa: F401
    except Exception:
        pp = _install_stub("matplotlib.pyplot")
        tk = _install_stub("matplotlib.ticker", FormatStrFormatter=object)
        _install_stub("matplotlib", pyplot=pp, ticker=tk)
    if reference_root not in sys.path:
        sys.path.insert(0, reference_root)


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
