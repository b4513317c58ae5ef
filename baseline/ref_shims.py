"""Stand-ins for the third-party packages the UNMODIFIED reference imports but this image lacks:
``pywt`` (filter taps + dwt_max_level only), ``pyevtk.hl.imageToVTK`` (no-op), ``configargparse`` (argparse subclass),
``mlflow`` / ``matplotlib`` (empty stubs).  Used by oracle/ref_harness.py (golden generation in the build container) and
by bench.py's reference arms (baseline/_ref on the GPU box).  Never imported by the product package."""
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
