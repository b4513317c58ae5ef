"""Orthogonal wavelet filter taps without PyWavelets.

The reference takes the 1-D taps from ``pywt.Wavelet(name).filter_bank`` (PyWavelets 1.4.1, Env.txt:178; call site
wavelet_transform/Torch_Wavelet_Transform.py:41) and one integer formula from ``pywt.dwt_max_level``
(model/Feature_Grid_Model.py:85).  PyWavelets is not part of this image, so the Daubechies filters are rebuilt here
from their definition (closed form for db1/db2, spectral factorisation of the Daubechies polynomial otherwise), in
pywt's ordering: ``rec_lo`` is the minimum-phase scaling filter, ``dec_lo = rec_lo[::-1]``,
``rec_hi[i] = (-1)^i dec_lo[i]``, ``dec_hi = rec_hi[::-1]``.
"""
from __future__ import annotations

import math
from functools import lru_cache

import numpy as np


@lru_cache(maxsize=None)
def _daubechies_rec_lo(order: int):
    if order == 1:
        return (1.0 / math.sqrt(2.0), 1.0 / math.sqrt(2.0))
    if order == 2:
        s2, s3 = math.sqrt(2.0), math.sqrt(3.0)
        return ((1 + s3) / (4 * s2), (3 + s3) / (4 * s2), (3 - s3) / (4 * s2), (1 - s3) / (4 * s2))
    if order > 8:
        raise ValueError('db%d needs %d taps; at most 16 taps are supported' % (order, 2 * order))
    # P(y) = sum_k C(N-1+k, k) y^k with y = (2 - z - 1/z) / 4; keep the roots inside the unit circle
    N = order
    coeffs = [math.comb(N - 1 + k, k) for k in range(N)]
    roots_y = np.roots(coeffs[::-1])
    zs = []
    for y in roots_y:
        b = 2.0 - 4.0 * y
        disc = np.sqrt(b * b - 4.0 + 0j)
        z1, z2 = (b + disc) / 2.0, (b - disc) / 2.0
        zs.append(z1 if abs(z1) < 1.0 else z2)
    poly = np.poly1d([1.0])
    for _ in range(N):
        poly = poly * np.poly1d([1.0, 1.0])
    for z in zs:
        poly = poly * np.poly1d([1.0, -z])
    h = np.real(poly.coeffs)
    h = h * (math.sqrt(2.0) / h.sum())
    return tuple(float(v) for v in h)


def filter_bank(name: str):
    """(dec_lo, dec_hi, rec_lo, rec_hi) as Python float lists, like ``pywt.Wavelet(name).filter_bank``."""
    name = name.lower()
    if name == 'haar':
        order = 1
    elif name.startswith('db') and name[2:].isdigit():
        order = int(name[2:])
    else:
        raise ValueError('unsupported wavelet %r (haar, db1..db8)' % (name,))
    rec_lo = list(_daubechies_rec_lo(order))
    dec_lo = rec_lo[::-1]
    rec_hi = [((-1) ** i) * dec_lo[i] for i in range(len(dec_lo))]
    dec_hi = rec_hi[::-1]
    return dec_lo, dec_hi, rec_lo, rec_hi


def dwt_max_level(data_len: int, filter_len: int) -> int:
    """``pywt.dwt_max_level``: floor(log2(n / (filter_len - 1))), 0 when the signal is shorter than the filter."""
    if filter_len < 2:
        raise ValueError('filter_len must be >= 2')
    if data_len < filter_len - 1:
        return 0
    return int(math.floor(math.log2(data_len / (filter_len - 1.0))))


def analysis_out_size(n: int, filter_len: int) -> int:
    """Output extent of one analysis level (Torch_Wavelet_Transform.py:59-67,87)."""
    pad = (2 * filter_len - 3) // 2
    return (n + 2 * pad + (n % 2) - filter_len) // 2 + 1
