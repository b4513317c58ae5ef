"""Test infrastructure (run by tests/test_storage_compat.py in a subprocess, build container only): a model built by
the REFERENCE is written by THIS repository's ``store_model_parameters`` and read by the REFERENCE's ``restore_model``
(model/model_utils.py:222-332), and the other way round at byte level: with the same deterministic quantiser plugged
into both writers the two ``binary_model_file`` / ``_mask.bnr`` pairs must be byte-identical."""
import json
import os
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_harness  # noqa: E402

ref_harness.install()
import model.model_utils as REF  # noqa: E402  (the reference's module)
from latent_feature_grid_compression_b200.model import model_utils as OURS  # noqa: E402


def uniform_quantiser(w, q):
    """Deterministic stand-in for KMeans in BOTH writers: q equally spaced centres over [min, max]."""
    w = np.asarray(w, dtype=np.float64).reshape(-1)
    lo, hi = float(w.min()), float(w.max())
    centres = np.linspace(lo, hi, q)
    labels = np.clip(np.rint((w - lo) / max(hi - lo, 1e-30) * (q - 1)), 0, q - 1).astype(np.int64)
    return labels.tolist(), centres.tolist()


def main():
    torch.manual_seed(3)
    np.random.seed(3)
    m = REF.setup_model(3, 32, 1, 4, 'fourier', 2, '', 0.025, 0.75, 'db2', 4, 15, '')
    with torch.no_grad():
        for f in m.feature_grid:
            f[torch.rand_like(f) < 0.35] = 0.0
    work = tempfile.mkdtemp()
    res = {}
    # (1) our writer (sklearn k-means, as shipped) -> the reference's reader
    p_ours = os.path.join(work, 'ours')
    OURS.store_model_parameters(m, p_ours)
    r = REF.restore_model(p_ours)
    sd0, sd1 = m.state_dict(), r.state_dict()
    exact, quant = {}, {}
    for k in sd0:
        if k.startswith('filter.'):
            continue
        a, b = sd0[k].numpy(), sd1[k].numpy()
        d = float(np.abs(a - b).max())
        unquantised = k.endswith('.bias') or k.startswith('net_layers.0.') or k.startswith('final_layer.')
        (exact if unquantised else quant)[k] = d / max(float(np.abs(a).max()), 1e-30)
        if k.startswith('feature_grid.'):
            res.setdefault('zero_pattern_kept', True)
            res['zero_pattern_kept'] &= bool(np.array_equal(a == 0.0, b == 0.0))
    res['max_err_unquantised'] = max(exact.values())
    res['max_relerr_quantised'] = max(quant.values())
    # (2) byte-level: both writers with the same deterministic quantiser
    REF.kmeans_quantization, OURS.kmeans_quantization = uniform_quantiser, uniform_quantiser
    p_a, p_b = os.path.join(work, 'ref_det'), os.path.join(work, 'ours_det')
    REF.store_model_parameters(m, p_a)
    OURS.store_model_parameters(m, p_b)
    res['file_identical'] = open(p_a, 'rb').read() == open(p_b, 'rb').read()
    res['mask_identical'] = open(p_a + '_mask.bnr', 'rb').read() == open(p_b + '_mask.bnr', 'rb').read()
    res['file_bytes'] = os.path.getsize(p_a)
    print(json.dumps(res))


if __name__ == '__main__':
    main()
