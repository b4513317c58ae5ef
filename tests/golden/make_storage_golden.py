"""On-disk format fixture written BY THE REFERENCE (model/model_utils.py:120-219 ``store_model_parameters``) and read
back BY THE REFERENCE (``restore_model``, :222-332), committed as tests/golden/storage_ref.npz:

    python tests/golden/make_storage_golden.py          # build container only (needs /root/reference)

Contents: the bytes of ``binary_model_file`` and ``binary_model_file_mask.bnr``, the model's tensors before the store,
and the tensors of the model the reference restored from those files.  tests/test_storage_compat.py feeds the same bytes
to this repository's ``restore_model`` (must reproduce the reference-restored tensors bit for bit).
"""
import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import ref_harness  # noqa: E402

ref_harness.install()
from model.model_utils import restore_model, setup_model, store_model_parameters  # noqa: E402


def main():
    torch.manual_seed(11)
    np.random.seed(11)                      # sklearn KMeans draws from numpy's global state
    # restore_model hard-codes fourier / 2 frequencies / db2 (model_utils.py:310-313): the fixture uses those
    model = setup_model(3, 32, 1, 4, 'fourier', 2, '', 0.025, 0.75, 'db2', 4, 15, '')
    with torch.no_grad():                   # prune ~40 % of the coefficients so the mask stream matters
        for f in model.feature_grid:
            f[torch.rand_like(f) < 0.4] = 0.0
    work = tempfile.mkdtemp()
    path = os.path.join(work, 'binary_model_file')
    store_model_parameters(model, path)
    restored = restore_model(path)
    out = dict(file_bytes=np.frombuffer(open(path, 'rb').read(), dtype=np.uint8),
               mask_bytes=np.frombuffer(open(path + '_mask.bnr', 'rb').read(), dtype=np.uint8))
    for k, v in model.state_dict().items():
        out['orig.' + k] = v.detach().numpy().copy()
    for k, v in restored.state_dict().items():
        out['restored.' + k] = v.detach().numpy().copy()
    np.savez_compressed(os.path.join(HERE, 'storage_ref.npz'), **out)
    print('file %d B, mask %d B' % (out['file_bytes'].size, out['mask_bytes'].size))


if __name__ == '__main__':
    main()
