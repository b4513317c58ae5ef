"""On-disk format compatibility with the reference (reference model/model_utils.py:120-332).

  * GPU: a ``binary_model_file`` + ``_mask.bnr`` pair WRITTEN BY THE REFERENCE (tests/golden/storage_ref.npz, made by
    tests/golden/make_storage_golden.py) is read by this repository's ``restore_model``; every tensor must equal what
    the reference's own ``restore_model`` produced from the same bytes, bit for bit.
  * CPU, build container only (needs /root/reference): this repository's writer -> the reference's reader, and
    byte-identical files when both writers use the same deterministic quantiser (tests/ref_storage_roundtrip.py).
"""
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, 'tests', 'golden')


def test_reader_helpers_on_the_reference_written_bytes(tmp_path):
    """Header, mask stream and code-book arithmetic of the reference-written file, decoded on the host only."""
    import struct
    from latent_feature_grid_compression_b200.model.model_utils import read_binary
    g = np.load(os.path.join(GOLD, 'storage_ref.npz'))
    raw = g['file_bytes'].tobytes()
    head = struct.unpack('9B', raw[:9])
    assert head == (4, 32, 3 + 12 + 4, 3, 1, 8, 15, 3, 4)
    nz = struct.unpack('3I', raw[9:21])
    zeros = struct.unpack('3I', raw[21:33])
    for i in range(3):
        t = g['orig.feature_grid.%d' % i]
        assert nz[i] == np.count_nonzero(t) and zeros[i] == t.size - np.count_nonzero(t)
    p = tmp_path / 'm_mask.bnr'
    p.write_bytes(g['mask_bytes'].tobytes())
    total = sum(nz) + sum(zeros)
    bits = read_binary(str(p), total)
    want = np.concatenate([(g['orig.feature_grid.%d' % i].reshape(-1) != 0) for i in range(3)])
    assert np.array_equal(np.frombuffer(bits.encode(), dtype=np.uint8)[:total] == ord('1'), want)


@pytest.mark.gpu
def test_restore_model_reads_the_reference_written_file(tmp_path):
    from latent_feature_grid_compression_b200.model.model_utils import restore_model
    g = np.load(os.path.join(GOLD, 'storage_ref.npz'))
    path = str(tmp_path / 'binary_model_file')
    open(path, 'wb').write(g['file_bytes'].tobytes())
    open(path + '_mask.bnr', 'wb').write(g['mask_bytes'].tobytes())
    m = restore_model(path)
    sd = m.state_dict()
    checked = 0
    for k in g.files:
        if not k.startswith('restored.') or k.startswith('restored.filter.'):
            continue
        got = sd[k[len('restored.'):]].detach().cpu().numpy()
        assert np.array_equal(got, g[k]), k          # centres[labels] look-ups and raw fp32: bit-exact
        checked += 1
    assert checked == 3 + 2 * 5
    # and the restored model runs: same output as the restored tensors pushed through a fresh model
    m.cuda().eval()
    tile = torch.rand(1, 4, 5, 6, 3, device='cuda') * 2 - 1
    with torch.no_grad():
        y = m(tile)
    assert tuple(y.shape) == (1, 4, 5, 6, 1) and bool(torch.isfinite(y).all())


@pytest.mark.skipif(not os.path.isdir('/root/reference'), reason='needs the reference checkout (build container)')
def test_reference_reader_reads_our_file_and_writers_agree_bytewise():
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'tests', 'ref_storage_roundtrip.py')], cwd='/tmp',
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    res = json.loads(out.stdout.strip().splitlines()[-1])
    assert res['file_identical'] and res['mask_identical']
    assert res['zero_pattern_kept']
    assert res['max_err_unquantised'] == 0.0
    assert res['max_relerr_quantised'] < 0.02        # 256-centre k-means of a few thousand values
