"""Debug aid: per-phase cycle breakdown of the fused training kernel (needs a -DLFGC_PHASE_TIMING build:
   nvcc ... -DLFGC_PHASE_TIMING -o /tmp/liblfgc_pt.so ; LFGC_LIB=/tmp/liblfgc_pt.so python profiles/phase_timing.py)."""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import bench
from latent_feature_grid_compression_b200 import _lib, ops
from latent_feature_grid_compression_b200.model.model_utils import setup_model

lib = _lib.load()
lib.lfgc_phase_timing.argtypes = [ctypes.c_void_p, ctypes.c_int]
dev = torch.device('cuda', 0)
cfg = bench.CFG
volume = bench.synthetic_volume(255, dev)
torch.manual_seed(0)
model = setup_model(3, cfg['H'], 1, cfg['L'], 'fourier', cfg['F'], '', 0.1, 0.9, cfg['wavelet'], cfg['C'], cfg['G'], '')
model.to(dev).train()
geom = model.geometry()
coeffs = [f.detach().contiguous() for f in model.feature_grid]
grid_cl = ops.decode_fwd(geom, coeffs, [None] * len(coeffs))
mlp = model.mlp_flat()
ws = torch.empty(geom.backward_workspace_bytes // 4, device=dev)
gg = torch.zeros_like(grid_cl)
gm = torch.empty(geom.mlp_param_count, device=dev)
loss = torch.zeros(1, device=dev)
names = ['setup', 'input', 'forward', 'dz+bias', 'barrier', 'dW', 'dh/dfeat', 'scatter', 'tile barrier', 'flush']
for n in (32768, 262144):
    for nw in (7, 8):
        os.environ['LFGC_BWD_WARPS'] = str(nw)
        for _ in range(3):
            ops.train_step(geom, volume, n, 1, 0, 1.0 / n, grid_cl, mlp, gg, gm, loss, ws)
        buf = (ctypes.c_ulonglong * 16)()
        lib.lfgc_phase_timing(None, 1)
        reps = 10
        for _ in range(reps):
            ops.train_step(geom, volume, n, 1, 0, 1.0 / n, grid_cl, mlp, gg, gm, loss, ws)
        lib.lfgc_phase_timing(buf, 1)
        tiles = -(-n // (16 * nw))
        ctas = min(148, tiles)
        warps = ctas * nw
        tot = sum(buf[:10])
        print('n=%d warps/CTA=%d: mean cycles per warp per launch = %.0f' % (n, nw, tot / reps / warps))
        for i, nm in enumerate(names):
            print('   %-13s %9.0f  %5.1f%%' % (nm, buf[i] / reps / warps, 100.0 * buf[i] / tot))
