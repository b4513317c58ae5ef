"""Time lfgc_train_step (fused sampler+fwd+loss+bwd) at the bench config for different CTA widths (tuning aid)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import bench
from latent_feature_grid_compression_b200 import ops
from latent_feature_grid_compression_b200.model.model_utils import setup_model

dev = torch.device('cuda', 0)
cfg = bench.CFG
R = int(os.environ.get('LFGC_PROFILE_R', '255'))
volume = bench.synthetic_volume(R, dev)
torch.manual_seed(0)
C = int(os.environ.get('LFGC_SWEEP_C', cfg['C']))
model = setup_model(3, cfg['H'], 1, cfg['L'], 'fourier', cfg['F'], '', 0.1, 0.9, cfg['wavelet'], C, cfg['G'], '')
model.to(dev).train()
geom = model.geometry()
coeffs = [f.detach().contiguous() for f in model.feature_grid]
grid_cl = ops.decode_fwd(geom, coeffs, [None] * len(coeffs))
mlp = model.mlp_flat()
ws = torch.empty(geom.backward_workspace_bytes // 4, device=dev)
gg = torch.zeros_like(grid_cl)
gm = torch.empty(geom.mlp_param_count, device=dev)
loss = torch.zeros(1, device=dev)
ref = None
for n in [int(v) for v in os.environ.get('LFGC_SWEEP_N', '32768,262144').split(',')]:
    for nw in [0, 4, 5, 6, 7, 8]:
        if nw:
            os.environ['LFGC_BWD_WARPS'] = str(nw)
        else:
            os.environ.pop('LFGC_BWD_WARPS', None)
        try:
            for _ in range(3):
                gg.zero_()
                ops.train_step(geom, volume, n, 1, 0, 1.0 / n, grid_cl, mlp, gg, gm, loss, ws)
            torch.cuda.synchronize()
        except Exception as e:
            print('n', n, 'warps', nw, 'FAILED', str(e)[:80])
            continue
        if ref is None:
            ref = (gm.clone(), gg.clone())
        reps = 50
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
        for a, b in evs:
            gg.zero_()
            a.record()
            ops.train_step(geom, volume, n, 1, 0, 1.0 / n, grid_cl, mlp, gg, gm, loss, ws)
            b.record()
        torch.cuda.synchronize()
        ms = sorted(a.elapsed_time(b) for a, b in evs)[reps // 2]
        print('n %7d warps %2d : %8.2f us  %7.1f Msamples/s  %5.2f TFLOP/s' %
              (n, nw, ms * 1e3, n / ms / 1e3, 23616 * n / ms / 1e9))
