import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from latent_feature_grid_compression_b200 import ops
from latent_feature_grid_compression_b200.model.model_utils import setup_model
torch.manual_seed(0)
for (C, G, H, n) in [(16, 15, 32, 1000), (16, 15, 32, 262144), (8, 15, 20, 5000), (32, 17, 32, 4096), (6, 15, 32, 300)]:
    m = setup_model(3, H, 1, 4, 'fourier', 2, '', 0.1, 0.9, 'db2', C, G, '').cuda()
    with torch.no_grad():
        for lyr in m.net_layers:
            lyr.bias.uniform_(-0.5, 0.5)
    geom = m.geometry()
    grid = ops.decode_fwd(geom, [f.detach().contiguous() for f in m.feature_grid], [None] * len(m.feature_grid))
    mlp = m.mlp_flat()
    coords = (torch.rand(n, 3, device='cuda') * 2.2 - 1.1)
    os.environ['LFGC_FORWARD_TC'] = '0'
    ref = ops.sample_forward(geom, coords, grid, mlp)
    os.environ['LFGC_FORWARD_TC'] = '1'
    out = ops.sample_forward(geom, coords, grid, mlp)
    torch.cuda.synchronize()
    err = float((out - ref).abs().max() / ref.abs().max())
    print('C%d G%d H%d n%d: rel err tc vs ffma = %.3e  (ref max %.3f)' % (C, G, H, n, err, float(ref.abs().max())))
    if n >= 100000:
        for flag in ('0', '1'):
            os.environ['LFGC_FORWARD_TC'] = flag
            for _ in range(3): ops.sample_forward(geom, coords, grid, mlp)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20): ops.sample_forward(geom, coords, grid, mlp)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 20
            print('   TC=%s: %.1f us  %.2f G samples/s' % (flag, ms * 1e3, n / ms / 1e6))
print('--- occupancy sweep (TC kernel, n=262144, C16)')
m = setup_model(3, 32, 1, 4, 'fourier', 2, '', 0.1, 0.9, 'db2', 16, 15, '').cuda()
geom = m.geometry()
grid = ops.decode_fwd(geom, [f.detach().contiguous() for f in m.feature_grid], [None] * 3)
mlp = m.mlp_flat(); n = 262144
coords = torch.rand(n, 3, device='cuda') * 2 - 1
os.environ['LFGC_FORWARD_TC'] = '1'
for occ in ('3', '4', '5'):
    os.environ['LFGC_TC_GROUPS'] = occ
    for _ in range(3): ops.sample_forward(geom, coords, grid, mlp)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): ops.sample_forward(geom, coords, grid, mlp)
    e1.record(); torch.cuda.synchronize()
    print('  groups %s: %.1f us' % (occ, e0.elapsed_time(e1) / 20 * 1e3))
