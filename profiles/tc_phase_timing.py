import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from latent_feature_grid_compression_b200 import _lib, ops
from latent_feature_grid_compression_b200.model.model_utils import setup_model
lib = _lib.load(); lib.lfgc_btc_timing.argtypes = [ctypes.c_void_p, ctypes.c_int]
torch.manual_seed(0)
m = setup_model(3, 32, 1, 4, 'fourier', 2, '', 0.1, 0.9, 'db2', 16, 15, '').cuda()
geom = m.geometry()
grid = ops.decode_fwd(geom, [f.detach().contiguous() for f in m.feature_grid], [None] * 3)
mlp = m.mlp_flat()
vol = torch.rand(255, 255, 255, device='cuda') * 2 - 1
ws = torch.empty(geom.backward_workspace_bytes // 4, device='cuda')
os.environ['LFGC_BACKWARD_TC'] = '1'
names = ['setup', 'input', 'barrier', 'mma issue', 'mma wait', 'fwd epilogue', 'output+loss', 'bwd staging', 'dz', 'scatter', 'flush (reduction)', 'dW wait', 'setup: fill+stage', 'setup: panels', 'flush: lo halves', 'flush: dW rows']
panels = ops.tc_panel_image(geom, mlp)
for tps in ('2',):
    os.environ['LFGC_TC_TPS'] = tps
    for n in (32768, 32768.5, 262144):
        img = panels if n != int(n) else None     # n = 32768.5: the same launch with the ready-made operand image
        n = int(n)
        gg = torch.zeros((*geom.G, geom.Cp), device='cuda'); gm = torch.empty(geom.mlp_param_count, device='cuda'); ls = torch.zeros(1, device='cuda')
        for _ in range(3): ops.train_step_partials(geom, vol, n, 7, 0, 1.0 / n, grid, mlp, gg, ws, tc_panels=img)
        lib.lfgc_btc_timing(None, 1)
        reps = 10
        for _ in range(reps): ops.train_step_partials(geom, vol, n, 7, 0, 1.0 / n, grid, mlp, gg, ws, tc_panels=img)
        buf = (ctypes.c_ulonglong * 16)(); lib.lfgc_btc_timing(buf, 1)
        ctas = min(148, (n + 127) // 128)
        tot = sum(buf[:16])
        print('tps=%s n=%d%s: cycles per CTA (thread 0) per launch: %.0f  (tiles per CTA %.2f)' % (tps, n, ' (operand image)' if img is not None else '', tot / reps / ctas, n / 128 / ctas))
        for i, nm in enumerate(names): print('  %-14s %9.0f %5.1f%%' % (nm, buf[i] / reps / ctas, 100.0 * buf[i] / tot))
