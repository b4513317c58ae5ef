import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from latent_feature_grid_compression_b200 import ops
from latent_feature_grid_compression_b200.model.model_utils import setup_model


def rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


torch.manual_seed(0)
cases = [(16, 15, 32, 4, 1000), (16, 15, 32, 4, 32768), (8, 15, 20, 3, 5000), (32, 17, 32, 4, 4096), (6, 15, 32, 2, 300)]
if len(sys.argv) > 1 and sys.argv[1] == 'quick':
    cases = cases[:1]
for (C, G, H, Lyr, n) in cases:
    m = setup_model(3, H, 1, Lyr, 'fourier', 2, '', 0.1, 0.9, 'db2', C, G, '').cuda()
    with torch.no_grad():
        for lyr in m.net_layers:
            lyr.bias.uniform_(-0.5, 0.5)
    geom = m.geometry()
    grid = ops.decode_fwd(geom, [f.detach().contiguous() for f in m.feature_grid], [None] * len(m.feature_grid))
    mlp = m.mlp_flat()
    coords = (torch.rand(n, 3, device='cuda') * 2.2 - 1.1)
    gout = torch.randn(n, device='cuda') / n
    res = {}
    for flag in ('0', '1'):
        os.environ['LFGC_BACKWARD_TC'] = flag
        gg, gm = ops.sample_backward(geom, coords, gout, grid, mlp)
        torch.cuda.synchronize()
        res[flag] = (gg.clone(), gm.clone())
    e_grid = rel(res['1'][0], res['0'][0])
    e_mlp = rel(res['1'][1], res['0'][1])
    print('C%d G%d H%d L%d n%d: backward-only  grid rel %.3e  mlp rel %.3e' % (C, G, H, Lyr, n, e_grid, e_mlp))
    def breakdown(res):
        in0 = 3 + 12 + C
        offs = [0]
        for l in range(Lyr):
            K = in0 if l == 0 else H
            offs.append(offs[-1] + K * H)
            offs.append(offs[-1] + H)
        offs.append(offs[-1] + H + 1)
        names = []
        for l in range(Lyr):
            names += ['W%d' % l, 'b%d' % l]
        names += ['Wf+bf']
        for i, nm in enumerate(names):
            a, b = res['1'][1][offs[i]:offs[i + 1]], res['0'][1][offs[i]:offs[i + 1]]
            print('   %s: rel %.3e  (max ref %.3e, max tc %.3e)' % (nm, rel(a, b), float(b.abs().max()), float(a.abs().max())))
    if e_mlp > 1e-5:
        breakdown(res)
    # fused mode with explicit samples
    vol = torch.rand(40, 41, 42, device='cuda') * 2 - 1
    ws = torch.empty(geom.backward_workspace_bytes // 4, device='cuda')
    fres = {}
    for flag in ('0', '1'):
        os.environ['LFGC_BACKWARD_TC'] = flag
        gg = torch.zeros((*geom.G, geom.Cp), device='cuda')
        gm = torch.empty(geom.mlp_param_count, device='cuda')
        ls = torch.zeros(1, device='cuda')
        ops.train_step(geom, vol, n, 7, 0, 1.0 / n, grid, mlp, gg, gm, ls, ws)
        torch.cuda.synchronize()
        fres[flag] = (gg, gm, ls)
    print('   fused: grid rel %.3e  mlp rel %.3e  loss %.6e vs %.6e' % (rel(fres['1'][0], fres['0'][0]), rel(fres['1'][1], fres['0'][1]),
                                                                     float(fres['1'][2]), float(fres['0'][2])))
    if rel(fres['1'][1], fres['0'][1]) > 5e-6:
        breakdown(fres)
    if n >= 30000:
        for flag, tps in (('0', ''), ('1', '2'), ('1', '4')):
            os.environ['LFGC_BACKWARD_TC'] = flag
            if tps:
                os.environ['LFGC_TC_TPS'] = tps
            for nn in (n, 8 * n):
                gg = torch.zeros((*geom.G, geom.Cp), device='cuda')
                gm = torch.empty(geom.mlp_param_count, device='cuda')
                ls = torch.zeros(1, device='cuda')
                for _ in range(3):
                    ops.train_step(geom, vol, nn, 7, 0, 1.0 / nn, grid, mlp, gg, gm, ls, ws)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(20):
                    ops.train_step(geom, vol, nn, 7, 0, 1.0 / nn, grid, mlp, gg, gm, ls, ws)
                e1.record(); torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / 20
                print('   TC=%s tps=%s n=%d: %.1f us  %.3f G samples/s' % (flag, tps or '-', nn, ms * 1e3, nn / ms / 1e6))
