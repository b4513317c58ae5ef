"""lfgc_grid_step alone against the launches it replaces (reduce_partials + decode_bwd + adam + decode_fwd), CUDA events,
hot L2, per configuration.  python profiles/grid_step_timing.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from latent_feature_grid_compression_b200 import ops
from latent_feature_grid_compression_b200.model.model_utils import setup_model
from latent_feature_grid_compression_b200.training.fast_loop import FastTrainer

vol = (torch.rand(64, 64, 64, device='cuda') * 2 - 1)
for C, G in ((16, 15), (8, 9), (32, 15), (16, 9), (8, 15), (16, 17), (16, 15)):
    res = {}
    for flag in ('1', 's', '0'):
        os.environ['LFGC_GRID_STEP'] = '0' if flag == '0' else '1'; os.environ['LFGC_GRID_STEP_SPLIT'] = '0' if flag == 's' else '1'
        torch.manual_seed(0)
        m = setup_model(3, 32, 1, 4, 'fourier', 2, '', 0.1, 0.9, 'db2', C, G, '').cuda().train()
        tr = FastTrainer(m, vol, 32768, lr=0.008, seed=1)
        tr.capture()
        for _ in range(50):
            tr.step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(500):
            tr.step()
        e1.record()
        torch.cuda.synchronize()
        step_us = e0.elapsed_time(e1) * 1e3 / 500
        # the per-sample kernel alone (with its partial reduction when grid_step is off)
        geom = tr.geom
        gm = torch.empty(geom.mlp_param_count, device='cuda')
        e0.record()
        for _ in range(200):
            if tr._gstep:
                ops.train_step_partials(geom, vol, 32768, 1, 0, 1.0 / 32768, tr.grid_cl, tr.mlp_flat, tr.grad_grid, tr.workspace)
            else:
                ops.train_step(geom, vol, 32768, 1, 0, 1.0 / 32768, tr.grid_cl, tr.mlp_flat, tr.grad_grid, gm, tr.loss_sum, tr.workspace)
        e1.record()
        torch.cuda.synchronize()
        k_us = e0.elapsed_time(e1) * 1e3 / 200
        res[flag] = (step_us, k_us, tr.launches_per_step)
        tr._graphs.clear()
    print('C%d G%d: grid_step split  step %.2f us (per-sample kernel %.2f, rest %.2f) | whole pyramid per CTA  step %.2f us '
          '(rest %.2f) | separate  step %.2f us (kernel+reduce %.2f, rest %.2f, %d launches)' % (
              C, G, res['1'][0], res['1'][1], res['1'][0] - res['1'][1], res['s'][0], res['s'][0] - res['s'][1],
              res['0'][0], res['0'][1], res['0'][0] - res['0'][1], res['0'][2]))
