"""Short eager (no CUDA graph) run of the hot path for ncu: a few training steps at the bench config plus one
reconstruction slab.  Usage (see B200_PROFILING.md):

    python profiles/profile_step.py                       # must exit 0 first
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
        python profiles/profile_step.py
    ncu --set full --clock-control none --import-source on -k regex:sample_backward -s 3 -c 2 -o gpurun_out/prof_bwd \
        python profiles/profile_step.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch

import bench
from latent_feature_grid_compression_b200.data.IndexDataset import IndexDataset
from latent_feature_grid_compression_b200.model.model_utils import setup_model
from latent_feature_grid_compression_b200.training.fast_loop import FastTrainer
from latent_feature_grid_compression_b200.visualization.OutputToVTK import field_from_net

R = int(os.environ.get('LFGC_PROFILE_R', '255'))
steps = int(os.environ.get('LFGC_PROFILE_STEPS', '6'))
dev = torch.device('cuda', 0)
cfg = bench.CFG
volume = bench.synthetic_volume(R, dev)
torch.manual_seed(0)
model = setup_model(3, cfg['H'], 1, cfg['L'], 'fourier', cfg['F'], '', 0.1, 0.9, cfg['wavelet'], cfg['C'], cfg['G'], '')
model.to(dev).train()
trainer = FastTrainer(model, volume, cfg['batch'], lr=cfg['lr'], seed=1234, use_graph=False)
for _ in range(steps):
    trainer.step()
torch.cuda.synchronize()
print('loss', trainer.last_loss())
ds = IndexDataset(volume.cpu(), 16)
model.eval()
out = field_from_net(ds, model, True, slab=(0, 64), to_cpu=False)
torch.cuda.synchronize()
print('slab', tuple(out.shape), float(out.mean()))
