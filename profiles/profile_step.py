"""Short eager (no CUDA graph) run of the hot path for ncu: a few training steps of one BASELINE configuration
(LFGC_PROFILE_CONFIG, default mhd_p_basic) plus one reconstruction slab.  Usage (see B200_PROFILING.md):

    python profiles/profile_step.py                       # must exit 0 first
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
        python profiles/profile_step.py
    ncu --set full --clock-control none --import-source on -k regex:backward_tc -s 3 -c 2 -o gpurun_out/prof_bwd \
        python profiles/profile_step.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch

import bench
from latent_feature_grid_compression_b200.data.IndexDataset import IndexDataset
from latent_feature_grid_compression_b200.training.fast_loop import make_trainer
from latent_feature_grid_compression_b200.visualization.OutputToVTK import field_from_net

name = os.environ.get('LFGC_PROFILE_CONFIG', 'mhd_p_basic')
steps = int(os.environ.get('LFGC_PROFILE_STEPS', '6'))
cfg = bench.CONFIGS[name]
R = int(os.environ.get('LFGC_PROFILE_R', str(min(cfg['R'], 512))))
dev = torch.device('cuda', 0)
volume = bench.synthetic_volume(R, dev)
model = bench.build_model(name, dev)
trainer = make_trainer(model, volume, R ** 3, cfg['args'], cfg['args']['lr'], seed=1234)
trainer._use_graph = False
for _ in range(steps):
    trainer.step()
torch.cuda.synchronize()
print('loss', trainer.last_loss(), 'launches per step', trainer.launches_per_step)
ds = IndexDataset(torch.zeros(1, 1, 1).expand(R, R, R), 16)
model.eval()
out = field_from_net(ds, model, True, slab=(0, min(64, R)), to_cpu=False)
torch.cuda.synchronize()
print('slab', tuple(out.shape), float(out.mean()))
