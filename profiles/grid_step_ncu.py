"""A few eager optimiser steps of the grid-step path (C16/G15, 32768 samples) for ncu:
ncu --set full --import-source on -k regex:grid_step -c 2 -o gpurun_out/gs python profiles/grid_step_ncu.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from latent_feature_grid_compression_b200.model.model_utils import setup_model
from latent_feature_grid_compression_b200.training.fast_loop import FastTrainer
vol = (torch.rand(64, 64, 64, device='cuda') * 2 - 1)
torch.manual_seed(0)
m = setup_model(3, 32, 1, 4, 'fourier', 2, '', 0.1, 0.9, 'db2', 16, 15, '').cuda().train()
tr = FastTrainer(m, vol, 32768, lr=0.008, seed=1, use_graph=False)
for _ in range(6):
    tr.step()
torch.cuda.synchronize()
print('ok', tr.last_loss())
