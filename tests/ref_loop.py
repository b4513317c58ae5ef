"""Test infrastructure: the call sequence of the reference's training driver (training/training.py:71-243,
solve_model + training), written against the reference's module API and executed with THIS repository's drop-in
modules.  It consumes the torch CPU generator exactly like the reference run that produced
tests/golden/psnr_run_*.npz (setup_model draws, DataLoader(shuffle=True, num_workers=0) permutation, per-item
randint), so both runs see the same initial parameters and the same sample stream.
TensorBoard logging, checkpoint writing and the binary store are left out (they do not touch the numerics)."""
from copy import deepcopy

import torch
from torch.utils.data import DataLoader

from latent_feature_grid_compression_b200.data.IndexDataset import IndexDataset, normalize_volume
from latent_feature_grid_compression_b200.data.Interpolation import trilinear_f_interpolation
from latent_feature_grid_compression_b200.model.model_utils import setup_model
from latent_feature_grid_compression_b200.model.Smallify_Dropout import SmallifyLoss
from latent_feature_grid_compression_b200.visualization.OutputToVTK import tiled_net_out


class NeurcompDecay:
    def __init__(self, optimizer, pass_decay, lr_decay):
        self.optimizer, self.pass_decay, self.lr_decay = optimizer, pass_decay, lr_decay

    def step(self, prior, cur):
        if prior != int(cur) and (int(cur) + 1) % self.pass_decay == 0:
            for g in self.optimizer.param_groups:
                g['lr'] *= self.lr_decay


def solve(model, optimizer, decay, drop_loss, volume, dataset, loader, args, device):
    voxel_seen, passes = 0.0, 0.0
    crit = torch.nn.MSELoss().to(device)
    steps = 0
    while int(passes) + 1 < args['max_pass']:
        for raw, norm in loader:
            raw = raw.to(device).view(-1, 3)
            norm = norm.to(device).view(-1, 3)
            norm.requires_grad = True
            optimizer.zero_grad()
            pred = model(norm).squeeze(-1)
            gt = trilinear_f_interpolation(raw, volume, dataset.min_idx.to(device), dataset.max_idx.to(device),
                                           dataset.vol_res.to(device))
            prior = int(voxel_seen / dataset.n_voxels)
            voxel_seen += gt.shape[0]
            passes = voxel_seen / dataset.n_voxels
            loss = crit(pred, gt)
            if drop_loss is not None:
                loss = loss + drop_loss(model)
            loss.backward()
            optimizer.step()
            steps += 1
            if decay is not None:
                decay.step(prior, passes)
            if int(passes) >= args['max_pass']:
                break
    return steps


def training(args, volume_raw, device=torch.device('cuda')):
    vol = torch.from_numpy(volume_raw)
    vol = normalize_volume(vol, torch.min(vol), torch.max(vol), -1.0, 1.0)       # get_tensor_from_numpy
    dataset = IndexDataset(vol, args['sample_size'])
    loader = DataLoader(dataset, batch_size=args['batch_size'], shuffle=True, num_workers=0)
    volume = vol.to(device)
    model = setup_model(args['d_in'], args['n_hidden_size'], args['d_out'], args['n_layers'], args['embedding_type'],
                        args['n_embedding_freq'], args['drop_type'], args['drop_momentum'], args['drop_threshold'],
                        args['wavelet_filter'], args['grid_features'], args['grid_size'], args['checkpoint_path'])
    model.to(device)
    model.train()
    optimizer = torch.optim.Adam(model.parameters(), lr=args['lr'])
    decay = NeurcompDecay(optimizer, args['pass_decay'], args['lr_decay'])
    drop_loss = SmallifyLoss(args['lambda_drop_loss'], args['lambda_weight_loss']) if args['drop_type'] else None
    first = deepcopy(args)
    first['max_pass'] *= (2.0 / 3.0)
    steps = solve(model, optimizer, decay, drop_loss, volume, dataset, loader, first, device)
    zeros = model.save_dropvalues_on_grid(device)
    second = deepcopy(args)
    second['max_pass'] *= (1.0 / 3.0)
    optimizer = torch.optim.Adam(model.parameters(), lr=args['lr'] / 10.0)
    steps += solve(model, optimizer, None, None, volume, dataset, loader, second, device)   # decay bound to optimizer 1
    model.remove_drop_layers(device)
    psnr, l1, mse, rmse = tiled_net_out(dataset, model, True, gt_vol=volume, evaluate=True, write_vols=False)
    n_params = sum(p.numel() for n, p in model.named_parameters() if 'drop' not in n)
    return dict(psnr=psnr, mse=mse, rmse=rmse, num_zeros=float(zeros), num_parameters=n_params, steps=steps,
                compression_ratio=dataset.n_voxels / (n_params - float(zeros)))
