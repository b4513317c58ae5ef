"""Generate the golden fixtures in this directory by RUNNING THE REFERENCE ITSELF.

Run in the build container only (needs /root/reference, CPU torch):

    python tests/golden/make_golden.py

Every ``*.npz`` written here is the output of the unmodified reference modules
(imported through ``oracle/ref_harness.py``) on seeded inputs; the numpy oracle
(``oracle/fvsrn_numpy.py``) and the CUDA path are both tested against them.
The reference holds no golden vectors or asserting tests of its own
(SURVEY.md section 4), so these fixtures are what pins parity.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_harness  # noqa: E402

ref_harness.install()

from model.model_utils import setup_model  # noqa: E402
from model.Smallify_Dropout import SmallifyLoss, SmallifyDropout  # noqa: E402
from model.Variational_Dropout_Layer import (  # noqa: E402
    VariationalDropoutLoss, VariationalDropout, Variance_Model)
from model.Straight_Through_Dropout import (  # noqa: E402
    Straight_Through_Dropout, MaskedWavelet_Straight_Through_Dropout)
from data.IndexDataset import IndexDataset, get_tensor  # noqa: E402
from data.Interpolation import trilinear_f_interpolation  # noqa: E402
from visualization.OutputToVTK import field_from_net, calculate_deviation_statistics  # noqa: E402
import model.Feature_Grid_Model as FGM  # noqa: E402

torch.set_num_threads(1)
torch.use_deterministic_algorithms(False)


def npy(t):
    if isinstance(t, torch.Tensor):
        return t.detach().cpu().numpy().copy()
    return np.asarray(t)


class NoiseRecorder:
    """Record every torch.randn_like / torch.rand draw made by the mask layers."""

    def __init__(self):
        self.draws = []

    def __enter__(self):
        self._randn_like = torch.randn_like
        self._rand = torch.rand

        def randn_like(x, *a, **k):
            r = self._randn_like(x, *a, **k)
            self.draws.append(npy(r))
            return r

        def rand(*a, **k):
            r = self._rand(*a, **k)
            self.draws.append(npy(r))
            return r

        torch.randn_like = randn_like
        torch.rand = rand
        return self

    def __exit__(self, *exc):
        torch.randn_like = self._randn_like
        torch.rand = self._rand


class NoisePlayer:
    """Replay recorded draws so that a second reference call sees the same noise."""

    def __init__(self, draws):
        self.draws = [torch.from_numpy(d.copy()) for d in draws]
        self.i = 0

    def __enter__(self):
        self._randn_like = torch.randn_like
        self._rand = torch.rand

        def nxt(*a, **k):
            r = self.draws[self.i]
            self.i += 1
            return r

        torch.randn_like = nxt
        torch.rand = nxt
        return self

    def __exit__(self, *exc):
        torch.randn_like = self._randn_like
        torch.rand = self._rand


def make_coords(n, gen):
    c = torch.rand(n, 3, generator=gen) * 2.0 - 1.0
    # edge cases: exact borders, centre, slightly outside the [-1,1] cube (zeros padding)
    special = torch.tensor([
        [-1.0, -1.0, -1.0], [1.0, 1.0, 1.0], [0.0, 0.0, 0.0], [1.0, -1.0, 0.5],
        [-1.1, 0.3, 0.2], [0.2, 1.15, -0.4], [0.9, 0.9, -1.2], [1.3, 1.3, 1.3],
        [0.999999, -0.999999, 0.5], [-0.5, 0.25, 1.0],
    ])
    c[: special.shape[0]] = special
    return c


def model_case(tag, drop_type, wavelet, C, G, H=32, L=4, F=2, N=96, p=0.1, thr=0.9, seed=0,
               perturb_masks=True):
    torch.manual_seed(seed)
    model = setup_model(3, H, 1, L, 'fourier', F, drop_type, p, thr, wavelet, C, G, '')
    model.train()
    gen = torch.Generator().manual_seed(seed + 1000)
    # make the mask parameters non-trivial so that pruning decisions are exercised
    if perturb_masks:
        with torch.no_grad():
            for d in model.drop:
                if isinstance(d, VariationalDropout):
                    d.log_thetas.add_(0.3 * torch.randn(d.log_thetas.shape, generator=gen))
                    d.log_var.add_(2.0 * torch.randn(d.log_var.shape, generator=gen))
                if isinstance(d, (Straight_Through_Dropout, MaskedWavelet_Straight_Through_Dropout)):
                    d.mask_values.copy_(torch.randn(d.mask_values.shape, generator=gen) * 1.5 + 0.3)
    coords = make_coords(N, gen).requires_grad_(True)
    wout = torch.randn(N, 1, generator=gen)

    out = {}
    for name, prm in model.state_dict().items():
        out['sd.' + name] = npy(prm)
    out['shape_array'] = np.asarray(model.shape_array, dtype=np.int64).reshape(-1, 3)
    out['coords'] = npy(coords)
    out['wout'] = npy(wout)
    out['meta'] = np.asarray([C, G, H, L, F, N], dtype=np.int64)

    with NoiseRecorder() as rec:
        y = model(coords)
    out['n_noise'] = np.asarray([len(rec.draws)])
    for i, d in enumerate(rec.draws):
        out['noise.%d' % i] = d
    with NoisePlayer(rec.draws):
        grid = model.decode_volume()
    out['grid'] = npy(grid)
    out['y_train'] = npy(y)
    loss = (y * wout).sum()
    loss.backward()
    for name, prm in model.named_parameters():
        out['grad.' + name] = npy(prm.grad) if prm.grad is not None else np.zeros(0, np.float32)
    out['grad.coords'] = npy(coords.grad)

    # regulariser losses (values + grads)
    model.zero_grad()
    if drop_type in ('smallify', 'straight_through', 'masked_straight_through'):
        sl = SmallifyLoss(weight_l1=0.37, weight_l2=1.9)
        v = sl(model)
        v.backward()
        out['smallify_loss'] = npy(v)
        for name, prm in model.named_parameters():
            if prm.grad is not None:
                out['sgrad.' + name] = npy(prm.grad)
        model.zero_grad()
    if 'variational' in drop_type:
        vl = VariationalDropoutLoss(size_volume=1000.0, batch_size=float(N), weight_dkl=1.3, weight_weights=0.7)
        with NoisePlayer(rec.draws):
            y2 = model(coords.detach()).squeeze(-1)
        gt = torch.rand(N, generator=gen) * 2 - 1
        logsig = torch.randn(N, generator=gen) * 0.2 - 3.0
        logsig.requires_grad_(True)
        tot, ll, mse, dkl, wsum = vl(model, y2, gt, logsig, 5e-5)
        tot.backward()
        out['vloss.gt'] = npy(gt)
        out['vloss.logsig'] = npy(logsig)
        out['vloss.values'] = np.asarray([tot.item(), ll.item(), mse.item(), dkl.item(), wsum.item(),
                                          vl.weight_dkl], dtype=np.float64)
        out['vloss.grad_logsig'] = npy(logsig.grad)
        for name, prm in model.named_parameters():
            if prm.grad is not None:
                out['vgrad.' + name] = npy(prm.grad)
        for i, d in enumerate(model.drop):
            out['dkl.%d' % i] = npy(d.calculate_Dkl())
            out['droprate.%d' % i] = npy(d.dropout_rates)
        model.zero_grad()

    # eval-mode forward on a small non-cubic tile (OutputToVTK.py:40-41 shape contract)
    model.eval()
    tile = (torch.rand(1, 1, 4, 5, 6, 3, generator=gen) * 2 - 1)
    with torch.no_grad():
        with NoiseRecorder() as rec2:
            ye = ref_harness.patched_eval_forward(model, tile)
    out['tile'] = npy(tile)
    out['y_eval'] = npy(ye)
    out['n_noise_eval'] = np.asarray([len(rec2.draws)])
    for i, d in enumerate(rec2.draws):
        out['noise_eval.%d' % i] = d
    model.train()

    # mask baking: save_dropvalues_on_grid + remove_drop_layers (Feature_Grid_Model.py:110-140)
    if drop_type and drop_type != 'straight_through':  # the Bernoulli layer has no size_layer (reference bug)
        zeros = model.save_dropvalues_on_grid(torch.device('cpu'))
        out['bake.zeros'] = npy(zeros)
        for i, f in enumerate(model.feature_grid):
            out['bake.feature_grid.%d' % i] = npy(f)
        for i, d in enumerate(model.drop):
            out['bake.d_mask.%d' % i] = npy(d.d_mask.float())
        with torch.no_grad():
            with NoiseRecorder():
                yb = model(coords.detach())
        out['y_baked_train'] = npy(yb)
        model.remove_drop_layers(torch.device('cpu'))
        for i, f in enumerate(model.feature_grid):
            out['final.feature_grid.%d' % i] = npy(f)
        with torch.no_grad():
            out['y_final_train'] = npy(model(coords.detach()))

    np.savez_compressed(os.path.join(HERE, 'model_%s.npz' % tag), **out)
    print('wrote model_%s.npz' % tag, 'levels', len(model.shape_array))


def smallify_tracker_case():
    torch.manual_seed(3)
    d = SmallifyDropout((3, 4, 5), 0.1, 0.05)
    d.train()
    x = torch.randn(2, 3, 4, 5)
    out = {'betas0': npy(d.betas), 'x': npy(x)}
    gen = torch.Generator().manual_seed(11)
    for step in range(12):
        y = d(x)
        with torch.no_grad():
            flip = torch.rand(d.betas.shape, generator=gen) < 0.25
            d.betas[flip] *= -1.0
        out['flip.%d' % step] = npy(flip)
    out['y_last'] = npy(y)
    out['EMA'] = npy(d.tracker.EMA)
    out['EMAVar'] = npy(d.tracker.EMAVar)
    out['mask'] = npy(d.calculate_pruning_mask(torch.device('cpu')))
    out['l1'] = npy(d.l1_loss())
    out['baked'] = npy(d.multiply_values_with_dropout(x, torch.device('cpu')))
    out['momentum_threshold'] = np.asarray([0.1, 0.05])
    np.savez_compressed(os.path.join(HERE, 'smallify_tracker.npz'), **out)
    print('wrote smallify_tracker.npz')


def dataset_case():
    gen = torch.Generator().manual_seed(5)
    out = {}
    for tag, shape in (('a', (7, 9, 6)), ('b', (12, 12, 12))):
        vol_raw = torch.randn(*shape, generator=gen) * 3.0 + 0.5
        path = '/tmp/_lfgc_golden_%s.npy' % tag
        np.save(path, vol_raw.numpy())
        vol = get_tensor(path)
        os.remove(path)
        ds = IndexDataset(vol, 16)
        out[tag + '.vol_raw'] = npy(vol_raw)
        out[tag + '.vol'] = npy(vol)
        out[tag + '.n_voxels'] = np.asarray([ds.n_voxels])
        out[tag + '.max_idx'] = npy(ds.max_idx)
        out[tag + '.scales'] = npy(ds.scales)
        out[tag + '.volume_indices'] = npy(ds.volume_indices)
        # __getitem__ with recorded indices (IndexDataset.py:90-96)
        rec = []
        _randint = torch.randint

        def randint(*a, **k):
            r = _randint(*a, **k)
            rec.append(npy(r))
            return r
        torch.randint = randint
        torch.manual_seed(17)
        raw, norm = ds[0]
        torch.randint = _randint
        out[tag + '.idx'] = rec[0]
        out[tag + '.raw'] = npy(raw)
        out[tag + '.norm'] = npy(norm)
        # ground-truth lookup at the integer positions and at arbitrary float positions
        gt_int = trilinear_f_interpolation(raw, vol, ds.min_idx, ds.max_idx, ds.vol_res)
        out[tag + '.gt_int'] = npy(gt_int)
        pf = torch.rand(200, 3, generator=gen) * ds.max_idx.unsqueeze(0)
        pf[:8] = torch.tensor([[0, 0, 0], [1, 2, 3], [0.5, 0.5, 0.5], [6, 8, 5], [5.999, 7.5, 0.001],
                               [3.25, 0, 4.75], [2, 2.5, 2], [6, 0, 5]], dtype=torch.float)[:8].clamp(
            max=ds.max_idx.unsqueeze(0))
        out[tag + '.pf'] = npy(pf)
        out[tag + '.gt_float'] = npy(trilinear_f_interpolation(pf, vol, ds.min_idx, ds.max_idx, ds.vol_res))
        # general bounding box form (min_bb != 0)
        min_bb = torch.tensor([-1.0, -2.0, 0.5])
        max_bb = torch.tensor([1.0, 3.0, 4.5])
        pb = min_bb + torch.rand(100, 3, generator=gen) * (max_bb - min_bb)
        out[tag + '.pb'] = npy(pb)
        out[tag + '.bb'] = np.stack([npy(min_bb), npy(max_bb)])
        out[tag + '.gt_bb'] = npy(trilinear_f_interpolation(pb, vol, min_bb, max_bb, ds.vol_res))
    np.savez_compressed(os.path.join(HERE, 'dataset.npz'), **out)
    print('wrote dataset.npz')


def reconstruct_case():
    torch.manual_seed(21)
    gen = torch.Generator().manual_seed(22)
    shape = (40, 37, 35)
    vol = torch.rand(*shape, generator=gen) * 2 - 1
    ds = IndexDataset(vol, 16)
    model = setup_model(3, 32, 1, 4, 'fourier', 2, '', 0.1, 0.9, 'db2', 4, 15, '')
    # reference eval forward is broken under torch>=2: route the one call through the patched forward
    model.eval()
    orig_forward = FGM.Feature_Grid_Model.forward
    seen_tiles = []

    def recording_forward(self, t):
        seen_tiles.append(t.detach().clone())
        return ref_harness.patched_eval_forward(self, t)
    FGM.Feature_Grid_Model.forward = recording_forward
    try:
        with torch.no_grad():
            full = field_from_net(ds, model, False, tiled_res=32)
    finally:
        FGM.Feature_Grid_Model.forward = orig_forward
    # the coordinates the reference fed to the network, re-assembled in its tile order (x, then y, then z)
    coords = torch.zeros(*shape, 3)
    it = iter(seen_tiles)
    for xb in range(0, shape[0], 32):
        for yb in range(0, shape[1], 32):
            for zb in range(0, shape[2], 32):
                t = next(it)[0]
                coords[xb:xb + t.shape[0], yb:yb + t.shape[1], zb:zb + t.shape[2]] = t
    psnr, l1, mse, rmse = calculate_deviation_statistics(full, vol)
    out = {'vol': npy(vol), 'full': npy(full), 'stats': np.asarray([psnr, l1, mse, rmse], dtype=np.float64),
           'scales': npy(ds.scales), 'axis0': npy(coords[:, 0, 0, 0]), 'axis1': npy(coords[0, :, 0, 1]),
           'axis2': npy(coords[0, 0, :, 2])}
    assert torch.equal(coords[..., 0], coords[:, :1, :1, 0].expand(*shape))
    for name, prm in model.state_dict().items():
        out['sd.' + name] = npy(prm)
    np.savez_compressed(os.path.join(HERE, 'reconstruct.npz'), **out)
    print('wrote reconstruct.npz')


def variance_model_case():
    torch.manual_seed(31)
    gen = torch.Generator().manual_seed(32)
    vm = Variance_Model()
    x = torch.rand(50, 3, generator=gen) * 2 - 1
    w = torch.randn(50, 1, generator=gen)
    y = vm(x)
    (y * w).sum().backward()
    out = {'x': npy(x), 'w': npy(w), 'y': npy(y)}
    for name, prm in vm.named_parameters():
        out['sd.' + name] = npy(prm)
        out['grad.' + name] = npy(prm.grad)
    np.savez_compressed(os.path.join(HERE, 'variance_model.npz'), **out)
    print('wrote variance_model.npz')


def trajectory_case(tag, drop_type, steps=12):
    """A few optimiser steps of the reference step (training/training.py:89-138) on fixed batches."""
    torch.manual_seed(41)
    gen = torch.Generator().manual_seed(42)
    shape = (20, 18, 22)
    xs = torch.linspace(0, 1, shape[0]).view(-1, 1, 1)
    ys = torch.linspace(0, 1, shape[1]).view(1, -1, 1)
    zs = torch.linspace(0, 1, shape[2]).view(1, 1, -1)
    vol = torch.sin(5 * xs) * torch.cos(3 * ys) + 0.5 * torch.sin(7 * zs * xs)
    vol = 2 * (vol - vol.min()) / (vol.max() - vol.min()) - 1
    ds = IndexDataset(vol, 16)
    model = setup_model(3, 32, 1, 4, 'fourier', 2, drop_type, 0.1, 0.9, 'db2', 8, 15, '')
    model.train()
    opt = torch.optim.Adam(model.parameters(), lr=0.008)
    crit = torch.nn.MSELoss()
    dl = SmallifyLoss(1e-4, 1e-5) if drop_type == 'smallify' else None
    N = 256
    out = {'vol': npy(vol)}
    for name, prm in model.state_dict().items():
        out['sd0.' + name] = npy(prm)
    losses = []
    idx_all = torch.randint(0, ds.n_voxels, (steps, N), generator=gen)
    out['idx'] = npy(idx_all)
    for s in range(steps):
        raw = ds.volume_indices[idx_all[s]]
        norm = ds.scales.unsqueeze(0) * (2.0 * (raw - ds.min_idx) / (ds.max_idx - ds.min_idx) - 1.0)
        opt.zero_grad()
        pred = model(norm).squeeze(-1)
        gt = trilinear_f_interpolation(raw, vol, ds.min_idx, ds.max_idx, ds.vol_res)
        loss = crit(pred, gt)
        if dl is not None:
            loss = loss + dl(model)
        loss.backward()
        opt.step()
        losses.append(loss.item())
    out['losses'] = np.asarray(losses, dtype=np.float64)
    for name, prm in model.state_dict().items():
        out['sd1.' + name] = npy(prm)
    if drop_type == 'smallify':
        for i, d in enumerate(model.drop):
            out['EMA.%d' % i] = npy(d.tracker.EMA)
            out['EMAVar.%d' % i] = npy(d.tracker.EMAVar)
    np.savez_compressed(os.path.join(HERE, 'trajectory_%s.npz' % tag), **out)
    print('wrote trajectory_%s.npz' % tag, losses[0], losses[-1])


if __name__ == '__main__':
    if len(sys.argv) > 1:  # regenerate selected fixtures only: python make_golden.py reconstruct_case ...
        for fn in sys.argv[1:]:
            globals()[fn]()
        sys.exit(0)
    model_case('basic_db2_c16_g15', '', 'db2', 16, 15)
    model_case('basic_haar_c4_g16', '', 'haar', 4, 16, N=64)
    model_case('basic_db2_c8_g17_h64_l3_f3', '', 'db2', 8, 17, H=64, L=3, F=3, N=64)
    model_case('basic_db2_c4_g5_nolevels', '', 'db2', 4, 5, N=48)
    model_case('smallify_db2_c6_g15', 'smallify', 'db2', 6, 15, p=0.025, thr=0.75)
    model_case('variational_db2_c8_g15', 'variational_dynamic', 'db2', 8, 15, p=0.1, thr=0.5)
    model_case('maskedste_db2_c8_g15', 'masked_straight_through', 'db2', 8, 15, p=0.5, thr=0.6, N=64)
    model_case('bernoulli_db2_c8_g15', 'straight_through', 'db2', 8, 15, p=0.5, thr=0.5, N=64)
    smallify_tracker_case()
    dataset_case()
    reconstruct_case()
    variance_model_case()
    trajectory_case('basic', '')
    trajectory_case('smallify', 'smallify')
