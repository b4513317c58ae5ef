"""Run the UNMODIFIED reference training driver (training/training.py: training()) on a small synthetic volume and
record the final PSNR, so that the GPU path can be held to the same run (same torch seed, same sample stream).

    python tests/golden/make_psnr_golden.py          # build container only (CPU, ~2 minutes)

Patches applied at run time only (no source edits): the torch>=2 eval `view` fix (oracle/ref_harness.py) and
`is_cuda=False` for the reconstruction call, which the reference hard-codes to True (training/training.py:20).
"""
import os
import shutil
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import ref_harness  # noqa: E402

ref_harness.install()
import model.Feature_Grid_Model as FGM  # noqa: E402
import training.training as T  # noqa: E402
import visualization.OutputToVTK as V  # noqa: E402

torch.set_num_threads(4)


def synthetic_volume(n=48):
    ax = np.linspace(0, 1, n, dtype=np.float64)
    x, y, z = np.meshgrid(ax, ax, ax, indexing='ij')
    v = np.sin(7 * x) * np.cos(5 * y) + 0.5 * np.sin(11 * z * x) + 0.3 * np.cos(9 * (y + z))
    return v.astype(np.float32)


def run(drop_type, tag, lam=1e-8):
    work = tempfile.mkdtemp()
    cwd = os.getcwd()
    os.chdir(work)
    os.makedirs('datasets')
    raw = synthetic_volume()
    np.save('datasets/vol.npy', raw)
    args = dict(expname='psnr_' + tag, data='datasets/vol.npy', basedir='/experiments/', Tensorboard_log_dir='',
                batch_size=256, sample_size=16, num_workers=0, max_pass=12, lr=0.008, pass_decay=20, lr_decay=0.2,
                smallify_decay=0, lambda_drop_loss=lam, lambda_weight_loss=lam, weight_dkl_multiplier=5e-4,
                variational_sigma=-7.0, d_in=3, d_out=1, n_hidden_size=32, n_layers=4, checkpoint_path='',
                binary_checkpoint_path='', embedding_type='fourier', n_embedding_freq=2, drop_type=drop_type,
                drop_momentum=0.025, drop_threshold=0.75, pruning_threshold_list=None, wavelet_filter='db2',
                grid_features=8, grid_size=9)
    orig_tiled = V.tiled_net_out
    orig_fwd = FGM.Feature_Grid_Model.forward

    def fwd(self, t):
        return ref_harness.patched_eval_forward(self, t) if not self.training else orig_fwd(self, t)
    FGM.Feature_Grid_Model.forward = fwd
    T.tiled_net_out = lambda ds, m, is_cuda, **kw: orig_tiled(ds, m, False, **{**kw, 'write_vols': False})
    T.store_model_parameters = lambda *a, **k: None          # storage format is exercised elsewhere
    try:
        torch.manual_seed(0)
        info = T.training(dict(args), verbose=False)
    finally:
        FGM.Feature_Grid_Model.forward = orig_fwd
        T.tiled_net_out = orig_tiled
        os.chdir(cwd)
        shutil.rmtree(work, ignore_errors=True)
    print(tag, info['psnr'], info['num_zeros'], info['compression_ratio'])
    np.savez_compressed(os.path.join(HERE, 'psnr_run_%s.npz' % tag), volume_raw=raw,
                        psnr=np.float64(info['psnr']), mse=np.float64(info['mse']), rmse=np.float64(info['rmse']),
                        num_zeros=np.float64(info['num_zeros']), num_parameters=np.int64(info['num_parameters']),
                        compression_ratio=np.float64(info['compression_ratio']),
                        args_keys=np.asarray(list(args.keys())), args_vals=np.asarray([repr(v) for v in args.values()]))


if __name__ == '__main__':
    run('', 'basic')
    run('smallify', 'smallify')
