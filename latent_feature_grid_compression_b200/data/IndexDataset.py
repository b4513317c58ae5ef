"""Volume loading and the random voxel sampler -- counterpart of the reference's data/IndexDataset.py.

The reference materialises an (n_voxels, 3) fp32 index table (199 MB for 255^3, 12.9 GB for 1024^3) and samples it
on CPU DataLoader workers.  Here the table never exists: ``IndexDataset.sample`` draws voxel indices and
evaluates raw positions, normalised positions and the ground-truth value in one CUDA launch (``lfgc_sample``).
``__getitem__`` keeps the reference contract (``sample_size`` random positions -> (raw, norm)) for code that still
drives the dataset through a ``DataLoader``; it uses the torch CPU generator exactly like the reference so that
the same seed yields the same sample stream.
"""
from __future__ import annotations

import numpy as np
import torch
from torch.utils.data.dataset import Dataset

from .. import ops


def normalize_volume(volume, minV, maxV, minN, maxN):
    """Affine map of [minV, maxV] onto [minN, maxN] in the reference's operation order (IndexDataset.py:7-8)."""
    return (maxN - minN) * ((volume - minV) / (maxV - minV)) + minN


def _to_unit_range(volume: torch.Tensor):
    return normalize_volume(volume, torch.min(volume), torch.max(volume), -1.0, 1.0)


def get_tensor_from_numpy(filepath):
    volume = _to_unit_range(torch.from_numpy(np.load(filepath).astype(np.float32)))
    print('Loaded Numpy Volume Successfully. Shape of: ', volume.shape)
    return volume


def get_tensor_from_hdf5(filepath):
    import h5py  # optional dependency, as in the reference
    with h5py.File(filepath, 'r') as f:
        first_key = list(f.keys())[0]
        volume = _to_unit_range(torch.from_numpy(np.squeeze(f[first_key][()])))
    print('Loaded HDF5 Volume Successfully. Shape of: ', volume.shape)
    return volume


def get_tensor(filepath):
    if filepath.endswith('.npy'):
        return get_tensor_from_numpy(filepath)
    if filepath.endswith('.h5'):
        return get_tensor_from_hdf5(filepath)
    if filepath.endswith('.cvol'):
        raise NotImplementedError('.cvol volumes need the third-party pyrenderer module (not part of this path)')
    raise ValueError('unsupported volume file %r' % (filepath,))


class _LazyVoxelIndices:
    """Stands in for the reference's (n_voxels, 3) ``volume_indices`` table: rows are computed on demand."""

    def __init__(self, shape):
        self.shape_3d = tuple(int(s) for s in shape)
        self.n = int(np.prod(self.shape_3d))

    def __len__(self):
        return self.n

    @property
    def shape(self):
        return (self.n, 3)

    def __getitem__(self, idx):
        idx = torch.as_tensor(idx, dtype=torch.long)
        r1, r2 = self.shape_3d[1], self.shape_3d[2]
        i = torch.div(idx, r1 * r2, rounding_mode='floor')
        j = torch.div(idx, r2, rounding_mode='floor') % r1
        k = idx % r2
        return torch.stack([i, j, k], dim=-1).to(torch.float32)

    def view(self, *shape):
        return self


class IndexDataset(Dataset):
    def __init__(self, volume, sampleSize=16):
        self.vol_res = torch.tensor(volume.shape, dtype=torch.float)
        self.vol_res_touple = volume.shape
        self.n_voxels = torch.prod(self.vol_res).int().item()
        self.min_idx = torch.tensor([0.0, 0.0, 0.0], dtype=torch.float)
        self.max_idx = self.vol_res - 1.0
        self.volume_indices = _LazyVoxelIndices(volume.shape)
        self.sample_size = sampleSize
        self.max_dim = torch.max(self.max_idx)
        self.scales = self.max_idx / self.max_dim

    def generate_indices(self, start, end, res):
        """(res0, res1, res2, 3) lattice of per-axis ``linspace(start, end, res)`` (IndexDataset.py:69-76)."""
        r = [int(v) for v in res]
        out = torch.zeros(r[0], r[1], r[2], 3)
        shapes = ((r[0], 1, 1), (1, r[1], 1), (1, 1, r[2]))
        for a in range(3):
            out[:, :, :, a] = torch.linspace(float(start[a]), float(end[a]), r[a], dtype=torch.float).view(shapes[a])
        return out

    def move_data_to_device(self, device):
        self.min_idx = self.min_idx.to(device)
        self.max_idx = self.max_idx.to(device)
        self.vol_res = self.vol_res.to(device)
        self.scales = self.scales.to(device)

    def __len__(self):
        return self.n_voxels

    def __getitem__(self, index):
        """``sample_size`` iid uniform voxel positions -> (raw (i,j,k) as float, normalised positions); the DataLoader
        index is ignored, as in the reference."""
        raw = self.volume_indices[torch.randint(0, self.n_voxels, (self.sample_size,))]
        norm = normalize_volume(raw, self.min_idx.cpu().unsqueeze(0), self.max_idx.cpu().unsqueeze(0), -1.0, 1.0)
        return raw, self.scales.cpu().unsqueeze(0) * norm

    # ---- B200 path ------------------------------------------------------------------------------------------------
    def sample(self, n, volume=None, seed=0, sample_offset=0, explicit_idx=None, device='cuda'):
        """n samples in one launch: (raw (n,3), norm (n,3), gt (n,) or None).  Philox(seed, sample_offset + i) picks
        the voxel of sample i unless ``explicit_idx`` (int64 flat voxel indices) is given."""
        return ops.sample(self.vol_res_touple, n, seed=seed, sample_offset=sample_offset, volume=volume,
                          explicit_idx=explicit_idx, want_gt=volume is not None, device=device)
