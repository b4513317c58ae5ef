"""world_size-2 CPU/gloo tests of the data-parallel recipe: per-rank shards of one global sample stream,
loss pre-scaled by 1/global batch, one sum all-reduce of the flat gradient == the single-process full-batch gradient;
slab partition of the reconstruction.  The numpy oracle stands in for the kernels on the CPU ranks."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from latent_feature_grid_compression_b200.training import parallel
from oracle import fvsrn_numpy as O
from oracle import torch_port as TP


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _flat(grads, names):
    return np.concatenate([np.asarray(grads[n], dtype=np.float64).reshape(-1) for n in names])


def _worker(rank, world, port, out_path):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    spec = O.Spec(4, 15, 32, 4, 2, 'db2', '')
    sd = TP.make_state(spec, seed=7)
    names = sorted(sd)
    shape = (9, 10, 11)
    rng = np.random.default_rng(3)
    vol = rng.uniform(-1, 1, size=shape).astype(np.float32)
    batch, step = 24, 5
    stream = np.random.default_rng(11).integers(0, vol.size, size=1000)      # the ONE global sample stream
    off = parallel.sample_stream_offset(step, rank, batch, world)
    idx = stream[off:off + batch]
    raw, norm = O.sample_positions(idx, shape)
    gt = vol.reshape(-1)[idx].astype(np.float64)
    y, ctx = O.model_forward(sd, spec, norm, training=True, keep=True)
    scale = parallel.loss_scale(batch, world)
    g = O.model_backward(2.0 * scale * (y - gt[:, None]), ctx, spec)
    flat = torch.from_numpy(_flat(g, names))
    dist.all_reduce(flat)                                                     # the path's only collective
    lo, hi = parallel.slab_bounds(shape[0], rank, world)
    slabs = [None] * world
    dist.all_gather_object(slabs, (lo, hi))
    if rank == 0:
        np.savez(out_path, flat=flat.numpy(), slabs=np.asarray(slabs))
    dist.destroy_process_group()


def test_two_rank_gradient_equals_full_batch(tmp_path):
    world = 2
    out = str(tmp_path / 'dp.npz')
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    got = np.load(out)
    # single-process reference over the global batch
    spec = O.Spec(4, 15, 32, 4, 2, 'db2', '')
    sd = TP.make_state(spec, seed=7)
    names = sorted(sd)
    shape = (9, 10, 11)
    vol = np.random.default_rng(3).uniform(-1, 1, size=shape).astype(np.float32)
    batch, step = 24, 5
    stream = np.random.default_rng(11).integers(0, vol.size, size=1000)
    off = parallel.sample_stream_offset(step, 0, batch, world)
    idx = stream[off:off + batch * world]
    raw, norm = O.sample_positions(idx, shape)
    gt = vol.reshape(-1)[idx].astype(np.float64)
    y, ctx = O.model_forward(sd, spec, norm, training=True, keep=True)
    g = O.model_backward((2.0 / (batch * world)) * (y - gt[:, None]), ctx, spec)
    want = _flat(g, names)
    assert np.abs(got['flat'] - want).max() <= 1e-12 * max(np.abs(want).max(), 1.0)
    slabs = got['slabs']
    assert slabs[0][0] == 0 and slabs[-1][1] == shape[0]
    assert all(slabs[i][1] == slabs[i + 1][0] for i in range(world - 1))


def test_slab_bounds_cover_exactly():
    for extent in (1, 7, 150, 255, 1024):
        for world in (1, 2, 3, 4, 8):
            bounds = [parallel.slab_bounds(extent, r, world) for r in range(world)]
            assert bounds[0][0] == 0 and bounds[-1][1] == extent
            sizes = [b - a for a, b in bounds]
            assert all(bounds[i][1] == bounds[i + 1][0] for i in range(world - 1))
            assert max(sizes) - min(sizes) <= 1


def test_rank_offsets_tile_the_stream():
    batch, world = 32768, 8
    seen = []
    for step in range(3):
        for r in range(world):
            o = parallel.sample_stream_offset(step, r, batch, world)
            seen.append((o, o + batch))
    seen.sort()
    assert seen[0][0] == 0
    assert all(seen[i][1] == seen[i + 1][0] for i in range(len(seen) - 1))
