"""Data-parallel FastTrainer ON HARDWARE (round-1 verdict, weak #3): two ranks, each with half of the global batch,
must reach the parameters of ONE rank stepping on the whole batch -- same global Philox sample stream, gradients summed
either by the NCCL all-reduce or by lfgc_grid_step's in-kernel peer reads (LFGC_ALLREDUCE=p2p, symmetric memory) --
and both ranks must hold bit-identical parameters.  Needs >= 2 GPUs (skipped on the single-GPU box):

    gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu -q
"""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _volume(dev):
    xs = [torch.linspace(0, 1, s) for s in (40, 36, 44)]
    v = torch.sin(6 * xs[0])[:, None, None] * torch.cos(4 * xs[1])[None, :, None] + 0.4 * torch.sin(9 * xs[2])[None, None, :]
    return (2 * (v - v.min()) / (v.max() - v.min()) - 1).contiguous().to(dev)


def _model(drop, dev):
    from latent_feature_grid_compression_b200.model.model_utils import setup_model
    torch.manual_seed(11)
    return setup_model(3, 32, 1, 4, 'fourier', 2, drop, 0.1, 0.9, 'db2', 8, 15, '').to(dev).train()


def _worker(rank, world, port, mode, drop, steps, n_global, out):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), LFGC_ALLREDUCE=mode)
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    try:
        from latent_feature_grid_compression_b200.training.fast_loop import FastTrainer
        kw = dict(weight_l1=1e-5, weight_l2=1e-5) if drop else {}
        tr = FastTrainer(_model(drop, dev), _volume(dev), n_global // world, lr=0.008, seed=9, rank=rank, world_size=world,
                         **kw)
        for _ in range(steps):
            tr.step()
        torch.cuda.synchronize()
        gathered = [torch.empty_like(tr.flat_p) for _ in range(world)]
        dist.all_gather(gathered, tr.flat_p)
        if rank == 0:
            out['p'] = tr.flat_p.cpu()
            out['identical'] = all(bool(torch.equal(g, gathered[0])) for g in gathered)
            out['p2p'] = tr._p2p is not None
            out['gstep'] = bool(tr._gstep)
        tr._graphs.clear()
        torch.cuda.synchronize()
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs two GPUs')
@pytest.mark.parametrize('mode,drop', [('p2p', ''), ('nccl', ''), ('nccl', 'smallify')])
def test_two_rank_training_equals_one_rank_on_the_whole_batch(mode, drop):
    import torch.multiprocessing as mp
    from latent_feature_grid_compression_b200.training.fast_loop import FastTrainer
    steps, n_global = 9, 4096
    dev = torch.device('cuda', 0)
    kw = dict(weight_l1=1e-5, weight_l2=1e-5) if drop else {}
    ref = FastTrainer(_model(drop, dev), _volume(dev), n_global, lr=0.008, seed=9, **kw)
    for _ in range(steps):
        ref.step()
    torch.cuda.synchronize()
    want = ref.flat_p.cpu()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), mode, drop, steps, n_global, out), nprocs=2, join=True)
    assert out['identical']                       # redundant Adam on identical sums: ranks stay bit-identical
    assert out['p2p'] == (mode == 'p2p') and out['gstep'] == (drop == '')
    got = out['p']
    assert float((got - want).abs().max()) <= 1e-5 * float(want.abs().max())
