"""Where the data-parallel step's extra time goes (needs the -DLFGC_PHASE_TIMING build, LFGC_LIB=...): per launch of
lfgc_peer_sum, the time CTA 0 spends in the in-kernel barrier and in the remote reads + sum, next to the graph-replayed
step time.    torchrun --nproc-per-node 2 profiles/peer_sum_timing.py"""
import ctypes
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import bench
from latent_feature_grid_compression_b200 import _lib
from latent_feature_grid_compression_b200.training.fast_loop import make_trainer

rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
torch.cuda.set_device(int(os.environ['LOCAL_RANK']))
dist.init_process_group('nccl')
dev = torch.device('cuda')
lib = _lib.load()
timed = hasattr(lib, 'lfgc_peer_timing')     # only the -DLFGC_PHASE_TIMING build exports it
if timed:
    lib.lfgc_peer_timing.argtypes = [ctypes.c_void_p, ctypes.c_int]
cfg = bench.CONFIGS['mhd_p_basic']
vol = bench.synthetic_volume(255, dev)
model = bench.build_model('mhd_p_basic', dev)
a = dict(cfg['args'])
a['batch_size'] = a['batch_size'] * world          # weak scaling: the per-rank batch stays 32768
tr = make_trainer(model, vol, 255 ** 3, a, a['lr'], seed=1, rank=rank, world=world)
for _ in range(20):
    tr.step()
torch.cuda.synchronize(); dist.barrier()
if timed:
    lib.lfgc_peer_timing(None, 1)
n = 2000
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(n):
    tr.step()
e1.record(); torch.cuda.synchronize()
buf = (ctypes.c_ulonglong * 8)()
if timed:
    lib.lfgc_peer_timing(buf, 1)
k = max(int(buf[2]), 1)
print('rank %d: step %.2f us; peer_sum CTA 0: barrier wait %.2f us, remote reads + sum %.2f us (%d launches)' % (
    rank, 1e3 * e0.elapsed_time(e1) / n, buf[0] / k / 1e3, buf[1] / k / 1e3, k), flush=True)
# timeline of the last of a few queued steps (timing build): per-sample kernel entry/exit, peer-sum entry/exit, step period
if timed:
    lib.lfgc_btc_time.argtypes = [ctypes.c_void_p, ctypes.c_int]
    torch.cuda.synchronize(); dist.barrier()
    for rep in range(3):
        for _ in range(8):
            tr.step()
        torch.cuda.synchronize()
        bt = (ctypes.c_ulonglong * 4)(); lib.lfgc_btc_time(bt, 0)
        pb = (ctypes.c_ulonglong * 8)(); lib.lfgc_peer_timing(pb, 0)
        us = lambda a, b: (int(a) - int(b)) / 1e3
        print('rank %d: period %.2f us = per-sample kernel %.2f + gap %.2f + peer sum %.2f + rest (grid step, 3 launches) %.2f'
              % (rank, us(bt[0], bt[2]), us(bt[1], bt[0]), us(pb[4], bt[1]), us(pb[5], pb[4]),
                 us(bt[0], bt[2]) - us(bt[1], bt[0]) - us(pb[4], bt[1]) - us(pb[5], pb[4])), flush=True)
# the per-sample kernel alone: gradient accumulators in ordinary device memory vs in the symmetric-memory message buffer
from latent_feature_grid_compression_b200 import ops  # noqa: E402
geom = tr.geom
pc = tr.n_mlp_elems
gg_local = torch.zeros_like(tr.grad_grid)
acc_local = torch.zeros(pc + 1, device=dev)
def run(gg, acc, reps=300):
    for _ in range(20):
        ops.train_step_accumulate(geom, tr.volume, tr.batch, 1, 0, 1e-9, tr.grid_cl, tr.mlp_flat, gg, acc, tr.workspace)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        ops.train_step_accumulate(geom, tr.volume, tr.batch, 1, 0, 1e-9, tr.grid_cl, tr.mlp_flat, gg, acc, tr.workspace)
    e1.record(); torch.cuda.synchronize()
    return 1e3 * e0.elapsed_time(e1) / reps
t_local = run(gg_local, acc_local)
t_symm = run(tr.grad_grid, tr.red_mlp)
print('rank %d: per-sample kernel (back-to-back launches) %.2f us with local accumulators, %.2f us with symmetric-memory ones'
      % (rank, t_local, t_symm), flush=True)
dist.barrier()
os._exit(0)
