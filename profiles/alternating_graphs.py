"""Does replaying two CUDA graphs alternately cost more than replaying one?  (The peer-sum data-parallel step used one graph
per step parity.)  Single GPU: the headline step captured twice, replayed as A A A A ... and as A B A B ..."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from latent_feature_grid_compression_b200.training.fast_loop import make_trainer

dev = torch.device('cuda')
cfg = bench.CONFIGS['mhd_p_basic']
vol = bench.synthetic_volume(255, dev)
model = bench.build_model('mhd_p_basic', dev)
tr = make_trainer(model, vol, 255 ** 3, cfg['args'], cfg['args']['lr'], seed=1)
for _ in range(10):
    tr.step()
torch.cuda.synchronize()
ga = tr._graphs[(False, 0)]
gb = torch.cuda.CUDAGraph()
with torch.cuda.graph(gb):
    tr._step_body(False)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for name, seq in (('A A A A', [ga, ga]), ('A B A B', [ga, gb]), ('A A A A', [ga, ga]), ('A B A B', [ga, gb])):
    for _ in range(50):
        seq[0].replay(); seq[1].replay()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(1000):
        seq[0].replay(); seq[1].replay()
    e1.record(); torch.cuda.synchronize()
    print('%s: %.2f us per step' % (name, 1e3 * e0.elapsed_time(e1) / 2000), flush=True)
