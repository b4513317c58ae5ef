#!/usr/bin/env python
"""Benchmark of the fV-SRN latent-feature-grid training hot path on B200 ("Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repository's CUDA path
    python bench.py --impl reference [--gpus N] ...                 # the reference's algorithm on the host cores

Workload (BASELINE.json configs[1], experiment-config-files/mhd_p_basic.txt): synthetic seeded 255^3 volume,
grid_features 16, grid_size 15, hidden 32 x 4 layers, 2 embedding frequencies, db2 wavelet, fp32,
batch 2048 x 16 = 32768 samples per optimiser step and GPU.  One bench "step" is ONE VOLUME PASS (the reference's
own unit of training length, training/training.py:112-114) = ceil(255^3 / 32768) = 507 optimiser steps, each the full
hot path: voxel sampler + ground truth + wavelet synthesis + forward + MSE + backward + synthesis adjoint
(+ one NCCL all-reduce of the flat gradient when N > 1) + Adam.  Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np
import torch

CFG = dict(R=255, C=16, G=15, H=32, L=4, F=2, wavelet='db2', batch=2048 * 16, lr=0.008)
METRIC = 'train_samples_per_s_fwd_bwd'
UNIT = 'samples/s'
FLOPS_PER_SAMPLE = 8192 + 15424      # fwd + bwd MLP FLOPs per sample at C=16, H=32, L=4
HBM_BYTES_PER_SAMPLE = 4             # fused sampler + loss: only the ground-truth voxel is read from HBM


def synthetic_volume(R: int, device):
    """Seeded band-limited field (64 sinusoids, amplitudes ~ 1/|f|), min/max-normalised to [-1, 1] like
    data/IndexDataset.py:15-17.  Same values on CPU and GPU up to fp32 rounding."""
    rng = np.random.default_rng(1234)
    freqs = rng.uniform(-8.0, 8.0, size=(64, 3))
    phase = rng.uniform(0, 2 * np.pi, size=64)
    amp = 1.0 / np.maximum(np.linalg.norm(freqs, axis=1), 1.0)
    ax = torch.linspace(0.0, 1.0, R, device=device)
    vol = torch.zeros(R, R, R, device=device)
    for f, p, a in zip(freqs, phase, amp):
        arg = (2 * math.pi * f[0]) * ax[:, None, None] + (2 * math.pi * f[1]) * ax[None, :, None] \
            + (2 * math.pi * f[2]) * ax[None, None, :] + p
        vol += float(a) * torch.sin(arg)
    mn, mx = vol.min(), vol.max()
    return (2.0 * ((vol - mn) / (mx - mn)) - 1.0).contiguous()


def peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return dict(hbm_gbs=float(d['hbm_gbs']), sm_max_mhz=float(d.get('sm_max_mhz', 1965.0)),
                    bf16_tflops=float(d.get('bf16_tflops', 1654.2)), source='measured')
    return dict(hbm_gbs=6650.0, sm_max_mhz=1965.0, bf16_tflops=1654.2, source='fallback')


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix='.csv')
            os.close(fd)
            self.proc = subprocess.Popen(['nvidia-smi', '--query-gpu=' + self.Q, '--format=csv,noheader,nounits',
                                          '-lms', '100', '-i', str(self.idx)],
                                         stdout=open(self.path, 'w'), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = dict(sm_mhz=None, sm_max_mhz=None, reasons=[], samples=0)
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        try:
            with open(self.path) as f:
                for line in f:
                    parts = [p.strip() for p in line.split(',')]
                    if len(parts) < 9:
                        continue
                    try:
                        sm.append(float(parts[1]))
                        mx.append(float(parts[2]))
                    except ValueError:
                        continue
                    for name, val in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'),
                                         parts[5:9]):
                        if val.lower().startswith('active'):
                            reasons.add(name)
            os.remove(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), samples=len(sm))
        out['reasons'] = sorted(reasons)
        return out


def host_threads() -> int:
    """CPU threads this process may really use: affinity mask and cgroup quota, not the machine's core count
    (a container on a 200-core host with a 16-CPU quota must not start 200 compute threads)."""
    n = os.cpu_count() or 1
    try:
        n = min(n, len(os.sched_getaffinity(0)))
    except Exception:
        pass
    try:
        with open('/sys/fs/cgroup/cpu.max') as f:
            quota, period = f.read().split()
        if quota != 'max':
            n = min(n, max(1, int(int(quota) / int(period))))
    except Exception:
        try:
            q = int(open('/sys/fs/cgroup/cpu/cpu.cfs_quota_us').read())
            per = int(open('/sys/fs/cgroup/cpu/cpu.cfs_period_us').read())
            if q > 0:
                n = min(n, max(1, q // per))
        except Exception:
            pass
    return max(1, n)


def dist_env():
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    return rank, local, world


# ---------------------------------------------------------------------------------------------------------------------
# reference arm: the reference's algorithm on the host cores (oracle port, kind "port")
# ---------------------------------------------------------------------------------------------------------------------

def cpu_port_rate(steps_budget_s: float, opt_steps: int = None, warmup: int = 1):
    """samples/s of the CPU port on the bench workload; runs about `steps_budget_s` seconds unless opt_steps given."""
    from oracle import fvsrn_numpy as O
    from oracle import torch_port as TP
    torch.set_num_threads(host_threads())
    spec = O.Spec(CFG['C'], CFG['G'], CFG['H'], CFG['L'], CFG['F'], CFG['wavelet'], '')
    vol = synthetic_volume(CFG['R'], 'cpu')
    port = TP.CpuPort(spec, TP.make_state(spec, 0), lr=CFG['lr'])
    gen = torch.Generator().manual_seed(0)
    n = CFG['batch']
    for _ in range(max(1, warmup)):
        port.train_step(vol, n, gen)
    t0 = time.perf_counter()
    port.train_step(vol, n, gen)
    t1 = time.perf_counter() - t0
    if opt_steps is None:
        opt_steps = int(min(400, max(3, steps_budget_s / max(t1, 1e-4))))
    t0 = time.perf_counter()
    for _ in range(opt_steps):
        port.train_step(vol, n, gen)
    dt = time.perf_counter() - t0
    # SURVEY 8(d) also asks for (1) forward + loss + backward on a pre-sampled batch and (3) the tiled reconstruction
    # (field_from_net: 32^3-voxel tiles, the grid decoded again for every tile as the reference does)
    coords = torch.rand(n, 3, generator=gen) * 2 - 1
    gt = torch.rand(n, generator=gen) * 2 - 1
    reps = int(min(100, max(3, 3.0 / max(t1, 1e-4))))
    t0 = time.perf_counter()
    for _ in range(reps):
        port.opt.zero_grad(set_to_none=True)
        torch.nn.functional.mse_loss(port.forward(coords).squeeze(-1), gt).backward()
    fb = reps * n / (time.perf_counter() - t0)
    tile = torch.rand(32 * 32 * 32, 3, generator=gen) * 2 - 1
    port.reconstruct(tile)
    t0 = time.perf_counter()
    tiles = 0
    while time.perf_counter() - t0 < 2.0:
        port.reconstruct(tile)
        tiles += 1
    rec = tiles * tile.shape[0] / (time.perf_counter() - t0)
    return dict(rate=opt_steps * n / dt, opt_steps=opt_steps, seconds=dt, threads=torch.get_num_threads(),
                ms_per_opt_step=1e3 * dt / opt_steps, fwd_bwd_rate=fb, reconstruct_rate=rec)


def run_reference(args):
    rank, local, world = dist_env()
    if rank != 0:
        return
    steps_per_pass = math.ceil(CFG['R'] ** 3 / CFG['batch'])
    from oracle import fvsrn_numpy as O
    from oracle import torch_port as TP
    torch.set_num_threads(host_threads())
    spec = O.Spec(CFG['C'], CFG['G'], CFG['H'], CFG['L'], CFG['F'], CFG['wavelet'], '')
    vol = synthetic_volume(CFG['R'], 'cpu')
    port = TP.CpuPort(spec, TP.make_state(spec, 0), lr=CFG['lr'])
    gen = torch.Generator().manual_seed(0)
    n = CFG['batch']
    port.train_step(vol, n, gen)
    t0 = time.perf_counter()
    port.train_step(vol, n, gen)
    t1 = max(time.perf_counter() - t0, 1e-4)
    # bounded sample: as many optimiser steps per bench step as fit the budget for the whole K+W run
    budget_s = float(os.environ.get('LFGC_BENCH_CPU_BUDGET_S', '90'))
    per_step = int(max(1, min(steps_per_pass, budget_s / ((args.steps + args.warmup) * t1))))
    deadline = time.perf_counter() + 2.0 * budget_s          # hard guard: never run away on a slow / shared host
    for _ in range(args.warmup * per_step):
        port.train_step(vol, n, gen)
        if time.perf_counter() > deadline:
            break
    done = 0
    t0 = time.perf_counter()
    for _ in range(args.steps * per_step):
        port.train_step(vol, n, gen)
        done += 1
        if time.perf_counter() > deadline and done >= args.steps:
            break
    dt = time.perf_counter() - t0
    rate = done * n / dt
    sample = '%.1f of the %d optimiser steps of a volume pass per bench step (32768 samples each: sampler + GT + ' \
             'synthesis + fwd + MSE + bwd + Adam), ATen-op port of the reference on %d host threads' % (
                 done / args.steps, steps_per_pass, torch.get_num_threads())
    line = dict(impl='reference', metric=METRIC, value=rate, unit=UNIT, n_gpus=args.gpus, steps=args.steps,
                warmup=args.warmup, ms_per_step=1e3 * dt / args.steps, higher_is_better=True, scaling='weak',
                vs_baseline=None, dtype='f32', data='synthetic',
                config=dict(workload='mhd_p_basic: 255^3 synthetic volume, C16 G15 H32 L4 F2 db2, 32768 samples/optimiser step',
                            step=sample),
                cpu_baseline=dict(value=rate, unit=UNIT, cores=torch.get_num_threads(), kind='port', sample=sample),
                e2e=dict(value=rate, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------------
# this repository's arm
# ---------------------------------------------------------------------------------------------------------------------

def run_native(args):
    import torch.distributed as dist
    from latent_feature_grid_compression_b200.build import build_library
    rank, local, world = dist_env()
    if rank == 0:
        build_library()
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device: the native arm has no CPU fallback')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        # a rank that dies leaves its peers waiting inside a collective: bound the whole run instead of hanging
        import threading
        limit = float(os.environ.get('LFGC_BENCH_TIME_LIMIT_S', '420'))

        def _give_up():
            sys.stderr.write('bench.py: rank %d exceeded %.0f s, aborting\n' % (rank, limit))
            sys.stderr.flush()
            os._exit(3)
        guard = threading.Timer(limit, _give_up)
        guard.daemon = True
        guard.start()
        dist.init_process_group('nccl', device_id=dev)
        dist.barrier()
    from latent_feature_grid_compression_b200 import ops
    from latent_feature_grid_compression_b200.model.model_utils import setup_model
    from latent_feature_grid_compression_b200.training.fast_loop import FastTrainer

    pk = peaks()
    volume = synthetic_volume(CFG['R'], dev)
    torch.manual_seed(0)
    model = setup_model(3, CFG['H'], 1, CFG['L'], 'fourier', CFG['F'], '', 0.1, 0.9, CFG['wavelet'], CFG['C'], CFG['G'], '')
    model.to(dev).train()
    n = CFG['batch']
    trainer = FastTrainer(model, volume, n, lr=CFG['lr'], seed=1234, rank=rank, world_size=world)
    trainer.capture()
    steps_per_pass = math.ceil(CFG['R'] ** 3 / n)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def one_pass():
        for _ in range(steps_per_pass):
            trainer.step()

    for _ in range(max(args.warmup, 3)):
        one_pass()
    torch.cuda.synchronize()

    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    wall0 = time.perf_counter()
    for s0, s1 in evs:
        flush.fill_(1)          # evict L2 between timed steps (outside the event pair)
        s0.record()
        one_pass()
        s1.record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - wall0
    if world > 1:
        dist.barrier()
    total_ms = sum(a.elapsed_time(b) for a, b in evs)
    # back-to-back (hot L2) number for information
    h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    h0.record()
    for _ in range(min(args.steps, 10)):
        one_pass()
    h1.record()
    torch.cuda.synchronize()
    hot_ms = h0.elapsed_time(h1) / min(args.steps, 10)
    clk = clocks.stop() if rank == 0 else None
    t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    samples = args.steps * steps_per_pass * n * world
    value = samples / (total_ms * 1e-3)
    final_loss = trainer.last_loss()

    # ---- dominant kernel alone: lfgc_train_step (fused sampler + fwd + loss + bwd) ---------------------------------
    geom = trainer.geom
    reps = 200
    ke = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for i in range(reps + 5):
        trainer.grad_grid.zero_()
        if i >= 5:
            ke[i - 5][0].record()
        ops.train_step(geom, volume, n, 1234, 0, 1.0 / n, trainer.grid_cl, trainer.mlp_flat, trainer.grad_grid,
                       trainer.flat_g[trainer.mlp_off:], trainer.loss_sum, trainer.workspace,
                       step_dev=trainer.step_dev, step_stride=n)
        if i >= 5:
            ke[i - 5][1].record()
    torch.cuda.synchronize()
    k_ms = float(np.mean([a.elapsed_time(b) for a, b in ke]))
    fp32_peak = 148 * 128 * 2 * pk['sm_max_mhz'] * 1e6 / 1e12
    ach_tflops = FLOPS_PER_SAMPLE * n / (k_ms * 1e-3) / 1e12
    traffic = None
    tpath = os.path.join(ROOT, 'profiles', 'traffic.json')
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get('train_step_dram_bytes_per_launch')
        except Exception:
            traffic = None
    tc_on = os.environ.get('LFGC_BACKWARD_TC', '1') != '0'
    kname = 'backward_tc_kernel<FUSED=1,TPS=2> (lfgc_train_step, tcgen05 3xTF32)' if tc_on \
        else 'backward_v2_kernel<FUSED=1> (lfgc_train_step, FFMA2)'
    roofline = dict(bound='fp32', kernel=kname, achieved=ach_tflops,
                    peak=fp32_peak, unit='TFLOP/s', frac=ach_tflops / fp32_peak, traffic=traffic,
                    kernel_us=k_ms * 1e3, peak_source='148 SMs x 128 FFMA x 2 x %s sm_max_mhz' % pk['source'],
                    note='algorithmic FLOPs = 23616/sample (SURVEY 8d).  Neither HBM nor the tensor pipe bounds this path: the '
                         'contractions run on tcgen05 (tensor pipe 7 % active in ncu, see roofline_tensor), the kernel time '
                         'is the per-sample fp32 work left on the SM (gather, Fourier, SnakeAlt, hi/lo splits, scatter), '
                         'so the CUDA-core FFMA2 peak stays the yardstick (it is what the FFMA2 kernel is bounded by)')
    tf32_peak = pk.get('bf16_tflops', 1654.2) / 2.0
    roofline_tensor = dict(bound='tensor', achieved=ach_tflops, peak=tf32_peak, unit='TFLOP/s',
                           frac=ach_tflops / tf32_peak, executed_over_algorithmic=2.6,
                           peak_source='dense tf32 = measured bf16 cuBLAS peak / 2 (%s)' % pk['source'],
                           note='3xTF32 executes 3 MMAs per sample-major product and 2 per weight-gradient product')
    hbm_ach = HBM_BYTES_PER_SAMPLE * n / (k_ms * 1e-3) / 1e9
    roofline_hbm = dict(bound='hbm', achieved=hbm_ach, peak=pk['hbm_gbs'], unit='GB/s', frac=hbm_ach / pk['hbm_gbs'],
                        traffic=traffic, peak_source=pk['source'])

    # ---- full-volume reconstruction (the path's second metric: decode voxels/s), this rank's slab ---------------------
    recon = reconstruct_rate(model, volume, rank, world, dev, fp32_peak)

    # ---- end to end through the reference-facing nn.Module API with HOST buffers ------------------------------------
    e2e = e2e_host_fed(volume, n, rank, world, dev)
    e2e_module = e2e_module_path(model, volume, n, rank, world, dev)

    line = None
    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            r = cpu_port_rate(15.0)
            cpu = dict(value=r['rate'], unit=UNIT, cores=r['threads'], kind='port',
                       sample='%d optimiser steps of 32768 samples (sampler + GT + synthesis + fwd + MSE + bwd + Adam), '
                              'ATen-op port of the reference, %.1f s' % (r['opt_steps'], r['seconds']),
                       fwd_bwd_samples_per_s=r['fwd_bwd_rate'], reconstruct_voxels_per_s=r['reconstruct_rate'],
                       extra='fwd_bwd: forward + MSE + backward on a pre-sampled batch of 32768; reconstruct: 32^3-voxel '
                             'tiles with the grid decoded per tile (visualization/OutputToVTK.py:7-47), ~2 s each')
        line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=max(args.warmup, 3),
                    ms_per_step=total_ms / args.steps, higher_is_better=True, scaling='weak', vs_baseline=None,
                    dtype='f32', data='synthetic',
                    config=dict(workload='mhd_p_basic: 255^3 synthetic volume, C16 G15 H32 L4 F2 db2, '
                                         '32768 samples/optimiser step/GPU',
                                step='one volume pass = %d optimiser steps (sampler + GT + synthesis + fwd + MSE + bwd '
                                     '+ adjoint%s + Adam), CUDA-graph replay' % (steps_per_pass, ' + NCCL all-reduce' if world > 1 else ''),
                                l2='flushed between timed steps (256 MiB write outside the event pairs)',
                                parallelism='dp%d' % world,
                                switches={k: v for k, v in sorted(os.environ.items()) if k.startswith('LFGC_')}),
                    e2e=e2e, gpu_launches=int(trainer.launches_per_step * steps_per_pass * args.steps),
                    clocks=dict(sm_mhz=clk['sm_mhz'], sm_max_mhz=clk['sm_max_mhz'], reasons=clk['reasons'],
                                samples=clk['samples']),
                    roofline=roofline, roofline_hbm=roofline_hbm, roofline_tensor=roofline_tensor, reconstruct=recon, cpu_baseline=cpu,
                    e2e_module_api=e2e_module,
                    extra=dict(us_per_optimiser_step=1e3 * total_ms / (args.steps * steps_per_pass),
                               hot_l2_samples_per_s=steps_per_pass * n * world / (hot_ms * 1e-3),
                               wall_s_timed_region=wall, final_mse=final_loss,
                               launches_per_optimiser_step=int(trainer.launches_per_step)))
        print(json.dumps(line), flush=True)
    _shutdown(world, dist, trainer)


def _shutdown(world, dist, trainer):
    """Leave promptly once the JSON line is out.  Observed on a 2-GPU box: with NCCL collectives captured in CUDA
    graphs the workers can sit in the process-group / interpreter teardown indefinitely after the result was printed
    (the launcher then waits for them).  So: release the graphs first, tear down cooperatively under a watchdog, and
    hard-exit with status 0 -- nothing after the printed line carries information."""
    import gc
    import threading
    sys.stdout.flush()
    sys.stderr.flush()
    if world <= 1:
        return
    threading.Timer(20.0, lambda: os._exit(0)).start()      # watchdog: never outlive the result by more than 20 s
    try:
        torch.cuda.synchronize()
        trainer._graphs.clear()      # CUDA graphs holding captured NCCL kernels go first
        gc.collect()
        torch.cuda.synchronize()
        dist.barrier()
        dist.destroy_process_group()
    except Exception:
        pass
    sys.stdout.flush()
    sys.stderr.flush()
    os._exit(0)


def reconstruct_rate(model, volume, rank, world, dev, fp32_peak, reps=10):
    """decode voxels/s: field_from_net over this rank's slab of the 255^3 volume (grid decoded once, one launch),
    output left on the device; whole-job rate = sum of the slabs / max time over ranks."""
    import torch.distributed as dist
    from latent_feature_grid_compression_b200.data.IndexDataset import IndexDataset
    from latent_feature_grid_compression_b200.training.parallel import slab_bounds
    from latent_feature_grid_compression_b200.visualization.OutputToVTK import field_from_net
    R = volume.shape[0]
    ds = IndexDataset(torch.zeros(1, 1, 1).expand(R, R, R), 16)
    slab = slab_bounds(R, rank, world)
    model.eval()
    for _ in range(3):
        out = field_from_net(ds, model, True, slab=slab, to_cpu=False)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = field_from_net(ds, model, True, slab=slab, to_cpu=False)
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    model.train()
    ms = float(t.item())
    vox = R ** 3
    tf = 8192.0 * (slab[1] - slab[0]) * R * R / (ms * 1e-3) / 1e12
    return dict(metric='decode_voxels_per_s', value=vox / (ms * 1e-3), unit='voxels/s', ms_per_volume=ms,
                volume='%d^3, z-slab sharded over %d GPU(s), no communication' % (R, world),
                includes='axis tables + mask multipliers + wavelet synthesis + fused sample kernel, output stays in HBM',
                roofline=dict(bound='fp32', achieved=tf, peak=fp32_peak, unit='TFLOP/s', frac=tf / fp32_peak,
                              note='per GPU; 8192 FLOP/voxel; HBM-algorithmic 4 B/voxel written'))


def e2e_host_fed(volume, n, rank, world, dev, steps=300, warmup=20):
    """samples/s of whole optimiser steps fed from HOST memory through the public trainer API
    (FastTrainer.step_host): every step copies that step's positions (n x 3 fp32) and target values (n fp32) from
    pinned host buffers, replays the captured step (synthesis + forward + MSE + backward + adjoint [+ all-reduce] +
    Adam) and reads the loss back to the host."""
    import torch.distributed as dist
    from latent_feature_grid_compression_b200 import ops
    from latent_feature_grid_compression_b200.model.model_utils import setup_model
    from latent_feature_grid_compression_b200.training.fast_loop import FastTrainer
    torch.manual_seed(0)
    m = setup_model(3, CFG['H'], 1, CFG['L'], 'fourier', CFG['F'], '', 0.1, 0.9, CFG['wavelet'], CFG['C'], CFG['G'], '')
    m.to(dev).train()
    tr = FastTrainer(m, volume, n, lr=CFG['lr'], seed=7, rank=rank, world_size=world)
    n_buf = 8
    host = []
    for b in range(n_buf):
        raw, norm, gt = ops.sample(volume.shape, n, seed=99 + rank, sample_offset=b * n, volume=volume, want_gt=True)
        host.append((norm.cpu().pin_memory(), gt.cpu().pin_memory()))
    last = 0.0
    for i in range(warmup):
        tr.step_host(*host[i % n_buf])
        last = tr.last_loss()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        tr.step_host(*host[i % n_buf])
        last = tr.last_loss()          # device -> host read of the step's result, every step
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_serial = float(t.item())
    # pipelined variant: the same per-step traffic (H2D of the step's samples, D2H of the step's loss, each read by the
    # host), with the copies on a second stream and the loss of step i read while step i+1 runs
    for i in range(warmup):
        tr.step_host_pipelined(*host[i % n_buf])
    tr.flush_host_pipeline()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    e0.record()
    for i in range(steps):
        prev = tr.step_host_pipelined(*host[i % n_buf])   # returns the loss of step i-1 (host float)
        if prev is not None:
            last = prev
    last = tr.flush_host_pipeline()
    e1.record()
    torch.cuda.synchronize()
    wall_ms = 1e3 * (time.perf_counter() - t0)
    t = torch.tensor([max(e0.elapsed_time(e1), wall_ms)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    return dict(value=steps * n * world / (ms * 1e-3), unit=UNIT, h2d_bytes_per_step=n * 16, d2h_bytes_per_step=4,
                api='FastTrainer.step_host_pipelined(coords, targets): one optimiser step of 32768 host-resident samples '
                    'per GPU per call (H2D on a copy stream into double-buffered staging, lfgc_train_step with '
                    'caller-supplied samples, CUDA-graph replay, every step\'s loss copied to pinned host memory and read '
                    'one call later); timed = max(CUDA events, host wall clock)',
                us_per_optimiser_step=1e3 * ms / steps, final_mse=last,
                serial=dict(value=steps * n * world / (ms_serial * 1e-3), us_per_optimiser_step=1e3 * ms_serial / steps,
                            api='FastTrainer.step_host + last_loss(): copies, replay and loss read serialised per step'))


def e2e_module_path(model, volume, n, rank, world, dev, steps=100, warmup=10):
    """samples/s through model(coords) / loss.backward() / Adam with pinned HOST inputs every step (coords + ground
    truth), one device->host read of the loss per step; data parallel ranks all-reduce the flattened gradients."""
    import torch.distributed as dist
    from latent_feature_grid_compression_b200 import ops
    from latent_feature_grid_compression_b200.model.model_utils import setup_model
    torch.manual_seed(0)
    m = setup_model(3, CFG['H'], 1, CFG['L'], 'fourier', CFG['F'], '', 0.1, 0.9, CFG['wavelet'], CFG['C'], CFG['G'], '')
    m.to(dev).train()
    opt = torch.optim.Adam(m.parameters(), lr=CFG['lr'])
    crit = torch.nn.MSELoss()
    n_buf = 8
    host = []
    for b in range(n_buf):
        raw, norm, gt = ops.sample(volume.shape, n, seed=99 + rank, sample_offset=b * n, volume=volume, want_gt=True)
        host.append((norm.cpu().pin_memory(), gt.cpu().pin_memory()))
    params = [p for p in m.parameters()]

    def step(i):
        hc, hg = host[i % n_buf]
        coords = hc.to(dev, non_blocking=True)
        gt = hg.to(dev, non_blocking=True)
        opt.zero_grad(set_to_none=True)
        pred = m(coords).squeeze(-1)
        loss = crit(pred, gt)
        loss.backward()
        if world > 1:
            flat = torch.cat([p.grad.reshape(-1) for p in params])
            dist.all_reduce(flat)
            flat /= world
            off = 0
            for p in params:
                p.grad.copy_(flat[off:off + p.numel()].view_as(p.grad))
                off += p.numel()
        opt.step()
        return loss.item()

    for i in range(warmup):
        step(i)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        step(i)
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    return dict(value=steps * n * world / (ms * 1e-3), unit=UNIT, h2d_bytes_per_step=n * 16, d2h_bytes_per_step=4,
                api='Feature_Grid_Model(coords) + MSELoss + backward + torch.optim.Adam, one optimiser step of 32768 '
                    'samples per GPU per call', us_per_optimiser_step=1e3 * ms / steps)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', type=str, default='native', choices=['native', 'reference'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_native(args)


if __name__ == '__main__':
    main()
