#!/usr/bin/env python
"""Benchmark of the fV-SRN latent-feature-grid hot path on B200 ("Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config NAME]   # this repository's CUDA path
    python bench.py --impl reference [--gpus N] ...                        # the reference itself on the host cores

Headline workload (BASELINE.json configs[1], experiment-config-files/mhd_p_basic.txt): synthetic seeded 255^3 volume,
grid_features 16, grid_size 15, hidden 32 x 4 layers, 2 embedding frequencies, db2 wavelet, fp32, batch 2048 x 16 =
32768 samples per optimiser step and GPU.  One bench "step" is ONE VOLUME PASS (the reference's own unit of training
length, training/training.py:112-114) = ceil(255^3 / 32768) = 507 optimiser steps, each the full hot path: voxel sampler
+ ground truth + wavelet synthesis + forward + MSE + backward + synthesis adjoint (+ the data-parallel gradient sum when
N > 1) + Adam.  Prints ONE JSON line (rank 0).

The other BASELINE.json configurations ride along under ``extra.configs`` (test_vol / turbulence 150^3, mhd_p_smallify,
mhd_p_dynamic_variational, the wide C32/G64 grid on a 1024^3 volume), each with samples/s, voxels/s and the roofline of
its dominant kernel; ``extra.psnr`` holds the final PSNR of the FAST loop (full two-phase schedule, 3 Philox seeds) next
to the reference's own PSNR on the same volume and config (tests/golden/psnr_configs.json, made by running the unmodified
reference driver in the build container) and both seed-to-seed spreads.
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

METRIC = 'train_samples_per_s_fwd_bwd'
UNIT = 'samples/s'
HBM_BYTES_PER_SAMPLE = 4             # fused sampler + loss: only the ground-truth voxel is read from HBM

# BASELINE.json `configs`, as the reference's experiment-config files define them (values restated here because
# /root/reference does not exist on the GPU box).  args = the dict Feature_Grid_Training.py would hand to training().
_COMMON = dict(d_in=3, d_out=1, n_layers=4, n_hidden_size=32, embedding_type='fourier', n_embedding_freq=2,
               wavelet_filter='db2', lr=0.008, pass_decay=20, lr_decay=0.2, smallify_decay=0, sample_size=16,
               grid_features=16, grid_size=15, checkpoint_path='', drop_type='', drop_momentum=0.1, drop_threshold=0.9,
               lambda_drop_loss=1.0, lambda_weight_loss=2.0, variational_sigma=-3.5, weight_dkl_multiplier=5e-05)
CONFIGS = {
    'mhd_p_basic': dict(R=255, file='mhd_p_basic.txt', args=dict(_COMMON, batch_size=2048, max_pass=60)),
    'mhd_p_smallify': dict(R=255, file='mhd_p_smallify.txt', args=dict(
        _COMMON, batch_size=2048, max_pass=60, drop_type='smallify', drop_momentum=0.025, drop_threshold=0.75,
        lambda_drop_loss=1e-08, lambda_weight_loss=1e-08, variational_sigma=0, weight_dkl_multiplier=0)),
    'mhd_p_dynamic_variational': dict(R=255, file='mhd_p_dynamic_variational.txt', args=dict(
        _COMMON, batch_size=2048, max_pass=60, drop_type='variational_dynamic', lambda_drop_loss=0.1,
        variational_sigma=0, weight_dkl_multiplier=3.0e-05)),
    'test_vol': dict(R=150, file='test_impl_test.txt', args=dict(
        _COMMON, batch_size=1024, max_pass=50, drop_type='smallify', drop_momentum=0.025, drop_threshold=0.75,
        lambda_drop_loss=1e-08, lambda_weight_loss=1e-08, variational_sigma=-3.2)),
    'turbulence_basic': dict(R=150, file='turbulence_basic.txt', args=dict(
        _COMMON, batch_size=1024, max_pass=50, drop_momentum=0.025, drop_threshold=0.75, lambda_drop_loss=1e-08,
        lambda_weight_loss=1e-08, variational_sigma=-3.2)),
    # BASELINE.json configs[4]: synthetic 1024^3 volume, wide latent grid; 2^18 samples per optimiser step and GPU
    'wide': dict(R=1024, file=None, args=dict(_COMMON, batch_size=16384, max_pass=2, grid_features=32, grid_size=64)),
}
GOLDEN_NAME = {'test_vol': 'test_impl_test'}      # name of the config in tests/golden/psnr_configs.json


def flops_per_sample(C, H=32, L=4, F=2):
    """2 (in H + (L-1) H^2 + H) forward + 2 (in H + C H + 2 (L-1) H^2 + 2 H) backward."""
    n_in = 3 + 6 * F + C
    fwd = 2 * (n_in * H + (L - 1) * H * H + H)
    bwd = 2 * (n_in * H + C * H + 2 * (L - 1) * H * H + 2 * H)
    return fwd, bwd


def synthetic_volume(R: int, device):
    """Seeded band-limited field (64 sinusoids, amplitudes ~ 1/|f|), min/max-normalised to [-1, 1] like
    data/IndexDataset.py:15-17.  Same values on CPU and GPU up to fp32 rounding."""
    rng = np.random.default_rng(1234)
    freqs = rng.uniform(-8.0, 8.0, size=(64, 3))
    phase = rng.uniform(0, 2 * np.pi, size=64)
    amp = 1.0 / np.maximum(np.linalg.norm(freqs, axis=1), 1.0)
    ax = torch.linspace(0.0, 1.0, R, device=device)
    vol = torch.zeros(R, R, R, device=device)
    slab = max(1, min(R, (1 << 26) // (R * R)))      # bound the temporaries (1024^3: 4 GiB per full-size temporary)
    for f, p, a in zip(freqs, phase, amp):
        ay = (2 * math.pi * f[1]) * ax[None, :, None] + (2 * math.pi * f[2]) * ax[None, None, :] + p
        for z0 in range(0, R, slab):
            z1 = min(R, z0 + slab)
            vol[z0:z1] += float(a) * torch.sin((2 * math.pi * f[0]) * ax[z0:z1, None, None] + ay)
    mn, mx = vol.min(), vol.max()
    vol -= mn
    vol *= 2.0 / float(mx - mn)
    vol -= 1.0
    return vol.contiguous()


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


def workload_string(name):
    c = CONFIGS[name]
    a = c['args']
    return '%s: %d^3 synthetic volume, C%d G%d H%d L%d F%d %s%s, %d samples/optimiser step/GPU' % (
        name, c['R'], a['grid_features'], a['grid_size'], a['n_hidden_size'], a['n_layers'], a['n_embedding_freq'],
        a['wavelet_filter'], (', ' + a['drop_type']) if a['drop_type'] else '', a['batch_size'] * a['sample_size'])


# ---------------------------------------------------------------------------------------------------------------------
# reference arms: the UNMODIFIED reference modules (baseline/_ref, installed by baseline/install_ref.py), timed on the
# host cores (--impl reference, cpu_baseline) or -- informative -- on the B200 through stock torch eager
# (torch_eager_b200).  Fallback when baseline/_ref is absent: the ATen-op port oracle/torch_port.py (kind "port").
# ---------------------------------------------------------------------------------------------------------------------

REF_DIR = os.path.join(ROOT, 'baseline', '_ref')


def reference_available():
    return os.path.isfile(os.path.join(REF_DIR, 'training', 'training.py'))


class ReferenceRunner:
    """One optimiser step written with the reference's own objects exactly as solve_model does (training/training.py:
    89-138): its IndexDataset + DataLoader, setup_model, trilinear_f_interpolation, MSELoss / SmallifyLoss /
    VariationalDropoutLoss, torch.optim.Adam.  Runs in the current process; call only after install_reference()."""

    def __init__(self, name, device, workers, volume=None):
        from torch.utils.data import DataLoader
        from data.IndexDataset import IndexDataset                      # the reference's modules (baseline/_ref)
        from data.Interpolation import trilinear_f_interpolation
        from model.model_utils import setup_model
        from model.Smallify_Dropout import SmallifyLoss
        from model.Variational_Dropout_Layer import Variance_Model, VariationalDropoutLoss
        cfg = CONFIGS[name]
        a = self.args = cfg['args']
        self.device = device
        vol = synthetic_volume(cfg['R'], 'cpu') if volume is None else volume
        self.dataset = IndexDataset(vol, a['sample_size'])
        self.loader = DataLoader(self.dataset, batch_size=a['batch_size'], shuffle=True, num_workers=workers,
                                 persistent_workers=workers > 0)
        self.it = iter(self.loader)
        self.volume = vol.to(device)
        torch.manual_seed(0)
        self.model = setup_model(a['d_in'], a['n_hidden_size'], a['d_out'], a['n_layers'], a['embedding_type'],
                                 a['n_embedding_freq'], a['drop_type'], a['drop_momentum'], a['drop_threshold'],
                                 a['wavelet_filter'], a['grid_features'], a['grid_size'], '').to(device).train()
        self.opt = torch.optim.Adam(self.model.parameters(), lr=a['lr'])
        self.crit = torch.nn.MSELoss().to(device)
        self.tri = trilinear_f_interpolation
        self.drop_loss, self.var_model = None, None
        drop = a['drop_type']
        if drop and 'variational' in drop:
            self.drop_loss = VariationalDropoutLoss(size_volume=self.dataset.n_voxels,
                                                    batch_size=a['batch_size'] * a['sample_size'],
                                                    weight_dkl=a['lambda_drop_loss'], weight_weights=a['lambda_weight_loss'])
            if 'dynamic' in drop:
                self.var_model = Variance_Model().to(device).train()
                self.opt.add_param_group({'params': self.var_model.parameters()})
        elif drop:
            self.drop_loss = SmallifyLoss(weight_l1=a['lambda_drop_loss'], weight_l2=a['lambda_weight_loss'])
        self.n = a['batch_size'] * a['sample_size']
        self.mi, self.ma, self.res = (self.dataset.min_idx.to(device), self.dataset.max_idx.to(device),
                                      self.dataset.vol_res.to(device))

    def step(self):
        a, dev = self.args, self.device
        try:
            raw, norm = next(self.it)
        except StopIteration:
            self.it = iter(self.loader)
            raw, norm = next(self.it)
        raw = raw.to(dev).view(-1, a['d_in'])
        norm = norm.to(dev).view(-1, a['d_in'])
        norm.requires_grad = True
        self.opt.zero_grad()
        pred = self.model(norm).squeeze(-1)
        gt = self.tri(raw, self.volume, self.mi, self.ma, self.res)
        if self.drop_loss is not None and 'variational' in a['drop_type']:
            if self.var_model is not None:
                var = self.var_model(norm).squeeze(-1)
            else:
                var = torch.ones_like(pred).fill_(a['variational_sigma'])
            loss = self.drop_loss(self.model, pred, gt, var, a['weight_dkl_multiplier'])[0]
        else:
            loss = self.crit(pred, gt)
            if self.drop_loss is not None:
                loss = loss + self.drop_loss(self.model)
        loss.backward()
        self.opt.step()
        return loss

    def close(self):
        self.it = None
        self.loader = None


def install_reference():
    """Put baseline/_ref and the third-party stand-ins on the import path (this process then resolves `model`, `data`,
    ... to the reference)."""
    sys.path.insert(0, os.path.join(ROOT, 'baseline'))
    import ref_shims
    ref_shims.install(REF_DIR)


def time_reference_steps(name, device, budget_s, max_steps=None, warmup=2, workers=None):
    """(steps, seconds, workers, threads) of the reference's optimiser step on `device`."""
    threads = host_threads()
    torch.set_num_threads(threads)
    if workers is None:
        workers = min(8, threads)            # num_workers = 8 in every shipped config
    run = ReferenceRunner(name, device, workers)
    sync = torch.cuda.synchronize if device != 'cpu' and str(device) != 'cpu' else (lambda: None)
    for _ in range(warmup):
        run.step()
    sync()
    t0 = time.perf_counter()
    run.step()
    sync()
    t1 = max(time.perf_counter() - t0, 1e-5)
    steps = int(max(3, budget_s / t1))
    if max_steps is not None:
        steps = min(steps, max_steps)
    t0 = time.perf_counter()
    for _ in range(steps):
        run.step()
    sync()
    dt = time.perf_counter() - t0
    n = run.n
    run.close()
    return dict(steps=steps, seconds=dt, workers=workers, threads=threads, samples=n, rate=steps * n / dt)


def _subprocess_json(argv, timeout):
    """Run `python bench.py <argv>` and return the last JSON line (the reference modules shadow the product's module
    names, so reference legs of the native arm live in their own process)."""
    try:
        out = subprocess.run([sys.executable, os.path.abspath(__file__)] + argv, capture_output=True, text=True,
                             timeout=timeout, cwd=ROOT)
        lines = [l for l in out.stdout.splitlines() if l.startswith('{')]
        if lines:
            return json.loads(lines[-1])
        return dict(error=(out.stderr or 'no output')[-400:])
    except Exception as e:   # noqa: BLE001
        return dict(error=repr(e)[:400])


def cpu_port_rate(name, steps_budget_s: float):
    """Fallback CPU baseline: the ATen-op port of the reference's step (oracle/torch_port.py)."""
    from oracle import fvsrn_numpy as O
    from oracle import torch_port as TP
    cfg = CONFIGS[name]
    a = cfg['args']
    torch.set_num_threads(host_threads())
    spec = O.Spec(a['grid_features'], a['grid_size'], a['n_hidden_size'], a['n_layers'], a['n_embedding_freq'],
                  a['wavelet_filter'], '')
    vol = synthetic_volume(cfg['R'], 'cpu')
    port = TP.CpuPort(spec, TP.make_state(spec, 0), lr=a['lr'])
    gen = torch.Generator().manual_seed(0)
    n = a['batch_size'] * a['sample_size']
    port.train_step(vol, n, gen)
    t0 = time.perf_counter()
    port.train_step(vol, n, gen)
    t1 = time.perf_counter() - t0
    steps = int(min(400, max(3, steps_budget_s / max(t1, 1e-4))))
    t0 = time.perf_counter()
    for _ in range(steps):
        port.train_step(vol, n, gen)
    dt = time.perf_counter() - t0
    return dict(steps=steps, seconds=dt, workers=0, threads=torch.get_num_threads(), samples=n, rate=steps * n / dt)


def run_cpu_leg(args):
    """`--leg cpu_baseline|torch_eager`: a bounded sample of the reference's step, one JSON line (child process of the
    native arm)."""
    name = args.config
    if args.leg == 'torch_eager':
        if not reference_available() or not torch.cuda.is_available():
            print(json.dumps(dict(unavailable='needs baseline/_ref and a GPU')))
            return
        install_reference()
        r = time_reference_steps(name, torch.device('cuda', 0), args.budget, max_steps=400)
        # also: forward + loss + backward only, on a pre-sampled batch (no DataLoader, no Adam)
        print(json.dumps(dict(value=r['rate'], unit=UNIT, ms_per_optimiser_step=1e3 * r['seconds'] / r['steps'],
                              steps=r['steps'], dataloader_workers=r['workers'],
                              what='the UNMODIFIED reference modules (baseline/_ref) with device=cuda on this B200: its '
                                   'DataLoader + IndexDataset, F.grid_sample, conv_transpose3d, nn.Linear, autograd, '
                                   'torch.optim.Adam -- the "existing Blackwell kernels" yardstick (SURVEY section 0)')))
        return
    if reference_available():
        install_reference()
        r = time_reference_steps(name, 'cpu', args.budget)
        kind = 'reference'
    else:
        r = cpu_port_rate(name, args.budget)
        kind = 'port'
    print(json.dumps(dict(value=r['rate'], unit=UNIT, cores=r['threads'], kind=kind,
                          sample='%d optimiser steps of %d samples (%s), %.1f s, %d torch threads + %d DataLoader workers'
                                 % (r['steps'], r['samples'],
                                    'the reference\'s own IndexDataset/DataLoader + model + trilinear_f_interpolation + loss '
                                    '+ backward + Adam' if kind == 'reference' else 'ATen-op port of the reference step',
                                    r['seconds'], r['threads'], r['workers']))))


def run_reference(args):
    rank, local, world = dist_env()
    if rank != 0:
        return
    name = args.config
    a = CONFIGS[name]['args']
    n = a['batch_size'] * a['sample_size']
    steps_per_pass = math.ceil(CONFIGS[name]['R'] ** 3 / n)
    budget_s = float(os.environ.get('LFGC_BENCH_CPU_BUDGET_S', '90'))
    if reference_available():
        install_reference()
        threads = host_threads()
        torch.set_num_threads(threads)
        workers = min(8, threads)
        runner = ReferenceRunner(name, 'cpu', workers)
        step = runner.step
        kind = 'reference'
        what = 'the UNMODIFIED reference (baseline/_ref): IndexDataset + DataLoader(%d workers) + Feature_Grid_Model + ' \
               'trilinear_f_interpolation + loss + backward + torch.optim.Adam on %d host threads' % (workers, threads)
    else:
        from oracle import fvsrn_numpy as O
        from oracle import torch_port as TP
        torch.set_num_threads(host_threads())
        spec = O.Spec(a['grid_features'], a['grid_size'], a['n_hidden_size'], a['n_layers'], a['n_embedding_freq'],
                      a['wavelet_filter'], '')
        vol = synthetic_volume(CONFIGS[name]['R'], 'cpu')
        port = TP.CpuPort(spec, TP.make_state(spec, 0), lr=a['lr'])
        gen = torch.Generator().manual_seed(0)
        step = lambda: port.train_step(vol, n, gen)   # noqa: E731
        kind = 'port'
        what = 'ATen-op port of the reference step (baseline/_ref absent) on %d host threads' % torch.get_num_threads()
    step()
    t0 = time.perf_counter()
    step()
    t1 = max(time.perf_counter() - t0, 1e-4)
    # bounded sample: as many optimiser steps per bench step as fit the budget for the whole K+W run
    per_step = int(max(1, min(steps_per_pass, budget_s / ((args.steps + args.warmup) * t1))))
    deadline = time.perf_counter() + 2.0 * budget_s          # hard guard: never run away on a slow / shared host
    for _ in range(args.warmup * per_step):
        step()
        if time.perf_counter() > deadline:
            break
    done = 0
    t0 = time.perf_counter()
    for _ in range(args.steps * per_step):
        step()
        done += 1
        if time.perf_counter() > deadline and done >= args.steps:
            break
    dt = time.perf_counter() - t0
    rate = done * n / dt
    sample = '%.1f of the %d optimiser steps of a volume pass per bench step (%d samples each): %s' % (
        done / args.steps, steps_per_pass, n, what)
    line = dict(impl='reference', metric=METRIC, value=rate, unit=UNIT, n_gpus=args.gpus, steps=args.steps,
                warmup=args.warmup, ms_per_step=1e3 * dt / args.steps, higher_is_better=True, scaling='weak',
                vs_baseline=None, dtype='f32', data='synthetic',
                config=dict(workload=workload_string(name), step=sample),
                cpu_baseline=dict(value=rate, unit=UNIT, cores=torch.get_num_threads(), kind=kind, sample=sample),
                e2e=dict(value=rate, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line), flush=True)
    if kind == 'reference':
        runner.close()
    sys.stdout.flush()
    os._exit(0)      # DataLoader worker processes must not keep the arm alive


# ---------------------------------------------------------------------------------------------------------------------
# this repository's arm
# ---------------------------------------------------------------------------------------------------------------------

def build_model(name, dev, seed=0):
    from latent_feature_grid_compression_b200.model.model_utils import setup_model
    a = CONFIGS[name]['args']
    torch.manual_seed(seed)
    m = setup_model(a['d_in'], a['n_hidden_size'], a['d_out'], a['n_layers'], a['embedding_type'],
                    a['n_embedding_freq'], a['drop_type'], a['drop_momentum'], a['drop_threshold'], a['wavelet_filter'],
                    a['grid_features'], a['grid_size'], '')
    return m.to(dev).train()


def kernel_name(trainer, a):
    tc = os.environ.get('LFGC_BACKWARD_TC', '1') != '0' and trainer.var_cfg is None
    if tc:
        return 'backward_tc_kernel<FUSED=1,TPS=2> (lfgc_train_step, tcgen05 3xTF32)'
    return 'backward_v2_kernel<FUSED=1> (lfgc_train_step%s, FFMA2)' % ('_weighted' if trainer.var_cfg is not None else '')


def measure_config(name, volume, rank, world, dev, pk, opt_steps=300, with_recon=True):
    """Whole-step throughput (CUDA-graph replay, L2 flushed before the timed run), the fused per-sample kernel alone
    and the reconstruction rate of one BASELINE configuration, with its rooflines."""
    import torch.distributed as dist
    from latent_feature_grid_compression_b200 import ops
    from latent_feature_grid_compression_b200.training.fast_loop import make_trainer
    cfg = CONFIGS[name]
    a = cfg['args']
    n = a['batch_size'] * a['sample_size']           # per GPU (weak scaling)
    model = build_model(name, dev)
    args_w = dict(a, batch_size=a['batch_size'] * world)
    trainer = make_trainer(model, volume, cfg['R'] ** 3, args_w, a['lr'], seed=1234, rank=rank, world=world)
    trainer.capture()
    for _ in range(20):
        trainer.step()
    torch.cuda.synchronize()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    flush.fill_(1)
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(opt_steps):
        trainer.step()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    us_step = 1e3 * ms / opt_steps
    fwd, bwd = flops_per_sample(a['grid_features'], a['n_hidden_size'], a['n_layers'], a['n_embedding_freq'])
    fp32_peak = 148 * 128 * 2 * pk['sm_max_mhz'] * 1e6 / 1e12
    out = dict(workload=workload_string(name), value=opt_steps * n * world / (ms * 1e-3), unit=UNIT,
               us_per_optimiser_step=us_step, launches_per_optimiser_step=int(trainer.launches_per_step),
               grid_step_kernel=bool(trainer._gstep), final_mse=trainer.last_loss())
    # the fused per-sample kernel alone (MSE objective; the variational step adds its log-likelihood terms)
    if trainer.var_cfg is None:
        geom = trainer.geom
        reps = 100
        gm = torch.empty(geom.mlp_param_count, device=dev)
        ke = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
        for i in range(reps + 5):
            trainer.grad_grid.zero_()
            if i >= 5:
                ke[i - 5][0].record()
            # the fused per-sample kernel ALONE (its MLP-gradient partial sums stay in the workspace; their reduction is a
            # separate launch / part of lfgc_grid_step)
            ops.train_step_partials(geom, volume, n, 1234, 0, 1.0 / n, trainer.grid_cl, trainer.mlp_flat,
                                    trainer.grad_grid, trainer.workspace, step_dev=trainer.step_dev, step_stride=n,
                                    tc_panels=getattr(trainer, '_tc_panels', None))   # as the trainer launches it
            if i >= 5:
                ke[i - 5][1].record()
        torch.cuda.synchronize()
        k_us = 1e3 * float(np.mean([x.elapsed_time(y) for x, y in ke]))
        ach = (fwd + bwd) * n / (k_us * 1e-6) / 1e12
        out['roofline'] = dict(bound='fp32', kernel=kernel_name(trainer, a), achieved=ach, peak=fp32_peak,
                               unit='TFLOP/s', frac=ach / fp32_peak, kernel_us=k_us,
                               flops_per_sample=fwd + bwd)
        # what is left of the step: synthesis + adjoint + Adam (+ reductions): for the wide grid this is HBM traffic
        n_params = trainer.flat_p.numel()
        grid_us = max(us_step - k_us, 1e-3)
        alg_bytes = 44.0 * n_params           # IDWT fwd + adjoint ~ 16 B, Adam 28 B per parameter
        out['roofline_grid_work'] = dict(bound='hbm', achieved=alg_bytes / (grid_us * 1e-6) / 1e9, peak=pk['hbm_gbs'],
                                         unit='GB/s', frac=alg_bytes / (grid_us * 1e-6) / 1e9 / pk['hbm_gbs'],
                                         us=grid_us, params=n_params,
                                         note='step time minus the per-sample kernel; 44 B per parameter algorithmic '
                                              '(synthesis + adjoint + Adam)')
    if with_recon:
        out['reconstruct'] = reconstruct_rate(model, volume, rank, world, dev, fp32_peak, fwd, reps=5)
    trainer._graphs.clear()
    del trainer, model
    torch.cuda.empty_cache()
    return out


def measure_strong(name, volume, rank, world, dev, opt_steps=300):
    """The reference's global batch (batch_size * sample_size) split over the ranks: what train_volume(args, world=N)
    runs.  At the shipped batch sizes a rank keeps only a fraction of a wave of tiles, so this cannot scale; it is
    reported because weak scaling alone (the headline) trains a different problem."""
    import torch.distributed as dist
    from latent_feature_grid_compression_b200.training.fast_loop import make_trainer
    cfg = CONFIGS[name]
    a = cfg['args']
    model = build_model(name, dev)
    tr = make_trainer(model, volume, cfg['R'] ** 3, a, a['lr'], seed=99, rank=rank, world=world)
    tr.capture()
    for _ in range(20):
        tr.step()
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(opt_steps):
        tr.step()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    n_global = a['batch_size'] * a['sample_size']
    tr._graphs.clear()
    return dict(value=opt_steps * n_global / (ms * 1e-3), unit=UNIT, us_per_optimiser_step=1e3 * ms / opt_steps,
                global_batch=n_global, per_gpu_batch=n_global // world, scaling='strong')


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
        limit = float(os.environ.get('LFGC_BENCH_TIME_LIMIT_S', '540'))

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
    from latent_feature_grid_compression_b200.training.fast_loop import make_trainer

    name = args.config
    cfg = CONFIGS[name]
    a = cfg['args']
    pk = peaks()
    volume = synthetic_volume(cfg['R'], dev)
    n = a['batch_size'] * a['sample_size']
    model = build_model(name, dev)
    trainer = make_trainer(model, volume, cfg['R'] ** 3, dict(a, batch_size=a['batch_size'] * world), a['lr'], seed=1234,
                           rank=rank, world=world)
    trainer.capture()
    steps_per_pass = math.ceil(cfg['R'] ** 3 / n)
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
    total_ms = sum(x.elapsed_time(y) for x, y in evs)
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
    launches = int(trainer.launches_per_step)
    gstep = bool(trainer._gstep)
    allreduce = 'none' if world == 1 else ('lfgc_grid_step peer reads over NVLink (in-kernel barrier, no collective)'
                                           if trainer._p2p is not None else 'NCCL all-reduce')

    # ---- dominant kernel alone: lfgc_train_step (fused sampler + fwd + loss + bwd) ---------------------------------
    fwd, bwd = flops_per_sample(a['grid_features'], a['n_hidden_size'], a['n_layers'], a['n_embedding_freq'])
    fp32_peak = 148 * 128 * 2 * pk['sm_max_mhz'] * 1e6 / 1e12
    roofline = roofline_hbm = roofline_tensor = None
    if trainer.var_cfg is None:
        geom = trainer.geom
        reps = 200
        gm = torch.empty(geom.mlp_param_count, device=dev)
        ke = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
        for i in range(reps + 5):
            trainer.grad_grid.zero_()
            if i >= 5:
                ke[i - 5][0].record()
            # the fused per-sample kernel ALONE (its MLP-gradient partial sums stay in the workspace; their reduction is a
            # separate launch / part of lfgc_grid_step)
            ops.train_step_partials(geom, volume, n, 1234, 0, 1.0 / n, trainer.grid_cl, trainer.mlp_flat,
                                    trainer.grad_grid, trainer.workspace, step_dev=trainer.step_dev, step_stride=n,
                                    tc_panels=getattr(trainer, '_tc_panels', None))   # as the trainer launches it
            if i >= 5:
                ke[i - 5][1].record()
        torch.cuda.synchronize()
        k_ms = float(np.mean([x.elapsed_time(y) for x, y in ke]))
        ach_tflops = (fwd + bwd) * n / (k_ms * 1e-3) / 1e12
        traffic = None
        tpath = os.path.join(ROOT, 'profiles', 'traffic.json')
        if os.path.exists(tpath):
            try:
                traffic = json.load(open(tpath)).get('train_step_dram_bytes_per_launch')
            except Exception:
                traffic = None
        roofline = dict(bound='fp32', kernel=kernel_name(trainer, a), achieved=ach_tflops,
                        peak=fp32_peak, unit='TFLOP/s', frac=ach_tflops / fp32_peak, traffic=traffic,
                        kernel_us=k_ms * 1e3, peak_source='148 SMs x 128 FFMA x 2 x %s sm_max_mhz' % pk['source'],
                        note='algorithmic FLOPs = %d/sample (SURVEY 8d).  Neither HBM nor the tensor pipe bounds this path: '
                             'the contractions run on tcgen05 (see roofline_tensor), the kernel time is the per-sample fp32 '
                             'work left on the SM (gather, Fourier, SnakeAlt, hi/lo splits, scatter) and its latency '
                             'chain, so the CUDA-core FFMA2 peak stays the yardstick (it is what the FFMA2 kernel is '
                             'bounded by).  Timed as lfgc_train_step_partials = this kernel alone' % (fwd + bwd))
        tf32_peak = pk.get('bf16_tflops', 1654.2) / 2.0
        roofline_tensor = dict(bound='tensor', achieved=ach_tflops, peak=tf32_peak, unit='TFLOP/s',
                               frac=ach_tflops / tf32_peak, executed_over_algorithmic=2.6,
                               peak_source='dense tf32 = measured bf16 cuBLAS peak / 2 (%s)' % pk['source'],
                               note='3xTF32 executes 3 MMAs per sample-major product and 2 per weight-gradient product')
        hbm_ach = HBM_BYTES_PER_SAMPLE * n / (k_ms * 1e-3) / 1e9
        roofline_hbm = dict(bound='hbm', achieved=hbm_ach, peak=pk['hbm_gbs'], unit='GB/s',
                            frac=hbm_ach / pk['hbm_gbs'], traffic=traffic, peak_source=pk['source'])

    # ---- full-volume reconstruction (the path's second metric: decode voxels/s), this rank's slab ---------------------
    recon = reconstruct_rate(model, volume, rank, world, dev, fp32_peak, fwd)

    # ---- end to end through the public API with HOST buffers -----------------------------------------------------------
    e2e = e2e_host_fed(name, volume, n, rank, world, dev) if trainer.var_cfg is None else None
    e2e_module = e2e_module_path(name, volume, n, rank, world, dev)
    if e2e is None:
        e2e = e2e_module
    trainer._graphs.clear()
    del trainer
    torch.cuda.empty_cache()

    # ---- strong scaling (N > 1): the reference's GLOBAL batch split over the ranks -- the same optimisation problem ------
    strong = None
    if world > 1 and (a['batch_size'] * a['sample_size']) % world == 0:
        try:
            strong = measure_strong(name, volume, rank, world, dev)
        except Exception as e:   # noqa: BLE001
            strong = dict(error=repr(e)[:300])

    # ---- the other BASELINE configurations and the PSNR of the fast loop (N = 1 ... world) ----------------------------
    extra_cfgs, psnr = {}, None
    if not args.no_extras:
        vols = {cfg['R']: volume}
        for other in ('turbulence_basic', 'test_vol', 'mhd_p_basic', 'mhd_p_smallify', 'mhd_p_dynamic_variational', 'wide'):
            if other == name:
                continue
            R = CONFIGS[other]['R']
            if R not in vols:
                for k in [k for k in vols if k != cfg['R']]:
                    del vols[k]
                torch.cuda.empty_cache()
                vols[R] = synthetic_volume(R, dev)
            try:
                extra_cfgs[other] = measure_config(other, vols[R], rank, world, dev, pk,
                                                   opt_steps=100 if other == 'wide' else 300)
            except Exception as e:   # noqa: BLE001  (an extra must never take the headline down)
                extra_cfgs[other] = dict(error=repr(e)[:300])
        del vols
        torch.cuda.empty_cache()
        if world == 1:
            psnr = psnr_section(dev)

    line = None
    if rank == 0:
        cpu = eager = None
        if world == 1 and not args.no_cpu_baseline:
            cpu = _subprocess_json(['--leg', 'cpu_baseline', '--config', name, '--budget', '15'], 240)
            eager = _subprocess_json(['--leg', 'torch_eager', '--config', name, '--budget', '6'], 240)
        line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=max(args.warmup, 3),
                    ms_per_step=total_ms / args.steps, higher_is_better=True, scaling='weak', vs_baseline=None,
                    dtype='f32', data='synthetic',
                    config=dict(workload=workload_string(name),
                                step='one volume pass = %d optimiser steps (sampler + GT + synthesis + fwd + loss + bwd '
                                     '+ adjoint%s + Adam), CUDA-graph replay, %d launches per optimiser step%s' % (
                                         steps_per_pass, ' + gradient sum over ranks' if world > 1 else '', launches,
                                         ' (per-sample kernel + lfgc_grid_step)' if gstep else ''),
                                l2='flushed between timed steps (256 MiB write outside the event pairs)',
                                parallelism='dp%d' % world, gradient_sum=allreduce,
                                scaling_note='weak: the per-GPU batch is fixed, the global batch grows with N (a different '
                                             'optimisation problem; train_volume(args, world=N) is the strong-scaling '
                                             'entry point that keeps the reference\'s global batch)',
                                switches={k: v for k, v in sorted(os.environ.items()) if k.startswith('LFGC_')}),
                    e2e=e2e, gpu_launches=int(launches * steps_per_pass * args.steps),
                    clocks=dict(sm_mhz=clk['sm_mhz'], sm_max_mhz=clk['sm_max_mhz'], reasons=clk['reasons'],
                                samples=clk['samples']),
                    roofline=roofline, roofline_hbm=roofline_hbm, roofline_tensor=roofline_tensor, reconstruct=recon,
                    cpu_baseline=cpu, torch_eager_b200=eager, e2e_module_api=e2e_module,
                    extra=dict(us_per_optimiser_step=1e3 * total_ms / (args.steps * steps_per_pass),
                               hot_l2_samples_per_s=steps_per_pass * n * world / (hot_ms * 1e-3),
                               wall_s_timed_region=wall, final_mse=final_loss,
                               launches_per_optimiser_step=launches, strong_scaling=strong, configs=extra_cfgs,
                               psnr=psnr))
        print(json.dumps(line), flush=True)
    _shutdown(world, dist)


def _shutdown(world, dist):
    """Leave promptly once the JSON line is out.  Observed on a 2-GPU box: with collectives captured in CUDA graphs the
    workers can sit in the process-group / interpreter teardown indefinitely after the result was printed (the launcher
    then waits for them).  So: tear down cooperatively under a watchdog and hard-exit with status 0 -- nothing after the
    printed line carries information."""
    import gc
    import threading
    sys.stdout.flush()
    sys.stderr.flush()
    if world <= 1:
        return
    threading.Timer(20.0, lambda: os._exit(0)).start()      # watchdog: never outlive the result by more than 20 s
    try:
        torch.cuda.synchronize()
        gc.collect()
        torch.cuda.synchronize()
        dist.barrier()
        dist.destroy_process_group()
    except Exception:
        pass
    sys.stdout.flush()
    sys.stderr.flush()
    os._exit(0)


def reconstruct_rate(model, volume, rank, world, dev, fp32_peak, fwd_flops, reps=10):
    """decode voxels/s: field_from_net over this rank's slab of the volume (grid decoded once, one launch);
    whole-job rate = sum of the slabs / max time over ranks.  `value`: output left on the device; `e2e`: the reference's
    contract (OutputToVTK.py:42 returns a CPU volume) -- the slab copied to pinned host memory inside the timed region."""
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
    ms = float(t.item())
    # host-buffer variant
    host = torch.empty((slab[1] - slab[0], R, R), dtype=torch.float32).pin_memory()
    field_from_net(ds, model, True, slab=slab, host_out=host)
    if world > 1:
        dist.barrier()
    reps_h = max(2, reps // 2)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(reps_h):
        field_from_net(ds, model, True, slab=slab, host_out=host)     # synchronises: the host owns the data on return
    e1.record()
    torch.cuda.synchronize()
    wall = 1e3 * (time.perf_counter() - t0) / reps_h
    t = torch.tensor([max(e0.elapsed_time(e1) / reps_h, wall)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_h = float(t.item())
    del host
    model.train()
    vox = R ** 3
    tf = float(fwd_flops) * (slab[1] - slab[0]) * R * R / (ms * 1e-3) / 1e12
    return dict(metric='decode_voxels_per_s', value=vox / (ms * 1e-3), unit='voxels/s', ms_per_volume=ms,
                volume='%d^3, z-slab sharded over %d GPU(s), no communication' % (R, world),
                includes='axis tables + mask multipliers + wavelet synthesis + fused sample kernel, output stays in HBM',
                e2e=dict(value=vox / (ms_h * 1e-3), unit='voxels/s', ms_per_volume=ms_h, h2d_bytes_per_step=0,
                         d2h_bytes_per_step=4 * (slab[1] - slab[0]) * R * R,
                         api='field_from_net(dataset, net, slab=..., host_out=<pinned tensor>): the reconstructed slab '
                             'lands in HOST memory inside the timed region, max(CUDA events, wall clock)'),
                roofline=dict(bound='fp32', achieved=tf, peak=fp32_peak, unit='TFLOP/s', frac=tf / fp32_peak,
                              note='per GPU; %d FLOP/voxel; HBM-algorithmic 4 B/voxel written' % fwd_flops))


def e2e_host_fed(name, volume, n, rank, world, dev, steps=300, warmup=20):
    """samples/s of whole optimiser steps fed from HOST memory through the public trainer API
    (FastTrainer.step_host): every step copies that step's positions (n x 3 fp32) and target values (n fp32) from
    pinned host buffers, replays the captured step (synthesis + forward + MSE + backward + adjoint [+ gradient sum] +
    Adam) and reads the loss back to the host."""
    import torch.distributed as dist
    from latent_feature_grid_compression_b200 import ops
    from latent_feature_grid_compression_b200.training.fast_loop import make_trainer
    cfg = CONFIGS[name]
    a = cfg['args']
    m = build_model(name, dev)
    tr = make_trainer(m, volume, cfg['R'] ** 3, dict(a, batch_size=a['batch_size'] * world), a['lr'], seed=7, rank=rank,
                      world=world)
    n_buf = 8
    host = []
    for b in range(n_buf):
        raw, norm, gt = ops.sample(volume.shape, n, seed=99 + rank, sample_offset=b * n, volume=volume, want_gt=True)
        buf = torch.empty(4 * n).pin_memory()          # [positions | targets] back to back: one H2D copy per step
        buf[:3 * n].view(n, 3).copy_(norm.cpu())
        buf[3 * n:].copy_(gt.cpu())
        host.append((buf[:3 * n].view(n, 3), buf[3 * n:]))
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
    tr._graphs.clear()
    return dict(value=steps * n * world / (ms * 1e-3), unit=UNIT, h2d_bytes_per_step=n * 16, d2h_bytes_per_step=4,
                api='FastTrainer.step_host_pipelined(coords, targets): one optimiser step of %d host-resident samples '
                    'per GPU per call (H2D on a copy stream into double-buffered staging, fused step with '
                    'caller-supplied samples, CUDA-graph replay, every step\'s loss copied to pinned host memory and read '
                    'one call later); timed = max(CUDA events, host wall clock)' % n,
                us_per_optimiser_step=1e3 * ms / steps, final_mse=last,
                serial=dict(value=steps * n * world / (ms_serial * 1e-3), us_per_optimiser_step=1e3 * ms_serial / steps,
                            api='FastTrainer.step_host + last_loss(): copies, replay and loss read serialised per step'))


def e2e_module_path(name, volume, n, rank, world, dev, steps=100, warmup=10):
    """samples/s through model(coords) / loss.backward() / Adam with pinned HOST inputs every step (coords + ground
    truth), one device->host read of the loss per step; data parallel ranks all-reduce the flattened gradients.
    This is the path the unchanged training/training.py takes (MSE objective)."""
    import torch.distributed as dist
    from latent_feature_grid_compression_b200 import ops
    a = CONFIGS[name]['args']
    m = build_model(name, dev)
    opt = torch.optim.Adam(m.parameters(), lr=a['lr'])
    crit = torch.nn.MSELoss()
    n_buf = 8
    host = []
    for b in range(n_buf):
        raw, norm, gt = ops.sample(volume.shape, n, seed=99 + rank, sample_offset=b * n, volume=volume, want_gt=True)
        buf = torch.empty(4 * n).pin_memory()          # [positions | targets] back to back: one H2D copy per step
        buf[:3 * n].view(n, 3).copy_(norm.cpu())
        buf[3 * n:].copy_(gt.cpu())
        host.append((buf[:3 * n].view(n, 3), buf[3 * n:]))
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
            flat = torch.cat([p.grad.reshape(-1) for p in params if p.grad is not None])
            dist.all_reduce(flat)
            flat /= world
            off = 0
            for p in params:
                if p.grad is None:
                    continue
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
                api='Feature_Grid_Model(coords) + MSELoss + backward + torch.optim.Adam, one optimiser step of %d '
                    'samples per GPU per call' % n, us_per_optimiser_step=1e3 * ms / steps)


def psnr_section(dev, seeds=tuple(range(8))):
    """Final PSNR of the FAST loop (train_volume: the reference's full two-phase schedule on the graph-captured step,
    Philox sample stream) per BASELINE config and seed, next to the reference's own runs on the same synthetic volume
    (tests/golden/psnr_configs.json: unmodified training/training.py:184, CPU, torch seeds 0-2).  The sample streams
    differ by construction, so the comparison is distribution against distribution: delta of the means, both standard
    deviations and the standard error of the delta (the reference spreads 0.2-1.3 dB between seeds on its own, SURVEY
    7.2; one and the same fast-loop seed spreads 0.2 dB between runs because the scatter atomics reorder the sums:
    profiles/r2_variational_psnr_distribution.txt).  Eight fast-loop seeds against the reference's three; the 60-pass
    record runs three."""
    from latent_feature_grid_compression_b200.training.fast_loop import train_volume
    path = os.path.join(ROOT, 'tests', 'golden', 'psnr_configs.json')
    if not os.path.exists(path):
        return None
    recs = json.load(open(path))
    groups = {}
    for r in recs:
        groups.setdefault((r['config'], r['max_pass']), []).append(r)
    out = {}
    vols = {}
    for (cname, max_pass), rs in sorted(groups.items()):
        a = dict(rs[0]['args'])
        R = int(rs[0]['volume'].split('(')[1].rstrip(')'))
        if R not in vols:
            vols.clear()
            torch.cuda.empty_cache()
            vols[R] = synthetic_volume(R, dev).cpu()
        mine, zeros, t_s = [], [], []
        for s in (seeds if max_pass <= 50 else seeds[:3]):
            torch.manual_seed(s)
            t0 = time.perf_counter()
            try:
                info = train_volume(dict(a, max_pass=max_pass), volume=vols[R], seed=1000 + s)
            except Exception as e:   # noqa: BLE001
                out['%s@%d' % (cname, max_pass)] = dict(error=repr(e)[:300])
                mine = None
                break
            t_s.append(time.perf_counter() - t0)
            mine.append(float(info['psnr']))
            zeros.append(float(info['num_zeros']))
        if not mine:
            continue
        ref = [float(r['psnr']) for r in rs]
        out['%s@%d' % (cname, max_pass)] = dict(
            config=cname, max_pass=max_pass, config_max_pass=rs[0]['config_max_pass'],
            fast_loop_psnr_db=mine, reference_psnr_db=ref, psnr_delta_db=float(np.mean(mine) - np.mean(ref)),
            fast_loop_spread_db=float(max(mine) - min(mine)), reference_spread_db=float(max(ref) - min(ref)),
            fast_loop_std_db=float(np.std(mine, ddof=1)) if len(mine) > 1 else None,
            reference_std_db=float(np.std(ref, ddof=1)) if len(ref) > 1 else None,
            psnr_delta_stderr_db=float(np.sqrt(np.var(mine, ddof=1) / len(mine) + np.var(ref, ddof=1) / len(ref)))
            if len(mine) > 1 and len(ref) > 1 else None,
            fast_loop_num_zeros=zeros, reference_num_zeros=[float(r['num_zeros']) for r in rs],
            fast_loop_seconds_per_run=float(np.mean(t_s)),
            reference_cpu_seconds_per_run=float(np.mean([r['cpu_seconds'] for r in rs])))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', type=str, default='native', choices=['native', 'reference'])
    ap.add_argument('--config', type=str, default='mhd_p_basic', choices=sorted(CONFIGS))
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-extras', action='store_true', help='skip the other BASELINE configs and the PSNR runs')
    ap.add_argument('--leg', type=str, default='', choices=['', 'cpu_baseline', 'torch_eager'], help=argparse.SUPPRESS)
    ap.add_argument('--budget', type=float, default=15.0, help=argparse.SUPPRESS)
    args = ap.parse_args()
    if args.leg:
        run_cpu_leg(args)
        sys.stdout.flush()
        os._exit(0)
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_native(args)


if __name__ == '__main__':
    main()
