"""Where does the fast loop's final PSNR on a variational config come from?  Runs the two-phase schedule of one golden
record (tests/golden/psnr_configs.json) under switches, one subprocess each, and prints the final PSNR / zero count:

    default      the shipped fast loop
    nograph      the same step body launched eagerly (no CUDA graph)
    novarfast    per-layer mask launches instead of the flat variational kernels
    simt         LFGC_BACKWARD_TC=0 LFGC_FORWARD_TC=0 (SIMT kernels instead of tcgen05)
    hostfed      samples drawn like data/IndexDataset.py:90-96 with torch.randint and fed through step_host

    python profiles/variational_diag.py mhd_p_dynamic_variational 0 12 [variant ...]
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
VARIANTS = ['default', 'nograph', 'novarfast', 'simt', 'hostfed']


def child(cfg, seed, max_pass, variant):
    sys.path.insert(0, ROOT)
    import torch
    import bench
    from latent_feature_grid_compression_b200.training import fast_loop as F
    rec = [r for r in json.load(open(os.path.join(ROOT, 'tests', 'golden', 'psnr_configs.json')))
           if r['config'] == cfg and r['seed'] == seed and r['max_pass'] == max_pass][0]
    R = int(rec['volume'].split('(')[1].rstrip(')'))
    vol = bench.synthetic_volume(R, 'cuda').cpu()
    init = F.FastTrainer.__init__
    if variant == 'nograph':
        def patched(self, *a, **k):
            k['use_graph'] = False
            init(self, *a, **k)
        F.FastTrainer.__init__ = patched
    if variant == 'novarfast':
        def patched(self, *a, **k):
            init(self, *a, **k)
            self._var_fast = False
        F.FastTrainer.__init__ = patched
    if variant == 'hostfed':
        def solve_phase(model, volume, n_voxels, args, max_pass, lr, sched=None, bind=True, seed=0, rank=0, world=1,
                        group=None, regularise=True, verbose=False):
            trainer = F.make_trainer(model, volume, n_voxels, args, lr, seed=seed, regularise=regularise,
                                     variational_sched_check=sched)
            if sched is not None:
                sched.bind(trainer if bind else None)
            res = torch.tensor(volume.shape, device='cuda')
            mx = (res - 1).float()
            scales = mx / mx.max()
            flat = volume.reshape(-1)
            for step, prior, passes, last in F.epoch_schedule(n_voxels, int(args['batch_size']),
                                                              int(args['sample_size']), max_pass):
                lin = torch.randint(0, n_voxels, (trainer.batch,), device='cuda')
                z = lin % res[2]
                y = (lin // res[2]) % res[1]
                x = lin // (res[2] * res[1])
                idx = torch.stack([x, y, z], 1).float()
                coords = scales[None] * (2.0 * idx / mx[None] - 1.0)
                trainer.step_host(coords.contiguous(), flat[lin].contiguous())
                if sched is not None and sched.update(prior, passes, trainer.complete_loss):
                    break
            torch.cuda.synchronize()
            return trainer, False
        F.solve_phase = solve_phase
    torch.manual_seed(seed)
    info = F.train_volume(dict(rec['args']), volume=vol, seed=1000 + seed)
    print('RESULT %s: fast loop %.3f dB / zeros %.0f / steps %d   reference %.3f dB / zeros %.0f / steps %d' % (
        variant, info['psnr'], info['num_zeros'], info['steps'], rec['psnr'], rec['num_zeros'], rec['optimiser_steps']))


if __name__ == '__main__':
    if sys.argv[1] == '--child':
        child(sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), sys.argv[5])
        sys.exit(0)
    cfg, seed, max_pass = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    for v in (sys.argv[4:] or VARIANTS):
        env = dict(os.environ)
        if v == 'simt':
            env.update(LFGC_BACKWARD_TC='0', LFGC_FORWARD_TC='0')
        out = subprocess.run([sys.executable, os.path.abspath(__file__), '--child', cfg, str(seed), str(max_pass), v],
                             env=env, capture_output=True, text=True)
        lines = [l for l in out.stdout.splitlines() if l.startswith('RESULT')]
        print(lines[0] if lines else '%s FAILED\n%s' % (v, out.stderr[-2000:]), flush=True)
