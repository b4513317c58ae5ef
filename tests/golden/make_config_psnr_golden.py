"""PSNR goldens AT THE BASELINE CONFIGS: run the UNMODIFIED reference driver (training/training.py:184 ``training``)
with the reference's own experiment-config files on the seeded synthetic volumes bench.py uses (the real datasets
are absent, .MISSING_LARGE_BLOBS) and record final PSNR / pruned-coefficient count / compression ratio per seed.

    python tests/golden/make_config_psnr_golden.py <config> <seed> [max_pass]      # build container only (CPU)
        config: turbulence_basic | test_impl_test | mhd_p_basic | mhd_p_smallify | mhd_p_dynamic_variational

Results are appended to tests/golden/psnr_configs.json (one record per (config, seed, max_pass)); the GPU tests and
bench.py read that file only.  Run-time patches (no source edits), as in make_psnr_golden.py: the torch>=2 eval
``view`` fix, ``is_cuda=False`` for the reconstruction (training/training.py:20 hard-codes True), and the binary store
skipped (storage is covered by tests/test_storage_compat.py).  ``max_pass`` overrides the config's value when the
full schedule is too slow for a CPU run; the record says so.
"""
import json
import os
import shutil
import sys
import tempfile
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import ref_harness  # noqa: E402

ref_harness.install()
import Feature_Grid_Training as FGT  # noqa: E402
import model.Feature_Grid_Model as FGM  # noqa: E402
import training.training as T  # noqa: E402
import visualization.OutputToVTK as V  # noqa: E402

OUT = os.path.join(HERE, 'psnr_configs.json')
VOLUME_EDGE = {'turbulence_basic': 150, 'test_impl_test': 150, 'mhd_p_basic': 255, 'mhd_p_smallify': 255,
               'mhd_p_dynamic_variational': 255}


def read_config(name):
    """key = value file -> args dict with the parser's defaults and types (what ConfigArgParse would produce)."""
    parser = FGT.config_parser()
    types = {a.dest: a for a in parser._actions}
    args = {a.dest: a.default for a in parser._actions if a.dest not in ('help', 'config')}
    path = os.path.join(ref_harness.REFERENCE_ROOT, 'experiment-config-files', name + '.txt')
    for line in open(path):
        line = line.strip()
        if not line or line.startswith('#') or '=' not in line:
            continue
        k, v = [s.strip() for s in line.split('=', 1)]
        if v in ("''", '""'):
            v = ''
        act = types[k]
        args[k] = act.type(v) if (act.type is not None and v != '') else v
    return args


def run(config, seed, max_pass=None, threads=4, workers=4):
    from bench import synthetic_volume
    torch.set_num_threads(threads)
    args = read_config(config)
    full_pass = args['max_pass']
    if max_pass is not None:
        args['max_pass'] = int(max_pass)
    args['num_workers'] = workers
    R = VOLUME_EDGE[config]
    work = tempfile.mkdtemp()
    cwd = os.getcwd()
    os.chdir(work)
    os.makedirs('datasets')
    np.save('datasets/vol.npy', synthetic_volume(R, 'cpu').numpy())
    args['data'] = 'datasets/vol.npy'
    args['basedir'] = '/experiments/'
    orig_tiled, orig_fwd, orig_store = V.tiled_net_out, FGM.Feature_Grid_Model.forward, T.store_model_parameters

    def fwd(self, t):
        return ref_harness.patched_eval_forward(self, t) if not self.training else orig_fwd(self, t)
    FGM.Feature_Grid_Model.forward = fwd
    T.tiled_net_out = lambda ds, m, is_cuda, **kw: orig_tiled(ds, m, False, **{**kw, 'write_vols': False})
    T.store_model_parameters = lambda *a, **k: None
    t0 = time.time()
    n_steps = [0]
    orig_step = torch.optim.Adam.step

    def counting_step(self, *a, **k):
        n_steps[0] += 1
        return orig_step(self, *a, **k)
    torch.optim.Adam.step = counting_step
    try:
        torch.manual_seed(seed)
        info = T.training(dict(args), verbose=False)
    finally:
        torch.optim.Adam.step = orig_step
        FGM.Feature_Grid_Model.forward, T.tiled_net_out, T.store_model_parameters = orig_fwd, orig_tiled, orig_store
        os.chdir(cwd)
        shutil.rmtree(work, ignore_errors=True)
    rec = dict(config=config, seed=int(seed), max_pass=int(args['max_pass']), config_max_pass=int(full_pass),
               volume='bench.synthetic_volume(%d)' % R, psnr=float(info['psnr']), mse=float(info['mse']),
               rmse=float(info['rmse']), num_zeros=float(info['num_zeros']),
               num_parameters=int(info['num_parameters']), compression_ratio=float(info['compression_ratio']),
               cpu_seconds=time.time() - t0, optimiser_steps=n_steps[0], torch_threads=threads, num_workers=workers,
               args={k: v for k, v in args.items() if k not in ('data', 'basedir')})
    recs = json.load(open(OUT)) if os.path.exists(OUT) else []
    recs = [r for r in recs if (r['config'], r['seed'], r['max_pass']) != (rec['config'], rec['seed'], rec['max_pass'])]
    recs.append(rec)
    recs.sort(key=lambda r: (r['config'], r['max_pass'], r['seed']))
    with open(OUT + '.tmp', 'w') as f:
        json.dump(recs, f, indent=1)
    os.replace(OUT + '.tmp', OUT)
    print(json.dumps({k: v for k, v in rec.items() if k != 'args'}))


if __name__ == '__main__':
    run(sys.argv[1], int(sys.argv[2]), int(sys.argv[3]) if len(sys.argv) > 3 else None)
