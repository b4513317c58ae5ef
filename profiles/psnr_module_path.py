"""Final PSNR of the UNCHANGED reference driver (baseline/_ref training/training.py:184) on top of the drop-in modules for
one BASELINE config / seed of tests/golden/psnr_configs.json -- same torch seed, same DataLoader sample stream as the
reference's own CPU run, so the two numbers are directly comparable (module path; the fast loop uses a Philox stream).

    python profiles/psnr_module_path.py mhd_p_dynamic_variational 0 12
"""
import json
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
cfg, seed, max_pass = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
recs = [r for r in json.load(open(os.path.join(ROOT, 'tests', 'golden', 'psnr_configs.json')))
        if r['config'] == cfg and r['seed'] == seed and r['max_pass'] == max_pass]
assert recs, 'no golden record'
rec = recs[0]
work = tempfile.mkdtemp()
for item in ('training',):
    shutil.copytree(os.path.join(ROOT, 'baseline', '_ref', item), os.path.join(work, item))
code = '''
import sys, json, numpy as np, torch
sys.path.insert(0, %r)
import bench
from training.training import training
import os
os.makedirs('datasets', exist_ok=True)
R = %d
np.save('datasets/vol.npy', bench.synthetic_volume(R, 'cpu').numpy())
args = json.loads(%r)
args.update(data='datasets/vol.npy', basedir='/experiments/', expname='run', Tensorboard_log_dir='', checkpoint_path='',
            binary_checkpoint_path='', pruning_threshold_list=None)
torch.manual_seed(%d)
info = training(args, verbose=False)
print('RESULT', json.dumps(dict(psnr=float(info['psnr']), num_zeros=float(info['num_zeros']))))
''' % (ROOT, int(rec['volume'].split('(')[1].rstrip(')')), json.dumps(rec['args']), seed)
env = dict(os.environ, PYTHONPATH=os.pathsep.join([os.path.join(ROOT, 'dropin'), ROOT, os.path.join(ROOT, 'tests', 'shims')]))
out = subprocess.run([sys.executable, '-c', code], cwd=work, env=env, capture_output=True, text=True)
line = [l for l in out.stdout.splitlines() if l.startswith('RESULT')]
if not line:
    print(out.stdout[-1500:], out.stderr[-3000:])
    sys.exit(1)
res = json.loads(line[0][7:])
print('%s seed %d max_pass %d: reference (CPU) %.3f dB / zeros %.0f   unchanged driver on the drop-in (B200) %.3f dB / zeros %.0f'
      % (cfg, seed, max_pass, rec['psnr'], rec['num_zeros'], res['psnr'], res['num_zeros']))
