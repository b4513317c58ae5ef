"""Build a variant of liblfgc.so with extra nvcc flags: python profiles/build_variant.py OUT.so [-DLFGC_PHASE_TIMING ...];
select it with LFGC_LIB=OUT.so."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import subprocess
from latent_feature_grid_compression_b200 import build as B
out = sys.argv[1]; defs = sys.argv[2:]
cmd = [B._nvcc()] + B.NVCC_FLAGS + defs + ['-I', B.INCLUDE, '-I', B.CSRC] + B.sources() + ['-o', out]
r = subprocess.run(cmd, capture_output=True, text=True); print(r.stdout[-2000:], r.stderr[-2000:]); sys.exit(r.returncode)
