"""Build recipe for liblfgc.so (hand-written sm_100a CUDA behind the C ABI of include/lfgc.h).

``python -m latent_feature_grid_compression_b200.build`` or ``build_library()``.  nvcc cross-compiles without a
GPU; the .so is written IN-TREE next to this file so it travels with the repository snapshot.
"""
from __future__ import annotations

import glob
import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
REPO_ROOT = os.path.dirname(PKG_DIR)
CSRC = os.path.join(PKG_DIR, 'csrc')
INCLUDE = os.path.join(REPO_ROOT, 'include')
LIB_PATH = os.path.join(PKG_DIR, 'liblfgc.so')

NVCC_FLAGS = [
    '-shared', '-Xcompiler', '-fPIC', '-std=c++17', '-O3', '-lineinfo', '-t', '8',
    '-gencode', 'arch=compute_100a,code=sm_100a',
]


def _nvcc():
    exe = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    if not os.path.exists(exe):
        raise RuntimeError('nvcc not found; liblfgc.so cannot be built')
    return exe


def sources():
    return sorted(glob.glob(os.path.join(CSRC, '*.cu')))


def is_stale():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = sources() + glob.glob(os.path.join(CSRC, '*.cuh')) + glob.glob(os.path.join(INCLUDE, '*.h'))
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force=False, verbose=False):
    """Compile every .cu under csrc/ into liblfgc.so for sm_100a.  Returns the library path."""
    if not force and not is_stale():
        return LIB_PATH
    cmd = [_nvcc()] + NVCC_FLAGS + ['-I', INCLUDE, '-I', CSRC] + sources() + ['-o', LIB_PATH + '.tmp']
    if verbose:
        cmd.insert(1, '-Xptxas')
        cmd.insert(2, '-v')
        print(' '.join(cmd))
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError('nvcc failed building liblfgc.so')
    if verbose:
        print(res.stdout + res.stderr)
    os.replace(LIB_PATH + '.tmp', LIB_PATH)
    return LIB_PATH


if __name__ == '__main__':
    print(build_library(force='--force' in sys.argv, verbose='-v' in sys.argv))
