"""Install the UNMODIFIED reference into baseline/_ref (git-ignored, travels to the GPU box with the snapshot).

The prescribed ``pip install --no-index --target baseline/_ref /root/reference`` fails ("Neither 'setup.py' nor
'pyproject.toml' found": the reference is a plain script tree, recorded in DESIGN.md), so the Python files of the hot
path and its drivers are copied verbatim: Feature_Grid_Training.py, Feature_Grid_Inference.py, training/, model/, data/,
wavelet_transform/, visualization/{OutputToVTK,pltUtils}.py and experiment-config-files/.  Nothing under baseline/_ref
is ever committed (.gitignore) or imported by the product package; users: bench.py --impl reference (timing the
reference on the host cores and, as an informative key, its own torch-eager path on the B200) and
tests/test_gpu_reference_drivers.py (the unchanged drivers on top of the drop-in modules).

    python baseline/install_ref.py [/root/reference]
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, '_ref')
ITEMS = ['Feature_Grid_Training.py', 'Feature_Grid_Inference.py', 'training', 'model', 'data', 'wavelet_transform',
         'visualization/OutputToVTK.py', 'visualization/pltUtils.py', 'experiment-config-files']


def install(src='/root/reference'):
    if not os.path.isdir(src):
        return None
    os.makedirs(DEST, exist_ok=True)
    for item in ITEMS:
        s, d = os.path.join(src, item), os.path.join(DEST, item)
        if os.path.isdir(s):
            shutil.copytree(s, d, dirs_exist_ok=True, ignore=shutil.ignore_patterns('__pycache__', '*.pyc'))
        else:
            os.makedirs(os.path.dirname(d), exist_ok=True)
            shutil.copy2(s, d)
    return DEST


if __name__ == '__main__':
    print(install(sys.argv[1] if len(sys.argv) > 1 else '/root/reference'))
