"""B200-native (sm_100a) implementation of the fV-SRN latent-feature-grid hot path of
Bussler/Latent_Feature_Grid_Compression, behind the reference's own nn.Module API.

Sub-packages mirror the reference's module names (``model``, ``data``, ``wavelet_transform``, ``visualization``) so
they can stand in for them; ``ops`` holds the tensor-level wrappers over the C ABI of
``liblfgc.so`` (include/lfgc.h) and ``training.fast_loop`` the graph-captured data-parallel trainer.
"""
from . import _lib  # noqa: F401

__all__ = ['ops', 'model', 'data', 'wavelet_transform', 'visualization', 'training']
__version__ = '0.1.0'
