"""ctypes binding of liblfgc.so (the C ABI declared in include/lfgc.h).

There is no fallback: if the library is missing or a call fails, an exception is raised.
"""
from __future__ import annotations

import ctypes as C
import os

from .build import LIB_PATH as _DEFAULT_LIB_PATH

LIB_PATH = os.environ.get('LFGC_LIB', _DEFAULT_LIB_PATH)  # debug builds only; the shipped path is the in-tree library

MAX_LEVELS = 12
MAX_TAPS = 16
MAX_LAYERS = 8
MAX_PEERS = 16

MASK_IDENTITY, MASK_DIRECT, MASK_VARIATIONAL, MASK_STE_SIGMOID, MASK_BERNOULLI = range(5)
F_CLAMP = 1
ABI_VERSION = 6   # include/lfgc.h LFGC_ABI_VERSION this binding was written against

_ERRORS = {-1: 'LFGC_E_INVALID', -2: 'LFGC_E_UNSUPPORTED', -3: 'LFGC_E_CUDA', -4: 'LFGC_E_WORKSPACE'}


class LfgcError(RuntimeError):
    pass


class ModelDesc(C.Structure):
    _fields_ = [('C', C.c_int32), ('Cp', C.c_int32), ('G', C.c_int32 * 3), ('H', C.c_int32), ('L', C.c_int32),
                ('F', C.c_int32)]


class WaveletDesc(C.Structure):
    _fields_ = [('n_coeff', C.c_int32), ('C', C.c_int32), ('n_taps', C.c_int32),
                ('rec_lo', C.c_float * MAX_TAPS), ('rec_hi', C.c_float * MAX_TAPS),
                ('dims', (C.c_int32 * 3) * MAX_LEVELS), ('target', (C.c_int32 * 3) * MAX_LEVELS)]


_f = C.c_void_p      # device pointers travel as integers
_i64 = C.c_int64


class GridStepArgs(C.Structure):
    """lfgc_grid_step_args (include/lfgc.h)."""
    _fields_ = [('n_srcs', C.c_int32), ('grad_grid', _f * MAX_PEERS), ('mlp_partials', _f * MAX_PEERS),
                ('nslices', C.c_int32), ('pstride', C.c_int32), ('pcount', C.c_int32), ('zero_grid', _f),
                ('grid_cl', _f), ('p', _f), ('g', _f), ('m', _f), ('v', _f), ('coeff_off', _i64 * MAX_LEVELS),
                ('mlp_off', _i64), ('loss_out', _f), ('lr', _f), ('step_count', _f), ('beta1', C.c_double),
                ('beta2', C.c_double), ('eps', C.c_double), ('grad_scale', C.c_double), ('weight_l2', C.c_double),
                ('scratch', _f), ('scratch_bytes', C.c_size_t), ('panel_model', C.POINTER(ModelDesc)),
                ('panel_image', _f)]


class PeerAnnounce(C.Structure):
    """lfgc_peer_announce (include/lfgc.h)."""
    _fields_ = [('n_peers', C.c_int32), ('rank', C.c_int32), ('flags', _f * MAX_PEERS), ('epoch', _f), ('ticket', _f)]


_SIGNATURES = {
    'lfgc_abi_version': (C.c_int, []),
    'lfgc_struct_sizes': (C.c_int, [C.POINTER(C.c_size_t), C.c_int]),
    'lfgc_last_error': (C.c_char_p, []),
    'lfgc_sm_count': (C.c_int, []),
    'lfgc_launch_count': (C.c_longlong, []),
    'lfgc_mask_multiplier': (C.c_int, [C.c_int, _i64, _f, _f, _f, C.c_float, _f, _f, _f]),
    'lfgc_mask_param_grad': (C.c_int, [C.c_int, _i64, _f, _f, _f, _f, _f, _f, C.c_int, _f]),
    'lfgc_smallify_ema': (C.c_int, [_f, _f, _f, _i64, C.c_float, _f]),
    'lfgc_dwt_level': (C.c_int, [_f, C.c_int, C.POINTER(C.c_int32), C.c_int, C.POINTER(C.c_float),
                                 C.POINTER(C.c_float), _f, C.POINTER(C.c_int32), _f]),
    'lfgc_decode_scratch_bytes': (C.c_size_t, [C.POINTER(WaveletDesc)]),
    'lfgc_decode_fwd': (C.c_int, [C.POINTER(WaveletDesc), C.POINTER(_f), C.POINTER(_f), _f, _f, C.c_int, _f, _f]),
    'lfgc_decode_bwd': (C.c_int, [C.POINTER(WaveletDesc), _f, C.c_int, C.POINTER(_f), C.POINTER(_f), _f,
                                  C.POINTER(_f), C.POINTER(_f), C.c_int, _f]),
    'lfgc_mlp_param_count': (_i64, [C.POINTER(ModelDesc)]),
    'lfgc_forward': (C.c_int, [C.POINTER(ModelDesc), _f, _i64, _f, _f, _f, C.c_int, _f]),
    'lfgc_backward_workspace_bytes': (C.c_size_t, [C.POINTER(ModelDesc)]),
    'lfgc_backward': (C.c_int, [C.POINTER(ModelDesc), _f, _i64, _f, _f, _f, _f, _f, _f, C.c_int, _f, C.c_size_t, _f]),
    'lfgc_train_step': (C.c_int, [C.POINTER(ModelDesc), _f, C.POINTER(C.c_int32), _i64, C.c_uint64, C.c_uint64, _f,
                                  C.c_uint64, _f, _f, _f, C.c_float, _f, _f, _f, _f, _f, C.c_int, _f, C.c_size_t, _f]),
    'lfgc_train_step_weighted': (C.c_int, [C.POINTER(ModelDesc), _f, C.POINTER(C.c_int32), _i64, C.c_uint64, C.c_uint64,
                                           _f, C.c_uint64, _f, _f, _f, C.c_float, _f, _f, _f, _f, _f, _f, _f, C.c_int, _f,
                                           C.c_size_t, _f]),
    'lfgc_plain_mlp_forward': (C.c_int, [C.c_int, C.c_int, _f, _i64, _f, _f, _f]),
    'lfgc_plain_mlp_workspace_bytes': (C.c_size_t, [C.c_int, C.c_int]),
    'lfgc_plain_mlp_backward': (C.c_int, [C.c_int, C.c_int, _f, _i64, _f, _f, _f, C.c_int, _f, C.c_size_t, _f]),
    'lfgc_sample': (C.c_int, [_f, C.POINTER(C.c_int32), _i64, C.c_uint64, C.c_uint64, _f, _f, _f, _f, _f]),
    'lfgc_sample_stream': (C.c_int, [_f, C.POINTER(C.c_int32), _i64, C.c_uint64, C.c_uint64, _f, C.c_uint64, _f, _f, _f,
                                     _f, _f]),
    'lfgc_trilinear': (C.c_int, [_f, _i64, _f, C.POINTER(C.c_int32), C.POINTER(C.c_float), C.POINTER(C.c_float), _f, _f]),
    'lfgc_reconstruct': (C.c_int, [C.POINTER(ModelDesc), _f, _f, C.POINTER(C.c_int32), _f, _f, _f, C.c_int32,
                                   C.c_int32, _f, C.c_int, _f]),
    'lfgc_deviation_stats': (C.c_int, [_f, _f, _i64, _f, _f]),
    'lfgc_adam': (C.c_int, [_f, _f, _f, _f, _i64, _f, _f, C.c_double, C.c_double, C.c_double, C.c_double, _f]),
    'lfgc_adam_reg': (C.c_int, [_f, _f, _f, _f, _i64, _f, _f, C.c_double, C.c_double, C.c_double, C.c_double, _i64, _i64,
                                C.c_double, _i64, _i64, C.c_double, _f, _f, C.c_float, C.c_int, _f]),
    'lfgc_peer_sum': (C.c_int, [C.POINTER(_f), C.POINTER(_f), C.c_int, C.c_int, _f, _f, _f, _i64, C.c_int, _f]),
    'lfgc_add_l2_grad': (C.c_int, [_f, _f, _i64, C.c_float, _f]),
    'lfgc_add_l1_grad': (C.c_int, [_f, _f, _i64, C.c_float, _f]),
    'lfgc_train_step_partials': (C.c_int, [C.POINTER(ModelDesc), _f, C.POINTER(C.c_int32), _i64, C.c_uint64, C.c_uint64, _f,
                                           C.c_uint64, _f, _f, _f, C.c_float, _f, _f, _f, _f, _f, C.c_size_t,
                                           C.POINTER(C.c_int32), _f]),
    'lfgc_tc_panel_bytes': (C.c_size_t, [C.POINTER(ModelDesc)]),
    'lfgc_tc_panel_build': (C.c_int, [C.POINTER(ModelDesc), _f, _f, _f]),
    'lfgc_train_step_accumulate': (C.c_int, [C.POINTER(ModelDesc), _f, C.POINTER(C.c_int32), _i64, C.c_uint64, C.c_uint64,
                                             _f, C.c_uint64, _f, _f, _f, C.c_float, _f, _f, _f, _f, C.c_int, C.POINTER(PeerAnnounce),
                                             _f, _f, C.c_size_t, _f]),
    'lfgc_grid_step': (C.c_int, [C.POINTER(WaveletDesc), C.c_int, C.POINTER(GridStepArgs), _f]),
    'lfgc_grid_step_smem_bytes': (C.c_size_t, [C.POINTER(WaveletDesc)]),
    'lfgc_grid_step_scratch_bytes': (C.c_size_t, [C.POINTER(WaveletDesc)]),
    'lfgc_grid_step_supported': (C.c_int, [C.POINTER(WaveletDesc)]),
    'lfgc_variational_multiplier': (C.c_int, [_f, _f, C.c_int, C.POINTER(C.c_int64), _f, _f, _f]),
    'lfgc_variational_param_grad': (C.c_int, [_f, _f, _f, C.c_int, C.POINTER(C.c_int64), _f, _f]),
    'lfgc_variational_dkl_grad': (C.c_int, [_f, _f, C.c_int, C.POINTER(C.c_int64), _f, _f, C.c_double, C.c_double,
                                            C.c_float, _f]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lib = None


def load():
    """Load liblfgc.so (once).  Raises if it has not been built -- there is no CPU path."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise LfgcError('liblfgc.so not found at %s: build it with '
                        '`python -m latent_feature_grid_compression_b200.build` (no CPU fallback exists)' % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.lfgc_abi_version() != ABI_VERSION:
        raise LfgcError('liblfgc.so ABI version mismatch; rebuild')
    sizes = (C.c_size_t * 4)()
    n = lib.lfgc_struct_sizes(sizes, 4)
    mine = [C.sizeof(WaveletDesc), C.sizeof(ModelDesc), C.sizeof(PeerAnnounce), C.sizeof(GridStepArgs)]
    if n != len(mine) or [int(v) for v in sizes] != mine:
        raise LfgcError('struct layouts of this binding %s differ from the library\'s %s (include/lfgc.h changed?)'
                        % (mine, [int(v) for v in sizes[:n]]))
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        msg = load().lfgc_last_error()
        raise LfgcError('%s failed: %s: %s' % (what, _ERRORS.get(rc, rc), msg.decode() if msg else ''))


def int3(v):
    return (C.c_int32 * 3)(int(v[0]), int(v[1]), int(v[2]))


def float3(v):
    return (C.c_float * 3)(float(v[0]), float(v[1]), float(v[2]))


def ptr_array(ptrs):
    """HOST array of device pointers (None -> NULL)."""
    arr = (_f * len(ptrs))()
    for i, p in enumerate(ptrs):
        arr[i] = p if p else None
    return arr
