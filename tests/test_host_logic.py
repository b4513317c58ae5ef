"""CPU-side checks: the C-ABI library loads and exports every symbol include/lfgc.h declares, host-side descriptors,
wavelet filters, config reader, on-disk format helpers (no GPU, no compute calls)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, 'include', 'lfgc.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(lfgc_[a-z0-9_]+)\s*\(', text)))


def test_library_exports_every_declared_symbol():
    from latent_feature_grid_compression_b200 import _lib
    from latent_feature_grid_compression_b200.build import LIB_PATH, build_library
    build_library()
    lib = ctypes.CDLL(LIB_PATH)
    declared = _declared_symbols()
    assert len(declared) >= 20
    for sym in declared:
        assert hasattr(lib, sym), sym
    assert set(declared) == set(_lib.EXPORTED_SYMBOLS)
    loaded = _lib.load()
    assert loaded.lfgc_abi_version() == _lib.ABI_VERSION == 6
    # the ctypes mirrors of the ABI structs have the library's sizes (checked again at every load)
    import ctypes as _ct
    sizes = (_ct.c_size_t * 4)()
    assert loaded.lfgc_struct_sizes(sizes, 4) == 4
    assert [int(v) for v in sizes] == [_ct.sizeof(_lib.WaveletDesc), _ct.sizeof(_lib.ModelDesc),
                                       _ct.sizeof(_lib.PeerAnnounce), _ct.sizeof(_lib.GridStepArgs)]
    assert loaded.lfgc_last_error() is not None


def test_argument_validation_without_a_gpu():
    """Bad arguments are rejected on the host side before any launch (error code + message, no exception)."""
    from latent_feature_grid_compression_b200 import _lib
    lib = _lib.load()
    rc = lib.lfgc_mask_multiplier(99, 4, None, None, None, 0.0, None, None, None)
    assert rc == -1 and b'bad arguments' in lib.lfgc_last_error()
    rc = lib.lfgc_smallify_ema(None, None, None, 4, 0.1, None)
    assert rc == -1
    with pytest.raises(_lib.LfgcError):
        _lib.check(rc, 'lfgc_smallify_ema')


def test_geometry_descriptors():
    from latent_feature_grid_compression_b200 import ops
    g = ops.Geometry(16, (15, 15, 15), 32, 4, 2, 'db2', [(6, 6, 6), (6, 6, 6), (9, 9, 9)], [[9, 9, 9], [15, 15, 15]])
    assert g.Cp == 16 and g.in0 == 31 and g.mlp_param_count == 4225
    assert [s for _, s in g.mlp_shapes()][:2] == [(32, 31), (32,)]
    assert g.coeff_shape(0) == (16, 6, 6, 6) and g.coeff_shape(2) == (16, 7, 9, 9, 9)
    assert g.backward_workspace_bytes >= 148 * 4226 * 4
    g6 = ops.Geometry(6, (15, 15, 15), 20, 3, 1, 'haar', [(15, 15, 15)], np.zeros((0, 3)))
    assert g6.Cp == 8 and g6.in0 == 15 and g6.mlp_param_count == 15 * 20 + 20 + 2 * (400 + 20) + 21
    with pytest.raises(ValueError):
        ops.Geometry(4, (8, 8, 8), 32, 4, 2, 'db2', [(4, 4, 4)], [[8, 8, 8]])


def test_wavelet_filters_and_levels():
    from latent_feature_grid_compression_b200 import wavelets
    from oracle import fvsrn_numpy as O
    for name in ('haar', 'db2'):
        for a, b in zip(wavelets.filter_bank(name), O.wavelet_taps(name)):
            assert np.allclose(np.asarray(a, dtype=np.float32), b, atol=0, rtol=0)
    # published db3/db4 leading taps (Daubechies 1992, table 6.1)
    assert np.allclose(wavelets.filter_bank('db3')[2][:3], [0.3326705529500825, 0.8068915093110924, 0.4598775021184914], atol=1e-9)
    assert np.allclose(wavelets.filter_bank('db4')[2][:3], [0.2303778133088964, 0.7148465705529154, 0.6308807679298587], atol=1e-9)
    for name in ('haar', 'db2', 'db3', 'db4', 'db6', 'db8'):
        dec_lo, dec_hi, rec_lo, rec_hi = (np.asarray(t) for t in wavelets.filter_bank(name))
        n = len(rec_lo)
        assert abs(rec_lo.sum() - np.sqrt(2)) < 1e-9 and abs(rec_hi.sum()) < 1e-9
        for k in range(n // 2):  # orthonormal shifts
            assert abs(np.dot(rec_lo[:n - 2 * k], rec_lo[2 * k:]) - (1.0 if k == 0 else 0.0)) < 1e-9
            assert abs(np.dot(rec_lo[:n - 2 * k], rec_hi[2 * k:])) < 1e-9
    for n, flen in ((15, 4), (16, 2), (5, 4), (2, 4), (64, 4), (17, 4)):
        assert wavelets.dwt_max_level(n, flen) == O.dwt_max_level(n, flen)
    assert wavelets.analysis_out_size(15, 4) == 9 and wavelets.analysis_out_size(9, 4) == 6


def test_dict_from_file_reads_reference_configs(tmp_path):
    from latent_feature_grid_compression_b200.visualization.pltUtils import dict_from_file
    p = tmp_path / 'cfg.txt'
    p.write_text('expname = basic\ndata = datasets/mhd1024.h5\nd_in = 3\nlr = 0.008\ndrop_type =\n'
                 'pruning_threshold_list = [0.9, 0.8]\ncheckpoint_path = \'\'\nflag = True\nweight_dkl_multiplier = 5e-05\n')
    d = dict_from_file(str(p))
    assert d['expname'] == 'basic' and d['d_in'] == 3 and d['lr'] == 0.008 and d['drop_type'] == ''
    assert d['pruning_threshold_list'] == [0.9, 0.8] and d['flag'] is True and d['weight_dkl_multiplier'] == 5e-05


def test_bit_packing_helpers_match_the_reference_layout(tmp_path):
    from latent_feature_grid_compression_b200.model import model_utils as mu
    packed, leftover = mu.ints_to_bits_to_bytes([5, 255, 0, 128], 8)
    assert bytes(packed) == bytes([5, 255, 0, 128]) and leftover is False
    packed, leftover = mu.ints_to_bits_to_bytes([5, 3], 3)          # '101' '011' -> one partial byte 0b101011
    assert bytes(packed) == bytes([0b101011]) and leftover is True
    packed, leftover = mu.ints_to_bits_to_bytes([300, 17, 511], 9)  # 27 bits: 3 full bytes + '111'
    bits = format(300, '09b') + format(17, '09b') + format(511, '09b')
    assert bytes(packed) == bytes([int(bits[0:8], 2), int(bits[8:16], 2), int(bits[16:24], 2), int(bits[24:], 2)])
    f = str(tmp_path / 'mask.bnr')
    mu.binary_writing('1011001110', f)
    assert open(f, 'rb').read() == bytes([0b10110011, 0b10000000])
    assert mu.read_binary(f, 10)[:10] == '1011001110'


def test_flat_pack_keeps_parameters_as_views():
    from latent_feature_grid_compression_b200.ops import FlatPack
    ps = [torch.nn.Parameter(torch.randn(3, 4)), torch.nn.Parameter(torch.randn(5))]
    before = [p.detach().clone() for p in ps]
    pack = FlatPack()
    flat = pack.ensure(ps)
    assert flat.numel() == 17 and all(torch.equal(a, b) for a, b in zip(before, ps))
    assert ps[0].data_ptr() == flat.data_ptr() and ps[1].data_ptr() == flat.data_ptr() + 48
    assert pack.ensure(ps) is flat                       # still views: no re-pack
    with torch.no_grad():
        ps[0].add_(1.0)
    assert torch.equal(flat[:12].view(3, 4), ps[0])      # in-place optimiser updates land in the flat buffer
    ps[1].data = torch.zeros(5)                          # restore_model-style rebinding breaks the view ...
    flat2 = pack.ensure(ps)
    assert flat2 is not flat and torch.equal(flat2[12:], torch.zeros(5))   # ... and is detected and re-packed


def test_model_refuses_cpu_tensors():
    """No CPU fallback: without CUDA the model cannot even be built, and CPU inputs are refused."""
    from latent_feature_grid_compression_b200._lib import LfgcError
    from latent_feature_grid_compression_b200.data.Interpolation import trilinear_f_interpolation
    if not torch.cuda.is_available():
        from latent_feature_grid_compression_b200.model.model_utils import setup_model
        with pytest.raises(LfgcError):
            setup_model(3, 32, 1, 4, 'fourier', 2, '', 0.1, 0.9, 'db2', 4, 15, '')
    with pytest.raises(LfgcError):
        trilinear_f_interpolation(torch.zeros(4, 3), torch.zeros(3, 3, 3), torch.zeros(3), torch.ones(3) * 2,
                                  torch.tensor([3.0, 3.0, 3.0]))


def test_dropin_aliases_expose_the_reference_module_names():
    """with dropin/ first on sys.path the reference's import statements resolve to this package."""
    import subprocess
    import sys
    code = ("import model.model_utils as mu, data.IndexDataset as d, visualization.OutputToVTK as v\n"
            "from model.Smallify_Dropout import SmallifyLoss\n"
            "from model.Variational_Dropout_Layer import VariationalDropoutLoss, Variance_Model, VariationalDropout\n"
            "from data.Interpolation import trilinear_f_interpolation, finite_difference_trilinear_grad\n"
            "from data.IndexDataset import get_tensor, IndexDataset\n"
            "from visualization.OutputToVTK import tiled_net_out\n"
            "from visualization.pltUtils import dict_from_file\n"
            "from model.model_utils import write_dict, setup_model, store_model_parameters, restore_model\n"
            "from wavelet_transform.Torch_Wavelet_Transform import WaveletFilter3d, _WaveletFilterNd\n"
            "assert mu.setup_model.__module__.startswith('latent_feature_grid_compression_b200')\n"
            "print('ok')\n")
    env = dict(os.environ, PYTHONPATH=os.path.join(ROOT, 'dropin') + os.pathsep + ROOT)
    out = subprocess.run([sys.executable, '-c', code], cwd='/tmp', env=env, capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.strip() == 'ok', out.stderr


def test_ctypes_signatures_match_the_header():
    """Every prototype of include/lfgc.h against the ctypes signature the package binds it with: same arity, and the
    same class (pointer / 32-bit int / 64-bit int / size_t / float / double) for every argument and the result.  A
    float-vs-double or int-vs-int64 slip would otherwise only show up as garbage on the GPU."""
    import ctypes as C
    import re
    from latent_feature_grid_compression_b200 import _lib
    text = open(os.path.join(ROOT, 'include', 'lfgc.h')).read()
    text = re.sub(r'/\*.*?\*/', ' ', text, flags=re.S)
    protos = re.findall(r'\n\s*((?:const\s+)?[A-Za-z_][\w\s\*]*?)\b(lfgc_\w+)\s*\(([^;{]*?)\)\s*;', text)
    assert len(protos) >= 30

    def c_class(decl):
        decl = decl.strip()
        if decl == 'void':
            return None
        if '*' in decl or '[' in decl:
            return 'ptr'
        base = re.sub(r'\b(const|unsigned)\b', '', decl).split()
        t = base[0] if base else ''
        return {'int': 'i32', 'int32_t': 'i32', 'int64_t': 'i64', 'uint64_t': 'i64', 'long': 'i64', 'size_t': 'size',
                'float': 'f32', 'double': 'f64'}.get(t, 'other:' + t)

    def ct_class(t):
        if t is None:
            return None
        if t in (C.c_void_p, C.c_char_p) or hasattr(t, 'contents') or (isinstance(t, type) and issubclass(t, C._Pointer)):
            return 'ptr'
        return {C.c_int: 'i32', C.c_int32: 'i32', C.c_uint32: 'i32', C.c_int64: 'i64', C.c_uint64: 'i64',
                C.c_longlong: 'i64', C.c_size_t: 'size', C.c_float: 'f32', C.c_double: 'f64'}.get(t, 'other:%r' % t)

    seen = set()
    for ret, name, args in protos:
        assert name in _lib._SIGNATURES, name + ' is declared in lfgc.h but not bound'
        res, argtypes = _lib._SIGNATURES[name]
        seen.add(name)
        ret = ret.strip()
        want_ret = 'ptr' if '*' in ret else c_class('long' if ret.startswith('long long') else ret)
        # size_t and 64-bit ints are the same register class on LP64; keep them distinct for arguments only
        got_ret = ct_class(res)
        assert (want_ret, got_ret) in ((want_ret, want_ret), ('size', 'i64'), ('i64', 'size')), (name, ret, res)
        decls = [a for a in (s.strip() for s in args.split(',')) if a]
        want = [c_class(a) for a in decls if c_class(a) is not None]
        got = [ct_class(t) for t in argtypes]
        assert len(want) == len(got), (name, len(want), len(got))
        for i, (w, g) in enumerate(zip(want, got)):
            assert w == g or {w, g} == {'size', 'i64'}, '%s: argument %d is %s in the header, %s in ctypes' % (name, i, w, g)
    assert seen == set(_lib._SIGNATURES), set(_lib._SIGNATURES) - seen
