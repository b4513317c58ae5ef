"""Model factory and on-disk format -- counterpart of the reference's model/model_utils.py.

``setup_model`` keeps the reference's positional signature and the order in which it draws from the torch
generator (grid ``uniform_`` -> prototype mask -> per-level masks -> ``nn.Linear`` inits, model_utils.py:23-59), so
that the same ``torch.manual_seed`` gives the same initial parameters.

``store_model_parameters`` / ``restore_model`` read and write the reference's binary layout
(model_utils.py:120-332): a uint8 header, fp32 first and last layer, 256-entry k-means codebooks with 8-bit
labels for the hidden weights and for the non-zero wavelet coefficients, and a separate ``_mask.bnr`` bit stream
(1 = coefficient kept).  The byte layout is the reference's; the implementation is vectorised numpy (the
reference builds the bit strings element by element in Python).
"""
from __future__ import annotations

import math
import os
import struct

import numpy as np
import torch

from .Feature_Embedding import FourierEmbedding
from .Feature_Grid_Model import Feature_Grid_Model
from .Smallify_Dropout import SmallifyDropout
from .Straight_Through_Dropout import MaskedWavelet_Straight_Through_Dropout, Straight_Through_Dropout
from .Variational_Dropout_Layer import VariationalDropout
from ..wavelet_transform.Torch_Wavelet_Transform import WaveletFilter3d


def write_dict(dictionary, filename, experiment_path=''):
    with open(os.path.join(experiment_path, filename), 'w') as f:
        for key, value in dictionary.items():
            f.write('%s = %s\n' % (key, value))


_MASK_TYPES = (
    (lambda t: t == 'smallify', SmallifyDropout),
    (lambda t: t == 'straight_through', Straight_Through_Dropout),
    (lambda t: t == 'masked_straight_through', MaskedWavelet_Straight_Through_Dropout),
    (lambda t: 'variational' in t, VariationalDropout),
)


def setup_model(input_channel, hidden_channel, out_channel, num_layer, embedding_type, n_embedding_freq, drop_type,
                drop_momentum, drop_threshold, wavelet_filter, grid_features, grid_size, checkpoint_path):
    grid = torch.empty((grid_features, grid_size, grid_size, grid_size)).uniform_(0, 1)
    filt = WaveletFilter3d(wavelet_filter)

    prototype = None
    if drop_type:
        for matches, cls in _MASK_TYPES:
            if matches(drop_type):
                prototype = cls(grid.shape[1:], drop_momentum, drop_threshold)
        if prototype is None:
            raise ValueError('unknown drop_type %r' % (drop_type,))

    embedder = FourierEmbedding(n_freqs=n_embedding_freq, input_dim=input_channel)  # the only embedding type
    model = Feature_Grid_Model(embedder, grid, prototype, filt, input_channel_data=input_channel,
                               hidden_channel=hidden_channel, out_channel=out_channel, num_layer=num_layer)
    if checkpoint_path:
        model.load_state_dict(torch.load(checkpoint_path))
    return model


def get_net_weights_biases(net):
    weights = [p.data for n, p in net.named_parameters() if n.endswith('.weight')]
    biases = [p.data for n, p in net.named_parameters() if n.endswith('.bias')]
    return weights, biases


# ----------------------------------------------------------------------------------------------------------------
# bit-level helpers (same names and byte results as the reference's, model_utils.py:78-117)
# ----------------------------------------------------------------------------------------------------------------

def kmeans_quantization(w, q):
    """1-D k-means codebook: (labels list, q centres list).  ``w`` is (n, 1)."""
    from sklearn.cluster import KMeans
    w = np.asarray(w, dtype=np.float64).reshape(-1, 1)
    if w.shape[0] < q:
        # fewer values than centres: every value is its own centre (the reference's KMeans call would raise here)
        centres = np.zeros(q, dtype=np.float64)
        centres[:w.shape[0]] = w[:, 0]
        return list(range(w.shape[0])), centres.tolist()
    km = KMeans(n_clusters=q, n_init=4).fit(w)
    return km.labels_.tolist(), km.cluster_centers_.reshape(q).tolist()


def ints_to_bits_to_bytes(all_ints, n_bits):
    """Concatenate ``n_bits``-wide big-endian codes and cut into bytes; the last byte is NOT padded on the right
    (it is the integer value of the remaining bits), as in the reference."""
    vals = np.asarray(all_ints, dtype=np.uint64).reshape(-1)
    shifts = np.arange(n_bits - 1, -1, -1, dtype=np.uint64)
    bits = ((vals[:, None] >> shifts[None, :]) & 1).astype(np.uint8).reshape(-1)
    n_full, rest = divmod(bits.size, 8)
    out = bytearray(np.packbits(bits[:n_full * 8]).tobytes())
    if rest:
        out.append(int(''.join(str(int(b)) for b in bits[n_full * 8:]), 2))
    return out, bool(rest)


def binary_writing(mask_string, filename):
    """Write a '0'/'1' string MSB-first, zero-padding the last byte on the right."""
    bits = np.frombuffer(mask_string.encode('ascii'), dtype=np.uint8) - ord('0') if isinstance(mask_string, str) \
        else np.asarray(mask_string, dtype=np.uint8)
    with open(filename, 'wb') as f:
        f.write(np.packbits(bits.astype(np.uint8)).tobytes())


def read_binary(filename, num_bits):
    n_bytes = (num_bits + 7) // 8
    with open(filename, 'rb') as f:
        raw = np.frombuffer(f.read(n_bytes), dtype=np.uint8)
    return ''.join(map(str, np.unpackbits(raw).tolist()))


def _f32_bytes(t):
    return np.asarray(t.detach().cpu().reshape(-1).numpy(), dtype='<f4').tobytes()


def store_model_parameters(model, filename):
    bit_precision = 8
    n_clusters = int(math.pow(2, bit_precision))
    grids = [g.detach().cpu().reshape(-1).numpy() for g in model.feature_grid]
    nonzero = [int(np.count_nonzero(g)) for g in grids]
    zeros = [int(g.size - nz) for g, nz in zip(grids, nonzero)]
    weights, biases = get_net_weights_biases(model)

    def quantised(values):
        labels, centres = kmeans_quantization(np.asarray(values, dtype=np.float32).reshape(-1, 1), n_clusters)
        blob = struct.pack('%df' % len(centres), *centres)
        packed, _ = ints_to_bits_to_bytes(labels, bit_precision)
        blob += bytes(packed)
        if bit_precision % 8 != 0:
            blob += struct.pack('I', labels[-1])
        return blob

    with open(filename, 'wb') as f:
        f.write(struct.pack('9B', model.num_layer, model.hidden_width, model.input_channel, model.d_in,
                            model.output_channel, bit_precision, int(model.shape_array[-1][0]),
                            len(model.feature_grid), int(model.feature_grid[0].shape[0])))
        f.write(struct.pack('%dI' % len(nonzero), *nonzero))
        f.write(struct.pack('%dI' % len(zeros), *zeros))
        f.write(_f32_bytes(weights[0]))
        f.write(_f32_bytes(biases[0]))
        for w, b in zip(weights[1:-1], biases[1:-1]):
            f.write(quantised(w.detach().cpu().reshape(-1).numpy()))
            f.write(_f32_bytes(b))
        f.write(_f32_bytes(weights[-1]))
        f.write(_f32_bytes(biases[-1]))
        for g in grids:
            f.write(quantised(g[g != 0.0]))
    mask_bits = np.concatenate([(g != 0.0).astype(np.uint8) for g in grids])
    binary_writing(mask_bits, filename + '_mask.bnr')


def restore_model(filename):
    with open(filename, 'rb') as f:
        n_layers, width, input_dim, d_in, output_dim, bit_precision, grid_size, n_grids, feature_size = \
            struct.unpack('9B', f.read(9))
        n_clusters = int(math.pow(2, bit_precision))
        grid_sizes = list(struct.unpack('%dI' % n_grids, f.read(4 * n_grids)))
        zeros = list(struct.unpack('%dI' % n_grids, f.read(4 * n_grids)))

        def read_f32(n):
            return np.frombuffer(f.read(4 * n), dtype='<f4').copy()

        def read_quantised(n):
            centres = read_f32(n_clusters)
            n_bytes = (n * bit_precision + 7) // 8
            raw = np.unpackbits(np.frombuffer(f.read(n_bytes), dtype=np.uint8))
            if (n * bit_precision) % 8:
                # the reference stores the trailing partial byte as a plain integer (not left-aligned)
                rest = (n * bit_precision) % 8
                tail = raw[-8:][8 - rest:]
                raw = np.concatenate([raw[:-8], tail])
            codes = raw[:n * bit_precision].reshape(n, bit_precision)
            labels = codes.dot(1 << np.arange(bit_precision - 1, -1, -1)).astype(np.int64)
            if bit_precision % 8 != 0:
                labels[-1] = struct.unpack('I', f.read(4))[0]
            return centres[labels]

        net_w = [read_f32(input_dim * width)]
        net_b = [read_f32(width)]
        for _ in range(n_layers - 1):
            net_w.append(read_quantised(width * width))
            net_b.append(read_f32(width))
        net_w.append(read_f32(output_dim * width))
        net_b.append(read_f32(output_dim))
        grid_vals = [read_quantised(n) for n in grid_sizes]

    total = sum(grid_sizes) + sum(zeros)
    mask = np.frombuffer(read_binary(filename + '_mask.bnr', total).encode('ascii'), dtype=np.uint8)[:total] == ord('1')
    grids, at = [], 0
    for n_keep, n_zero, vals in zip(grid_sizes, zeros, grid_vals):
        n = n_keep + n_zero
        full = np.zeros(n, dtype=np.float32)
        full[mask[at:at + n]] = vals
        grids.append(full)
        at += n

    # the reference hard-codes fourier / 2 frequencies / db2 here (model_utils.py:310-313)
    model = setup_model(d_in, width, output_dim, n_layers, 'fourier', 2, '', 0.025, 0.75, 'db2', feature_size,
                        grid_size, '')
    wi = bi = gi = 0
    for name, prm in model.named_parameters():
        if 'grid' in name.lower():
            prm.data = torch.from_numpy(grids[gi]).view(prm.shape)
            gi += 1
        elif name.endswith('.weight'):
            prm.data = torch.from_numpy(net_w[wi]).view(prm.shape)
            wi += 1
        elif name.endswith('.bias'):
            prm.data = torch.from_numpy(net_b[bi]).view(prm.shape)
            bi += 1
    return model
