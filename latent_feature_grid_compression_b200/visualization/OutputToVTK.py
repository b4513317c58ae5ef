"""Full-volume reconstruction and error statistics -- counterpart of the reference's visualization/OutputToVTK.py.

The reference walks the volume in 32^3 tiles: linspace coordinates on the CPU, host->device copy, an eval forward
that re-synthesises the latent grid for every tile, device->host copy (512 tiles for 255^3).  Here the grid is
decoded once and the whole volume (or one slab of it, for multi-GPU) is evaluated by a single launch
(``lfgc_reconstruct``) that writes the output volume directly; the per-voxel coordinates are bit-identical to the
reference's because the per-axis tables are built with the same per-tile fp32 ``linspace``.
"""
from __future__ import annotations

import struct

import numpy as np
import torch as th

from .. import ops
from ..model.Feature_Grid_Model import _multipliers


def axis_tables(dataset, tiled_res=32, device='cuda'):
    """Normalised coordinate of every voxel index along each axis, exactly as field_from_net builds them
    (OutputToVTK.py:23-37): per tile ``linspace(b/(R-1), (e-1)/(R-1), e-b) * 2 - 1``, times ``scales``."""
    res = dataset.vol_res_touple
    cache = dataset.__dict__.setdefault('_lfgc_axis_tables', {})
    key = (tuple(int(r) for r in res), int(tiled_res), str(device))
    if key in cache:
        return cache[key]
    min_idx, max_idx, scales = dataset.min_idx.cpu(), dataset.max_idx.cpu(), dataset.scales.cpu()
    span = max_idx - min_idx
    tables = []
    for a, R in enumerate(res):
        vals = th.zeros(R, dtype=th.float)
        for b in range(0, R, tiled_res):
            e = min(b + tiled_res, R)
            lo = (min_idx[a] + th.tensor(b / (R - 1), dtype=th.float) * span[a]) / span[a]
            hi = (min_idx[a] + th.tensor((e - 1) / (R - 1), dtype=th.float) * span[a]) / span[a]
            vals[b:e] = th.linspace(float(lo), float(hi), e - b, dtype=th.float)
        tables.append((scales[a] * (2.0 * vals - 1.0)).to(device).contiguous())
    cache[key] = tables
    return tables


_COPY_STREAMS = {}


def _copy_stream(dev):
    key = (dev.type, dev.index)
    if key not in _COPY_STREAMS:
        _COPY_STREAMS[key] = th.cuda.Stream(device=dev)
    return _COPY_STREAMS[key]


def field_from_net(dataset, net, is_cuda=True, tiled_res=32, verbose=False, slab=None, to_cpu=True, host_out=None):
    """Reconstructed volume (R0, R1, R2).  ``slab=(begin, end)`` restricts the work to rows [begin, end) of dim 0
    and returns only that slab (the multi-GPU sharding unit).  ``host_out``: a (pinned) CPU tensor of the slab's shape
    that receives the result (one asynchronous D2H copy + a stream synchronise instead of a pageable ``.cpu()``)."""
    if not th.cuda.is_available():
        raise ops.L.LfgcError('field_from_net runs on CUDA only (no CPU fallback)')
    res = tuple(int(r) for r in dataset.vol_res_touple)
    begin, end = (0, res[0]) if slab is None else slab
    with th.no_grad():
        geom = net.geometry()
        dev = next(net.parameters()).device
        mults, _ = _multipliers(net.mask_specs())
        grid_cl = ops.decode_fwd(geom, [f.detach().contiguous() for f in net.feature_grid], mults)
        axes = axis_tables(dataset, tiled_res, dev)
        if host_out is not None:
            # Result wanted in HOST memory: the slab is computed in chunks of rows and every chunk's device -> host copy runs
            # on a second stream while the next chunk is computed (66 MB at 255^3: the copy takes about as long as the
            # compute; one copy at the end made the call 3.7 ms, 2.4 ms of it compute).
            if tuple(host_out.shape) != (end - begin, res[1], res[2]):
                raise ops.L.LfgcError('host_out has shape %s, the slab is %s' % (tuple(host_out.shape), (end - begin, res[1], res[2])))
            out = th.empty((end - begin, res[1], res[2]), device=dev, dtype=th.float32)
            rows = max(1, min(end - begin, (1 << 21) // max(1, res[1] * res[2])))     # ~2 M voxels per chunk
            main, side = th.cuda.current_stream(), _copy_stream(dev)
            mlp = net.mlp_flat()
            for b in range(begin, end, rows):
                e = min(end, b + rows)
                ops.reconstruct(geom, grid_cl, mlp, res, axes, b, e, clamp=True, out=out[b - begin:e - begin])
                ev = th.cuda.Event()
                ev.record(main)
                side.wait_event(ev)
                with th.cuda.stream(side):
                    host_out[b - begin:e - begin].copy_(out[b - begin:e - begin], non_blocking=True)
            side.synchronize()
            return host_out
        out = ops.reconstruct(geom, grid_cl, net.mlp_flat(), res, axes, begin, end, clamp=True)
    return out.cpu() if to_cpu else out


def calculate_deviation_statistics(prediction, ground_truth):
    """psnr, l1, mse, rmse with PSNR = 10 log10((max gt - min gt)^2 / mse) (OutputToVTK.py:53-60); CUDA inputs are
    reduced on the device (fp64 accumulators), CPU inputs are moved there first."""
    if not th.cuda.is_available():
        raise ops.L.LfgcError('calculate_deviation_statistics runs on CUDA only (no CPU fallback)')
    pred = prediction.to('cuda', th.float32).contiguous()
    gt = ground_truth.to('cuda', th.float32).contiguous()
    acc = th.tensor([0.0, 0.0, -float('inf'), float('inf')], dtype=th.float64, device='cuda')
    ops.deviation_stats_accumulate(pred, gt, acc)
    sq, ab, mx, mn = acc.tolist()
    n = pred.numel()
    mse = sq / n
    l1 = ab / n
    psnr = 10.0 * np.log10((mx - mn) ** 2 / mse)
    rmse = float(np.sqrt(mse))
    print('PSNR:', psnr, 'l1:', l1, 'mse:', mse, 'rmse:', rmse)
    return float(psnr), float(l1), float(mse), rmse


def write_vti(filename, volume):
    """Minimal VTK ImageData writer (appended raw fp32 point data 'sf'), standing in for pyevtk.hl.imageToVTK."""
    vol = np.ascontiguousarray(volume, dtype='<f4')
    nx, ny, nz = vol.shape
    data = vol.transpose(2, 1, 0).tobytes()  # VTK wants x fastest
    header = ('<?xml version="1.0"?>\n<VTKFile type="ImageData" version="1.0" byte_order="LittleEndian" '
              'header_type="UInt64">\n<ImageData WholeExtent="0 %d 0 %d 0 %d" Origin="0 0 0" Spacing="1 1 1">\n'
              '<Piece Extent="0 %d 0 %d 0 %d">\n<PointData Scalars="sf">\n'
              '<DataArray type="Float32" Name="sf" format="appended" offset="0"/>\n</PointData>\n<CellData/>\n'
              '</Piece>\n</ImageData>\n<AppendedData encoding="raw">\n_' % ((nx - 1, ny - 1, nz - 1) * 2))
    with open(filename + '.vti', 'wb') as f:
        f.write(header.encode('ascii'))
        f.write(struct.pack('<Q', len(data)))
        f.write(data)
        f.write(b'\n</AppendedData>\n</VTKFile>\n')


def tiled_net_out(dataset, net, is_cuda, gt_vol=None, evaluate=True, write_vols=False, filename='vol'):
    """Same contract as the reference (OutputToVTK.py:64-82): eval-mode reconstruction, optional statistics and
    .vti dumps, model back in train mode; returns (psnr, l1, mse, rmse)."""
    if is_cuda:
        net = net.cuda()
    net.eval()
    full_vol = field_from_net(dataset, net, is_cuda, tiled_res=32, to_cpu=False)
    psnr = l1_diff = mse = rmse = 0
    if evaluate and gt_vol is not None:
        psnr, l1_diff, mse, rmse = calculate_deviation_statistics(full_vol, gt_vol)
    if write_vols:
        write_vti(filename, full_vol.cpu().numpy())
        if gt_vol is not None:
            write_vti('gt', gt_vol.cpu().numpy())
    net.train()
    return psnr, l1_diff, mse, rmse
