"""Host-side sharding rules of the data-parallel path (no device code; shared by the trainer, the reconstruction
driver and the CPU/gloo tests).

Training: sample i of optimiser step s on rank r is element ``s * global_batch + r * batch + i`` of ONE global
counter-based (Philox) stream, so the union over ranks is exactly the batch a single GPU would draw with
``batch * world`` samples; the loss is pre-scaled by 1/global_batch and the flat gradients are summed.
Reconstruction: contiguous slabs along volume dim 0, no communication.
"""
from __future__ import annotations


def sample_stream_offset(step: int, rank: int, batch: int, world: int) -> int:
    return step * batch * world + rank * batch


def loss_scale(batch: int, world: int) -> float:
    return 1.0 / float(batch * world)


def slab_bounds(extent: int, rank: int, world: int):
    """[begin, end) rows of dim 0 owned by `rank`: sizes differ by at most one, earlier ranks get the extra row."""
    base, extra = divmod(int(extent), int(world))
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)
