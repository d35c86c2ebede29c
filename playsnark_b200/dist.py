"""Multi-GPU MSM: point-range shards, one process per GPU.

Every rank holds the bases of its own index range resident and receives the matching scalars; it runs
the whole Pippenger pipeline on its range (ps_msm_device) and leaves one XYZZ partial (192 B G1 /
384 B G2) in device memory.  The only exchange is an all-gather of those records; rank 0 adds them and
emits the compressed point (ps_msm_combine).  There is no NCCL reduction for curve points, hence
gather-then-add.  The backend's stream must be torch's current stream (Backend.set_stream) so that the
collective is ordered after the kernels.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

from . import _lib as L

PARTIAL_BYTES = {L.PS_G1: 192, L.PS_G2: 384}


def shard_range(n_total: int, rank: int, world: int):
    """contiguous slice [lo, hi) of the point range owned by `rank`"""
    base, rem = divmod(n_total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def msm_partial(be, bases, scalars_le, n: int, out_partial, first: int = 0):
    """scalars_le: torch int32/uint8 tensor with n x 8 little-endian limbs on the backend's device;
    out_partial: uint8 tensor of PARTIAL_BYTES[group]."""
    be._check(be.lib.ps_msm_device(be.ctx, bases.handle, first, C.c_void_p(scalars_le.data_ptr()), n,
                                   C.c_void_p(out_partial.data_ptr())))


def msm_sharded(be, bases, scalars_le, n: int, dist=None, first: int = 0) -> Optional[bytes]:
    """MSM over this rank's shard, gathered and summed on rank 0 (returns the compressed point there,
    None elsewhere).  `dist` is torch.distributed (initialised) or None for a single process."""
    import torch
    group = bases.group
    nb = PARTIAL_BYTES[group]
    part = torch.zeros(nb, dtype=torch.uint8, device=scalars_le.device)
    msm_partial(be, bases, scalars_le, n, part, first)
    world = dist.get_world_size() if dist is not None else 1
    rank = dist.get_rank() if dist is not None else 0
    if world > 1:
        parts = [torch.zeros(nb, dtype=torch.uint8, device=part.device) for _ in range(world)]
        dist.all_gather(parts, part)
        allp = torch.cat(parts)
    else:
        allp = part
    if rank != 0:
        return None
    out = C.create_string_buffer(48 if group == L.PS_G1 else 96)
    be._check(be.lib.ps_msm_combine(be.ctx, group, C.c_void_p(allp.data_ptr()), world, out))
    return out.raw
