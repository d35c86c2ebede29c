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


class _KeyBases:
    """view of a base set owned by a resident key (not freed here)"""

    def __init__(self, handle, group):
        self.handle, self.group = handle, group


def groth16_prove_sharded(be, tr, q, witness, r: int, s: int, dist=None, device="cpu"):
    """Groth16Prove (groth16.go:122-211) with its three MSMs sharded by point range over the ranks
    of `dist`.  Every rank holds the proving key; rank 0 holds the QAP, runs the quotient and
    broadcasts the scalar vectors; partial points are all-gathered and summed on rank 0, which
    returns (A, B, C) as compressed bytes (other ranks return None).  `witness` is only read on rank 0."""
    import torch
    from .api import _fr_bytes
    lib = be.lib
    world = dist.get_world_size() if dist is not None else 1
    rank = dist.get_rank() if dist is not None else 0
    kh = tr._resident(be)
    counts = [int(lib.ps_g16_scalar_count(kh, w)) for w in (0, 1, 2)]
    groups = [L.PS_G1, L.PS_G1, L.PS_G2]
    bufs = [torch.zeros((c, 8), dtype=torch.int32, device=device) for c in counts]
    status = torch.zeros(1, dtype=torch.int32, device=device)
    if rank == 0:
        st = lib.ps_g16_scalars(be.ctx, kh, q._resident(be), _fr_bytes(witness), _fr_bytes([r]), _fr_bytes([s]),
                                C.c_void_p(bufs[0].data_ptr()), C.c_void_p(bufs[1].data_ptr()), C.c_void_p(bufs[2].data_ptr()))
        status[0] = st
    if world > 1:
        dist.broadcast(status, src=0)
    st = int(status[0])
    if st == L.PS_ERR_REMAINDER:
        raise ArithmeticError("apocalypse")
    be._check(st)
    if world > 1:
        for b in bufs:
            dist.broadcast(b, src=0)
    sizes = [PARTIAL_BYTES[g] for g in groups]
    part = torch.zeros(sum(sizes), dtype=torch.uint8, device=device)
    off = 0
    for w in range(3):
        lo, hi = shard_range(counts[w], rank, world)
        kb = _KeyBases(C.c_void_p(lib.ps_g16_key_bases(kh, w)), groups[w])
        msm_partial(be, kb, bufs[w][lo:hi], hi - lo, part[off:off + sizes[w]], first=lo)
        off += sizes[w]
    if world > 1:
        parts = [torch.zeros_like(part) for _ in range(world)]
        dist.all_gather(parts, part)
    else:
        parts = [part]
    if rank != 0:
        return None
    res, off = [], 0
    for w in range(3):
        allp = torch.cat([p[off:off + sizes[w]] for p in parts]).contiguous()
        out = C.create_string_buffer(48 if groups[w] == L.PS_G1 else 96)
        be._check(lib.ps_msm_combine(be.ctx, groups[w], C.c_void_p(allp.data_ptr()), world, out))
        res.append(out.raw)
        off += sizes[w]
    return res[0], res[2], res[1]   # A, B, C
