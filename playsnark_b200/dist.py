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


def groth16_prove_sharded(be, tr, q, witness, r: int, s: int, dist=None, device="cpu", split_quotient: bool = True):
    """Groth16Prove (groth16.go:122-211) over the ranks of `dist`.  Every rank holds the proving key.
    Quotient: rank 0 (and, with split_quotient and a sparse QAP, rank 1 for the second aggregate
    polynomial) -- `q` and `witness` are only read there.  The three scalar vectors are broadcast, every
    rank sums its index range of each MSM (G2 on its second stream), the 768-byte partials are
    all-gathered and rank 0 adds and encodes.  Returns (A, B, C) compressed on rank 0, None elsewhere."""
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
    split = bool(split_quotient and world > 1 and not getattr(q, "left", None) is None and type(q).__name__ == "SparseQAP")
    ptr = lambda t: C.c_void_p(t.data_ptr())
    if split:
        n = q.nbGates
        wb = _fr_bytes(witness) if rank in (0, 1) else None
        coef_b = torch.zeros((n, 8), dtype=torch.int32, device=device)
        if rank == 1:
            status[0] = lib.ps_qap_aggregate_one(be.ctx, q._resident(be), wb, 1, ptr(coef_b))
            be.sync()
        if rank == 0:
            coef_a = torch.zeros((n, 8), dtype=torch.int32, device=device)
            st0 = lib.ps_qap_aggregate_one(be.ctx, q._resident(be), wb, 0, ptr(coef_a))
            be.sync()
        dist.broadcast(coef_b, src=1)
        if rank == 0:
            st = st0 or lib.ps_g16_scalars_from_ab(be.ctx, kh, q._resident(be), wb, _fr_bytes([r]), _fr_bytes([s]), ptr(coef_a),
                                                   ptr(coef_b), ptr(bufs[0]), ptr(bufs[1]), ptr(bufs[2]))
            status[0] = st
        st1 = status.clone()
        dist.broadcast(st1, src=1)
        dist.broadcast(status, src=0)
        if int(status[0]) == 0 and int(st1[0]) != 0:
            status = st1
    else:
        if rank == 0:
            status[0] = lib.ps_g16_scalars(be.ctx, kh, q._resident(be), _fr_bytes(witness), _fr_bytes([r]), _fr_bytes([s]),
                                           ptr(bufs[0]), ptr(bufs[1]), ptr(bufs[2]))
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
    ranges = [shard_range(counts[w], rank, world) for w in range(3)]
    first = (C.c_size_t * 3)(*[lo for lo, _ in ranges])
    cnt = (C.c_size_t * 3)(*[hi - lo for lo, hi in ranges])
    views = [bufs[w][ranges[w][0]:ranges[w][1]] for w in range(3)]
    be._check(lib.ps_g16_msm_partials(be.ctx, kh, ptr(views[0]) if cnt[0] else ptr(bufs[0]), ptr(views[1]) if cnt[1] else ptr(bufs[1]),
                                      ptr(views[2]) if cnt[2] else ptr(bufs[2]), first, cnt, ptr(part)))
    if world > 1:
        parts = [torch.zeros_like(part) for _ in range(world)]
        dist.all_gather(parts, part)
    else:
        be.sync()
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
