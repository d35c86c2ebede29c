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


def weighted_ranges(count: int, weights):
    """contiguous slices of [0, count) with sizes proportional to `weights` (one per rank)"""
    total = float(sum(weights))
    cuts, acc = [0], 0.0
    for w in weights:
        acc += w
        cuts.append(min(count, int(round(count * acc / total))))
    cuts[-1] = count
    return [(cuts[i], max(cuts[i], cuts[i + 1])) for i in range(len(weights))]


def _raise_status(st: int):
    if st & 2:
        raise ArithmeticError("apocalypse")
    if st & 1:
        raise L.PlaysnarkError(L.PS_ERR_ENCODING, "bad scalar or point encoding")


def _groth16_pipelined(be, tr, q, witness, r: int, s: int, dist, device, rank0_share: float, trace=None,
                       gather_witness: bool = True):
    """world = 2 * parts ranks.  Ranks [0, parts) fold polynomial a, ranks [parts, 2 parts) polynomial b:
    each one the subtree over its n/parts gates (ps_qap_interp_part); an all-gather hands the subtree
    roots to the two leaders, which run the top levels (ps_qap_interp_finish) and broadcast a and b.
    Every rank then derives the scalars of A, B and of C's tail locally and starts those MSM shards
    while rank 0 divides (ps_g16_h_from_ab); h is broadcast into C's scalar vector and the remaining
    shard [NioLP | XiT] follows.  One all-gather of the partial points (and of the device status words)
    ends the proof; no host synchronisation in between."""
    import torch
    from .api import _fr_bytes, HostBuffer
    lib = be.lib
    world, rank = dist.get_world_size(), dist.get_rank()
    parts = world // 2
    g, part = divmod(rank, parts)

    def mark(name):
        # optional stage timeline (CUDA events on the current stream), read by the caller after a sync
        if trace is not None:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            trace.append((name, ev))
    mark("start")
    kh, qh = load_key_sharded(be, tr, world), q._resident(be)
    n, nio = q.nbGates, q.nbIO
    nA, nC, nB = (int(lib.ps_g16_scalar_count(kh, w)) for w in (0, 1, 2))
    head = nio + n - 1                                   # [w_nio | h] ; tail = [s a + r b | s | r | r s]
    ptr = lambda t: C.c_void_p(t.data_ptr())
    # every buffer below is completely written before it is read (no zero fill: ~0.4 GB of memsets per proof)
    new = lambda rows: torch.empty((rows, 8), dtype=torch.int32, device=device)
    bufA, bufC, bufB = new(nA), new(nC), new(nB)
    status = torch.zeros(1, dtype=torch.int32, device=device)
    rb, sb = _fr_bytes([r]), _fr_bytes([s])
    np_ = _tree_leaves(n)                                # leaves of the interpolation tree (power of two >= n)
    e_rows = 2 * np_ // parts if parts > 1 else n
    e_part = new(e_rows)
    m = q.nbVars
    if gather_witness:
        # every rank uploads 1/world of the witness; the slices travel over NVLink (one upload per node)
        chunk = (m + world - 1) // world
        wfull = new(world * chunk)
        lo, hi = min(m, rank * chunk), min(m, (rank + 1) * chunk)
        if isinstance(witness, HostBuffer):
            src = C.c_void_p(witness.ptr.value + 32 * lo)
        else:
            src = _fr_bytes(witness)[32 * lo:32 * hi]
        mine = wfull[rank * chunk:(rank + 1) * chunk]
        be._check(lib.ps_fr_upload(be.ctx, src, hi - lo, ptr(mine), ptr(status)))
        dist.all_gather(list(wfull.view(world, chunk, 8).unbind(0)), mine.clone())
        mark("witness")
        be._check(lib.ps_qap_interp_part_dev(be.ctx, qh, ptr(wfull), g, part, parts, ptr(e_part), ptr(bufC) if nio else None,
                                             ptr(status)))
    else:
        be._check(lib.ps_qap_interp_part(be.ctx, qh, _fr_bytes(witness), g, part, parts, ptr(e_part), ptr(bufC) if nio else None,
                                         ptr(status)))
    mark("interp_part")
    if parts > 1:
        e_all = [new(e_rows) for _ in range(world)]
        dist.all_gather(e_all, e_part)
        mark("gather_roots")
        coef = [new(n), new(n)]
        if part == 0:
            mine = torch.cat(e_all[g * parts:(g + 1) * parts])
            be._check(lib.ps_qap_interp_finish(be.ctx, qh, parts, ptr(mine), ptr(coef[g])))
    else:
        coef = [e_part if g == 0 else new(n), e_part if g == 1 else new(n)]
    mark("interp_finish")
    dist.broadcast(coef[0], src=0)
    dist.broadcast(coef[1], src=parts)
    mark("bcast_ab")
    be._check(lib.ps_g16_scalars_ab(be.ctx, kh, rb, sb, ptr(coef[0]), ptr(coef[1]), ptr(bufA), ptr(bufB), ptr(bufC[head:])))
    hview = bufC[nio:head]
    if rank == 0:
        be._check(lib.ps_g16_h_from_ab(be.ctx, qh, ptr(coef[0]), ptr(coef[1]), ptr(hview)))
    mark("scalars_and_h")
    weights = [rank0_share] + [1.0] * (world - 1)       # rank 0 also divides: it takes a smaller MSM share
    rA = weighted_ranges(nA, weights)[rank]
    rB = weighted_ranges(nB, weights)[rank]
    rT = weighted_ranges(nC - head, weights)[rank]
    rH = weighted_ranges(head, weights)[rank]

    def partials(spans):
        # spans: (lo, hi) inside the base sets of A, C, B; scalar pointers offset to the span
        out = torch.empty(768, dtype=torch.uint8, device=device)
        first = (C.c_size_t * 3)(*[lo for lo, _ in spans])
        cnt = (C.c_size_t * 3)(*[hi - lo for lo, hi in spans])
        views = [buf[lo:hi] if hi > lo else buf for buf, (lo, hi) in zip((bufA, bufC, bufB), spans)]
        be._check(lib.ps_g16_msm_partials(be.ctx, kh, ptr(views[0]), ptr(views[1]), ptr(views[2]), first, cnt, ptr(out)))
        return out

    early = partials([rA, (head + rT[0], head + rT[1]), rB])
    mark("msm_early")
    dist.broadcast(hview, src=0)
    mark("bcast_h")
    late = partials([(0, 0), rH, (0, 0)])
    mark("msm_late")
    pad = torch.zeros(12, dtype=torch.uint8, device=device)
    rec = torch.cat([early, late[192:384], status.view(torch.uint8), pad])   # A | C tail | B | C head | status | pad
    recs = torch.empty((world, rec.numel()), dtype=torch.uint8, device=device)
    dist.all_gather(list(recs.unbind(0)), rec)
    mark("gather_partials")
    _raise_status(int(recs[:, 960:964].contiguous().view(torch.int32).max().item()))
    if rank != 0:
        return None
    oA, oB, oC = C.create_string_buffer(48), C.create_string_buffer(96), C.create_string_buffer(48)
    be._check(lib.ps_g16_combine(be.ctx, ptr(recs), world, rec.numel(), oA, oB, oC))
    return oA.raw, oB.raw, oC.raw


def _tree_leaves(n: int) -> int:
    """leaves of the device's interpolation tree for n gates (csrc/interp.cuh): the power of two >= max(n, 2)"""
    np_ = 2
    while np_ < n:
        np_ <<= 1
    return np_


def load_key_sharded(be, tr, world: int):
    """proving key resident on this rank with MSM windows sized for a 1/world share of every base set"""
    be.set_option("msm_shards", max(1, world))
    try:
        return tr._resident(be)
    finally:
        be.set_option("msm_shards", 1)


def groth16_prove_sharded(be, tr, q, witness, r: int, s: int, dist=None, device="cpu", split_quotient: bool = True,
                          rank0_share: Optional[float] = None, trace=None, gather_witness: bool = True):
    """Groth16Prove (groth16.go:122-211) over the ranks of `dist`.  Every rank holds the proving key.
    With a sparse QAP and an even, power-of-two-halved world the whole proof is pipelined across the
    ranks (_groth16_pipelined: every rank holds the QAP and reads `witness`).  Otherwise the quotient
    runs on rank 0 (`q` and `witness` are only read there), the three scalar vectors are broadcast,
    every rank sums its index range of each MSM (G2 on its second stream), the 768-byte partials are
    all-gathered and rank 0 adds and encodes.  Returns (A, B, C) compressed on rank 0, None elsewhere."""
    import torch
    from .api import _fr_bytes, _sanity
    lib = be.lib
    world = dist.get_world_size() if dist is not None else 1
    rank = dist.get_rank() if dist is not None else 0
    parts = world // 2
    # QAP.sanityCheck (qap.go:177-189) on every rank that reads the witness, BEFORE any C call or collective: the C side
    # trusts qap->m when it reads witness bytes
    if witness is not None and q is not None:
        _sanity(q, witness)
    if rank0_share is None:
        # rank 0 also divides (about 1/16 of the single-GPU MSM time): its MSM share shrinks with the world
        # size so that it reaches the h broadcast together with the others; 0.4 measured best on 8 GPUs
        rank0_share = max(0.2, 1.0 - 0.075 * world)
    if (split_quotient and world >= 2 and world % 2 == 0 and parts & (parts - 1) == 0 and type(q).__name__ == "SparseQAP"
            and parts <= _tree_leaves(q.nbGates) // 2):
        return _groth16_pipelined(be, tr, q, witness, r, s, dist, device, rank0_share, trace, gather_witness)
    kh = load_key_sharded(be, tr, world)
    counts = [int(lib.ps_g16_scalar_count(kh, w)) for w in (0, 1, 2)]
    groups = [L.PS_G1, L.PS_G1, L.PS_G2]
    bufs = [torch.zeros((c, 8), dtype=torch.int32, device=device) for c in counts]
    status = torch.zeros(1, dtype=torch.int32, device=device)
    ptr = lambda t: C.c_void_p(t.data_ptr())
    if rank == 0:
        try:
            status[0] = lib.ps_g16_scalars(be.ctx, kh, q._resident(be), _fr_bytes(witness), _fr_bytes([r]), _fr_bytes([s]),
                                           ptr(bufs[0]), ptr(bufs[1]), ptr(bufs[2]))
        except Exception:
            # the other ranks are about to enter the status broadcast: hand them an error instead of leaving them blocked
            status[0] = L.PS_ERR_ARG
            if world > 1:
                dist.broadcast(status, src=0)
            raise
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
        parts_l = [torch.zeros_like(part) for _ in range(world)]
        dist.all_gather(parts_l, part)
    else:
        be.sync()
        parts_l = [part]
    if rank != 0:
        return None
    res, off = [], 0
    for w in range(3):
        allp = torch.cat([p[off:off + sizes[w]] for p in parts_l]).contiguous()
        out = C.create_string_buffer(48 if groups[w] == L.PS_G1 else 96)
        be._check(lib.ps_msm_combine(be.ctx, groups[w], C.c_void_p(allp.data_ptr()), world, out))
        res.append(out.raw)
        off += sizes[w]
    return res[0], res[2], res[1]   # A, B, C
