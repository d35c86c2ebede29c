"""world_size-2 test of the sharded MSM host logic on CPU: two gloo ranks, each running the C ABI in
host emulation on its point range, one all-gather of the 192-byte partials, sum on rank 0."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys, random
sys.path.insert(0, %(root)r)
import numpy as np, torch, torch.distributed as dist
from oracle import ps_oracle as O
from playsnark_b200 import _lib as L, api, build as B, dist as D
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
lib = L.bind(B.build_host_emulation(os.path.join(%(root)r, "tests", "_build")))
be = api.Backend(0, lib=lib)
n_total = 75
rng = random.Random(99)                      # same stream on every rank
ks = [rng.randrange(1, O.R) for _ in range(n_total)]
sc = [rng.randrange(O.R) for _ in range(n_total)]
FULL = os.environ.get("PS_GLOO_PART", "all") == "all"     # the 4-rank run repeats only what differs with 4 ranks
for group, F, gen, comp in ((L.PS_G1, O.F1, O.G1_GEN, O.g1_compress), (L.PS_G2, O.F2, O.G2_GEN, O.g2_compress)) if FULL else ():
    lo, hi = D.shard_range(n_total, rank, world)
    bases = be.bases_from_scalars(group, ks[lo:hi], 7, -1)          # this rank's range only
    limbs = np.array([[(s >> (32 * j)) & 0xFFFFFFFF for j in range(8)] for s in sc[lo:hi]], dtype=np.uint32)
    t = torch.from_numpy(limbs.view(np.int32))
    res = D.msm_sharded(be, bases, t, hi - lo, dist)
    if rank == 0:
        want = comp(O.pt_mul(F, sum(k * s for k, s in zip(ks, sc)) %% O.R, gen))
        assert res == want, (group, res.hex(), want.hex())
    else:
        assert res is None
assert D.shard_range(10, 0, 3) == (0, 4) and D.shard_range(10, 2, 3) == (7, 10)
# Groth16 with the three MSMs sharded over the ranks, against the oracle prover (same seeds on all ranks)
from tests import helpers as H
r1cs, wit = H.mixed_circuit(12, 5, 6)
oq = O.to_qap(r1cs)
smp = O.Sampler(5)
tr = O.groth16_setup(oq, smp)
rr, ss = smp.fr(), smp.fr()
res = D.groth16_prove_sharded(be, H.mirror_g16_setup(tr), H.mirror_qap(oq), wit, rr, ss, dist) if FULL else None
if rank == 0 and FULL:
    want = O.groth16_prove(tr, oq, wit, rr, ss)
    assert res == (O.g1_compress(want["A"]), O.g2_compress(want["B"]), O.g1_compress(want["C"]))
# the same through the sparse form with the quotient split over ranks 0 and 1
r2, wit2 = H.mixed_circuit(16, 9, 8)
oq2 = O.to_qap(r2)
sq2 = api.SparseQAP.from_dense_rows(len(r2.vars), r2.nb_io(), r2.left, r2.right, r2.out)
smp2 = O.Sampler(9)
tr2 = O.groth16_setup(oq2, smp2)
r3, s3 = smp2.fr(), smp2.fr()
res2 = D.groth16_prove_sharded(be, H.mirror_g16_setup(tr2), sq2, wit2, r3, s3, dist)
if rank == 0:
    want2 = O.groth16_prove(tr2, oq2, wit2, r3, s3)
    assert res2 == (O.g1_compress(want2["A"]), O.g2_compress(want2["B"]), O.g1_compress(want2["C"]))
bad2 = list(wit2); bad2[-1] = (bad2[-1] + 1) %% O.R
try:
    D.groth16_prove_sharded(be, H.mirror_g16_setup(tr2), sq2, bad2, r3, s3, dist)
    raise SystemExit("expected apocalypse on every rank (split quotient)")
except ArithmeticError:
    pass
bad = list(wit); bad[-1] = (bad[-1] + 1) %% O.R
try:
    if FULL:
        D.groth16_prove_sharded(be, H.mirror_g16_setup(tr), H.mirror_qap(oq), bad, rr, ss, dist)
        raise SystemExit("expected apocalypse on every rank")
except ArithmeticError:
    pass
dist.barrier()
if rank == 0:
    print("SHARDED_OK")
dist.destroy_process_group()
'''


def _run(tmp_path, nproc, port, part="all"):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT})
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1", PS_GLOO_PART=part)
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=%d" % nproc,
                          "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                         capture_output=True, text=True, timeout=1200, env=env)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    assert "SHARDED_OK" in res.stdout


def test_sharded_msm_two_ranks(tmp_path):
    """2 ranks: sharded MSM; Groth16 with a dense QAP (quotient on rank 0) and with a sparse QAP
    (pipelined flow, one aggregate polynomial per rank)."""
    _run(tmp_path, 2, 29533)


def test_sharded_groth16_four_ranks(tmp_path):
    """4 ranks: the interpolation of each aggregate polynomial is split over two subtrees
    (ps_qap_interp_part / ps_qap_interp_finish), uneven MSM shares, h broadcast into C's scalars."""
    _run(tmp_path, 4, 29534, part="sparse")


def test_weighted_ranges_cover_exactly():
    """MSM shares of the pipelined prover: contiguous, exhaustive, proportional, robust to tiny counts."""
    from playsnark_b200 import dist as D
    for count in (0, 1, 5, 1000, (1 << 20) + 2):
        for world in (2, 4, 8):
            w = [max(0.2, 1.0 - 0.075 * world)] + [1.0] * (world - 1)
            r = D.weighted_ranges(count, w)
            assert r[0][0] == 0 and r[-1][1] == count and len(r) == world
            assert all(a[1] == b[0] for a, b in zip(r, r[1:])) and all(lo <= hi for lo, hi in r)
            if count >= 1000:
                sizes = [hi - lo for lo, hi in r]
                assert abs(sizes[0] / count - w[0] / sum(w)) < 0.01 and max(sizes[1:]) - min(sizes[1:]) <= 1
