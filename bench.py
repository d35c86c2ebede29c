#!/usr/bin/env python3
"""Benchmark of the prover hot path (contract: see the task statement / DESIGN.md section 6).

Metric (BASELINE.json): G1 MSM points/s -- the operation that dominates Groth16Prove / PHGR13Prove
(Poly.BlindEval, algebra.go:348-359) -- on synthetic data: bases k_i*G built on the device, scalars
uniform below 2^254 (< r).  One "step" = one MSM over the rank's point range.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--log-n L] [--impl reference]

N = 1: one MSM of 2^L points (default L = 24, configs[3] of BASELINE.json, the size the north-star's
roofline target is quoted on).  N > 1 (torchrun, one rank per GPU): every rank owns its own 2^L-point
range (weak scaling; the MSM shards by point range with no data-path collective), the partial points
are all-gathered over NCCL (192 B per rank) and summed on rank 0.
`value` = points/s with scalars resident in HBM (ps_msm_device); `e2e` = the same through the
reference-facing call with HOST buffers (ps_msm: pinned big-endian scalars in, compressed point out).
`--impl reference` times the CPU port of the reference algorithm (oracle/ps_oracle.c: one bit-serial
scalar multiplication per term, single-threaded like the Go code) on a bounded sample.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

R = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001
METRIC = "g1_msm_points_per_s"
UNIT = "points/s"
IMAD_PER_FP_MUL = 600.0      # 300 32x32->64 MACs as IMAD.LO/IMAD.HI pairs
FP_MUL_PER_MADD = 10.0       # XYZZ mixed addition, 8M + 2S


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def random_scalars_be(n: int, seed: int):
    """n uniform 254-bit scalars as big-endian 32-byte rows (numpy uint8 [n, 32])."""
    import numpy as np
    rng = np.random.default_rng(seed)
    a = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
    a[:, 0] &= 0x3F
    return a


def be_to_le_limbs(a):
    """[n, 32] big-endian bytes -> [n, 8] little-endian uint32 limbs (standard form)."""
    import numpy as np
    return np.ascontiguousarray(a[:, ::-1]).view("<u4").reshape(-1, 8).copy()


def expected_exponent(ks_be, sc_be) -> int:
    """sum_i k_i s_i mod r for the bases k_i*G and scalars s_i (oracle's C dot product; checker, untimed)"""
    from oracle import c_oracle as CO
    return CO.fr_dot(ks_be, sc_be)


def expected_point(group: int, e: int) -> bytes:
    from oracle import ps_oracle as O
    return O.g1_compress(O.g1_mul(e)) if group == 1 else O.g2_compress(O.g2_mul(e))


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        super().__init__(daemon=True)
        self.gpu_index, self.rows, self.proc = gpu_index, [], None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu_index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
            for line in self.proc.stdout:
                self.rows.append([x.strip() for x in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc:
            self.proc.terminate()
        self.join(timeout=2)
        sm, mx, reasons = [], 0.0, set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx = max(mx, float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


def reference_arm(args):
    """CPU port of the reference's BlindEval, single-threaded, bounded sample of the same workload."""
    rank = env_int("RANK", 0)
    if rank != 0:
        return
    from oracle import c_oracle as CO, ps_oracle as O
    sample = 1 << args.ref_log_n
    # bases k_i * G for the sample, built with the oracle (untimed)
    sc = random_scalars_be(sample, 3)
    pts = []
    acc = None
    # cheap distinct multiples of G (untimed setup): P_i = (i+1) * k0 * G by repeated addition
    p0 = O.g1_mul(0x1234567 | 1)
    for _ in range(sample):
        acc = O.g1_add(acc, p0)
        pts.append(acc)
    pbytes = b"".join(O.g1_affine_bytes(p) for p in pts)
    sbytes = sc.tobytes()
    for _ in range(args.warmup if args.warmup < 2 else 1):
        CO.blind_eval_g1(pbytes[:96 * 64], sbytes[:32 * 64])
    t0 = time.perf_counter()
    for _ in range(args.steps):
        CO.blind_eval_g1(pbytes, sbytes)
    dt = time.perf_counter() - t0
    val = sample * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": {"workload": "G1 MSM 2^%d points/GPU, random 254-bit scalars" % args.log_n,
                   "sample": "2^%d points per step" % args.ref_log_n},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": 1, "kind": "port",
                         "sample": "Poly.BlindEval (algebra.go:348-359) restated in C (oracle/ps_oracle.c), one bit-serial "
                                   "scalar-mul per term, 2^%d of the 2^%d points per step, %d host cores present, 1 used "
                                   "(the reference is single-threaded); Go toolchain absent, reference not buildable"
                                   % (args.ref_log_n, args.log_n, os.cpu_count() or 0)},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--log-n", type=int, default=24)
    ap.add_argument("--ref-log-n", type=int, default=12)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--tables", type=int, default=-1, help="precomputed window tables (-1 = all windows)")
    ap.add_argument("--window-bits", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--g2-log-n", type=int, default=20, help="also time a G2 MSM of 2^k points per GPU (0 = skip)")
    ap.add_argument("--no-small-configs", action="store_true", help="skip the 2^10 dense / 2^16 sparse prove latencies")
    ap.add_argument("--groth16-log-n", type=int, default=20,
                    help="also time a full Groth16 prove on a sparse synthetic circuit of 2^k constraints (0 = skip)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl != "reference":
        args.warmup = 3
    if args.impl == "reference":
        return reference_arm(args)

    import numpy as np
    import torch
    import playsnark_b200 as ps
    from playsnark_b200 import _lib as L

    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    be = ps.Backend(local)
    stream = torch.cuda.current_stream()
    be.set_stream(stream.cuda_stream)
    lib = be.lib

    n = 1 << args.log_n
    # synthetic inputs: bases k_i*G (fixed-base kernel on the device, untimed), scalars uniform < 2^254
    ks = random_scalars_be(n, 1000 + rank)
    sc_be = random_scalars_be(n, 2000 + rank)
    t0 = time.time()
    bases = be.bases_from_scalars(L.PS_G1, ks.tobytes(), args.window_bits, args.tables)
    be.sync()
    t_bases = time.time() - t0
    d_scalars = torch.from_numpy(be_to_le_limbs(sc_be).view(np.int32)).to(dev)
    h_scalars = torch.from_numpy(sc_be).pin_memory()            # e2e input: pinned host, wire format
    d_part = torch.zeros(192, dtype=torch.uint8, device=dev)
    d_all = torch.zeros(192 * world, dtype=torch.uint8, device=dev)
    out = C.create_string_buffer(48)

    def step_resident():
        be._check(lib.ps_msm_device(be.ctx, bases.handle, 0, C.c_void_p(d_scalars.data_ptr()), n, C.c_void_p(d_part.data_ptr())))
        if world > 1:
            dist.all_gather_into_tensor(d_all, d_part)

    def step_e2e():
        if world == 1:
            be._check(lib.ps_msm(be.ctx, bases.handle, C.c_void_p(h_scalars.data_ptr()), n, out))
            return out.raw
        # host scalars -> device limbs -> partial -> gather -> sum on rank 0 -> compressed point on host
        d_be = h_scalars.to(dev, non_blocking=True)
        limbs = d_be.flip(1).contiguous().view(torch.int32)
        be._check(lib.ps_msm_device(be.ctx, bases.handle, 0, C.c_void_p(limbs.data_ptr()), n, C.c_void_p(d_part.data_ptr())))
        dist.all_gather_into_tensor(d_all, d_part)
        if rank == 0:
            be._check(lib.ps_msm_combine(be.ctx, L.PS_G1, C.c_void_p(d_all.data_ptr()), world, out))
        return out.raw

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step_resident()
    barrier()
    # Correctness gate on the exact configuration that is timed (same bases, scalars, window, scatter path): the
    # rank's partial must equal (sum_i k_i s_i mod r) * G, the exponent-level check of groth16_test.go:32-107.
    # The expectation comes from the oracle (checker only, untimed): a C dot product over Fr and one scalar-mul.
    exp_e = expected_exponent(ks, sc_be)
    del ks
    be._check(lib.ps_msm_combine(be.ctx, L.PS_G1, C.c_void_p(d_part.data_ptr()), 1, out))
    parity = {"g1_2p%d_resident" % args.log_n: out.raw == expected_point(L.PS_G1, exp_e)}
    barrier()

    # integer-multiply peak measured on this device (MEASURED_PEAKS.json has no integer figure)
    v, ms_ = C.c_double(), C.c_double()
    be._check(lib.ps_bench_intpipe(be.ctx, 0, 2000, C.byref(v), C.byref(ms_)))
    imad_peak = v.value
    be._check(lib.ps_bench_fieldmul(be.ctx, 1, 1000, C.byref(v), C.byref(ms_)))
    fpmul_rate = v.value

    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.3)
    launches0 = be.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    phase = {"sort_ms": 0.0, "accumulate_ms": 0.0, "combine_ms": 0.0, "reduce_ms": 0.0, "total_ms": 0.0}
    barrier()
    e0.record(stream)
    for _ in range(args.steps):
        step_resident()
    e1.record(stream)
    barrier()
    ms_total = e0.elapsed_time(e1)
    launches = be.launch_count() - launches0
    tm = be.msm_timing()            # phases of the last step (CUDA events on the launching stream)
    for k in phase:
        phase[k] = tm[k]
    clocks = sampler.stop() if sampler else None

    # e2e: host buffers in, host bytes out, copies inside the timed region
    for _ in range(2):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        res = step_e2e()
    barrier()
    e2e_s = time.perf_counter() - t0
    if world == 1:
        parity["g1_2p%d_e2e" % args.log_n] = res == expected_point(L.PS_G1, exp_e)
    else:
        exps = [None] * world
        dist.all_gather_object(exps, exp_e)
        if rank == 0:
            parity["g1_2p%d_x%d_e2e" % (args.log_n, world)] = res == expected_point(L.PS_G1, sum(exps) % R)

    # secondary figure: G2 MSM (same pipeline over Fp2), resident scalars, every rank its own range
    g2 = None
    if args.g2_log_n:
        n2 = 1 << args.g2_log_n
        bases2 = be.bases_from_scalars(L.PS_G2, random_scalars_be(n2, 3000 + rank).tobytes(), args.window_bits, args.tables)
        ks2, sc2 = random_scalars_be(n2, 3000 + rank), random_scalars_be(n2, 4000 + rank)
        d_sc2 = torch.from_numpy(be_to_le_limbs(sc2).view(np.int32)).to(dev)
        d_part2 = torch.zeros(384, dtype=torch.uint8, device=dev)
        step2 = lambda: be._check(lib.ps_msm_device(be.ctx, bases2.handle, 0, C.c_void_p(d_sc2.data_ptr()), n2, C.c_void_p(d_part2.data_ptr())))
        for _ in range(3):
            step2()
        barrier()
        out2 = C.create_string_buffer(96)
        be._check(lib.ps_msm_combine(be.ctx, L.PS_G2, C.c_void_p(d_part2.data_ptr()), 1, out2))
        parity["g2_2p%d_resident" % args.g2_log_n] = out2.raw == expected_point(L.PS_G2, expected_exponent(ks2, sc2))
        del ks2, sc2
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record(stream)
        for _ in range(args.steps):
            step2()
        f1.record(stream)
        barrier()
        info2 = (C.c_int * 4)()
        be._check(lib.ps_bases_info(bases2.handle, info2))
        g2 = {"ms": f0.elapsed_time(f1), "phases": be.msm_timing(), "c": info2[0], "W": info2[1]}
        bases2.close()
        del d_sc2

    t = torch.tensor([ms_total, e2e_s * 1e3, g2["ms"] if g2 else 0.0], dtype=torch.float64, device=dev)
    ok = torch.tensor([1.0 if all(parity.values()) else 0.0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    ms_total, e2e_ms = float(t[0]), float(t[1])
    parity["all_ranks_ok"] = bool(ok.item() > 0.5)
    if g2:
        g2["ms"] = float(t[2])
    if rank != 0:
        if args.groth16_log_n:
            bases.close()
            del d_scalars, h_scalars
            torch.cuda.empty_cache()
            dist.broadcast(torch.ones(1, device=dev), src=0)
            try:
                groth16_section(be, args, dist, dev)
            except Exception as e:
                print("rank %d groth16 section failed: %r" % (rank, e), file=sys.stderr)
        dist.destroy_process_group()
        return

    total_points = float(n) * world * args.steps
    value = total_points / (ms_total * 1e-3)
    e2e_value = total_points / (e2e_ms * 1e-3)

    # roofline of the dominant kernel (MsmAccumK): algorithmic IMAD = entries * madd * Fp-mul cost
    info = (C.c_int * 4)()
    be._check(lib.ps_bases_info(bases.handle, info))
    c_bits, W = info[0], info[1]
    madds = float(n) * W
    imad_alg = madds * FP_MUL_PER_MADD * IMAD_PER_FP_MUL
    accum_s = phase["accumulate_ms"] * 1e-3
    achieved = imad_alg / accum_s if accum_s > 0 else 0.0
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
            tr = json.load(f)
        key = "g1_msm_2^%d" % args.log_n
        traffic = tr.get(key, {}).get("MsmAccumK_dram_bytes_per_launch")
    except (OSError, ValueError):
        pass
    gather_bytes = madds * 96.0 + madds * 4.0
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u32", "data": "synthetic",
        "config": {"workload": "G1 MSM 2^%d points/GPU, random 254-bit scalars, resident bases" % args.log_n,
                   "window_bits": c_bits, "windows": W, "precomputed_tables": args.tables,
                   "l2": "inputs larger than L2 (scalars %d MB, bases >= %d MB per table)" % (n * 32 >> 20, n * 96 >> 20),
                   "parallelism": "point-range shard x%d, all-gather of 192 B partials" % world if world > 1 else "single GPU",
                   "bases_build_s": round(t_bases, 2)},
        "clocks": clocks,
        "parity": dict(parity, how="every timed configuration checked against (sum k_i s_i mod r)*G from the oracle "
                                   "(C dot product over Fr + one scalar multiplication), untimed"),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": n * 32 * world, "d2h_bytes_per_step": 48,
                "ms_per_step": e2e_ms / args.steps},
        "gpu_launches": int(launches),
        "phases_ms_last_step": phase,
        "roofline": {"bound": "imad", "kernel": "MsmAccumK<Fp>", "achieved": achieved / 1e12, "peak": imad_peak / 1e12,
                     "unit": "TIMAD/s", "frac": achieved / imad_peak if imad_peak else None, "traffic": traffic,
                     "how": "algorithmic IMAD = n*W mixed adds x 10 Fp mul x 600 IMAD (IMAD.LO/HI pair accounting) / "
                            "CUDA-event time of the kernel; peak = IMAD issue rate measured live on this GPU "
                            "(ps_bench_intpipe, 'of measured'; MEASURED_PEAKS.json has no integer figure)",
                     "whole_msm_frac": (imad_alg / (phase["total_ms"] * 1e-3)) / imad_peak if phase["total_ms"] else None,
                     "fp_mul_per_s_microbench": fpmul_rate,
                     "hbm": {"achieved_gbs": gather_bytes / accum_s / 1e9 if accum_s else None,
                             "what": "base gather 96 B + entry 4 B per mixed add"}},
    }
    if g2:
        n2 = 1 << args.g2_log_n
        acc2 = g2["phases"]["accumulate_ms"] * 1e-3
        imad2 = float(n2) * g2["W"] * 28.0 * IMAD_PER_FP_MUL     # G2 mixed add: 8 Fp2 mul + 2 Fp2 sqr = 28 Fp mul
        line["g2_msm"] = {"metric": "g2_msm_points_per_s", "value": float(n2) * world * args.steps / (g2["ms"] * 1e-3),
                          "unit": UNIT, "points_per_gpu": n2, "ms_per_step": g2["ms"] / args.steps, "window_bits": g2["c"],
                          "windows": g2["W"], "phases_ms_last_step": g2["phases"],
                          "roofline_frac_accumulate_kernel": (imad2 / acc2) / imad_peak if acc2 > 0 and imad_peak else None}
    if not args.no_cpu_baseline and world == 1:
        line["cpu_baseline"] = cpu_baseline(args)
    if args.groth16_log_n:
        bases.close()
        del d_scalars, h_scalars
        torch.cuda.empty_cache()
        if world > 1:
            dist.broadcast(torch.ones(1, device=dev), src=0)   # release the other ranks into the section
        try:
            line["groth16"] = groth16_section(be, args, dist if world > 1 else None, dev)
        except Exception as e:  # the headline metric must still be reported
            line["groth16"] = {"error": repr(e)}
    if world == 1 and not args.no_small_configs:
        try:
            cb = line.get("cpu_baseline") or {}
            line["small_configs"] = small_configs_section(be, cb.get("value"))
        except Exception as e:
            line["small_configs"] = {"error": repr(e)}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    if not parity["all_ranks_ok"]:
        print("bench.py: PARITY MISMATCH %r -- the numbers above are invalid" % parity, file=sys.stderr)
        sys.exit(3)


def groth16_section(be, args, dist=None, dev="cuda"):
    """BASELINE configs[2]/[4]: full Groth16 prove (sparse R1CS -> interpolation on {1..n} -> quotient ->
    3 MSMs) on a synthetic circuit of 2^k constraints; parity = exponent-level recomputation from the
    toxic waste (groth16_test.go:32-107 at scale).  Timed through the reference-facing call with the
    witness in host memory (wire format) and the proof bytes back on the host.  With several ranks
    the three MSMs are sharded by point range (playsnark_b200/dist.py); every rank holds the key."""
    import torch
    import playsnark_b200 as ps
    from playsnark_b200 import dist as D
    from oracle import ps_oracle as O
    from tests import helpers as H
    world = dist.get_world_size() if dist is not None else 1
    rank = dist.get_rank() if dist is not None else 0
    k = args.groth16_log_n
    n = 1 << k
    sq, wit = H.sparse_circuit(n, 7, n // 2)
    tr, tw = H.sparse_groth16_setup(be, sq, 7)
    smp = O.Sampler(99)
    r, s = smp.fr(), smp.fr()
    wb = b"".join(v.to_bytes(32, "big") for v in wit)
    t0 = time.perf_counter()
    sq._resident(be)         # every rank folds a subtree of one aggregate polynomial
    D.load_key_sharded(be, tr, world); be.sync()
    wb = ps.HostBuffer(be, wb)   # the caller marshals its witness into page-locked memory (ps_host_alloc)
    t_load = time.perf_counter() - t0

    def prove():
        if world == 1:
            p = ps.Groth16Prove(tr, sq, wb, r, s, backend=be)
            return p.A, p.B, p.C
        return D.groth16_prove_sharded(be, tr, sq, wb, r, s, dist, dev)

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(2):
        pr = prove()
    reps = max(3, args.steps)
    l0 = be.launch_count()
    sync_all()
    t0 = time.perf_counter()
    for _ in range(reps):
        pr = prove()
    sync_all()
    wall = (time.perf_counter() - t0) / reps
    launches = (be.launch_count() - l0) // reps
    if rank != 0:
        return None
    out = {"constraints": n, "variables": sq.nbVars, "nio_points": sq.nbIO, "n_gpus": world, "proof_ms_e2e": wall * 1e3,
           "proofs_per_s_e2e": 1.0 / wall, "gpu_launches_per_proof_rank0": int(launches),
           "h2d_bytes_per_proof": wb.nbytes + 64, "d2h_bytes_per_proof": 192, "key_and_qap_load_s": round(t_load, 2)}
    if world == 1:
        out["device_ms"] = be.prove_timing()
    A, B, Cc, _ = H.sparse_groth16_expected(sq, wit, tw, r, s)
    out["parity"] = ("A, B, C equal the exponent-level recomputation from the toxic waste"
                     if tuple(pr) == (A, B, Cc) else "MISMATCH")
    out["cpu_reference"] = ("does not finish at this size: ToQAP is O(m n^3) field multiplications and the dense QAP "
                            "would need 3*m*n*32 bytes (BASELINE.md section 2)")
    return out


def small_configs_section(be, cpu_points_per_s):
    """BASELINE configs[1] and [2]: the 2^10 repeated-squaring circuit with its dense QAP (Groth16 and
    PHGR13, checked against the oracle inside tests.parity_cases.config_c2) and a 2^16-constraint sparse
    circuit (Groth16, checked in the exponent); end-to-end latency through the reference-facing calls.
    The reference's own cost at 2^10 is an ESTIMATE from its operation counts (3*m*n scalar
    multiplications in sumBlind, groth16.go:134-141) and the measured rate of the CPU port."""
    import playsnark_b200 as ps
    from oracle import ps_oracle as O
    from tests import helpers as H, parity_cases as P
    out = {}
    t = {}
    P.config_c2(be, 1 << 10, timings=t)
    m, n = t["variables"], t["gates"]
    t["parity"] = "h, A, B, C and the 8 PHGR13 elements equal the oracle's (tests.parity_cases.config_c2)"
    if cpu_points_per_s:
        t["reference_cpu_groth16_estimate_s"] = round((3 * m * n + 2 * n) / cpu_points_per_s, 1)
        t["reference_cpu_estimate_how"] = ("(3*m*n + 2n) bit-serial scalar multiplications (sumBlind + BlindEval) / measured "
                                          "CPU-port rate; Div2's n^3/2 field operations not included")
    out["c2_dense_2p10"] = t
    k = 16
    nn = 1 << k
    sq, wit = H.sparse_circuit(nn, 7, nn // 2)
    tr, tw = H.sparse_groth16_setup(be, sq, 7)
    smp = O.Sampler(99)
    r, s = smp.fr(), smp.fr()
    wb = ps.HostBuffer(be, b"".join(v.to_bytes(32, "big") for v in wit))
    pr = ps.Groth16Prove(tr, sq, wb, r, s, backend=be)
    best = 1e9
    for _ in range(5):
        t0 = time.perf_counter(); pr = ps.Groth16Prove(tr, sq, wb, r, s, backend=be); best = min(best, time.perf_counter() - t0)
    A, B, Cc, _ = H.sparse_groth16_expected(sq, wit, tw, r, s)
    out["c3_sparse_2p16"] = {"groth16_prove_ms": best * 1e3, "device_ms": be.prove_timing(),
                             "parity": "A, B, C equal the exponent-level recomputation" if (pr.A, pr.B, pr.C) == (A, B, Cc) else "MISMATCH"}
    return out


def cpu_baseline(args):
    """the oracle's C port of BlindEval on this box's host cores, bounded sample (about 10-20 s)"""
    from oracle import c_oracle as CO, ps_oracle as O
    sample = 1 << args.ref_log_n
    sc = random_scalars_be(sample, 3)
    p0 = O.g1_mul(0x1234567 | 1)
    pts, acc = [], None
    for _ in range(sample):
        acc = O.g1_add(acc, p0)
        pts.append(acc)
    pbytes = b"".join(O.g1_affine_bytes(p) for p in pts)
    CO.blind_eval_g1(pbytes[:96 * 32], sc.tobytes()[:32 * 32])
    reps, t0 = 0, time.perf_counter()
    while True:
        CO.blind_eval_g1(pbytes, sc.tobytes())
        reps += 1
        if time.perf_counter() - t0 > 10.0 or reps >= 20:
            break
    dt = time.perf_counter() - t0
    return {"value": sample * reps / dt, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": "%d x Poly.BlindEval over 2^%d of the workload's points (oracle/ps_oracle.c, bit-serial scalar-mul "
                      "per term as algebra.go:348-359; single-threaded like the reference; %d host cores present)"
                      % (reps, args.ref_log_n, os.cpu_count() or 0)}


if __name__ == "__main__":
    main()
