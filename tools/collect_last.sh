#!/bin/bash
# Last GPU call of round 2 (one B200, through gpurun from the repo root), most important first:
#  1. the GPU parity suite on the shipped library;
#  2. MSM sweep G1 and G2, 2^12..2^24, every random row checked against the oracle (SURVEY 8 d2 / configs[3]);
#  3. A/B of NTT launch shapes (resident blocks per SM held at 7 x 64 / 14 x 32 threads by a shared-memory pad, so that
#     the 2^18-thread passes fill whole waves), quotient time at 2^16 and 2^20 constraints, proofs compared;
#  4. the proof-path tests against the most promising variant.
set -x
timeout 170 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 200 python tools/sweep_msm.py 24 24 r02 > gpurun_out/sweep_r02.log 2>&1; echo "sweep rc=$?"
timeout 150 python tools/ab_ntt.py > gpurun_out/ab_ntt_r02e.log 2>&1; echo "ab rc=$?"; cp gpurun_out/ab_ntt.json gpurun_out/ab_ntt_r02e.json
PLAYSNARK_B200_LIB=$PWD/playsnark_b200/variants/lib_ntt_b64p7.so timeout 100 python -m pytest tests/test_gpu_parity.py -x -q -m gpu \
  -k "ntt or quotient or groth16 or phgr13 or sparse or sharded or setups" 2>&1 | tail -3
tail -4 gpurun_out/ab_ntt_r02e.log
