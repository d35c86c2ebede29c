#!/bin/bash
# End-of-round evidence on ONE B200 (run through gpurun from the repo root): GPU parity suite, bench line, ncu launch lists of
# the bench's headline section and of the device setup + two Groth16 proofs at 2^20, `ncu --set full` of the G1
# accumulation kernel (the one kernel that changed since collect_evidence.sh / collect_full.sh ran).  Every ncu pass runs
# only after the same command has exited 0 without ncu.  Outputs under gpurun_out/ (tag = $1).
tag=${1:-r02d}
set -x
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_${tag}_1gpu.json 2> gpurun_out/bench_${tag}_1gpu.err; echo "bench rc=$?"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_${tag}_bench.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --g2-log-n "" --groth16-log-n 0 --phgr13-log-n "" --no-small-configs > gpurun_out/ncu_list_${tag}.log 2>&1; echo "list rc=$?"
timeout 200 python tools/profile_groth16.py 20 > gpurun_out/g16_plain_${tag}.log 2>&1 && \
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_${tag}_g16.csv \
  python tools/profile_groth16.py 20 > gpurun_out/g16_ncu_${tag}.log 2>&1; echo "g16 list rc=$?"
timeout 120 python tools/profile_target.py 24 g1 > gpurun_out/target_g1_${tag}.log 2>&1 && \
timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:MsmAccumK -s 1 -c 1 -f -o gpurun_out/prof_accum_g1_${tag} \
  python tools/profile_target.py 24 g1 > gpurun_out/ncu_full_g1_${tag}.log 2>&1; echo "full g1 rc=$?"
python tools/summarize_profiles.py full gpurun_out/prof_accum_g1_${tag}.ncu-rep gpurun_out/${tag}_MsmAccumK_g1_2p24_full.txt "g1_msm_2^24"
cp profiles/roofline_traffic.json gpurun_out/roofline_traffic_${tag}.json
timeout 100 python tools/setup_probe.py 20 2>&1 | tail -1
ls -la gpurun_out | tail -8
