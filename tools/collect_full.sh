#!/bin/bash
# `ncu --set full` captures of the three dominant kernels on ONE B200 (each after its command ran clean in
# collect_evidence.sh); summaries are written on the box, only the G1 report travels back.  tag = $1
tag=${1:-r02}
set -x
NCU="ncu --set full --clock-control none --import-source on --kernel-name-base demangled"
timeout 300 $NCU -k regex:MsmAccumK -s 1 -c 1 -f -o gpurun_out/prof_accum_g1_${tag} python tools/profile_target.py 24 g1 > gpurun_out/ncu_full_g1_${tag}.log 2>&1; echo "full g1 rc=$?"
timeout 300 $NCU -k regex:MsmAccumK -s 1 -c 1 -f -o gpurun_out/prof_accum_g2_${tag} python tools/profile_target.py 20 g2 > gpurun_out/ncu_full_g2_${tag}.log 2>&1; echo "full g2 rc=$?"
timeout 300 $NCU -k regex:NttDifK -s 400 -c 1 -f -o gpurun_out/prof_nttdif_${tag} python tools/profile_groth16.py 20 > gpurun_out/ncu_full_ntt_${tag}.log 2>&1; echo "full ntt rc=$?"
python tools/summarize_profiles.py full gpurun_out/prof_accum_g1_${tag}.ncu-rep gpurun_out/${tag}_MsmAccumK_g1_2p24_full.txt "g1_msm_2^24"
python tools/summarize_profiles.py full gpurun_out/prof_accum_g2_${tag}.ncu-rep gpurun_out/${tag}_MsmAccumK_g2_2p20_full.txt "g2_msm_2^20"
python tools/summarize_profiles.py full gpurun_out/prof_nttdif_${tag}.ncu-rep gpurun_out/${tag}_NttDifK_groth16_2p20_full.txt "ntt_pass_groth16_2^20"
cp profiles/roofline_traffic.json gpurun_out/roofline_traffic_${tag}.json
ncu -i gpurun_out/prof_accum_g1_${tag}.ncu-rep --page source --csv > gpurun_out/${tag}_MsmAccumK_g1_source.csv 2>/dev/null
rm -f gpurun_out/prof_accum_g2_${tag}.ncu-rep gpurun_out/prof_nttdif_${tag}.ncu-rep
ls -la gpurun_out | tail -12
