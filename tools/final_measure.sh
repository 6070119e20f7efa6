#!/bin/bash
# One gpurun call that refreshes every number / artefact quoted in DESIGN.md and profiles/ (run from the repo root).
#   tools/final_measure.sh <tag>            e.g. r2
set -u
mkdir -p gpurun_out
TAG=${1:-r2}
rm -f gpurun_out/parity_report.jsonl
timeout 600 python -m pytest tests -q -m gpu --timeout 400 > gpurun_out/pytest_${TAG}.log 2>&1
grep -E "passed|failed|FAILED|^E  " gpurun_out/pytest_${TAG}.log | tail -8 | cut -c1-300
timeout 200 python __graft_entry__.py smoke > gpurun_out/smoke_${TAG}.log 2>&1; tail -1 gpurun_out/smoke_${TAG}.log
for c in hg_fpd hg_mse_s1 hrnet_fpd hg_infer; do
  timeout 400 python bench.py --config $c --steps 20 --warmup 5 > gpurun_out/bench_${TAG}_$c.json 2> gpurun_out/bench_${TAG}_$c.err
  cut -c1-330 gpurun_out/bench_${TAG}_$c.json
  timeout 300 python bench.py --config $c --impl reference --steps 4 --warmup 1 > gpurun_out/bench_ref_${TAG}_$c.json 2>/dev/null
  cut -c1-200 gpurun_out/bench_ref_${TAG}_$c.json
done
timeout 200 python tools/timeline_step.py --tag ${TAG} > gpurun_out/timeline_${TAG}.log 2>&1
timeout 200 python tools/timeline_step.py --config hrnet_fpd --batch 0 --tag ${TAG}_hrnet > gpurun_out/timeline_${TAG}_hrnet.log 2>&1
FPD_OVERLAP_TEACHER=0 FPD_FORK_UP1=0 FPD_WGRAD_STREAM=0 timeout 300 python tools/profile_step.py > gpurun_out/profile_stdout.log 2>&1
cp gpurun_out/profile_step.txt gpurun_out/step_cupti_${TAG}.txt
FPD_OVERLAP_TEACHER=0 FPD_FORK_UP1=0 FPD_WGRAD_STREAM=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches_${TAG}.csv python tools/profile_step.py --ncu > gpurun_out/ncu_launch.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_tc_h_kernel -s 2 -c 1 -f \
    -o gpurun_out/prof_conv_h_3x3 python tools/profile_kernel.py conv_h_f16 32 64 64 128 128 3 > gpurun_out/ncu_conv_h3.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_tc_h_kernel -s 2 -c 1 -f \
    -o gpurun_out/prof_conv_h_1x1 python tools/profile_kernel.py conv_h_f16 32 64 64 128 256 1 > gpurun_out/ncu_conv_h1.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:wgrad_tc3_kernel -s 2 -c 1 -f \
    -o gpurun_out/prof_wgrad_tc3 python tools/profile_kernel.py wgrad_fused 32 64 64 64 64 3 > gpurun_out/ncu_wgrad.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:wgrad_tc3_kernel -s 2 -c 1 -f \
    -o gpurun_out/prof_wgrad_tc3_pair python tools/profile_kernel.py wgrad_fused 24 64 48 32 32 3 > gpurun_out/ncu_wgrad_pair.log 2>&1
head -14 gpurun_out/step_cupti_${TAG}.txt | cut -c1-150
head -8 gpurun_out/timeline_${TAG}.txt | cut -c1-200
ls -la gpurun_out/*.ncu-rep gpurun_out/launches_${TAG}.csv 2>/dev/null | awk '{print $5, $9}'
