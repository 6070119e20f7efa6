#!/bin/bash
# One gpurun call that refreshes every number / artefact quoted in DESIGN.md and profiles/ (run from the repo root).
set -u
mkdir -p gpurun_out
TAG=${1:-r1}
timeout 1200 python -m pytest tests -q -m gpu --timeout 900 > gpurun_out/pytest_all.log 2>&1
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_${TAG}.json 2>/dev/null
FPD_OVERLAP_TEACHER=0 FPD_FORK_UP1=0 FPD_WGRAD_STREAM=0 timeout 600 python tools/profile_step.py > gpurun_out/profile_stdout.log 2>&1
cp gpurun_out/profile_step.txt gpurun_out/step_cupti_${TAG}.txt
FPD_OVERLAP_TEACHER=0 FPD_FORK_UP1=0 FPD_WGRAD_STREAM=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches_${TAG}.csv python tools/profile_step.py --ncu > gpurun_out/ncu_launch.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc_h_kernel -s 2 -c 1 -f \
    -o gpurun_out/prof_conv_h_3x3 python tools/profile_kernel.py conv_h_f16 32 64 64 128 128 3 > gpurun_out/ncu_conv_h3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc_h_kernel -s 2 -c 1 -f \
    -o gpurun_out/prof_conv_h_1x1 python tools/profile_kernel.py conv_h_f16 32 64 64 128 256 1 > gpurun_out/ncu_conv_h1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wgrad_tc3_kernel -s 2 -c 1 -f \
    -o gpurun_out/prof_wgrad_tc3 python tools/profile_kernel.py wgrad_fused 32 64 64 64 64 3 > gpurun_out/ncu_wgrad.log 2>&1
grep -E "passed|failed|FAILED|rror" gpurun_out/pytest_all.log | tail -5 | cut -c1-300
tail -1 gpurun_out/smoke.log
cat gpurun_out/bench_${TAG}.json | cut -c1-2500
cat gpurun_out/bench_ref_${TAG}.json | cut -c1-600
head -12 gpurun_out/step_cupti_${TAG}.txt
