#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --timeout 600 > gpurun_out/pytest_h.log 2>&1
tail -6 gpurun_out/pytest_h.log | cut -c1-400
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_h.json 2> gpurun_out/bench_h.err
cut -c1-330 gpurun_out/bench_h.json; tail -3 gpurun_out/bench_h.err
FPD_OVERLAP_TEACHER=0 timeout 600 python tools/profile_step.py > gpurun_out/profile_stdout.log 2>&1
head -22 gpurun_out/profile_step.txt | cut -c1-110
timeout 300 python tools/diag_conv_h.py --stalls > gpurun_out/stalls.log 2>&1; grep "dbg=0 " gpurun_out/stalls.log | grep "1)" | cut -c1-330
