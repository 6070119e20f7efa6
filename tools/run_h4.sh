#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --timeout 600 > gpurun_out/pytest_h.log 2>&1
tail -5 gpurun_out/pytest_h.log | cut -c1-400
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_h.json 2> gpurun_out/bench_h.err
cut -c1-330 gpurun_out/bench_h.json; tail -3 gpurun_out/bench_h.err
FPD_OVERLAP_TEACHER=0 timeout 600 python tools/profile_step.py > gpurun_out/profile_stdout.log 2>&1
head -24 gpurun_out/profile_step.txt | cut -c1-110
