#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --timeout 600 > gpurun_out/pytest_h.log 2>&1
tail -4 gpurun_out/pytest_h.log | cut -c1-300
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_h.json 2> gpurun_out/bench_h.err
cut -c1-250 gpurun_out/bench_h.json; python -c "
import json;d=json.load(open('gpurun_out/bench_h.json'));print(d['roofline']);print(d['e2e'])"; tail -3 gpurun_out/bench_h.err
FPD_OVERLAP_TEACHER=0 FPD_FORK_UP1=0 timeout 600 python tools/profile_step.py > gpurun_out/profile_stdout.log 2>&1
head -22 gpurun_out/profile_step.txt | cut -c1-110
