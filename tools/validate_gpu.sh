#!/bin/bash
# validation: GPU tests + smoke + bench with the default switches
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --timeout 600 > gpurun_out/pytest_h.log 2>&1
tail -3 gpurun_out/pytest_h.log | cut -c1-300
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_h.json 2> gpurun_out/bench_h.err
cut -c1-200 gpurun_out/bench_h.json; tail -3 gpurun_out/bench_h.err
