#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python tools/diag_conv_h.py > gpurun_out/diag_conv_h.log 2>&1
echo "diag rc=$?"; grep -c "^ok" gpurun_out/diag_conv_h.log; grep -E "^FAIL|^EXC|failures|^SKIP" gpurun_out/diag_conv_h.log | head -12
grep "4, 4" gpurun_out/diag_conv_h.log | tail -3
timeout 900 python -m pytest tests -q -m gpu --timeout 600 > gpurun_out/pytest_h.log 2>&1
tail -4 gpurun_out/pytest_h.log | cut -c1-300
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_h.json 2> gpurun_out/bench_h.err
cut -c1-200 gpurun_out/bench_h.json; tail -3 gpurun_out/bench_h.err
FPD_PIPELINE_TEACHER=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-200
