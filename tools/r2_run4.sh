#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu --timeout 800 > gpurun_out/pytest_r2c.log 2>&1
grep -n "^E   \|FAILED\|passed\|failed" gpurun_out/pytest_r2c.log | cut -c1-500 | head -20
timeout 400 python bench.py --config hrnet_fpd --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2c_hrnet.json 2>gpurun_out/bench_r2c_hrnet.err; tail -2 gpurun_out/bench_r2c_hrnet.err | cut -c1-300; cut -c1-300 gpurun_out/bench_r2c_hrnet.json
timeout 400 python tools/timeline_step.py --config hrnet_fpd --batch 0 --tag r2c_hrnet > gpurun_out/timeline_r2c_hrnet.log 2>&1; tail -38 gpurun_out/timeline_r2c_hrnet.log | cut -c1-150
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | cut -c1-200
FPD_CONV_STATS=3x3 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | cut -c1-200
