#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu --timeout 800 -x > gpurun_out/pytest_r2b.log 2>&1
grep -n "^E   \|FAILED\|passed\|failed" gpurun_out/pytest_r2b.log | cut -c1-500 | head -20
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_r2b_on.json 2>gpurun_out/bench_r2b_on.err; cut -c1-400 gpurun_out/bench_r2b_on.json
FPD_CONV_STATS=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | cut -c1-200
FPD_BN_APPLY_SUM=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | cut -c1-200
FPD_CONV_STATS=0 FPD_BN_APPLY_SUM=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | cut -c1-200
timeout 300 python tools/timeline_step.py --tag r2b > gpurun_out/timeline_r2b.log 2>&1; tail -32 gpurun_out/timeline_r2b.log | cut -c1-150
timeout 400 python tools/timeline_step.py --config hrnet_fpd --batch 0 --tag r2b_hrnet > gpurun_out/timeline_r2b_hrnet.log 2>&1; tail -36 gpurun_out/timeline_r2b_hrnet.log | cut -c1-150
timeout 600 python tools/cpu_thread_sweep.py small > gpurun_out/cpu_thread_sweep_small.txt 2>&1; cat gpurun_out/cpu_thread_sweep_small.txt
