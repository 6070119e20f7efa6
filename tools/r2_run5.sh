#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_hrnet_gpu.py tests/test_parity_bench_config_gpu.py tests/test_ops_gpu.py -q -m gpu --timeout 800 > gpurun_out/pytest_r2d.log 2>&1
grep -n "^E   \|FAILED\|passed\|failed" gpurun_out/pytest_r2d.log | cut -c1-500 | head -20
timeout 400 python bench.py --config hrnet_fpd --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2d_hrnet.json 2>gpurun_out/bench_r2d_hrnet.err; tail -2 gpurun_out/bench_r2d_hrnet.err | cut -c1-300; cut -c1-300 gpurun_out/bench_r2d_hrnet.json
timeout 400 python tools/timeline_step.py --config hrnet_fpd --batch 0 --tag r2d_hrnet > gpurun_out/timeline_r2d_hrnet.log 2>&1; tail -36 gpurun_out/timeline_r2d_hrnet.log | cut -c1-150
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | cut -c1-200
timeout 300 python bench.py --config diag_student --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | cut -c1-200
timeout 300 python bench.py --config diag_teacher --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | cut -c1-200
