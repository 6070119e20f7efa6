#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/parity_report.jsonl
timeout 900 python -m pytest tests/test_parity_bench_config_gpu.py tests/test_hrnet_gpu.py -q -m gpu --timeout 800 > gpurun_out/pytest_new2.log 2>&1
grep -n "^E   \|FAILED\|passed\|failed" gpurun_out/pytest_new2.log | cut -c1-600 | head -30
timeout 400 python bench.py --config hrnet_fpd --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2a_hrnet_fpd.json 2> gpurun_out/bench_r2a_hrnet_fpd.err
tail -3 gpurun_out/bench_r2a_hrnet_fpd.err | cut -c1-300; cut -c1-1500 gpurun_out/bench_r2a_hrnet_fpd.json
timeout 900 python tools/cpu_thread_sweep.py > gpurun_out/cpu_thread_sweep.txt 2>&1; cat gpurun_out/cpu_thread_sweep.txt
