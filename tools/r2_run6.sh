#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_hrnet_gpu.py tests/test_hourglass_gpu.py -q -m gpu --timeout 800 > gpurun_out/pytest_r2e.log 2>&1
grep -n "^E   \|FAILED\|passed\|failed" gpurun_out/pytest_r2e.log | cut -c1-400 | head -20
timeout 400 python bench.py --config hrnet_fpd --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | cut -c1-220
FPD_WGRAD3_PAIR=0 timeout 400 python bench.py --config hrnet_fpd --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | cut -c1-220
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | cut -c1-220
FPD_WGRAD3_PAIR=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | cut -c1-220
