#!/bin/bash
# round-2 measurement pass 1: new parity tests, step timeline, every bench config (short)
mkdir -p gpurun_out; rm -f gpurun_out/parity_report.jsonl
timeout 900 python -m pytest tests/test_parity_bench_config_gpu.py tests/test_decode_tail_gpu.py -q -m gpu --timeout 800 -s > gpurun_out/pytest_new.log 2>&1
tail -40 gpurun_out/pytest_new.log | cut -c1-400
timeout 300 python -m pytest tests -q -m gpu --timeout 600 --deselect tests/test_parity_bench_config_gpu.py --deselect tests/test_decode_tail_gpu.py > gpurun_out/pytest_old.log 2>&1
tail -5 gpurun_out/pytest_old.log | cut -c1-300
timeout 300 python tools/timeline_step.py --tag r2a > gpurun_out/timeline_r2a.log 2>&1; tail -45 gpurun_out/timeline_r2a.log
for c in hg_fpd hg_mse_s1 hrnet_fpd hg_infer; do
  timeout 400 python bench.py --config $c --steps 10 --warmup 3 > gpurun_out/bench_r2a_$c.json 2> gpurun_out/bench_r2a_$c.err
  tail -3 gpurun_out/bench_r2a_$c.err | cut -c1-300; cut -c1-1800 gpurun_out/bench_r2a_$c.json
done
