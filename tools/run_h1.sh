#!/bin/bash
# diag + tests + bench of the generation-5 conv kernel (one gpurun call)
set -u
mkdir -p gpurun_out
timeout 600 python tools/diag_conv_h.py > gpurun_out/diag_conv_h.log 2>&1
echo "diag rc=$?"
grep -c "^ok" gpurun_out/diag_conv_h.log; grep -E "^FAIL|^EXC|failures" gpurun_out/diag_conv_h.log | head -40
sed -n '/B,H,W,Cin,Cout,k/,$p' gpurun_out/diag_conv_h.log
timeout 900 python -m pytest tests -q -m gpu --timeout 600 -x -k "not conv2d_tc_h" > gpurun_out/pytest_h.log 2>&1
tail -5 gpurun_out/pytest_h.log | cut -c1-400
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_h.json 2> gpurun_out/bench_h.err
cut -c1-700 gpurun_out/bench_h.json; tail -3 gpurun_out/bench_h.err
FPD_CONV_F16=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_h_tf32.json 2> gpurun_out/bench_h_tf32.err
cut -c1-300 gpurun_out/bench_h_tf32.json; tail -3 gpurun_out/bench_h_tf32.err
