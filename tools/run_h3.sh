#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python tools/diag_conv_h.py > gpurun_out/diag_conv_h.log 2>&1
echo "diag rc=$?"
grep -c "^ok" gpurun_out/diag_conv_h.log; grep -E "^FAIL|^EXC|failures" gpurun_out/diag_conv_h.log | head -40
sed -n '/B,H,W,Cin,Cout,k/,$p' gpurun_out/diag_conv_h.log
timeout 300 python tools/diag_conv_h.py --stalls > gpurun_out/stalls.log 2>&1; cat gpurun_out/stalls.log | tail -45
