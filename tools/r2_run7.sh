#!/bin/bash
# 2-GPU pass: DDP / SyncBN tests, weak-scaling bench with and without SyncBN
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ddp_gpu.py -q -m gpu --timeout 800 > gpurun_out/pytest_r2_ddp.log 2>&1
grep -n "^E   \|FAILED\|passed\|failed\|Error" gpurun_out/pytest_r2_ddp.log | cut -c1-500 | head -20
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_r2_n2.json 2> gpurun_out/bench_r2_n2.err; tail -2 gpurun_out/bench_r2_n2.err | cut -c1-300; cut -c1-260 gpurun_out/bench_r2_n2.json
FPD_BN_SYNC=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29712 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_r2_n2_syncbn.json 2> gpurun_out/bench_r2_n2_syncbn.err; tail -3 gpurun_out/bench_r2_n2_syncbn.err | cut -c1-400; cut -c1-260 gpurun_out/bench_r2_n2_syncbn.json
