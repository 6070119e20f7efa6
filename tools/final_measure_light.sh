#!/bin/bash
# Lighter sibling of final_measure.sh (no ncu --set full captures): tests, smoke, every bench line, CUPTI step profile,
# timeline, ncu launch list -- most important first, so a cut-off run still leaves the essentials in gpurun_out/.
set -u
mkdir -p gpurun_out
TAG=${1:-r2f}
rm -f gpurun_out/parity_report.jsonl
timeout 300 python -m pytest tests -q -m gpu --timeout 200 > gpurun_out/pytest_${TAG}.log 2>&1
grep -E "passed|failed|FAILED|^E  " gpurun_out/pytest_${TAG}.log | tail -5 | cut -c1-300
timeout 100 python __graft_entry__.py smoke > gpurun_out/smoke_${TAG}.log 2>&1; tail -1 gpurun_out/smoke_${TAG}.log | cut -c1-200
for c in hg_fpd hg_infer hrnet_fpd res50_mse hg_mse_s1; do
  timeout 150 python bench.py --config $c --steps 20 --warmup 5 > gpurun_out/bench_${TAG}_$c.json 2> gpurun_out/bench_${TAG}_$c.err
  python - "$c" gpurun_out/bench_${TAG}_$c.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    print(sys.argv[1], "value", round(d["value"], 1), "ms", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"], 1),
          "roof", d.get("roofline", {}).get("frac"), "cpu", (d.get("cpu_baseline") or {}).get("value"))
except Exception as exc:
    print(sys.argv[1], "FAILED", exc)
PY
done
timeout 120 python bench.py --config res50_mse --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_${TAG}_res50_mse.json 2>/dev/null
timeout 120 python tools/timeline_step.py --tag ${TAG} > gpurun_out/timeline_${TAG}.log 2>&1
head -6 gpurun_out/timeline_${TAG}.txt 2>/dev/null | cut -c1-160
FPD_OVERLAP_TEACHER=0 FPD_FORK_UP1=0 FPD_WGRAD_STREAM=0 timeout 150 python tools/profile_step.py > gpurun_out/profile_stdout.log 2>&1
cp gpurun_out/profile_step.txt gpurun_out/step_cupti_${TAG}.txt 2>/dev/null
FPD_OVERLAP_TEACHER=0 FPD_FORK_UP1=0 FPD_WGRAD_STREAM=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches_${TAG}.csv python tools/profile_step.py --ncu > gpurun_out/ncu_launch.log 2>&1
ls -la gpurun_out/launches_${TAG}.csv 2>/dev/null | awk '{print $5, $9}'
