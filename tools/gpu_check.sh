#!/bin/bash
# GPU-box check used during development: full GPU parity suite, a short bench, and an ncu launch list.
# Run as: gpurun --timeout 1500 -- 'bash tools/gpu_check.sh [tag]'
tag=${1:-dev}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -x > gpurun_out/tests_$tag.log 2>&1
echo "tests rc=$?"; tail -5 gpurun_out/tests_$tag.log
timeout 300 python bench.py --steps 2 --warmup 1 --seqs-per-step 1024 --no-cpu-baseline --no-secondary > gpurun_out/bench_$tag.json 2>gpurun_out/bench_$tag.err
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_$tag.json"))
    print(round(d["value"], 1), round(d["roofline"]["achieved"], 1), d["roofline"].get("share_of_step"), d["clocks"])
except Exception as e:
    print("bench failed", e)
PY
timeout 300 ncu --metrics gpu__time_duration.sum,sm__cycles_elapsed.max,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -c 700 --csv --log-file gpurun_out/launches_$tag.csv python bench.py --steps 1 --warmup 1 --seqs-per-step 128 --max-tokens 65536 --no-cpu-baseline --no-secondary > gpurun_out/ncu_$tag.log 2>&1
echo "ncu rc=$?"
