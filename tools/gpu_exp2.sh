#!/bin/bash
# attention experiment: kernel tests + short bench with per-class shares
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 300 -x -k "attention or modernbert or splade" > gpurun_out/tests_exp2.log 2>&1
echo "tests rc=$?"; tail -5 gpurun_out/tests_exp2.log
timeout 300 python bench.py --steps 2 --warmup 1 --seqs-per-step 1024 --no-cpu-baseline --no-secondary > gpurun_out/bench_exp2.json 2>gpurun_out/bench_exp2.err
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_exp2.json"))
    print(round(d["value"], 1), "ms/step", round(d["ms_per_step"], 2), round(d["roofline"]["achieved"], 1), d["roofline"].get("share_of_step"), d["clocks"])
except Exception as e:
    print("bench failed", e)
PY
