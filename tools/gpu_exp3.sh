#!/bin/bash
# development cycle: GPU parity suite, GEMM probe table, short bench
tag=${1:-exp}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -x > gpurun_out/tests_$tag.log 2>&1
echo "tests rc=$?"; tail -5 gpurun_out/tests_$tag.log
timeout 300 python tools/gemm_probe.py > gpurun_out/gemm_probe_$tag.log 2>&1; tail -16 gpurun_out/gemm_probe_$tag.log
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/bench_$tag.json 2>gpurun_out/bench_$tag.err
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_$tag.json"))
    print(round(d["value"], 1), round(d["roofline"]["achieved"], 1), d["roofline"].get("share_of_step"), d["clocks"])
except Exception as e:
    print("bench failed", e)
PY
