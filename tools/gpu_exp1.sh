#!/bin/bash
# experiment 1: GEMM probe + deferred-LN parity + A/B timing
mkdir -p gpurun_out
timeout 300 python tools/gemm_probe.py > gpurun_out/gemm_probe.log 2>&1
echo "probe rc=$?"; tail -60 gpurun_out/gemm_probe.log
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 300 -x -k "gemm or epilogue or modernbert" > gpurun_out/tests_exp1.log 2>&1
echo "tests rc=$?"; tail -15 gpurun_out/tests_exp1.log
for dl in 0 1; do
  VRAG_DEFERRED_LN=$dl timeout 300 python bench.py --steps 2 --warmup 1 --seqs-per-step 1024 --no-cpu-baseline --no-secondary > gpurun_out/bench_dl$dl.json 2>gpurun_out/bench_dl$dl.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_dl$dl.json"))
    print("deferred_ln=$dl", round(d["value"], 1), round(d["roofline"]["achieved"], 1), d["roofline"].get("share_of_step"), d["clocks"])
except Exception as e:
    print("bench failed", e)
PY
done
