#!/bin/bash
# pass-size sweep: does keeping a pass's activations inside the 126 MB L2 beat larger GEMM grids?
mkdir -p gpurun_out
for mt in 8192 12288 16384 24576 32768 37888 65536 131072; do
  timeout 200 python bench.py --steps 2 --warmup 1 --seqs-per-step 1776 --max-tokens $mt --no-cpu-baseline --no-secondary > gpurun_out/sweep_$mt.json 2>gpurun_out/sweep_$mt.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/sweep_$mt.json"))
    print($mt, round(d["value"],1), round(d["roofline"]["achieved"],1), d["roofline"].get("share_of_step"), d["clocks"]["sm_mhz"])
except Exception as e:
    print($mt, "failed", e)
PY
done
