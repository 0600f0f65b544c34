#!/bin/bash
# Round-end style check on one B200: GPU parity suite, smoke(), the default bench line, the reference arm.
# Run as: gpurun --timeout 1500 -- 'bash tools/gpu_round.sh [tag]'
tag=${1:-round}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -x > gpurun_out/tests_$tag.log 2>&1
echo "tests rc=$?"; tail -3 gpurun_out/tests_$tag.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$tag.log 2>&1
echo "smoke rc=$?"; tail -1 gpurun_out/smoke_$tag.log
timeout 600 python bench.py > gpurun_out/bench_$tag.json 2>gpurun_out/bench_$tag.err
echo "bench rc=$?"; cut -c1-600 gpurun_out/bench_$tag.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$tag.json 2>gpurun_out/bench_ref_$tag.err
echo "ref rc=$?"; cut -c1-300 gpurun_out/bench_ref_$tag.json
