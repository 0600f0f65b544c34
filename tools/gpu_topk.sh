#!/bin/bash
# top-k development check: parity tests of the vector-store path + the secondary bench (dense 1M x 768)
tag=${1:-dev}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout 300 -x -k "topk or store or sharded or dense or sparse" > gpurun_out/tests_topk_$tag.log 2>&1
echo "tests rc=$?"; tail -5 gpurun_out/tests_topk_$tag.log
timeout 600 python tools/topk_bench.py > gpurun_out/quick_topk_$tag.json 2>gpurun_out/quick_topk_$tag.err
echo "quickbench rc=$?"; tail -c 1500 gpurun_out/quick_topk_$tag.json; tail -3 gpurun_out/quick_topk_$tag.err
# launch list of one 16-query and one 1-query search (scan / select / finish kernels)
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/topk_launches_$tag.csv python tools/topk_ncu.py > gpurun_out/topk_ncu.log 2>&1
