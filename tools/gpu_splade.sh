#!/bin/bash
# SPLADE / BERT development check: GEMM epilogue self tests, parity tests of the BERT-MLM / SPLADE / dense-provider
# paths, the encode timing and its launch list
tag=${1:-dev}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -x -k "fused_epilogues or residual_stats or splade or provider or end_to_end" > gpurun_out/tests_splade_$tag.log 2>&1
echo "tests rc=$?"; tail -6 gpurun_out/tests_splade_$tag.log
timeout 300 python tools/splade_bench.py > gpurun_out/splade_$tag.json 2>gpurun_out/splade_$tag.err
head -5 gpurun_out/splade_$tag.json; tail -2 gpurun_out/splade_$tag.err
SPLADE_CHUNKS=256 timeout 300 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__cycles_elapsed.max --clock-control none --csv --log-file gpurun_out/splade_launches_$tag.csv python tools/splade_bench.py > gpurun_out/splade_ncu.log 2>&1
python tools/summarize_launches.py gpurun_out/splade_launches_$tag.csv
