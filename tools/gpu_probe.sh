#!/bin/bash
# GEMM probe only (timing experiments of the epilogues)
tag=${1:-probe}
mkdir -p gpurun_out
timeout 300 python tools/gemm_probe.py > gpurun_out/gemm_probe_$tag.log 2>&1; tail -24 gpurun_out/gemm_probe_$tag.log
