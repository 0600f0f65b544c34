#!/bin/bash
# ncu evidence for profiles/ (one GPU, never under a timed bench): launch list of one bench pass + --set full captures
# of one layer's four GEMM launches and two attention launches (global + local layer) at the bench's pass size.
tag=${1:-prof}
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 1 --seqs-per-step 256 --max-tokens 131072 --no-cpu-baseline --no-secondary"
timeout 400 ncu --metrics gpu__time_duration.sum,sm__cycles_elapsed.max,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_$tag.csv $B > gpurun_out/ncu_launches_$tag.log 2>&1
echo "launch list rc=$?"
timeout 400 ncu --set full --import-source on --clock-control none --kernel-name regex:gemm_tcgen05_kernel --launch-skip 96 --launch-count 4 -o gpurun_out/gemm_full_$tag -f $B > gpurun_out/ncu_gemm_$tag.log 2>&1
echo "gemm full rc=$?"
timeout 400 ncu --set full --import-source on --clock-control none --kernel-name regex:attention_tc_kernel --launch-skip 24 --launch-count 2 -o gpurun_out/attn_full_$tag -f $B > gpurun_out/ncu_attn_$tag.log 2>&1
echo "attention full rc=$?"
ls -la gpurun_out/*_$tag*
