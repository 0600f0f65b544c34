#!/bin/bash
# Round-2 ncu evidence for profiles/ (one GPU, never under a timed bench).  Launch lists (time, cycles, tensor %, DRAM
# bytes per launch) of one bench pass in fast and in precise mode, of one batched + one 16-query dense search, of one
# sparse search over 1 M documents; `--set full` captures of the kernels DESIGN.md discusses.
tag=${1:-r2}
mkdir -p gpurun_out
M="gpu__time_duration.sum,sm__cycles_elapsed.max,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum"
B="python bench.py --steps 1 --warmup 1 --seqs-per-step 256 --max-tokens 131072 --no-cpu-baseline --no-secondary --no-precise"
timeout 400 ncu --metrics $M --clock-control none -c 400 --csv --log-file gpurun_out/launches_fast_$tag.csv $B > gpurun_out/ncu_l1_$tag.log 2>&1
echo "fast launch list rc=$?"
timeout 400 ncu --metrics $M --clock-control none -s 250 -c 300 --csv --log-file gpurun_out/launches_precise_$tag.csv python tools/profile_targets.py precise_pass > gpurun_out/ncu_l2_$tag.log 2>&1
echo "precise launch list rc=$?"
timeout 400 ncu --metrics $M --clock-control none -c 80 --csv --log-file gpurun_out/launches_topk_$tag.csv python tools/profile_targets.py topk_batched > gpurun_out/ncu_l3_$tag.log 2>&1
echo "topk launch list rc=$?"
timeout 600 ncu --metrics $M --clock-control none -c 60 --csv --log-file gpurun_out/launches_sparse_$tag.csv python tools/profile_targets.py sparse_1m > gpurun_out/ncu_l4_$tag.log 2>&1
echo "sparse launch list rc=$?"
timeout 400 ncu --set full --import-source on --clock-control none --kernel-name regex:gemm_tcgen05_kernel --launch-skip 96 --launch-count 4 -o gpurun_out/gemm_fast_full_$tag -f $B > gpurun_out/ncu_f1_$tag.log 2>&1
echo "gemm fast full rc=$?"
timeout 400 ncu --set full --import-source on --clock-control none --kernel-name regex:gemm_tcgen05_kernel --launch-skip 96 --launch-count 4 -o gpurun_out/gemm_split_full_$tag -f python tools/profile_targets.py precise_pass > gpurun_out/ncu_f2_$tag.log 2>&1
echo "gemm split full rc=$?"
timeout 400 ncu --set full --import-source on --clock-control none --kernel-name regex:attention_tc_kernel --launch-skip 24 --launch-count 2 -o gpurun_out/attn_fast_full_$tag -f $B > gpurun_out/ncu_f3_$tag.log 2>&1
echo "attention fast full rc=$?"
timeout 400 ncu --set full --import-source on --clock-control none --kernel-name regex:attention_tc_kernel --launch-skip 24 --launch-count 2 -o gpurun_out/attn_split_full_$tag -f python tools/profile_targets.py precise_pass > gpurun_out/ncu_f4_$tag.log 2>&1
echo "attention split full rc=$?"
timeout 400 ncu --set full --import-source on --clock-control none --kernel-name regex:gemm_tcgen05_kernel --launch-skip 2 --launch-count 2 -o gpurun_out/topk_gemm_full_$tag -f python tools/profile_targets.py topk_batched > gpurun_out/ncu_f5_$tag.log 2>&1
echo "topk gemm full rc=$?"
timeout 600 ncu --set full --import-source on --clock-control none --kernel-name regex:sparse_scan_kernel --launch-skip 2 --launch-count 1 -o gpurun_out/sparse_scan_full_$tag -f python tools/profile_targets.py sparse_1m > gpurun_out/ncu_f6_$tag.log 2>&1
echo "sparse scan full rc=$?"
ls -la gpurun_out/*_$tag* | head -30
# gpurun brings back at most 64 MiB: export the raw pages here and keep only the small reports
for f in gemm_fast_full gemm_split_full attn_fast_full attn_split_full topk_gemm_full sparse_scan_full; do
  ncu -i gpurun_out/${f}_$tag.ncu-rep --page raw --csv > gpurun_out/${f}_${tag}_raw.csv 2>/dev/null
done
rm -f gpurun_out/gemm_fast_full_$tag.ncu-rep gpurun_out/gemm_split_full_$tag.ncu-rep gpurun_out/topk_gemm_full_$tag.ncu-rep
ls -la gpurun_out/*_raw.csv
