#!/bin/bash
# Re-capture of the round-2 evidence that changed after tools/gpu_profile_r2.sh ran: the precise-mode launch list and
# the split attention kernel (now two CTAs per SM), the sparse scan (final lane layout, one launch per selection group).
tag=${1:-r2c}
mkdir -p gpurun_out
M="gpu__time_duration.sum,sm__cycles_elapsed.max,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum"
timeout 400 ncu --metrics $M --clock-control none -s 250 -c 300 --csv --log-file gpurun_out/launches_precise_$tag.csv python tools/profile_targets.py precise_pass > gpurun_out/ncu_l2_$tag.log 2>&1
echo "precise launch list rc=$?"
timeout 400 ncu --set full --import-source on --clock-control none --kernel-name regex:attention_tc_kernel --launch-skip 24 --launch-count 2 -o gpurun_out/attn_split_full_$tag -f python tools/profile_targets.py precise_pass > gpurun_out/ncu_f4_$tag.log 2>&1
echo "attention split full rc=$?"
timeout 600 ncu --set full --import-source on --clock-control none --kernel-name regex:sparse_scan_kernel --launch-skip 1 --launch-count 1 -o gpurun_out/sparse_scan_full_$tag -f python tools/profile_targets.py sparse_1m > gpurun_out/ncu_f6_$tag.log 2>&1
echo "sparse scan full rc=$?"
for f in attn_split_full sparse_scan_full; do
  ncu -i gpurun_out/${f}_$tag.ncu-rep --page raw --csv > gpurun_out/${f}_${tag}_raw.csv 2>/dev/null
  rm -f gpurun_out/${f}_$tag.ncu-rep
done
ls -la gpurun_out/*_${tag}*
