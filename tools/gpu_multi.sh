#!/bin/bash
# N-GPU validation: sharded top-k parity over NCCL + the torchrun bench line (driver-style launch).
# Run as: gpurun --gpus N --timeout 900 -- 'bash tools/gpu_multi.sh N [tag] [ref]'
n=${1:-2}; tag=${2:-multi}; ref=${3:-}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_golden.py -m gpu -q --timeout 600 -k "sharded" > gpurun_out/tests_${tag}.log 2>&1
echo "sharded tests rc=$?"; tail -3 gpurun_out/tests_${tag}.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 \
  bench.py --gpus $n --steps 2 --warmup 3 > gpurun_out/bench_${tag}.json 2>gpurun_out/bench_${tag}.err
echo "bench rc=$? stdout lines: $(wc -l < gpurun_out/bench_${tag}.json)"; cut -c1-400 gpurun_out/bench_${tag}.json; tail -3 gpurun_out/bench_${tag}.err
if [ -n "$ref" ]; then
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29518 \
  bench.py --impl reference --gpus $n --steps 1 --warmup 1 > gpurun_out/bench_ref_${tag}.json 2>gpurun_out/bench_ref_${tag}.err
echo "ref rc=$?"; cut -c1-200 gpurun_out/bench_ref_${tag}.json
fi
