#!/bin/bash
# Round-2 starting point: validate and time the experimental two-tiles-per-CTA attention kernel (attention_tc2.cu).
# A protocol bug shows up as "[vrag] mbarrier wait timed out: tag N" (2 s watchdog in every wait) and a CUDA error,
# not as a hung box; every step still runs under its own timeout.
# Run as: gpurun --timeout 600 -- 'bash tools/gpu_attn_v2.sh'
mkdir -p gpurun_out
echo "== v1 timing"; timeout 120 python tools/attn_probe.py 2>&1 | tail -8
echo "== v2 parity (attention self test vs float64, then the 4-layer model vs the oracle)"
VRAG_ATTENTION_V2=1 timeout 200 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x --timeout 120 \
  -k "attention_kernel_vs_float64 or modernbert_forward_vs_oracle or multi_pass" > gpurun_out/tests_attn_v2.log 2>&1
echo "rc=$?"; tail -15 gpurun_out/tests_attn_v2.log
echo "== v2 timing"; VRAG_ATTENTION_V2=1 timeout 120 python tools/attn_probe.py 2>&1 | tail -8
