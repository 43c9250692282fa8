#!/bin/bash
# Round-2 GPU call B: fixed micro-benchmark, ping-pong self-attention variants
mkdir -p gpurun_out
timeout 120 tools/micro/ubench > gpurun_out/r2b_ubench.log 2>&1; echo "ubench rc=$?"; cat gpurun_out/r2b_ubench.log
: > gpurun_out/r2b_attn_variants.log
for v in 0 8 9 10 11 12 13; do
  HEDIT_ATTN_V3=$v timeout 300 python tools/op_bench.py attn --iters 20 >> gpurun_out/r2b_attn_variants.log 2>&1
  HEDIT_ATTN_V3=$v timeout 300 python -m pytest tests/test_gpu_ops.py -q -k "self_attention" 2>&1 | tail -1 >> gpurun_out/r2b_attn_variants.log
done
cat gpurun_out/r2b_attn_variants.log
