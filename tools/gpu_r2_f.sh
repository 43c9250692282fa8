#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/r2f_attn.log
for v in 0 20 24 26 27 28 29 30; do
  HEDIT_ATTN_V3=$v timeout 120 python tools/op_bench.py attn --iters 10 --samples 40 2>&1 | grep "N=4096\|N=1024\|rror" >> gpurun_out/r2f_attn.log
  HEDIT_ATTN_V3=$v timeout 200 python -m pytest tests/test_gpu_ops.py -q -k "self_attention" 2>&1 | tail -1 >> gpurun_out/r2f_attn.log
done
cat gpurun_out/r2f_attn.log | cut -c1-175
