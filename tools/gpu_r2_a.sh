#!/bin/bash
# Round-2 GPU call A: micro-benchmarks, self-attention variants (speed + accuracy + op tests), full GPU test suite, per-shape forward profiles
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2a_smi.txt 2>&1
timeout 120 tools/micro/ubench > gpurun_out/r2a_ubench.log 2>&1; echo "ubench rc=$?"; cat gpurun_out/r2a_ubench.log
: > gpurun_out/r2a_attn_variants.log
for v in 0 1 2 3 4 5 6 7; do
  HEDIT_ATTN_V3=$v timeout 300 python tools/op_bench.py attn --iters 20 >> gpurun_out/r2a_attn_variants.log 2>&1
  HEDIT_ATTN_V3=$v timeout 300 python -m pytest tests/test_gpu_ops.py -q -k "self_attention" 2>&1 | tail -1 >> gpurun_out/r2a_attn_variants.log
done
cat gpurun_out/r2a_attn_variants.log
timeout 1500 python -m pytest tests -m gpu -q --tb=short -x > gpurun_out/r2a_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2a_pytest_gpu.log
for s in 2 5 40; do timeout 300 python tools/shape_prof.py --samples $s > gpurun_out/r2a_shape_prof_s$s.log 2>&1; head -12 gpurun_out/r2a_shape_prof_s$s.log; done
