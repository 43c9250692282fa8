#!/bin/bash
# compute-sanitizer over the operator-level GPU tests (memcheck and racecheck), incl. the cta_group::2 GEMM variant
mkdir -p gpurun_out
S=/usr/local/cuda/bin/compute-sanitizer
for tool in memcheck racecheck; do
  timeout 1500 $S --tool $tool --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_ops.py -q -x > gpurun_out/r02_sanitizer_${tool}_ops.log 2>&1; echo "$tool ops rc=$?"
  tail -4 gpurun_out/r02_sanitizer_${tool}_ops.log
done
HEDIT_GEMM_CLUSTER=1 timeout 1200 $S --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_ops.py -q -x -k "linear or conv or gemm" > gpurun_out/r02_sanitizer_memcheck_cluster.log 2>&1; echo "memcheck cluster rc=$?"; tail -3 gpurun_out/r02_sanitizer_memcheck_cluster.log
timeout 1500 $S --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_unet.py -q -x -k "tiny_refine_blend or dedup" > gpurun_out/r02_sanitizer_memcheck_loop.log 2>&1; echo "memcheck loop rc=$?"; tail -3 gpurun_out/r02_sanitizer_memcheck_loop.log
