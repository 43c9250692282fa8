#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_unet.py tests/test_gpu_compat.py tests/test_gpu_baselines.py -q -x > gpurun_out/r2q_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2q_pytest.log | cut -c1-300
for v in 1 0; do
  HEDIT_GEMM_SPLITK=$v timeout 600 python bench.py --batch 1 --steps 3 --warmup 3 --no-cpu-baseline --no-single-image > gpurun_out/r2q_bench_b1_splitk$v.json 2> gpurun_out/r2q_bench_b1_splitk$v.err; echo "bench b1 splitk=$v rc=$?"; python tools/show_bench.py gpurun_out/r2q_bench_b1_splitk$v.json | head -1 | cut -c1-110
done
