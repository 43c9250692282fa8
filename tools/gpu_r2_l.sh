#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_unet.py -q -x > gpurun_out/r2l_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2l_pytest.log | cut -c1-300
for v in 1 0; do
  HEDIT_GEMM_TMA_STORE=$v timeout 300 python tools/op_bench.py linear > gpurun_out/r2l_op_bench_tma$v.log 2>&1; echo "op_bench tma=$v rc=$?"; grep "f32out=False" gpurun_out/r2l_op_bench_tma$v.log | cut -c1-200
done
for v in 1 0; do
  HEDIT_GEMM_TMA_STORE=$v timeout 600 python bench.py --steps 2 --warmup 3 > gpurun_out/r2l_bench_tma$v.json 2> gpurun_out/r2l_bench_tma$v.err; echo "bench tma=$v rc=$?"
  python tools/show_bench.py gpurun_out/r2l_bench_tma$v.json 2>/dev/null | grep -E "value|qkv|tf.q2|total" | cut -c1-200
done
