#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_style.py tests/test_gpu_vae.py -m gpu -q --tb=short -x -s > gpurun_out/pytest_style.log 2>&1; echo "pytest rc=$?"; grep -vi "warn" gpurun_out/pytest_style.log | tail -25
timeout 600 python tools/op_bench.py vae --iters 5 > gpurun_out/op_bench_vae.log 2>&1; cat gpurun_out/op_bench_vae.log | tail
