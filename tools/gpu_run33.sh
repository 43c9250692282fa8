#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
timeout 600 python tools/op_bench.py vae --iters 5 > gpurun_out/op_bench_vae.log 2>&1; cat gpurun_out/op_bench_vae.log | tail -3
