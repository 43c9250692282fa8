#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/op_bench.py face --iters 5 > gpurun_out/op_bench_face.log 2>&1; tail -3 gpurun_out/op_bench_face.log
timeout 1200 python -m pytest tests -m gpu -q --tb=short > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
