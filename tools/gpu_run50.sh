#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --tb=short > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --profile > gpurun_out/bench_silu.json 2>/dev/null; python tools/show_bench.py gpurun_out/bench_silu.json | head -12
