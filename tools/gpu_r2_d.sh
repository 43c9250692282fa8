#!/bin/bash
# Round-2 GPU call D: full GPU suite with the new tests, bench (sparse replace path), UNet leg of the library comparator
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --tb=short -x > gpurun_out/r2d_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r2d_pytest_gpu.log
timeout 900 python bench.py --profile > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r2d_bench.err; python tools/show_bench.py gpurun_out/r2d_bench.json 2>/dev/null | head -12
timeout 900 python tests/library_comparator.py --iters 10 --out gpurun_out/r02_vs_library.json 2>&1 | grep -v Warning | tail -4
