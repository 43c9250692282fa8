#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --tb=short -s > gpurun_out/r2h_pytest_gpu.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|error" gpurun_out/r2h_pytest_gpu.log | tail -5
grep -E "sd15_config3|th09|golden mask px|cuda :|16-bit-operand oracle:|pixels on the other side|tiny_refine_blend_substruct|tiny_masactrl" gpurun_out/r2h_pytest_gpu.log | cut -c1-250
timeout 900 python bench.py --profile > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r2h_bench.err; python tools/show_bench.py gpurun_out/r2h_bench.json 2>/dev/null | head -14 | cut -c1-600
