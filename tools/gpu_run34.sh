#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_unet.py -m gpu -q --tb=short -x -k "skip_uncond or refine_blend" -s > gpurun_out/pytest_s2.log 2>&1; echo "pytest rc=$?"; grep -E "sched=|passed|failed|Error" gpurun_out/pytest_s2.log | tail
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench_s2.json 2> gpurun_out/bench_s2.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_s2.err; cat gpurun_out/bench_s2.json
