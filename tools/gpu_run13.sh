#!/bin/bash
mkdir -p gpurun_out
echo "--- v2"; timeout 300 python tools/op_bench.py cross
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q --tb=short -x -k cross > gpurun_out/ops.log 2>&1; echo "ops rc=$?"; tail -3 gpurun_out/ops.log
timeout 900 python -m pytest tests/test_gpu_unet.py -m gpu -q -s --tb=short -k "replace_mos2 or sd15 or explicit" > gpurun_out/loop.log 2>&1; echo "loop rc=$?"; grep -E "rel |passed|failed|Error" gpurun_out/loop.log
timeout 900 python bench.py --steps 1 --warmup 2 --profile --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_quick.err
