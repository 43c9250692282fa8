#!/bin/bash
# quick iteration: operator tests + one loop test + profiled quick bench
mkdir -p gpurun_out
: > gpurun_out/summary.txt
run() { local name=$1 to=$2; shift 2; timeout "$to" "$@" > "gpurun_out/$name.log" 2>&1; echo "$name rc=$?" >> gpurun_out/summary.txt; }
run ops 600 python -m pytest tests/test_gpu_ops.py -m gpu -q --tb=short -x
run loop 600 python -m pytest tests/test_gpu_unet.py -m gpu -q -s --tb=short -k "forward_tiny or sd15_config1 or replace_mos2"
timeout 900 python bench.py --steps 1 --warmup 1 --profile --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err
echo "bench rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -3 gpurun_out/ops.log; grep -E "rel " gpurun_out/loop.log; tail -3 gpurun_out/bench_quick.err
