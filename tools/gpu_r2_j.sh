#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2j_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2j_pytest_gpu.log | cut -c1-300
for c in 5 3 4; do
  timeout 600 python bench.py --config $c --steps 2 --warmup 1 > gpurun_out/r2j_bench_config$c.json 2> gpurun_out/r2j_bench_config$c.err; echo "config $c rc=$?"
  python tools/show_bench.py gpurun_out/r2j_bench_config$c.json 2>/dev/null | cut -c1-600; tail -2 gpurun_out/r2j_bench_config$c.err
done
HEDIT_NATIVE_REWARD=0 timeout 600 python bench.py --config 5 --steps 1 --warmup 1 > gpurun_out/r2j_bench_config5_torchrewards.json 2> gpurun_out/r2j_bench_config5_torchrewards.err; echo "config 5 torch rewards rc=$?"
python tools/show_bench.py gpurun_out/r2j_bench_config5_torchrewards.json 2>/dev/null | cut -c1-400
