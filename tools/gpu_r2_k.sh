#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_reward.py tests/test_gpu_face.py tests/test_gpu_ops.py -q > gpurun_out/r2k_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2k_pytest.log | cut -c1-300
timeout 300 python tools/reward_bench.py 8 > gpurun_out/r2k_reward_bench.json 2> gpurun_out/r2k_reward_bench.err; cat gpurun_out/r2k_reward_bench.json
for w in arcface lpips; do
HEDIT_NET_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2k_${w}_launches.csv python tools/reward_prof.py $w > /dev/null 2>&1
done
timeout 600 python bench.py --config 5 --steps 1 --warmup 1 > gpurun_out/r2k_bench_config5.json 2> gpurun_out/r2k_bench_config5.err; echo "config 5 rc=$?"
python tools/show_bench.py gpurun_out/r2k_bench_config5.json 2>/dev/null | cut -c1-200
timeout 600 python bench.py --steps 2 --warmup 3 > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err; echo "bench rc=$?"
python tools/show_bench.py gpurun_out/r2k_bench.json 2>/dev/null | cut -c1-200
