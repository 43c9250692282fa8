#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_reward.py -q > gpurun_out/r2m_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2m_pytest.log | cut -c1-300
timeout 300 python tools/reward_bench.py 8 > gpurun_out/r2m_reward_bench.json 2> gpurun_out/r2m_reward_bench.err; cat gpurun_out/r2m_reward_bench.json
HEDIT_NET_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2m_arcface_launches.csv python tools/reward_prof.py arcface > /dev/null 2>&1
