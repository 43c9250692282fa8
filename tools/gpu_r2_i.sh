#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_reward.py -q -s > gpurun_out/r2i_pytest_reward.log 2>&1; echo "pytest rc=$?"
grep -E '^arcface|^lpips|^h_Edit|passed|failed|Error|assert'  gpurun_out/r2i_pytest_reward.log | cut -c1-500
