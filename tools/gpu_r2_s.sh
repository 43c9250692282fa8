#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/single_image_probe.py 2>&1 | tail -3
S=/usr/local/cuda/bin/compute-sanitizer
timeout 900 $S --tool racecheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_ops.py -q -x -k "(self_attention and 2-256-8-40) or (self_attention and 2-64-8-160) or (cross_attention and 256-8-8) or (cross_attention and 64-8-160)" > gpurun_out/r02_sanitizer_racecheck_attention.log 2>&1; echo "racecheck attention rc=$?"; tail -4 gpurun_out/r02_sanitizer_racecheck_attention.log
timeout 600 $S --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_ops.py tests/test_gpu_reward.py -q -x -k "(linear and 128-1280-5120) or (conv3x3 and 2-8-8-1280) or (arcface and 1-False) or (lpips and 128-2-1)" > gpurun_out/r02_sanitizer_memcheck_splitk_reward.log 2>&1; echo "memcheck splitk+reward rc=$?"; tail -4 gpurun_out/r02_sanitizer_memcheck_splitk_reward.log
