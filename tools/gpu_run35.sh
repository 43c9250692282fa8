#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_clip.py tests/test_gpu_style.py -m gpu -q --tb=short -s > gpurun_out/pytest_clip.log 2>&1; echo "pytest rc=$?"; grep -vi "warn" gpurun_out/pytest_clip.log | tail -40
