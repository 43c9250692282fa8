#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_vae.py -m gpu -q --tb=short -x -s > gpurun_out/pytest_vae.log 2>&1; echo "pytest rc=$?"; grep -vi "warn" gpurun_out/pytest_vae.log | tail -40
