#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_face.py tests/test_gpu_vae.py -m gpu -q --tb=short -s > gpurun_out/pytest_face.log 2>&1; echo "pytest rc=$?"; grep -vi "warn" gpurun_out/pytest_face.log | tail -40
