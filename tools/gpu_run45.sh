#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_vae.py tests/test_gpu_unet.py -m gpu -q --tb=short -s -k "encode or forward_tiny or sd15" > gpurun_out/pytest_enc.log 2>&1; echo "pytest rc=$?"; grep -vi "warn" gpurun_out/pytest_enc.log | tail -20
