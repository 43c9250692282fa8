#!/bin/bash
# PnP + skipped-schedule variants against the reference goldens
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_unet.py -m gpu -q --tb=short -k "variants" -s > gpurun_out/pytest_variants.log 2>&1; echo "pytest rc=$?"; grep -E "rel|passed|failed|Error|error" gpurun_out/pytest_variants.log | tail -20
