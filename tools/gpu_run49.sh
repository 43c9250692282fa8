#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_unet.py -m gpu -q --tb=short -k "batched_edit" 2>&1 | tail -8
