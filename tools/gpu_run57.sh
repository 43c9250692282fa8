#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_text.py tests/test_gpu_clip.py -m gpu -q --tb=short -s 2>&1 | grep -v -i warn | tail -14
