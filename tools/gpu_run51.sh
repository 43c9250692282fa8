#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_face.py -m gpu -q --tb=short -s -k "inversion" 2>&1 | grep -v -i warn | tail -8
