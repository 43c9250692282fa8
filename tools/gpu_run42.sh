#!/bin/bash
for P in 0 2 4; do echo "HEDIT_ATTN_POLY=$P"; HEDIT_ATTN_POLY=$P timeout 300 python tools/op_bench.py attn --iters 20 2>&1 | head -2; done
HEDIT_ATTN_POLY=4 timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q -k "attention" 2>&1 | tail -3
