#!/bin/bash
for c in 0 1 2; do echo "--- HEDIT_ATTN_CFG=$c"; HEDIT_ATTN_CFG=$c timeout 300 python tools/op_bench.py attn > /tmp/o.txt 2>&1; head -1 /tmp/o.txt; done
HEDIT_ATTN_CFG=1 timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q --tb=short -x -k "attention" > /tmp/p.txt 2>&1; tail -2 /tmp/p.txt
