#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q --tb=short -x -k "attention" > gpurun_out/ops.log 2>&1; echo "ops rc=$?"; tail -2 gpurun_out/ops.log
echo "--- NT=2 BKV=128"; timeout 300 python tools/op_bench.py attn
echo "--- NT=4 BKV=64"; HEDIT_ATTN_NT4=1 timeout 300 python tools/op_bench.py attn
HEDIT_ATTN_NT4=1 timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q --tb=short -x -k "attention" > gpurun_out/ops4.log 2>&1; echo "ops(nt4) rc=$?"; tail -2 gpurun_out/ops4.log
