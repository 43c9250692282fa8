#!/bin/bash
# Round-2 GPU call E: v4 attention variants (P in TMEM), then the full suite / bench / comparator after the cross-attention smem fix
mkdir -p gpurun_out
: > gpurun_out/r2e_attn.log
for v in 0 20 21 22 23 24 25; do
  HEDIT_ATTN_V3=$v timeout 120 python tools/op_bench.py attn --iters 10 --samples 40 2>&1 | grep "N=4096\|N=1024\|rror" >> gpurun_out/r2e_attn.log
  HEDIT_ATTN_V3=$v timeout 200 python -m pytest tests/test_gpu_ops.py -q -k "self_attention" 2>&1 | tail -1 >> gpurun_out/r2e_attn.log
done
cat gpurun_out/r2e_attn.log
timeout 2400 python -m pytest tests -m gpu -q --tb=short > gpurun_out/r2e_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r2e_pytest_gpu.log
timeout 900 python bench.py --profile > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r2e_bench.err; python tools/show_bench.py gpurun_out/r2e_bench.json 2>/dev/null | head -12
timeout 900 python tests/library_comparator.py --iters 10 --out gpurun_out/r02_vs_library.json 2>&1 | grep "unet_forward"
