#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/r2g_attn.log
for v in 24 31 32 33 34 35; do
  HEDIT_ATTN_V3=$v timeout 120 python tools/op_bench.py attn --iters 10 --samples 40 2>&1 | grep "N=4096\|N=1024\|rror" >> gpurun_out/r2g_attn.log
  HEDIT_ATTN_V3=$v timeout 200 python -m pytest tests/test_gpu_ops.py -q -k "self_attention" 2>&1 | tail -1 >> gpurun_out/r2g_attn.log
done
cat gpurun_out/r2g_attn.log | cut -c1-175
timeout 900 python -m pytest tests/test_gpu_unet.py -q --tb=short -x -k "dedup or tiny or variants or batched" > gpurun_out/r2g_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r2g_pytest.log
HEDIT_PREFIX_DEDUP=0 timeout 600 python bench.py --no-cpu-baseline --no-single-image --steps 2 > gpurun_out/r2g_bench_nodedup.json 2>gpurun_out/r2g_bench_nodedup.err; python tools/show_bench.py gpurun_out/r2g_bench_nodedup.json | head -1 | cut -c1-200
timeout 600 python bench.py --no-cpu-baseline --no-single-image --steps 2 > gpurun_out/r2g_bench_dedup.json 2>gpurun_out/r2g_bench_dedup.err; python tools/show_bench.py gpurun_out/r2g_bench_dedup.json | head -1 | cut -c1-200
HEDIT_ATTN_V3=24 timeout 600 python bench.py --no-cpu-baseline --no-single-image --steps 2 > gpurun_out/r2g_bench_dedup_v24.json 2>gpurun_out/r2g_bench_v24.err; python tools/show_bench.py gpurun_out/r2g_bench_dedup_v24.json | head -1 | cut -c1-200
tail -3 gpurun_out/r2g_bench_dedup.err
