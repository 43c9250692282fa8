#!/bin/bash
# Round-2 GPU call C: headline parity (T=50 goldens, calibrated), new bench.py, library comparator, attention variants at 40 samples
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_headline.py -q -s --tb=short > gpurun_out/r2c_headline.log 2>&1; echo "headline rc=$?"; tail -60 gpurun_out/r2c_headline.log
timeout 900 python bench.py --profile > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r2c_bench.err; python tools/show_bench.py gpurun_out/r2c_bench.json 2>/dev/null | head -40
timeout 900 python tests/library_comparator.py --iters 10 > gpurun_out/r2c_vs_library.log 2>&1; echo "comparator rc=$?"; tail -30 gpurun_out/r2c_vs_library.log
for v in 0 10 11 3; do HEDIT_ATTN_V3=$v timeout 300 python tools/op_bench.py attn --iters 10 --samples 40 2>&1 | grep "N=4096\|N=1024"; done | tee gpurun_out/r2c_attn_s40.log
