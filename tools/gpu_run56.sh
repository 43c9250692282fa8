#!/bin/bash
mkdir -p gpurun_out
echo "== single"; timeout 300 python tools/op_bench.py conv --iters 20 2>&1 | tail -5
echo "== pair"; HEDIT_GEMM_CLUSTER=1 timeout 300 python tools/op_bench.py conv --iters 20 2>&1 | tail -5
HEDIT_GEMM_CLUSTER=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 2 -c 1 -o gpurun_out/prof_conv_pair -f python tools/op_bench.py conv --iters 1 > gpurun_out/ncu_conv_pair.log 2>&1; echo "ncu rc=$?"
ncu -i gpurun_out/prof_conv_pair.ncu-rep --page raw --csv > gpurun_out/prof_conv_pair_raw.csv 2>/dev/null
