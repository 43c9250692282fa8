#!/bin/bash
mkdir -p gpurun_out
echo "--- v2"; timeout 300 python tools/op_bench.py cross
echo "--- v1"; HEDIT_CROSS_V1=1 timeout 300 python tools/op_bench.py cross
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cross_attn2 -s 2 -c 1 -o gpurun_out/prof_cross2 python tools/op_bench.py cross --iters 1 > gpurun_out/ncu_cross.log 2>&1; echo "ncu rc=$?"
